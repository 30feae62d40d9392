// awm_cli.cc -- the `audiowmark` command line of the B200 build.
//
// Drop-in for the reference CLI (src/audiowmark.cc): same argv grammar, same messages, same exit codes for add, get, cmp, gen-key
// and the helper commands the reference's tests/*.sh call.  The implementation is table driven:
//   * kOptions  one row per option: spelling, arity, the commands that accept it and a small handler that stores the value
//   * kCommands one row per command: the option groups it accepts, its key policy, the names of its positional arguments and the
//               function that runs it
// A command line is processed by ONE generic routine: options are taken out of the token list row by row (table order, so that
// e.g. --format still overrides --input-format wherever it stands), every handler validates its own value, what is left must be
// exactly the positional arguments.  Not available here: hls-*, MP3 / FLAC and other libsndfile formats.
#include <errno.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <initializer_list>
#include <string>
#include <vector>

#include "awm_wm.hh"
#include "awm_speed.hh"
#include "awm_engine.hh"
#include "awm_tables.hh"
#include "awm_util.hh"
#include "awm_cli_tools.hh"

using std::string;
using std::vector;

#define AWM_VERSION "0.6.5-b200"

namespace {

[[noreturn]] void
die (const char *fmt, ...)
{
  va_list ap;
  va_start (ap, fmt);
  char buf[1024];
  vsnprintf (buf, sizeof (buf), fmt, ap);
  va_end (ap);
  error ("%s", buf);
  exit (1);
}

/* Numbers: what the reference accepts (src/audiowmark.cc:175-199) -- strtol with base 0 (so 0x10 and 010 work) / strtof, the
 * whole token must be consumed, an empty token counts as 0, overflow is not diagnosed. */
template<class T, class Convert> T
number_or_die (const string& s, const char *type_name, Convert convert)
{
  char *end = nullptr;
  const T v = T (convert (s.c_str(), &end));
  if (end && *end)
    die ("audiowmark: error during string->%s conversion: %s\n", type_name, s.c_str());
  return v;
}

int   to_int (const string& s)   { return number_or_die<int> (s, "int", [] (const char *p, char **e) { return strtol (p, e, 0); }); }
float to_float (const string& s) { return number_or_die<float> (s, "float", [] (const char *p, char **e) { return strtof (p, e); }); }

bool looks_like_option (const string& t) { return t.size() > 1 && t[0] == '-'; }     // a lone "-" is stdin / stdout

/* ------------------------------------------------------------------------------------------------ value vocabularies */

template<class T> struct Word { const char *text; T value; };

template<class T, size_t N> T
lookup (const Word<T> (&words)[N], const string& s, const char *what)
{
  for (const auto& w : words)
    if (s == w.text)
      return w.value;
  die ("audiowmark: unsupported %s '%s'\n", what, s.c_str());
}

const Word<Format> kFormats[] = { { "raw", Format::RAW }, { "auto", Format::AUTO }, { "rf64", Format::RF64 }, { "wav-pipe", Format::WAV_PIPE } };
const Word<RawFormat::Endian> kEndians[] = { { "little", RawFormat::LITTLE }, { "big", RawFormat::BIG } };
struct EncodingChoice { Encoding enc; int forced_bits; };      // float / double carry their width
const Word<EncodingChoice> kEncodings[] = { { "signed", { Encoding::SIGNED, 0 } }, { "unsigned", { Encoding::UNSIGNED, 0 } },
                                            { "float", { Encoding::FLOAT, 32 } }, { "double", { Encoding::FLOAT, 64 } } };

enum Side { IN = 1, OUT = 2, BOTH = 3 };

template<class F> void
each_raw (int side, F f)
{
  if (side & IN)  f (Params::raw_input_format);
  if (side & OUT) f (Params::raw_output_format);
}

void
set_encoding (RawFormat& fmt, const string& s)
{
  const EncodingChoice c = lookup (kEncodings, s, "encoding");
  fmt.set_encoding (c.enc);
  if (c.forced_bits)
    fmt.set_bit_depth (c.forced_bits);
}

void
set_bits (RawFormat& fmt, int bits)
{
  if (fmt.encoding() == Encoding::FLOAT)
    die ("audiowmark: bit depth can not be changed for float / double encoding\n");
  fmt.set_bit_depth (bits);                   // widths the raw converter has no code for are reported when the stream is opened
}

/* ------------------------------------------------------------------------------------------------ option table */

enum Group : unsigned
{
  G_SHARED = 1,      // payload / code options every watermark command takes
  G_ADD    = 2,
  G_GET    = 4,      // get and cmp
  G_CMP    = 8,
  G_KEYS   = 16,     // --key / --test-key
  G_NAME   = 32,     // gen-key
  G_BITS   = 64,     // test-gen-noise
};

enum Arity { FLAG, VALUE, REPEATED, CHECK };

struct Session                     // what option handlers collect besides Params
{
  vector<Key> keys;
  int         speed_option_count = 0;
  string      key_name;
  int         noise_bits = 16;
};

struct Option
{
  const char *name;                // spelling; for CHECK rows a comment
  Arity       arity;
  unsigned    groups;
  std::function<void (const vector<string>&, Session&)> apply;     // FLAG: empty vector; VALUE: one element (the last given); REPEATED: all
};

void
load_key_file (const vector<string>& files, Session& s)
{
  for (const auto& f : files)
    {
      Key key;
      key.load_key (f);
      s.keys.push_back (key);
    }
}

void
add_test_keys (const vector<string>& seeds, Session& s)
{
  for (const auto& t : seeds)
    {
      Key key;
      key.set_test_key (to_int (t));
      s.keys.push_back (key);
    }
}

#define V(expr) [] (const vector<string>& v, Session& s) { (void) v; (void) s; expr; }

const Option kOptions[] = {
  /* ---- payload / code */
  { "--short", VALUE, G_SHARED, V (
      Params::payload_size = to_int (v[0]);
      if (!short_code_init (Params::payload_size))
        die ("audiowmark: unsupported short payload size %zd\n", Params::payload_size);
      Params::payload_short = true) },
  { "--frames-per-bit", VALUE, G_SHARED, V (Params::frames_per_bit = to_int (v[0])) },
  { "--linear", FLAG, G_SHARED, V (Params::mix = false) },

  /* ---- add */
  { "--set-input-label", VALUE, G_ADD, V (Params::input_label = v[0]) },
  { "--set-output-label", VALUE, G_ADD, V (Params::output_label = v[0]) },
  { "--snr", FLAG, G_ADD, V (Params::snr = true) },
  { "--input-format", VALUE, G_ADD, V (Params::input_format = lookup (kFormats, v[0], "format")) },
  { "--output-format", VALUE, G_ADD, V (Params::output_format = lookup (kFormats, v[0], "format")) },
  { "--format", VALUE, G_ADD, V (Params::input_format = Params::output_format = lookup (kFormats, v[0], "format")) },
  { "--raw-input-endian", VALUE, G_ADD, V (Params::raw_input_format.set_endian (lookup (kEndians, v[0], "endianness"))) },
  { "--raw-output-endian", VALUE, G_ADD, V (Params::raw_output_format.set_endian (lookup (kEndians, v[0], "endianness"))) },
  { "--raw-endian", VALUE, G_ADD, V (const auto e = lookup (kEndians, v[0], "endianness"); each_raw (BOTH, [&] (RawFormat& f) { f.set_endian (e); })) },
  { "--raw-input-encoding", VALUE, G_ADD, V (set_encoding (Params::raw_input_format, v[0])) },
  { "--raw-output-encoding", VALUE, G_ADD, V (set_encoding (Params::raw_output_format, v[0])) },
  { "--raw-encoding", VALUE, G_ADD, V (each_raw (BOTH, [&] (RawFormat& f) { set_encoding (f, v[0]); })) },
  { "--raw-input-bits", VALUE, G_ADD, V (set_bits (Params::raw_input_format, to_int (v[0]))) },
  { "--raw-output-bits", VALUE, G_ADD, V (set_bits (Params::raw_output_format, to_int (v[0]))) },
  { "--raw-bits", VALUE, G_ADD, V (const int b = to_int (v[0]); each_raw (BOTH, [&] (RawFormat& f) { set_bits (f, b); })) },
  { "--raw-channels", VALUE, G_ADD, V (const int c = to_int (v[0]); each_raw (BOTH, [&] (RawFormat& f) { f.set_channels (c); })) },
  { "--raw-rate", VALUE, G_ADD, V (const int r = to_int (v[0]); each_raw (BOTH, [&] (RawFormat& f) { f.set_sample_rate (r); })) },
  { "--test-no-limiter", FLAG, G_ADD, V (Params::test_no_limiter = true) },
  { "rf64 is an output format", CHECK, G_ADD, V (
      if (Params::input_format == Format::RF64)
        die ("audiowmark: using rf64 as input format has no effect\n")) },
  { "--strength", VALUE, G_ADD, V (Params::water_delta = to_float (v[0]) / 1000) },      // add only: get normalises with the default strength

  /* ---- get / cmp */
  { "--test-cut", VALUE, G_GET, V (Params::test_cut = to_int (v[0])) },
  { "--test-truncate", VALUE, G_GET, V (Params::test_truncate = to_int (v[0])) },
  { "--hard", FLAG, G_GET, V (Params::hard = true) },
  { "--test-no-sync", FLAG, G_GET, V (Params::test_no_sync = true) },
  { "--detect-speed", FLAG, G_GET, V (Params::detect_speed = true; s.speed_option_count++) },
  { "--detect-speed-patient", FLAG, G_GET, V (Params::detect_speed_patient = true; s.speed_option_count++) },
  { "--try-speed", VALUE, G_GET, V (Params::try_speed = to_float (v[0]); s.speed_option_count++) },
  { "one speed option", CHECK, G_GET, V (
      if (s.speed_option_count > 1)
        die ("audiowmark: can only use one option: --detect-speed or --detect-speed-patient or --try-speed\n")) },
  { "--test-speed", VALUE, G_GET, V (Params::test_speed = to_float (v[0])) },
  { "--json", VALUE, G_GET, V (Params::json_output = v[0]) },
  { "--chunk-size", VALUE, G_GET, V (
      const float minutes = to_float (v[0]);
      if (minutes < 10)
        die ("audiowmark: --chunk-size needs to be at least 10 minutes\n");
      Params::get_chunk_size = minutes) },
  { "--sync-threshold", VALUE, G_GET, V (Params::sync_threshold2 = to_float (v[0])) },
  { "--n-best", VALUE, G_GET, V (
      const int n = to_int (v[0]);
      if (n < 0)
        die ("audiowmark: --n-best should not be a negative number\n");
      Params::get_n_best = n) },
  { "--expect-matches", VALUE, G_CMP, V (Params::expect_matches = to_int (v[0])) },

  /* ---- keys, gen-key, test-gen-noise */
  { "--key", REPEATED, G_KEYS, load_key_file },
  { "--test-key", REPEATED, G_KEYS, add_test_keys },
  { "--name", VALUE, G_NAME, V (s.key_name = v[0]) },
  { "--bits", VALUE, G_BITS, V (s.noise_bits = to_int (v[0])) },
};
#undef V

/* Takes the occurrences of one option out of `tokens`.  FLAG: the first occurrence only (a repeated flag is left over and reported
 * as unsupported, like in the reference); VALUE / REPEATED: every "--opt value" and "--opt=value". */
bool
extract (vector<string>& tokens, const Option& opt, vector<string>& values)
{
  const string name = opt.name, with_eq = name + "=";
  vector<string> kept;
  bool found = false;
  for (size_t i = 0; i < tokens.size(); i++)
    {
      const string& t = tokens[i];
      if (opt.arity == FLAG)
        {
          if (t == name && !found)
            found = true;
          else
            kept.push_back (t);
        }
      else if (t == name && i + 1 < tokens.size())
        {
          values.push_back (tokens[++i]);
          found = true;
        }
      else if (t.compare (0, with_eq.size(), with_eq) == 0)
        {
          values.push_back (t.substr (with_eq.size()));
          found = true;
        }
      else
        kept.push_back (t);
    }
  tokens.swap (kept);
  return found;
}

void
apply_options (vector<string>& tokens, unsigned groups, Session& session)
{
  for (const Option& opt : kOptions)
    {
      if (!(opt.groups & groups))
        continue;
      if (opt.arity == CHECK)
        {
          opt.apply ({}, session);
          continue;
        }
      vector<string> values;
      if (!extract (tokens, opt, values))
        continue;
      if (opt.arity == VALUE)
        values.erase (values.begin(), values.end() - 1);       // the last one wins
      opt.apply (values, session);
    }
}

/* ------------------------------------------------------------------------------------------------ command table */

enum KeyPolicy { NO_KEY, ONE_KEY, KEY_LIST };

struct Invocation
{
  Session        session;
  vector<string> args;             // positional arguments
  const Key&     key() const { return session.keys[0]; }
};

struct Command
{
  const char *name;
  unsigned    groups;
  KeyPolicy   keys;
  vector<const char *> positional;
  int (*run) (Invocation&);
};

const Command kCommands[] = {
  { "add", G_SHARED | G_ADD | G_KEYS, ONE_KEY, { "input_wav", "watermarked_wav", "message_hex" },
    [] (Invocation& inv) { return add_watermark (inv.key(), inv.args[0], inv.args[1], inv.args[2]); } },
  { "get", G_SHARED | G_GET | G_KEYS, KEY_LIST, { "watermarked_wav" },
    [] (Invocation& inv) { return get_watermark (inv.session.keys, inv.args[0], ""); } },
  { "cmp", G_SHARED | G_GET | G_CMP | G_KEYS, KEY_LIST, { "watermarked_wav", "message_hex" },
    [] (Invocation& inv) { return get_watermark (inv.session.keys, inv.args[0], inv.args[1]); } },
  { "gen-key", G_NAME, NO_KEY, { "key_file" },
    [] (Invocation& inv) { return cli_tools::gen_key (inv.args[0], inv.session.key_name); } },
  { "gentest", 0, NO_KEY, { "input_wav", "output_wav" },
    [] (Invocation& inv) { return cli_tools::gentest (inv.args[0], inv.args[1]); } },
  { "cut-start", 0, NO_KEY, { "input_wav", "output_wav", "cut_samples" },
    [] (Invocation& inv) { return cli_tools::cut_start (inv.args[0], inv.args[1], size_t (to_int (inv.args[2]))); } },
  { "test-subtract", 0, NO_KEY, { "input1_wav", "input2_wav", "output_wav" },
    [] (Invocation& inv) { return cli_tools::subtract (inv.args[0], inv.args[1], inv.args[2]); } },
  { "test-snr", 0, NO_KEY, { "orig_wav", "watermarked_wav" },
    [] (Invocation& inv) { return cli_tools::snr (inv.args[0], inv.args[1]); } },
  { "test-clip", G_SHARED | G_KEYS, ONE_KEY, { "input_wav", "output_wav", "seed", "seconds" },
    [] (Invocation& inv) { return cli_tools::clip (inv.key(), inv.args[0], inv.args[1], to_int (inv.args[2]), to_int (inv.args[3])); } },
  { "test-gen-noise", G_SHARED | G_BITS | G_KEYS, ONE_KEY, { "output_wav", "seconds", "sample_rate" },
    [] (Invocation& inv) { return cli_tools::gen_noise (inv.key(), inv.args[0], to_float (inv.args[1]), to_int (inv.args[2]), inv.session.noise_bits); } },
  { "test-info", G_SHARED, NO_KEY, { "input_wav", "property" },
    [] (Invocation& inv) { return cli_tools::info (inv.args[0], inv.args[1]); } },
  { "test-speed", G_SHARED | G_KEYS, ONE_KEY, { "seed" },
    [] (Invocation& inv) { return cli_tools::speed (inv.key(), to_int (inv.args[0])); } },
  { "test-change-speed", G_SHARED, NO_KEY, { "input_wav", "output_wav", "speed" },
    [] (Invocation& inv) { return cli_tools::change_speed (inv.args[0], inv.args[1], to_float (inv.args[2])); } },
  { "test-resample", G_SHARED, NO_KEY, { "input_wav", "output_wav", "new_rate" },
    [] (Invocation& inv) { return cli_tools::resample (inv.args[0], inv.args[1], to_int (inv.args[2])); } },
};

int
run_command (const Command& cmd, vector<string> tokens)
{
  Invocation inv;
  apply_options (tokens, cmd.groups, inv.session);
  if (cmd.keys != NO_KEY && inv.session.keys.empty())
    inv.session.keys.push_back (Key());                      // the zero key
  if (cmd.keys == ONE_KEY && inv.session.keys.size() > 1)
    die ("audiowmark %s: watermark key can at most be set once (--key / --test-key option)\n", cmd.name);

  bool clean = tokens.size() == cmd.positional.size();
  for (const auto& t : tokens)
    clean = clean && !looks_like_option (t);
  if (!clean)
    {
      for (const auto& t : tokens)
        if (looks_like_option (t))
          die ("audiowmark: unsupported option '%s' for command '%s' (use audiowmark -h)\n", t.c_str(), cmd.name);
      string usage = string ("usage: audiowmark ") + cmd.name + " [options...]";
      for (const char *p : cmd.positional)
        usage += string (" <") + p + ">";
      die ("audiowmark: error parsing arguments for command '%s' (use audiowmark -h)\n\n%s\n", cmd.name, usage.c_str());
    }
  inv.args = tokens;
  return cmd.run (inv);
}

void
print_usage()
{
  static const char *const text[] = {
    "usage: audiowmark <command> [ <args>... ]",
    "",
    "Commands:",
    "  * create a watermarked wav file with a message",
    "    audiowmark add <input_wav> <watermarked_wav> <message_hex>",
    "",
    "  * retrieve message",
    "    audiowmark get <watermarked_wav>",
    "",
    "  * compare watermark message with expected message",
    "    audiowmark cmp <watermarked_wav> <message_hex>",
    "",
    "  * generate 128-bit watermarking key, to be used with --key option",
    "    audiowmark gen-key <key_file> [ --name <key_name> ]",
    "",
    "Global options:",
    "  -q, --quiet             disable information messages",
    "  --strict                treat (minor) problems as errors",
    "  --gpu-device <n>        CUDA device to run on                [0]",
    "",
    "Options for get / cmp:",
    "  --detect-speed          detect and correct replay speed difference",
    "  --detect-speed-patient  slower, more accurate speed detection",
    "  --json <file>           write JSON results into file",
    "",
    "Options for add / get / cmp:",
    "  --key <file>            load watermarking key from file",
    "  --short <bits>          enable short payload mode",
  };
  for (const char *line : text)
    printf ("%s\n", line);
  printf ("  --strength <s>          set watermark strength              [%.6g]\n", Params::water_delta * 1000);
  static const char *const tail[] = {
    "",
    "  --input-format raw      use raw stream as input",
    "  --output-format raw     use raw stream as output",
    "  --format raw            use raw stream as input and output",
    "",
    "The options to set the raw stream parameters (such as --raw-rate",
    "or --raw-channels) follow the reference audiowmark README.",
  };
  for (const char *line : tail)
    printf ("%s\n", line);
}

/* a global flag may stand anywhere on the command line */
bool
take_flag (vector<string>& tokens, std::initializer_list<const char *> spellings)
{
  for (const char *sp : spellings)
    for (auto it = tokens.begin(); it != tokens.end(); ++it)
      if (*it == sp)
        {
          tokens.erase (it);
          return true;
        }
  return false;
}

} // namespace

int
main (int argc, char **argv)
{
  vector<string> tokens (argv + 1, argv + argc);

  if (take_flag (tokens, { "--help", "-h" }))
    {
      print_usage();
      return 0;
    }
  if (take_flag (tokens, { "--version", "-v" }))
    {
      printf ("audiowmark %s\n", AWM_VERSION);
      return 0;
    }
  if (take_flag (tokens, { "--quiet", "-q" }))
    set_log_level (Log::WARNING);
  if (take_flag (tokens, { "--strict" }))
    Params::strict = true;
  {
    const Option gpu_device { "--gpu-device", VALUE, 0, nullptr };
    vector<string> values;
    if (extract (tokens, gpu_device, values))
      Params::gpu_device = to_int (values.back());
  }

  int rc = 1;
  if (tokens.empty())
    error ("audiowmark: error parsing commandline args (use audiowmark -h)\n");
  else
    {
      const string word = tokens.front();
      const Command *cmd = nullptr;
      for (const Command& c : kCommands)
        if (word == c.name)
          cmd = &c;
      if (cmd)
        rc = run_command (*cmd, vector<string> (tokens.begin() + 1, tokens.end()));
      else if (word == "hls-add" || word == "hls-prepare")
        error ("audiowmark: command '%s' is not available in this build (HLS is out of scope)\n", word.c_str());
      else if (looks_like_option (word))
        error ("audiowmark: unsupported global option '%s' (use audiowmark -h)\n", word.c_str());
      else
        error ("audiowmark: unsupported command '%s' (use audiowmark -h)\n", word.c_str());
    }
  Engine::shutdown();
  return rc;
}
