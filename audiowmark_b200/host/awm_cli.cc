// awm_cli.cc -- the `audiowmark` command line of the B200 build.
// Same argv grammar, messages and exit codes as the reference CLI (src/audiowmark.cc:47-88,540-1079)
// for: add, get, cmp, gen-key and the test helpers the reference's tests/*.sh use (test-gen-noise,
// cut-start, test-snr, test-info, test-clip, test-subtract, gentest).  Not available here: hls-*,
// MP3/FLAC and other libsndfile formats.
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <errno.h>

#include <string>
#include <vector>

#include "awm_wm.hh"
#include "awm_speed.hh"
#include "awm_engine.hh"
#include "awm_tables.hh"
#include "awm_util.hh"

using std::string;
using std::vector;

#define AWM_VERSION "0.6.5-b200"

static void
print_usage()
{
  printf ("usage: audiowmark <command> [ <args>... ]\n");
  printf ("\n");
  printf ("Commands:\n");
  printf ("  * create a watermarked wav file with a message\n");
  printf ("    audiowmark add <input_wav> <watermarked_wav> <message_hex>\n");
  printf ("\n");
  printf ("  * retrieve message\n");
  printf ("    audiowmark get <watermarked_wav>\n");
  printf ("\n");
  printf ("  * compare watermark message with expected message\n");
  printf ("    audiowmark cmp <watermarked_wav> <message_hex>\n");
  printf ("\n");
  printf ("  * generate 128-bit watermarking key, to be used with --key option\n");
  printf ("    audiowmark gen-key <key_file> [ --name <key_name> ]\n");
  printf ("\n");
  printf ("Global options:\n");
  printf ("  -q, --quiet             disable information messages\n");
  printf ("  --strict                treat (minor) problems as errors\n");
  printf ("  --gpu-device <n>        CUDA device to run on                [0]\n");
  printf ("\n");
  printf ("Options for get / cmp:\n");
  printf ("  --detect-speed          detect and correct replay speed difference\n");
  printf ("  --detect-speed-patient  slower, more accurate speed detection\n");
  printf ("  --json <file>           write JSON results into file\n");
  printf ("\n");
  printf ("Options for add / get / cmp:\n");
  printf ("  --key <file>            load watermarking key from file\n");
  printf ("  --short <bits>          enable short payload mode\n");
  printf ("  --strength <s>          set watermark strength              [%.6g]\n", Params::water_delta * 1000);
  printf ("\n");
  printf ("  --input-format raw      use raw stream as input\n");
  printf ("  --output-format raw     use raw stream as output\n");
  printf ("  --format raw            use raw stream as input and output\n");
  printf ("\n");
  printf ("The options to set the raw stream parameters (such as --raw-rate\n");
  printf ("or --raw-channels) follow the reference audiowmark README.\n");
}

static int
atoi_or_die (const char *s)
{
  char *end;
  errno = 0;
  const long l = strtol (s, &end, 10);
  if (errno || *end || !*s)
    {
      error ("audiowmark: error during string->int conversion: %s\n", s);
      exit (1);
    }
  return l;
}

static float
atof_or_die (const char *s)
{
  char *end;
  errno = 0;
  const float d = strtof (s, &end);          /* single precision like the reference (src/audiowmark.cc:188-199) */
  if (errno || *end || !*s)
    {
      error ("audiowmark: error during string->float conversion: %s\n", s);
      exit (1);
    }
  return d;
}

static bool
is_option (const string& arg)
{
  return arg.size() > 1 && arg[0] == '-';     // a single "-" means stdin / stdout
}

class ArgParser
{
  vector<string> m_args;
  string         m_command;
public:
  ArgParser (int argc, char **argv) : m_args (argv + 1, argv + argc) {}
  bool
  parse_cmd (const string& cmd)
  {
    if (m_args.empty() || m_args[0] != cmd)
      return false;
    m_args.erase (m_args.begin());
    m_command = cmd;
    return true;
  }
  vector<string>
  parse_multi_opt (const string& option)       // --option value  or  --option=value, any number of times
  {
    vector<string> values, rest;
    for (size_t i = 0; i < m_args.size(); i++)
      {
        if (m_args[i] == option && i + 1 < m_args.size())
          values.push_back (m_args[++i]);
        else if (m_args[i].compare (0, option.size() + 1, option + "=") == 0)
          values.push_back (m_args[i].substr (option.size() + 1));
        else
          rest.push_back (m_args[i]);
      }
    m_args = rest;
    return values;
  }
  bool
  parse_opt (const string& option, string& out)
  {
    const vector<string> v = parse_multi_opt (option);
    if (v.empty())
      return false;
    out = v.back();
    return true;
  }
  bool parse_opt (const string& option, int& out)   { string s; if (!parse_opt (option, s)) return false; out = atoi_or_die (s.c_str()); return true; }
  bool parse_opt (const string& option, float& out) { string s; if (!parse_opt (option, s)) return false; out = atof_or_die (s.c_str()); return true; }
  bool
  parse_opt (const string& option)
  {
    for (auto it = m_args.begin(); it != m_args.end(); it++)
      if (*it == option)
        {
          m_args.erase (it);
          return true;
        }
    return false;
  }
  bool
  parse_args (size_t expected, vector<string>& out)
  {
    if (m_args.size() != expected)
      return false;
    for (const auto& a : m_args)
      if (is_option (a))
        return false;
    out = m_args;
    return true;
  }
  const vector<string>& remaining_args() const { return m_args; }
  const string&         command() const        { return m_command; }
};

static Format
parse_format (const string& s)
{
  if (s == "raw") return Format::RAW;
  if (s == "auto") return Format::AUTO;
  if (s == "rf64") return Format::RF64;
  if (s == "wav-pipe") return Format::WAV_PIPE;
  error ("audiowmark: unsupported format '%s'\n", s.c_str());
  exit (1);
}

static RawFormat::Endian
parse_endian (const string& s)
{
  if (s == "little") return RawFormat::LITTLE;
  if (s == "big") return RawFormat::BIG;
  error ("audiowmark: unsupported endianness '%s'\n", s.c_str());
  exit (1);
}

static void
parse_encoding (const string& s, RawFormat& fmt)
{
  if (s == "signed") fmt.set_encoding (Encoding::SIGNED);
  else if (s == "unsigned") fmt.set_encoding (Encoding::UNSIGNED);
  else if (s == "float") fmt.set_encoding (Encoding::FLOAT);
  else if (s == "double")
    {
      fmt.set_encoding (Encoding::FLOAT);
      fmt.set_bit_depth (64);
    }
  else
    {
      error ("audiowmark: unsupported encoding '%s'\n", s.c_str());
      exit (1);
    }
  if (s == "float")
    fmt.set_bit_depth (32);
}

static void
update_raw_bits (RawFormat& fmt, int bits)
{
  if (fmt.encoding() == Encoding::FLOAT)
    return;                                   // float / double fix the width themselves
  if (bits != 8 && bits != 16 && bits != 24 && bits != 32)
    {
      error ("audiowmark: unsupported bit depth %d (use 8, 16, 24 or 32)\n", bits);
      exit (1);
    }
  fmt.set_bit_depth (bits);
}

static void
parse_shared_options (ArgParser& ap)
{
  int i;
  if (ap.parse_opt ("--short", i))
    {
      Params::payload_size = i;
      if (!short_code_init (Params::payload_size))
        {
          error ("audiowmark: unsupported short payload size %zd\n", Params::payload_size);
          exit (1);
        }
      Params::payload_short = true;
    }
  ap.parse_opt ("--frames-per-bit", Params::frames_per_bit);
  if (ap.parse_opt ("--linear"))
    Params::mix = false;
}

static vector<Key>
parse_key_list (ArgParser& ap)
{
  vector<Key> key_list;
  for (const auto& f : ap.parse_multi_opt ("--key"))
    {
      Key key;
      key.load_key (f);
      key_list.push_back (key);
    }
  for (const auto& t : ap.parse_multi_opt ("--test-key"))
    {
      Key key;
      key.set_test_key (atoi_or_die (t.c_str()));
      key_list.push_back (key);
    }
  if (key_list.empty())
    key_list.push_back (Key());               // zero key
  return key_list;
}

static Key
parse_key (ArgParser& ap)
{
  auto key_list = parse_key_list (ap);
  if (key_list.size() > 1)
    {
      error ("audiowmark %s: watermark key can at most be set once (--key / --test-key option)\n", ap.command().c_str());
      exit (1);
    }
  return key_list[0];
}

static void
parse_add_options (ArgParser& ap)
{
  string s;
  int i;
  float f;
  ap.parse_opt ("--set-input-label", Params::input_label);
  ap.parse_opt ("--set-output-label", Params::output_label);
  if (ap.parse_opt ("--snr"))
    Params::snr = true;
  if (ap.parse_opt ("--input-format", s))  Params::input_format = parse_format (s);
  if (ap.parse_opt ("--output-format", s)) Params::output_format = parse_format (s);
  if (ap.parse_opt ("--format", s))        Params::input_format = Params::output_format = parse_format (s);
  if (ap.parse_opt ("--raw-input-endian", s))  Params::raw_input_format.set_endian (parse_endian (s));
  if (ap.parse_opt ("--raw-output-endian", s)) Params::raw_output_format.set_endian (parse_endian (s));
  if (ap.parse_opt ("--raw-endian", s))
    {
      Params::raw_input_format.set_endian (parse_endian (s));
      Params::raw_output_format.set_endian (parse_endian (s));
    }
  if (ap.parse_opt ("--raw-input-encoding", s))  parse_encoding (s, Params::raw_input_format);
  if (ap.parse_opt ("--raw-output-encoding", s)) parse_encoding (s, Params::raw_output_format);
  if (ap.parse_opt ("--raw-encoding", s))
    {
      parse_encoding (s, Params::raw_input_format);
      parse_encoding (s, Params::raw_output_format);
    }
  if (ap.parse_opt ("--raw-input-bits", i))  update_raw_bits (Params::raw_input_format, i);
  if (ap.parse_opt ("--raw-output-bits", i)) update_raw_bits (Params::raw_output_format, i);
  if (ap.parse_opt ("--raw-bits", i))
    {
      update_raw_bits (Params::raw_input_format, i);
      update_raw_bits (Params::raw_output_format, i);
    }
  if (ap.parse_opt ("--raw-channels", i))
    {
      Params::raw_input_format.set_channels (i);
      Params::raw_output_format.set_channels (i);
    }
  if (ap.parse_opt ("--raw-rate", i))
    {
      Params::raw_input_format.set_sample_rate (i);
      Params::raw_output_format.set_sample_rate (i);
    }
  if (ap.parse_opt ("--test-no-limiter"))
    Params::test_no_limiter = true;
  if (Params::input_format == Format::RF64)
    {
      error ("audiowmark: using rf64 as input format has no effect\n");
      exit (1);
    }
  if (ap.parse_opt ("--strength", f))
    Params::water_delta = f / 1000;
}

static void
parse_get_options (ArgParser& ap)
{
  string s;
  float f;
  int i;
  ap.parse_opt ("--test-cut", Params::test_cut);
  ap.parse_opt ("--test-truncate", Params::test_truncate);
  if (ap.parse_opt ("--hard"))
    Params::hard = true;
  if (ap.parse_opt ("--test-no-sync"))
    Params::test_no_sync = true;
  int speed_options = 0;
  if (ap.parse_opt ("--detect-speed"))
    {
      Params::detect_speed = true;
      speed_options++;
    }
  if (ap.parse_opt ("--detect-speed-patient"))
    {
      Params::detect_speed_patient = true;
      speed_options++;
    }
  if (ap.parse_opt ("--try-speed", f))
    {
      Params::try_speed = f;
      speed_options++;
    }
  if (speed_options > 1)
    {
      error ("audiowmark: can only use one option: --detect-speed or --detect-speed-patient or --try-speed\n");
      exit (1);
    }
  if (ap.parse_opt ("--test-speed", f))
    Params::test_speed = f;
  if (ap.parse_opt ("--input-format", s) || ap.parse_opt ("--format", s))
    Params::input_format = parse_format (s);
  if (ap.parse_opt ("--raw-input-endian", s) || ap.parse_opt ("--raw-endian", s))   Params::raw_input_format.set_endian (parse_endian (s));
  if (ap.parse_opt ("--raw-input-encoding", s) || ap.parse_opt ("--raw-encoding", s)) parse_encoding (s, Params::raw_input_format);
  if (ap.parse_opt ("--raw-input-bits", i) || ap.parse_opt ("--raw-bits", i))       update_raw_bits (Params::raw_input_format, i);
  if (ap.parse_opt ("--raw-channels", i)) Params::raw_input_format.set_channels (i);
  if (ap.parse_opt ("--raw-rate", i))     Params::raw_input_format.set_sample_rate (i);
  if (ap.parse_opt ("--json", s))
    Params::json_output = s;
  if (ap.parse_opt ("--chunk-size", f))
    {
      if (f < 10)
        {
          error ("audiowmark: --chunk-size needs to be at least 10 minutes\n");
          exit (1);
        }
      Params::get_chunk_size = f;
    }
  if (ap.parse_opt ("--sync-threshold", f))
    Params::sync_threshold2 = f;
  if (ap.parse_opt ("--n-best", i))
    {
      if (i < 0)
        {
          error ("audiowmark: --n-best should not be a negative number\n");
          exit (1);
        }
      Params::get_n_best = i;
    }
  /* --strength is an `add` option only: `get` normalises sync qualities with the default strength (src/audiowmark.cc:806-809) */
}

static vector<string>
parse_positional (ArgParser& ap, const vector<string>& names)
{
  vector<string> args;
  if (ap.parse_args (names.size(), args))
    return args;
  for (const auto& arg : ap.remaining_args())
    if (is_option (arg))
      {
        error ("audiowmark: unsupported option '%s' for command '%s' (use audiowmark -h)\n", arg.c_str(), ap.command().c_str());
        exit (1);
      }
  error ("audiowmark: error parsing arguments for command '%s' (use audiowmark -h)\n\n", ap.command().c_str());
  string msg = "usage: audiowmark " + ap.command() + " [options...]";
  for (const auto& s : names)
    msg += " <" + s + ">";
  error ("%s\n", msg.c_str());
  exit (1);
}

/* ---------------------------------------------------------------- test helpers (src/audiowmark.cc:201-481) */

static int
load_or_complain (WavData& wav, const string& file)
{
  Error err = wav.load (file);
  if (err)
    {
      error ("audiowmark: error loading %s: %s\n", file.c_str(), err.message());
      return 1;
    }
  return 0;
}

static int
save_or_complain (const WavData& wav, const string& file)
{
  Error err = wav.save (file);
  if (err)
    {
      error ("audiowmark: error saving %s: %s\n", file.c_str(), err.message());
      return 1;
    }
  return 0;
}

static int
test_gen_noise (const Key& key, const string& out_file, double seconds, int rate, int bits)
{
  const int channels = 2;
  vector<float> noise;
  Random rng (key, 0, Random::Stream::data_up_down);
  const size_t n = size_t (rate * seconds) * channels;
  noise.reserve (n);
  for (size_t i = 0; i < n; i++)
    noise.push_back (rng.random_double() * 2 - 1);
  return save_or_complain (WavData (noise, channels, rate, bits), out_file);
}

static int
test_speed (const Key& key, int seed)              /* src/audiowmark.cc:389-397 */
{
  Random rng (key, seed, Random::Stream::data_up_down);
  const double low = 0.85, high = 1.15;
  printf ("%.6f\n", low + (rng() / double (UINT64_MAX)) * (high - low));
  return 0;
}

static int
test_change_speed (const string& in_file, const string& out_file, double speed)      /* src/audiowmark.cc:419-438 */
{
  WavData in_data;
  if (load_or_complain (in_data, in_file))
    return 1;
  vector<float> out;
  if (!resample_ratio (in_data.samples().data(), in_data.n_frames(), in_data.n_channels(), 1 / speed, out))
    return 1;
  return save_or_complain (WavData (out, in_data.n_channels(), in_data.sample_rate(), in_data.bit_depth()), out_file);
}

static int
test_resample (const string& in_file, const string& out_file, int new_rate)           /* src/audiowmark.cc:440-458 */
{
  WavData in_data;
  if (load_or_complain (in_data, in_file))
    return 1;
  if (new_rate == in_data.sample_rate())
    {
      error ("audiowmark: test-resample: input already has sample rate %d\n", new_rate);
      return 1;
    }
  vector<float> out;
  if (!resample_ratio (in_data.samples().data(), in_data.n_frames(), in_data.n_channels(), double (new_rate) / in_data.sample_rate(), out))
    return 1;
  return save_or_complain (WavData (out, in_data.n_channels(), new_rate, in_data.bit_depth()), out_file);
}

static int
cut_start (const string& infile, const string& outfile, const string& start_str)
{
  WavData wav;
  if (load_or_complain (wav, infile))
    return 1;
  const size_t start = size_t (atoi_or_die (start_str.c_str())) * wav.n_channels();
  const vector<float>& in = wav.samples();
  vector<float> out (in.begin() + std::min (start, in.size()), in.end());
  return save_or_complain (WavData (out, wav.n_channels(), wav.sample_rate(), wav.bit_depth()), outfile);
}

static int
gentest (const string& infile, const string& outfile)
{
  printf ("generating test sample from '%s' to '%s'\n", infile.c_str(), outfile.c_str());
  WavData wav;
  if (load_or_complain (wav, infile))
    return 1;
  const size_t n = size_t (165) * wav.n_channels() * wav.sample_rate();      // 2:45, room for three blocks
  if (wav.n_values() < n)
    {
      error ("audiowmark: input file %s too short\n", infile.c_str());
      return 1;
    }
  vector<float> out (wav.samples().begin(), wav.samples().begin() + n);
  return save_or_complain (WavData (out, wav.n_channels(), wav.sample_rate(), wav.bit_depth()), outfile);
}

static int
test_subtract (const string& f1, const string& f2, const string& outfile)
{
  WavData a, b;
  if (load_or_complain (a, f1) || load_or_complain (b, f2))
    return 1;
  if (a.n_values() != b.n_values())
    {
      const size_t delta = a.n_values() > b.n_values() ? a.n_values() - b.n_values() : b.n_values() - a.n_values();
      warning ("audiowmark: size mismatch: %zd frames\n", delta / a.n_channels());
      warning (" - %s frames: %zd\n", f1.c_str(), a.n_frames());
      warning (" - %s frames: %zd\n", f2.c_str(), b.n_frames());
    }
  const size_t len = std::min (a.n_values(), b.n_values());
  vector<float> out (len);
  for (size_t i = 0; i < len; i++)
    out[i] = a.samples()[i] - b.samples()[i];
  return save_or_complain (WavData (out, a.n_channels(), a.sample_rate(), a.bit_depth()), outfile);
}

static int
test_snr (const string& orig_file, const string& wm_file)
{
  WavData orig, wm;
  if (load_or_complain (orig, orig_file) || load_or_complain (wm, wm_file))
    return 1;
  if (orig.n_values() != wm.n_values())
    {
      error ("audiowmark: test-snr: files differ in length\n");
      return 1;
    }
  double delta_power = 0, signal_power = 0;
  for (size_t i = 0; i < orig.n_values(); i++)
    {
      const double o = orig.samples()[i], d = orig.samples()[i] - wm.samples()[i];
      delta_power += d * d;
      signal_power += o * o;
    }
  printf ("%f\n", 10 * log10 (signal_power / delta_power));
  return 0;
}

static int
test_clip (const Key& key, const string& in_file, const string& out_file, int seed, int time_seconds)
{
  WavData in;
  if (load_or_complain (in, in_file))
    return 1;
  Random rng (key, seed, Random::Stream::data_up_down);
  size_t start_point, end_point;
  for (;;)
    {
      const size_t values_per_block = frames_per_block() * Params::frame_size * in.n_channels();
      start_point = 2 * values_per_block * rng.random_double();
      start_point /= in.n_channels();
      end_point = start_point + size_t (time_seconds) * in.sample_rate();
      if (end_point < in.n_values() / in.n_channels())
        break;
    }
  vector<float> out (in.samples().begin() + start_point * in.n_channels(), in.samples().begin() + end_point * in.n_channels());
  return save_or_complain (WavData (out, in.n_channels(), in.sample_rate(), in.bit_depth()), out_file);
}

static int
test_info (const string& in_file, const string& property)
{
  WavData in;
  if (load_or_complain (in, in_file))
    return 1;
  if (property == "bit_depth")
    {
      printf ("%d\n", in.bit_depth());
      return 0;
    }
  if (property == "frames")
    {
      printf ("%zd\n", in.n_frames());
      return 0;
    }
  error ("audiowmark: unsupported property for test_info: %s\n", property.c_str());
  return 1;
}

static int
gen_key (const string& outfile, const string& key_name)
{
  string ename;
  for (unsigned char ch : key_name)
    {
      if (ch == '"' || ch == '\\')
        ename += '\\';
      else if (ch < 32)
        {
          error ("audiowmark: bad key name: %d is not allowed as character in key names\n", ch);
          exit (1);
        }
      ename += ch;
    }
  const int fd = open (outfile.c_str(), O_WRONLY | O_CREAT | O_TRUNC, S_IRUSR | S_IWUSR);
  FILE *f = fd >= 0 ? fdopen (fd, "w") : nullptr;
  if (!f)
    {
      if (fd >= 0)
        close (fd);
      error ("audiowmark: error opening file %s: %s\n", outfile.c_str(), strerror (errno));
      return 1;
    }
  fprintf (f, "# watermarking key for audiowmark\n\nkey %s\n", Random::gen_key().c_str());
  if (!key_name.empty())
    fprintf (f, "name \"%s\"\n", ename.c_str());
  fclose (f);
  return 0;
}

int
main (int argc, char **argv)
{
  ArgParser ap (argc, argv);
  vector<string> args;

  if (ap.parse_opt ("--help") || ap.parse_opt ("-h"))
    {
      print_usage();
      return 0;
    }
  if (ap.parse_opt ("--version") || ap.parse_opt ("-v"))
    {
      printf ("audiowmark %s\n", AWM_VERSION);
      return 0;
    }
  if (ap.parse_opt ("--quiet") || ap.parse_opt ("-q"))
    set_log_level (Log::WARNING);
  if (ap.parse_opt ("--strict"))
    Params::strict = true;
  ap.parse_opt ("--gpu-device", Params::gpu_device);

  int rc = 1;
  if (ap.parse_cmd ("add"))
    {
      parse_shared_options (ap);
      parse_add_options (ap);
      Key key = parse_key (ap);
      args = parse_positional (ap, { "input_wav", "watermarked_wav", "message_hex" });
      rc = add_watermark (key, args[0], args[1], args[2]);
    }
  else if (ap.parse_cmd ("get"))
    {
      parse_shared_options (ap);
      parse_get_options (ap);
      vector<Key> key_list = parse_key_list (ap);
      args = parse_positional (ap, { "watermarked_wav" });
      rc = get_watermark (key_list, args[0], "");
    }
  else if (ap.parse_cmd ("cmp"))
    {
      parse_shared_options (ap);
      parse_get_options (ap);
      ap.parse_opt ("--expect-matches", Params::expect_matches);
      vector<Key> key_list = parse_key_list (ap);
      args = parse_positional (ap, { "watermarked_wav", "message_hex" });
      rc = get_watermark (key_list, args[0], args[1]);
    }
  else if (ap.parse_cmd ("gen-key"))
    {
      string key_name;
      ap.parse_opt ("--name", key_name);
      args = parse_positional (ap, { "key_file" });
      rc = gen_key (args[0], key_name);
    }
  else if (ap.parse_cmd ("gentest"))
    {
      args = parse_positional (ap, { "input_wav", "output_wav" });
      rc = gentest (args[0], args[1]);
    }
  else if (ap.parse_cmd ("cut-start"))
    {
      args = parse_positional (ap, { "input_wav", "output_wav", "cut_samples" });
      rc = cut_start (args[0], args[1], args[2]);
    }
  else if (ap.parse_cmd ("test-subtract"))
    {
      args = parse_positional (ap, { "input1_wav", "input2_wav", "output_wav" });
      rc = test_subtract (args[0], args[1], args[2]);
    }
  else if (ap.parse_cmd ("test-snr"))
    {
      args = parse_positional (ap, { "orig_wav", "watermarked_wav" });
      rc = test_snr (args[0], args[1]);
    }
  else if (ap.parse_cmd ("test-clip"))
    {
      parse_shared_options (ap);
      Key key = parse_key (ap);
      args = parse_positional (ap, { "input_wav", "output_wav", "seed", "seconds" });
      rc = test_clip (key, args[0], args[1], atoi_or_die (args[2].c_str()), atoi_or_die (args[3].c_str()));
    }
  else if (ap.parse_cmd ("test-gen-noise"))
    {
      parse_shared_options (ap);
      int bits = 16;
      ap.parse_opt ("--bits", bits);
      Key key = parse_key (ap);
      args = parse_positional (ap, { "output_wav", "seconds", "sample_rate" });
      rc = test_gen_noise (key, args[0], atof_or_die (args[1].c_str()), atoi_or_die (args[2].c_str()), bits);
    }
  else if (ap.parse_cmd ("test-info"))
    {
      parse_shared_options (ap);
      args = parse_positional (ap, { "input_wav", "property" });
      rc = test_info (args[0], args[1]);
    }
  else if (ap.parse_cmd ("test-speed"))
    {
      parse_shared_options (ap);
      Key key = parse_key (ap);
      args = parse_positional (ap, { "seed" });
      rc = test_speed (key, atoi_or_die (args[0].c_str()));
    }
  else if (ap.parse_cmd ("test-change-speed"))
    {
      parse_shared_options (ap);
      args = parse_positional (ap, { "input_wav", "output_wav", "speed" });
      rc = test_change_speed (args[0], args[1], atof_or_die (args[2].c_str()));
    }
  else if (ap.parse_cmd ("test-resample"))
    {
      parse_shared_options (ap);
      args = parse_positional (ap, { "input_wav", "output_wav", "new_rate" });
      rc = test_resample (args[0], args[1], atoi_or_die (args[2].c_str()));
    }
  else if (ap.parse_cmd ("hls-add") || ap.parse_cmd ("hls-prepare"))
    {
      error ("audiowmark: command '%s' is not available in this build (HLS is out of scope)\n", ap.command().c_str());
      rc = 1;
    }
  else if (!ap.remaining_args().empty())
    {
      const string s = ap.remaining_args().front();
      if (is_option (s))
        error ("audiowmark: unsupported global option '%s' (use audiowmark -h)\n", s.c_str());
      else
        error ("audiowmark: unsupported command '%s' (use audiowmark -h)\n", s.c_str());
      rc = 1;
    }
  else
    {
      error ("audiowmark: error parsing commandline args (use audiowmark -h)\n");
      rc = 1;
    }
  Engine::shutdown();
  return rc;
}
