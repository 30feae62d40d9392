// awm_cli_tools.cc -- see awm_cli_tools.hh.  Most helpers are "load a WAV, derive a second sample vector, save it with the
// input's channel count / bit depth": that shape lives in rewrite(), the helpers only say what the new samples are.
#include "awm_cli_tools.hh"

#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include <errno.h>

#include <functional>
#include <vector>

#include "awm_wm.hh"
#include "awm_speed.hh"
#include "awm_tables.hh"
#include "awm_util.hh"

using std::string;
using std::vector;

namespace {

bool
load (WavData& wav, const string& file)
{
  const Error err = wav.load (file);
  if (err)
    error ("audiowmark: error loading %s: %s\n", file.c_str(), err.message());
  return !err;
}

int
save (const vector<float>& samples, int channels, int rate, int bits, const string& file)
{
  const Error err = WavData (samples, channels, rate, bits).save (file);
  if (err)
    error ("audiowmark: error saving %s: %s\n", file.c_str(), err.message());
  return err ? 1 : 0;
}

/* in_file -> make (input, new sample rate) -> out_file; make returns false after reporting a problem itself */
int
rewrite (const string& in_file, const string& out_file, const std::function<bool (const WavData&, vector<float>&, int&)>& make)
{
  WavData in;
  if (!load (in, in_file))
    return 1;
  vector<float> out;
  int rate = in.sample_rate();
  if (!make (in, out, rate))
    return 1;
  return save (out, in.n_channels(), rate, in.bit_depth(), out_file);
}

} // namespace

namespace cli_tools {

int
gen_noise (const Key& key, const string& out_file, double seconds, int rate, int bits)
{
  const int channels = 2;
  Random rng (key, 0, Random::Stream::data_up_down);        // there is no stream of its own for this test signal
  vector<float> noise (size_t (rate * seconds) * channels);
  for (float& v : noise)
    v = rng.random_double() * 2 - 1;
  return save (noise, channels, rate, bits, out_file);
}

int
speed (const Key& key, int seed)
{
  Random rng (key, seed, Random::Stream::data_up_down);
  const double low = 0.85, high = 1.15;
  printf ("%.6f\n", low + (rng() / double (UINT64_MAX)) * (high - low));
  return 0;
}

int
change_speed (const string& in_file, const string& out_file, double speed)
{
  return rewrite (in_file, out_file, [&] (const WavData& in, vector<float>& out, int&)
    {
      return resample_ratio (in.samples().data(), in.n_frames(), in.n_channels(), 1 / speed, out);
    });
}

int
resample (const string& in_file, const string& out_file, int new_rate)
{
  return rewrite (in_file, out_file, [&] (const WavData& in, vector<float>& out, int& rate)
    {
      if (new_rate == in.sample_rate())
        {
          error ("audiowmark: test-resample: input already has sample rate %d\n", new_rate);
          return false;
        }
      rate = new_rate;
      return resample_ratio (in.samples().data(), in.n_frames(), in.n_channels(), double (new_rate) / in.sample_rate(), out);
    });
}

int
cut_start (const string& in_file, const string& out_file, size_t frames)
{
  return rewrite (in_file, out_file, [&] (const WavData& in, vector<float>& out, int&)
    {
      const vector<float>& all = in.samples();
      out.assign (all.begin() + std::min (frames * in.n_channels(), all.size()), all.end());
      return true;
    });
}

int
gentest (const string& in_file, const string& out_file)
{
  printf ("generating test sample from '%s' to '%s'\n", in_file.c_str(), out_file.c_str());
  return rewrite (in_file, out_file, [&] (const WavData& in, vector<float>& out, int&)
    {
      const size_t want = size_t (165) * in.n_channels() * in.sample_rate();      // 2:45, room for three blocks
      if (in.n_values() < want)
        {
          error ("audiowmark: input file %s too short\n", in_file.c_str());
          return false;
        }
      out.assign (in.samples().begin(), in.samples().begin() + want);
      return true;
    });
}

int
clip (const Key& key, const string& in_file, const string& out_file, int seed, int seconds)
{
  return rewrite (in_file, out_file, [&] (const WavData& in, vector<float>& out, int&)
    {
      /* a keyed random start inside the first two blocks; drawn again until the clip fits into the file */
      Random rng (key, seed, Random::Stream::data_up_down);
      const size_t two_blocks = 2 * frames_per_block() * Params::frame_size * in.n_channels();
      size_t first, last;
      do
        {
          first = size_t (two_blocks * rng.random_double()) / in.n_channels();
          last = first + size_t (seconds) * in.sample_rate();
        }
      while (last >= in.n_values() / in.n_channels());
      out.assign (in.samples().begin() + first * in.n_channels(), in.samples().begin() + last * in.n_channels());
      return true;
    });
}

int
subtract (const string& file1, const string& file2, const string& out_file)
{
  WavData b;
  return rewrite (file1, out_file, [&] (const WavData& a, vector<float>& out, int&)
    {
      if (!load (b, file2))
        return false;
      if (a.n_values() != b.n_values())
        {
          const size_t longer = std::max (a.n_values(), b.n_values()), shorter = std::min (a.n_values(), b.n_values());
          warning ("audiowmark: size mismatch: %zd frames\n", (longer - shorter) / a.n_channels());
          warning (" - %s frames: %zd\n", file1.c_str(), a.n_frames());
          warning (" - %s frames: %zd\n", file2.c_str(), b.n_frames());
        }
      out.resize (std::min (a.n_values(), b.n_values()));
      for (size_t i = 0; i < out.size(); i++)
        out[i] = a.samples()[i] - b.samples()[i];
      return true;
    });
}

int
snr (const string& orig_file, const string& wm_file)
{
  WavData orig, wm;
  if (!load (orig, orig_file) || !load (wm, wm_file))
    return 1;
  if (orig.n_values() != wm.n_values())
    {
      error ("audiowmark: test-snr: files differ in length\n");
      return 1;
    }
  double noise = 0, signal = 0;
  for (size_t i = 0; i < orig.n_values(); i++)
    {
      const double o = orig.samples()[i], d = orig.samples()[i] - wm.samples()[i];
      noise += d * d;
      signal += o * o;
    }
  printf ("%f\n", 10 * log10 (signal / noise));
  return 0;
}

int
info (const string& in_file, const string& property)
{
  WavData in;
  if (!load (in, in_file))
    return 1;
  if (property == "bit_depth")
    printf ("%d\n", in.bit_depth());
  else if (property == "frames")
    printf ("%zd\n", in.n_frames());
  else
    {
      error ("audiowmark: unsupported property for test_info: %s\n", property.c_str());
      return 1;
    }
  return 0;
}

int
gen_key (const string& key_file, const string& key_name)
{
  string quoted;
  for (unsigned char ch : key_name)
    {
      if (ch < 32)
        {
          error ("audiowmark: bad key name: %d is not allowed as character in key names\n", ch);
          exit (1);
        }
      if (ch == '"' || ch == '\\')
        quoted += '\\';
      quoted += ch;
    }
  const int fd = open (key_file.c_str(), O_WRONLY | O_CREAT | O_TRUNC, S_IRUSR | S_IWUSR);      // readable by the owner only
  FILE *f = fd >= 0 ? fdopen (fd, "w") : nullptr;
  if (!f)
    {
      if (fd >= 0)
        close (fd);
      error ("audiowmark: error opening file %s: %s\n", key_file.c_str(), strerror (errno));
      return 1;
    }
  fprintf (f, "# watermarking key for audiowmark\n\nkey %s\n", Random::gen_key().c_str());
  if (!key_name.empty())
    fprintf (f, "name \"%s\"\n", quoted.c_str());
  fclose (f);
  return 0;
}

} // namespace cli_tools
