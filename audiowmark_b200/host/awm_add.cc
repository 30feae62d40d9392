// awm_add.cc -- `audiowmark add`: add_stream_watermark / add_watermark (reference src/wmadd.cc:448-657).
// The reference pulls 1024-frame blocks through WatermarkGen / WatermarkSynth / Limiter on one CPU
// thread; here the whole stream is handed to awm_embed, which reproduces the same stream semantics
// (frame numbering from 2*fpb - 250, zero padding of the tail, limiter on the zero-extended signal).
#include "awm_wm.hh"
#include "awm_speed.hh"
#include "awm_engine.hh"
#include "awm_tables.hh"
#include "awm_util.hh"

#include <math.h>
#include <functional>

using std::string;
using std::vector;

/* number of WatermarkGen::run calls the reference loop makes (src/wmadd.cc:520-589): zero frames are fed
 * after EOF until every input frame has been written; synth delays one frame (src/wmadd.cc:240-249) and the
 * limiter keeps one block of look-ahead (src/limiter.cc:53-58).  Only needed for the "Data Blocks" line. */
static size_t
gen_runs (size_t n_frames, bool limiter_on, size_t limiter_block)
{
  const size_t N = Params::frame_size;
  if (n_frames == 0)
    return 0;
  if (!limiter_on)
    return 1 + (n_frames + N - 1) / N;
  const size_t need = ((n_frames + limiter_block - 1) / limiter_block + 1) * limiter_block;
  return 1 + (need + N - 1) / N;
}

/* Frame counts of the add loop when the input is not at the watermark rate (src/wmadd.cc:520-589 with a
 * WatermarkResampler, :353-430).  Every iteration pushes one 1024-frame block (zero frames after EOF) into the input
 * resampler; whole 1024-frames at the watermark rate go through WatermarkGen::run and the output resampler; whatever
 * that has delivered is mixed, counted for --snr and handed to the limiter, which holds back until two complete blocks
 * are buffered (src/limiter.cc:53-58).  The loop stops at the first short read with everything written.
 *   n_emit: frames that went through the mixer;  gen_runs: WatermarkGen::run calls */
void
resampled_add_plan (size_t n_frames, int sample_rate, bool limiter_on, size_t limiter_block, size_t& n_emit, size_t& gen_runs_out)
{
  const size_t N = Params::frame_size;
  const double r_in = double (Params::mark_sample_rate) / sample_rate, r_out = double (sample_rate) / Params::mark_sample_rate;
  size_t total_in = 0, total_out = 0;
  n_emit = gen_runs_out = 0;
  for (size_t iteration = 1;; iteration++)
    {
      const size_t real = std::min (N, n_frames - total_in);
      total_in += real;
      if (real < N && total_in == total_out)
        break;
      const size_t at_mark_rate = resample_stream_available (N * iteration, r_in);
      gen_runs_out = at_mark_rate / N;
      n_emit = gen_runs_out ? resample_stream_available (gen_runs_out * N, r_out) : 0;
      size_t released = n_emit;
      if (limiter_on)
        released = n_emit / limiter_block >= 2 ? (n_emit / limiter_block - 1) * limiter_block : 0;
      total_out = std::min (released, total_in);
    }
}

static void
info_format (const string& label, const RawFormat& format)
{
  const char *e = format.encoding() == Encoding::SIGNED ? "signed" : format.encoding() == Encoding::UNSIGNED ? "unsigned" : "float";
  info ("%-13s %d Hz, %d Channels, %d Bit (%s %s-endian)\n", (label + ":").c_str(), format.sample_rate(), format.n_channels(),
        format.bit_depth(), e, format.endian() == RawFormat::LITTLE ? "little" : "big");
}

int
add_watermark_buffer (const Key& key, const float *in, float *out, size_t n_frames, int n_channels, int sample_rate,
                      const string& bits, AddStats *stats, uint64_t first_frame_number)
{
  const vector<int> bitvec = parse_payload (bits);
  if (bitvec.empty())
    return 1;
  awm_ctx *ctx = Engine::ctx();
  if (!ctx || !Engine::set_embed_tables (key, bitvec))
    return 1;
  const int limiter_block = Params::test_no_limiter ? 0 : int (sample_rate * int (Params::limiter_block_size_ms) / 1000);
  double snr_power[2] = { 0, 0 };
  if (sample_rate != Params::mark_sample_rate)
    {
      if (first_frame_number)
        {
          error ("audiowmark: sharded embedding is only available at %d Hz\n", Params::mark_sample_rate);
          return 1;
        }
      size_t n_emit = 0, runs = 0;
      resampled_add_plan (n_frames, sample_rate, !Params::test_no_limiter, sample_rate * int (Params::limiter_block_size_ms) / 1000, n_emit, runs);
      if (awm_embed_resampled (ctx, in, out, n_frames, n_channels, sample_rate, Params::mark_sample_rate, std::max (n_emit, n_frames), Params::frames_pad_start,
                               Params::water_delta, limiter_block, Params::limiter_ceiling, (Params::snr || stats) ? snr_power : nullptr))
        {
          error ("audiowmark: embedding failed: %s\n", awm_last_error (ctx));
          return 1;
        }
      if (stats)
        {
          const size_t fpb = frames_per_block(), f0 = 2 * fpb - Params::frames_pad_start;
          const int blocks = int ((f0 + runs) / fpb - f0 / fpb);
          stats->data_blocks = std::max (blocks - 1, 0);
          stats->snr_db = snr_power[0] > 0 ? 10 * log10 (snr_power[1] / snr_power[0]) : INFINITY;
        }
      return 0;
    }
  if (awm_embed (ctx, in, out, n_frames, n_channels, first_frame_number, Params::frames_pad_start, Params::water_delta,
                 limiter_block, Params::limiter_ceiling, (Params::snr || stats) ? snr_power : nullptr))
    {
      error ("audiowmark: embedding failed: %s\n", awm_last_error (ctx));
      return 1;
    }
  if (stats)
    {
      const size_t fpb = frames_per_block(), f0 = 2 * fpb - Params::frames_pad_start;
      const size_t runs = gen_runs (n_frames, !Params::test_no_limiter, sample_rate * int (Params::limiter_block_size_ms) / 1000);
      const int blocks = int ((f0 + runs) / fpb - f0 / fpb);
      stats->data_blocks = std::max (blocks - 1, 0);         // the first (partial B) block is padding
      stats->snr_db = snr_power[0] > 0 ? 10 * log10 (snr_power[1] / snr_power[0]) : INFINITY;
    }
  return 0;
}

/* `add` between two 16 bit PCM buffers in host memory (44.1 kHz): the int <-> float conversions of SFInputStream / SFOutputStream
 * run on the device (awm_embed_s16), only 2 + 2 bytes per sample cross PCIe */
int
add_watermark_buffer_s16 (const Key& key, const int16_t *in, int16_t *out, size_t n_frames, int n_channels, int sample_rate,
                          const string& bits, AddStats *stats, uint64_t first_frame_number)
{
  const vector<int> bitvec = parse_payload (bits);
  if (bitvec.empty())
    return 1;
  if (sample_rate != Params::mark_sample_rate)
    {
      error ("audiowmark: 16 bit buffers are only accepted at %d Hz (use the float entry point for other rates)\n", Params::mark_sample_rate);
      return 1;
    }
  awm_ctx *ctx = Engine::ctx();
  if (!ctx || !Engine::set_embed_tables (key, bitvec))
    return 1;
  const int limiter_block = Params::test_no_limiter ? 0 : int (sample_rate * int (Params::limiter_block_size_ms) / 1000);
  double snr_power[2] = { 0, 0 };
  if (awm_embed_s16 (ctx, in, out, n_frames, n_channels, first_frame_number, Params::frames_pad_start, Params::water_delta,
                     limiter_block, Params::limiter_ceiling, (Params::snr || stats) ? snr_power : nullptr))
    {
      error ("audiowmark: embedding failed: %s\n", awm_last_error (ctx));
      return 1;
    }
  if (stats)
    {
      const size_t fpb = frames_per_block(), f0 = 2 * fpb - Params::frames_pad_start;
      const size_t runs = gen_runs (n_frames, !Params::test_no_limiter, sample_rate * int (Params::limiter_block_size_ms) / 1000);
      const int blocks = int ((f0 + runs) / fpb - f0 / fpb);
      stats->data_blocks = std::max (blocks - 1, 0);
      stats->snr_db = snr_power[0] > 0 ? 10 * log10 (snr_power[1] / snr_power[0]) : INFINITY;
    }
  return 0;
}

/* Part of the stream a window has to embed so that [own_start, own_end) comes out exactly as in a run over the whole stream:
 * one 1024-frame before the first limiter block that influences the owned range (synthesis-window tail of the frame before) and
 * everything up to the end of the limiter block after the last owned one (the gain ramp looks one block ahead), rounded to whole
 * 1024-frames.  n_total = ~0 while the end of the stream is not known. */
static void
window_range (size_t own_start, size_t own_end, size_t n_total, size_t limiter_block, size_t& ext_start, size_t& ext_end)
{
  const size_t N = Params::frame_size;
  if (own_start == 0)
    ext_start = 0;
  else
    {
      const size_t b = own_start / limiter_block;                   // first owned limiter block
      const size_t lo = (b ? b - 1 : 0) * limiter_block;            // the block before it must be complete
      ext_start = lo / N * N >= N ? lo / N * N - N : 0;
    }
  if (own_end >= n_total)
    ext_end = n_total;
  else
    {
      const size_t hi = ((own_end - 1) / limiter_block + 2) * limiter_block;
      ext_end = std::min (ext_start + (hi - ext_start + N - 1) / N * N + N, n_total);
    }
}

/* add_stream_watermark's loop (src/wmadd.cc:504-589) with bounded memory: the stream is embedded window by window.  A window
 * is run through awm_embed with the halo window_range() asks for and with first_frame_number = its position in the stream, so the
 * table row of every frame and the limiter blocks are those of a run over the whole stream and the owned part of the result
 * is bit identical to it (tests/test_gpu_sharding.py checks exactly this property for sharded embedding).  Reads block only
 * until one window + about two seconds of look-ahead are buffered: a pipe gets its first output after ~100 s of input, not at EOF.
 * zero_frames: the input continues a stream that began zero_frames earlier with silence (HLS segments, src/wmadd.cc:504-519):
 * positions, table rows and limiter blocks count from there, nothing is written for the silent part.
 * window_frames = 0: default window (4096 1024-frames = 95 s); tests pass small windows. */
int
add_watermark_windowed (const Key& key, const std::function<Error (vector<float>&, size_t)>& read, const std::function<Error (const vector<float>&)>& write,
                        int n_channels, int sample_rate, const string& bits, size_t zero_frames, size_t window_frames, AddStats *stats, size_t *frames_written)
{
  const size_t N = Params::frame_size, C = size_t (n_channels);
  const size_t L = size_t (sample_rate) * size_t (Params::limiter_block_size_ms) / 1000;      // limiter block; also sizes the halo without limiter
  const size_t W = (window_frames ? window_frames : 4096) * N;
  const size_t unknown = ~size_t (0);
  const vector<int> bitvec = parse_payload (bits);
  if (bitvec.empty())
    return 1;
  awm_ctx *ctx = Engine::ctx();
  if (!ctx || !Engine::set_embed_tables (key, bitvec))
    return 1;
  const int limiter_block = Params::test_no_limiter ? 0 : int (L);
  vector<float> pending;                         // stream positions [p0, p1), interleaved
  size_t p0 = 0, p1 = 0, own_start = zero_frames, written = 0;
  double snr_delta = 0, snr_signal = 0;
  {
    /* the silent prefix is never materialised beyond what the first window's halo reaches into */
    size_t e0, e1;
    window_range (own_start, own_start + 1, unknown, L, e0, e1);
    p0 = e0;
    p1 = zero_frames;
    pending.assign ((p1 - p0) * C, 0.f);
  }
  bool eof = false;
  vector<float> block, out;
  while (!eof || own_start < p1)
    {
      size_t e0, e1;
      window_range (own_start, own_start + W, unknown, L, e0, e1);
      while (!eof && p1 < e1)                    // the window and its look-ahead
        {
          const Error err = read (block, std::min<size_t> (e1 - p1, 1 << 16));
          if (err)
            {
              error ("audiowmark: input stream read failed: %s\n", err.message());
              return 1;
            }
          if (block.empty())
            eof = true;
          pending.insert (pending.end(), block.begin(), block.end());
          p1 += block.size() / C;
        }
      const size_t n_total = eof ? p1 : unknown;
      const size_t own_end = eof ? p1 : own_start + W;
      if (own_end <= own_start)
        break;
      window_range (own_start, own_end, n_total, L, e0, e1);
      out.resize ((e1 - e0) * C);
      double snr_power[2] = { 0, 0 };
      /* the --snr sums count the watermark signal before the limiter, over this window's own part (the last window also takes the
       * zero padded frames after the end of the input, like the reference loop) */
      if (awm_embed_window (ctx, pending.data() + (e0 - p0) * C, out.data(), e1 - e0, n_channels, e0 / N, Params::frames_pad_start, Params::water_delta,
                            limiter_block, Params::limiter_ceiling, own_start - e0, eof ? ~uint64_t (0) : own_end - e0, (Params::snr || stats) ? snr_power : nullptr))
        {
          error ("audiowmark: embedding failed: %s\n", awm_last_error (ctx));
          return 1;
        }
      snr_delta += snr_power[0];
      snr_signal += snr_power[1];
      const float *o = out.data() + (own_start - e0) * C;
      const Error err = write (vector<float> (o, o + (own_end - own_start) * C));
      if (err)
        {
          error ("audiowmark output write failed: %s\n", err.message());
          return 1;
        }
      written += own_end - own_start;
      own_start = own_end;
      if (eof)
        break;
      window_range (own_start, own_start + 1, unknown, L, e0, e1);   // what the next window still needs of the buffer
      pending.erase (pending.begin(), pending.begin() + (e0 - p0) * C);
      p0 = e0;
    }
  if (frames_written)
    *frames_written = written;
  if (stats)
    {
      const size_t fpb = frames_per_block(), f0 = 2 * fpb - Params::frames_pad_start;
      const size_t runs = gen_runs (p1, !Params::test_no_limiter, L);
      const int blocks = int ((f0 + runs) / fpb - f0 / fpb);
      stats->data_blocks = std::max (blocks - 1, 0);         // the first (partial B) block is padding
      stats->snr_db = snr_delta > 0 ? 10 * log10 (snr_signal / snr_delta) : INFINITY;
    }
  return 0;
}

int
add_stream_watermark (const Key& key, AudioInputStream *in_stream, AudioOutputStream *out_stream, const string& bits, size_t zero_frames)
{
  auto bitvec = parse_payload (bits);
  if (bitvec.empty())
    return 1;
  if (in_stream->sample_rate() != out_stream->sample_rate())
    {
      error ("audiowmark: input sample rate (%d) and output sample rate (%d) don't match\n", in_stream->sample_rate(), out_stream->sample_rate());
      return 1;
    }
  if (in_stream->n_channels() != out_stream->n_channels())
    {
      error ("audiowmark: input channels (%d) and output channels (%d) don't match\n", in_stream->n_channels(), out_stream->n_channels());
      return 1;
    }
  info ("Message:      %s\n", bit_vec_to_str (bitvec).c_str());
  info ("Strength:     %.6g\n\n", Params::water_delta * 1000);
  if (in_stream->n_frames() == AudioInputStream::N_FRAMES_UNKNOWN)
    info ("Time:         unknown\n");
  else
    {
      const size_t orig_seconds = in_stream->n_frames() / in_stream->sample_rate();
      info ("Time:         %zd:%02zd\n", orig_seconds / 60, orig_seconds % 60);
    }
  info ("Sample Rate:  %d\n", in_stream->sample_rate());
  info ("Channels:     %d\n", in_stream->n_channels());

  AddStats stats;
  size_t total_output_frames = 0;
  Error err;
  if (in_stream->sample_rate() != Params::mark_sample_rate)
    {
      /* other sample rates go through the resamplers around the embedder (awm_embed_resampled): whole stream at once */
      if (zero_frames)
        {
          error ("audiowmark: stream offsets are only supported at %d Hz\n", Params::mark_sample_rate);
          return 1;
        }
      WavData wav;
      err = wav.load (in_stream);
      if (err)
        {
          error ("audiowmark: input stream read failed: %s\n", err.message());
          return 1;
        }
      vector<float> out (wav.n_values());
      if (add_watermark_buffer (key, wav.samples().data(), out.data(), wav.n_frames(), wav.n_channels(), wav.sample_rate(), bits, &stats))
        return 1;
      vector<float>().swap (wav.mutable_samples());
      err = out_stream->write_frames (out);
      if (err)
        {
          error ("audiowmark output write failed: %s\n", err.message());
          return 1;
        }
      total_output_frames = out.size() / std::max (in_stream->n_channels(), 1);
    }
  else
    {
      auto read = [&] (vector<float>& samples, size_t count) { return in_stream->read_frames (samples, count); };
      auto write = [&] (const vector<float>& samples) { return out_stream->write_frames (samples); };
      if (add_watermark_windowed (key, read, write, in_stream->n_channels(), in_stream->sample_rate(), bits, zero_frames, 0, &stats, &total_output_frames))
        return 1;
    }
  if (Params::snr)
    info ("SNR:          %f dB\n", stats.snr_db);
  info ("Data Blocks:  %d\n", stats.data_blocks);

  if (in_stream->n_frames() != AudioInputStream::N_FRAMES_UNKNOWN && total_output_frames != in_stream->n_frames())
    {
      auto msg = string_printf ("unexpected EOF; input frames (%zd) != output frames (%zd)", in_stream->n_frames(), total_output_frames);
      if (Params::strict)
        {
          error ("audiowmark: error: %s\n", msg.c_str());
          return 1;
        }
      warning ("audiowmark: warning: %s\n", msg.c_str());
    }
  err = out_stream->close();
  if (err)
    {
      error ("audiowmark: closing output stream failed: %s\n", err.message());
      return 1;
    }
  return 0;
}

int
add_watermark (const Key& key, const string& infile, const string& outfile, const string& bits)
{
  Error err;
  std::unique_ptr<AudioInputStream> in_stream = AudioInputStream::create (infile, err);
  if (err)
    {
      error ("audiowmark: error opening %s: %s\n", infile.c_str(), err.message());
      return 1;
    }
  /* output keeps the input's depth / encoding, but at least 16 bit signed */
  int out_bit_depth = in_stream->bit_depth();
  Encoding out_encoding = in_stream->encoding();
  if (in_stream->bit_depth() < 16)
    {
      out_bit_depth = 16;
      out_encoding = Encoding::SIGNED;
    }
  std::unique_ptr<AudioOutputStream> out_stream = AudioOutputStream::create (outfile, in_stream->n_channels(), in_stream->sample_rate(),
                                                                           out_bit_depth, out_encoding, in_stream->n_frames(), err);
  if (err)
    {
      error ("audiowmark: error writing to %s: %s\n", outfile.c_str(), err.message());
      return 1;
    }
  info ("Input:        %s\n", Params::input_label.size() ? Params::input_label.c_str() : infile.c_str());
  if (Params::input_format == Format::RAW)
    info_format ("Raw Input", Params::raw_input_format);
  info ("Output:       %s\n", Params::output_label.size() ? Params::output_label.c_str() : outfile.c_str());
  if (Params::output_format == Format::RAW)
    info_format ("Raw Output", Params::raw_output_format);
  return add_stream_watermark (key, in_stream.get(), out_stream.get(), bits, 0);
}
