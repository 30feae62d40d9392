/* oracle_c.cc -- float32-exact inner loops of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md): imported by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker;
 * never by the product path.
 *
 * Each function restates one inner loop of the reference (file:line cited) in
 * batch/array form; nothing is copied.  Arithmetic types follow the reference
 * (fp32 sums, fp64 where the reference uses double; glibc log2f/powf/hypotf as
 * the reference links them).  The FFT is the same in-repo FFT that backs the
 * reference build in oracle/_ref (ref_shims/fftw_shim.cc), so that the oracle
 * and oracle/_ref/audiowmark agree bit-for-bit and the pin is exact.
 *
 * Built by oracle/build_oracle.py:  g++ -O2 -shared -fPIC oracle_c.cc ref_shims/fftw_shim.cc
 * (no -ffast-math, no -march: baseline x86-64 has no FMA, so every * and + rounds
 * separately exactly like the reference build).
 */
#include "ref_shims/fftw3.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

struct Plans
{
  int n = 0;
  fftwf_plan fwd = nullptr, inv = nullptr;
};

Plans&
plans_for (int n)
{
  static Plans p1024, p512, pother;
  Plans *p = (n == 1024) ? &p1024 : (n == 512) ? &p512 : &pother;
  if (p->n != n)
    {
      if (p->fwd) fftwf_destroy_plan (p->fwd);
      if (p->inv) fftwf_destroy_plan (p->inv);
      p->fwd = fftwf_plan_dft_r2c_1d (n, nullptr, nullptr, 0);
      p->inv = fftwf_plan_dft_c2r_1d (n, nullptr, nullptr, 0);
      p->n = n;
    }
  return *p;
}

} // namespace

extern "C" {

/* FFTProcessor::fft on raw buffers (src/fft.cc:82-86): out = float[n+2] interleaved re/im */
void
orc_rfft (const float *in, float *out, int n, int64_t count)
{
  Plans& p = plans_for (n);
  std::vector<float> buf (n + 2);
  for (int64_t i = 0; i < count; i++)
    {
      memcpy (buf.data(), in + i * n, sizeof (float) * n);
      fftwf_execute_dft_r2c (p.fwd, buf.data(), (fftwf_complex *) (out + i * (n + 2)));
    }
}

/* FFTProcessor::ifft (src/fft.cc:88-92): unnormalised c2r; in = float[n+2], out = float[n] */
void
orc_irfft (const float *in, float *out, int n, int64_t count)
{
  Plans& p = plans_for (n);
  std::vector<float> buf (n + 2);
  for (int64_t i = 0; i < count; i++)
    {
      memcpy (buf.data(), in + i * (n + 2), sizeof (float) * (n + 2));
      fftwf_execute_dft_c2r (p.inv, (fftwf_complex *) buf.data(), out + i * n);
    }
}

/* FFTAnalyzer::run_fft (src/wmcommon.cc:91-121) for a list of frame start positions.
 *   samples: interleaved, n_channels; starts[j] = first sample-frame of job j
 *   out: [n_jobs][n_channels][n+2] floats (re/im interleaved, n/2+1 complex)
 */
void
orc_analyze_frames (const float *samples, int n_channels, const int64_t *starts, int64_t n_jobs,
                    const float *window, int n, float *out)
{
  Plans& p = plans_for (n);
  std::vector<float> frame (n + 2);
  for (int64_t j = 0; j < n_jobs; j++)
    for (int ch = 0; ch < n_channels; ch++)
      {
        int64_t pos = starts[j] * n_channels + ch;
        for (int x = 0; x < n; x++)
          {
            frame[x] = samples[pos] * window[x];
            pos += n_channels;
          }
        fftwf_execute_dft_r2c (p.fwd, frame.data(), (fftwf_complex *) (out + (j * n_channels + ch) * (n + 2)));
      }
}

/* db_from_complex (src/wmcommon.hh:204-224) */
static inline float
db_from_complex (float re, float im, float min_db)
{
  const float abs2 = re * re + im * im;
  if (abs2 > 0)
    {
      const float log2_db_factor = 3.01029995663981;
      return log2f (abs2) * log2_db_factor;
    }
  return min_db;
}

/* elementwise db_from_complex over an array of complex values (re/im interleaved) */
void
orc_db (const float *spect, float *db, int64_t count, float min_db)
{
  for (int64_t i = 0; i < count; i++)
    db[i] = db_from_complex (spect[2 * i], spect[2 * i + 1], min_db);
}

/* channel-summed band dB as in SyncFinder::sync_fft (src/syncfinder.cc:592-598):
 *   spect: [n_jobs][n_channels][n+2]; out: [n_jobs][n_bands] (+= over channels, starting from 0)
 */
void
orc_db_bands (const float *spect, int64_t n_jobs, int n_channels, int n, int min_band, int max_band, float *out)
{
  const int n_bands = max_band - min_band + 1;
  for (int64_t j = 0; j < n_jobs; j++)
    {
      float *o = out + j * n_bands;
      for (int b = 0; b < n_bands; b++)
        o[b] = 0;
      for (int ch = 0; ch < n_channels; ch++)
        {
          const float *s = spect + (j * n_channels + ch) * (n + 2);
          for (int i = min_band; i <= max_band; i++)
            o[i - min_band] += db_from_complex (s[2 * i], s[2 * i + 1], -96);
        }
    }
}

/* apply_frame_mod (src/wmadd.cc:61-84): frame_mod[n_bins] in {0 keep, 1 up, 2 down};
 * spect/delta: [n+2] floats for one channel; delta must be zero-initialised by the caller */
void
orc_apply_frame_mod (const uint8_t *frame_mod, int n_mod, const float *spect, float *delta, double water_delta)
{
  const float min_mag = 1e-7;
  for (int i = 0; i < n_mod; i++)
    {
      if (frame_mod[i] == 0)
        continue;
      const int data_bit_sign = (frame_mod[i] == 1) ? 1 : -1;
      const float re = spect[2 * i], im = spect[2 * i + 1];
      const float mag = hypotf (re, im);          /* std::abs (std::complex<float>) */
      if (mag > min_mag)
        {
          const float mag_factor = powf (mag, -water_delta * data_bit_sign);
          const float f = mag_factor - 1;
          delta[2 * i]     = re * f;               /* complex<float> * float */
          delta[2 * i + 1] = im * f;
        }
    }
}

/* SyncFinder::sync_decode (src/syncfinder.cc:116-153) for many start frames.
 *   db: [n_frames][n_bands] row-major by frame, have[n_frames]
 *   sync layout: bit b owns entries ent_off[b] .. ent_off[b+1]-1; entry e has frame ent_frame[e]
 *   and 30 up / 30 down band indices ent_up[e*30+i], ent_down[e*30+i]
 *   norm_div = min (water_delta, 0.080) * 2.9 applied as two divisions like :90
 */
void
orc_sync_decode (const float *db, const char *have, int n_bands,
                 int n_bits, const int *ent_off, const int *ent_frame, const int *ent_up, const int *ent_down, int n_ud,
                 const int64_t *start_frames, int64_t n_starts, double water_delta, double *quality_out)
{
  for (int64_t s = 0; s < n_starts; s++)
    {
      const int64_t start_frame = start_frames[s];
      double sync_quality = 0;
      int bit_count = 0;
      for (int bit = 0; bit < n_bits; bit++)
        {
          float umag = 0, dmag = 0;
          int frame_bit_count = 0;
          for (int e = ent_off[bit]; e < ent_off[bit + 1]; e++)
            {
              const int64_t f = start_frame + ent_frame[e];
              if (have[f])
                {
                  const float *row = db + f * n_bands;
                  for (int i = 0; i < n_ud; i++)
                    {
                      umag += row[ent_up[e * n_ud + i]];
                      dmag += row[ent_down[e * n_ud + i]];
                    }
                  frame_bit_count++;
                }
            }
          /* bit_quality (src/syncfinder.cc:94-114) */
          const int expect_data_bit = bit & 1;
          double raw_bit;
          if (umag == 0 || dmag == 0)
            raw_bit = 0;
          else if (umag < dmag)
            raw_bit = 1 - umag / dmag;
          else
            raw_bit = dmag / umag - 1;
          sync_quality += (expect_data_bit ? raw_bit : -raw_bit) * frame_bit_count;
          bit_count += frame_bit_count;
        }
      if (bit_count)
        sync_quality /= bit_count;
      /* normalize_sync_quality (src/syncfinder.cc:80-91) */
      const double wd = water_delta < 0.080 ? water_delta : 0.080;
      quality_out[s] = sync_quality / wd / 2.9;
    }
}

/* mix_decode (src/wmget.cc:67-108)
 *   spect: [n_spect][n+2] with n_spect = frames_per_block * n_channels (frame-major, then channel)
 *   mix entries: frame/up/down arrays of length frame_count * bands_per_frame
 *   out: frame_count / frames_per_bit floats
 */
void
orc_mix_decode (const float *spect, int64_t n_spect, int n, int n_channels,
                const int *mix_frame, const int *mix_up, const int *mix_down,
                int frame_count, int bands_per_frame, int frames_per_bit, float *out)
{
  const double min_db = -96;
  auto db = [&] (int64_t idx, int bin) { const float *s = spect + idx * (n + 2); return db_from_complex (s[2 * bin], s[2 * bin + 1], min_db); };
  double umag = 0, dmag = 0;
  int o = 0;
  for (int f = 0; f < frame_count; f++)
    {
      for (int ch = 0; ch < n_channels; ch++)
        for (int frame_b = 0; frame_b < bands_per_frame; frame_b++)
          {
            const int b = f * bands_per_frame + frame_b;
            const int64_t index = int64_t (mix_frame[b]) * n_channels + ch;
            const int64_t next_index = (index + n_channels) < n_spect ? index + n_channels : index - n_channels;
            const int64_t prev_index = (index - n_channels) >= 0 ? index - n_channels : index + n_channels;
            const int u = mix_up[b], d = mix_down[b];

            umag += db (index, u);
            umag -= (db (prev_index, u) + db (next_index, u)) * 0.5;   /* float + float, then * double 0.5 */
            dmag += db (index, d);
            dmag -= (db (prev_index, d) + db (next_index, d)) * 0.5;
          }
      if ((f % frames_per_bit) == (frames_per_bit - 1))
        {
          out[o++] = umag - dmag;
          umag = 0;
          dmag = 0;
        }
    }
}

/* linear_decode (src/wmget.cc:110-152): same gather but per-frame up/down lists in frame order */
void
orc_linear_decode (const float *spect, int64_t n_spect, int n, int n_channels,
                   const int *data_frame, const int *up, const int *down,
                   int frame_count, int bands_per_frame, int frames_per_bit, float *out)
{
  const double min_db = -96;
  auto db = [&] (int64_t idx, int bin) { const float *s = spect + idx * (n + 2); return db_from_complex (s[2 * bin], s[2 * bin + 1], min_db); };
  double umag = 0, dmag = 0;
  int o = 0;
  for (int f = 0; f < frame_count; f++)
    {
      for (int ch = 0; ch < n_channels; ch++)
        {
          const int64_t index = int64_t (data_frame[f]) * n_channels + ch;
          const int64_t next_index = (index + n_channels) < n_spect ? index + n_channels : index - n_channels;
          const int64_t prev_index = (index - n_channels) >= 0 ? index - n_channels : index + n_channels;
          for (int i = 0; i < bands_per_frame; i++)
            {
              const int u = up[f * bands_per_frame + i];
              umag += db (index, u);
              umag -= 0.5 * (db (prev_index, u) + db (next_index, u));
            }
          for (int i = 0; i < bands_per_frame; i++)
            {
              const int d = down[f * bands_per_frame + i];
              dmag += db (index, d);
              dmag -= 0.5 * (db (prev_index, d) + db (next_index, d));
            }
        }
      if ((f % frames_per_bit) == (frames_per_bit - 1))
        {
          out[o++] = umag - dmag;
          umag = 0;
          dmag = 0;
        }
    }
}

/* conv_decode_soft (src/convcode.cc:128-213): Viterbi over 2^order states.
 *   generators[rate]; coded[n_coded] with n_coded % rate == 0
 *   decoded: n_coded / rate ints (including the `order` termination bits; caller strips them)
 *   returns error = final metric of state 0 / n_coded
 * Update order and tie rule as the reference: old states ascending, bit 0 then 1, strict '<'.
 */
float
orc_viterbi (const unsigned *generators, int rate, int order, const float *coded, int64_t n_coded, int *decoded)
{
  const unsigned state_count = 1u << order, state_mask = state_count - 1;
  const int64_t steps = n_coded / rate;

  std::vector<float> state2bits (size_t (state_count) * rate);
  for (unsigned state = 0; state < state_count; state++)
    for (int p = 0; p < rate; p++)
      state2bits[size_t (state) * rate + p] = __builtin_parity (state & generators[p]);

  std::vector<float>    delta_old (state_count, -1.f), delta_new (state_count);
  std::vector<uint16_t> last_state (size_t (steps + 1) * state_count);  /* order <= 16 */
  std::vector<uint8_t>  last_bit (size_t (steps + 1) * state_count);
  delta_old[0] = 0;

  for (int64_t t = 0; t < steps; t++)
    {
      std::fill (delta_new.begin(), delta_new.end(), -1.f);
      uint16_t *ls = &last_state[size_t (t + 1) * state_count];
      uint8_t  *lb = &last_bit[size_t (t + 1) * state_count];
      const float *c = coded + t * rate;
      for (unsigned state = 0; state < state_count; state++)
        {
          if (delta_old[state] >= 0)
            for (int bit = 0; bit < 2; bit++)
              {
                const unsigned new_state = ((state << 1) | bit) & state_mask;
                float delta = delta_old[state];
                const float *sb = &state2bits[size_t (new_state) * rate];
                for (int p = 0; p < rate; p++)
                  delta += (c[p] - sb[p]) * (c[p] - sb[p]);
                if (delta < delta_new[new_state] || delta_new[new_state] < 0)
                  {
                    delta_new[new_state] = delta;
                    ls[new_state] = state;
                    lb[new_state] = bit;
                  }
              }
        }
      delta_old.swap (delta_new);
    }
  unsigned state = 0;
  const float err = delta_old[state] / n_coded;
  for (int64_t idx = steps; idx > 0; idx--)
    {
      decoded[idx - 1] = last_bit[size_t (idx) * state_count + state];
      state = last_state[size_t (idx) * state_count + state];
    }
  return err;
}

/* glibc transcendental probes so tests can pin device math against the host libm */
void orc_log2f (const float *in, float *out, int64_t n) { for (int64_t i = 0; i < n; i++) out[i] = log2f (in[i]); }
void orc_powf (const float *a, const float *b, float *out, int64_t n) { for (int64_t i = 0; i < n; i++) out[i] = powf (a[i], b[i]); }
void orc_hypotf (const float *a, const float *b, float *out, int64_t n) { for (int64_t i = 0; i < n; i++) out[i] = hypotf (a[i], b[i]); }

} // extern "C"
