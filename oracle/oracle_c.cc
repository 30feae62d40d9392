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
#include <time.h>
#include <vector>
#include <algorithm>

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
  /* independent jobs: spread over threads (execute keeps its scratch on the caller's stack), same arithmetic per job */
#pragma omp parallel
  {
    std::vector<float> frame (n + 2);
#pragma omp for schedule(static)
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

/* std::sort as the reference's selectors call it (src/syncfinder.cc:366,388: descending |q - mean|, comparator on the value only).
 * std::sort is not stable: with equal keys (digital silence: every quality is 0) the surviving elements depend on libstdc++'s
 * introsort and on the input sequence, so the oracle asks the same library for the permutation.  perm[i] = input position of the
 * element that ends up at output position i. */
void
orc_std_sort_desc (const double *key, int64_t n, int64_t *perm)
{
  struct Item { int64_t pos; double key; double pad; };     /* three words like SyncFinder::SearchScore (moves are by value either way) */
  std::vector<Item> v (n);
  for (int64_t i = 0; i < n; i++)
    v[i] = Item { i, key[i], 0 };
  std::sort (v.begin(), v.end(), [] (Item& a, Item& b) { return a.key > b.key; });
  for (int64_t i = 0; i < n; i++)
    perm[i] = v[i].pos;
}

/* microseconds per 1024-point r2c transform of the FFT shim, one thread (bench.py reports it next to pocketfft's, so that the
 * distance between the reference build's FFT and a tuned library is on record) */
double
orc_fft_r2c_us (int reps)
{
  const int N = 1024;
  float *in = (float *) fftwf_malloc (sizeof (float) * (N + 2));
  fftwf_complex *out = (fftwf_complex *) fftwf_malloc (sizeof (float) * (N + 2));
  for (int i = 0; i < N; i++)
    in[i] = float ((i * 7919) % 1000) / 1000.f - 0.5f;
  fftwf_plan p = fftwf_plan_dft_r2c_1d (N, in, out, 0);
  struct timespec a, b;
  volatile float sink = 0;
  clock_gettime (CLOCK_MONOTONIC, &a);
  for (int r = 0; r < reps; r++)
    {
      fftwf_execute_dft_r2c (p, in, out);
      sink += out[5][0];
      in[3] += 1e-9f;
    }
  clock_gettime (CLOCK_MONOTONIC, &b);
  fftwf_destroy_plan (p);
  fftwf_free (in);
  fftwf_free (out);
  return ((b.tv_sec - a.tv_sec) * 1e9 + (b.tv_nsec - a.tv_nsec)) / reps / 1e3;
}

/* glibc transcendental probes so tests can pin device math against the host libm */
void orc_log2f (const float *in, float *out, int64_t n) { for (int64_t i = 0; i < n; i++) out[i] = log2f (in[i]); }
void orc_powf (const float *a, const float *b, float *out, int64_t n) { for (int64_t i = 0; i < n; i++) out[i] = powf (a[i], b[i]); }
void orc_hypotf (const float *a, const float *b, float *out, int64_t n) { for (int64_t i = 0; i < n; i++) out[i] = hypotf (a[i], b[i]); }

} // extern "C"

/* ------------------------------------------------------------------------------------------------ resampler
 * process_resampler (src/resample.cc:27-50) driving the in-repo VResampler stand-in (ref_shims/awm_vresampler.hh --
 * zita-resampler itself is third party and absent, see that header): k/2 - 1 frames of zero pre-roll, the input,
 * k/2 frames of zero post-roll; outputs that cannot be produced stay 0 like the zero-initialised vector of the caller.
 */
#include "ref_shims/awm_vresampler.hh"
extern "C" {

int
orc_resample (const float *in, int64_t n_in, int n_channels, double ratio, int hlen, float *out, int64_t n_out)
{
  AwmVResampler r;
  if (r.setup (ratio, n_channels, hlen) != 0)
    return 1;
  memset (out, 0, sizeof (float) * n_out * n_channels);
  r.out_count = n_out;
  r.out_data = out;
  r.inp_count = r.inpsize() / 2 - 1;
  r.inp_data = nullptr;
  r.process();
  r.inp_count = n_in;
  r.inp_data = const_cast<float *> (in);
  r.process();
  r.inp_count = r.inpsize() / 2;
  r.inp_data = nullptr;
  r.process();
  return 0;
}

/* ------------------------------------------------------------------------------------------------ speed detection
 * SpeedSync::prepare_mags (src/wmspeed.cc:203-268) on the already resampled half-rate clip.
 *   sub: [n_sub][n_channels]; window: gen_normalized_window (512); entries sorted by frame (SpeedSync ctor :156-160),
 *   up/down: [n_ent][n_ud] band indices (bin - min_band).  mags: [n_ent][rows][2] (column major like MagMatrix :75-79).
 * returns the number of rows.
 */
int64_t
orc_speed_rows (int64_t n_sub)
{
  int64_t rows = 0;
  for (int64_t ppos = 0; ppos + 512 < n_sub; ppos += 128)
    rows++;
  return rows;
}

void
orc_speed_mags (const float *sub, int64_t n_sub, int n_channels, const float *window, int n_ent, int n_ud,
                const int *up, const int *down, float *mags)
{
  const int sub_frame_size = 512, sub_sync_search_step = 128, min_band = 20, max_band = 100;
  const int64_t rows = orc_speed_rows (n_sub);
  Plans& p = plans_for (sub_frame_size);
#pragma omp parallel for schedule(static)
  for (int64_t row = 0; row < rows; row++)
    {
      const int64_t pos = row * sub_sync_search_step;
      float in[512 + 2], out[512 + 2];
      float fft_out_db[max_band - min_band + 1];
      for (int i = 0; i <= max_band - min_band; i++)
        fft_out_db[i] = 0;
      for (int ch = 0; ch < n_channels; ch++)
        {
          for (int i = 0; i < sub_frame_size; i++)
            in[i] = sub[ch + (pos + i) * n_channels] * window[i];
          fftwf_execute_dft_r2c (p.fwd, in, (fftwf_complex *) out);
          for (int i = min_band; i <= max_band; i++)
            fft_out_db[i - min_band] += db_from_complex (out[i * 2], out[i * 2 + 1], -96);
        }
      for (int col = 0; col < n_ent; col++)
        {
          float umag = 0, dmag = 0;
          for (int i = 0; i < n_ud; i++)
            {
              umag += fft_out_db[up[col * n_ud + i]];
              dmag += fft_out_db[down[col * n_ud + i]];
            }
          mags[(int64_t (col) * rows + row) * 2] = umag;
          mags[(int64_t (col) * rows + row) * 2 + 1] = dmag;
        }
    }
}

namespace {
struct OrcBitValue { float umag = 0, dmag = 0; int count = 0; };
struct OrcCmpState { int offset = 0; OrcBitValue bit_values[8]; };

/* SpeedSync::compare_bits<BLOCK> (src/wmspeed.cc:270-326) */
void
orc_compare_bits (int BLOCK, std::vector<OrcCmpState>& cmp_states, double relative_speed, const float *mags, int64_t rows,
                  int n_ent, const int *frame, const int *bit, int frames_per_block)
{
  const int steps_per_frame = 4, OFFSET_SHIFT = 16;
  const double relative_speed_inv = 1 / relative_speed;
  size_t begin = cmp_states.size(), end = cmp_states.size();
  for (int mi = 0; mi < n_ent; mi++)
    {
      const int frame_offset = ((BLOCK * frames_per_block + frame[mi]) * steps_per_frame * relative_speed_inv + 0.5) * (1 << OFFSET_SHIFT);
      while (begin > 0)
        {
          const int index = cmp_states[begin - 1].offset + frame_offset;
          if (index < 0)
            break;
          begin--;
        }
      while (end > 0)
        {
          const int index = (cmp_states[end - 1].offset + frame_offset) >> OFFSET_SHIFT;
          if (index < rows)
            break;
          end--;
        }
      for (size_t it = begin; it < end; it++)
        {
          const int index = (cmp_states[it].offset + frame_offset) >> OFFSET_SHIFT;
          OrcBitValue& bv = cmp_states[it].bit_values[bit[mi]];
          const float mu = mags[(int64_t (mi) * rows + index) * 2], md = mags[(int64_t (mi) * rows + index) * 2 + 1];
          if (BLOCK & 1)
            {
              bv.umag += md;
              bv.dmag += mu;
            }
          else
            {
              bv.umag += mu;
              bv.dmag += md;
            }
          bv.count++;
        }
    }
}
}

/* SpeedSync::compare (src/wmspeed.cc:328-375) for n_rel relative speeds on one MagMatrix; quality_out[r] = best_score.quality */
void
orc_speed_compare (const float *mags, int64_t rows, int n_ent, const int *frame, const int *bit, int n_sync_bits, int frames_per_block,
                   const double *relative_speeds, int n_rel, double water_delta, double *quality_out)
{
#pragma omp parallel for schedule(dynamic)
  for (int r = 0; r < n_rel; r++)
    {
      const double relative_speed = relative_speeds[r];
      const int steps_per_frame = 4;
      const int pad_start = frames_per_block * steps_per_frame + steps_per_frame;
      std::vector<OrcCmpState> cmp_states;
      for (int offset = -pad_start; offset < 0; offset++)
        {
          OrcCmpState cs;
          cs.offset = offset * ((1 << 16) / relative_speed);
          cmp_states.push_back (cs);
        }
      orc_compare_bits (0, cmp_states, relative_speed, mags, rows, n_ent, frame, bit, frames_per_block);
      orc_compare_bits (1, cmp_states, relative_speed, mags, rows, n_ent, frame, bit, frames_per_block);
      orc_compare_bits (2, cmp_states, relative_speed, mags, rows, n_ent, frame, bit, frames_per_block);
      double best_quality = 0;
      for (const auto& cs : cmp_states)
        {
          double sync_quality = 0;
          int bit_count = 0;
          for (int b = 0; b < n_sync_bits; b++)
            {
              const OrcBitValue& bv = cs.bit_values[b];
              /* SyncFinder::bit_quality (src/syncfinder.cc:94-114) */
              double raw_bit;
              if (bv.umag == 0 || bv.dmag == 0)
                raw_bit = 0;
              else if (bv.umag < bv.dmag)
                raw_bit = 1 - bv.umag / bv.dmag;
              else
                raw_bit = bv.dmag / bv.umag - 1;
              sync_quality += ((b & 1) ? raw_bit : -raw_bit) * bv.count;
              bit_count += bv.count;
            }
          if (bit_count)
            {
              sync_quality /= bit_count;
              const double wd = water_delta < 0.080 ? water_delta : 0.080;
              sync_quality = fabs (sync_quality / wd / 2.9);
              if (sync_quality > best_quality)
                best_quality = sync_quality;
            }
        }
      quality_out[r] = best_quality;
    }
}

/* score_smooth_find_best (src/wmspeed.cc:385-419); scores sorted by speed by the caller */
double
orc_score_smooth_find_best (const double *speeds, const double *qualities, int n, double step, double distance)
{
  double best_speed = 0, best_quality = 0;
  for (double speed = speeds[0]; speed < speeds[n - 1]; speed += 0.000001)
    {
      double quality_sum = 0, quality_div = 0;
      for (int i = 0; i < n; i++)
        {
          const double x = (speeds[i] - speed) / (step * distance);
          const double w = fabs (x) > 1 ? 0 : 0.5 * cos (x * M_PI) + 0.5;
          quality_sum += qualities[i] * w;
          quality_div += w;
        }
      quality_sum /= quality_div;
      if (quality_sum > best_quality)
        {
          best_speed = speed;
          best_quality = quality_sum;
        }
    }
  return best_speed;
}

} // extern "C"
