/* fftw_shim.cc -- in-repo single precision FFT behind the seven FFTW entry
 * points the reference uses (src/fft.cc:57-91).
 *
 * TEST INFRASTRUCTURE ONLY.  This file is linked into oracle/_ref/audiowmark
 * (the unmodified reference sources built by oracle/Makefile.ref) because
 * FFTW3f is not installed in this image.  It is not FFTW: it is a Stockham
 * autosort radix-4/2 complex FFT (split re/im arrays so gcc vectorises the
 * inner loops) plus the usual N/2-point packing for real transforms.
 *
 * Semantics reproduced (FFTW manual, "One-Dimensional DFTs of Real Data"):
 *   r2c: out[k] = sum_n in[n] exp(-2 pi i k n / N), k = 0..N/2   (N/2+1 complex)
 *   c2r: out[n] = sum_k Hermitian-extended in[k] exp(+2 pi i k n / N)  (unnormalised)
 *   plans are shared between threads by the reference (src/fft.cc:29-49), so
 *   execute() keeps all scratch on the caller's stack.
 */
#include "fftw3.h"

#include <immintrin.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

struct Stage { int n, s, radix; std::vector<float> wr[3], wi[3]; };

/* The first two passes of the Stockham transform have inner loops of length s = 1 and s = 4, too short for the compiler to
 * vectorise.  These AVX2 versions run the butterflies of 8 (s = 1) or 2 (s = 4) values of p side by side: per output element the
 * SAME sequence of IEEE single precision operations as the scalar loops in CPlan::run (separate multiply and add/sub, no FMA),
 * so the results are bit identical -- checked by tests/test_oracle_golden.py::test_fft_shim_simd_bit_identical. */
struct R4 { __m256 o0r, o0i, o1r, o1i, o2r, o2i, o3r, o3i; };

__attribute__ ((target ("avx2"))) static inline R4
radix4_avx2 (__m256 ar, __m256 ai, __m256 br, __m256 bi, __m256 cr, __m256 ci, __m256 dr, __m256 di,
             __m256 w1r, __m256 w1i, __m256 w2r, __m256 w2i, __m256 w3r, __m256 w3i)
{
  const __m256 sign = _mm256_set1_ps (-0.0f);
  const __m256 apcr = _mm256_add_ps (ar, cr), apci = _mm256_add_ps (ai, ci);
  const __m256 amcr = _mm256_sub_ps (ar, cr), amci = _mm256_sub_ps (ai, ci);
  const __m256 bpdr = _mm256_add_ps (br, dr), bpdi = _mm256_add_ps (bi, di);
  const __m256 jr = _mm256_xor_ps (_mm256_sub_ps (bi, di), sign), ji = _mm256_sub_ps (br, dr);      /* j * (b - d) */
  R4 o;
  o.o0r = _mm256_add_ps (apcr, bpdr);
  o.o0i = _mm256_add_ps (apci, bpdi);
  const __m256 t1r = _mm256_sub_ps (amcr, jr), t1i = _mm256_sub_ps (amci, ji);
  o.o1r = _mm256_sub_ps (_mm256_mul_ps (t1r, w1r), _mm256_mul_ps (t1i, w1i));
  o.o1i = _mm256_add_ps (_mm256_mul_ps (t1r, w1i), _mm256_mul_ps (t1i, w1r));
  const __m256 t2r = _mm256_sub_ps (apcr, bpdr), t2i = _mm256_sub_ps (apci, bpdi);
  o.o2r = _mm256_sub_ps (_mm256_mul_ps (t2r, w2r), _mm256_mul_ps (t2i, w2i));
  o.o2i = _mm256_add_ps (_mm256_mul_ps (t2r, w2i), _mm256_mul_ps (t2i, w2r));
  const __m256 t3r = _mm256_add_ps (amcr, jr), t3i = _mm256_add_ps (amci, ji);
  o.o3r = _mm256_sub_ps (_mm256_mul_ps (t3r, w3r), _mm256_mul_ps (t3i, w3i));
  o.o3i = _mm256_add_ps (_mm256_mul_ps (t3r, w3i), _mm256_mul_ps (t3i, w3r));
  return o;
}

/* y[4 p + t] = o_t[p] for 8 consecutive p: 4 x 8 transpose */
__attribute__ ((target ("avx2"))) static inline void
store_interleaved4 (float *y, __m256 o0, __m256 o1, __m256 o2, __m256 o3)
{
  const __m256 t0 = _mm256_unpacklo_ps (o0, o1), t1 = _mm256_unpackhi_ps (o0, o1);
  const __m256 t2 = _mm256_unpacklo_ps (o2, o3), t3 = _mm256_unpackhi_ps (o2, o3);
  const __m256 u0 = _mm256_shuffle_ps (t0, t2, 0x44), u1 = _mm256_shuffle_ps (t0, t2, 0xee);
  const __m256 u2 = _mm256_shuffle_ps (t1, t3, 0x44), u3 = _mm256_shuffle_ps (t1, t3, 0xee);
  _mm256_storeu_ps (y, _mm256_permute2f128_ps (u0, u1, 0x20));
  _mm256_storeu_ps (y + 8, _mm256_permute2f128_ps (u2, u3, 0x20));
  _mm256_storeu_ps (y + 16, _mm256_permute2f128_ps (u0, u1, 0x31));
  _mm256_storeu_ps (y + 24, _mm256_permute2f128_ps (u2, u3, 0x31));
}

/* radix-4 pass with s = 1 (m = n / 4 butterflies, m a multiple of 8) */
__attribute__ ((target ("avx2"))) static void
pass_r4_s1_avx2 (const float *xr, const float *xi, float *yr, float *yi, int m, const float *const *wr, const float *const *wi)
{
  for (int p = 0; p < m; p += 8)
    {
      const R4 o = radix4_avx2 (_mm256_loadu_ps (xr + p), _mm256_loadu_ps (xi + p), _mm256_loadu_ps (xr + p + m), _mm256_loadu_ps (xi + p + m),
                                _mm256_loadu_ps (xr + p + 2 * m), _mm256_loadu_ps (xi + p + 2 * m), _mm256_loadu_ps (xr + p + 3 * m), _mm256_loadu_ps (xi + p + 3 * m),
                                _mm256_loadu_ps (wr[0] + p), _mm256_loadu_ps (wi[0] + p), _mm256_loadu_ps (wr[1] + p), _mm256_loadu_ps (wi[1] + p),
                                _mm256_loadu_ps (wr[2] + p), _mm256_loadu_ps (wi[2] + p));
      store_interleaved4 (yr + 4 * p, o.o0r, o.o1r, o.o2r, o.o3r);
      store_interleaved4 (yi + 4 * p, o.o0i, o.o1i, o.o2i, o.o3i);
    }
}

__attribute__ ((target ("avx2"))) static inline __m256
dup2 (const float *w, int p)
{
  return _mm256_set_m128 (_mm_set1_ps (w[p + 1]), _mm_set1_ps (w[p]));
}

__attribute__ ((target ("avx2"))) static inline void
put (float *y, __m256 v)
{
  _mm_storeu_ps (y, _mm256_castps256_ps128 (v));
  _mm_storeu_ps (y + 16, _mm256_extractf128_ps (v, 1));
}

/* radix-4 pass with s = 4 (m even): lanes = [p: q 0..3 | p + 1: q 0..3] */
__attribute__ ((target ("avx2"))) static void
pass_r4_s4_avx2 (const float *xr, const float *xi, float *yr, float *yi, int m, const float *const *wr, const float *const *wi)
{
  for (int p = 0; p < m; p += 2)
    {
      const float *ar = xr + 4 * p, *ai = xi + 4 * p;
      const R4 o = radix4_avx2 (_mm256_loadu_ps (ar), _mm256_loadu_ps (ai), _mm256_loadu_ps (ar + 4 * m), _mm256_loadu_ps (ai + 4 * m),
                                _mm256_loadu_ps (ar + 8 * m), _mm256_loadu_ps (ai + 8 * m), _mm256_loadu_ps (ar + 12 * m), _mm256_loadu_ps (ai + 12 * m),
                                dup2 (wr[0], p), dup2 (wi[0], p), dup2 (wr[1], p), dup2 (wi[1], p), dup2 (wr[2], p), dup2 (wi[2], p));
      float *o_r = yr + 16 * p, *o_i = yi + 16 * p;
      put (o_r, o.o0r); put (o_r + 4, o.o1r); put (o_r + 8, o.o2r); put (o_r + 12, o.o3r);
      put (o_i, o.o0i); put (o_i + 4, o.o1i); put (o_i + 8, o.o2i); put (o_i + 12, o.o3i);
    }
}

static const bool have_avx2 = __builtin_cpu_supports ("avx2");
bool awm_shim_use_simd = true;          /* tests switch the explicit AVX2 passes off to compare against the plain loops */

struct CPlan              /* complex FFT of size M (power of two) */
{
  int M = 0;
  std::vector<Stage> stages;

  explicit CPlan (int m) : M (m)
  {
    int n = M, s = 1;
    while (n > 1)
      {
        Stage st;
        st.n = n; st.s = s;
        st.radix = (n % 4 == 0) ? 4 : 2;
        const int m4 = n / st.radix;
        for (int t = 0; t < st.radix - 1; t++)
          {
            st.wr[t].resize (m4); st.wi[t].resize (m4);
            for (int p = 0; p < m4; p++)
              {
                const double a = -2.0 * M_PI * double (p) * (t + 1) / n;
                st.wr[t][p] = float (cos (a));
                st.wi[t][p] = float (sin (a));
              }
          }
        stages.push_back (st);
        n /= st.radix; s *= st.radix;
      }
  }
  /* forward transform; result ends up in (xr,xi) or (yr,yi): returns true if in y.
   * Built twice (function multiversioning, picked at load time): baseline x86-64 and AVX2.  The AVX2 clone runs the same IEEE
   * operations per element on wider vectors -- it has no FMA (the clone enables avx2 only), so nothing is contracted and the
   * results are bit identical on every host; the golden hashes do not depend on the CPU. */
  __attribute__ ((target_clones ("avx2", "default"))) bool
  run (float *xr, float *xi, float *yr, float *yi) const
  {
    bool in_y = false;
    for (const Stage& st : stages)
      {
        const int s = st.s;
        if (st.radix == 4 && have_avx2 && awm_shim_use_simd && (s == 1 || s == 4) && st.n % 32 == 0)
          {
            const float *wr[3] = { st.wr[0].data(), st.wr[1].data(), st.wr[2].data() }, *wi[3] = { st.wi[0].data(), st.wi[1].data(), st.wi[2].data() };
            if (s == 1)
              pass_r4_s1_avx2 (xr, xi, yr, yi, st.n / 4, wr, wi);
            else
              pass_r4_s4_avx2 (xr, xi, yr, yi, st.n / 4, wr, wi);
          }
        else if (st.radix == 4)
          {
            const int m = st.n / 4;
            for (int p = 0; p < m; p++)
              {
                const float w1r = st.wr[0][p], w1i = st.wi[0][p];
                const float w2r = st.wr[1][p], w2i = st.wi[1][p];
                const float w3r = st.wr[2][p], w3i = st.wi[2][p];
                const float *ar = xr + s * p, *ai = xi + s * p;
                const float *br = ar + s * m, *bi = ai + s * m;
                const float *cr = br + s * m, *ci = bi + s * m;
                const float *dr = cr + s * m, *di = ci + s * m;
                float *o0r = yr + s * 4 * p, *o0i = yi + s * 4 * p;
                float *o1r = o0r + s, *o1i = o0i + s;
                float *o2r = o1r + s, *o2i = o1i + s;
                float *o3r = o2r + s, *o3i = o2i + s;
                for (int q = 0; q < s; q++)
                  {
                    const float apcr = ar[q] + cr[q], apci = ai[q] + ci[q];
                    const float amcr = ar[q] - cr[q], amci = ai[q] - ci[q];
                    const float bpdr = br[q] + dr[q], bpdi = bi[q] + di[q];
                    /* j * (b - d) */
                    const float jr = -(bi[q] - di[q]), ji = br[q] - dr[q];
                    o0r[q] = apcr + bpdr;
                    o0i[q] = apci + bpdi;
                    const float t1r = amcr - jr, t1i = amci - ji;
                    o1r[q] = t1r * w1r - t1i * w1i;
                    o1i[q] = t1r * w1i + t1i * w1r;
                    const float t2r = apcr - bpdr, t2i = apci - bpdi;
                    o2r[q] = t2r * w2r - t2i * w2i;
                    o2i[q] = t2r * w2i + t2i * w2r;
                    const float t3r = amcr + jr, t3i = amci + ji;
                    o3r[q] = t3r * w3r - t3i * w3i;
                    o3i[q] = t3r * w3i + t3i * w3r;
                  }
              }
          }
        else
          {
            const int m = st.n / 2;
            for (int p = 0; p < m; p++)
              {
                const float wr = st.wr[0][p], wi = st.wi[0][p];
                const float *ar = xr + s * p, *ai = xi + s * p;
                const float *br = ar + s * m, *bi = ai + s * m;
                float *o0r = yr + s * 2 * p, *o0i = yi + s * 2 * p;
                float *o1r = o0r + s, *o1i = o0i + s;
                for (int q = 0; q < s; q++)
                  {
                    o0r[q] = ar[q] + br[q];
                    o0i[q] = ai[q] + bi[q];
                    const float tr = ar[q] - br[q], ti = ai[q] - bi[q];
                    o1r[q] = tr * wr - ti * wi;
                    o1i[q] = tr * wi + ti * wr;
                  }
              }
          }
        float *t;
        t = xr; xr = yr; yr = t;
        t = xi; xi = yi; yi = t;
        in_y = !in_y;
      }
    return in_y;
  }
};

} // namespace

struct awm_shim_plan
{
  int   N = 0;        /* real transform size */
  bool  inverse = false;
  CPlan cplan;
  std::vector<float> tw_r, tw_i;  /* exp(-2 pi i k / N), k = 0..N/2 */

  awm_shim_plan (int n, bool inv) : N (n), inverse (inv), cplan (n / 2)
  {
    tw_r.resize (N / 2 + 1); tw_i.resize (N / 2 + 1);
    for (int k = 0; k <= N / 2; k++)
      {
        const double a = -2.0 * M_PI * k / N;
        tw_r[k] = float (cos (a));
        tw_i[k] = float (sin (a));
      }
  }
};

static const int MAX_N = 1 << 16;

/* one output bin of the real-input post-pass; k1 = k mod M, k2 = (M - k) mod M */
static inline void
unpack_bin (const float *xr, const float *xi, int k1, int k2, float wr, float wi, float *o)
{
  const float er = 0.5f * (xr[k1] + xr[k2]), ei = 0.5f * (xi[k1] - xi[k2]);
  const float orr = 0.5f * (xr[k1] - xr[k2]), oi = 0.5f * (xi[k1] + xi[k2]);
  /* -i * (orr + i oi) = oi - i orr */
  const float tr = oi, ti = -orr;
  o[0] = er + (tr * wr - ti * wi);
  o[1] = ei + (tr * wi + ti * wr);
}

__attribute__ ((target_clones ("avx2", "default"))) static void
unpack_r2c (const float *xr, const float *xi, const float *tw_r, const float *tw_i, int M, float *out)
{
  unpack_bin (xr, xi, 0, 0, tw_r[0], tw_i[0], out);
  for (int k = 1; k < M; k++)                         /* no index arithmetic modulo M inside the loop */
    unpack_bin (xr, xi, k, M - k, tw_r[k], tw_i[k], out + 2 * k);
  unpack_bin (xr, xi, 0, 0, tw_r[M], tw_i[M], out + 2 * M);
}

extern "C" {

/* test hook: 0 = plain loops only, 1 = explicit AVX2 passes where the CPU has them (default) */
void
awm_shim_set_simd (int on)
{
  awm_shim_use_simd = on != 0;
}

void *
fftwf_malloc (size_t n)
{
  void *p = nullptr;
  if (posix_memalign (&p, 64, n ? n : 64))
    return nullptr;
  return p;
}

void
fftwf_free (void *p)
{
  free (p);
}

fftwf_plan
fftwf_plan_dft_r2c_1d (int n, float *, fftwf_complex *, unsigned)
{
  if (n < 4 || n > MAX_N || (n & (n - 1)))
    return nullptr;
  return new awm_shim_plan (n, false);
}

fftwf_plan
fftwf_plan_dft_c2r_1d (int n, fftwf_complex *, float *, unsigned)
{
  if (n < 4 || n > MAX_N || (n & (n - 1)))
    return nullptr;
  return new awm_shim_plan (n, true);
}

void
fftwf_destroy_plan (fftwf_plan p)
{
  delete p;
}

void
fftwf_execute_dft_r2c (const fftwf_plan p, float *in, fftwf_complex *out)
{
  const int N = p->N, M = N / 2;
  static thread_local std::vector<float> buf;           /* scratch per calling thread: plans are shared between threads */
  if (buf.size() < size_t (4 * M))
    buf.resize (4 * M);
  float *xr = &buf[0], *xi = &buf[M], *yr = &buf[2 * M], *yi = &buf[3 * M];
  for (int n = 0; n < M; n++)
    {
      xr[n] = in[2 * n];
      xi[n] = in[2 * n + 1];
    }
  if (p->cplan.run (xr, xi, yr, yi))
    {
      xr = yr; xi = yi;
    }
  /* X[k] = (Z[k] + conj Z[M-k]) / 2  -  (i/2) W^k (Z[k] - conj Z[M-k]);  Z[M] = Z[0] (k = 0 and k = M both pair Z[0] with itself) */
  unpack_r2c (xr, xi, &p->tw_r[0], &p->tw_i[0], M, &out[0][0]);
}

void
fftwf_execute_dft_c2r (const fftwf_plan p, fftwf_complex *in, float *out)
{
  const int N = p->N, M = N / 2;
  static thread_local std::vector<float> buf;           /* scratch per calling thread: plans are shared between threads */
  if (buf.size() < size_t (4 * M))
    buf.resize (4 * M);
  float *xr = &buf[0], *xi = &buf[M], *yr = &buf[2 * M], *yi = &buf[3 * M];
  /* Z[k] = (X[k] + conj X[M-k]) + i conj(W^k) (X[k] - conj X[M-k]);  x[2n] + i x[2n+1] = IDFT_M (Z) */
  for (int k = 0; k < M; k++)
    {
      const float ar = in[k][0], ai = in[k][1];
      const float br = in[M - k][0], bi = -in[M - k][1];
      const float er = ar + br, ei = ai + bi;
      const float dr = ar - br, di = ai - bi;
      const float wr = p->tw_r[k], wi = -p->tw_i[k];
      const float tr = dr * wr - di * wi, ti = dr * wi + di * wr;
      /* i * (tr + i ti) = -ti + i tr ; store swapped (re<->im) so a forward FFT computes the inverse */
      const float zr = er - ti, zi = ei + tr;
      xr[k] = zi;
      xi[k] = zr;
    }
  if (p->cplan.run (xr, xi, yr, yi))
    {
      xr = yr; xi = yi;
    }
  for (int n = 0; n < M; n++)
    {
      out[2 * n]     = xi[n];
      out[2 * n + 1] = xr[n];
    }
}

} // extern "C"
