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

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {

struct Stage { int n, s, radix; std::vector<float> wr[3], wi[3]; };

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
  /* forward transform; result ends up in (xr,xi) or (yr,yi): returns true if in y */
  bool
  run (float *xr, float *xi, float *yr, float *yi) const
  {
    bool in_y = false;
    for (const Stage& st : stages)
      {
        const int s = st.s;
        if (st.radix == 4)
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

extern "C" {

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
  std::vector<float> buf (4 * M);
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
  /* X[k] = (Z[k] + conj Z[M-k]) / 2  -  (i/2) W^k (Z[k] - conj Z[M-k]) */
  for (int k = 0; k <= M; k++)
    {
      const int k1 = k % M, k2 = (M - k) % M;
      const float er = 0.5f * (xr[k1] + xr[k2]), ei = 0.5f * (xi[k1] - xi[k2]);
      const float orr = 0.5f * (xr[k1] - xr[k2]), oi = 0.5f * (xi[k1] + xi[k2]);
      /* -i * (orr + i oi) = oi - i orr */
      const float tr = oi, ti = -orr;
      const float wr = p->tw_r[k], wi = p->tw_i[k];
      out[k][0] = er + (tr * wr - ti * wi);
      out[k][1] = ei + (tr * wi + ti * wr);
    }
}

void
fftwf_execute_dft_c2r (const fftwf_plan p, fftwf_complex *in, float *out)
{
  const int N = p->N, M = N / 2;
  std::vector<float> buf (4 * M);
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
