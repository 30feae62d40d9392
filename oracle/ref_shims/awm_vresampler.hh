/* awm_vresampler.hh -- in-repo variable-ratio resampler behind zita-resampler's VResampler interface.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/ref_shims).  zita-resampler is a third-party library the reference links
 * (configure.ac:35-41; call sites src/resample.cc:30-50,80-118,150-201,249-259) and that is neither vendored nor
 * installed here, so parity for everything downstream of a resampler is UNPINNED against a stock build.  This stand-in
 * keeps the interface and the conventions the reference relies on (windowed-sinc polyphase filter of half length `hlen`
 * at the lower of the two rates, k = inpsize() input samples per output, k/2 - 1 samples of pre-roll, k/2 of post-roll,
 * inp_/out_ count + data members, process() consumes / produces as far as it can), not zita's coefficients.
 *
 *   y(t) = sum_i x[i] g(i - t),  g(d) = fc sinc(fc d) w(d / h),  fc = min(1, ratio),  h = hlen / fc
 *   256 phases, coefficients linearly interpolated between neighbouring phases, float accumulation in tap order.
 */
#ifndef AWM_REF_SHIM_VRESAMPLER_HH
#define AWM_REF_SHIM_VRESAMPLER_HH

#include <math.h>
#include <string.h>
#include <vector>

class AwmVResampler
{
public:
  static constexpr int NP = 256;                  /* phases */
  unsigned int inp_count = 0, out_count = 0;
  float       *inp_data = nullptr, *out_data = nullptr;

  int
  setup (double ratio, unsigned int nchan, unsigned int hlen)
  {
    if (!(ratio > 1.0 / 64 && ratio < 64) || nchan < 1 || hlen < 8 || hlen > 96)
      return 1;
    m_ratio = ratio;
    m_nchan = nchan;
    const double fc = ratio < 1 ? ratio : 1;
    m_h = int (ceil (hlen / fc));
    m_step = 1.0 / ratio;
    /* coefficient table: phase p (fraction p / NP), tap j <-> offset d = (j - (h - 1)) - p / NP, j = 0 .. 2h - 1 */
    m_coef.assign (size_t (NP + 1) * 2 * m_h, 0.f);
    for (int p = 0; p <= NP; p++)
      for (int j = 0; j < 2 * m_h; j++)
        {
          const double d = (j - (m_h - 1)) - double (p) / NP;
          m_coef[size_t (p) * 2 * m_h + j] = float (fc * sinc (fc * d) * wind (d / m_h));
        }
    reset();
    return 0;
  }
  void
  reset()
  {
    m_buf.clear();
    m_base = 0;
    m_nout = 0;
    m_t = m_h - 1;          /* centre of the first output: the sample after k/2 - 1 frames of pre-roll */
    inp_count = out_count = 0;
    inp_data = out_data = nullptr;
  }
  int nchan() const   { return m_nchan; }
  int inpsize() const { return 2 * m_h; }
  double ratio() const { return m_ratio; }

  /* consume up to inp_count frames (inp_data == nullptr: zeros), produce up to out_count frames (out_data == nullptr: discard) */
  int
  process()
  {
    for (;;)
      {
        /* produce while the taps of the next output are available */
        while (out_count)
          {
            const double fl = floor (m_t);
            const long long c = (long long) fl;               /* taps c - h + 1 .. c + h */
            if (c + m_h >= m_base + (long long) (m_buf.size() / m_nchan))
              break;
            const double frac = (m_t - fl) * NP;
            const int p = int (frac);
            const float a = float (frac - p), b = 1.0f - a;
            const float *c0 = &m_coef[size_t (p) * 2 * m_h], *c1 = c0 + 2 * m_h;
            const float *x = &m_buf[size_t (c - m_h + 1 - m_base) * m_nchan];
            for (int ch = 0; ch < m_nchan; ch++)
              {
                float s = 0;
                for (int j = 0; j < 2 * m_h; j++)
                  s += x[j * m_nchan + ch] * (b * c0[j] + a * c1[j]);
                if (out_data)
                  out_data[ch] = s;
              }
            if (out_data)
              out_data += m_nchan;
            out_count--;
            m_nout++;
            m_t = (m_h - 1) + double (m_nout) * m_step;      /* closed form: no drift, reproducible by a parallel implementation */
          }
        if (!out_count || !inp_count)
          break;
        /* take just the input the next output needs (a stream that is consumed frame by frame delivers every output as
         * early as possible), drop history that is no longer needed */
        const long long need_from = (long long) floor (m_t) - m_h + 1;
        if (need_from > m_base + 4096)
          {
            const size_t drop = size_t (need_from - m_base);
            m_buf.erase (m_buf.begin(), m_buf.begin() + drop * m_nchan);
            m_base += drop;
          }
        const long long have = m_base + (long long) (m_buf.size() / m_nchan);
        const long long missing = (long long) floor (m_t) + m_h + 1 - have;        /* frames up to tap c + h */
        const unsigned int want = missing > 1 ? (unsigned int) (missing < 4096 ? missing : 4096) : 1;
        const unsigned int take = inp_count < want ? inp_count : want;
        const size_t old = m_buf.size();
        m_buf.resize (old + size_t (take) * m_nchan);
        if (inp_data)
          {
            memcpy (&m_buf[old], inp_data, size_t (take) * m_nchan * sizeof (float));
            inp_data += size_t (take) * m_nchan;
          }
        else
          memset (&m_buf[old], 0, size_t (take) * m_nchan * sizeof (float));
        inp_count -= take;
      }
    return 0;
  }
  static double
  sinc (double x)
  {
    x = fabs (x);
    if (x < 1e-9)
      return 1;
    x *= M_PI;
    return sin (x) / x;
  }
  static double
  wind (double x)     /* three-term cosine window on [-1, 1] */
  {
    x = fabs (x);
    if (x >= 1)
      return 0;
    x *= M_PI;
    return 0.384 + 0.5 * cos (x) + 0.116 * cos (2 * x);
  }
  const std::vector<float>& coef() const { return m_coef; }
  int half_length() const { return m_h; }
private:
  double m_ratio = 1, m_step = 1, m_t = 0;
  int    m_nchan = 1, m_h = 16;
  long long m_base = 0;                       /* stream index of m_buf[0] */
  long long m_nout = 0;                       /* outputs produced so far */
  std::vector<float> m_coef, m_buf;
};

#endif
