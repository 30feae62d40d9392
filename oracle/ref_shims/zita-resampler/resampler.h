/* Stand-in for zita-resampler's fixed-ratio Resampler: setup() always fails, which makes the reference fall back
 * to VResampler (src/resample.cc:80-92,233-245) -- one in-repo algorithm serves both. */
#ifndef AWM_REF_SHIM_ZITA_RESAMPLER_H
#define AWM_REF_SHIM_ZITA_RESAMPLER_H
class Resampler
{
public:
  unsigned int inp_count = 0, out_count = 0;
  float       *inp_data = nullptr, *out_data = nullptr;
  int  setup (unsigned int, unsigned int, unsigned int, unsigned int) { return 1; }
  int  nchan() const { return 1; }
  int  inpsize() const { return 2; }
  int  process() { return 1; }
};
#endif
