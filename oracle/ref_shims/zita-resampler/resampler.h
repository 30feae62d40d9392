/* Stand-in for zita-resampler's Resampler (library absent, out of scope):
 * setup() always fails so the reference reports "resampling not available"
 * (call sites src/resample.cc:80-92,233-245). */
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
