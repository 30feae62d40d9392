/* Stand-in for zita-resampler's VResampler (see resampler.h). */
#ifndef AWM_REF_SHIM_ZITA_VRESAMPLER_H
#define AWM_REF_SHIM_ZITA_VRESAMPLER_H
class VResampler
{
public:
  unsigned int inp_count = 0, out_count = 0;
  float       *inp_data = nullptr, *out_data = nullptr;
  int  setup (double, unsigned int, unsigned int) { return 1; }
  int  nchan() const { return 1; }
  int  inpsize() const { return 2; }
  int  process() { return 1; }
};
#endif
