/* Stand-in for zita-resampler's VResampler: interface of the library (absent here), algorithm of
 * ../awm_vresampler.hh (in-repo, NOT zita's coefficients -> parity downstream of a resampler is unpinned). */
#ifndef AWM_REF_SHIM_ZITA_VRESAMPLER_H
#define AWM_REF_SHIM_ZITA_VRESAMPLER_H
#include "../awm_vresampler.hh"
class VResampler : public AwmVResampler
{
};
#endif
