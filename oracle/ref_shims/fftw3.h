/* Minimal stand-in for <fftw3.h>: exactly the seven entry points the
 * reference calls (src/fft.cc:57-91).  Implemented by fftw_shim.cc with an
 * in-repo single-precision FFT; NOT FFTW.  Test infrastructure only. */
#ifndef AWM_REF_SHIM_FFTW3_H
#define AWM_REF_SHIM_FFTW3_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef float fftwf_complex[2];
typedef struct awm_shim_plan *fftwf_plan;
#define FFTW_ESTIMATE       (1U << 6)
#define FFTW_PRESERVE_INPUT (1U << 4)
void      *fftwf_malloc (size_t n);
void       fftwf_free (void *p);
fftwf_plan fftwf_plan_dft_r2c_1d (int n, float *in, fftwf_complex *out, unsigned flags);
fftwf_plan fftwf_plan_dft_c2r_1d (int n, fftwf_complex *in, float *out, unsigned flags);
void       fftwf_execute_dft_r2c (const fftwf_plan p, float *in, fftwf_complex *out);
void       fftwf_execute_dft_c2r (const fftwf_plan p, fftwf_complex *in, float *out);
void       fftwf_destroy_plan (fftwf_plan p);
#ifdef __cplusplus
}
#endif
#endif
