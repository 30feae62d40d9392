// awm_f32x2.cuh -- packed fp32 arithmetic of sm_100a (PTX add / mul / fma .rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2): one instruction
// produces two IEEE round-to-nearest results, each bit for bit what the scalar __fadd_rn / __fmul_rn / __fmaf_rn gives.  The kernels
// that are bound by fp32 issue (Viterbi add-compare-select, FFT butterflies, sliding DFT) use it to halve their instruction count
// without touching the arithmetic the parity tests pin.  tools/ffma2_probe.cu measures the issue rates.
#pragma once
#include <cuda_runtime.h>

namespace awm {

struct f2 { unsigned long long v; };          // two floats in one aligned 64-bit register pair: lo = x, hi = y

__device__ __forceinline__ f2
f2_make (float x, float y)
{
  f2 r;
  asm ("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void
f2_split (f2 a, float& x, float& y)
{
  asm ("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
}
__device__ __forceinline__ float f2_lo (f2 a) { float x, y; f2_split (a, x, y); return x; }
__device__ __forceinline__ float f2_hi (f2 a) { float x, y; f2_split (a, x, y); return y; }
__device__ __forceinline__ f2
f2_add (f2 a, f2 b)
{
  f2 r;
  asm ("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2
f2_sub (f2 a, f2 b)
{
  f2 r;
  asm ("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2
f2_mul (f2 a, f2 b)
{
  f2 r;
  asm ("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2
f2_fma (f2 a, f2 b, f2 c)
{
  f2 r;
  asm ("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}

} // namespace awm
