// tools/ffma2_probe.cu -- issue rate of scalar vs packed fp32 arithmetic on sm_100a (FFMA vs FFMA2, FADD vs FADD2, FMUL vs FMUL2).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/ffma2_probe tools/ffma2_probe.cu ; run on the GPU box.
// Prints fp32 operations per clock per SM for every variant: decides whether the FFT butterflies are worth restating on float2 pairs.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pk (float2 a) { return *reinterpret_cast<unsigned long long *> (&a); }
__device__ __forceinline__ float2 upk (unsigned long long a) { return *reinterpret_cast<float2 *> (&a); }
__device__ __forceinline__ float2 ffma2 (float2 a, float2 b, float2 c)
{ unsigned long long d; asm volatile ("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk (a)), "l"(pk (b)), "l"(pk (c))); return upk (d); }
__device__ __forceinline__ float2 fadd2 (float2 a, float2 b)
{ unsigned long long d; asm volatile ("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk (a)), "l"(pk (b))); return upk (d); }
__device__ __forceinline__ float2 fmul2 (float2 a, float2 b)
{ unsigned long long d; asm volatile ("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk (a)), "l"(pk (b))); return upk (d); }

constexpr int ILP = 8, ITER = 4096;

template<int MODE> __global__ void
k (float *out, float s0, float s1)
{
  float2 a[ILP];
  for (int i = 0; i < ILP; i++)
    a[i] = make_float2 (threadIdx.x + i, threadIdx.x - i);
  const float2 b = make_float2 (s0, s0), c = make_float2 (s1, s1);
  for (int it = 0; it < ITER; it++)
    {
#pragma unroll
      for (int i = 0; i < ILP; i++)
        {
          if (MODE == 0) { a[i].x = __fmaf_rn (a[i].x, s0, s1); a[i].y = __fmaf_rn (a[i].y, s0, s1); }        // 2 FFMA
          if (MODE == 1) a[i] = ffma2 (a[i], b, c);                                                                // 1 FFMA2
          if (MODE == 2) { a[i].x = __fadd_rn (a[i].x, s1); a[i].y = __fadd_rn (a[i].y, s1); }
          if (MODE == 3) a[i] = fadd2 (a[i], c);
          if (MODE == 4) { a[i].x = __fmul_rn (a[i].x, s0); a[i].y = __fmul_rn (a[i].y, s0); }
          if (MODE == 5) a[i] = fmul2 (a[i], b);
          if (MODE == 6) { a[i] = ffma2 (a[i], b, c); a[i].x = __shfl_xor_sync (0xffffffffu, a[i].x, 1); }      // FFMA2 + SHFL
          if (MODE == 7) { a[i].x = __fmaf_rn (a[i].x, s0, s1); a[i].y = __fmaf_rn (a[i].y, s0, s1); a[i].x = __shfl_xor_sync (0xffffffffu, a[i].x, 1); }
        }
    }
  float r = 0;
  for (int i = 0; i < ILP; i++)
    r += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template<int MODE> void
run (const char *name, int warps_per_sm, int sms, float *out, double clk_ghz)
{
  const int threads = 256, blocks = sms * warps_per_sm * 32 / threads;
  cudaEvent_t e0, e1;
  cudaEventCreate (&e0); cudaEventCreate (&e1);
  k<MODE><<<blocks, threads>>> (out, 0.999f, 0.001f);
  cudaEventRecord (e0);
  k<MODE><<<blocks, threads>>> (out, 0.999f, 0.001f);
  cudaEventRecord (e1);
  cudaEventSynchronize (e1);
  float ms; cudaEventElapsedTime (&ms, e0, e1);
  const double ops = double (blocks) * threads * ITER * ILP * 2;           // fp32 results produced
  printf ("%-22s warps/SM %2d  %8.3f ms  %7.1f fp32 results / clk / SM (at %.3f GHz)\n", name, warps_per_sm, ms, ops / (ms * 1e-3) / (clk_ghz * 1e9) / sms, clk_ghz);
}

int main ()
{
  cudaDeviceProp p; cudaGetDeviceProperties (&p, 0);
  int khz; cudaDeviceGetAttribute (&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6;
  float *out; cudaMalloc (&out, size_t (p.multiProcessorCount) * 64 * 32 * 4);
  for (int wps : { 8, 16, 32 })
    {
      run<0> ("2 x FFMA", wps, p.multiProcessorCount, out, ghz);
      run<1> ("FFMA2", wps, p.multiProcessorCount, out, ghz);
      run<2> ("2 x FADD", wps, p.multiProcessorCount, out, ghz);
      run<3> ("FADD2", wps, p.multiProcessorCount, out, ghz);
      run<4> ("2 x FMUL", wps, p.multiProcessorCount, out, ghz);
      run<5> ("FMUL2", wps, p.multiProcessorCount, out, ghz);
      run<6> ("FFMA2 + SHFL", wps, p.multiProcessorCount, out, ghz);
      run<7> ("2 x FFMA + SHFL", wps, p.multiProcessorCount, out, ghz);
    }
  return 0;
}
