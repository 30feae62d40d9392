// tools/tc_probe.cu -- development probe for audiowmark_b200/csrc/awm_tc.cuh: one 128 x 256 x 96 tcgen05.mma (kind::f16, fp32
// accumulators in TMEM) with hand-laid K-major operands, the B operand fetched by a 1-D bulk copy (TMA), result read back with
// tcgen05.ld and compared with a host computation.  Small integers and 0/1 masks make every product and sum exact, so any
// difference is a layout / descriptor error, not rounding.  Tries the descriptor encodings that are plausible from the public
// headers and reports which reproduce the host result.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tc_probe tools/tc_probe.cu && timeout 60 /tmp/tc_probe
#include "../audiowmark_b200/csrc/awm_tc.cuh"
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

using namespace awm::tc;

constexpr int M = 128, N = 256, K = 96;

__global__ void __launch_bounds__ (160, 1)
k_probe (const __half *a_rowmajor /* [M][K] */, const unsigned char *b_laid_out /* operand layout, N * K * 2 bytes */, float *d_out /* [M][N] */,
         uint32_t lbo_a, uint32_t sbo_a, uint32_t lbo_b, uint32_t sbo_b, int version_bit, int k_steps)
{
  extern __shared__ __align__ (1024) unsigned char smem[];
  unsigned char *sa = smem;                        // M * K * 2 = 24576
  unsigned char *sb = smem + M * K * 2;            // N * K * 2 = 49152
  uint64_t *bars = reinterpret_cast<uint64_t *> (sb + N * K * 2);   // [0] B landed, [1] MMA done
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *> (bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0)
    {
      mbar_init (&bars[0], 1);
      mbar_init (&bars[1], 1);
      fence_mbar_init();
    }
  if (warp == 4)
    tmem_alloc (tmem_slot, 256);
  for (int i = tid; i < M * K; i += blockDim.x)
    {
      const int r = i / K, k = i % K;
      *reinterpret_cast<__half *> (sa + operand_offset (M, r, k)) = a_rowmajor[i];
    }
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (tid == 128)
    {
      mbar_arrive_expect_tx (&bars[0], N * K * 2);
      bulk_load (sb, b_laid_out, N * K * 2, &bars[0]);
      mbar_wait (&bars[0], 0);
      tc_fence_after_sync();
      const uint64_t vmask = version_bit ? ~uint64_t (0) : ~(uint64_t (1) << 46);
      for (int j = 0; j < k_steps; j++)
        {
          const uint64_t da = smem_desc_kmajor (smem_u32 (sa) + j * 2 * (lbo_a > sbo_a ? lbo_a : sbo_a), lbo_a, sbo_a) & vmask;
          const uint64_t db = smem_desc_kmajor (smem_u32 (sb) + j * 2 * (lbo_b > sbo_b ? lbo_b : sbo_b), lbo_b, sbo_b) & vmask;
          mma_f16 (tmem, da, db, idesc_f16_f32 (M, N), j > 0);
        }
      mma_commit (&bars[1]);
    }
  if (warp < 4)
    {
      mbar_wait (&bars[1], 0);
      tc_fence_after_sync();
      for (int c0 = 0; c0 < N; c0 += 32)
        {
          uint32_t r[32];
          tmem_ld_32x32 (tmem + (uint32_t (warp * 32) << 16) + c0, r);
          tmem_ld_wait();
          for (int c = 0; c < 32; c++)
            d_out[(warp * 32 + lane) * N + c0 + c] = __uint_as_float (r[c]);
        }
      tc_fence_before_sync();
    }
  __syncthreads();
  if (warp == 4)
    tmem_dealloc (tmem, 256);
}

int
main()
{
  std::vector<__half> a (M * K);
  std::vector<unsigned char> b01 (N * K);
  std::vector<unsigned char> b_laid (N * K * 2, 0);
  srand (1);
  for (int i = 0; i < M * K; i++)
    a[i] = __float2half (float (rand() % 17 - 8));
  for (int n = 0; n < N; n++)
    for (int k = 0; k < K; k++)
      {
        b01[n * K + k] = (rand() % 3 == 0) ? 1 : 0;
        const __half v = __float2half (float (b01[n * K + k]));
        *reinterpret_cast<__half *> (&b_laid[operand_offset (N, n, k)]) = v;
      }
  std::vector<float> want (M * N);
  for (int m = 0; m < M; m++)
    for (int n = 0; n < N; n++)
      {
        float s = 0;
        for (int k = 0; k < K; k++)
          s += __half2float (a[m * K + k]) * b01[n * K + k];
        want[m * N + n] = s;
      }
  __half *d_a; unsigned char *d_b; float *d_d;
  cudaMalloc (&d_a, a.size() * 2); cudaMalloc (&d_b, b_laid.size()); cudaMalloc (&d_d, M * N * 4);
  cudaMemcpy (d_a, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy (d_b, b_laid.data(), b_laid.size(), cudaMemcpyHostToDevice);
  const size_t smem = M * K * 2 + N * K * 2 + 64;
  cudaFuncSetAttribute (k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, int (smem));
  int best = -1;
  for (int variant = 0; variant < 4; variant++)
    {
      const bool swapped = variant & 1;
      const int version_bit = (variant & 2) ? 0 : 1;
      const uint32_t slab_a = M * 16, slab_b = N * 16;
      cudaMemset (d_d, 0xff, M * N * 4);
      k_probe<<<1, 160, smem>>> (d_a, d_b, d_d, swapped ? 128 : slab_a, swapped ? slab_a : 128, swapped ? 128 : slab_b, swapped ? slab_b : 128, version_bit, K / 16);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess)
        {
          printf ("variant %d (swapped=%d version_bit=%d): CUDA error %s\n", variant, swapped, version_bit, cudaGetErrorString (e));
          return 2;
        }
      std::vector<float> got (M * N);
      cudaMemcpy (got.data(), d_d, M * N * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first = -1;
      for (int i = 0; i < M * N; i++)
        if (got[i] != want[i])
          {
            if (first < 0) first = i;
            bad++;
          }
      printf ("variant %d (LBO/SBO %s, descriptor version bit %d): %d of %d wrong", variant, swapped ? "swapped" : "as documented", version_bit, bad, M * N);
      if (first >= 0)
        printf ("  first at (%d, %d): got %g want %g", first / N, first % N, got[first], want[first]);
      printf ("\n");
      if (!bad && best < 0)
        best = variant;
    }
  printf ("RESULT: %s (variant %d)\n", best == 0 ? "awm_tc.cuh encoding verified" : best > 0 ? "another encoding works" : "no encoding works", best);
  return best == 0 ? 0 : 1;
}
