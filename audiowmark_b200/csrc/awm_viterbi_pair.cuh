// awm_viterbi_pair.cuh -- k_viterbi_pair: conv_decode_soft (src/convcode.cc:128-213) with ONE code word on a CLUSTER OF TWO CTAs.
//
// k_viterbi (awm_kernels.cuh) gives a code word one SM; its time is the fp32 add pipe of that SM (2^15 states x 12 or 24 adds x 143
// steps), and a `get` of one hour has ~110 words for 148 SMs -- of which the third that are AB words (twice the adds) set the time.
// Here the 2^15 states of a word are split over the two CTAs of a thread-block cluster:
//   CTA h computes the new states n in [h 2^14, (h + 1) 2^14).  Their predecessors ps0 = n >> 1 and ps1 = ps0 + 2^14 lie in the
//   quarters Q_h = [h 2^13, (h + 1) 2^13) and Q_(h + 2), so a CTA keeps just those two quarters of the old metrics ("lo", "hi").
//   A new state n is a predecessor in the NEXT step for the CTA given by bit 13 of n: a thread stores its 32 new metrics either
//   into its own shared memory or into the peer's (distributed shared memory: mapa + st.shared::cluster), into the buffer of the
//   next step (the quarters are double buffered, so nothing is overwritten while it is still read).
//   One cluster barrier per trellis step (arrive.release / wait.acquire) replaces the two CTA barriers of k_viterbi.
// The arithmetic is k_viterbi's (viterbi_step: metrics accumulated in the reference's order with packed adds, strict '<' tie rule),
// so bits and error are identical; AB words are put first in the grid, the A / B words (half the adds) fill in behind them.
//
// MEASURED (B200, 110 words of a 1 h `get`): 1.16 ms against 0.72 ms for k_viterbi.  The adds per SM halve as planned, but
// barrier.cluster.arrive.release / wait.acquire compiles to UCGABAR_ARV + MEMBAR.ALL.GPU + UCGABAR_WAIT and costs ~3 us per trellis
// step (143 steps), and 110 clusters need two rounds on 74 SM pairs.  Kept as a tested alternative (AWM_VITERBI=pair); what would make
// it pay is handing the metrics over with st.async + mbarrier complete_tx (no cluster-wide fence) and pairing only the AB words.
#pragma once
#include "awm_kernels.cuh"

namespace awm {

constexpr int kPairQuarter = kVitStates / 4;                              // 8192 states per quarter
constexpr int kPairQuarterPadded = kPairQuarter + kPairQuarter / 32 * 4;  // vit_pos padding
constexpr size_t viterbi_pair_smem_bytes (int steps) { return size_t (4) * kPairQuarterPadded * sizeof (float) + size_t (steps) * 12 * sizeof (float); }

namespace pair {
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile ("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync()
{
  asm volatile ("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
/* address of the same shared-memory location in the CTA `rank` of the cluster */
__device__ __forceinline__ uint32_t peer_address (const void *local, uint32_t rank)
{
  uint32_t r;
  asm volatile ("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"((uint32_t) __cvta_generic_to_shared (local)), "r"(rank));
  return r;
}
__device__ __forceinline__ void store_cluster (uint32_t addr, float4 v)
{
  asm volatile ("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
}

template<int TYPE> __device__ __forceinline__ void
viterbi_pair_run (float *quarters /* [2 buffers][lo, hi][kPairQuarterPadded] */, uint32_t *dec, const float *coded, float (*m0)[12], float (*m1)[12],
                  int steps, int tid, uint32_t h)
{
  constexpr int RATE = TYPE == AWM_BLOCK_AB ? 12 : 6;
  const int u_word = int (h) * kVitThreads + tid;                         // decision word = group of 32 new states this thread owns
  unsigned hi = 0;
#pragma unroll
  for (int p = 0; p < RATE; p++)
    hi |= unsigned (__popc ((32u * unsigned (u_word)) & type_generator<TYPE> (p)) & 1) << p;
  /* where the 32 new metrics go: CTA (bit 13 of the state) = (tid >= 256), array lo for the states of CTA 0, hi for those of CTA 1 */
  const uint32_t dest_cta = tid >= kVitThreads / 2 ? 1u : 0u;
  const int dest_index = vit_pos (32 * (tid & (kVitThreads / 2 - 1)));
  for (int t = 0; t < steps; t++)
    {
      const int cur = t & 1;
      const float *lo = quarters + size_t (cur) * 2 * kPairQuarterPadded, *hi_arr = lo + kPairQuarterPadded;
      float *next = quarters + size_t (cur ^ 1) * 2 * kPairQuarterPadded + size_t (h) * kPairQuarterPadded + dest_index;
      float outv[32];
      dec[(size_t) t * kVitWords + u_word] = viterbi_step<TYPE> (lo, hi_arr, outv, hi, m0[cur], m1[cur], tid);
      if (tid < RATE && t + 1 < steps)                                    // branch metrics of the next step (the other half of m0 / m1)
        {
          const float c = coded[(t + 1) * RATE + tid];
          m0[cur ^ 1][tid] = __fmul_rn (c, c);
          m1[cur ^ 1][tid] = __fmul_rn (c - 1.0f, c - 1.0f);
        }
      if (dest_cta == h)
        {
#pragma unroll
          for (int k = 0; k < 8; k++)
            reinterpret_cast<float4 *> (next)[k] = make_float4 (outv[4 * k], outv[4 * k + 1], outv[4 * k + 2], outv[4 * k + 3]);
        }
      else
        {
          const uint32_t remote = pair::peer_address (next, dest_cta);
#pragma unroll
          for (int k = 0; k < 8; k++)
            pair::store_cluster (remote + 16 * k, make_float4 (outv[4 * k], outv[4 * k + 1], outv[4 * k + 2], outv[4 * k + 3]));
        }
      pair::cluster_sync();                                               // new metrics (local and remote) in place, old ones no longer read
    }
}

__global__ void __cluster_dims__ (2, 1, 1) __launch_bounds__ (kVitThreads)
k_viterbi_pair (const float *__restrict__ raw, const long long *__restrict__ raw_off, int n_msg, const int *__restrict__ block_types, int hard,
                int max_steps, const int *__restrict__ job_order /* grid position -> job: AB words first */,
                uint32_t *__restrict__ dec_buf /* [job][steps][kVitWords] */, unsigned char *__restrict__ bits_out, float *__restrict__ err_out)
{
  extern __shared__ __align__ (16) unsigned char smem[];
  float *quarters = reinterpret_cast<float *> (smem);
  float *coded = quarters + 4 * kPairQuarterPadded;
  __shared__ float m0[2][12], m1[2][12];
  __shared__ double s_mean;
  const uint32_t h = pair::cluster_rank();
  const int job = job_order[blockIdx.x >> 1], tid = threadIdx.x;
  const int btype = block_types[job];
  const int rate = (btype == AWM_BLOCK_AB) ? 12 : 6;
  const int steps = n_msg + AWM_VITERBI_ORDER;
  const int n_coded = steps * rate;

  const float *rj = raw + raw_off[job];
  if (tid == 0)
    {
      double mean = 0;
      for (int i = 0; i < n_coded; i++)
        mean += fabs (double (rj[i]));
      s_mean = mean / n_coded;
    }
  __syncthreads();
  /* digital silence: see k_viterbi.  Both CTAs of the pair take this exit (same data), before the first cluster barrier */
  if (!hard && !(s_mean > 0))
    {
      if (h == 0)
        {
          for (int i = tid; i < n_msg; i += blockDim.x)
            bits_out[(size_t) job * n_msg + i] = 0;
          if (tid == 0)
            err_out[job] = -1.0f / float (n_coded);
        }
      return;
    }
  for (int i = tid; i < n_coded; i += blockDim.x)
    coded[i] = hard ? (rj[i] > 0 ? 1.0f : 0.0f) : float (0.5 * (double (rj[i]) / s_mean + 1));
  /* step 0 reads buffer 0: state 0 (quarter 0 = "lo" of CTA 0) has metric 0, every other state is unreachable */
  for (int i = tid; i < 2 * kPairQuarter; i += blockDim.x)
    quarters[(i >> 13) * kPairQuarterPadded + vit_pos (i & (kPairQuarter - 1))] = (i == 0 && h == 0) ? 0.f : INFINITY;
  __syncthreads();
  if (tid < rate)
    {
      const float c = coded[tid];
      m0[0][tid] = __fmul_rn (c, c);
      m1[0][tid] = __fmul_rn (c - 1.0f, c - 1.0f);
    }
  pair::cluster_sync();                                                   // both CTAs are set up before anybody writes into the peer

  uint32_t *dec = dec_buf + (size_t) job * max_steps * kVitWords;
  if (btype == AWM_BLOCK_A)
    viterbi_pair_run<AWM_BLOCK_A> (quarters, dec, coded, m0, m1, steps, tid, h);
  else if (btype == AWM_BLOCK_B)
    viterbi_pair_run<AWM_BLOCK_B> (quarters, dec, coded, m0, m1, steps, tid, h);
  else
    viterbi_pair_run<AWM_BLOCK_AB> (quarters, dec, coded, m0, m1, steps, tid, h);
  /* the last cluster barrier of the run has made CTA 1's decision words visible; state 0 ends up in "lo" of CTA 0 */
  if (h == 0)
    {
      if (tid == 0)
        err_out[job] = quarters[size_t (steps & 1) * 2 * kPairQuarterPadded] / float (n_coded);
      if (tid < 32)
        viterbi_traceback (dec, steps, n_msg, bits_out + (size_t) job * n_msg, tid);
    }
}

} // namespace awm
