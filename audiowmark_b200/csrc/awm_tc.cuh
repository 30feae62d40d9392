// awm_tc.cuh -- Blackwell (sm_100a) plumbing used by the tensor-core kernels of this library: mbarriers, the bulk-copy engine
// (TMA, 1-D form), tensor memory (TMEM) and tcgen05.mma, as thin inline-PTX wrappers, plus the shared-memory operand layout the
// kernels build by hand.
//
// Operand layout (K-major, no swizzle -- "interleaved" canonical layout of the tcgen05 shared-memory descriptor): an operand of
// R rows (M or N) and K columns of 16-bit elements is stored as 8 x 8 core matrices of 128 contiguous bytes (8 rows x 16 bytes);
// core matrices that are neighbours in the row direction lie SBO = 128 bytes apart, neighbours in the K direction LBO = R * 16
// bytes apart (one "slab" of 8 K-columns for all rows):
//     byte offset of element (r, k) = (k / 8) * (R * 16) + (r / 8) * 128 + (r % 8) * 16 + (k % 8) * 2
// One tcgen05.mma of kind::f16 consumes K = 16, i.e. two slabs; the descriptor of K-step j starts at slab 2 j.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace awm { namespace tc {

__device__ __forceinline__ uint32_t smem_u32 (const void *p) { return (uint32_t) __cvta_generic_to_shared (p); }

// ---- mbarrier ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init (uint64_t *bar, uint32_t count)
{
  asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32 (bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive (uint64_t *bar)
{
  asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32 (bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx (uint64_t *bar, uint32_t bytes)
{
  asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32 (bar)), "r"(bytes) : "memory");
}
// spin until the phase with the given parity has completed (a fresh barrier passes a wait on parity 1)
__device__ __forceinline__ void mbar_wait (uint64_t *bar, uint32_t parity)
{
  const uint32_t addr = smem_u32 (bar);
  uint32_t done;
  do
    {
      asm volatile ("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
  while (!done);
}
// the same for warps that wait for work which takes microseconds (the MMA thread, the epilogue warps): try_wait with a suspend-time
// hint parks the thread in hardware until the phase completes or the hint (nanoseconds) runs out, so a waiting warp wakes up at once
// when its barrier flips but issues next to nothing meanwhile -- polling in a loop (even with nanosleep between polls) cost the
// schedulers of k_stft_mags_tc ~15 % of their issue slots (ncu source page: 13 M loop iterations per 10 min launch)
__device__ __forceinline__ void mbar_wait_relaxed (uint64_t *bar, uint32_t parity, uint32_t suspend_ns = 20000)
{
  const uint32_t addr = smem_u32 (bar);
  uint32_t done;
  do
    {
      asm volatile ("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done) : "r"(addr), "r"(parity), "r"(suspend_ns) : "memory");
    }
  while (!done);
}
// make generic-proxy writes to shared memory (st.shared) visible to the async proxy (tcgen05.mma / bulk copies read through it)
__device__ __forceinline__ void fence_proxy_async() { asm volatile ("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMA, 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP) ------------------------------
__device__ __forceinline__ void bulk_load (void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
  asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                :: "r"(smem_u32 (smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32 (bar)) : "memory");
}

// ---- tensor memory -------------------------------------------------------------------------------------------------------------
// whole warp; the base address (lane 0, first column) is written to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc (uint32_t *slot, uint32_t columns)
{
  asm volatile ("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32 (slot)), "r"(columns) : "memory");
  asm volatile ("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc (uint32_t taddr, uint32_t columns)
{
  asm volatile ("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(columns) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile ("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync()  { asm volatile ("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 columns of 32 bit: thread i of the warp receives row (lane field of taddr) + i, columns (column field) .. + 31.
// A warp may only touch the 32 TMEM lanes of its quadrant: lanes 32 * (warp id % 4) .. + 31.
__device__ __forceinline__ void tmem_ld_32x32 (uint32_t taddr, uint32_t (&r)[32])
{
  asm volatile ("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile ("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- tcgen05.mma -----------------------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, no swizzle: start address, LBO (K direction), SBO (row direction), all >> 4;
// bits 46..47 = 1 (descriptor version of sm_100), layout type (bits 61..63) = 0
__device__ __forceinline__ uint64_t smem_desc_kmajor (uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
  return uint64_t ((smem_addr >> 4) & 0x3fff) | (uint64_t ((lbo_bytes >> 4) & 0x3fff) << 16) | (uint64_t ((sbo_bytes >> 4) & 0x3fff) << 32)
       | (uint64_t (1) << 46);
}
// instruction descriptor of kind::f16: D = f32 (bits 4..5 = 1), A = B = f16 (formats 0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t idesc_f16_f32 (int M, int N) { return (1u << 4) | (uint32_t (N >> 3) << 17) | (uint32_t (M >> 4) << 24); }

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread for the whole CTA
__device__ __forceinline__ void mma_f16 (uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate)
{
  asm volatile ("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate ? 1u : 0u) : "memory");
}
// the mbarrier receives one arrival when every tcgen05.mma this thread issued so far has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit (uint64_t *bar)
{
  asm volatile ("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32 (bar)) : "memory");
}

// byte offset of element (r, k) of an R-row operand in the layout described at the top of this file
__host__ __device__ constexpr uint32_t operand_offset (int R, int r, int k)
{
  return uint32_t (k >> 3) * uint32_t (R * 16) + uint32_t (r >> 3) * 128u + uint32_t (r & 7) * 16u + uint32_t (k & 7) * 2u;
}

} } // namespace awm::tc
