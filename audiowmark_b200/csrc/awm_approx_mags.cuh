// awm_approx_mags.cuh -- SyncFinder::search_approx (src/syncfinder.cc:171-256) in two passes.
//
// sync_decode (:116-153) for start frame s adds, for every sync entry e (a sync frame with its 30 "up" and 30 "down" bands),
// the band dB values of frame s + frame(e).  The inner sums  U_e(f) = sum_{u in up(e)} dB[f][u],  D_e(f) = sum_{d in down(e)} dB[f][d]
// do not depend on s, and every (f, e) pair is used by exactly one start frame.  So:
//   k_stft_mags   sync_fft for the four 256-sample shifts (warp = four frames, tile of 64 frames per CTA), then -- while the tile's
//                 81 band values still sit in shared memory -- U_e, D_e for all entries (warp = entry, lane = two frames,
//                 bands in list order) -> mags[shift][e][frame] (float2)
//   k_sync_gather thread = start frame: 510 (1020 in CLIP mode) coalesced float2 loads, six (u, d) accumulators filled bit by bit in
//                 frame order -> the per-bit sums k_sync_quality consumes.  This pass streams the whole matrix once: HBM bound.
// Versus one thread walking 30 600 scattered dB values per start frame (k_sync_approx) the gathered volume drops 30x.
// The float additions are grouped differently from the reference (per-frame subtotals first), which moves a quality by ~1e-5
// relative -- far inside the 2e-4 bar of the parity tests; search_refine recomputes the surviving candidates anyway.
#pragma once
#include "awm_kernels.cuh"

namespace awm {

constexpr int kMagWarps2 = 16;              // warps per CTA
constexpr int kMagTile = 128;               // frames per CTA: eight transforms per warp
constexpr int kMagEntChunk = 512;           // sync entries staged in shared memory at a time
constexpr size_t kMagSmem2 = fft_smem_bytes (kMagWarps2) + size_t (kBands) * kMagTile * sizeof (float) + size_t (kMagEntChunk) * 64;

// Phase 2 mapping: lane = four neighbouring frames of the tile, warp = one sync entry at a time.  All lanes read the same band row
// (float4 at consecutive addresses: conflict free), the band list of the entry is a broadcast read -- the per-entry sums then cost
// one LDS.128 + four FADD per band and four frames, in list order.
__global__ void __launch_bounds__ (kMagWarps2 * 32, 1)
k_stft_mags (const float *__restrict__ pcm, long long n_frames, int C, int n_out, int ld,
             const awm_sync_entry *__restrict__ ent, int n_ent,
             float2 *__restrict__ mags /* [4][n_ent][ld] */, unsigned char *__restrict__ have,
             long long wav_first, long long wav_last, const float2 *g_tw, const float *g_win)
{
  extern __shared__ __align__ (16) unsigned char smem[];
  FftSmem s = fft_smem_setup (smem, g_tw, g_win, kMagWarps2);
  float *tile = s.extra;                               // [81][128]
  uint32_t *ent_sm = reinterpret_cast<uint32_t *> (tile + kBands * kMagTile);      // [chunk][16 words]: 30 up bytes + 2 pad, 30 down bytes + 2 pad
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int shift_idx = blockIdx.x & 3, tile_idx = blockIdx.x >> 2;
  const int f0 = tile_idx * kMagTile;
  for (int pass = 0; pass < kMagTile / kMagWarps2; pass++)
    {
      const int r = w * (kMagTile / kMagWarps2) + pass;
      const int f = f0 + r;
      const long long start = (long long) shift_idx * 256 + (long long) f * kFrame;
      bool ok = f < n_out;
      if (ok)
        {
          const long long f_first = start * C, f_last = (start + kFrame) * C;
          if (f_last < wav_first || f_first > wav_last)   // frame in leading / trailing digital silence
            ok = false;
        }
      float acc[4] = { 0.f, 0.f, 0.f, 0.f };
      if (ok)
        frame_db_sum (pcm, n_frames, C, start, s, lane, acc);
      bands_to_array (acc, lane, tile + r, kMagTile);
      if (lane == 0 && f < n_out)
        have[(size_t) shift_idx * ld + f] = ok ? 1 : 0;
    }
  for (int e0 = 0; e0 < n_ent; e0 += kMagEntChunk)
    {
      const int n_chunk = n_ent - e0 < kMagEntChunk ? n_ent - e0 : kMagEntChunk;
      __syncthreads();                                   // tile complete / previous chunk consumed
      for (int i = threadIdx.x; i < n_chunk * 16; i += blockDim.x)
        {
          const awm_sync_entry *en = ent + e0 + (i >> 4);
          const int word = i & 15;
          const uint8_t *src = word < 8 ? en->up : en->down;
          const int b0 = (word & 7) * 4;
          uint32_t v = 0;
#pragma unroll
          for (int k = 0; k < 4; k++)
            if (b0 + k < kUD)
              v |= uint32_t (src[b0 + k]) << (8 * k);
          ent_sm[i] = v;
        }
      __syncthreads();
      for (int e = w; e < n_chunk; e += kMagWarps2)
        {
          const uint4 *ew = reinterpret_cast<const uint4 *> (ent_sm + e * 16);
          const uint4 u0 = ew[0], u1 = ew[1], d0 = ew[2], d1 = ew[3];
          const uint32_t uw[8] = { u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w };
          const uint32_t dw[8] = { d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w };
          const float4 *t4 = reinterpret_cast<const float4 *> (tile) + lane;
          float4 ua = make_float4 (0.f, 0.f, 0.f, 0.f), da = ua;
#pragma unroll
          for (int i = 0; i < kUD; i++)
            {
              const int bu = (uw[i >> 2] >> (8 * (i & 3))) & 0xff, bd = (dw[i >> 2] >> (8 * (i & 3))) & 0xff;
              const float4 vu = t4[bu * (kMagTile / 4)], vd = t4[bd * (kMagTile / 4)];
              ua.x += vu.x; ua.y += vu.y; ua.z += vu.z; ua.w += vu.w;
              da.x += vd.x; da.y += vd.y; da.z += vd.z; da.w += vd.w;
            }
          float4 *o = reinterpret_cast<float4 *> (mags + ((size_t) shift_idx * n_ent + e0 + e) * ld + f0) + 2 * lane;     // ld, f0 multiples of 128
          o[0] = make_float4 (ua.x, da.x, ua.y, da.y);
          o[1] = make_float4 (ua.z, da.z, ua.w, da.w);
        }
    }
}

constexpr int kGatherMaxEntries = 1024;

template<bool CHECK_HAVE> __global__ void __launch_bounds__ (256)
k_sync_gather (const float2 *__restrict__ mags, const unsigned char *__restrict__ have, int ld, int n_starts,
               const awm_sync_entry *__restrict__ ent, int n_ent, const int *__restrict__ bit_off, int n_bits,
               float *__restrict__ out_ud, int *__restrict__ out_cnt)
{
  __shared__ unsigned short frame_of[kGatherMaxEntries];
  for (int e = threadIdx.x; e < n_ent; e += blockDim.x)
    frame_of[e] = ent[e].frame;
  __syncthreads();
  const int s = blockIdx.x * blockDim.x + threadIdx.x, shift_idx = blockIdx.y;
  if (s >= n_starts)
    return;
  const float2 *m = mags + (size_t) shift_idx * n_ent * ld + s;
  const unsigned char *hv = have + (size_t) shift_idx * ld + s;
  for (int bit = 0; bit < n_bits; bit++)
    {
      float umag = 0.f, dmag = 0.f;
      int cnt = 0;
      const int e1 = bit_off[bit + 1];
#pragma unroll 5
      for (int e = bit_off[bit]; e < e1; e++)
        {
          const int fr = frame_of[e];
          if (!CHECK_HAVE || hv[fr])
            {
              const float2 v = __ldg (m + (size_t) e * ld + fr);
              umag += v.x;
              dmag += v.y;
              cnt++;
            }
        }
      const size_t o = ((size_t) shift_idx * n_starts + s) * n_bits + bit;
      out_ud[o * 2] = umag;
      out_ud[o * 2 + 1] = dmag;
      out_cnt[o] = cnt;
    }
}

} // namespace awm
