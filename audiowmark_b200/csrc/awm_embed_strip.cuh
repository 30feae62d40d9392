// awm_embed_strip.cuh -- k_embed_strip: the embed pipeline of k_embed (awm_kernels.cuh: FFTAnalyzer::run_fft + apply_frame_mod +
// WatermarkSynth::run + "+ orig" + Limiter::block_max; src/wmcommon.cc:91-121, src/wmadd.cc:61-84,215-250,564-565,
// src/limiter.cc:90-97) for stereo audio, reorganised as a streaming kernel:
//
//   * a warp owns a STRIP of consecutive frames and walks it front to back.  What an output frame needs from its neighbours -- the
//     first 104 samples of the previous frame's watermark signal and the last 104 of the next one's (the synthesis window spans
//     three frames, only its outer 103 samples on each side are non-zero) -- lives in fixed registers of fixed lanes (sample
//     x = lane + 32 brev5 (i) sits in register i), so it is simply carried from one iteration to the next: no shared-memory
//     exchange, no CTA barrier, and one halo frame per strip end instead of two per 14 frames (2 % recomputation instead of 14 %)
//   * the PCM of a frame arrives by a bulk copy (TMA, cp.async.bulk + mbarrier) in a per-warp landing buffer in shared memory, issued
//     as soon as the previous frame has been emitted; the forward transform reads it from there and so does the final "+ orig", so
//     every input byte crosses HBM -> SM exactly once, no global load ever stalls the butterflies, and the raw samples cost no
//     registers (twelve warps per SM instead of eight)
//   * an iteration stores one contiguous 8 KB span: the last 104 samples of frame m - 1 (now that frame m's contribution is known)
//     and the first 920 of frame m
// Arithmetic and rounding order are those of k_embed (products and sums of the synthesis rounded separately, in the reference's
// order); tests hold both kernels to the same oracle.  Mono / multichannel audio, the ragged last frame and unaligned buffers
// stay on k_embed.
#pragma once
#include "awm_kernels.cuh"
#include "awm_tc.cuh"

namespace awm {

constexpr int kStripWarps = 12;
constexpr size_t kStripSmem = fft_smem_bytes (kStripWarps) + 3 * kFrame * sizeof (float) + size_t (kStripWarps) * kFrame * sizeof (float2)
                              + 64 + kStripWarps * sizeof (uint64_t);

__global__ void __launch_bounds__ (kStripWarps * 32, 1)
k_embed_strip (EmbedArgs A, int strip_len)
{
  using namespace tc;
  extern __shared__ __align__ (16) unsigned char smem[];
  FftSmem s = fft_smem_setup (smem, A.tw, A.win, kStripWarps);
  float *synth = s.extra;                                   // [3072]
  float2 *pcm_all = reinterpret_cast<float2 *> (synth + 3 * kFrame);       // [warps][1024]: landing buffers of the bulk copies
  uint64_t *bars = reinterpret_cast<uint64_t *> (pcm_all + size_t (kStripWarps) * kFrame);
  for (int i = threadIdx.x; i < 3 * kFrame; i += blockDim.x)
    synth[i] = A.synth[i];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x < kStripWarps)
    mbar_init (&bars[threadIdx.x], 1);
  if (threadIdx.x == 0)
    fence_mbar_init();
  __syncthreads();
  uint64_t *bar = &bars[w];
  uint32_t phase = 0;
  float2 *pcmbuf = pcm_all + size_t (w) * kFrame;

  const long long m_first = A.frame_begin + ((long long) blockIdx.x * kStripWarps + w) * strip_len;
  const long long m_last = m_first + strip_len < A.frame_end ? m_first + strip_len : A.frame_end;      // exclusive
  if (m_first >= m_last)
    return;
  const long long n_real = (A.n_frames + kFrame - 1) / kFrame;          // frames that contain input
  const float2 *in2 = reinterpret_cast<const float2 *> (A.in);
  float2 *out2 = reinterpret_cast<float2 *> (A.out);
  auto exists = [&] (long long m) { return m >= 0 && m < n_real; };
  auto whole = [&] (long long m) { return m >= 0 && (m + 1) * kFrame <= A.n_frames; };      // the bulk copy needs all 1024 sample-frames
  auto prefetch = [&] (long long m)             // caller: after __syncwarp, no lane reads the landing buffer any more
    {
      if (lane == 0)
        {
          fence_proxy_async();
          mbar_arrive_expect_tx (bar, kFrame * sizeof (float2));
          bulk_load (pcmbuf, in2 + m * kFrame, kFrame * sizeof (float2), bar);
        }
    };
  // carried between iterations; the lane's head samples are registers brev5 (i) = 0..3 (the 4th only in lanes < 8: x < 104),
  // its tail samples brev5 (i) = 28..31 (the 1st only in lanes >= 24: x >= 920)
  float2 prev_head[4], tail_wm[4], tail_orig[4];
#pragma unroll
  for (int q = 0; q < 4; q++)
    prev_head[q] = tail_wm[q] = tail_orig[q] = make_float2 (0.f, 0.f);
  // Limiter::block_max: an iteration emits 1024 consecutive positions, which touch at most two 1-s limiter blocks; every lane keeps
  // the running maximum of the block it is in (pk_cur of block blk_cur, relative to A.blk0) and hands it to the block's global
  // maximum when the block changes
  float pk_cur = 0.f, pk_nxt = 0.f;
  long long blk_cur = -1, pos_boundary = 0;       // positions >= pos_boundary belong to block blk_cur + 1
  double snr_d = 0, snr_s = 0;
  auto give = [&] (long long blk, float v)
    {
      if (blk >= 0 && v > 0.f)
        atomicMax (A.peaks + blk, __float_as_uint (v));
    };
  auto begin_span = [&] (long long pos0)        // pos0: first position of the span, buffer coordinates
    {
      const long long blk = (A.stream_pos0 + pos0) / A.limiter_block - A.blk0;
      if (blk != blk_cur)
        {
          give (blk_cur, pk_cur);
          blk_cur = blk;
          pk_cur = 0.f;
        }
      pos_boundary = (blk + A.blk0 + 1) * (long long) A.limiter_block - A.stream_pos0;
    };
  auto end_span = [&] (long long pos_end)
    {
      if (pos_end > pos_boundary)                 // the span crossed into the next block: that one is current from now on
        {
          give (blk_cur, pk_cur);
          blk_cur++;
          pk_cur = pk_nxt;
          pk_nxt = 0.f;
        }
    };
  auto track_peak = [&] (long long pos, float ya, float yb)
    {
      const float v = fmaxf (fabsf (ya), fabsf (yb));
      if (pos < pos_boundary) pk_cur = fmaxf (pk_cur, v); else pk_nxt = fmaxf (pk_nxt, v);
    };

  if (whole (m_first - 1) && exists (m_first - 1))
    prefetch (m_first - 1);
  for (long long m = m_first - 1; m <= m_last; m++)
    {
      float re[32], im[32];
      const bool ex = exists (m), by_tma = ex && whole (m);
      const bool nxt_tma = m + 1 <= m_last && exists (m + 1) && whole (m + 1);
      if (by_tma)
        {
          mbar_wait (bar, phase);
          phase ^= 1;
        }
      else
        {
          /* ragged last frame / beyond the input: the lanes fill the landing buffer themselves (zeros where there is no input) */
#pragma unroll
          for (int j = 0; j < 32; j++)
            {
              const long long pos = m * kFrame + 32 * j + lane;
              pcmbuf[32 * j + lane] = (ex && pos < A.n_frames) ? __ldg (in2 + pos) : make_float2 (0.f, 0.f);
            }
          __syncwarp();
        }
      if (ex)
        {
#pragma unroll
          for (int j = 0; j < 32; j++)
            {
              const float wn = s.win[32 * j + lane];
              const float2 v = pcmbuf[32 * j + lane];
              re[j] = __fmul_rn (v.x, wn);
              im[j] = __fmul_rn (v.y, wn);
            }
          fft1024_warp (re, im, s.tw, s.xbuf, lane);
          const long long r = (A.frame_number0 + m) % (2LL * A.fpb);
          const uint8_t *fm = A.frame_mod + (size_t) r * (kMaxBand + 1);   // rows [0,fpb) = A, [fpb,2fpb) = B
          float inr[32], ini[32];
#pragma unroll
          for (int j = 0; j < 32; j++)
            inr[j] = ini[j] = 0.f;
          embed_mod_group<0> (re, im, inr, ini, fm, A.pow_up, A.pow_down, true, lane);
          embed_mod_group<1> (re, im, inr, ini, fm, A.pow_up, A.pow_down, true, lane);
          embed_mod_group<2> (re, im, inr, ini, fm, A.pow_up, A.pow_down, true, lane);
          embed_mod_group<3> (re, im, inr, ini, fm, A.pow_up, A.pow_down, true, lane);
          fft1024_warp (inr, ini, s.tw, s.xbuf, lane);
          // inverse result: sample x = lane + 32*brev5(i): channel A = ini[i], channel B = inr[i]
#pragma unroll
          for (int i = 0; i < 32; i++)
            {
              re[i] = ini[i];
              im[i] = inr[i];
            }
        }
      else
        {
#pragma unroll
          for (int i = 0; i < 32; i++)
            re[i] = im[i] = 0.f;
        }
      // ---- emit: the tail of frame m - 1 (x >= 920), then head and middle of frame m (x < 920)
      const bool emit_tail = m - 1 >= m_first && m - 1 < m_last && m - 1 < A.n_proc;
      const bool emit_body = m >= m_first && m < m_last && m < A.n_proc;
      if (A.limiter_block > 0 && (emit_tail || emit_body))
        begin_span (emit_tail ? (m - 1) * kFrame + kEdgeHi : m * kFrame);
#pragma unroll
      for (int i = 0; i < 32; i++)
        {
          const int b = brev5 (i);                 // x = lane + 32 b
          const int x = lane + 32 * b;
          if (b >= 28)                             // ---- tail region of frames m - 1 (finish) and m (defer)
            {
              const int q = b - 28;
              const bool in_tail = x >= kEdgeHi;
              if (emit_tail && in_tail)
                {
                  // ((prev*w2) + cur*w1) + next*w0 with prev = 0 here: tail_wm holds cur*w1 of frame m - 1, this frame is its "next"
                  const float wa = __fadd_rn (tail_wm[q].x, __fmul_rn (re[i], synth[x]));
                  const float wb = __fadd_rn (tail_wm[q].y, __fmul_rn (im[i], synth[x]));
                  const float oa = tail_orig[q].x, ob = tail_orig[q].y;
                  const float ya = A.delta_only ? wa : __fadd_rn (wa, oa), yb = A.delta_only ? wb : __fadd_rn (wb, ob);
                  const long long pos = (m - 1) * kFrame + x;
                  if (A.snr && m - 1 < A.snr_frames && pos >= A.snr_pos0 && pos < A.snr_pos1)
                    {
                      snr_d += double (wa) * double (wa) + double (wb) * double (wb);
                      snr_s += double (oa) * double (oa) + double (ob) * double (ob);
                    }
                  if (A.limiter_block > 0)
                    track_peak (pos, ya, yb);
                  if (pos < A.n_frames)
                    out2[pos] = make_float2 (ya, yb);
                }
              // this frame's own tail: cur*w1 now, the next iteration adds next*w0
              tail_wm[q] = make_float2 (__fmul_rn (re[i], synth[kFrame + x]), __fmul_rn (im[i], synth[kFrame + x]));
              tail_orig[q] = pcmbuf[x];
              if (emit_body && !in_tail)           // x in [896, 920): middle of frame m
                {
                  const float wa = tail_wm[q].x, wb = tail_wm[q].y;
                  const float2 og = pcmbuf[x]; const float oa = og.x, ob = og.y;
                  const float ya = A.delta_only ? wa : __fadd_rn (wa, oa), yb = A.delta_only ? wb : __fadd_rn (wb, ob);
                  const long long pos = m * kFrame + x;
                  if (A.snr && m < A.snr_frames && pos >= A.snr_pos0 && pos < A.snr_pos1)
                    {
                      snr_d += double (wa) * double (wa) + double (wb) * double (wb);
                      snr_s += double (oa) * double (oa) + double (ob) * double (ob);
                    }
                  if (A.limiter_block > 0)
                    track_peak (pos, ya, yb);
                  if (pos < A.n_frames)
                    out2[pos] = make_float2 (ya, yb);
                }
            }
          else
            {
              float wa = __fmul_rn (re[i], synth[kFrame + x]);
              float wb = __fmul_rn (im[i], synth[kFrame + x]);
              if (b < 4)                           // ---- head region: the previous frame reaches into it
                {
                  if (x < kEdge)
                    {
                      wa = __fadd_rn (__fmul_rn (prev_head[b].x, synth[2 * kFrame + x]), wa);
                      wb = __fadd_rn (__fmul_rn (prev_head[b].y, synth[2 * kFrame + x]), wb);
                    }
                  prev_head[b] = make_float2 (re[i], im[i]);
                }
              if (emit_body)
                {
                  const float2 og = pcmbuf[x]; const float oa = og.x, ob = og.y;
                  const float ya = A.delta_only ? wa : __fadd_rn (wa, oa), yb = A.delta_only ? wb : __fadd_rn (wb, ob);
                  const long long pos = m * kFrame + x;
                  if (A.snr && m < A.snr_frames && pos >= A.snr_pos0 && pos < A.snr_pos1)
                    {
                      snr_d += double (wa) * double (wa) + double (wb) * double (wb);
                      snr_s += double (oa) * double (oa) + double (ob) * double (ob);
                    }
                  if (A.limiter_block > 0)
                    track_peak (pos, ya, yb);
                  if (pos < A.n_frames)
                    out2[pos] = make_float2 (ya, yb);
                }
            }
        }
      if (A.limiter_block > 0 && (emit_tail || emit_body))
        end_span (emit_body ? m * kFrame + kEdgeHi : m * kFrame);
      __syncwarp();                                 // every lane has taken what it needs from the landing buffer
      if (nxt_tma)
        prefetch (m + 1);
    }
  if (A.limiter_block > 0)
    give (blk_cur, pk_cur);
  if (A.snr)
    {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
        {
          snr_d += __shfl_xor_sync (0xffffffffu, snr_d, off);
          snr_s += __shfl_xor_sync (0xffffffffu, snr_s, off);
        }
      if (lane == 0)
        {
          atomicAdd (A.snr, snr_d);
          atomicAdd (A.snr + 1, snr_s);
        }
    }
}

} // namespace awm
