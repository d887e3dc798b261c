// mptc_sparse.cu -- K3s: intra search for the few blocks of an INTER frame that the inter search
// left over (DXTImage::Reencode, codec/dxt_image.cpp:910-955, reached when :890 fails).
//
// On ordinary content these are a few dozen blocks per frame, scattered over the frame, so the
// row wavefront (mptc_intra.cu) would spend its time handing rows from CTA to CTA.  Here every
// leftover block is one work item, dependencies are tracked per BLOCK:
//   * every CTA builds the frame's raster-ordered list of leftover blocks itself (rows flagged
//     by K2 in row_todo, then the flags of those rows);
//   * items are handed out in raster order through a ticket; a target's window only contains
//     blocks that precede it in raster order, so a CTA only ever waits for items whose tickets
//     were taken before its own: no deadlock for any grid size;
//   * a window position whose flag is still 0 (leftover, undecided) is polled until its owner
//     publishes flag = 2 behind a fence; everything else is final already.
// Each window position is evaluated directly (the same code as the direct kernels).  Frames with
// more than kSparseMaxItems leftovers are left to the row wavefront, which de-duplicates; the
// count is handed to it through n_unique[f] (free until K4 runs).
#include "mptc_kernels.h"
#include "mptc_device.cuh"

namespace mptc {

namespace {

#ifndef MPTC_SPARSE_THREADS
#define MPTC_SPARSE_THREADS 1024
#endif
constexpr int kThreads = MPTC_SPARSE_THREADS, kWarps = kThreads / 32;
constexpr int kMaxRows = 4096;   // block rows per frame the row list can hold

__device__ __forceinline__ uint8_t ld_flag(const uint8_t *p) { return *reinterpret_cast<const volatile uint8_t *>(p); }

}  // namespace

__global__ void __launch_bounds__(kThreads)
k_intra_sparse(SeqView v, int k_in_gop, int sa, int thr, int *__restrict__ tickets) {
  __shared__ uint32_t s_list[kSparseMaxItems];
  __shared__ uint16_t s_rows[kMaxRows];
  __shared__ int s_row_off[kMaxRows + 1];
  __shared__ int s_wsum[kWarps];
  __shared__ int s_n, s_item;
  __shared__ TargetCtx s_t;
  __shared__ WinnerState s_red[kWarps];

  const int f = v.first + blockIdx.y * v.gop + k_in_gop;
  if (f >= v.first + v.count) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint8_t *frame = v.rgb + v.frame_bytes * f;
  uint64_t *cur = v.final_blocks + (size_t)f * v.nb;
  const uint64_t *init = v.init_blocks + (size_t)f * v.nb;
  uint8_t *flags = v.flags + (size_t)f * v.nb;
  uint8_t *motion = v.motion + (size_t)f * v.nb * 2;
  const uint8_t *row_todo = v.row_todo + (size_t)f * v.bh;
  int *ticket = tickets + blockIdx.y;

  // ---- rows with leftovers, in order ------------------------------------------------------------
  int n_rows = 0;
  for (int r0 = 0; r0 < v.bh; r0 += kThreads) {
    const int r = r0 + tid;
    const bool has = r < v.bh && row_todo[r] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, has);
    if (lane == 0) s_wsum[wid] = __popc(m);
    __syncthreads();
    int before = n_rows, total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const int c = s_wsum[w];
      if (w < wid) before += c;
      total += c;
    }
    if (has) {
      const int at = before + __popc(m & ((1u << lane) - 1u));
      if (at < kMaxRows) s_rows[at] = (uint16_t)r;
    }
    n_rows += total;
    __syncthreads();
  }
  bool overflow = n_rows > kMaxRows || v.bh > 65535;

  // ---- leftovers per row (warp = row), exclusive prefix, then the ordered list ---------------------
  // The flags of leftover blocks change from 0 to 2 while other CTAs work; a CTA that starts late
  // could miss them, so "leftover" is taken from the motion bytes K2 did NOT write... they are
  // not reset either.  Instead every decided leftover keeps row_todo and gets flag 2, and the
  // list is built from flags != 1.
  if (!overflow) {
    for (int q = wid; q < n_rows; q += kWarps) {
      const uint8_t *fr = flags + (size_t)s_rows[q] * v.bw;
      int c = 0;
      for (int x = lane; x < v.bw; x += 32) c += ld_flag(fr + x) != 1;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
      if (lane == 0) s_row_off[q + 1] = c;
    }
    if (tid == 0) s_row_off[0] = 0;
    __syncthreads();
    if (wid == 0) {   // inclusive scan over rows by one warp
      int carry = 0;
      for (int q0 = 0; q0 < n_rows; q0 += 32) {
        const int q = q0 + lane;
        int x = q < n_rows ? s_row_off[q + 1] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, x, d);
          if (lane >= d) x += y;
        }
        if (q < n_rows) s_row_off[q + 1] = carry + x;
        carry += __shfl_sync(0xffffffffu, x, 31);
      }
      if (lane == 0) s_n = carry;
    }
    __syncthreads();
    overflow = s_n > kSparseMaxItems;
  }
  const int n_items = overflow ? 0 : s_n;
  if (blockIdx.x == 0 && tid == 0) v.n_unique[f] = overflow ? 0xFFFFFFFFu : (uint32_t)n_items;
  if (overflow || n_items == 0) return;   // the row wavefront takes the frame / nothing to do
  for (int q = wid; q < n_rows; q += kWarps) {
    const int row = s_rows[q];
    const uint8_t *fr = flags + (size_t)row * v.bw;
    int at = s_row_off[q];
    for (int x0 = 0; x0 < v.bw; x0 += 32) {
      const int x = x0 + lane;
      const bool left = x < v.bw && ld_flag(fr + x) != 1;
      const unsigned m = __ballot_sync(0xffffffffu, left);
      if (left) s_list[at + __popc(m & ((1u << lane) - 1u))] = (uint32_t)(row * v.bw + x);
      at += __popc(m);
    }
  }
  __syncthreads();

  // ---- work items in raster order ---------------------------------------------------------------------
  const int W = 2 * sa;
  for (;;) {
    if (tid == 0) s_item = atomicAdd(ticket, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= n_items) return;
    const int b = (int)s_list[item];
    const int bx = b % v.bw, by = b / v.bw;
    if (tid == 0) build_target(s_t, frame, v.w, bx, by, init[b]);
    __syncthreads();
    WinnerState s;
    winner_init(s);
    for (int p = tid; p < W * W; p += kThreads) {
      const int row = p / W, col = p - row * W;          // scan order: j downwards, i downwards
      const int j = by - row, i = bx + sa - 1 - col;
      if (i < 0 || j < 0 || i >= v.bw || (row == 0 && i >= bx)) continue;
      const size_t idx = (size_t)j * v.bw + i;
      if (ld_flag(flags + idx) == 0) {                   // an earlier leftover, still undecided
        while (ld_flag(flags + idx) == 0) __nanosleep(64);
        __threadfence();
      }
      const uint32_t word = __ldcg(reinterpret_cast<const uint32_t *>(cur) + 2 * idx + 1);
      winner_update(s, eval_candidate(s_t, word), row, col, W);
    }
    winner_block_reduce<kWarps>(s, s_red);
    if (tid == 0) {
      int row, col;
      const int min_err = winner_resolve(s, W, row, col);
      if (min_err <= thr) {
        const size_t src = (size_t)(by - row) * v.bw + (bx + sa - 1 - col);
        cur[b] = winning_block(s_t, __ldcg(reinterpret_cast<const uint32_t *>(cur) + 2 * src + 1));
        motion[2 * b + 0] = (uint8_t)(2 * sa - 1 - col);   // x = (i - bx) + sa
        motion[2 * b + 1] = (uint8_t)(2 * sa - 1 - row);   // y = (j - by) + 2sa - 1
      } else {
        motion[2 * b + 0] = 255;
        motion[2 * b + 1] = 255;
      }
      __threadfence();
      *reinterpret_cast<volatile uint8_t *>(flags + b) = 2;
    }
    // s_item / s_t are rewritten only after the next barrier pair; s_red was consumed above
  }
}

void launch_intra_sparse(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, int *tickets, int ctas_per_frame,
                         cudaStream_t s) {
  dim3 grid(ctas_per_frame, n_gops);
  k_intra_sparse<<<grid, kThreads, 0, s>>>(v, k_in_gop, sa, thr, tickets);
}

}  // namespace mptc
