// mptc_pixel.cu -- K2p: pixel-granular inter search (DXTImage::InterPixelSearch,
// codec/dxt_image.cpp:776-832; pattern DXTImage::SetPattern, codec/dxt_image.h:135-164).
//
// The reference compiles this function but never calls it (the call site in Reencode is commented
// out, dxt_image.cpp:930-951), so it is offered as an analysis entry point of its own and is NOT part
// of the sequence encode.  Candidates are 4x4 cut-outs of the previous frame's index picture at pixel
// offsets (i, j), |i|, |j| < search_area, visited ring by ring; each is scored exactly like a block
// candidate (AssignIndices, ==, RecalculateEndpoints, swap check, Error: eval_candidate), the first
// candidate with err_diff <= 0 wins, otherwise the first strict minimum (:816-825).
//
// Candidate word: the 16 gathered 2-bit indices packed as they are.  The reference obtains it through
// Get4X4InterpolationBlock (:619-634), whose result depends on uninitialised memory (it may or may not
// be XORed with 0x55555555 depending on stack garbage); the defined reading is
// implemented here and pinned against the reference's own CompressedBlock methods by the tests.
//
// One CTA per target block: the (at most 35 x 35) previous-frame words its candidates can touch are
// staged in shared memory, every thread scores pattern entries tid, tid + 256, ...; the single-loop
// winner rule is two associative minima (first index with err_diff <= 0; minimum of (err_diff, index)).
#include "mptc_kernels.h"
#include "mptc_device.cuh"

namespace mptc {

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
k_inter_pixel_search(const uint8_t *__restrict__ frame, int w, int h, int bw, int bh, int sa, int n_pat,
                     const int8_t *__restrict__ pat, const uint64_t *__restrict__ cur_blocks,
                     const uint64_t *__restrict__ prev_blocks, int32_t *__restrict__ min_err_out,
                     uint8_t *__restrict__ motion_out, uint32_t *__restrict__ index_out, uint8_t *__restrict__ reassigned_out) {
  extern __shared__ uint32_t s_words[];   // [side][side] previous-frame index words around the target
  __shared__ TargetCtx t;
  __shared__ uint32_t s_first[kThreads / 32], s_best[kThreads / 32];
  const int b = blockIdx.x, bx = b % bw, by = b / bw;
  const int lo = -((sa - 1 + 3) / 4), hi = (sa + 2) / 4, side = hi - lo + 1;   // block offsets the cut-outs can reach
  if (threadIdx.x == 0) build_target(t, frame, w, bx, by, cur_blocks[b]);
  for (int p = threadIdx.x; p < side * side; p += kThreads) {
    const int i = bx + lo + p % side, j = by + lo + p / side;
    s_words[p] = (i >= 0 && j >= 0 && i < bw && j < bh) ? (uint32_t)(__ldg(prev_blocks + (size_t)j * bw + i) >> 32) : 0u;
  }
  __syncthreads();
  // the cut-out at pixel (x, y): row v comes from index-word byte (y + v) & 3 of the blocks in block row
  // (y + v) >> 2; its four 2-bit indices straddle two horizontally adjacent blocks
  auto gather = [&](int x, int y) -> uint32_t {
    const int cx = (x >> 2) - (bx + lo), ox = 2 * (x & 3);
    uint32_t word = 0;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int yy = y + v;
      const uint32_t *row = s_words + ((yy >> 2) - (by + lo)) * side + cx;
      const uint32_t left = (row[0] >> (8 * (yy & 3))) & 0xFFu;
      const uint32_t right = ox ? ((row[1] >> (8 * (yy & 3))) & 0xFFu) : 0u;
      word |= (((left >> ox) | (right << (8 - ox))) & 0xFFu) << (8 * v);
    }
    return word;
  };
  uint32_t first = 0xffffffffu, best = 0xffffffffu;
  for (int k = threadIdx.x; k < n_pat; k += kThreads) {
    const int x = 4 * bx + pat[2 * k], y = 4 * by + pat[2 * k + 1];
    if (x < 0 || y < 0 || x > w - 4 || y > h - 4) continue;          // dxt_image.cpp:795
    const int e = eval_candidate(t, gather(x, y));
    if (e == kRejected) continue;
    if (e <= 0) first = min(first, (uint32_t)k);
    best = min(best, ((uint32_t)(e + 65536) << 14) | (uint32_t)k);   // k < 2^14 (search_area <= 63)
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    first = min(first, __shfl_xor_sync(0xffffffffu, first, d));
    best = min(best, __shfl_xor_sync(0xffffffffu, best, d));
  }
  if ((threadIdx.x & 31) == 0) { s_first[threadIdx.x >> 5] = first; s_best[threadIdx.x >> 5] = best; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < kThreads / 32; ++q) { first = min(first, s_first[q]); best = min(best, s_best[q]); }
    int min_err = 0x7fffffff, k = -1;
    if (first != 0xffffffffu) { k = (int)first; min_err = 0; }
    else if (best != 0xffffffffu) { k = (int)(best & 0x3FFFu); min_err = (int)(best >> 14) - 65536; }
    uint32_t word = 0;
    int mx = 0, my = 0;
    if (k >= 0) {
      const int i = pat[2 * k], j = pat[2 * k + 1];
      word = gather(4 * bx + i, 4 * by + j);
      mx = i + 64; my = j + 64;                                         // :818
    }
    min_err_out[b] = min_err;
    motion_out[2 * b + 0] = (uint8_t)mx;
    motion_out[2 * b + 1] = (uint8_t)my;
    index_out[b] = word;
    reassigned_out[b] = (uint8_t)(k >= 0 && word != t.own_word);
  }
}

}  // namespace

// DXTImage::SetPattern (dxt_image.h:135-164): (0, 0), then the rings at Chebyshev distance 1 ..
// search_area - 1: top row right to left, the two side columns downwards, bottom row left to right.
int inter_pixel_pattern(int sa, int8_t *ij) {
  int n = 0;
  auto push = [&](int x, int y) { if (ij) { ij[2 * n] = (int8_t)x; ij[2 * n + 1] = (int8_t)y; } ++n; };
  push(0, 0);
  for (int ring = 1; ring < sa; ++ring) {
    for (int x = ring; x >= -ring; --x) push(x, ring);
    for (int y = ring - 1; y > -ring; --y) { push(ring, y); push(-ring, y); }
    for (int x = -ring; x <= ring; ++x) push(x, -ring);
  }
  return n;
}

void launch_inter_pixel_search(const uint8_t *frame, int w, int h, int sa, int n_pat, const int8_t *pat,
                               const uint64_t *cur_blocks, const uint64_t *prev_blocks, int32_t *min_err,
                               uint8_t *motion, uint32_t *index, uint8_t *reassigned, cudaStream_t s) {
  const int bw = w / 4, bh = h / 4;
  const int side = (sa + 2) / 4 + (sa - 1 + 3) / 4 + 1;
  k_inter_pixel_search<<<bw * bh, kThreads, (size_t)side * side * 4, s>>>(frame, w, h, bw, bh, sa, n_pat, pat, cur_blocks,
                                                                           prev_blocks, min_err, motion, index, reassigned);
}

}  // namespace mptc
