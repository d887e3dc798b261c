// mptc_kernels.cu -- sm_100a kernels of the MPTC encoder hot path.
//
//   K1 k_dxt1_fit          stb_compress_dxt_block(HIGHQUAL) per 4x4 block   (Include/stb_dxt.h:467-538)
//   K2 k_inter_search      DXTImage::InterBlockSearch + winner apply        (codec/dxt_image.cpp:715-774, :885-908)
//                          direct form (one CTA per target); the tiled form is in mptc_inter.cu
//   K3 k_intra_wavefront   DXTImage::IntraSearch + winner apply, raster     (codec/dxt_image.cpp:652-713, :912-955)
//                          dependency resolved by a row-staggered wavefront; direct form, the tiled
//                          form is in mptc_intra_rows.cu, the leftover kernel of inter frames in mptc_sparse.cu
//   K4 k_compact_count /   _unique_palette push_backs as an ordered prefix  (codec/dxt_image.cpp:953-954)
//      k_compact_unique    sum over chunks of 1024 blocks
//   K5 k_endpoint_planes   RGB565 -> YCoCg667 -> 64x64 5/3 wavelet -> u8    (codec/codec.cpp:804-839, wavelet.cpp:30-131)
//
// No tensor cores: nothing here is a dense contraction (integer / ordered-FP32 work).
#include "mptc_kernels.h"

#include <cstring>

#include <cstdlib>
#include "mptc_device.cuh"


namespace mptc {

// ------------------------------------------------------------------------------------------
// stb single-colour tables (stb__OMatch5/6, stb_dxt.h:111-137), filled by the host at
// context creation with the same search stb__PrepareOptTable runs.
// ------------------------------------------------------------------------------------------
__constant__ uint8_t c_omatch5[256][2];
__constant__ uint8_t c_omatch6[256][2];

cudaError_t upload_tables(const uint8_t *omatch5, const uint8_t *omatch6) {
  cudaError_t e = cudaMemcpyToSymbol(c_omatch5, omatch5, 512);
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbol(c_omatch6, omatch6, 512);
}

// ------------------------------------------------------------------------------------------
// K1: DXT1 endpoint fit
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int mul8bit(int a, int b) {  // stb_dxt.h:64-68
  int t = a * b + 128;
  return (t + (t >> 8)) >> 8;
}
__device__ __forceinline__ uint32_t quant565(int r, int g, int b) {  // stb__As16Bit :82-85
  return (uint32_t)((mul8bit(r, 31) << 11) + (mul8bit(g, 63) << 5) + mul8bit(b, 31));
}
__device__ __forceinline__ int exp5(uint32_t v) { return (int)((v << 3) | (v >> 2)); }
__device__ __forceinline__ int exp6(uint32_t v) { return (int)((v << 2) | (v >> 4)); }

// stb__EvalColors + stb__MatchColorsBlock, non-dither branch (:139-145, :176-215)
__device__ __forceinline__ uint32_t match_indices(const uint32_t *px, uint32_t c0, uint32_t c1) {
  int col[4][3];
  col[0][0] = exp5(c0 >> 11); col[0][1] = exp6((c0 >> 5) & 63u); col[0][2] = exp5(c0 & 31u);
  col[1][0] = exp5(c1 >> 11); col[1][1] = exp6((c1 >> 5) & 63u); col[1][2] = exp5(c1 & 31u);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    col[2][k] = (2 * col[0][k] + col[1][k]) / 3;
    col[3][k] = (2 * col[1][k] + col[0][k]) / 3;
  }
  int dr = col[0][0] - col[1][0], dg = col[0][1] - col[1][1], db = col[0][2] - col[1][2];
  int stops[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) stops[i] = col[i][0] * dr + col[i][1] * dg + col[i][2] * db;
  int c0pt = (stops[1] + stops[3]) >> 1;
  int half = (stops[3] + stops[2]) >> 1;
  int c3pt = (stops[2] + stops[0]) >> 1;
  uint32_t mask = 0;
#pragma unroll
  for (int i = 15; i >= 0; --i) {
    int dot = (int)(px[i] & 0xFF) * dr + (int)((px[i] >> 8) & 0xFF) * dg + (int)((px[i] >> 16) & 0xFF) * db;
    mask <<= 2;
    if (dot < half) mask |= (dot < c0pt) ? 1u : 3u;
    else            mask |= (dot < c3pt) ? 2u : 0u;
  }
  return mask;
}

// stb__OptimizeColorsBlock (:273-375)
__device__ __forceinline__ void pca_endpoints(const uint32_t *px, uint32_t &mx16, uint32_t &mn16) {
  int mu[3], lo[3], hi[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    int s = 0, mn = 255, mx = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      int v = (int)((px[i] >> (8 * ch)) & 0xFF);
      s += v; mn = min(mn, v); mx = max(mx, v);
    }
    mu[ch] = (s + 8) >> 4; lo[ch] = mn; hi[ch] = mx;
  }
  int cov[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    int r = (int)(px[i] & 0xFF) - mu[0], g = (int)((px[i] >> 8) & 0xFF) - mu[1], b = (int)((px[i] >> 16) & 0xFF) - mu[2];
    cov[0] += r * r; cov[1] += r * g; cov[2] += r * b;
    cov[3] += g * g; cov[4] += g * b; cov[5] += b * b;
  }
  float cf[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) cf[i] = __fdiv_rn(__int2float_rn(cov[i]), 255.0f);
  float vr = __int2float_rn(hi[0] - lo[0]), vg = __int2float_rn(hi[1] - lo[1]), vb = __int2float_rn(hi[2] - lo[2]);
#pragma unroll
  for (int it = 0; it < 4; ++it) {  // nIterPower = 4, each op rounded (:331-340)
    float r = __fadd_rn(__fadd_rn(__fmul_rn(vr, cf[0]), __fmul_rn(vg, cf[1])), __fmul_rn(vb, cf[2]));
    float g = __fadd_rn(__fadd_rn(__fmul_rn(vr, cf[1]), __fmul_rn(vg, cf[3])), __fmul_rn(vb, cf[4]));
    float b = __fadd_rn(__fadd_rn(__fmul_rn(vr, cf[2]), __fmul_rn(vg, cf[4])), __fmul_rn(vb, cf[5]));
    vr = r; vg = g; vb = b;
  }
  double magn = fabs((double)vr);  // the one FP64 spot (:342-354)
  magn = fmax(magn, fabs((double)vg));
  magn = fmax(magn, fabs((double)vb));
  int ar, ag, ab;
  if (magn < 4.0) { ar = 299; ag = 587; ab = 114; }
  else {
    magn = __ddiv_rn(512.0, magn);
    ar = __double2int_rz(__dmul_rn((double)vr, magn));
    ag = __double2int_rz(__dmul_rn((double)vg, magn));
    ab = __double2int_rz(__dmul_rn((double)vb, magn));
  }
  int dmin = 0x7fffffff, dmax = -0x7fffffff;
  uint32_t pmin = 0, pmax = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    int dot = (int)(px[i] & 0xFF) * ar + (int)((px[i] >> 8) & 0xFF) * ag + (int)((px[i] >> 16) & 0xFF) * ab;
    if (dot < dmin) { dmin = dot; pmin = px[i]; }
    if (dot > dmax) { dmax = dot; pmax = px[i]; }
  }
  mx16 = quant565((int)(pmax & 0xFF), (int)((pmax >> 8) & 0xFF), (int)((pmax >> 16) & 0xFF));
  mn16 = quant565((int)(pmin & 0xFF), (int)((pmin >> 8) & 0xFF), (int)((pmin >> 16) & 0xFF));
}

__device__ __forceinline__ int sclamp(float y, int hi) {  // stb__sclamp :377-383 (values stay tiny)
  int x = __float2int_rz(y);
  return min(max(x, 0), hi);
}

__device__ __forceinline__ uint32_t omatch_pair(int r, int g, int b, int which) {
  return ((uint32_t)c_omatch5[r][which] << 11) | ((uint32_t)c_omatch6[g][which] << 5) | (uint32_t)c_omatch5[b][which];
}

// stb__RefineBlock (:388-464)
__device__ __forceinline__ bool refine_endpoints(const uint32_t *px, uint32_t &mx16, uint32_t &mn16, uint32_t mask) {
  uint32_t old_mx = mx16, old_mn = mn16;
  if ((mask ^ (mask << 2)) < 4u) {
    int r = 8, g = 8, b = 8;
#pragma unroll
    for (int i = 0; i < 16; ++i) { r += (int)(px[i] & 0xFF); g += (int)((px[i] >> 8) & 0xFF); b += (int)((px[i] >> 16) & 0xFF); }
    r >>= 4; g >>= 4; b >>= 4;
    mx16 = omatch_pair(r, g, b, 0);
    mn16 = omatch_pair(r, g, b, 1);
  } else {
    int a1r = 0, a1g = 0, a1b = 0, a2r = 0, a2g = 0, a2b = 0, akku = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      uint32_t step = (mask >> (2 * i)) & 3u;
      int w1 = (step == 0) ? 3 : (step == 1 ? 0 : (step == 2 ? 2 : 1));
      akku += (step == 0) ? 0x090000 : (step == 1 ? 0x000900 : (step == 2 ? 0x040102 : 0x010402));
      int r = (int)(px[i] & 0xFF), g = (int)((px[i] >> 8) & 0xFF), b = (int)((px[i] >> 16) & 0xFF);
      a1r += w1 * r; a1g += w1 * g; a1b += w1 * b;
      a2r += r; a2g += g; a2b += b;
    }
    a2r = 3 * a2r - a1r; a2g = 3 * a2g - a1g; a2b = 3 * a2b - a1b;
    int xx = akku >> 16, yy = (akku >> 8) & 0xff, xy = akku & 0xff;
    // 3.0f*31.0f/255.0f folds to one FP32 constant in the reference build as well
    float frb = __fdiv_rn(3.0f * 31.0f / 255.0f, __int2float_rn(xx * yy - xy * xy));
    float fg = __fdiv_rn(__fmul_rn(frb, 63.0f), 31.0f);
    mx16  = (uint32_t)sclamp(__fadd_rn(__fmul_rn(__int2float_rn(a1r * yy - a2r * xy), frb), 0.5f), 31) << 11;
    mx16 |= (uint32_t)sclamp(__fadd_rn(__fmul_rn(__int2float_rn(a1g * yy - a2g * xy), fg), 0.5f), 63) << 5;
    mx16 |= (uint32_t)sclamp(__fadd_rn(__fmul_rn(__int2float_rn(a1b * yy - a2b * xy), frb), 0.5f), 31);
    mn16  = (uint32_t)sclamp(__fadd_rn(__fmul_rn(__int2float_rn(a2r * xx - a1r * xy), frb), 0.5f), 31) << 11;
    mn16 |= (uint32_t)sclamp(__fadd_rn(__fmul_rn(__int2float_rn(a2g * xx - a1g * xy), fg), 0.5f), 63) << 5;
    mn16 |= (uint32_t)sclamp(__fadd_rn(__fmul_rn(__int2float_rn(a2b * xx - a1b * xy), frb), 0.5f), 31);
  }
  return old_mn != mn16 || old_mx != mx16;
}

// stb__CompressColorBlock (:467-538)
__device__ __forceinline__ uint64_t fit_block(const uint32_t *px) {
  uint32_t mx, mn, mask;
  bool constant = true;
#pragma unroll
  for (int i = 1; i < 16; ++i) constant = constant && (px[i] == px[0]);
  if (constant) {
    int r = (int)(px[0] & 0xFF), g = (int)((px[0] >> 8) & 0xFF), b = (int)((px[0] >> 16) & 0xFF);
    mask = 0xAAAAAAAAu;
    mx = omatch_pair(r, g, b, 0);
    mn = omatch_pair(r, g, b, 1);
  } else {
    pca_endpoints(px, mx, mn);
    mask = (mx != mn) ? match_indices(px, mx, mn) : 0u;
    for (int pass = 0; pass < 2; ++pass) {  // HIGHQUAL: refinecount = 2
      uint32_t last = mask;
      if (refine_endpoints(px, mx, mn, mask)) {
        if (mx != mn) mask = match_indices(px, mx, mn);
        else { mask = 0; break; }
      }
      if (mask == last) break;
    }
  }
  if (mx < mn) { uint32_t t = mn; mn = mx; mx = t; mask ^= 0x55555555u; }
  return (uint64_t)mx | ((uint64_t)mn << 16) | ((uint64_t)mask << 32);
}

__global__ void __launch_bounds__(128) k_dxt1_fit(const uint8_t *__restrict__ rgb, size_t frame_bytes, int w, int bw,
                                                   int nb, uint64_t *__restrict__ init_blocks,
                                                   uint64_t *__restrict__ final_blocks, int fstride) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  int f = blockIdx.y * fstride;
  if (b >= nb) return;
  uint32_t px[16];
  load_block_rgbx(rgb + frame_bytes * f, w, b % bw, b / bw, px);
  uint64_t blk = fit_block(px);
  init_blocks[(size_t)f * nb + b] = blk;
  final_blocks[(size_t)f * nb + b] = blk;
}

// ------------------------------------------------------------------------------------------
// K2: inter search.  One CTA per target block; threads stride over the (2sa)^2 window of the
// previous frame's FINAL index words.
// ------------------------------------------------------------------------------------------
constexpr int kSearchThreads = 256;

__global__ void __launch_bounds__(kSearchThreads)
k_inter_search(SeqView v, int k_in_gop, int sa, int thr) {
  __shared__ TargetCtx t;
  __shared__ WinnerState red[kSearchThreads / 32];
  const int b = blockIdx.x;
  const int f = v.first + blockIdx.y * v.gop + k_in_gop;
  if (f >= v.first + v.count) return;
  const int bx = b % v.bw, by = b / v.bw;
  const uint64_t *prev = v.final_blocks + (size_t)(f - 1) * v.nb;
  if (threadIdx.x == 0) build_target(t, v.rgb + v.frame_bytes * f, v.w, bx, by, v.init_blocks[(size_t)f * v.nb + b]);
  __syncthreads();

  const int W = 2 * sa;
  WinnerState s;
  winner_init(s);
  for (int p = threadIdx.x; p < W * W; p += kSearchThreads) {
    int row = p / W, col = p - row * W;
    int j = by - sa + row, i = bx - sa + col;
    if (i < 0 || j < 0 || i >= v.bw || j >= v.bh) continue;
    uint32_t word = (uint32_t)(__ldg(prev + (size_t)j * v.bw + i) >> 32);
    winner_update(s, eval_candidate(t, word), row, col, W);
  }
  winner_block_reduce<kSearchThreads / 32>(s, red);

  if (threadIdx.x == 0) {
    int row, col;
    int min_err = winner_resolve(s, W, row, col);
    uint8_t flag = 0;
    if (min_err <= thr) {
      uint32_t word = (uint32_t)(prev[(size_t)(by - sa + row) * v.bw + (bx - sa + col)] >> 32);
      v.final_blocks[(size_t)f * v.nb + b] = winning_block(t, word);
      v.motion[((size_t)f * v.nb + b) * 2 + 0] = (uint8_t)(col | 0x80);  // x = (i - bx) + sa
      v.motion[((size_t)f * v.nb + b) * 2 + 1] = (uint8_t)(row | 0x80);  // y = (j - by) + sa
      flag = 1;
    }
    v.flags[(size_t)f * v.nb + b] = flag;
    if (!flag) v.row_todo[(size_t)f * v.bh + by] = 1;
  }
}

// ------------------------------------------------------------------------------------------
// K3: intra search as a row-staggered wavefront.
//
// Block (x, y) reads the FINAL index words of (x-sa..x-1, y) and (x-sa..x+sa-1, y-1..y-2sa+1)
// (dxt_image.cpp:672-679), which Reencode overwrites as it goes (:926).  One CTA walks one
// block row left to right; row y may process block x once row y-1 has finished block
// min(x+sa, bw)-1.  progress[f][y] = number of leading blocks of row y that are final.
// Rows are handed out in increasing order by an atomic ticket, so every CTA a waiter depends
// on already holds a ticket and is resident: no deadlock for any grid size.
// Blocks with flags != 0 (found by the inter search) are already final and are skipped.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSearchThreads)
k_intra_wavefront(SeqView v, int k_in_gop, int n_gops, int sa, int thr, int *__restrict__ ticket) {
  __shared__ TargetCtx t;
  __shared__ WinnerState red[kSearchThreads / 32];
  __shared__ int s_item;
  const int W = 2 * sa;
  const int n_items = n_gops * v.bh;
  if (k_in_gop > 0) {   // inter frames: anything left that K3s did not take?
    bool any = false;
    for (int g = 0; g < n_gops; ++g) {
      const int f = v.first + g * v.gop + k_in_gop;
      any = any || (f < v.first + v.count && v.n_unique[f] == kSparseNotHandled);
    }
    if (!any) return;
  }
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(ticket, 1);
    __syncthreads();
    const int item = s_item;
    __syncthreads();
    if (item >= n_items) return;
    const int g = item % n_gops, by = item / n_gops;
    const int f = v.first + g * v.gop + k_in_gop;
    if (f >= v.first + v.count) continue;
    if (k_in_gop > 0 && v.n_unique[f] != kSparseNotHandled) continue;
    const uint8_t *frame = v.rgb + v.frame_bytes * f;
    uint64_t *cur = v.final_blocks + (size_t)f * v.nb;
    const uint8_t *flags = v.flags + (size_t)f * v.nb;
    int *progress = v.progress + (size_t)f * v.bh;
    int published = 0, seen_above = 0;  // thread 0 only

    for (int bx = 0; bx < v.bw; ++bx) {
      const int b = by * v.bw + bx;
      if (flags[b]) continue;  // uniform: already final (inter search hit)
      if (threadIdx.x == 0) {
        // Invariant: progress[y] = p implies progress[y-1] >= min(p-1+sa, bw), so that by
        // induction every row of the window is final -- also when blocks are skipped.
        if (by > 0) {
          const int need = min(bx + sa, v.bw);
          while (seen_above < need) {
            seen_above = ld_acquire(progress + by - 1);
            if (seen_above < need) __nanosleep(20);
          }
        }
        if (published < bx) { st_release(progress + by, bx); published = bx; }
        build_target(t, frame, v.w, bx, by, v.init_blocks[(size_t)f * v.nb + b]);
      }
      __syncthreads();

      WinnerState s;
      winner_init(s);
      for (int p = threadIdx.x; p < W * W; p += kSearchThreads) {
        int row = p / W, col = p - row * W;          // scan order: j downwards, i downwards
        int j = by - row, i = bx + sa - 1 - col;
        if (i < 0 || j < 0 || i >= v.bw || (row == 0 && i >= bx)) continue;
        uint32_t word = (uint32_t)(__ldcg(cur + (size_t)j * v.bw + i) >> 32);
        winner_update(s, eval_candidate(t, word), row, col, W);
      }
      winner_block_reduce<kSearchThreads / 32>(s, red);

      if (threadIdx.x == 0) {
        int row, col;
        int min_err = winner_resolve(s, W, row, col);
        size_t mo = ((size_t)f * v.nb + b) * 2;
        if (min_err <= thr) {
          uint32_t word = (uint32_t)(__ldcg(cur + (size_t)(by - row) * v.bw + (bx + sa - 1 - col)) >> 32);
          cur[b] = winning_block(t, word);
          v.motion[mo + 0] = (uint8_t)(2 * sa - 1 - col);  // x = (i - bx) + sa
          v.motion[mo + 1] = (uint8_t)(2 * sa - 1 - row);  // y = (j - by) + 2sa - 1
        } else {
          v.motion[mo + 0] = 255;
          v.motion[mo + 1] = 255;
        }
        __threadfence();
        published = bx + 1;
        st_release(progress + by, published);
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      if (by > 0)
        while (seen_above < v.bw) {
          seen_above = ld_acquire(progress + by - 1);
          if (seen_above < v.bw) __nanosleep(20);
        }
      st_release(progress + by, v.bw);
    }
  }
}

// ------------------------------------------------------------------------------------------
// K4: unique-palette compaction: ordered prefix sum over the "(255,255)" motion entries, emits
// the interp words in raster order.  Two passes over chunks of 1024 raster-ordered blocks (any
// number of CTAs per frame): count the unique blocks of every chunk, then every chunk emits its
// words behind the chunks before it.
// ------------------------------------------------------------------------------------------
constexpr int kCompactChunk = 1024;

__global__ void __launch_bounds__(kCompactChunk)
k_compact_count(SeqView v, int sa, unsigned long long *__restrict__ cand_counts, int f0, int fstride) {
  __shared__ int warp_cnt[kCompactChunk / 32];
  __shared__ unsigned long long warp_cand[2][kCompactChunk / 32];
  const int f = f0 + blockIdx.y * fstride;
  const int b = blockIdx.x * kCompactChunk + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool intra_frame = ((f - v.first) % v.gop) == 0;
  unsigned long long n_inter = 0, n_intra = 0;  // full-window candidate positions
  bool uniq = false;
  if (b < v.nb) {
    const uint16_t m = reinterpret_cast<const uint16_t *>(v.motion + (size_t)f * v.nb * 2)[b];
    uniq = (m == 0xFFFFu);
    const int bx = b % v.bw, by = b / v.bw;
    const bool inter_hit = !intra_frame && (m & 0x8080u) == 0x8080u && !uniq;
    if (!intra_frame)
      n_inter = (unsigned long long)(min(bx + sa, v.bw) - max(bx - sa, 0)) * (min(by + sa, v.bh) - max(by - sa, 0));
    if (!inter_hit)
      n_intra = (unsigned long long)min(by, 2 * sa - 1) * (min(bx + sa - 1, v.bw - 1) - max(bx - sa, 0) + 1) + min(bx, sa);
  }
  const unsigned bal = __ballot_sync(0xffffffffu, uniq);
  for (int d = 16; d > 0; d >>= 1) {
    n_inter += __shfl_xor_sync(0xffffffffu, n_inter, d);
    n_intra += __shfl_xor_sync(0xffffffffu, n_intra, d);
  }
  if (lane == 0) { warp_cnt[wid] = __popc(bal); warp_cand[0][wid] = n_inter; warp_cand[1][wid] = n_intra; }
  __syncthreads();
  if (wid == 0) {
    int c = warp_cnt[lane];
    unsigned long long ci = warp_cand[0][lane], ca = warp_cand[1][lane];
    for (int d = 16; d > 0; d >>= 1) {
      c += __shfl_xor_sync(0xffffffffu, c, d);
      ci += __shfl_xor_sync(0xffffffffu, ci, d);
      ca += __shfl_xor_sync(0xffffffffu, ca, d);
    }
    if (lane == 0) {
      v.chunk_counts[(size_t)f * gridDim.x + blockIdx.x] = (uint32_t)c;
      if (ci) atomicAdd(cand_counts + 0, ci);
      if (ca) atomicAdd(cand_counts + 1, ca);
    }
  }
}

// host_unique / host_n_unique (optional): the caller's page-locked result buffers, mapped into the
// device's address space.  The unique words then go straight to the host from here -- n_unique words
// per frame over PCIe instead of a copy of the whole nb-word slot (a 1080p frame has ~100 unique
// blocks of 129 600).  host_first = the sequence frame that host frame 0 corresponds to.
__global__ void __launch_bounds__(kCompactChunk)
k_compact_unique(SeqView v, int f0, int fstride, uint32_t *__restrict__ host_unique, uint32_t *__restrict__ host_n_unique,
                 int host_first) {
  __shared__ int warp_cnt[kCompactChunk / 32];
  __shared__ int chunk_base;
  const int f = f0 + blockIdx.y * fstride;
  const int b = blockIdx.x * kCompactChunk + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool uniq = b < v.nb && reinterpret_cast<const uint16_t *>(v.motion + (size_t)f * v.nb * 2)[b] == 0xFFFFu;
  const unsigned bal = __ballot_sync(0xffffffffu, uniq);
  if (lane == 0) warp_cnt[wid] = __popc(bal);
  if (wid == 0) {   // unique blocks of the chunks before this one
    int c = 0;
    for (int i = lane; i < (int)blockIdx.x; i += 32) c += (int)v.chunk_counts[(size_t)f * gridDim.x + i];
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (lane == 0) chunk_base = c;
  }
  __syncthreads();
  int before = chunk_base + __popc(bal & ((1u << lane) - 1u)), total = chunk_base;
  for (int k = 0; k < kCompactChunk / 32; ++k) {
    const int c = warp_cnt[k];
    if (k < wid) before += c;
    total += c;
  }
  if (uniq) {
    const uint32_t word = (uint32_t)(v.final_blocks[(size_t)f * v.nb + b] >> 32);
    v.unique[(size_t)f * v.nb + before] = word;
    if (host_unique) host_unique[(size_t)(f - host_first) * v.nb + before] = word;
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    v.n_unique[f] = (uint32_t)total;
    if (host_n_unique) host_n_unique[f - host_first] = (uint32_t)total;
  }
}

// ------------------------------------------------------------------------------------------
// K5: endpoint planes.  One CTA per (64x64 tile, frame) with all six planes of the tile in shared
// memory (6 x 2 x 8 KB, ping-pong per lifting direction): the blocks are read once, the colour
// transform, six lifting levels (columns then rows, wavelet.cpp:110-130) and the +128 symbol
// mapping happen on chip; 8 B/block in, 6 B/block out.
// ------------------------------------------------------------------------------------------
constexpr int kPlaneTile = 64, kPlaneElems = kPlaneTile * kPlaneTile;

__global__ void __launch_bounds__(1024)
k_endpoint_planes(SeqView v, int pbw, int pbh, int f0, int fstride) {
  extern __shared__ int16_t plane_sm[];
  int16_t *A = plane_sm, *B = plane_sm + 6 * kPlaneElems;   // [plane][y][x]
  const int tiles_x = pbw / kPlaneTile;
  const int tx = (blockIdx.x % tiles_x) * kPlaneTile, ty = (blockIdx.x / tiles_x) * kPlaneTile;
  const int f = f0 + blockIdx.y * fstride;
  const uint64_t *blocks = v.final_blocks + (size_t)f * v.nb;
  for (int e = threadIdx.x; e < kPlaneElems; e += 1024) {
    const int y = e >> 6, x = e & 63;
    const int sx = min(tx + x, v.bw - 1), sy = min(ty + y, v.bh - 1);  // edge replication (extension)
    const uint32_t eps = (uint32_t)blocks[(size_t)sy * v.bw + sx];
#pragma unroll
    for (int ep = 0; ep < 2; ++ep) {
      const uint32_t c = (eps >> (16 * ep)) & 0xFFFFu;
      const int r = (int)(c >> 11), g = (int)((c >> 5) & 63u), b = (int)(c & 31u);
      const int co = r - b, tt = r + b + (b >> 4), cg = g - tt, yy = tt + cg / 2;  // image_processing.cpp:10-27
      A[(3 * ep + 0) * kPlaneElems + e] = (int16_t)yy;
      A[(3 * ep + 1) * kPlaneElems + e] = (int16_t)co;
      A[(3 * ep + 2) * kPlaneElems + e] = (int16_t)cg;
    }
  }
  __syncthreads();
  for (int dim = kPlaneTile; dim > 1; dim >>= 1) {
    const int half = dim >> 1, items = 6 * half * dim;
    const int sh = 31 - __clz(half);                       // half is a power of two
    // columns (wavelet.cpp:110-119): predict the odd rows into the lower half of B ...
    for (int e = threadIdx.x; e < items; e += 1024) {
      const int c = e & (dim - 1), k = (e >> (sh + 1)) & (half - 1), pl = e >> (2 * sh + 1);
      const int16_t *a = A + pl * kPlaneElems + c;
      const int i = 2 * k + 1, nx = i + 1 < dim ? i + 1 : dim - 2;   // mirror at the end
      B[pl * kPlaneElems + (half + k) * kPlaneTile + c] = (int16_t)(a[i * kPlaneTile] - (a[(i - 1) * kPlaneTile] + a[nx * kPlaneTile]) / 2);
    }
    __syncthreads();
    // ... update the even rows into the upper half
    for (int e = threadIdx.x; e < items; e += 1024) {
      const int c = e & (dim - 1), k = (e >> (sh + 1)) & (half - 1), pl = e >> (2 * sh + 1);
      const int16_t *d = B + pl * kPlaneElems + half * kPlaneTile + c;
      B[pl * kPlaneElems + k * kPlaneTile + c] =
          (int16_t)(A[pl * kPlaneElems + 2 * k * kPlaneTile + c] + (d[max(k - 1, 0) * kPlaneTile] + d[k * kPlaneTile] + 2) / 4);
    }
    __syncthreads();
    // rows (:121-130), from B back into A: predict the odd columns into the right half ...
    for (int e = threadIdx.x; e < items; e += 1024) {
      const int k = e & (half - 1), r = (e >> sh) & (dim - 1), pl = e >> (2 * sh + 1);
      const int16_t *b = B + pl * kPlaneElems + r * kPlaneTile;
      const int i = 2 * k + 1, nx = i + 1 < dim ? i + 1 : dim - 2;
      A[pl * kPlaneElems + r * kPlaneTile + half + k] = (int16_t)(b[i] - (b[i - 1] + b[nx]) / 2);
    }
    __syncthreads();
    // ... update the even columns into the left half
    for (int e = threadIdx.x; e < items; e += 1024) {
      const int k = e & (half - 1), r = (e >> sh) & (dim - 1), pl = e >> (2 * sh + 1);
      int16_t *a = A + pl * kPlaneElems + r * kPlaneTile;
      a[k] = (int16_t)(B[pl * kPlaneElems + r * kPlaneTile + 2 * k] + (a[half + max(k - 1, 0)] + a[half + k] + 2) / 4);
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < 6 * kPlaneElems / 4; e += 1024) {
    const int pl = e >> 10, idx = (e & 1023) * 4, y = idx >> 6, x4 = idx & 63;
    uint32_t w = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) w |= (uint32_t)((uint8_t)((int8_t)A[pl * kPlaneElems + idx + q] + 128)) << (8 * q);
    uint8_t *out = v.planes + ((size_t)f * 6 + pl) * (size_t)pbw * pbh;
    *reinterpret_cast<uint32_t *>(out + (size_t)(ty + y) * pbw + tx + x4) = w;
  }
}

// ------------------------------------------------------------------------------------------
// Launch wrappers
// ------------------------------------------------------------------------------------------
// K1/K4/K5 run over frames f0, f0 + fstride, ... (nf of them): a contiguous range (fstride 1) or
// frame k of every GOP of a lane (fstride = gop).
void launch_dxt1_fit(const SeqView &v, int f0, int fstride, int nf, cudaStream_t s) {
  dim3 grid((v.nb + 127) / 128, nf);
  k_dxt1_fit<<<grid, 128, 0, s>>>(v.rgb + v.frame_bytes * f0, v.frame_bytes, v.w, v.bw, v.nb,
                                  v.init_blocks + (size_t)f0 * v.nb, v.final_blocks + (size_t)f0 * v.nb, fstride);
}

// K2 dispatch: the 16x16-target kernel (mptc_inter_wide.cu) while two of its CTAs fit an SM, else the
// 8x4-target kernel of round 1 (mptc_inter.cu), else the wide kernel at one CTA per SM, else one CTA per
// target.  MPTC_K2 = wide | tiled forces the choice (A/B measurements, and the parity tests of both).
int launch_inter_search(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, cudaStream_t s) {
  static const int forced = [] {
    const char *e = getenv("MPTC_K2");
    return !e ? 0 : (!strcmp(e, "wide") ? 1 : (!strcmp(e, "tiled") ? 2 : 0));
  }();
  if (forced != 2 && launch_inter_search_wide(v, k_in_gop, n_gops, sa, thr, /*two_per_sm_only=*/forced == 0, s)) return 1;
  if (launch_inter_search_tiled(v, k_in_gop, n_gops, sa, thr, s)) return 1;
  if (forced == 0 && launch_inter_search_wide(v, k_in_gop, n_gops, sa, thr, false, s)) return 1;
  dim3 grid(v.nb, n_gops);  // direct (one CTA per target) fallback for very large windows
  k_inter_search<<<grid, kSearchThreads, 0, s>>>(v, k_in_gop, sa, thr);
  return 1;
}

void launch_intra_wavefront(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, int *ticket,
                            int max_ctas, int grid_cap, cudaStream_t s) {
  if (launch_intra_rows(v, k_in_gop, n_gops, sa, thr, ticket, grid_cap, s)) return;
  int items = n_gops * v.bh;  // direct (one target at a time) fallback for very large windows
  int grid = items < max_ctas ? items : max_ctas;
  if (grid_cap > 0 && grid > grid_cap) grid = grid_cap;
  k_intra_wavefront<<<grid, kSearchThreads, 0, s>>>(v, k_in_gop, n_gops, sa, thr, ticket);
}

void launch_compact_unique(const SeqView &v, int sa, unsigned long long *cand_counts, int f0, int fstride, int nf,
                           cudaStream_t s, uint32_t *host_unique, uint32_t *host_n_unique, int host_first) {
  dim3 grid((v.nb + kCompactChunk - 1) / kCompactChunk, nf);
  k_compact_count<<<grid, kCompactChunk, 0, s>>>(v, sa, cand_counts, f0, fstride);
  k_compact_unique<<<grid, kCompactChunk, 0, s>>>(v, f0, fstride, host_unique, host_n_unique, host_first);
}

std::mutex &launch_cfg_mutex() {
  static std::mutex m;
  return m;
}

void launch_endpoint_planes(const SeqView &v, int pbw, int pbh, int f0, int fstride, int nf, cudaStream_t s) {
  static bool configured[kMaxDevices] = {false};   // per device: one context per GPU may live in one process
  const int bytes = 12 * kPlaneElems * (int)sizeof(int16_t);
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  {
    std::lock_guard<std::mutex> lock(launch_cfg_mutex());
    if (!configured[cur_dev & (kMaxDevices - 1)]) {
      cudaFuncSetAttribute(k_endpoint_planes, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      configured[cur_dev & (kMaxDevices - 1)] = true;
    }
  }
  dim3 grid((pbw / kPlaneTile) * (pbh / kPlaneTile), nf);
  k_endpoint_planes<<<grid, 1024, bytes, s>>>(v, pbw, pbh, f0, fstride);
}

int intra_wavefront_max_ctas(int device) {
  int per_sm = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_intra_wavefront, kSearchThreads, 0);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (per_sm < 1) per_sm = 1;
  return per_sm * sms;
}

}  // namespace mptc
