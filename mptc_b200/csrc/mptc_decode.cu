// mptc_decode.cu -- decoder-side kernels for sm_100a: from the symbols the
// arithmetic decoder produced to ready-to-upload DXT1 blocks (and, optionally, RGB pixels).
//
//   D1 k_dec_count / k_dec_links / k_dec_jump   ReconstructDXTData / ReconstructDXTFrame
//        (codec/codec.cpp:393-500, codec/decoder.cpp:198-254): every block's index word is a unique
//        word, a copy from the previous frame (inter) or a copy from an earlier block of its own
//        frame (intra).  The reference resolves that sequentially, block after block, frame after
//        frame.  It is pure copying, so here every block of a GOP becomes a node with one parent
//        link and the chains are collapsed by pointer jumping: O(log depth) data-parallel passes
//        over ALL frames of the GOP at once, no raster or frame order left.
//   D2 k_inverse_planes    ReconstructEndPoints (codec.cpp:697-800; decoder.cpp:68-196): symbols ->
//        MakeSigned -> inverse 5/3 wavelet on 64x64 tiles (image_processing.h:337-399, levels
//        dim = 2..64, rows then columns wavelet.cpp:133-155, :64-95) -> ycocg667_to_rgb565
//        (codec.cpp:46-63) -> 565 packing (:764-771), fused with the gather of the index word, so
//        the 8-byte PhysicalDXTBlock is written once, coalesced.
//   D3 k_dxt1_to_rgb       DXTImage::DecompressedImage (dxt_image.cpp:463-479) over
//        PhysicalToLogical (:198-227): the decoded picture, 8 B in / 48 B out per block -- the one
//        HBM-bound kernel of the codec.
//
// No tensor cores: integer copies and lifting steps only.
#include "mptc_kernels.h"
#include "mptc_device.cuh"

#include <cstdlib>

namespace mptc {

namespace {
constexpr int kChunk = 1024;   // blocks per CTA of D1's count / link kernels
}

int dec_chunks(int nb) { return (nb + kChunk - 1) / kChunk; }

// Unique blocks per chunk of 1024 raster-ordered blocks.
__global__ void __launch_bounds__(kChunk)
k_dec_count(DecView v) {
  __shared__ int warp_cnt[kChunk / 32];
  const int f = v.first + blockIdx.y;
  const int b = blockIdx.x * kChunk + threadIdx.x;
  bool uniq = false;
  if (b < v.nb) {
    const uchar2 m = reinterpret_cast<const uchar2 *>(v.motion)[(size_t)f * v.nb + b];
    uniq = m.x == 255 && m.y == 255;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, uniq);
  if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = __popc(bal);
  __syncthreads();
  if (threadIdx.x < 32) {
    int c = warp_cnt[threadIdx.x];
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (threadIdx.x == 0) v.chunk_counts[(size_t)f * gridDim.x + blockIdx.x] = (uint32_t)c;
  }
}

// One node per block: link = the node its index word is copied from (itself for unique blocks,
// whose word is fetched from the palette right here).
__global__ void __launch_bounds__(kChunk)
k_dec_links(DecView v) {
  __shared__ int warp_cnt[kChunk / 32];
  __shared__ int chunk_base;
  const int f = v.first + blockIdx.y;
  const int b = blockIdx.x * kChunk + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uchar2 m = make_uchar2(0, 0);
  if (b < v.nb) m = reinterpret_cast<const uchar2 *>(v.motion)[(size_t)f * v.nb + b];
  const bool uniq = b < v.nb && m.x == 255 && m.y == 255;
  const unsigned bal = __ballot_sync(0xffffffffu, uniq);
  if (lane == 0) warp_cnt[warp] = __popc(bal);
  if (warp == 0) {   // uniques of the chunks before this one
    int c = 0;
    for (int i = lane; i < (int)blockIdx.x; i += 32) c += (int)v.chunk_counts[(size_t)f * gridDim.x + i];
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) chunk_base = c;
  }
  __syncthreads();
  if (b >= v.nb) return;
  int rank = chunk_base + __popc(bal & ((1u << lane) - 1u));
  for (int i = 0; i < warp; ++i) rank += warp_cnt[i];
  const size_t n = (size_t)f * v.nb + b;
  const int bx = b % v.bw, by = b / v.bw;
  int parent = -1;
  if (uniq) {
    if ((uint32_t)rank < v.n_unique[f]) {                      // codec.cpp:451-455
      v.words[n] = v.unique[(size_t)v.unique_off[f] + rank];
      parent = (int)n;
    }
  } else if ((m.x & 0x80) && (m.y & 0x80)) {                   // inter (:456-471)
    const int rx = bx + (m.x & 0x7F) - v.sa, ry = by + (m.y & 0x7F) - v.sa;
    if ((f - v.first) % v.gop != 0 && rx >= 0 && ry >= 0 && rx < v.bw && ry < v.bh)
      parent = (int)(n - v.nb) - b + ry * v.bw + rx;
  } else {                                                     // intra (:472-489)
    const int rx = bx + m.x - v.sa, ry = by + m.y - (2 * v.sa - 1);
    if (rx >= 0 && ry >= 0 && rx < v.bw && ry * v.bw + rx < b)
      parent = (int)n - b + ry * v.bw + rx;
  }
  if (parent < 0) {           // a vector the encoder cannot emit: corrupt stream
    atomicAdd(v.errors, 1);
    v.words[n] = 0;
    parent = (int)n;
  }
  v.link[n] = parent;
}

// One pointer-jumping pass, in place: every node walks up to kHops links towards its root and
// stores where it got to.  Any value a racing thread reads is an ancestor of its node, so stale
// reads only cost passes, never correctness.  status[1+pass] is raised while some node still has
// not reached a root; later passes return at once when the previous one did not raise it.
constexpr int kHops = 16;

__global__ void __launch_bounds__(256)
k_dec_jump(DecView v, int pass) {
  if (pass > 0 && v.status[pass] == 0) return;
  const size_t total = (size_t)v.count * v.nb, base = (size_t)v.first * v.nb;
  bool more = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = base + i;
    const int p0 = __ldcg(v.link + n);
    int p = p0, q = __ldcg(v.link + p);
    for (int hop = 1; q != p && hop < kHops; ++hop) {
      p = q;
      q = __ldcg(v.link + p);
    }
    if (q != p0) v.link[n] = q;
    more |= q != p && __ldcg(v.link + q) != q;
  }
  if (__any_sync(0xffffffffu, more) && (threadIdx.x & 31) == 0) v.status[1 + pass] = 1;
}

// ------------------------------------------------------------------------------------------
// D2: inverse endpoint planes + final block assembly.  One CTA per (64x64 tile, frame), all six
// planes of the tile in shared memory (6 x 2 x 8 KB), ping-pong per lifting direction.
// ------------------------------------------------------------------------------------------
constexpr int kTile = 64, kTileElems = kTile * kTile;

__device__ __forceinline__ uint32_t ycocg_to_565(int y16, int co16, int cg16) {
  // codec.cpp:46-63 and :764-771 as written (int8 arithmetic, sign-extending uint16 casts)
  const int8_t y = (int8_t)y16, co = (int8_t)co16, cg = (int8_t)cg16;
  const int8_t t = (int8_t)(y - cg / 2);
  const int8_t g = (int8_t)(cg + t), b = (int8_t)((t - co) / 2), r = (int8_t)(b + co);
  uint16_t x = (uint16_t)r;
  x = (uint16_t)(x << 6); x |= (uint16_t)g;
  x = (uint16_t)(x << 5); x |= (uint16_t)b;
  return x;
}

__global__ void __launch_bounds__(1024)
k_inverse_planes(DecView v) {
  extern __shared__ int16_t sm[];
  int16_t *A = sm, *B = sm + 6 * kTileElems;     // [plane][y][x]
  const int tiles_x = v.pbw / kTile;
  const int tx = (blockIdx.x % tiles_x) * kTile, ty = (blockIdx.x / tiles_x) * kTile;
  const int f = v.first + blockIdx.y;
  const size_t pn = (size_t)v.pbw * v.pbh;
  // MakeSigned (image_utils.h:243-266): symbol - 128 as int8
  for (int e = threadIdx.x; e < 6 * kTileElems / 4; e += 1024) {
    const int pl = e >> 10, idx = (e & 1023) * 4, y = idx >> 6, x = idx & 63;
    const uint32_t s4 = *reinterpret_cast<const uint32_t *>(v.planes + ((size_t)f * 6 + pl) * pn + (size_t)(ty + y) * v.pbw + tx + x);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      A[pl * kTileElems + idx + q] = (int16_t)(int8_t)(uint8_t)(((s4 >> (8 * q)) & 0xFFu) - 128u);
  }
  __syncthreads();
  for (int dim = 2; dim <= kTile; dim <<= 1) {
    const int half = dim >> 1, items = 6 * half * dim;
    const int sh = 31 - __clz(half);                       // half is a power of two
    // rows (wavelet.cpp:141-144): even samples ...
    for (int e = threadIdx.x; e < items; e += 1024) {
      const int k = e & (half - 1), r = (e >> sh) & (dim - 1), pl = e >> (2 * sh + 1);
      const int16_t *a = A + pl * kTileElems + r * kTile;
      B[pl * kTileElems + r * kTile + 2 * k] = (int16_t)(a[k] - (a[half + max(k - 1, 0)] + a[half + k] + 2) / 4);
    }
    __syncthreads();
    // ... odd samples
    for (int e = threadIdx.x; e < items; e += 1024) {
      const int k = e & (half - 1), r = (e >> sh) & (dim - 1), pl = e >> (2 * sh + 1);
      int16_t *b = B + pl * kTileElems + r * kTile;
      b[2 * k + 1] = (int16_t)(A[pl * kTileElems + r * kTile + half + k] + (b[2 * k] + b[min(2 * k + 2, dim - 2)]) / 2);
    }
    __syncthreads();
    // columns (:148-152): even samples ...
    for (int e = threadIdx.x; e < items; e += 1024) {
      const int c = e & (dim - 1), k = (e >> (sh + 1)) & (half - 1), pl = e >> (2 * sh + 1);
      const int16_t *b = B + pl * kTileElems + c;
      A[pl * kTileElems + 2 * k * kTile + c] =
          (int16_t)(b[k * kTile] - (b[(half + max(k - 1, 0)) * kTile] + b[(half + k) * kTile] + 2) / 4);
    }
    __syncthreads();
    // ... odd samples
    for (int e = threadIdx.x; e < items; e += 1024) {
      const int c = e & (dim - 1), k = (e >> (sh + 1)) & (half - 1), pl = e >> (2 * sh + 1);
      int16_t *a = A + pl * kTileElems + c;
      a[(2 * k + 1) * kTile] =
          (int16_t)(B[pl * kTileElems + (half + k) * kTile + c] + (a[2 * k * kTile] + a[min(2 * k + 2, dim - 2) * kTile]) / 2);
    }
    __syncthreads();
  }
  // endpoints + the index word gathered through the collapsed link -> PhysicalDXTBlock
  for (int e = threadIdx.x; e < kTileElems; e += 1024) {
    const int y = e >> 6, x = e & 63;
    if (ty + y >= v.bh || tx + x >= v.bw) continue;
    const uint32_t ep1 = ycocg_to_565(A[e], A[kTileElems + e], A[2 * kTileElems + e]);
    const uint32_t ep2 = ycocg_to_565(A[3 * kTileElems + e], A[4 * kTileElems + e], A[5 * kTileElems + e]);
    const size_t n = (size_t)f * v.nb + (size_t)(ty + y) * v.bw + tx + x;
    const uint32_t word = v.words[v.link[n]];
    v.blocks[n] = (uint64_t)(ep1 | (ep2 << 16)) | ((uint64_t)word << 32);
  }
}

// ------------------------------------------------------------------------------------------
// D3: DXT1 blocks -> RGB8 rows.  A CTA takes a run of 128 blocks of a block row (the loop lets a
// capped grid walk several): palettes in registers, the four pixel rows of the run staged in
// shared memory (double buffered), written
// back as coalesced 16-byte words when the frame rows allow it (a row of a run is 1536 contiguous
// bytes of the frame), 4-byte words otherwise.
// ------------------------------------------------------------------------------------------
constexpr int kRun = 128;

// Store modes of the staged rows: 0 = 4-byte words (any width), 1 = 16-byte words, 2 = one bulk
// asynchronous copy per pixel row (cp.async.bulk shared -> global, issued by one thread: the copy
// engine streams the 1536-byte row while the CTA retires).  1 and 2 need w % 16 == 0.
template <int MODE>
__global__ void __launch_bounds__(kRun)
k_dxt1_to_rgb(DecView v, int runs_x, int n_runs) {
  __shared__ __align__(128) uint32_t rows[2][4][kRun * 3];
  int buf = 0;
  for (int run = blockIdx.x; run < n_runs; run += gridDim.x, buf ^= 1) {
    const int fi = run / (runs_x * v.bh), rr = run - fi * runs_x * v.bh;
    const int by = rr / runs_x, bx0 = (rr - by * runs_x) * kRun;
    const int f = v.first + fi;
    const int bx = bx0 + threadIdx.x;
    const int nrun = min(kRun, v.bw - bx0);
    if (bx < v.bw) {
      const uint64_t blk = __ldcs(reinterpret_cast<const unsigned long long *>(v.blocks) + (size_t)f * v.nb + (size_t)by * v.bw + bx);
      uint32_t pal[4];
      palette_of_block(blk, pal);
      const uint32_t word = (uint32_t)(blk >> 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t c[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t s = (word >> (2 * (4 * j + i))) & 3u;
          c[i] = (s & 2u) ? ((s & 1u) ? pal[3] : pal[2]) : ((s & 1u) ? pal[1] : pal[0]);
        }
        // four RGBX pixels -> 12 bytes R G B R | G B R G | B R G B
        rows[buf][j][3 * threadIdx.x + 0] = c[0] | (c[1] << 24);
        rows[buf][j][3 * threadIdx.x + 1] = (c[1] >> 8) | (c[2] << 16);
        rows[buf][j][3 * threadIdx.x + 2] = (c[2] >> 16) | (c[3] << 8);
      }
    }
    if (MODE == 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the rows, for the copy engine
    __syncthreads();   // one barrier per run: the next run fills the other buffer
    uint8_t *frame = v.rgb + (size_t)f * v.w * v.h * 3;
    if (MODE == 2) {   // every run row is a 16-byte aligned multiple of 16 bytes, contiguous in the frame
      if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint8_t *dst = frame + ((size_t)(4 * by + j) * v.w + 4 * bx0) * 3;
          const uint32_t src = (uint32_t)__cvta_generic_to_shared(rows[buf][j]);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(12 * nrun) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        // the buffer two runs back (and, before the CTA retires, this one) must have been read
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    } else if (MODE == 1) {   // 3*nrun/4 uint4 per row
      const int per_row = 3 * nrun / 4;
      for (int i = threadIdx.x; i < 4 * per_row; i += kRun) {
        const int j = i / per_row, c = i - j * per_row;
        uint4 *dst = reinterpret_cast<uint4 *>(frame + ((size_t)(4 * by + j) * v.w + 4 * bx0) * 3);
        __stcs(dst + c, reinterpret_cast<const uint4 *>(rows[buf][j])[c]);
      }
    } else {
      for (int j = 0; j < 4; ++j) {
        uint32_t *dst = reinterpret_cast<uint32_t *>(frame + ((size_t)(4 * by + j) * v.w + 4 * bx0) * 3);
        for (int i = threadIdx.x; i < 3 * nrun; i += kRun) dst[i] = rows[buf][j][i];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Launch wrappers
// ------------------------------------------------------------------------------------------
int dec_jump_passes(int gop, int nb) {
  // a chain is at most gop * nb links long and every pass shortens the longest one kHops times
  const long long depth = (long long)gop * nb;
  int p = 1;
  for (long long reach = kHops; reach < depth; reach *= kHops) ++p;
  return p + 1;
}

int launch_decode_words(const DecView &v, cudaStream_t s) {
  dim3 grid(dec_chunks(v.nb), v.count);
  k_dec_count<<<grid, kChunk, 0, s>>>(v);
  k_dec_links<<<grid, kChunk, 0, s>>>(v);
  const int passes = dec_jump_passes(v.gop, v.nb);
  const size_t total = (size_t)v.count * v.nb;
  int ctas = (int)((total + 1023) / 1024);
  if (ctas > 148 * 8) ctas = 148 * 8;
  for (int p = 0; p < passes; ++p) k_dec_jump<<<ctas, 256, 0, s>>>(v, p);
  return 2 + passes;
}

cudaError_t decode_kernels_init() {
  return cudaFuncSetAttribute(k_inverse_planes, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)(12 * kTileElems * sizeof(int16_t)));
}

int launch_inverse_planes(const DecView &v, cudaStream_t s) {
  dim3 grid((v.pbw / kTile) * (v.pbh / kTile), v.count);
  k_inverse_planes<<<grid, 1024, 12 * kTileElems * sizeof(int16_t), s>>>(v);
  return 1;
}

int launch_dxt1_to_rgb(const DecView &v, cudaStream_t s) {
  const int runs_x = (v.bw + kRun - 1) / kRun;
  const long long n_runs = (long long)runs_x * v.bh * v.count;
  // one run per CTA: measured 3.2 TB/s against 2.95 TB/s for 16 persistent CTAs per SM looping over
  // runs and 3.1 TB/s with 4-byte stores (profiles/micro/rgb_ab.py).  Round 2: the rows leave through
  // cp.async.bulk (UBLKCP in the SASS) instead of 16-byte stores: 3.27 -> 4.26 TB/s = 0.65 of the
  // measured copy peak; 86 % of the traffic are writes, 3.65 TB/s of them against the 3.86 TB/s a
  // write-only fill reaches on the same box (profiles/micro/write_bw.py) -- 94 % of the write ceiling.
  const int ctas = (int)(n_runs < 0x7fffffff ? n_runs : 0x7fffffff);
  static const int mode16 = [] { const char *e = getenv("MPTC_RGB_STORE"); return (e && e[0] == 'v') ? 1 : 2; }();   // v = 16-byte words
  if (v.w % 16 != 0) k_dxt1_to_rgb<0><<<ctas, kRun, 0, s>>>(v, runs_x, (int)n_runs);
  else if (mode16 == 2) k_dxt1_to_rgb<2><<<ctas, kRun, 0, s>>>(v, runs_x, (int)n_runs);
  else k_dxt1_to_rgb<1><<<ctas, kRun, 0, s>>>(v, runs_x, (int)n_runs);
  return 1;
}

}  // namespace mptc
