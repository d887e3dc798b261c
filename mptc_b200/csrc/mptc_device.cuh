// mptc_device.cuh -- device-side building blocks shared by the sm_100a kernels.
//
// Bit-exactness rules: every FP32 operation of the reference is
// individually rounded (x86-64 SSE, no FMA contraction), so all float math here goes through
// __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn, which nvcc never contracts into FFMA.  Float->int
// casts follow cvttss2si (NaN / out of range -> INT_MIN).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda/atomic>

namespace mptc {

constexpr int kRejected = 0x7fffffff;  // candidate not accepted (dxt_image.cpp:753-755)

// ---- RGB565 helpers (codec/dxt_image.cpp:43-69) -------------------------------------
__device__ __forceinline__ uint32_t expand565_rgbx(uint32_t v) {
  uint32_t r = v >> 11, g = (v >> 5) & 63u, b = v & 31u;
  r = (r << 3) | (r >> 2);
  g = (g << 2) | (g >> 4);
  b = (b << 3) | (b >> 2);
  return r | (g << 8) | (b << 16);
}

__device__ __forceinline__ uint32_t pack565_rgbx(uint32_t c) {
  return ((c & 0xF8u) << 8) | (((c >> 8) & 0xFCu) << 3) | (((c >> 16) & 0xFFu) >> 3);
}

// per-byte (2a+b)/3 and (a+2b)/3 on packed RGBX (LerpChannels, dxt_image.cpp:34-41)
__device__ __forceinline__ uint32_t lerp_bytes(uint32_t a, uint32_t b, int wa, int wb, int div) {
  uint32_t out = 0;
#pragma unroll
  for (int s = 0; s < 24; s += 8) {
    int x = (int)((a >> s) & 0xFF), y = (int)((b >> s) & 0xFF);
    out |= (uint32_t)((wa * x + wb * y) / div) << s;
  }
  return out;
}

// Palette of a physical block (PhysicalToLogical, dxt_image.cpp:198-214), packed RGBX.
__device__ __forceinline__ void palette_of_block(uint64_t blk, uint32_t pal[4]) {
  uint32_t e1 = (uint32_t)(blk & 0xFFFF), e2 = (uint32_t)((blk >> 16) & 0xFFFF);
  pal[0] = expand565_rgbx(e1);
  pal[1] = expand565_rgbx(e2);
  if (e1 <= e2) {
    pal[2] = lerp_bytes(pal[0], pal[1], 1, 1, 2);
    pal[3] = 0;
  } else {
    pal[2] = lerp_bytes(pal[0], pal[1], 2, 1, 3);
    pal[3] = lerp_bytes(pal[0], pal[1], 1, 2, 3);
  }
}

// CompressedBlock::Error (dxt_image.cpp:258-276): sum of squared byte differences / 48.
// px[k] are RGBX-packed pixels (X = 0), pal[v] RGBX-packed palette entries (X = 0).
__device__ __forceinline__ int block_error(const uint32_t *px, const uint32_t pal[4], uint32_t word) {
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    uint32_t v = (word >> (2 * k)) & 3u;
    uint32_t c = (v & 2u) ? ((v & 1u) ? pal[3] : pal[2]) : ((v & 1u) ? pal[1] : pal[0]);
    uint32_t d = __vabsdiffu4(px[k], c);
    sum = __dp4a(d, d, sum);
  }
  return (int)(sum / 48u);
}

// ToFiveBits / ToSixBits exactly as written (dxt_image.cpp:72-121): neighbours at +-4 / +-2.
template <int KEEP, int STEP, int SHIFT>
__device__ __forceinline__ int snap_bits(int x) {
  int base = x & KEEP;
  int high = (base + STEP) & 0xFF;  // base is never 255
  int low = base == 0 ? 0 : base - STEP;
  base |= base >> SHIFT;
  high |= high >> SHIFT;
  low |= low >> SHIFT;
  int db = abs(x - base), dh = abs(x - high), dl = abs(x - low);
  return db <= dh ? (db < dl ? base : low) : high;
}

// (int32)(p + 0.5f) then clamp to 0..255 with cvttss2si semantics (dxt_image.cpp:330-331):
// NaN, +-inf and anything outside int32 become INT_MIN, which clamps to 0.
__device__ __forceinline__ int quantise_endpoint(float p) {
  float x = __fadd_rn(p, 0.5f);
  int v = __float2int_rz(x);
  v = min(max(v, 0), 255);
  return (x < 2147483648.0f) ? v : 0;  // false for NaN as well
}

// CompressedBlock::RecalculateEndpoints (dxt_image.cpp:290-351) for index word `word` over
// the 16 pixels pf[k*3+ch] (already converted to float).  Returns the snapped 8-bit
// endpoints packed RGBX.  Summation order and rounding follow the reference exactly.
__device__ __forceinline__ void refit_endpoints(const float *pf, uint32_t word, uint32_t &ep1, uint32_t &ep2) {
  const float w23 = 2.0f / 3.0f, w13 = 1.0f / 3.0f;  // (3-order)/3, order/3 in FP32
  float asq = 0.f, bsq = 0.f, ab = 0.f;
  float ax0 = 0.f, ax1 = 0.f, ax2 = 0.f, bx0 = 0.f, bx1 = 0.f, bx2 = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    uint32_t v = (word >> (2 * k)) & 3u;
    // idx_to_order = {0,3,1,2}: a = {1, 0, 2/3, 1/3}[v], b = {0, 1, 1/3, 2/3}[v]
    float a = (v & 2u) ? ((v & 1u) ? w13 : w23) : ((v & 1u) ? 0.0f : 1.0f);
    float b = (v & 2u) ? ((v & 1u) ? w23 : w13) : ((v & 1u) ? 1.0f : 0.0f);
    asq = __fadd_rn(asq, __fmul_rn(a, a));
    bsq = __fadd_rn(bsq, __fmul_rn(b, b));
    ab = __fadd_rn(ab, __fmul_rn(a, b));
    float p0 = pf[3 * k + 0], p1 = pf[3 * k + 1], p2 = pf[3 * k + 2];
    ax0 = __fadd_rn(ax0, __fmul_rn(p0, a));
    bx0 = __fadd_rn(bx0, __fmul_rn(p0, b));
    ax1 = __fadd_rn(ax1, __fmul_rn(p1, a));
    bx1 = __fadd_rn(bx1, __fmul_rn(p1, b));
    ax2 = __fadd_rn(ax2, __fmul_rn(p2, a));
    bx2 = __fadd_rn(bx2, __fmul_rn(p2, b));
  }
  float f = __fdiv_rn(1.0f, __fsub_rn(__fmul_rn(asq, bsq), __fmul_rn(ab, ab)));
  int r1 = quantise_endpoint(__fmul_rn(f, __fsub_rn(__fmul_rn(ax0, bsq), __fmul_rn(bx0, ab))));
  int r2 = quantise_endpoint(__fmul_rn(f, __fsub_rn(__fmul_rn(bx0, asq), __fmul_rn(ax0, ab))));
  int g1 = quantise_endpoint(__fmul_rn(f, __fsub_rn(__fmul_rn(ax1, bsq), __fmul_rn(bx1, ab))));
  int g2 = quantise_endpoint(__fmul_rn(f, __fsub_rn(__fmul_rn(bx1, asq), __fmul_rn(ax1, ab))));
  int b1 = quantise_endpoint(__fmul_rn(f, __fsub_rn(__fmul_rn(ax2, bsq), __fmul_rn(bx2, ab))));
  int b2 = quantise_endpoint(__fmul_rn(f, __fsub_rn(__fmul_rn(bx2, asq), __fmul_rn(ax2, ab))));
  r1 = snap_bits<0xF8, 4, 5>(r1);  r2 = snap_bits<0xF8, 4, 5>(r2);
  g1 = snap_bits<0xFC, 2, 6>(g1);  g2 = snap_bits<0xFC, 2, 6>(g2);
  b1 = snap_bits<0xF8, 4, 5>(b1);  b2 = snap_bits<0xF8, 4, 5>(b2);
  ep1 = (uint32_t)r1 | ((uint32_t)g1 << 8) | ((uint32_t)b1 << 16);
  ep2 = (uint32_t)r2 | ((uint32_t)g2 << 8) | ((uint32_t)b2 << 16);
}

// Loads the 4x4 block (bx, by) of an RGB8 frame as 16 RGBX words.  Each thread reads 4 rows
// x 12 bytes as three aligned 32-bit words; consecutive threads read consecutive 12-byte
// runs, so a warp covers 384 contiguous bytes per row.
__device__ __forceinline__ void load_block_rgbx(const uint8_t *frame, int w, int bx, int by, uint32_t *px) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t *row = reinterpret_cast<const uint32_t *>(frame + ((size_t)(by * 4 + j) * w + bx * 4) * 3);
    uint32_t a = __ldg(row), b = __ldg(row + 1), c = __ldg(row + 2);
    px[4 * j + 0] = a & 0x00FFFFFFu;
    px[4 * j + 1] = (a >> 24) | ((b & 0xFFFFu) << 8);
    px[4 * j + 2] = (b >> 16) | ((c & 0xFFu) << 16);
    px[4 * j + 3] = c >> 8;
  }
}

// Per-target-block context, built once per target and shared by every candidate evaluation.
struct TargetCtx {
  float pf[48];        // pixels as float, [k*3 + ch]      (Get4X4ColorsBlock, dxt_image.cpp:636-650)
  uint32_t px[16];     // pixels RGBX-packed
  uint64_t own_block;  // the block's initial stb fit
  uint32_t own_word;
  int orig_err;        // blk.Error() with the initial logical block (dxt_image.cpp:668, :728)
};

// One candidate evaluation (dxt_image.cpp:739-758 / :679-698).  Returns err_diff, or
// kRejected when the refit would need an endpoint swap.
__device__ __forceinline__ int eval_candidate(const TargetCtx &t, uint32_t word) {
  if (word == t.own_word) return 0;  // blk == blk2: palette untouched
  uint32_t ep1, ep2;
  refit_endpoints(t.pf, word, ep1, ep2);
  if (!(pack565_rgbx(ep1) > pack565_rgbx(ep2))) return kRejected;
  uint32_t pal[4] = {ep1, ep2, lerp_bytes(ep1, ep2, 2, 1, 3), lerp_bytes(ep1, ep2, 1, 2, 3)};
  return block_error(t.px, pal, word) - t.orig_err;
}

// The block emitted when candidate `word` wins (dxt_image.cpp:900-905 / :922-926).
__device__ __forceinline__ uint64_t winning_block(const TargetCtx &t, uint32_t word) {
  if (word == t.own_word) return t.own_block;
  uint32_t ep1, ep2;
  refit_endpoints(t.pf, word, ep1, ep2);
  return (uint64_t)pack565_rgbx(ep1) | ((uint64_t)pack565_rgbx(ep2) << 16) | ((uint64_t)word << 32);
}

// ---- order-independent form of the reference's stateful winner scan ----
// Scan position p = row * W + col in the reference's loop order.  Three associative
// reductions replace the sequential loop with its inner-loop-only `break`:
//   first  = min p over candidates with err_diff <= 0
//   lastneg= max (row, -col) over candidates with err_diff < 0
//   best   = min (err_diff, p) over all accepted candidates
struct WinnerState {
  uint32_t first;    // 0xffffffff = none
  int lastneg;       // -1 = none; (row << 7) | (127 - col)
  uint32_t best;     // 0xffffffff = none; (err_diff + 65536) << 14 | p
};

__device__ __forceinline__ void winner_init(WinnerState &s) {
  s.first = 0xffffffffu; s.lastneg = -1; s.best = 0xffffffffu;
}

__device__ __forceinline__ void winner_update(WinnerState &s, int e, int row, int col, int W) {
  if (e == kRejected) return;
  uint32_t p = (uint32_t)(row * W + col);
  if (e <= 0) s.first = min(s.first, p);
  if (e < 0) s.lastneg = max(s.lastneg, (row << 7) | (127 - col));
  s.best = min(s.best, ((uint32_t)(e + 65536) << 14) | p);
}

// Branch-free form for the tiled kernels.  e must be a real err_diff (|e| <= 65025) or 65535
// for "rejected".  Positions are encoded as p = (row << 7) | col (col < 128), which orders
// exactly like the reference's row-major scan and needs no division to decode; the
// "last row, first column" key of the strictly-negative rule is then p ^ 127.
__device__ __forceinline__ void winner_update_fast(WinnerState &s, int e, uint32_t p) {
  s.best = min(s.best, (uint32_t)(e * 16384) + (p + (65536u << 14)));
  s.first = min(s.first, p | ((uint32_t)(-e) & 0x80000000u));   // sign(-e) set <=> e > 0
  s.lastneg = max(s.lastneg, (int)(p ^ 127u) | ~(e >> 31));       // -1 unless e < 0
}

// Scans rows [row_begin, row_end) of one target's W-wide window.  The word id of window
// position (row, col) is pos[row * stride + cdir * col]; errcol[id * kStride] is that word's err_diff
// for this target.  Lanes stride over the flattened scan order, so there is no per-position
// division.  kRemap handles the multi-chunk case: ids outside [c0, c0 + cn) read the
// all-rejected row `dummy`.
template <bool kRemap, typename E = int, int kStride = 33>
__device__ __forceinline__ void scan_window(WinnerState &ws, const uint16_t *pos, int stride, int cdir,
                                            const E *errcol, int W, int row_begin, int row_end, int lane,
                                            int c0, int cn, int dummy) {
  if (W == 32) {
    // default search area (16): one window row per step, lane = column
    const uint16_t *q = pos + row_begin * stride + cdir * lane;
    uint32_t p = (uint32_t)((row_begin << 7) | lane);
#pragma unroll 4
    for (int row = row_begin; row < row_end; ++row, q += stride, p += 128u) {
      int u = *q;
      if (kRemap) {
        u -= c0;
        u = ((unsigned)u < (unsigned)cn) ? u : dummy;
      }
      const int e = (int)errcol[u * kStride];
      // (a per-row vote that skips the first / lastneg reductions for rows without an err_diff <= 0
      // candidate was measured: 17.6 vs 14.5 ms -- such rows are rare on the benchmark content)
      winner_update_fast(ws, e, p);
    }
    return;
  }
  const int rstep = 32 / W, cstep = 32 - rstep * W;
  int row = row_begin + lane / W, col = lane - (lane / W) * W;
#pragma unroll 2
  for (; row < row_end;) {
    int u = pos[row * stride + cdir * col];
    if (kRemap) {
      u -= c0;
      u = ((unsigned)u < (unsigned)cn) ? u : dummy;
    }
    winner_update_fast(ws, (int)errcol[u * kStride], (uint32_t)((row << 7) | col));
    col += cstep;
    row += rstep;
    if (col >= W) { col -= W; ++row; }
  }
}

// Decodes a WinnerState built with winner_update_fast.
__device__ __forceinline__ int winner_resolve_fast(const WinnerState &s, int &row, int &col) {
  if (s.first < 0x80000000u) {
    row = (int)(s.first >> 7);
    col = (int)(s.first & 127u);
    if (s.lastneg >= 0 && (s.lastneg >> 7) > row) {
      row = s.lastneg >> 7;
      col = 127 - (s.lastneg & 127);
    }
    return 0;
  }
  if ((s.best >> 14) < 131071u) {
    row = (int)((s.best >> 7) & 127u);
    col = (int)(s.best & 127u);
    return (int)(s.best >> 14) - 65536;
  }
  row = col = 0;
  return 0x7fffffff;
}

__device__ __forceinline__ void winner_merge(WinnerState &s, const WinnerState &o) {
  s.first = min(s.first, o.first);
  s.lastneg = max(s.lastneg, o.lastneg);
  s.best = min(s.best, o.best);
}

__device__ __forceinline__ void winner_warp_reduce(WinnerState &s) {
  // REDUX.MIN / REDUX.MAX: one instruction per component instead of five shuffle + min rounds
  s.first = __reduce_min_sync(0xffffffffu, s.first);
  s.lastneg = __reduce_max_sync(0xffffffffu, s.lastneg);
  s.best = __reduce_min_sync(0xffffffffu, s.best);
}

// Returns the search's return value (min_err) and the winning (row, col); INT_MAX if nothing
// was accepted.
__device__ __forceinline__ int winner_resolve(const WinnerState &s, int W, int &row, int &col) {
  if (s.first < 0x80000000u) {
    row = (int)(s.first / (uint32_t)W);
    col = (int)(s.first % (uint32_t)W);
    if (s.lastneg >= 0 && (s.lastneg >> 7) > row) {
      row = s.lastneg >> 7;
      col = 127 - (s.lastneg & 127);
    }
    return 0;
  }
  if ((s.best >> 14) < 131071u) {  // 131071 = rejected marker (65535 + 65536), also the init value's range
    uint32_t p = s.best & 0x3FFFu;
    row = (int)(p / (uint32_t)W);
    col = (int)(p % (uint32_t)W);
    return (int)(s.best >> 14) - 65536;
  }
  row = col = 0;
  return 0x7fffffff;
}

// ------------------------------------------------------------------------------------------
// Shared by K2/K3: build the per-target context in shared memory.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void build_target(TargetCtx &t, const uint8_t *frame, int w, int bx, int by,
                                             uint64_t own_block) {
  // called by one thread
  uint32_t px[16];
  load_block_rgbx(frame, w, bx, by, px);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    t.px[k] = px[k];
    t.pf[3 * k + 0] = __uint2float_rn(px[k] & 0xFF);
    t.pf[3 * k + 1] = __uint2float_rn((px[k] >> 8) & 0xFF);
    t.pf[3 * k + 2] = __uint2float_rn((px[k] >> 16) & 0xFF);
  }
  t.own_block = own_block;
  t.own_word = (uint32_t)(own_block >> 32);
  uint32_t pal[4];
  palette_of_block(own_block, pal);
  t.orig_err = block_error(px, pal, t.own_word);
}

template <int NWARPS>
__device__ __forceinline__ void winner_block_reduce(WinnerState &s, WinnerState *smem) {
  winner_warp_reduce(s);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) smem[wid] = s;
  __syncthreads();
  if (wid == 0) {
    WinnerState o;
    winner_init(o);
    if (lane < NWARPS) o = smem[lane];
    winner_warp_reduce(o);
    if (lane == 0) smem[0] = o;
  }
  __syncthreads();
  s = smem[0];
}


// device-scope acquire / release on the wavefront progress counters
__device__ __forceinline__ int ld_acquire(const int *p) {
  cuda::atomic_ref<int, cuda::thread_scope_device> r(*const_cast<int *>(p));
  return r.load(cuda::memory_order_acquire);
}
__device__ __forceinline__ void st_release(int *p, int val) {
  cuda::atomic_ref<int, cuda::thread_scope_device> r(*p);
  r.store(val, cuda::memory_order_release);
}


}  // namespace mptc
