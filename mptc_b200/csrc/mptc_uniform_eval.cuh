// mptc_uniform_eval.cuh -- candidate evaluation with a WARP-UNIFORM index word.
//
// The tiled search kernels evaluate one de-duplicated index word against 32 target blocks
// at a time: lane = target (pixels live in that lane's registers), word = the same for the
// whole warp.  Every branch on the word's 2-bit indices is therefore non-divergent, so the
// ordered FP32 accumulation of RecalculateEndpoints (dxt_image.cpp:298-318) only issues the
// operations the index actually needs:
//   index 0 (a=1, b=0): ax += P          (P*1 is exact, bx + P*0 = bx exactly)
//   index 1 (a=0, b=1): bx += P
//   index 2 (a=2/3, b=1/3) / index 3 (a=1/3, b=2/3): separately rounded products, then adds
// which is bit-identical to the reference's "always multiply, always add" loop.
// NOTE: __fmul2_rn + __fadd2_rn must NOT be used for the products: ptxas (12.9) contracts
// mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (profiles/micro/f32x2_test.cu), which changes results.
#pragma once
#include "mptc_device.cuh"

namespace mptc {

// Terms of RecalculateEndpoints that depend on the index word only (dxt_image.cpp:310-312,
// :320): asq, bsq, ab accumulated in pixel order, and f = 1 / (asq*bsq - ab*ab).
__device__ __forceinline__ float4 word_coefs(uint32_t word) {
  const float w23 = 2.0f / 3.0f, w13 = 1.0f / 3.0f;
  float asq = 0.f, bsq = 0.f, ab = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    uint32_t v = (word >> (2 * k)) & 3u;
    float a = (v & 2u) ? ((v & 1u) ? w13 : w23) : ((v & 1u) ? 0.0f : 1.0f);
    float b = (v & 2u) ? ((v & 1u) ? w23 : w13) : ((v & 1u) ? 1.0f : 0.0f);
    asq = __fadd_rn(asq, __fmul_rn(a, a));
    bsq = __fadd_rn(bsq, __fmul_rn(b, b));
    ab = __fadd_rn(ab, __fmul_rn(a, b));
  }
  float f = __fdiv_rn(1.0f, __fsub_rn(__fmul_rn(asq, bsq), __fmul_rn(ab, ab)));
  return make_float4(asq, bsq, ab, f);
}

// One lane's target block held in registers.
struct LaneTarget {
  float pf[48];       // [k*3 + ch]
  uint32_t px[16];    // RGBX packed
  uint64_t own_block;
  uint32_t own_word;
  int orig_err;
};

__device__ __forceinline__ void load_lane_target(LaneTarget &t, const uint8_t *frame, int w, int bx, int by,
                                                 uint64_t own_block) {
  load_block_rgbx(frame, w, bx, by, t.px);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    t.pf[3 * k + 0] = __uint2float_rn(t.px[k] & 0xFF);
    t.pf[3 * k + 1] = __uint2float_rn((t.px[k] >> 8) & 0xFF);
    t.pf[3 * k + 2] = __uint2float_rn((t.px[k] >> 16) & 0xFF);
  }
  t.own_block = own_block;
  t.own_word = (uint32_t)(own_block >> 32);
  uint32_t pal[4];
  palette_of_block(own_block, pal);
  t.orig_err = block_error(t.px, pal, t.own_word);
}

// floor(x / 3) for 0 <= x < 2^31 in one IMAD.HI
__device__ __forceinline__ uint32_t div3(uint32_t x) { return __umulhi(x, 0x55555556u); }

// err_diff of (this lane's target, uniform `word`), or kRejected.  cf = word_coefs(word).
__device__ __forceinline__ int eval_uniform(const LaneTarget &t, uint32_t word, float4 cf) {
  const float w23 = 2.0f / 3.0f, w13 = 1.0f / 3.0f;
  float ax0 = 0.f, ax1 = 0.f, ax2 = 0.f, bx0 = 0.f, bx1 = 0.f, bx2 = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint32_t v = (word >> (2 * k)) & 3u;  // warp-uniform
    const float p0 = t.pf[3 * k + 0], p1 = t.pf[3 * k + 1], p2 = t.pf[3 * k + 2];
    if (v == 0u) {
      ax0 = __fadd_rn(ax0, p0); ax1 = __fadd_rn(ax1, p1); ax2 = __fadd_rn(ax2, p2);
    } else if (v == 1u) {
      bx0 = __fadd_rn(bx0, p0); bx1 = __fadd_rn(bx1, p1); bx2 = __fadd_rn(bx2, p2);
    } else if (v == 2u) {
      ax0 = __fadd_rn(ax0, __fmul_rn(p0, w23)); bx0 = __fadd_rn(bx0, __fmul_rn(p0, w13));
      ax1 = __fadd_rn(ax1, __fmul_rn(p1, w23)); bx1 = __fadd_rn(bx1, __fmul_rn(p1, w13));
      ax2 = __fadd_rn(ax2, __fmul_rn(p2, w23)); bx2 = __fadd_rn(bx2, __fmul_rn(p2, w13));
    } else {
      ax0 = __fadd_rn(ax0, __fmul_rn(p0, w13)); bx0 = __fadd_rn(bx0, __fmul_rn(p0, w23));
      ax1 = __fadd_rn(ax1, __fmul_rn(p1, w13)); bx1 = __fadd_rn(bx1, __fmul_rn(p1, w23));
      ax2 = __fadd_rn(ax2, __fmul_rn(p2, w13)); bx2 = __fadd_rn(bx2, __fmul_rn(p2, w23));
    }
  }
  const float asq = cf.x, bsq = cf.y, ab = cf.z, f = cf.w;
  int r1 = quantise_endpoint(__fmul_rn(f, __fsub_rn(__fmul_rn(ax0, bsq), __fmul_rn(bx0, ab))));
  int r2 = quantise_endpoint(__fmul_rn(f, __fsub_rn(__fmul_rn(bx0, asq), __fmul_rn(ax0, ab))));
  int g1 = quantise_endpoint(__fmul_rn(f, __fsub_rn(__fmul_rn(ax1, bsq), __fmul_rn(bx1, ab))));
  int g2 = quantise_endpoint(__fmul_rn(f, __fsub_rn(__fmul_rn(bx1, asq), __fmul_rn(ax1, ab))));
  int b1 = quantise_endpoint(__fmul_rn(f, __fsub_rn(__fmul_rn(ax2, bsq), __fmul_rn(bx2, ab))));
  int b2 = quantise_endpoint(__fmul_rn(f, __fsub_rn(__fmul_rn(bx2, asq), __fmul_rn(ax2, ab))));
  r1 = snap_bits<0xF8, 4, 5>(r1);  r2 = snap_bits<0xF8, 4, 5>(r2);
  g1 = snap_bits<0xFC, 2, 6>(g1);  g2 = snap_bits<0xFC, 2, 6>(g2);
  b1 = snap_bits<0xF8, 4, 5>(b1);  b2 = snap_bits<0xF8, 4, 5>(b2);
  const uint32_t pk1 = ((uint32_t)(r1 & 0xF8) << 8) | ((uint32_t)(g1 & 0xFC) << 3) | ((uint32_t)b1 >> 3);
  const uint32_t pk2 = ((uint32_t)(r2 & 0xF8) << 8) | ((uint32_t)(g2 & 0xFC) << 3) | ((uint32_t)b2 >> 3);
  uint32_t pal[4];
  pal[0] = (uint32_t)r1 | ((uint32_t)g1 << 8) | ((uint32_t)b1 << 16);
  pal[1] = (uint32_t)r2 | ((uint32_t)g2 << 8) | ((uint32_t)b2 << 16);
  pal[2] = div3(2u * r1 + r2) | (div3(2u * g1 + g2) << 8) | (div3(2u * b1 + b2) << 16);
  pal[3] = div3(r1 + 2u * r2) | (div3(g1 + 2u * g2) << 8) | (div3(b1 + 2u * b2) << 16);
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint32_t v = (word >> (2 * k)) & 3u;  // warp-uniform
    uint32_t c;
    if (v == 0u) c = pal[0];
    else if (v == 1u) c = pal[1];
    else if (v == 2u) c = pal[2];
    else c = pal[3];
    const uint32_t d = __vabsdiffu4(t.px[k], c);
    sum = __dp4a(d, d, sum);
  }
  int e = (int)(sum / 48u) - t.orig_err;
  e = (pk1 > pk2) ? e : kRejected;
  return (word == t.own_word) ? 0 : e;
}

}  // namespace mptc
