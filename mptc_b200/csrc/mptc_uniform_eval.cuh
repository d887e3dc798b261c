// mptc_uniform_eval.cuh -- candidate evaluation with a WARP-UNIFORM index word.
//
// The tiled search kernels evaluate one de-duplicated index word against 32 target blocks
// at a time: lane = target (pixels live in that lane's registers), word = the same for the
// whole warp.  Everything that depends on the word only (the per-pixel weights a, b, the
// asq/bsq/ab/f terms, the index selectors for the error sum) is computed once per distinct word
// into shared memory (WordInfo) and read back with broadcast loads; the ordered FP32
// accumulation of RecalculateEndpoints (dxt_image.cpp:298-318) is then 2 FMUL + 1 packed FADD2
// per pixel and channel, each operation individually rounded exactly like the reference.
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

// Per distinct word: everything that depends on the index word only.
struct WordInfo {
  float4 cf;        // asq, bsq, ab, f = 1/(asq*bsq - ab*ab)
  uint32_t sel[4];  // per block row: PRMT selector, nibble i = 2-bit index of pixel 4*row + i
  float2 w[16];     // per pixel: (a, b) = ((3-order)/3, order/3), idx_to_order = {0,3,1,2}
};

__device__ __forceinline__ void word_info(uint32_t word, WordInfo &wi) {
  const float w23 = 2.0f / 3.0f, w13 = 1.0f / 3.0f;
  wi.cf = word_coefs(word);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    uint32_t x = (word >> (8 * r)) & 0xFFu;   // 4 x 2 bits -> 4 nibbles
    x = (x | (x << 4)) & 0x0F0Fu;
    x = (x | (x << 2)) & 0x3333u;
    wi.sel[r] = x;
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint32_t v = (word >> (2 * k)) & 3u;
    const float a = (v & 2u) ? ((v & 1u) ? w13 : w23) : ((v & 1u) ? 0.0f : 1.0f);
    const float b = (v & 2u) ? ((v & 1u) ? w23 : w13) : ((v & 1u) ? 1.0f : 0.0f);
    wi.w[k] = make_float2(a, b);
  }
}

// All 16 indices equal <=> the least-squares system is singular (exact determinant 0; for any
// other word it is >= 1/9, so f <= 9 and every intermediate stays far inside int32).  Only
// these four words can reach the x86 float->int overflow semantics of dxt_image.cpp:330-331.
__device__ __forceinline__ bool word_is_degenerate(uint32_t word) { return (word ^ (word << 2)) < 4u; }

// One lane's target block held in registers.
struct LaneTarget {
  float pf[48];       // pixels as float, [k*3 + ch]
  uint32_t pl[12];    // planar bytes: pl[ch*4 + row] = channel ch of pixels 4*row .. 4*row+3
  uint64_t own_block;
  uint32_t own_word;
  int orig_err;
};

__device__ __forceinline__ void load_lane_target(LaneTarget &t, const uint8_t *frame, int w, int bx, int by,
                                                 uint64_t own_block) {
  uint32_t px[16];
  load_block_rgbx(frame, w, bx, by, px);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    t.pf[3 * k + 0] = __uint2float_rn(px[k] & 0xFF);
    t.pf[3 * k + 1] = __uint2float_rn((px[k] >> 8) & 0xFF);
    t.pf[3 * k + 2] = __uint2float_rn((px[k] >> 16) & 0xFF);
  }
#pragma unroll
  for (int ch = 0; ch < 3; ++ch)
#pragma unroll
    for (int r = 0; r < 4; ++r)
      t.pl[ch * 4 + r] = ((px[4 * r + 0] >> (8 * ch)) & 0xFFu) | (((px[4 * r + 1] >> (8 * ch)) & 0xFFu) << 8) |
                         (((px[4 * r + 2] >> (8 * ch)) & 0xFFu) << 16) | (((px[4 * r + 3] >> (8 * ch)) & 0xFFu) << 24);
  t.own_block = own_block;
  t.own_word = (uint32_t)(own_block >> 32);
  uint32_t pal[4];
  palette_of_block(own_block, pal);
  t.orig_err = block_error(px, pal, t.own_word);
}

// floor(x / 3) for 0 <= x < 2^31 in one IMAD.HI
__device__ __forceinline__ uint32_t div3(uint32_t x) { return __umulhi(x, 0x55555556u); }

// Sum over the 16 pixels of (pixel - palette[index])^2 for one channel: plane = 4 words of
// pixel bytes, palch = that channel of the 4 palette entries packed (entry v in byte v).
__device__ __forceinline__ uint32_t plane_error(const uint32_t *plane, uint32_t palch, const uint32_t *sel, uint32_t sum) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    uint32_t c;   // __byte_perm masks the selector first; the nibbles here are 0..3 by construction
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(c) : "r"(palch), "r"(0u), "r"(sel[r]));
    const uint32_t d = __vabsdiffu4(plane[r], c);
    sum = __dp4a(d, d, sum);
  }
  return sum;
}

constexpr int kRejectedSmall = 65535;  // "rejected" marker that fits the packed winner keys

// err_diff of (this lane's target, warp-uniform `word`), or kRejectedSmall.
// lut5/lut6: 256-entry tables of ToFiveBits / ToSixBits (dxt_image.cpp:72-121) in shared memory.
// packed_out (optional): the refitted endpoints as they would be emitted, Pack565(ep1) | Pack565(ep2) << 16,
// so that the winner's 8-byte block needs no second refit.
__device__ __forceinline__ int eval_uniform(const LaneTarget &t, uint32_t word, const WordInfo &wi,
                                            const uint8_t *lut5, const uint8_t *lut6, uint32_t *packed_out = nullptr) {
  // wi lives in shared memory: every read below is a warp-uniform (broadcast) load
  // (ax_j, bx_j) accumulated as pairs: two individually rounded products, one packed add.
  // (Scalar FMUL + FADD2 is not contracted by ptxas; FMUL2 + FADD2 would be.)
  float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float2 w = wi.w[k];   // warp-uniform broadcast load
    const float p0 = t.pf[3 * k + 0], p1 = t.pf[3 * k + 1], p2 = t.pf[3 * k + 2];
    s0 = __fadd2_rn(s0, make_float2(__fmul_rn(p0, w.x), __fmul_rn(p0, w.y)));
    s1 = __fadd2_rn(s1, make_float2(__fmul_rn(p1, w.x), __fmul_rn(p1, w.y)));
    s2 = __fadd2_rn(s2, make_float2(__fmul_rn(p2, w.x), __fmul_rn(p2, w.y)));
  }
  const float ax0 = s0.x, bx0 = s0.y, ax1 = s1.x, bx1 = s1.y, ax2 = s2.x, bx2 = s2.y;
  const float asq = wi.cf.x, bsq = wi.cf.y, ab = wi.cf.z, f = wi.cf.w;
  const float q0 = __fmul_rn(f, __fsub_rn(__fmul_rn(ax0, bsq), __fmul_rn(bx0, ab)));
  const float q1 = __fmul_rn(f, __fsub_rn(__fmul_rn(bx0, asq), __fmul_rn(ax0, ab)));
  const float q2 = __fmul_rn(f, __fsub_rn(__fmul_rn(ax1, bsq), __fmul_rn(bx1, ab)));
  const float q3 = __fmul_rn(f, __fsub_rn(__fmul_rn(bx1, asq), __fmul_rn(ax1, ab)));
  const float q4 = __fmul_rn(f, __fsub_rn(__fmul_rn(ax2, bsq), __fmul_rn(bx2, ab)));
  const float q5 = __fmul_rn(f, __fsub_rn(__fmul_rn(bx2, asq), __fmul_rn(ax2, ab)));
  uint32_t r1, r2, g1, g2, b1, b2;
  if (word_is_degenerate(word)) {  // warp-uniform, rare: exact cvttss2si emulation
    r1 = (uint32_t)quantise_endpoint(q0); r2 = (uint32_t)quantise_endpoint(q1);
    g1 = (uint32_t)quantise_endpoint(q2); g2 = (uint32_t)quantise_endpoint(q3);
    b1 = (uint32_t)quantise_endpoint(q4); b2 = (uint32_t)quantise_endpoint(q5);
  } else {  // finite and far inside int32: saturating F2I (negative -> 0) + min == cast + clamp
    r1 = min(__float2uint_rz(__fadd_rn(q0, 0.5f)), 255u); r2 = min(__float2uint_rz(__fadd_rn(q1, 0.5f)), 255u);
    g1 = min(__float2uint_rz(__fadd_rn(q2, 0.5f)), 255u); g2 = min(__float2uint_rz(__fadd_rn(q3, 0.5f)), 255u);
    b1 = min(__float2uint_rz(__fadd_rn(q4, 0.5f)), 255u); b2 = min(__float2uint_rz(__fadd_rn(q5, 0.5f)), 255u);
  }
  r1 = lut5[r1]; r2 = lut5[r2]; g1 = lut6[g1]; g2 = lut6[g2]; b1 = lut5[b1]; b2 = lut5[b2];
  // Pack565(ep1) > Pack565(ep2) (dxt_image.cpp:344) compares (r5, g6, b5) lexicographically: the same order as
  // the three bytes (r, g, b) with the bits Pack565 drops masked off.  (Not the unmasked bytes: ToFiveBits /
  // ToSixBits as written return values such as 12 whose low bits are not a replica of the high ones.)  The 565
  // words themselves are only assembled when the caller wants them.  Bytes 1-3 of the six values are zero:
  // selector nibbles 1 and 5 fetch a zero byte.
  const uint32_t key1 = __byte_perm(__byte_perm(b1, g1, 0x1140), r1, 0x5410) & 0x00F8FCF8u;   // b | g << 8 | r << 16
  const uint32_t key2 = __byte_perm(__byte_perm(b2, g2, 0x1140), r2, 0x5410) & 0x00F8FCF8u;
  // palette per channel, entry v in byte v: ep1, ep2, (2*ep1+ep2)/3, (ep1+2*ep2)/3
  const uint32_t palr = r1 | (r2 << 8) | (div3(2u * r1 + r2) << 16) | (div3(r1 + 2u * r2) << 24);
  const uint32_t palg = g1 | (g2 << 8) | (div3(2u * g1 + g2) << 16) | (div3(g1 + 2u * g2) << 24);
  const uint32_t palb = b1 | (b2 << 8) | (div3(2u * b1 + b2) << 16) | (div3(b1 + 2u * b2) << 24);
  uint32_t sum = plane_error(t.pl + 0, palr, wi.sel, 0u);
  sum = plane_error(t.pl + 4, palg, wi.sel, sum);
  sum = plane_error(t.pl + 8, palb, wi.sel, sum);
  if (packed_out) {
    const uint32_t pk1 = ((r1 & 0xF8u) << 8) | ((g1 & 0xFCu) << 3) | (b1 >> 3);
    const uint32_t pk2 = ((r2 & 0xF8u) << 8) | ((g2 & 0xFCu) << 3) | (b2 >> 3);
    *packed_out = pk1 | (pk2 << 16);
  }
  int e = (int)(sum / 48u) - t.orig_err;
  e = (key1 > key2) ? e : kRejectedSmall;
  return (word == t.own_word) ? 0 : e;
}

// The block this lane's target emits when `word` (not necessarily uniform) wins.
__device__ __forceinline__ uint64_t lane_winning_block(const LaneTarget &t, uint32_t word) {
  if (word == t.own_word) return t.own_block;
  uint32_t ep1, ep2;
  refit_endpoints(t.pf, word, ep1, ep2);
  return (uint64_t)pack565_rgbx(ep1) | ((uint64_t)pack565_rgbx(ep2) << 16) | ((uint64_t)word << 32);
}

}  // namespace mptc
