/* oracle/mptc_oracle.c -- TEST INFRASTRUCTURE, see mptc_oracle.h.
 *
 * Plain-C restatement of the reference encoder hot path.  Written from the behaviour
 * of the reference (citations are path:line under /root/reference), not from its text.
 * Compile with -ffp-contract=off: the reference's results depend on every FP32 op
 * being rounded individually.
 */
#include "mptc_oracle.h"

#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------
 * 565 helpers (codec/dxt_image.cpp:43-69)
 * ---------------------------------------------------------------------------------- */
static void decode565(unsigned v, uint8_t out[3]) {
  unsigned r = v >> 11, g = (v >> 5) & 63, b = v & 31;
  out[0] = (uint8_t)((r << 3) | (r >> 2));
  out[1] = (uint8_t)((g << 2) | (g >> 4));
  out[2] = (uint8_t)((b << 3) | (b >> 2));
}

static unsigned pack565(const uint8_t c[3]) {
  return ((unsigned)(c[0] & 0xF8) << 8) | ((unsigned)(c[1] & 0xFC) << 3) | ((unsigned)c[2] >> 3);
}

/* ------------------------------------------------------------------------------------
 * stb_dxt tables (Include/stb_dxt.h:593-611, :111-137)
 * ---------------------------------------------------------------------------------- */
static uint8_t g_expand5[32], g_expand6[64];
static uint8_t g_omatch5[256][2], g_omatch6[256][2];
static pthread_once_t g_tables_once = PTHREAD_ONCE_INIT;

static int mul8bit(int a, int b) { /* stb_dxt.h:64-68 */
  int t = a * b + 128;
  return (t + (t >> 8)) >> 8;
}

static void build_single_colour_table(uint8_t tab[256][2], const uint8_t *expand, int size) {
  for (int target = 0; target < 256; ++target) {
    int best = 256;
    for (int lo = 0; lo < size; ++lo)
      for (int hi = 0; hi < size; ++hi) {
        int e_lo = expand[lo], e_hi = expand[hi];
        int err = abs((2 * e_hi + e_lo) / 3 - target) + abs(e_hi - e_lo) * 3 / 100;
        if (err < best) { tab[target][0] = (uint8_t)hi; tab[target][1] = (uint8_t)lo; best = err; }
      }
  }
}

static void build_tables(void) {
  for (int i = 0; i < 32; ++i) g_expand5[i] = (uint8_t)((i << 3) | (i >> 2));
  for (int i = 0; i < 64; ++i) g_expand6[i] = (uint8_t)((i << 2) | (i >> 4));
  build_single_colour_table(g_omatch5, g_expand5, 32);
  build_single_colour_table(g_omatch6, g_expand6, 64);
}

/* Exposed so that the product's table generator can be pinned against it in tests. */
void mptc_oracle_tables(uint8_t *omatch5 /*512*/, uint8_t *omatch6 /*512*/) {
  pthread_once(&g_tables_once, build_tables);
  memcpy(omatch5, g_omatch5, 512);
  memcpy(omatch6, g_omatch6, 512);
}

/* ------------------------------------------------------------------------------------
 * stb DXT1 block fit, HIGHQUAL, no dither, no alpha (stb_dxt.h:467-538)
 * ---------------------------------------------------------------------------------- */
static unsigned quant565(int r, int g, int b) { /* stb__As16Bit :82-85 */
  return (unsigned)((mul8bit(r, 31) << 11) + (mul8bit(g, 63) << 5) + mul8bit(b, 31));
}

static void palette_of(unsigned c0, unsigned c1, int col[4][3]) { /* stb__EvalColors :139-145 */
  col[0][0] = g_expand5[c0 >> 11]; col[0][1] = g_expand6[(c0 >> 5) & 63]; col[0][2] = g_expand5[c0 & 31];
  col[1][0] = g_expand5[c1 >> 11]; col[1][1] = g_expand6[(c1 >> 5) & 63]; col[1][2] = g_expand5[c1 & 31];
  for (int k = 0; k < 3; ++k) {
    col[2][k] = (2 * col[0][k] + col[1][k]) / 3;
    col[3][k] = (2 * col[1][k] + col[0][k]) / 3;
  }
}

static uint32_t match_indices(const uint8_t px[16][3], int col[4][3]) { /* :176-215 */
  int dr = col[0][0] - col[1][0], dg = col[0][1] - col[1][1], db = col[0][2] - col[1][2];
  int stops[4];
  for (int i = 0; i < 4; ++i) stops[i] = col[i][0] * dr + col[i][1] * dg + col[i][2] * db;
  int c0pt = (stops[1] + stops[3]) >> 1;
  int half = (stops[3] + stops[2]) >> 1;
  int c3pt = (stops[2] + stops[0]) >> 1;
  uint32_t mask = 0;
  for (int i = 15; i >= 0; --i) {
    int dot = px[i][0] * dr + px[i][1] * dg + px[i][2] * db;
    mask <<= 2;
    if (dot < half) mask |= (dot < c0pt) ? 1u : 3u;
    else            mask |= (dot < c3pt) ? 2u : 0u;
  }
  return mask;
}

static void pca_endpoints(const uint8_t px[16][3], unsigned *pmax, unsigned *pmin) { /* :273-375 */
  int mu[3], lo[3], hi[3];
  for (int ch = 0; ch < 3; ++ch) {
    int s = 0, mn = 255, mx = 0;
    for (int i = 0; i < 16; ++i) {
      int v = px[i][ch];
      s += v;
      if (v < mn) mn = v;
      if (v > mx) mx = v;
    }
    mu[ch] = (s + 8) >> 4; lo[ch] = mn; hi[ch] = mx;
  }
  int cov[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 16; ++i) {
    int r = px[i][0] - mu[0], g = px[i][1] - mu[1], b = px[i][2] - mu[2];
    cov[0] += r * r; cov[1] += r * g; cov[2] += r * b;
    cov[3] += g * g; cov[4] += g * b; cov[5] += b * b;
  }
  float cf[6];
  for (int i = 0; i < 6; ++i) cf[i] = (float)cov[i] / 255.0f;
  float vr = (float)(hi[0] - lo[0]), vg = (float)(hi[1] - lo[1]), vb = (float)(hi[2] - lo[2]);
  for (int it = 0; it < 4; ++it) {
    float r = vr * cf[0] + vg * cf[1] + vb * cf[2];
    float g = vr * cf[1] + vg * cf[3] + vb * cf[4];
    float b = vr * cf[2] + vg * cf[4] + vb * cf[5];
    vr = r; vg = g; vb = b;
  }
  double magn = fabs((double)vr);
  if (fabs((double)vg) > magn) magn = fabs((double)vg);
  if (fabs((double)vb) > magn) magn = fabs((double)vb);
  int ar, ag, ab;
  if (magn < 4.0) { ar = 299; ag = 587; ab = 114; }
  else {
    magn = 512.0 / magn;
    ar = (int)((double)vr * magn); ag = (int)((double)vg * magn); ab = (int)((double)vb * magn);
  }
  int dmin = INT_MAX, dmax = -INT_MAX, imin = 0, imax = 0;
  for (int i = 0; i < 16; ++i) {
    int dot = px[i][0] * ar + px[i][1] * ag + px[i][2] * ab;
    if (dot < dmin) { dmin = dot; imin = i; }
    if (dot > dmax) { dmax = dot; imax = i; }
  }
  *pmax = quant565(px[imax][0], px[imax][1], px[imax][2]);
  *pmin = quant565(px[imin][0], px[imin][1], px[imin][2]);
}

static int clampi_trunc(float y, int lo, int hi) { /* stb__sclamp :377-383 */
  int x = (int)y;
  return x < lo ? lo : (x > hi ? hi : x);
}

static int refine_endpoints(const uint8_t px[16][3], unsigned *pmax, unsigned *pmin, uint32_t mask) { /* :388-464 */
  static const int w1tab[4] = {3, 0, 2, 1};
  static const int prods[4] = {0x090000, 0x000900, 0x040102, 0x010402};
  unsigned old_min = *pmin, old_max = *pmax, mn, mx;
  if ((mask ^ (mask << 2)) < 4) { /* every pixel has the same index */
    int r = 8, g = 8, b = 8;
    for (int i = 0; i < 16; ++i) { r += px[i][0]; g += px[i][1]; b += px[i][2]; }
    r >>= 4; g >>= 4; b >>= 4;
    mx = ((unsigned)g_omatch5[r][0] << 11) | ((unsigned)g_omatch6[g][0] << 5) | g_omatch5[b][0];
    mn = ((unsigned)g_omatch5[r][1] << 11) | ((unsigned)g_omatch6[g][1] << 5) | g_omatch5[b][1];
  } else {
    int a1[3] = {0, 0, 0}, a2[3] = {0, 0, 0}, akku = 0;
    uint32_t cm = mask;
    for (int i = 0; i < 16; ++i, cm >>= 2) {
      int step = (int)(cm & 3), w1 = w1tab[step];
      akku += prods[step];
      for (int k = 0; k < 3; ++k) { a1[k] += w1 * px[i][k]; a2[k] += px[i][k]; }
    }
    for (int k = 0; k < 3; ++k) a2[k] = 3 * a2[k] - a1[k];
    int xx = akku >> 16, yy = (akku >> 8) & 0xff, xy = akku & 0xff;
    float frb = 3.0f * 31.0f / 255.0f / (float)(xx * yy - xy * xy);
    float fg = frb * 63.0f / 31.0f;
    mx  = (unsigned)clampi_trunc((float)(a1[0] * yy - a2[0] * xy) * frb + 0.5f, 0, 31) << 11;
    mx |= (unsigned)clampi_trunc((float)(a1[1] * yy - a2[1] * xy) * fg  + 0.5f, 0, 63) << 5;
    mx |= (unsigned)clampi_trunc((float)(a1[2] * yy - a2[2] * xy) * frb + 0.5f, 0, 31);
    mn  = (unsigned)clampi_trunc((float)(a2[0] * xx - a1[0] * xy) * frb + 0.5f, 0, 31) << 11;
    mn |= (unsigned)clampi_trunc((float)(a2[1] * xx - a1[1] * xy) * fg  + 0.5f, 0, 63) << 5;
    mn |= (unsigned)clampi_trunc((float)(a2[2] * xx - a1[2] * xy) * frb + 0.5f, 0, 31);
  }
  *pmin = mn; *pmax = mx;
  return old_min != mn || old_max != mx;
}

static uint64_t fit_block(const uint8_t px[16][3]) { /* stb__CompressColorBlock :467-538 */
  unsigned mx, mn;
  uint32_t mask;
  int constant = 1;
  for (int i = 1; i < 16; ++i)
    if (px[i][0] != px[0][0] || px[i][1] != px[0][1] || px[i][2] != px[0][2]) { constant = 0; break; }
  if (constant) {
    int r = px[0][0], g = px[0][1], b = px[0][2];
    mask = 0xAAAAAAAAu;
    mx = ((unsigned)g_omatch5[r][0] << 11) | ((unsigned)g_omatch6[g][0] << 5) | g_omatch5[b][0];
    mn = ((unsigned)g_omatch5[r][1] << 11) | ((unsigned)g_omatch6[g][1] << 5) | g_omatch5[b][1];
  } else {
    int col[4][3];
    pca_endpoints(px, &mx, &mn);
    if (mx != mn) { palette_of(mx, mn, col); mask = match_indices(px, col); }
    else mask = 0;
    for (int pass = 0; pass < 2; ++pass) { /* HIGHQUAL: refinecount = 2 */
      uint32_t last = mask;
      if (refine_endpoints(px, &mx, &mn, mask)) {
        if (mx != mn) { palette_of(mx, mn, col); mask = match_indices(px, col); }
        else { mask = 0; break; }
      }
      if (mask == last) break;
    }
  }
  if (mx < mn) { unsigned t = mn; mn = mx; mx = t; mask ^= 0x55555555u; }
  return (uint64_t)mx | ((uint64_t)mn << 16) | ((uint64_t)mask << 32);
}

static void load_block(const uint8_t *rgb, int w, int bx, int by, uint8_t px[16][3]) {
  /* CompressRGB / Get4X4ColorsBlock (dxt_image.cpp:123-140, :636-650) */
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) {
      const uint8_t *s = rgb + ((size_t)(by * 4 + j) * w + bx * 4 + i) * 3;
      px[j * 4 + i][0] = s[0]; px[j * 4 + i][1] = s[1]; px[j * 4 + i][2] = s[2];
    }
}

void mptc_oracle_dxt1_fit(const uint8_t *rgb, int w, int h, uint64_t *blocks_out) {
  pthread_once(&g_tables_once, build_tables);
  int bw = w >> 2, bh = h >> 2;
  uint8_t px[16][3];
  for (int by = 0; by < bh; ++by)
    for (int bx = 0; bx < bw; ++bx) {
      load_block(rgb, w, bx, by, px);
      blocks_out[(size_t)by * bw + bx] = fit_block(px);
    }
}

/* ------------------------------------------------------------------------------------
 * Candidate evaluation (dxt_image.cpp:244-353, :739-758)
 * ---------------------------------------------------------------------------------- */
typedef struct { uint8_t c[4][3]; } pal4;

static void palette_from_physical(uint64_t blk, pal4 *p) { /* PhysicalToLogical :198-214 */
  unsigned e1 = (unsigned)(blk & 0xFFFF), e2 = (unsigned)((blk >> 16) & 0xFFFF);
  decode565(e1, p->c[0]);
  decode565(e2, p->c[1]);
  for (int k = 0; k < 3; ++k) {
    int a = p->c[0][k], b = p->c[1][k];
    if (e1 <= e2) { p->c[2][k] = (uint8_t)((a + b) / 2); p->c[3][k] = 0; }
    else          { p->c[2][k] = (uint8_t)((2 * a + b) / 3); p->c[3][k] = (uint8_t)((a + 2 * b) / 3); }
  }
}

static int block_error(const uint8_t *px48, const pal4 *p, uint32_t word) { /* Error() :258-276 */
  unsigned sum = 0;
  for (int k = 0; k < 16; ++k) {
    const uint8_t *c = p->c[(word >> (2 * k)) & 3];
    for (int ch = 0; ch < 3; ++ch) {
      int d = (int)px48[3 * k + ch] - (int)c[ch];
      sum += (unsigned)(d * d);
    }
  }
  return (int)(sum / 48u);
}

static int cvt_x86(float x) { /* cvttss2si: out-of-range / NaN -> INT_MIN */
  if (!(x >= -2147483648.0f && x < 2147483648.0f)) return INT_MIN;
  return (int)x;
}

static uint8_t snap_bits(uint8_t x, int keep_mask, int step, int shift) {
  /* ToFiveBits / ToSixBits exactly as written (dxt_image.cpp:72-121): the neighbours
   * are base +- step with step = 4 (5 bit) or 2 (6 bit), all arithmetic in uint8. */
  uint8_t base = (uint8_t)(x & keep_mask);
  uint8_t high = (uint8_t)(base == 255 ? base : base + step);
  uint8_t low  = (uint8_t)(base == 0 ? base : base - step);
  base = (uint8_t)(base | (base >> shift));
  high = (uint8_t)(high | (high >> shift));
  low  = (uint8_t)(low | (low >> shift));
  uint8_t db = (uint8_t)(x > base ? x - base : base - x);
  uint8_t dh = (uint8_t)(x > high ? x - high : high - x);
  uint8_t dl = (uint8_t)(x > low ? x - low : low - x);
  if (db <= dh) return db < dl ? base : low;
  return high;
}

static void refit_endpoints(const uint8_t *px48, uint32_t word, pal4 *out) { /* RecalculateEndpoints :290-351 */
  static const float order_of[4] = {0.f, 3.f, 1.f, 2.f};
  float asq = 0.f, bsq = 0.f, ab = 0.f, ax[3] = {0.f, 0.f, 0.f}, bx[3] = {0.f, 0.f, 0.f};
  for (int k = 0; k < 16; ++k) {
    float order = order_of[(word >> (2 * k)) & 3];
    float a = (3.0f - order) / 3.0f, b = order / 3.0f;
    asq += a * a; bsq += b * b; ab += a * b;
    for (int ch = 0; ch < 3; ++ch) {
      float p = (float)px48[3 * k + ch];
      ax[ch] += p * a;
      bx[ch] += p * b;
    }
  }
  float f = 1.0f / (asq * bsq - ab * ab);
  for (int ch = 0; ch < 3; ++ch) {
    float p1 = f * (ax[ch] * bsq - bx[ch] * ab);
    float p2 = f * (bx[ch] * asq - ax[ch] * ab);
    int e1 = cvt_x86(p1 + 0.5f), e2 = cvt_x86(p2 + 0.5f);
    e1 = e1 > 255 ? 255 : e1; e1 = e1 < 0 ? 0 : e1;
    e2 = e2 > 255 ? 255 : e2; e2 = e2 < 0 ? 0 : e2;
    if (ch == 1) { out->c[0][ch] = snap_bits((uint8_t)e1, 0xFC, 2, 6); out->c[1][ch] = snap_bits((uint8_t)e2, 0xFC, 2, 6); }
    else         { out->c[0][ch] = snap_bits((uint8_t)e1, 0xF8, 4, 5); out->c[1][ch] = snap_bits((uint8_t)e2, 0xF8, 4, 5); }
  }
  for (int ch = 0; ch < 3; ++ch) {
    int a = out->c[0][ch], b = out->c[1][ch];
    out->c[2][ch] = (uint8_t)((2 * a + b) / 3);
    out->c[3][ch] = (uint8_t)((a + 2 * b) / 3);
  }
}

int mptc_oracle_eval_candidate(const uint8_t *px48, uint64_t own, uint32_t word, int *err_diff,
                               uint64_t *new_block) {
  pal4 own_pal, pal;
  uint32_t own_word = (uint32_t)(own >> 32);
  palette_from_physical(own, &own_pal);
  int orig = block_error(px48, &own_pal, own_word);
  if (word == own_word) { /* blk == blk2 (:742): palette untouched, err_diff == 0 */
    *err_diff = 0;
    if (new_block) *new_block = own;
    return 1;
  }
  refit_endpoints(px48, word, &pal);
  unsigned p1 = pack565(pal.c[0]), p2 = pack565(pal.c[1]);
  if (!(p1 > p2)) return 0; /* L2P would swap / 3-colour mode (:750-755, :154-161) */
  *err_diff = block_error(px48, &pal, word) - orig;
  if (new_block) *new_block = (uint64_t)p1 | ((uint64_t)p2 << 16) | ((uint64_t)word << 32);
  return 1;
}

/* ------------------------------------------------------------------------------------
 * Searches and the frame driver (dxt_image.cpp:652-774, :868-957)
 * ---------------------------------------------------------------------------------- */
typedef struct { int x, y; uint64_t blk; } hit;

static int search_inter(const uint8_t *px48, uint64_t own, const uint64_t *prev, int bw, int bh,
                        int bx, int by, int sa, hit *out) {
  int best = INT_MAX;
  for (int j = by - sa; j < by + sa; ++j)
    for (int i = bx - sa; i < bx + sa; ++i) {
      if (i < 0 || j < 0 || i >= bw || j >= bh) continue;
      int d; uint64_t nb;
      if (!mptc_oracle_eval_candidate(px48, own, (uint32_t)(prev[(size_t)j * bw + i] >> 32), &d, &nb)) continue;
      if (d < best) {
        best = d; out->x = (i - bx) + sa; out->y = (j - by) + sa; out->blk = nb;
        if (d <= 0) { best = 0; break; }
      }
    }
  return best;
}

static int search_intra(const uint8_t *px48, uint64_t own, const uint64_t *cur, int bw, int bh,
                        int bx, int by, int sa, hit *out) {
  int best = INT_MAX;
  for (int j = by; j >= by - 2 * sa + 1; --j)
    for (int i = bx + sa - 1; i >= bx - sa; --i) {
      if (i < 0 || j < 0 || i >= bw || j >= bh || (j == by && i >= bx)) continue;
      int d; uint64_t nb;
      if (!mptc_oracle_eval_candidate(px48, own, (uint32_t)(cur[(size_t)j * bw + i] >> 32), &d, &nb)) continue;
      if (d < best) {
        best = d; out->x = (i - bx) + sa; out->y = (j - by) + 2 * sa - 1; out->blk = nb;
        if (d <= 0) { best = 0; break; }
      }
    }
  return best;
}

int mptc_oracle_reencode(const uint8_t *rgb, int w, int h, int is_intra, int sa, int thr,
                         uint64_t *blocks, const uint64_t *prev, uint8_t *motion, uint32_t *unique) {
  int bw = w >> 2, bh = h >> 2, nu = 0;
  uint8_t px[16][3];
  for (int by = 0; by < bh; ++by)
    for (int bx = 0; bx < bw; ++bx) {
      size_t b = (size_t)by * bw + bx;
      hit ht;
      load_block(rgb, w, bx, by, px);
      if (!is_intra && prev) {
        if (search_inter(&px[0][0], blocks[b], prev, bw, bh, bx, by, sa, &ht) <= thr) {
          motion[2 * b] = (uint8_t)(ht.x | 0x80); motion[2 * b + 1] = (uint8_t)(ht.y | 0x80);
          blocks[b] = ht.blk;
          continue;
        }
      }
      if (search_intra(&px[0][0], blocks[b], blocks, bw, bh, bx, by, sa, &ht) <= thr) {
        motion[2 * b] = (uint8_t)ht.x; motion[2 * b + 1] = (uint8_t)ht.y;
        blocks[b] = ht.blk;
        continue;
      }
      motion[2 * b] = 255; motion[2 * b + 1] = 255;
      unique[nu++] = (uint32_t)(blocks[b] >> 32);
    }
  return nu;
}

/* Decoder-side index reconstruction (ReconstructDXTData, codec/codec.cpp:441-500): rebuilds
 * every block's interp word from the motion bytes, the unique list and the previous
 * frame's words.  Returns the number of unique words consumed, or -1 on a bad vector. */
/* ---- DXTImage::InterPixelSearch (dxt_image.cpp:776-832; call site commented out at :930-951) ---------
 * Pixel-granular inter search: candidate index words are 4x4 cut-outs of the previous frame's index
 * picture at pixel offsets (i, j) in the order of DXTImage::SetPattern (dxt_image.h:135-164).
 *
 * UNDEFINED BEHAVIOUR IN THE REFERENCE: Get4X4InterpolationBlock (dxt_image.cpp:619-634) fills only
 * the indices of a LogicalDXTBlock and passes it to LogicalToPhysical, which decides whether to flip
 * the word (^= 0x55555555) from the block's UNINITIALISED endpoints / palette (:150-170).  Compiled
 * with the canonical flags of oracle/Makefile (-O3 -DNDEBUG, as with -O0) g++ 13 leaves the word
 * unflipped; -O2 flips every word (probed).  This restatement follows the canonical build: the 16
 * gathered 2-bit indices packed as they are. */
int mptc_oracle_ips_pattern(int sa, int8_t *ij /* 2 * count, may be NULL */) {
  int n = 0;
#define PUSH(x, y) do { if (ij) { ij[2 * n] = (int8_t)(x); ij[2 * n + 1] = (int8_t)(y); } ++n; } while (0)
  PUSH(0, 0);
  for (int level = 1; level < sa; ++level)
    for (int cur = level; cur >= -level; --cur) {
      if (cur == level) { for (int x = cur; x >= -cur; --x) PUSH(x, cur); }
      else if (cur == -level) { for (int x = cur; x <= -cur; ++x) PUSH(x, cur); }
      else { PUSH(level, cur); PUSH(-level, cur); }
    }
#undef PUSH
  return n;
}

static uint32_t gather_word(const uint64_t *prev, int bw, int x, int y) { /* Get4X4InterpolationBlock, canonical build */
  uint32_t word = 0;
  for (int v = 0; v < 4; ++v)
    for (int u = 0; u < 4; ++u) {
      int px = x + u, py = y + v;
      uint32_t interp = (uint32_t)(prev[(py >> 2) * bw + (px >> 2)] >> 32);
      uint32_t idx = (interp >> (2 * ((py & 3) * 4 + (px & 3)))) & 3u; /* InterpolationValueAt :604-608 */
      word |= idx << (2 * (4 * v + u));
    }
  return word;
}

void mptc_oracle_inter_pixel_search(const uint8_t *rgb, int w, int h, int sa, const uint64_t *cur_blocks,
                                    const uint64_t *prev_blocks, int32_t *min_err_out, uint8_t *motion_out,
                                    uint32_t *index_out, uint8_t *reassigned_out) {
  const int bw = w / 4, bh = h / 4;
  const int n = mptc_oracle_ips_pattern(sa, NULL);
  int8_t *pat = (int8_t *)malloc((size_t)2 * n);
  mptc_oracle_ips_pattern(sa, pat);
  for (int b = 0; b < bw * bh; ++b) {
    const int bx = b % bw, by = b / bw;
    uint8_t px[16][3];
    load_block(rgb, w, bx, by, px);
    int min_err = INT_MAX, mx = 0, my = 0, re = 0;
    uint32_t index = 0;
    for (int k = 0; k < n; ++k) {
      const int i = pat[2 * k], j = pat[2 * k + 1];
      const int x = 4 * bx + i, y = 4 * by + j;
      if (!(x >= 0 && x <= w - 4 && y >= 0 && y <= h - 4)) continue;              /* :795 */
      const uint32_t word = gather_word(prev_blocks, bw, x, y);
      int ed;
      if (!mptc_oracle_eval_candidate(&px[0][0], cur_blocks[b], word, &ed, NULL)) continue;   /* :800-813 */
      if (ed < min_err) {                                                            /* :816-825 */
        min_err = ed; mx = i + 64; my = j + 64; index = word;
        re = word != (uint32_t)(cur_blocks[b] >> 32);
        if (ed <= 0) { min_err = 0; break; }
      }
    }
    min_err_out[b] = min_err;
    motion_out[2 * b] = (uint8_t)mx; motion_out[2 * b + 1] = (uint8_t)my;
    index_out[b] = index;
    reassigned_out[b] = (uint8_t)re;
  }
  free(pat);
}

int mptc_oracle_reconstruct_words(const uint8_t *motion, const uint32_t *unique, int n_unique,
                                  const uint32_t *prev_words, int bw, int bh, int sa, uint32_t *out) {
  int nu = 0;
  for (int b = 0; b < bw * bh; ++b) {
    int x = motion[2 * b], y = motion[2 * b + 1], bx = b % bw, by = b / bw;
    if (x == 255 && y == 255) {
      if (nu >= n_unique) return -1;
      out[b] = unique[nu++];
    } else if ((x & 0x80) && (y & 0x80)) {
      int rx = bx + (x & 0x7F) - sa, ry = by + (y & 0x7F) - sa;
      if (!prev_words || rx < 0 || ry < 0 || rx >= bw || ry >= bh) return -1;
      out[b] = prev_words[ry * bw + rx];
    } else {
      int rx = bx + x - sa, ry = by + y - (2 * sa - 1);
      if (rx < 0 || ry < 0 || rx >= bw || ry >= bh || ry * bw + rx >= b) return -1;
      out[b] = out[ry * bw + rx];
    }
  }
  return nu;
}

/* Inductive spot check for frames too large to run the whole oracle on: for each listed
 * block, redo the reference's decision for THAT block given the final blocks of its window
 * (cur_final for already-visited positions, prev_final for the inter window) and compare it
 * with what `cur_final`/`motion` hold.  If every block of a frame passes, the frame equals
 * the reference's output by induction over raster order.  Returns the number of mismatches. */
int mptc_oracle_check_blocks(const uint8_t *rgb, int w, int h, int is_intra, int sa, int thr,
                             const uint64_t *init_blocks, const uint64_t *cur_final,
                             const uint64_t *prev_final, const uint8_t *motion,
                             const int *which, int n_which) {
  int bw = w >> 2, bh = h >> 2, bad = 0;
  uint8_t px[16][3];
  for (int q = 0; q < n_which; ++q) {
    int b = which[q], bx = b % bw, by = b / bw;
    hit ht;
    uint64_t want_blk;
    uint8_t want_mx, want_my;
    load_block(rgb, w, bx, by, px);
    if (!is_intra && prev_final &&
        search_inter(&px[0][0], init_blocks[b], prev_final, bw, bh, bx, by, sa, &ht) <= thr) {
      want_blk = ht.blk; want_mx = (uint8_t)(ht.x | 0x80); want_my = (uint8_t)(ht.y | 0x80);
    } else if (search_intra(&px[0][0], init_blocks[b], cur_final, bw, bh, bx, by, sa, &ht) <= thr) {
      want_blk = ht.blk; want_mx = (uint8_t)ht.x; want_my = (uint8_t)ht.y;
    } else {
      want_blk = init_blocks[b]; want_mx = 255; want_my = 255;
    }
    bad += (want_blk != cur_final[b]) || want_mx != motion[2 * b] || want_my != motion[2 * b + 1];
  }
  return bad;
}

/* ------------------------------------------------------------------------------------
 * Endpoint planes: RGB565 -> YCoCg667 -> 64x64-tiled 5/3 integer wavelet -> uint8 symbols
 * (dxt_image.cpp:496-530, image_processing.cpp:10-27, image_processing.h:292-333,
 *  wavelet.cpp:11-131, image_utils.h:228-239)
 * ---------------------------------------------------------------------------------- */
static void lift53(const int16_t *src, int16_t *dst, int len) { /* ForwardWavelet1D, even len >= 2 */
  int mid = len - len / 2;
  for (int i = 1; i < len; i += 2) {
    int nxt = (i + 1 < len) ? i + 1 : 2 * len - 2 - (i + 1);
    dst[mid + i / 2] = (int16_t)(src[i] - (src[i - 1] + src[nxt]) / 2);
  }
  for (int i = 0; i < len; i += 2) {
    int dp = (i == 0) ? 0 : (i - 1) / 2; /* mirror: index -1 -> 1 */
    int dn = (i + 1 < len) ? (i + 1) / 2 : (2 * len - 2 - (i + 1)) / 2;
    dst[i / 2] = (int16_t)(src[i] + (dst[mid + dp] + dst[mid + dn] + 2) / 4);
  }
}

static void wavelet_tile(int16_t t[64][64]) {
  int16_t line[64], outl[64];
  for (int dim = 64; dim > 1; dim >>= 1) {
    for (int c = 0; c < dim; ++c) { /* columns first (wavelet.cpp:110-116) */
      for (int r = 0; r < dim; ++r) line[r] = t[r][c];
      lift53(line, outl, dim);
      for (int r = 0; r < dim; ++r) t[r][c] = outl[r];
    }
    for (int r = 0; r < dim; ++r) { /* then rows (:126-130) */
      memcpy(line, t[r], sizeof(int16_t) * dim);
      lift53(line, outl, dim);
      memcpy(t[r], outl, sizeof(int16_t) * dim);
    }
  }
}

void mptc_oracle_endpoint_planes(const uint64_t *blocks, int bw, int bh, uint8_t *planes) {
  int pbw = (bw + 63) / 64 * 64, pbh = (bh + 63) / 64 * 64;
  size_t pn = (size_t)pbw * pbh;
  int16_t tile[64][64];
  for (int ep = 0; ep < 2; ++ep)
    for (int ch = 0; ch < 3; ++ch) {
      uint8_t *dst = planes + (size_t)(ep * 3 + ch) * pn;
      for (int ty = 0; ty < pbh; ty += 64)
        for (int tx = 0; tx < pbw; tx += 64) {
          for (int y = 0; y < 64; ++y)
            for (int x = 0; x < 64; ++x) {
              int sx = tx + x < bw ? tx + x : bw - 1; /* edge replication = extension */
              int sy = ty + y < bh ? ty + y : bh - 1;
              unsigned v = (unsigned)((blocks[(size_t)sy * bw + sx] >> (16 * ep)) & 0xFFFF);
              int r = (int)(v >> 11), g = (int)((v >> 5) & 63), b = (int)(v & 31);
              int co = r - b, t = r + b + (b >> 4), cg = g - t, yy = t + cg / 2;
              tile[y][x] = (int16_t)(ch == 0 ? yy : (ch == 1 ? co : cg));
            }
          wavelet_tile(tile);
          for (int y = 0; y < 64; ++y)
            for (int x = 0; x < 64; ++x)
              dst[(size_t)(ty + y) * pbw + tx + x] = (uint8_t)((int8_t)tile[y][x] + 128);
        }
    }
}

/* ------------------------------------------------------------------------------------
 * FastAC adaptive data model + encoder (entropy/arithmetic_codec.cpp:81-97, :360-387,
 * :498-509, :547-571, :749-829), alphabet 257 as in codec.cpp:192
 * ---------------------------------------------------------------------------------- */
enum { AC_SYMS = 257, AC_SHIFT = 15 };
typedef struct {
  uint32_t dist[AC_SYMS], count[AC_SYMS];
  uint32_t total, cycle, until;
} ac_model;

static void ac_model_update(ac_model *m) {
  if ((m->total += m->cycle) > (1u << AC_SHIFT)) {
    m->total = 0;
    for (int k = 0; k < AC_SYMS; ++k) m->total += (m->count[k] = (m->count[k] + 1) >> 1);
  }
  uint32_t sum = 0, scale = 0x80000000u / m->total;
  for (int k = 0; k < AC_SYMS; ++k) { m->dist[k] = (scale * sum) >> (31 - AC_SHIFT); sum += m->count[k]; }
  m->cycle = (5 * m->cycle) >> 2;
  uint32_t cap = (AC_SYMS + 6) << 3;
  if (m->cycle > cap) m->cycle = cap;
  m->until = m->cycle;
}

static void ac_model_init(ac_model *m) {
  m->total = 0; m->cycle = AC_SYMS;
  for (int k = 0; k < AC_SYMS; ++k) m->count[k] = 1;
  ac_model_update(m);
  m->until = m->cycle = (AC_SYMS + 6) >> 1;
}

static void ac_carry(uint8_t *p) { for (--p; *p == 0xFF; --p) *p = 0; ++*p; }

int mptc_oracle_arith_encode(const uint8_t *sym, int n, uint8_t *out, int cap) {
  ac_model *m = (ac_model *)malloc(sizeof *m);
  uint8_t *buf = (uint8_t *)malloc((size_t)n * 2 + 64), *p = buf;
  uint32_t base = 0, length = 0xFFFFFFFFu;
  ac_model_init(m);
  for (int i = 0; i < n; ++i) {
    uint32_t s = sym[i], x, b0 = base;
    if (s == AC_SYMS - 1) { x = m->dist[s] * (length >> AC_SHIFT); base += x; length -= x; }
    else { length >>= AC_SHIFT; x = m->dist[s] * length; base += x; length = m->dist[s + 1] * length - x; }
    if (b0 > base) ac_carry(p);
    if (length < 0x01000000u)
      do { *p++ = (uint8_t)(base >> 24); base <<= 8; } while ((length <<= 8) < 0x01000000u);
    ++m->count[s];
    if (--m->until == 0) ac_model_update(m);
  }
  uint32_t b0 = base;
  if (length > 2u * 0x01000000u) { base += 0x01000000u; length = 0x01000000u >> 1; }
  else { base += 0x01000000u >> 1; length = 0x01000000u >> 9; }
  if (b0 > base) ac_carry(p);
  do { *p++ = (uint8_t)(base >> 24); base <<= 8; } while ((length <<= 8) < 0x01000000u);
  int nbytes = (int)(p - buf);
  if (nbytes <= cap) memcpy(out, buf, (size_t)nbytes);
  free(buf); free(m);
  return nbytes <= cap ? nbytes : -nbytes;
}

/* Arithmetic_Codec::decode(Adaptive_Data_Model&) with start_decoder / renorm_dec_interval
 * (entropy/arithmetic_codec.cpp:100-105, :391-444, :511-522), as driven by EntropyDecode
 * (codec/codec.cpp:560-577).  The reference finds the symbol through its decoder table plus
 * bisection (:399-414); any search for the s with dist[s] <= value/length < dist[s+1] gives the
 * same symbol.  `code` must be readable up to code[nbytes + 3].  Returns 0, or -1 on overrun. */
int mptc_oracle_arith_decode(const uint8_t *code, int nbytes, uint8_t *sym_out, int n) {
  ac_model *m = (ac_model *)malloc(sizeof *m);
  ac_model_init(m);
  uint32_t length = 0xFFFFFFFFu;
  const uint8_t *p = code + 3, *end = code + nbytes + 4;
  uint32_t value = ((uint32_t)code[0] << 24) | ((uint32_t)code[1] << 16) | ((uint32_t)code[2] << 8) | code[3];
  int rc = 0;
  for (int i = 0; i < n; ++i) {
    uint32_t y = length;
    length >>= AC_SHIFT;
    uint32_t dv = value / length, s = 0, hi = AC_SYMS;   /* largest s with dist[s] <= dv */
    while (hi > s + 1) { uint32_t mid = (s + hi) >> 1; if (m->dist[mid] > dv) hi = mid; else s = mid; }
    uint32_t x = m->dist[s] * length;
    if (s != AC_SYMS - 1) y = m->dist[s + 1] * length;
    value -= x;
    length = y - x;
    if (length < 0x01000000u)
      do {
        if (p + 1 >= end) { rc = -1; break; }
        value = (value << 8) | *++p;
      } while ((length <<= 8) < 0x01000000u);
    if (rc) break;
    ++m->count[s];
    if (--m->until == 0) ac_model_update(m);
    sym_out[i] = (uint8_t)s;
  }
  free(m);
  return rc;
}

/* ------------------------------------------------------------------------------------
 * Decoder side of the endpoint planes: ReconstructEndPoints (codec/codec.cpp:697-800) =
 * MakeSigned (image_utils.h:243-266) -> IWavelet2D<.,64> (image_processing.h:337-399, levels
 * dim = 2..64, InverseWavelet2D rows-then-columns wavelet.cpp:133-155, InverseWavelet1D :64-95)
 * -> ycocg667_to_rgb565 (codec.cpp:46-63) -> 565 packing (:764-771).
 * planes = 6 x pbw*pbh symbols as written by mptc_oracle_endpoint_planes; ep1/ep2 = bw*bh u16.
 * ---------------------------------------------------------------------------------- */
static void unlift53(const int16_t *src, int16_t *dst, int len) { /* InverseWavelet1D, even len >= 2 */
  int mid = len - len / 2;
  for (int i = 0; i < len; i += 2) {
    int pv = (i == 0) ? 1 : i - 1;                        /* NormalizeIndex(-1) = 1 */
    int nx = (i + 1 < len) ? i + 1 : 2 * len - 2 - (i + 1);
    dst[i] = (int16_t)(src[i / 2] - (src[mid + pv / 2] + src[mid + nx / 2] + 2) / 4);
  }
  for (int i = 1; i < len; i += 2) {
    int nx = (i + 1 < len) ? i + 1 : 2 * len - 2 - (i + 1);
    dst[i] = (int16_t)(src[mid + i / 2] + (dst[i - 1] + dst[nx]) / 2);
  }
}

static void inverse_wavelet_tile(int16_t t[64][64]) {
  int16_t line[64], outl[64];
  for (int dim = 2; dim <= 64; dim <<= 1) {
    for (int r = 0; r < dim; ++r) { /* rows first (wavelet.cpp:141-144) */
      memcpy(line, t[r], sizeof(int16_t) * dim);
      unlift53(line, outl, dim);
      memcpy(t[r], outl, sizeof(int16_t) * dim);
    }
    for (int c = 0; c < dim; ++c) { /* then columns (:148-152) */
      for (int r = 0; r < dim; ++r) line[r] = t[r][c];
      unlift53(line, outl, dim);
      for (int r = 0; r < dim; ++r) t[r][c] = outl[r];
    }
  }
}

void mptc_oracle_inverse_planes(const uint8_t *planes, int bw, int bh, uint16_t *ep1, uint16_t *ep2) {
  int pbw = (bw + 63) / 64 * 64, pbh = (bh + 63) / 64 * 64;
  size_t pn = (size_t)pbw * pbh;
  int16_t tile[3][64][64];
  for (int ep = 0; ep < 2; ++ep) {
    uint16_t *out = ep ? ep2 : ep1;
    for (int ty = 0; ty < pbh; ty += 64)
      for (int tx = 0; tx < pbw; tx += 64) {
        for (int ch = 0; ch < 3; ++ch) {
          const uint8_t *src = planes + (size_t)(ep * 3 + ch) * pn;
          for (int y = 0; y < 64; ++y)
            for (int x = 0; x < 64; ++x)
              tile[ch][y][x] = (int16_t)(int8_t)(uint8_t)(src[(size_t)(ty + y) * pbw + tx + x] - 128);
          inverse_wavelet_tile(tile[ch]);
        }
        for (int y = 0; y < 64 && ty + y < bh; ++y)
          for (int x = 0; x < 64 && tx + x < bw; ++x) {
            int8_t yy = (int8_t)tile[0][y][x], co = (int8_t)tile[1][y][x], cg = (int8_t)tile[2][y][x];
            int8_t t = (int8_t)(yy - cg / 2);
            int8_t g = (int8_t)(cg + t), b = (int8_t)((t - co) / 2), r = (int8_t)(b + co);
            uint16_t v = (uint16_t)r;                      /* codec.cpp:764-771, as written */
            v = (uint16_t)(v << 6); v |= (uint16_t)g;
            v = (uint16_t)(v << 5); v |= (uint16_t)b;
            out[(size_t)(ty + y) * bw + tx + x] = v;
          }
      }
  }
}

/* ------------------------------------------------------------------------------------
 * The decoded picture: DXTImage::DecompressedImage (dxt_image.cpp:463-479) = GetColorAt over
 * SetLogicalBlocks / PhysicalToLogical (:198-227); transparent black of the 3-colour mode
 * reads as (0,0,0).  rgb_out: w*h*3 bytes, row-major.
 * ---------------------------------------------------------------------------------- */
void mptc_oracle_decode_rgb(const uint64_t *blocks, int w, int h, uint8_t *rgb_out) {
  int bw = w >> 2;
  pal4 p;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      uint64_t blk = blocks[(size_t)(y >> 2) * bw + (x >> 2)];
      palette_from_physical(blk, &p);
      const uint8_t *c = p.c[((uint32_t)(blk >> 32) >> (2 * ((y & 3) * 4 + (x & 3)))) & 3];
      memcpy(rgb_out + ((size_t)y * w + x) * 3, c, 3);
    }
}

/* ------------------------------------------------------------------------------------
 * PSNR of the decoded blocks (dxt_image.cpp:363-383 over PhysicalToLogical(blocks))
 * ---------------------------------------------------------------------------------- */
double mptc_oracle_psnr(const uint8_t *rgb, int w, int h, const uint64_t *blocks) {
  int bw = w >> 2, bh = h >> 2;
  double mse = 0.0;
  pal4 p;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      uint64_t blk = blocks[(size_t)(y >> 2) * bw + (x >> 2)];
      palette_from_physical(blk, &p);
      const uint8_t *c = p.c[((uint32_t)(blk >> 32) >> (2 * ((y & 3) * 4 + (x & 3)))) & 3];
      for (int ch = 0; ch < 3; ++ch) {
        double d = (double)rgb[((size_t)y * w + x) * 3 + ch] - (double)c[ch];
        mse += d * d;
      }
    }
  (void)bh;
  mse /= (double)(w * h);
  return 10.0 * log10((3.0 * 255.0 * 255.0) / mse);
}

/* ------------------------------------------------------------------------------------
 * CPU baseline driver: one GOP per worker thread (what ThreadedCompressMultiUnique does,
 * codec.cpp:1781-1793, minus its 5-thread cap); stages A+B only.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  const uint8_t *frames; int n_frames, w, h, gop, sa, thr;
  uint64_t *out_blocks; uint8_t *out_motion;
  int next_gop; pthread_mutex_t mu;
} gop_job;

static void *gop_worker(void *arg) {
  gop_job *J = (gop_job *)arg;
  int bw = J->w >> 2, bh = J->h >> 2;
  size_t nb = (size_t)bw * bh, fsz = (size_t)J->w * J->h * 3;
  uint64_t *cur = (uint64_t *)malloc(nb * 8), *prev = (uint64_t *)malloc(nb * 8);
  uint8_t *mo = (uint8_t *)malloc(nb * 2);
  uint32_t *un = (uint32_t *)malloc(nb * 4);
  for (;;) {
    pthread_mutex_lock(&J->mu);
    int g = J->next_gop++;
    pthread_mutex_unlock(&J->mu);
    int f0 = g * J->gop;
    if (f0 >= J->n_frames) break;
    for (int f = f0; f < f0 + J->gop && f < J->n_frames; ++f) {
      const uint8_t *rgb = J->frames + fsz * f;
      mptc_oracle_dxt1_fit(rgb, J->w, J->h, cur);
      mptc_oracle_reencode(rgb, J->w, J->h, f == f0, J->sa, J->thr, cur, prev, mo, un);
      if (J->out_blocks) memcpy(J->out_blocks + nb * f, cur, nb * 8);
      if (J->out_motion) memcpy(J->out_motion + 2 * nb * f, mo, nb * 2);
      uint64_t *t = cur; cur = prev; prev = t;
    }
  }
  free(cur); free(prev); free(mo); free(un);
  return NULL;
}

double mptc_oracle_encode_gops(const uint8_t *frames, int n_frames, int w, int h, int gop, int sa,
                               int thr, int threads, uint64_t *out_blocks, uint8_t *out_motion) {
  pthread_once(&g_tables_once, build_tables);
  gop_job J = {frames, n_frames, w, h, gop, sa, thr, out_blocks, out_motion, 0, PTHREAD_MUTEX_INITIALIZER};
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pthread_t th[256];
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, gop_worker, &J);
  for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
