// mptc_inter_wide.cu -- K2, second generation: 16x16 targets per CTA, one 8x4 sub-tile per warp.
//
// Same restatement as mptc_inter.cu (DXTImage::InterBlockSearch + the winner apply in Reencode,
// codec/dxt_image.cpp:715-774, :885-908) and the same three ideas -- de-duplicate the window's index
// words, evaluate each DISTINCT word once per target with the word warp-uniform (mptc_uniform_eval.cuh),
// then let every target walk its own window through the (word id -> err_diff) table with the
// order-independent winner scan (WinnerState) -- but arranged so that the per-tile
// overhead is paid once per 256 targets instead of once per 32:
//
//   * K2 of round 1 gave a CTA 32 targets: all eight warps held THE SAME 32 targets in registers (eight
//     copies of the pixel loads, conversions and own-error computation), shared the words of one 39x35
//     union window (one hash build, one id pass, one set of per-word constants per 32 targets) and met at
//     seven barriers per tile.  Measured per tile (profiles/r2_k2_phases.txt): 33 % of the time in the
//     phases before the first evaluation.
//   * Here a CTA takes 16x16 targets.  Warp w owns the 8x4 sub-tile (w & 1, w >> 1): its 32 targets live in
//     its lanes' registers and nowhere else.  The hash build, the id pass and the per-word constants cover
//     the 47x47 union window of all 256 targets (2209 positions instead of 8 x 1365), after which each
//     warp works alone: it marks the words that occur in ITS sub-tile's 39x35 window, evaluates those
//     against its 32 targets into a warp-private table, scans its targets' windows, and applies its
//     winners -- no CTA barrier after the per-word constants.
//   * The table holds {-1, 0, min(err_diff, max)} (the winner rule needs the sign of a negative err_diff and
//     exact positive values only up to the threshold, dxt_image.cpp:890): int8 entries (max 127) for
//     thresholds below 127 -- 224 words x 36 targets x 8 warps fit two CTAs per SM --, int16 entries
//     (max 32767, 128 words) above.  (Thresholds >= 32767 keep round 1's kernel and its int32 table.)
//   * The scan walks COLUMNS of targets: the four targets (x, y0 .. y0+3) of a sub-tile column look at the
//     same union-window positions one row apart, so one id read and one 8-byte table read (the four
//     targets' entries of a word are adjacent: table[word][x][y]) serve four window positions.  The
//     scan is bound by shared-memory wavefronts (random-bank table reads), not by issue slots: this cuts
//     them by about three.
//   * A tile whose union window holds more distinct words than the table (word-diverse content: noise,
//     err_threshold 0) is handed, sub-tile by sub-tile, to round 1's tile search (mptc_inter_tile.cuh),
//     whose window -- and therefore word count -- is that of 32 targets.  (A second launch that spread
//     such tiles' sub-tiles over the GPU was measured and dropped: with four lanes of frames in flight it
//     only lengthened every lane's chain of kernels; profiles/r2_k2_wide.txt.)
//
// Results are bit-identical to the position-by-position loop (tests/test_gpu_parity_small.py,
// tests/test_gpu_full_golden.py run through whichever K2 the launcher picks; MPTC_K2=tiled|wide forces one).
#include "mptc_inter_tile.cuh"

#include <cstdio>
#include <cstdlib>

namespace mptc {

namespace {

constexpr int kSubX = 8, kSubY = 4;            // a warp's targets: lane = ly * 8 + lx
constexpr int kSubsX = 2, kSubsY = 4;          // sub-tiles per CTA
constexpr int kTileX = kSubX * kSubsX, kTileY = kSubY * kSubsY;   // 16 x 16 targets
constexpr int kWarps = kSubsX * kSubsY, kThreads = kWarps * 32;
constexpr int kBatch = 9;                      // window loads in flight per thread (2209 positions / 256 threads at search_area 16)
constexpr int kRow = 36;                       // table entries per word: [lx][ly] + 4 pad; rows of 9 / 18 words: an odd multiple of 1 / 2 banks
constexpr uint32_t kEmpty = 0xFFFFFFFFu;       // hash-table empty marker (the real word 0xFFFFFFFF lives in slot HT)
constexpr uint16_t kNoPos = 0xFFFFu;           // window position outside the frame

// Table entry types: E = int8_t for err_threshold < 127, int16_t below 32767.
template <typename E> struct Table;
template <> struct Table<int8_t>  { static constexpr int kWords = 224, kMax = 127; };     // word ids fit the uint8 lists
template <> struct Table<int16_t> { static constexpr int kWords = 128, kMax = 32767; };

__host__ __device__ inline int wide_round_up_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

struct WideLayout {
  int NP, HT;
  size_t off_keys, off_slot_uid;                       // phases 0-2 ...
  size_t off_info, off_err, off_need, off_wlist;       // ... share their bytes with phases 3-5
  size_t err_per_warp, need_per_warp;
  size_t off_lut, off_ulist, off_pos, bytes;
};

template <typename E>
__host__ __device__ inline WideLayout wide_layout(int sa) {
  constexpr int kWords = Table<E>::kWords;
  WideLayout L;
  const int UW = 2 * sa + kTileX - 1, UH = 2 * sa + kTileY - 1;
  L.NP = UW * UH;
  L.HT = wide_round_up_pow2(L.NP + L.NP / 4);
  const size_t a = (((size_t)(L.HT + 1) * 4 + (size_t)(L.HT + 1) * 2) + 15) & ~(size_t)15;
  L.err_per_warp = ((size_t)(kWords + 1) * kRow * sizeof(E) + 15) & ~(size_t)15;   // + the all-rejected row
  L.need_per_warp = (size_t)kWords + 16;                                           // + the dummy id's flag
  L.off_keys = 0;
  L.off_slot_uid = (size_t)(L.HT + 1) * 4;
  L.off_info = 0;
  L.off_err = (size_t)kWords * sizeof(WordInfo);
  L.off_need = L.off_err + kWarps * L.err_per_warp;
  L.off_wlist = L.off_need + kWarps * L.need_per_warp;
  const size_t e = L.off_wlist + (size_t)kWarps * kWords;
  size_t b = a > e ? a : e;
  L.off_lut = b;   b += 512;
  L.off_ulist = b; b += (size_t)(kWords + 2) * 4;      // only the words of a tile that stays on this path
  L.off_pos = b;   b += (size_t)L.NP * 2;
  L.bytes = (b + 15) & ~(size_t)15;
  const size_t t32 = tile32::tile_smem_bytes(sa, nullptr, nullptr);   // the word-diverse path's carve-up of the same bytes
  if (t32 > L.bytes) L.bytes = t32;
  return L;
}

// One union-window row of a column of four targets: target y's window row is R - y.  e = the four
// targets' table entries of the word at (R, column); kMask = which of the four have row R in their window.
template <int kMask>
__device__ __forceinline__ void column_step(WinnerState (&ws)[4], uint2 e, uint32_t p) {   // int16 entries
  if (kMask & 1) winner_update_fast(ws[0], (int)(int16_t)(e.x & 0xFFFFu), p);
  if (kMask & 2) winner_update_fast(ws[1], (int)e.x >> 16, p);
  if (kMask & 4) winner_update_fast(ws[2], (int)(int16_t)(e.y & 0xFFFFu), p);
  if (kMask & 8) winner_update_fast(ws[3], (int)e.y >> 16, p);
}
__device__ __forceinline__ uint2 load_entries(const int16_t *row) { return *reinterpret_cast<const uint2 *>(row); }
__device__ __forceinline__ uint32_t load_entries(const int8_t *row) { return *reinterpret_cast<const uint32_t *>(row); }

// Table entry of an err_diff.  int16: {-1, 0, min(e, 32767)}.  int8: bit 7 = "negative", bits 0-6 = min(max(e, 0), 127),
// i.e. 0x80 for a negative err_diff -- read as a signed byte that is still negative, which is all winner_update_fast
// needs -- and the form the packed scan below takes apart.
__device__ __forceinline__ int16_t encode_entry(int e, int16_t) { return (int16_t)(e < 0 ? -1 : min(e, 32767)); }
__device__ __forceinline__ int8_t encode_entry(int e, int8_t) { return (int8_t)(e < 0 ? -128 : min(e, 127)); }

// The int8 column scan keeps, per lane (= window column) and for the four targets of a sub-tile column at once:
//   key  (16 bits per target, two registers): min over the rows of (c << 6 | R), c = the entry's low 7 bits, R = the
//        union-window row.  A lane that saw an err_diff <= 0 ends with c == 0 and R = its first such row (the
//        reference's `first`); otherwise with its smallest positive err_diff and that one's first row (`best`) --
//        `best` is only consulted when no candidate is <= 0, so one minimum serves both.
//   last (8 bits per target, one register): R + 1 of the last row with a negative entry, 0 = none (rows ascend, so the
//        newest negative row simply replaces the byte).
// 18 instructions per four window positions instead of 39 with one WinnerState per target.
struct PackedColumn {
  uint32_t key01, key23, last;
};
template <int kMask>   // which of the four targets have union row R inside their window
__device__ __forceinline__ void packed_step(PackedColumn &pc, uint32_t w, uint32_t rr, uint32_t r1) {
  if (kMask != 15) {   // rows above / below a target's window: "rejected, not negative"
    constexpr uint32_t keep = (kMask & 1 ? 0xFFu : 0u) | (kMask & 2 ? 0xFF00u : 0u) | (kMask & 4 ? 0xFF0000u : 0u) | (kMask & 8 ? 0xFF000000u : 0u);
    w = (w & keep) | (0x7F7F7F7Fu & ~keep);
  }
  const uint32_t h01 = __byte_perm(w, 0u, 0x4140), h23 = __byte_perm(w, 0u, 0x4342);   // bytes -> halves
  pc.key01 = __vminu2(pc.key01, (h01 & 0x007F007Fu) * 64u + rr);
  pc.key23 = __vminu2(pc.key23, (h23 & 0x007F007Fu) * 64u + rr);
  const uint32_t n = (w & 0x80808080u) >> 7;             // 1 per byte with a negative entry
  pc.last = (pc.last & ~(n * 255u)) | (n * r1);
}
// One target's result of a packed column scan as a WinnerState in the target's own rows (y = its row in the sub-tile).
__device__ __forceinline__ WinnerState unpack_column(const PackedColumn &pc, int y, int lane) {
  const uint32_t key = ((y & 2) ? pc.key23 : pc.key01) >> (16 * (y & 1)) & 0xFFFFu;
  const uint32_t lastb = (pc.last >> (8 * y)) & 0xFFu;
  const uint32_t c = key >> 6, p = (((key & 63u) - (uint32_t)y) << 7) | (uint32_t)lane;
  WinnerState ws;
  ws.first = c == 0u ? p : 0xFFFFFFFFu;
  ws.best = ((c + 65536u) << 14) | p;
  ws.lastneg = lastb ? (int)(((lastb - 1u - (uint32_t)y) << 7) | (127u - (uint32_t)lane)) : -1;
  return ws;
}

}  // namespace

#ifdef MPTC_K2_PHASE_TIMING
__device__ unsigned long long g_k2w_cycles[12];
#define K2W_MARK(i) do { if (tid == 0) { const long long now_ = clock64(); atomicAdd(&g_k2w_cycles[i], (unsigned long long)(now_ - t_mark_)); t_mark_ = now_; } } while (0)
extern "C" void mptc_debug_k2w_cycles(unsigned long long *out12, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out12, g_k2w_cycles, sizeof(unsigned long long) * 12);
  if (reset) { unsigned long long z[12] = {0}; cudaMemcpyToSymbol(g_k2w_cycles, z, sizeof z); }
}
// per CTA since the last reset, in order of completion: {start ns, end ns, SM | distinct words << 16 | tile << 32}
constexpr unsigned kTraceCap = 1u << 16;
__device__ unsigned long long g_k2w_trace[3 * kTraceCap];
__device__ unsigned g_k2w_trace_n;
extern "C" int mptc_debug_k2w_trace(unsigned long long *out, int cap_entries, int reset) {
  cudaDeviceSynchronize();
  unsigned n = 0;
  cudaMemcpyFromSymbol(&n, g_k2w_trace_n, sizeof n);
  if (n > kTraceCap) n = kTraceCap;
  if ((int)n > cap_entries) n = (unsigned)cap_entries;
  if (out && n) cudaMemcpyFromSymbol(out, g_k2w_trace, sizeof(unsigned long long) * 3 * n);
  if (reset) { const unsigned z = 0; cudaMemcpyToSymbol(g_k2w_trace_n, &z, sizeof z); }
  return (int)n;
}
__device__ __forceinline__ unsigned long long k2w_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned k2w_smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#define K2W_TRACE_BEGIN() const unsigned long long trace_t0_ = k2w_now()
#define K2W_TRACE_END(words) do { if (tid == 0) { const unsigned i_ = atomicAdd(&g_k2w_trace_n, 1u); if (i_ < kTraceCap) { g_k2w_trace[3 * i_] = trace_t0_; \
    g_k2w_trace[3 * i_ + 1] = k2w_now(); g_k2w_trace[3 * i_ + 2] = k2w_smid() | ((unsigned long long)(words) << 16) | ((unsigned long long)blockIdx.x << 32); } } } while (0)
#else
#define K2W_MARK(i) do { } while (0)
#define K2W_TRACE_BEGIN() do { } while (0)
#define K2W_TRACE_END(words) do { } while (0)
#endif

template <typename E>
__global__ void __launch_bounds__(kThreads, 2)
k_inter_search_wide(SeqView v, int k_in_gop, int sa, int thr) {
  constexpr int kWords = Table<E>::kWords, kMax = Table<E>::kMax;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_count, s_special;

  const int W = 2 * sa;
  const int UW = W + kTileX - 1;
  const WideLayout L = wide_layout<E>(sa);
  const int NP = L.NP, HT = L.HT;
  uint32_t *const keys = reinterpret_cast<uint32_t *>(smem_raw + L.off_keys);
  uint16_t *const slot_uid = reinterpret_cast<uint16_t *>(smem_raw + L.off_slot_uid);
  WordInfo *const info = reinterpret_cast<WordInfo *>(smem_raw + L.off_info);
  uint8_t *const lut5 = smem_raw + L.off_lut, *const lut6 = lut5 + 256;
  uint32_t *const ulist = reinterpret_cast<uint32_t *>(smem_raw + L.off_ulist);
  uint16_t *const pos_uid = reinterpret_cast<uint16_t *>(smem_raw + L.off_pos);

  const int f = v.first + blockIdx.y * v.gop + k_in_gop;
  if (f >= v.first + v.count) return;
  const int tiles_x = (v.bw + kTileX - 1) / kTileX;
  const int tx0 = (blockIdx.x % tiles_x) * kTileX, ty0 = (blockIdx.x / tiles_x) * kTileY;
  const int ux0 = tx0 - sa, uy0 = ty0 - sa;   // union-window origin in block coordinates
  const uint64_t *prev = v.final_blocks + (size_t)(f - 1) * v.nb;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // this warp's sub-tile inside the CTA's tile, and this lane's target
  const int sx0 = (wid & (kSubsX - 1)) * kSubX, sy0 = (wid / kSubsX) * kSubY;
  const int lx = lane & (kSubX - 1), ly = lane >> 3;
  const int tbx = tx0 + sx0 + lx, tby = ty0 + sy0 + ly;
  const bool t_valid = tbx < v.bw && tby < v.bh;
  const int tb = tby * v.bw + tbx;
  const int n_valid = max(0, min(kSubX, v.bw - tx0 - sx0)) * max(0, min(kSubY, v.bh - ty0 - sy0));   // warp-uniform
  E *const err = reinterpret_cast<E *>(smem_raw + L.off_err + (size_t)wid * L.err_per_warp);
  uint8_t *const need = smem_raw + L.off_need + (size_t)wid * L.need_per_warp;
  uint8_t *const wlist = smem_raw + L.off_wlist + (size_t)wid * kWords;

#ifdef MPTC_K2_PHASE_TIMING
  long long t_mark_ = clock64();
#endif
  K2W_TRACE_BEGIN();
  // ---- phase 0: clear the hash table -------------------------------------------------------
  for (int s = tid; s <= HT; s += kThreads) keys[s] = kEmpty;
  lut5[tid] = (uint8_t)snap_bits<0xF8, 4, 5>(tid);   // kThreads == 256
  lut6[tid] = (uint8_t)snap_bits<0xFC, 2, 6>(tid);
  if (tid == 0) { s_count = 0; s_special = 0; }
  __syncthreads();
  K2W_MARK(0);   // clear

  // ---- phase 1: load the union window (all loads of a thread first, so their latencies overlap),
  // this lane's target block, then insert the words --------------------------------------------
  const uint32_t hmask = (uint32_t)HT - 1u;
  const int hshift = 32 - __ffs(HT) + 1;   // HT = 2^(ffs-1)
  const uint32_t uw_magic = 0xFFFFFFFFu / (uint32_t)UW + 1u;   // p / UW == umulhi(p, magic) for p < 2^16
  LaneTarget t;
  for (int p0 = 0; p0 < NP; p0 += kBatch * kThreads) {
    uint32_t wv[kBatch];
    bool ok[kBatch];
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      const int p = p0 + q * kThreads + tid;
      const int ur = (int)__umulhi((uint32_t)p, uw_magic), uc = p - ur * UW;
      const int i = ux0 + uc, j = uy0 + ur;
      ok[q] = p < NP && i >= 0 && j >= 0 && i < v.bw && j < v.bh;
      wv[q] = ok[q] ? (uint32_t)(__ldg(prev + (size_t)j * v.bw + i) >> 32) : 0u;
    }
    if (p0 == 0) {   // the target's pixel loads go out behind the first batch of window loads
      if (t_valid) {
        load_lane_target(t, v.rgb + v.frame_bytes * f, v.w, tbx, tby, v.init_blocks[(size_t)f * v.nb + tb]);
      } else {
#pragma unroll
        for (int k = 0; k < 48; ++k) t.pf[k] = 0.f;
#pragma unroll
        for (int k = 0; k < 12; ++k) t.pl[k] = 0u;
        t.own_block = 0; t.own_word = 0; t.orig_err = 0;
      }
    }
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      const int p = p0 + q * kThreads + tid;
      if (p >= NP) break;
      uint16_t slot = kNoPos;
      if (ok[q]) {
        const uint32_t word = wv[q];
        // the thread that claims a slot also hands out the word's dense id
        if (word == kEmpty) {
          if (atomicExch(&s_special, 1) == 0) {
            const int uid = atomicAdd(&s_count, 1);
            slot_uid[HT] = (uint16_t)uid;
            if (uid <= kWords) ulist[uid] = kEmpty;
          }
          slot = (uint16_t)HT;
        } else {
          uint32_t h = (word * 0x9E3779B1u) >> hshift;
          for (;;) {
            const uint32_t old = atomicCAS(&keys[h], kEmpty, word);
            if (old == kEmpty) {
              const int uid = atomicAdd(&s_count, 1);
              slot_uid[h] = (uint16_t)uid;
              if (uid <= kWords) ulist[uid] = word;
              break;
            }
            if (old == word) break;
            h = (h + 1u) & hmask;
          }
          slot = (uint16_t)h;
        }
      }
      pos_uid[p] = slot;
    }
  }
  __syncthreads();
  K2W_MARK(1);   // window load + hash

  const int U = s_count;
  if (U > kWords) {
    // word-diverse tile: round 1's search, one 8x4 sub-tile after the other, by the whole CTA
    for (int sub_i = 0; sub_i < kWarps; ++sub_i) {
      const int bx0 = tx0 + (sub_i & (kSubsX - 1)) * kSubX, by0 = ty0 + (sub_i / kSubsX) * kSubY;
      if (bx0 >= v.bw || by0 >= v.bh) continue;
      __syncthreads();                         // the shared memory changes hands
      tile32::search_tile(v, f, bx0, by0, sa, thr, smem_raw);
    }
    K2W_TRACE_END(U);
    return;
  }

  // ---- phase 2: position -> dense word id (out-of-frame positions: the all-rejected row) ----------
  for (int p = tid; p < NP; p += kThreads) {
    const uint16_t slot = pos_uid[p];
    pos_uid[p] = (slot != kNoPos) ? slot_uid[slot] : (uint16_t)kWords;
  }
  __syncthreads();   // slot_uid / keys are dead from here on: their bytes become info / err / need / wlist
  K2W_MARK(2);   // ids

  // ---- phase 3: per-word constants for the CTA; each warp lists the words of ITS sub-tile's window ----
  for (int u = tid; u < U; u += kThreads) word_info(ulist[u], info[u]);
  for (int u = lane; u < kRow; u += 32) err[kWords * kRow + u] = (E)kMax;   // the all-rejected row
  const uint16_t *const sub = pos_uid + sy0 * UW + sx0;   // the sub-tile's window: (W + 7) x (W + 3) positions
  const int SW = W + kSubX - 1, SH = W + kSubY - 1;
  int n_need = 0;
  if (n_valid > 0) {
    for (int u = lane; u < (int)L.need_per_warp / 4; u += 32) reinterpret_cast<uint32_t *>(need)[u] = 0u;
    __syncwarp();
    // plain byte stores (equal values may race); the dummy id kWords has a flag of its own
    for (int c = lane; c < SW; c += 32) {
      const uint16_t *q = sub + c;
#pragma unroll 5
      for (int r = 0; r < SH; ++r) need[q[r * UW]] = 1;
    }
    __syncwarp();
    for (int base = 0; base < U; base += 32) {
      const bool nd = base + lane < U && need[base + lane];
      const unsigned m = __ballot_sync(0xffffffffu, nd);
      if (nd) wlist[n_need + __popc(m & ((1u << lane) - 1u))] = (uint8_t)(base + lane);
      n_need += __popc(m);
    }
  }
  __syncthreads();   // the constants are there (and, for the warp, its list)
  K2W_MARK(3);   // per-word constants + lists

  if (n_valid > 0) {
    // ---- phase 4: evaluate the listed words against this warp's 32 targets ------------------------
    E *const my_err = err + lx * kSubY + ly;
    for (int i = 0; i < n_need; ++i) {
      const int u = wlist[i];
      const int e = eval_uniform(t, ulist[u], info[u], lut5, lut6);
      my_err[u * kRow] = encode_entry(e, E());
    }
    __syncwarp();
    K2W_MARK(4);   // evaluation (warp 0's)

    // ---- phase 5: scan.  Every target walks its own window in the reference's order (j up, i up) ----
    WinnerState mine;          // lane = target
    winner_init(mine);
    if (W == 32) {
      // lane = window column; the four targets of a sub-tile column share every id and table read.
#pragma unroll 1
      for (int ttx = 0; ttx < kSubX; ++ttx) {
        if (tx0 + sx0 + ttx >= v.bw) break;    // warp-uniform
        const uint16_t *q = sub + ttx + lane;
        const E *tab = err + ttx * kSubY;
        auto entries = [&](int R) { return load_entries(tab + (int)q[R * UW] * kRow); };
        if constexpr (sizeof(E) == 1) {
          PackedColumn pc = {0xFFFFFFFFu, 0xFFFFFFFFu, 0u};
          uint32_t rr = 0u, r1 = 1u;           // R in both halves; R + 1
          packed_step<1>(pc, entries(0), rr, r1);  rr += 0x00010001u; ++r1;
          packed_step<3>(pc, entries(1), rr, r1);  rr += 0x00010001u; ++r1;
          packed_step<7>(pc, entries(2), rr, r1);  rr += 0x00010001u; ++r1;
#pragma unroll 4
          for (int R = 3; R < 32; ++R, rr += 0x00010001u, ++r1) packed_step<15>(pc, entries(R), rr, r1);
          packed_step<14>(pc, entries(32), rr, r1);  rr += 0x00010001u; ++r1;
          packed_step<12>(pc, entries(33), rr, r1);  rr += 0x00010001u; ++r1;
          packed_step<8>(pc, entries(34), rr, r1);
#pragma unroll
          for (int y = 0; y < 4; ++y) {
            if (ty0 + sy0 + y >= v.bh) break;    // warp-uniform
            WinnerState ws = unpack_column(pc, y, lane);
            winner_warp_reduce(ws);
            if (lane == y * kSubX + ttx) mine = ws;
          }
        } else {
          // Positions are keyed by the UNION row R (target y's own row is R - y: subtracted after the reduction).
          WinnerState ws[4];
#pragma unroll
          for (int y = 0; y < 4; ++y) winner_init(ws[y]);
          uint32_t p = (uint32_t)lane;
          column_step<1>(ws, entries(0), p);  p += 128u;
          column_step<3>(ws, entries(1), p);  p += 128u;
          column_step<7>(ws, entries(2), p);  p += 128u;
#pragma unroll 4
          for (int R = 3; R < 32; ++R, p += 128u) column_step<15>(ws, entries(R), p);
          column_step<14>(ws, entries(32), p);  p += 128u;
          column_step<12>(ws, entries(33), p);  p += 128u;
          column_step<8>(ws, entries(34), p);
#pragma unroll
          for (int y = 0; y < 4; ++y) {
            if (ty0 + sy0 + y >= v.bh) break;    // warp-uniform
            winner_warp_reduce(ws[y]);
            if (lane == y * kSubX + ttx) {       // back to the target's own rows
              const uint32_t d = (uint32_t)y << 7;
              ws[y].first -= d;  ws[y].lastneg -= (int)d;  ws[y].best -= d;
              mine = ws[y];
            }
          }
        }
      }
    } else {
#pragma unroll 1
      for (int tt = 0; tt < kSubX * kSubY; ++tt) {
        const int ttx = tt & (kSubX - 1), tty = tt >> 3;
        if (tx0 + sx0 + ttx >= v.bw || ty0 + sy0 + tty >= v.bh) continue;   // warp-uniform
        WinnerState ws;
        winner_init(ws);
        scan_window<false, E, kRow>(ws, sub + tty * UW + ttx, UW, 1, err + ttx * kSubY + tty, W, 0, W, lane, 0, 0, kWords);
        winner_warp_reduce(ws);
        if (lane == tt) mine = ws;
      }
    }
    K2W_MARK(5);   // window scan (warp 0's)

    // ---- phase 6: resolve and apply (lane = target) ---------------------------------------------
    if (t_valid) {
      int row, col;
      const int min_err = winner_resolve_fast(mine, row, col);
      uint8_t flag = 0;
      if (min_err <= thr) {
        const uint32_t word = ulist[sub[(ly + row) * UW + lx + col]];
        v.final_blocks[(size_t)f * v.nb + tb] = lane_winning_block(t, word);
        v.motion[((size_t)f * v.nb + tb) * 2 + 0] = (uint8_t)(col | 0x80);   // x = (i - bx) + sa
        v.motion[((size_t)f * v.nb + tb) * 2 + 1] = (uint8_t)(row | 0x80);   // y = (j - by) + sa
        flag = 1;
      }
      v.flags[(size_t)f * v.nb + tb] = flag;
      if (!flag) v.row_todo[(size_t)f * v.bh + tby] = 1;   // the intra wavefront has work in this row
    }
    // executed work (bench.py's roofline): every word a warp evaluated once per valid target, every
    // window position of every valid target once
    if (lane == 0) {
      atomicAdd(v.work + kWorkInterEvals, (unsigned long long)n_need * (unsigned long long)n_valid);
      atomicAdd(v.work + kWorkInterScanned, (unsigned long long)n_valid * (unsigned long long)(W * W));
      if (wid == 0) atomicAdd(v.work + kWorkInterTiles, 1ull);
    }
    K2W_MARK(6);   // resolve + apply
  }
  K2W_TRACE_END(U);   // (warp 0's end)
}

// Returns false when the kernel does not apply (nothing launched): the caller then uses the round-1 tiling.
// two_per_sm_only: decline unless two CTAs fit an SM.
template <typename E>
static bool launch_wide(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, bool two_per_sm_only, cudaStream_t s) {
  static int max_optin = -1, max_sm = -1;
  static size_t configured[kMaxDevices] = {0};
  const WideLayout L = wide_layout<E>(sa);
  if (L.NP >= 0xFFFF) return false;            // 16-bit word ids
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  {
    std::lock_guard<std::mutex> lock(launch_cfg_mutex());
    if (max_optin < 0) {
      cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cur_dev);
      cudaDeviceGetAttribute(&max_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, cur_dev);
    }
    if (L.bytes + 1024 > (size_t)max_optin) return false;
    if (two_per_sm_only && 2 * (L.bytes + 1024 + 64) > (size_t)max_sm) return false;
    size_t &conf = configured[cur_dev & (kMaxDevices - 1)];
    if (L.bytes > conf) {
      if (cudaFuncSetAttribute(k_inter_search_wide<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.bytes) != cudaSuccess)
        return false;
      // two CTAs need nearly all of the SM's shared memory: ask for the largest carve-out outright
      cudaFuncSetAttribute(k_inter_search_wide<E>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (getenv("MPTC_DEBUG_OCC")) {
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_inter_search_wide<E>, kThreads, L.bytes);
        fprintf(stderr, "mptc: k_inter_search_wide<int%d>: %zu B dynamic shared memory, %d CTAs per SM (sa %d)\n",
                (int)sizeof(E) * 8, L.bytes, per_sm, sa);
      }
      conf = L.bytes;
    }
  }
  const int tiles = ((v.bw + kTileX - 1) / kTileX) * ((v.bh + kTileY - 1) / kTileY);
  k_inter_search_wide<E><<<dim3(tiles, n_gops), kThreads, L.bytes, s>>>(v, k_in_gop, sa, thr);
  return true;
}

template <typename E>
static bool wide_fits_twice(int sa) {
  int dev = 0, max_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  return 2 * (wide_layout<E>(sa).bytes + 1024 + 64) <= (size_t)max_sm;
}

bool launch_inter_search_wide(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, bool two_per_sm_only, cudaStream_t s) {
  static const bool no_int8 = getenv("MPTC_K2_NO_INT8") != nullptr;   // A/B measurements
  // The window decides, not the table type: where the 224-word int8 layout no longer fits twice on an SM (search
  // areas above ~20) the 128-word int16 one still would, but nearly every tile of such a window overflows 128
  // words and is handed to the 8x4 tile search -- search area 32 measured 15.4 (threshold 50, falling through
  // from int8 to int16) and 24.4 (threshold 200) instead of 12.6 ms per GOP with round 1's kernel alone.
  if (two_per_sm_only && !wide_fits_twice<int8_t>(sa)) return false;
  if (thr < Table<int8_t>::kMax && !no_int8) return launch_wide<int8_t>(v, k_in_gop, n_gops, sa, thr, two_per_sm_only, s);
  if (thr < Table<int16_t>::kMax) return launch_wide<int16_t>(v, k_in_gop, n_gops, sa, thr, two_per_sm_only, s);
  return false;
}

}  // namespace mptc
