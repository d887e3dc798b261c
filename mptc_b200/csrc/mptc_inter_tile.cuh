// mptc_inter_tile.cuh -- the 8x4-target tile search of K2 (round 1's tiling), as a device function: the
// body of k_inter_search_tiled (mptc_inter.cu) and the path k_inter_search_wide (mptc_inter_wide.cu)
// takes for word-diverse tiles.  See mptc_inter.cu for the description.
#pragma once
#include "mptc_kernels.h"
#include "mptc_uniform_eval.cuh"

#include <type_traits>

namespace mptc {
namespace tile32 {


constexpr int kTileX = 8, kTileY = 4;         // 32 targets per CTA: lane = ty*8 + tx
constexpr int kThreads = 256, kWarps = kThreads / 32;
constexpr int kChunk = 128;                   // distinct words evaluated per pass (WordInfo / epk scratch)
// Words per SCAN pass.  The winner rule only needs the sign of a negative err_diff, and exact positive
// values only up to the threshold (a candidate above it can never be "found", dxt_image.cpp:890), so
// with err_threshold < 32767 the (word, target) table holds int16 {-1, 0, min(err_diff, 32767)} and
// twice the words fit the same shared memory: word-diverse tiles (large windows, err_threshold 0,
// noise) walk their window scan half as often.  Larger thresholds keep the exact int32 table.
template <bool kErr16> struct ErrTable { typedef int type; static constexpr int kWords = kChunk; };
template <> struct ErrTable<true> { typedef int16_t type; static constexpr int kWords = 2 * kChunk; };
constexpr int kErr16Max = 32767;
constexpr uint32_t kEmpty = 0xFFFFFFFFu;      // hash-table empty marker (the real word
                                              // 0xFFFFFFFF lives in the extra slot HT)
constexpr uint16_t kNoPos = 0xFFFFu;          // window position outside the frame

struct TileSmem {
  // dynamic shared memory carve-up (all sizes depend on search_area)
  uint16_t *pos_uid;   // [NP]   hash slot, then dense id of that position's word
  uint32_t *keys;      // [HT+1] open-addressing table of words
  uint16_t *slot_uid;  // [HT+1]
  uint32_t *ulist;     // [NP]   dense list of distinct words
  WordInfo *info;      // [kChunk]
  void *err;           // [kWords + 1][33] int32 or int16; last row = "rejected" for every target
  uint32_t *epk;       // [kChunk][33]; the refitted endpoints of (word, target), packed 565 | 565 << 16
  uint8_t *lut5, *lut6;  // ToFiveBits / ToSixBits tables
};

__host__ __device__ inline int round_up_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

__host__ __device__ inline size_t tile_smem_bytes(int sa, int *np_out, int *ht_out) {
  const int UW = 2 * sa + kTileX - 1, UH = 2 * sa + kTileY - 1;
  const int NP = UW * UH;
  const int HT = round_up_pow2(NP + NP / 4);
  if (np_out) *np_out = NP;
  if (ht_out) *ht_out = HT;
  // {keys, slot_uid} are dead once every position has its dense word id (phase 2); {info, err,
  // epk} are only written after that: the two groups share the same bytes, which (with the window's words
  // no longer kept: the winner's word is ulist[its id]) lets two CTAs per SM run at search_area 32.
  const size_t a = (((size_t)(HT + 1) * 4 + (size_t)(HT + 1) * 2) + 15) & ~(size_t)15;   // keys, slot_uid
  const size_t e = (size_t)kChunk * sizeof(WordInfo) + (size_t)(kChunk + 1) * 33 * sizeof(int) + (size_t)kChunk * 33 * sizeof(uint32_t);
  size_t b = a > e ? a : e;
  b += 512;                                        // lut5, lut6
  b += (size_t)NP * 4;                             // ulist
  b += (size_t)NP * 2;                             // pos_uid
  return (b + 15) & ~(size_t)15;
}

#ifdef MPTC_K2_PHASE_TIMING
static __device__ unsigned long long g_k2_cycles[12];   // per translation unit (no relocatable device code)
#define K2_MARK(i) do { if (tid == 0) { const long long now_ = clock64(); atomicAdd(&g_k2_cycles[i], (unsigned long long)(now_ - t_mark_)); t_mark_ = now_; } } while (0)
#else
#define K2_MARK(i) do { } while (0)
#endif

// One 8x4 tile of targets of frame f (f >= 1 within its GOP), by a whole CTA of kThreads threads.
__device__ __forceinline__ void search_tile(const SeqView &v, int f, int tx0, int ty0, int sa, int thr, unsigned char *smem_raw) {
  __shared__ int s_count, s_special;
  __shared__ int s_res_err[kTileX * kTileY], s_res_pos[kTileX * kTileY];

  const int W = 2 * sa;
  const int UW = W + kTileX - 1, UH = W + kTileY - 1;
  int NP, HT;
  tile_smem_bytes(sa, &NP, &HT);
  TileSmem sm;
  {
    unsigned char *p = smem_raw;
    const size_t a = (((size_t)(HT + 1) * 4 + (size_t)(HT + 1) * 2) + 15) & ~(size_t)15;
    const size_t e = (size_t)kChunk * sizeof(WordInfo) + (size_t)(kChunk + 1) * 33 * sizeof(int) + (size_t)kChunk * 33 * sizeof(uint32_t);
    // phases 3-6 (16-byte aligned first) ...
    sm.info = reinterpret_cast<WordInfo *>(p);
    sm.err = p + (size_t)kChunk * sizeof(WordInfo);   // (kChunk + 1) * 33 int32 == (2 kChunk + 1) * 33 int16 rounded up
    sm.epk = reinterpret_cast<uint32_t *>(p + (size_t)kChunk * sizeof(WordInfo) + (size_t)(kChunk + 1) * 33 * sizeof(int));
    // ... over the same bytes as phases 0-2
    sm.keys = reinterpret_cast<uint32_t *>(p);
    sm.slot_uid = reinterpret_cast<uint16_t *>(p + (size_t)(HT + 1) * 4);
    p += a > e ? a : e;
    sm.lut5 = p; sm.lut6 = p + 256;            p += 512;
    sm.ulist = reinterpret_cast<uint32_t *>(p); p += (size_t)NP * 4;
    sm.pos_uid = reinterpret_cast<uint16_t *>(p);
  }

  const int ux0 = tx0 - sa, uy0 = ty0 - sa;   // union-window origin in block coordinates
  const uint64_t *prev = v.final_blocks + (size_t)(f - 1) * v.nb;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

#ifdef MPTC_K2_PHASE_TIMING
  long long t_mark_ = clock64();
#endif
  // ---- phase 0: clear the hash table -------------------------------------------------------
  for (int s = tid; s <= HT; s += kThreads) sm.keys[s] = kEmpty;
  sm.lut5[tid] = (uint8_t)snap_bits<0xF8, 4, 5>(tid);   // kThreads == 256
  sm.lut6[tid] = (uint8_t)snap_bits<0xFC, 2, 6>(tid);
  if (tid == 0) { s_count = 0; s_special = 0; }
  __syncthreads();
  K2_MARK(0);   // clear

  // ---- phase 1: load the union window (all loads of a thread first, so their latencies overlap),
  // this lane's target block, then insert the words --------------------------------------------
  const uint32_t hmask = (uint32_t)HT - 1u;
  const int hshift = 32 - __ffs(HT) + 1;   // HT = 2^(ffs-1)
  const uint32_t uw_magic = 0xFFFFFFFFu / (uint32_t)UW + 1u;   // p / UW == umulhi(p, magic) for p < 2^16

  // this lane's target block (every warp holds the same 32 targets)
  const int tbx = tx0 + (lane & (kTileX - 1)), tby = ty0 + (lane >> 3);
  const bool t_valid = tbx < v.bw && tby < v.bh;
  const int tb = tby * v.bw + tbx;
  LaneTarget t;
  for (int p0 = 0; p0 < NP; p0 += 8 * kThreads) {
    uint32_t wv[8];
    bool ok[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
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
    // (Measured and dropped, profiles/r2_k2_phases.txt: electing one lane per distinct word of a warp with
    // __match_any_sync before the insert, and looking at the slot before paying for the atomic -- this
    // phase is the latency of the window and pixel loads, not the atomics.)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int p = p0 + q * kThreads + tid;
      if (p >= NP) break;
      uint16_t slot = kNoPos;
      if (ok[q]) {
        const uint32_t word = wv[q];
        // the thread that claims a slot also hands out the word's dense id
        if (word == kEmpty) {
          if (atomicExch(&s_special, 1) == 0) {
            const int uid = atomicAdd(&s_count, 1);
            sm.slot_uid[HT] = (uint16_t)uid;
            sm.ulist[uid] = kEmpty;
          }
          slot = (uint16_t)HT;
        } else {
          uint32_t h = (word * 0x9E3779B1u) >> hshift;
          for (;;) {
            const uint32_t old = atomicCAS(&sm.keys[h], kEmpty, word);
            if (old == kEmpty) {
              const int uid = atomicAdd(&s_count, 1);
              sm.slot_uid[h] = (uint16_t)uid;
              sm.ulist[uid] = word;
              break;
            }
            if (old == word) break;
            h = (h + 1u) & hmask;
          }
          slot = (uint16_t)h;
        }
      }
      sm.pos_uid[p] = slot;
    }
  }
  __syncthreads();
  K2_MARK(1);   // window load + hash

  const int U = s_count;
  WinnerState ws[kTileX * kTileY / kWarps];   // this warp scans targets wid, wid+8, wid+16, wid+24
#pragma unroll
  for (int q = 0; q < kTileX * kTileY / kWarps; ++q) winner_init(ws[q]);
  // Phases 2-5 for one representation of the (word, target) table.  Tiles whose words fit one
  // evaluation chunk -- all but a handful on ordinary content -- keep the exact int32 table (this path
  // is the headline's and must not pay for the other); word-diverse tiles use the int16 table, which
  // holds twice the words per scan pass (see ErrTable).
  auto run = [&](auto tag) {
    constexpr bool kErr16 = decltype(tag)::value;
    typedef typename ErrTable<kErr16>::type E;
    constexpr int kWords = ErrTable<kErr16>::kWords;
    E *const err = static_cast<E *>(sm.err);
    // ---- phase 2: position -> dense word id ------------------------------------------------------
    for (int p = tid; p < NP; p += kThreads) {
      const uint16_t slot = sm.pos_uid[p];
      // out-of-frame positions: the all-rejected row when everything fits one chunk, otherwise an
      // id no chunk contains
      sm.pos_uid[p] = (slot != kNoPos) ? sm.slot_uid[slot] : (uint16_t)(U <= kWords ? kWords : kNoPos);
    }
    __syncthreads();   // slot_uid / keys are dead from here on: their bytes become info / err / epk
    K2_MARK(2);   // ids
    if (tid < 33) err[kWords * 33 + tid] = (E)(kErr16 ? kErr16Max : kRejectedSmall);

    // ---- phases 3-5 per chunk of distinct words ----------------------------------------------

    for (int c0 = 0; c0 < U; c0 += kWords) {
      const int cn = min(kWords, U - c0);
      // evaluate in sub-chunks of kChunk words (the per-word constants' scratch): warp = one distinct word, lane = target
      for (int s0 = 0; s0 < cn; s0 += kChunk) {
        const int sn = min(kChunk, cn - s0);
        if (s0 > 0) __syncthreads();            // the previous sub-chunk's constants are no longer read
        for (int u = tid; u < sn; u += kThreads) word_info(sm.ulist[c0 + s0 + u], sm.info[u]);
        __syncthreads();
        K2_MARK(3);   // per-word constants
        for (int u = wid; u < sn; u += kWarps) {
          const uint32_t word = sm.ulist[c0 + s0 + u];
          uint32_t packed;
          const int e = eval_uniform(t, word, sm.info[u], sm.lut5, sm.lut6, &packed);
          err[(s0 + u) * 33 + lane] = (E)(kErr16 ? (e < 0 ? -1 : min(e, kErr16Max)) : e);
          if (c0 + s0 == 0) sm.epk[u * 33 + lane] = packed;   // the first kChunk words: enough when U <= kChunk
        }
      }
      __syncthreads();
      K2_MARK(4);   // evaluation

      // scan: each target walks its own window in the reference's order (j up, i up).
      // Positions outside the frame (and, when the words do not fit one chunk, words of other
      // chunks) read the all-rejected row kWords.
      const bool single = (U <= kWords);
#pragma unroll
      for (int q = 0; q < kTileX * kTileY / kWarps; ++q) {
        const int tt = wid + q * kWarps;
        const int ttx = tt & (kTileX - 1), tty = tt >> 3;
        if (tx0 + ttx >= v.bw || ty0 + tty >= v.bh) continue;   // warp-uniform
        const uint16_t *pos = sm.pos_uid + tty * UW + ttx;
        if (single) scan_window<false, E>(ws[q], pos, UW, 1, err + tt, W, 0, W, lane, 0, 0, kWords);
        else        scan_window<true, E>(ws[q], pos, UW, 1, err + tt, W, 0, W, lane, c0, cn, kWords);
      }
      __syncthreads();
      K2_MARK(5);   // window scan
    }

    // executed work (bench.py's roofline): every distinct word once per valid target, every window
    // position of every valid target once per chunk pass
    if (tid == 0) {
      const unsigned long long nt = (unsigned long long)(min(kTileX, v.bw - tx0) * min(kTileY, v.bh - ty0));
      atomicAdd(v.work + kWorkInterEvals, nt * (unsigned long long)U);
      atomicAdd(v.work + kWorkInterScanned, nt * (unsigned long long)(W * W) * (unsigned long long)((U + kWords - 1) / kWords));
      atomicAdd(v.work + kWorkInterTiles, 1ull);
    }
  };
  if (U <= kChunk || thr >= kErr16Max) run(std::false_type());
  else run(std::true_type());

  // ---- phase 6: resolve and apply ------------------------------------------------------------
#pragma unroll
  for (int q = 0; q < kTileX * kTileY / kWarps; ++q) {
    winner_warp_reduce(ws[q]);
    if (lane == 0) {
      int row, col;
      const int tt = wid + q * kWarps;
      s_res_err[tt] = winner_resolve_fast(ws[q], row, col);
      s_res_pos[tt] = (row << 8) | col;
    }
  }
  __syncthreads();
  K2_MARK(6);   // resolve
  if (wid == 0 && t_valid) {
    const int min_err = s_res_err[lane];
    const int row = s_res_pos[lane] >> 8, col = s_res_pos[lane] & 0xFF;
    uint8_t flag = 0;
    if (min_err <= thr) {
      const int at = ((lane >> 3) + row) * UW + (lane & (kTileX - 1)) + col;
      const uint32_t word = sm.ulist[sm.pos_uid[at]];   // (the window's words themselves are gone: see tile_smem_bytes)
      // the winner's endpoints were computed when its word was evaluated: with a single chunk of
      // words the table still holds them and the block needs no second refit
      uint64_t blk;
      if (word == t.own_word) blk = t.own_block;
      else if (U <= kChunk) blk = (uint64_t)sm.epk[(int)sm.pos_uid[at] * 33 + lane] | ((uint64_t)word << 32);
      else blk = lane_winning_block(t, word);
      v.final_blocks[(size_t)f * v.nb + tb] = blk;
      v.motion[((size_t)f * v.nb + tb) * 2 + 0] = (uint8_t)(col | 0x80);   // x = (i - bx) + sa
      v.motion[((size_t)f * v.nb + tb) * 2 + 1] = (uint8_t)(row | 0x80);   // y = (j - by) + sa
      flag = 1;
    }
    v.flags[(size_t)f * v.nb + tb] = flag;
    if (!flag) v.row_todo[(size_t)f * v.bh + tby] = 1;   // the intra wavefront has work in this row
  }
  K2_MARK(7);   // apply
}

}  // namespace tile32
}  // namespace mptc
