// mptc_intra.cu -- K3: intra search (DXTImage::IntraSearch + the winner apply of Reencode,
// codec/dxt_image.cpp:652-713, :912-955) as a row-staggered wavefront over GROUPS of 32
// consecutive blocks.
//
// Dependency: block (x, y) reads the FINAL index words of (x-sa..x-1, y) and of
// (x-sa..x+sa-1, y-1..y-2sa+1).  One CTA walks one block row left to right, a group of 32
// targets at a time (lane = target):
//   1. wait until every row of the group's window has published enough final blocks;
//   2. load the union window's words, plus the group's own initial words (what a block keeps
//      when it turns out unique), and de-duplicate them in shared memory;
//   3. evaluate every distinct word once per target (warp = word, lane = target);
//   4. all warps scan the rows ABOVE for every target (those words are already final);
//   5. one warp resolves the targets in order: only the <= sa positions to the LEFT in the same
//      row depend on earlier decisions of this group, and their err_diff values are already in
//      the table, because a block's final word is either a word of its window or its own
//      initial word.  Each decision is a 16-position scan + a warp reduction.
// The index word of every decided block is published immediately (the high half of the 8-byte
// block); endpoints are refitted for the whole group afterwards (nobody waits on them).
// progress[f][y] = number of leading blocks of row y whose index words are final; rows are
// handed out in increasing order through a ticket, so a waiting CTA only ever depends on CTAs
// that already hold a ticket: no deadlock for any grid size.
#include "mptc_kernels.h"
#include "mptc_uniform_eval.cuh"

#include <cstdlib>

namespace mptc {

namespace {

constexpr int kG = 32;                        // targets per group
#ifndef MPTC_K3_THREADS
#define MPTC_K3_THREADS 512
#endif
constexpr int kThreads = MPTC_K3_THREADS, kWarps = kThreads / 32;
constexpr int kCtasPerSm = kThreads <= 256 ? 2 : 1;
constexpr int kMaxWords = 256;                // distinct words per group on the fast path
constexpr uint32_t kEmpty = 0xFFFFFFFFu;
constexpr uint16_t kNone = 0xFFFFu;
#ifndef MPTC_PUBLISH_EVERY
#define MPTC_PUBLISH_EVERY 8
#endif
constexpr int kPublishEvery = MPTC_PUBLISH_EVERY;   // decisions between progress publications (power of 2)
constexpr int kSparseTodo = 2;                // groups with this few targets take the direct path
constexpr int kChunkedTodo = 6;               // word-diverse groups with at least this many targets take the chunked path
#ifndef MPTC_NEAR_ROWS
#define MPTC_NEAR_ROWS 2
#endif
constexpr int kNear = MPTC_NEAR_ROWS;         // rows directly above that are consumed incrementally (<= 31)

struct GroupSmem {
  WordInfo *info;      // [kMaxWords]
  int *err;            // [kMaxWords + 1][33]; last row = rejected for every target
  uint8_t *lut5, *lut6;
  uint32_t *keys;      // [HT + 1]
  uint32_t *ulist;     // [kMaxWords + kG]
  uint16_t *pos_uid;   // [R][UW]: row 0 = the group's own row, row r = r rows above
  uint16_t *slot_uid;  // [HT + 1]
  uint32_t *ulist_all; // [NP + kG]: every distinct word of the window (word-diverse groups, chunked path)
};

__host__ __device__ inline int pow2_at_least(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

__host__ __device__ inline size_t group_smem_bytes(int sa, int *np_out, int *ht_out) {
  const int R = 2 * sa, UW = kG + 2 * sa - 1;
  const int NP = R * UW;
  const int HT = pow2_at_least(NP + kG + (NP + kG) / 4);
  if (np_out) *np_out = NP;
  if (ht_out) *ht_out = HT;
  size_t b = 0;
  b += (size_t)kMaxWords * sizeof(WordInfo);
  b += (size_t)(kMaxWords + 1) * 33 * sizeof(int);
  b += 512;
  b += (size_t)(HT + 1) * 4;
  b += (size_t)(kMaxWords + kG) * 4;
  b += (size_t)NP * 2;
  b += (size_t)(HT + 2) * 2;
  b += (size_t)(NP + kG) * 4;
  return (b + 15) & ~(size_t)15;
}

// Inserts `word` into the open-addressing set; the thread that claims a slot also hands out the
// word's dense id (ids >= kMaxWords only count: the group then takes the direct path).
__device__ __forceinline__ uint16_t wordset_insert(uint32_t *keys, uint32_t hmask, int hshift, int HT, int *special,
                                                   int *count, uint16_t *slot_uid, uint32_t *ulist, uint32_t word,
                                                   int cap = kMaxWords) {
  if (word == kEmpty) {
    if (atomicExch(special, 1) == 0) {
      const int uid = atomicAdd(count, 1);
      slot_uid[HT] = (uint16_t)uid;
      if (uid < cap) ulist[uid] = kEmpty;
    }
    return (uint16_t)HT;
  }
  uint32_t h = (word * 0x9E3779B1u) >> hshift;
  for (;;) {
    const uint32_t old = atomicCAS(&keys[h], kEmpty, word);
    if (old == kEmpty) {
      const int uid = atomicAdd(count, 1);
      slot_uid[h] = (uint16_t)uid;
      if (uid < cap) ulist[uid] = word;
      break;
    }
    if (old == word) break;
    h = (h + 1u) & hmask;
  }
  return (uint16_t)h;
}

__device__ __forceinline__ uint32_t ldcg_word(const uint64_t *blocks, size_t idx) {
  // index word = high half of the little-endian 8-byte block; L2 load (other CTAs write it)
  return __ldcg(reinterpret_cast<const uint32_t *>(blocks) + 2 * idx + 1);
}

#ifdef MPTC_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[16];
#define PHASE_MARK(i) do { if (tid == 0) { long long now_ = clock64(); atomicAdd(&g_phase_cycles[i], (unsigned long long)(now_ - t_mark_)); t_mark_ = now_; } } while (0)
#define PHASE_DECL long long t_mark_ = clock64()
#else
#define PHASE_MARK(i) do { } while (0)
#define PHASE_DECL do { } while (0)
#endif

}  // namespace

__global__ void __launch_bounds__(kThreads, kCtasPerSm)
k_intra_wavefront_tiled(SeqView v, int k_in_gop, int n_gops, int sa, int thr, int split, int *__restrict__ ticket) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_item, s_count, s_special, s_done, s_abort;
  __shared__ int s_avail[kNear + 1];   // [r]: leading blocks of row by-r that are in the word table
  __shared__ WinnerState s_partial[kG];
  __shared__ TargetCtx s_t;          // slow path only
  __shared__ WinnerState s_red[kWarps];

  const int W = 2 * sa, R = 2 * sa, UW = kG + 2 * sa - 1;
  int NP, HT;
  group_smem_bytes(sa, &NP, &HT);
  GroupSmem sm;
  {
    unsigned char *p = smem_raw;
    sm.info = reinterpret_cast<WordInfo *>(p);  p += (size_t)kMaxWords * sizeof(WordInfo);
    sm.err = reinterpret_cast<int *>(p);        p += (size_t)(kMaxWords + 1) * 33 * sizeof(int);
    sm.lut5 = p; sm.lut6 = p + 256;             p += 512;
    sm.keys = reinterpret_cast<uint32_t *>(p);  p += (size_t)(HT + 1) * 4;
    sm.ulist = reinterpret_cast<uint32_t *>(p); p += (size_t)(kMaxWords + kG) * 4;
    sm.pos_uid = reinterpret_cast<uint16_t *>(p); p += (size_t)NP * 2;
    sm.slot_uid = reinterpret_cast<uint16_t *>(p); p += (size_t)(HT + 2) * 2;
    sm.ulist_all = reinterpret_cast<uint32_t *>(p);
  }
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t hmask = (uint32_t)HT - 1u;
  const int hshift = 33 - __ffs(HT);
  const int n_items = n_gops * v.bh * split;   // one item = one row, or 1/split of its groups

  if (tid < 256) {
    sm.lut5[tid] = (uint8_t)snap_bits<0xF8, 4, 5>(tid);
    sm.lut6[tid] = (uint8_t)snap_bits<0xFC, 2, 6>(tid);
  }
  if (tid < 33) sm.err[kMaxWords * 33 + tid] = kRejectedSmall;

  if (k_in_gop > 0) {   // inter frames: anything left that K3s did not take?
    bool any = false;
    for (int g = 0; g < n_gops; ++g) {
      const int f = v.first + g * v.gop + k_in_gop;
      any = any || (f < v.first + v.count && v.n_unique[f] == kSparseNotHandled);
    }
    if (!any) return;
  }

  for (;;) {
    __syncthreads();
    if (tid == 0) s_item = atomicAdd(ticket, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= n_items) return;
    // the `split` CTAs of a row hold ADJACENT tickets: a part only ever waits for rows above (lower
    // tickets, handed out earlier) and for its partner, whose ticket is the next one to be taken
    const int part = item % split, gop_i = (item / split) % n_gops, by = item / (split * n_gops);
    const int f = v.first + gop_i * v.gop + k_in_gop;
    if (f >= v.first + v.count) continue;
    if (k_in_gop > 0 && v.n_unique[f] != kSparseNotHandled) continue;
    const uint8_t *frame = v.rgb + v.frame_bytes * f;
    uint64_t *cur = v.final_blocks + (size_t)f * v.nb;
    const uint64_t *init = v.init_blocks + (size_t)f * v.nb;
    const uint8_t *flags = v.flags + (size_t)f * v.nb;
    uint8_t *motion = v.motion + (size_t)f * v.nb * 2;
    int *progress = v.progress + (size_t)f * v.bh;
    // Inter frames: only rows in which the inter search left blocks over have work; every other
    // row is final already and counts as complete for its dependants without any publication.
    const bool all_rows = (k_in_gop == 0);
    const uint8_t *row_todo = v.row_todo + (size_t)f * v.bh;
    if (!all_rows && !row_todo[by]) continue;
    int published = 0;                 // tid 0: last value stored to progress[by]

    PHASE_DECL;
    // split > 1 (intra frames only): `split` CTAs share the row, CTA `part` takes every split-th
    // group.  The part of the own row that lies left of the group is then just another "near
    // row" (r = 0): the next group's tables are built while the neighbour CTA still decides.
    for (int x0 = part * kG; x0 < v.bw; x0 += split * kG) {
      const int x_end = min(x0 + kG, v.bw);
      PHASE_MARK(0);
      // ---- which blocks of the group still need the intra search ----------------------------
      const int gx = x0 + lane;
      const bool in_row = gx < v.bw;
      const bool todo = in_row && flags[(size_t)by * v.bw + gx] == 0;
      const unsigned todo_mask = __ballot_sync(0xffffffffu, todo);   // identical in every warp
      if (todo_mask == 0u) continue;    // nothing to do here; published lazily below / at the row end
      if (split == 1 && tid == 0 && published < x0) {  // blocks skipped so far are final
        st_release(progress + by, x0);
        published = x0;
      }

      // ---- wait for the FAR window rows; clear the word table meanwhile.  The kNear rows directly
      // above are not waited for: what they have published so far goes into the word table now,
      // the rest is consumed by the decider warp as it is published (near_update below), so a row
      // can follow the row above at a distance of ~sa blocks instead of a whole group + sa. ---------
      const bool sparse = __popc(todo_mask) <= kSparseTodo;
      const int need = min(x_end - 1 + sa, v.bw);
      if (wid == 0) {
        for (int base = sparse ? 1 : kNear + 1; base < R; base += 32) {
          const int r = base + lane;
          bool ok = r >= R || by - r < 0 || (!all_rows && !row_todo[by - r]);
          while (!__all_sync(0xffffffffu, ok)) {
            if (!ok) ok = ld_acquire(progress + by - r) >= need;
            if (!ok) __nanosleep(32);
          }
        }
        if (lane >= 1 && lane <= kNear) {
          const int r = lane;
          const bool complete = r >= R || by - r < 0 || (!all_rows && !row_todo[by - r]);
          s_avail[r] = (sparse || complete) ? need : min(ld_acquire(progress + by - r), need);
        } else if (lane == 0) {
          int a = x0;
          if (split > 1) {
            a = min(ld_acquire(progress + by), x0);
            while (sparse && a < x0) { __nanosleep(32); a = min(ld_acquire(progress + by), x0); }
          }
          s_avail[0] = a;
        }
      } else {
        for (int s = tid - 32; s <= HT; s += kThreads - 32) sm.keys[s] = kEmpty;
      }
      if (tid == 0) { s_count = 0; s_special = 0; s_done = 0; s_abort = -1; }
      __syncthreads();
      PHASE_MARK(1);   // wait for rows above

      // Few targets in the group (typical for the leftovers of an inter frame): evaluating every
      // distinct word for 32 lanes would cost more than evaluating their windows directly.
      int U = 0;
      LaneTarget t;
      if (!sparse) {
        // ---- load the union window (loads first, then the hash inserts) -------------------------
        for (int p0 = 0; p0 < NP; p0 += 4 * kThreads) {
          uint32_t wv[4];
          bool ok[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int p = p0 + q * kThreads + tid;
            const int r = p / UW, uc = p - r * UW;
            const int j = by - r, i = x0 - sa + uc;
            bool valid = p < NP && i >= 0 && i < v.bw && j >= 0;
            if (r == 0) valid = valid && (i < s_avail[0] || (i >= x0 && i < x_end && flags[(size_t)by * v.bw + i] != 0));
            if (r >= 1 && r <= kNear) valid = valid && i < s_avail[r];
            ok[q] = valid;
            wv[q] = valid ? ldcg_word(cur, (size_t)j * v.bw + i) : 0u;
          }
          if (p0 == 0) {   // the target's pixel loads go out behind the first batch of window loads
            if (in_row) {
              load_lane_target(t, frame, v.w, gx, by, init[(size_t)by * v.bw + gx]);
            } else {
#pragma unroll
              for (int k = 0; k < 48; ++k) t.pf[k] = 0.f;
#pragma unroll
              for (int k = 0; k < 12; ++k) t.pl[k] = 0u;
              t.own_block = 0; t.own_word = 0; t.orig_err = 0;
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int p = p0 + q * kThreads + tid;
            if (p < NP) sm.pos_uid[p] = ok[q] ? wordset_insert(sm.keys, hmask, hshift, HT, &s_special, &s_count, sm.slot_uid, sm.ulist, wv[q]) : kNone;
          }
        }
        __syncthreads();

        U = s_count;
        PHASE_MARK(2);   // window load + hash + ids
      }

      // ---- direct path: one target at a time, every window position evaluated (the same code as
      // the direct kernel).  Taken for sparse groups, when the window has little duplication
      // (noise), and to finish a group whose word table overflowed in the decider. ----------------
      auto direct_targets = [&](int g_begin, bool wait_near) {
        if (wait_near) {   // the direct evaluation reads every window row from global memory
          if (wid == 0 && lane >= 1 && lane <= kNear) {
            const int r = lane;
            const bool complete = r >= R || by - r < 0 || (!all_rows && !row_todo[by - r]);
            if (!complete)
              while (ld_acquire(progress + by - r) < need) __nanosleep(32);
          } else if (wid == 0 && lane == 0 && split > 1) {
            while (ld_acquire(progress + by) < x0) __nanosleep(32);
          }
          __syncthreads();
        }
        for (int g = g_begin; g < x_end - x0; ++g) {
          if (!((todo_mask >> g) & 1u)) continue;
          const int bx = x0 + g, b = by * v.bw + bx;
          if (tid == 0) build_target(s_t, frame, v.w, bx, by, init[b]);
          __syncthreads();
          WinnerState s;
          winner_init(s);
          for (int p = tid; p < W * W; p += kThreads) {
            const int row = p / W, col = p - row * W;
            const int j = by - row, i = bx + sa - 1 - col;
            if (i < 0 || j < 0 || i >= v.bw || (row == 0 && i >= bx)) continue;
            winner_update(s, eval_candidate(s_t, ldcg_word(cur, (size_t)j * v.bw + i)), row, col, W);
          }
          winner_block_reduce<kWarps>(s, s_red);
          if (tid == 0) {
            atomicAdd(v.work + kWorkIntraEvals, (unsigned long long)(W * W));   // every window position evaluated
            int row, col;
            const int min_err = winner_resolve(s, W, row, col);
            if (min_err <= thr) {
              cur[b] = winning_block(s_t, ldcg_word(cur, (size_t)(by - row) * v.bw + (bx + sa - 1 - col)));
              motion[2 * b + 0] = (uint8_t)(2 * sa - 1 - col);
              motion[2 * b + 1] = (uint8_t)(2 * sa - 1 - row);
            } else {
              motion[2 * b + 0] = 255;
              motion[2 * b + 1] = 255;
            }
            __threadfence();
            st_release(progress + by, bx + 1);
          }
          __syncthreads();
        }
        if (tid == 0) st_release(progress + by, x_end);
      };
      // ---- word-diverse group (more distinct words than the table holds: noise, err_threshold 0,
      // anything that keeps the frame's own ~distinct index words).  Same plan as the fast path,
      // restructured so that no table has to hold all words at once:
      //   1. wait until EVERYTHING the group's window can contain is final (no incremental near rows);
      //   2. reload the complete window, de-duplicate, all distinct words into ulist_all;
      //   3. rows above: chunks of kMaxWords words through the uniform evaluation (warp = word,
      //      lane = target) and a remapped window scan; then (3b) the <= sa blocks of the own row left
      //      of the group, once the neighbour CTA has decided them;
      //   4. the group's own blocks in order by one warp, lane = target: block g resolves from its
      //      winner state, then its final word is evaluated for the 32 lanes at once and pushed to
      //      the <= sa targets on its right (no table: the err_diff stays in a register).
      // ~30 us per group instead of ~600 us on the direct path. ------------------------------------
      auto overflow_group = [&]() {
        if (wid == 0) {
          if (lane >= 1 && lane <= kNear) {
            const int r = lane;
            const bool complete = r >= R || by - r < 0 || (!all_rows && !row_todo[by - r]);
            if (!complete)
              while (ld_acquire(progress + by - r) < need) __nanosleep(32);
          }
          __syncwarp();
          if (lane == 0) { s_count = 0; s_special = 0; }
        } else {
          for (int s = tid - 32; s <= HT; s += kThreads - 32) sm.keys[s] = kEmpty;
        }
        __syncthreads();
        for (int p0 = 0; p0 < NP; p0 += 4 * kThreads) {
          uint32_t wv[4];
          bool ok[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int p = p0 + q * kThreads + tid;
            const int r = p / UW, uc = p - r * UW;
            const int j = by - r, i = x0 - sa + uc;
            ok[q] = p < NP && i >= 0 && i < v.bw && j >= 0 && r > 0;   // the own row comes later (steps 3b, 4)
            wv[q] = ok[q] ? ldcg_word(cur, (size_t)j * v.bw + i) : 0u;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int p = p0 + q * kThreads + tid;
            if (p < NP) sm.pos_uid[p] = ok[q] ? wordset_insert(sm.keys, hmask, hshift, HT, &s_special, &s_count, sm.slot_uid, sm.ulist_all, wv[q], NP + kG) : kNone;
          }
        }
        __syncthreads();
        const int UA = s_count;
        if (tid == 0) {   // executed work: all distinct words + the pushed final words, once per block of the group
          const unsigned long long nn = (unsigned long long)(x_end - x0);
          atomicAdd(v.work + kWorkIntraEvals, ((unsigned long long)UA + (unsigned long long)min(sa, x0) + nn) * nn);
          atomicAdd(v.work + kWorkIntraScanned, (unsigned long long)__popc(todo_mask) * (unsigned long long)(min(R - 1, by) * W) *
                                                   (unsigned long long)((UA + kMaxWords - 1) / kMaxWords));
          atomicAdd(v.work + kWorkIntraGroups, 1ull);
        }
        for (int p = tid; p < NP; p += kThreads) {
          const uint16_t slot = sm.pos_uid[p];
          if (slot != kNone) sm.pos_uid[p] = sm.slot_uid[slot];   // kNone = 0xFFFF is outside every chunk
        }
        __syncthreads();
        constexpr int kPerWarp = kG / kWarps > 0 ? kG / kWarps : 1;
        WinnerState wsq[kPerWarp];
#pragma unroll
        for (int q = 0; q < kPerWarp; ++q) winner_init(wsq[q]);
        const int n = x_end - x0;
        for (int c0 = 0; c0 < UA; c0 += kMaxWords) {
          const int cn = min(kMaxWords, UA - c0);
          for (int u = tid; u < cn; u += kThreads) word_info(sm.ulist_all[c0 + u], sm.info[u]);
          __syncthreads();
          for (int u = wid; u < cn; u += kWarps)
            sm.err[u * 33 + lane] = eval_uniform(t, sm.ulist_all[c0 + u], sm.info[u], sm.lut5, sm.lut6);
          __syncthreads();
#pragma unroll
          for (int q = 0; q < kPerWarp; ++q) {
            const int g = wid + q * kWarps;
            if (g >= n || !((todo_mask >> g) & 1u)) continue;
            scan_window<true>(wsq[q], sm.pos_uid + g + W - 1, UW, -1, sm.err + g, W, 1, min(R - 1, by) + 1, lane, c0, cn, kMaxWords);
          }
          __syncthreads();
        }
#pragma unroll
        for (int q = 0; q < kPerWarp; ++q) {
          const int g = wid + q * kWarps;
          if (g >= n || !((todo_mask >> g) & 1u)) continue;
          winner_warp_reduce(wsq[q]);
          if (lane == 0) s_partial[g] = wsq[q];
        }
        // 3b. the own row left of the group: only now does the group wait for its neighbour CTA (two
        // CTAs per row on intra frames), so the chunk phase above overlapped the neighbour's decisions.
        // Its <= sa words are evaluated one per warp, lane = target, into rows 0.. of the err table.
        if (split > 1 && tid == 0)
          while (ld_acquire(progress + by) < x0) __nanosleep(32);
        __syncthreads();
        const int n_left = min(sa, x0);
        for (int j = wid; j < n_left; j += kWarps) {
          const uint32_t word = ldcg_word(cur, (size_t)by * v.bw + x0 - 1 - j);
          if (lane == 0) { sm.ulist[j] = word; word_info(word, sm.info[j]); }
          __syncwarp();
          sm.err[j * 33 + lane] = eval_uniform(t, word, sm.info[j], sm.lut5, sm.lut6);
        }
        __syncthreads();
        if (wid == 0) {
          WinnerState ws;
          winner_init(ws);
          if (todo) {
            ws = s_partial[lane];
            for (int j = 0; j < n_left && lane + 1 + j <= sa; ++j)   // block x0-1-j is lane+1+j to the left
              winner_update_fast(ws, sm.err[j * 33 + lane], (uint32_t)(sa + lane + j));
          }
          uint32_t final_word = (in_row && !todo) ? ldcg_word(cur, (size_t)by * v.bw + gx) : 0u;
          int dec = -1, stored = 0;
          uint32_t w_prev = 0u;
          int e_prev = kRejectedSmall;
          bool have_prev = false;
          for (int g = 0; g < n; ++g) {
            // lane g has seen all its candidates: resolve (every lane does, lane g's result counts)
            int row, col;
            const int min_err = winner_resolve_fast(ws, row, col);
            const bool fnd = todo && min_err <= thr;
            int gi = -1;
            uint32_t wt = 0u;
            if (fnd) {
              const int d = col - sa + 1;                       // row 0: the block d to the left
              if (row == 0 && lane - d >= 0) gi = lane - d;     // inside the group: that lane's final word
              else if (row == 0) wt = sm.ulist[d - lane - 1];   // left of the group (step 3b)
              else wt = sm.ulist_all[sm.pos_uid[row * UW + lane + W - 1 - col]];
            }
            const int gi_g = __shfl_sync(0xffffffffu, gi, g);
            const uint32_t w_in = __shfl_sync(0xffffffffu, final_word, gi_g >= 0 ? gi_g : 0);
            const uint32_t w_tab = __shfl_sync(0xffffffffu, wt, g);
            const uint32_t w_own = __shfl_sync(0xffffffffu, (in_row && !todo) ? final_word : t.own_word, g);
            const bool fnd_g = __shfl_sync(0xffffffffu, (int)fnd, g) != 0;
            const uint32_t w = fnd_g ? (gi_g >= 0 ? w_in : w_tab) : w_own;   // final word of block g
            if (lane == g) {
              final_word = w;
              dec = fnd ? ((row << 8) | col) : -1;
            }
            // push it to the <= sa undecided targets on its right
            const unsigned right = (todo_mask >> g) >> 1;
            if (right & ((sa >= 32) ? 0xffffffffu : ((1u << sa) - 1u))) {
              int e;
              if (have_prev && w == w_prev) e = e_prev;        // runs of the same word are common
              else {
                if (lane == 0) word_info(w, sm.info[0]);
                __syncwarp();
                e = eval_uniform(t, w, sm.info[0], sm.lut5, sm.lut6);
                __syncwarp();
                w_prev = w; e_prev = e; have_prev = true;
              }
              const int d = lane - g;
              const bool acc = d >= 1 && d <= sa && todo && e != kRejectedSmall;
              winner_update_fast(ws, acc ? e : kRejectedSmall, (uint32_t)(sa + d - 1));
            }
            if ((g & (kPublishEvery - 1)) == kPublishEvery - 1 || g == n - 1) {
              if (lane >= stored && lane <= g && todo)
                reinterpret_cast<uint32_t *>(cur)[2 * ((size_t)by * v.bw + x0 + lane) + 1] = final_word;
              stored = g + 1;
              __threadfence();
              __syncwarp();
              if (lane == 0) st_release(progress + by, x0 + g + 1);
            }
          }
          if (todo) {   // endpoints + motion: nobody waits on these inside the kernel
            const size_t b = (size_t)by * v.bw + gx;
            if (dec >= 0) {
              const int row = dec >> 8, col = dec & 0xFF;
              cur[b] = lane_winning_block(t, final_word);
              motion[2 * b + 0] = (uint8_t)(2 * sa - 1 - col);   // x = (i - bx) + sa
              motion[2 * b + 1] = (uint8_t)(2 * sa - 1 - row);   // y = (j - by) + 2sa - 1
            } else {
              motion[2 * b + 0] = 255;
              motion[2 * b + 1] = 255;
            }
          }
        }
        __syncthreads();
      };
      if (sparse) {
        direct_targets(0, false);
        continue;
      }
      if (U + kG > kMaxWords) {
        // the chunked path costs ~U evaluations x 32 lanes whatever the number of targets; a handful
        // of targets is cheaper one at a time (measured: leftovers of inter frames at err_threshold 0)
        if (__popc(todo_mask) >= kChunkedTodo) overflow_group();
        else direct_targets(0, true);
        continue;
      }

      // ---- ids per position, per-word constants ------------------------------------------------------
      for (int p = tid; p < NP; p += kThreads) {
        const uint16_t slot = sm.pos_uid[p];
        sm.pos_uid[p] = (slot != kNone) ? sm.slot_uid[slot] : (uint16_t)kMaxWords;
      }
      for (int u = tid; u < U; u += kThreads) word_info(sm.ulist[u], sm.info[u]);
      __syncthreads();
      PHASE_MARK(3);   // remap + word info

      // ---- evaluate: warp = one distinct word, lane = target ------------------------------------------
      for (int u = wid; u < U; u += kWarps)
        sm.err[u * 33 + lane] = eval_uniform(t, sm.ulist[u], sm.info[u], sm.lut5, sm.lut6);
      __syncthreads();
      PHASE_MARK(4);   // evaluate
#ifdef MPTC_PHASE_TIMING
      if (tid == 0) { atomicAdd(&g_phase_cycles[10], (unsigned long long)U); atomicAdd(&g_phase_cycles[11], 1ull); }
#endif

      // ---- everything that is already final: the rows above and the part of the own row that
      // lies left of the group.  Every target, all warps (scan order: j downwards, i downwards). ----
      for (int g = wid; g < x_end - x0; g += kWarps) {
        if (!((todo_mask >> g) & 1u)) continue;
        WinnerState ws;
        winner_init(ws);
        // uc = g + W - 1 - col: positions run right to left
        scan_window<false>(ws, sm.pos_uid + g + W - 1, UW, -1, sm.err + g, W, 1, min(R - 1, by) + 1, lane, 0, 0, kMaxWords);
        for (int l = lane; l < sa; l += 32)      // own row: position i = x0 + g - 1 - l, scan column sa + l
          if (l >= g && x0 + g - 1 - l >= 0)
            winner_update_fast(ws, sm.err[(int)sm.pos_uid[g + sa - 1 - l] * 33 + g], (uint32_t)(sa + l));
        winner_warp_reduce(ws);
        if (lane == 0) s_partial[g] = ws;
      }
      __syncthreads();
      PHASE_MARK(5);   // rows above

      // ---- the group's own row, in order, by one warp.  Lane l owns target l: its WinnerState and
      // its current candidate decision live in registers.  When block g is final its word is
      // PUSHED to the <= sa targets on its right (one table read + update per lane) and every lane
      // re-resolves its candidate, so step g only costs: shuffle lane g's candidate -> table read
      // -> update -> resolve.  Warp 1 publishes the progress counter (the device-scope fence is
      // off the critical path). ---------------------------------------------------------------------
      if (wid == 0) {
        constexpr int kNeedOwn = 0x7fff0000;
        const int n = x_end - x0;
        WinnerState ws;            // everything target `lane` has seen so far (rows above, near rows, pushes)
        winner_init(ws);
        if (todo) ws = s_partial[lane];
        // Lane state.  The pushes reach target l in DECREASING scan position (block g sits at
        // row 0, column sa + (l - g) - 1 of l's scan, and row 0 is scanned first), so every pushed
        // candidate is the earliest one seen so far.  With the reference's rule:
        //   err <= 0  -> it becomes the first non-positive candidate; the winner is then the
        //                "last row with a negative" candidate if one exists in a row above
        //                (fixed between near-row updates), else this candidate itself;
        //   err  > 0  -> it only matters while no non-positive candidate exists, and then wins
        //                ties against everything scanned later (err <= best so far).
        // A step is therefore a handful of selects; no shared-memory traffic besides the table read.
        int cand_uid = 0, cand_dec = 0, best_e = 0, ln_uid = 0, ln_dec = 0;
        bool found = true, has_first = false, ln_valid = false;
        // (Re)derives the select-only state of the still undecided lanes (lane >= g) from ws.
        auto derive = [&](int g) {
          if (todo && lane < g) return;   // decided: cand_* are final
          int row, col;
          const int min_err = winner_resolve_fast(ws, row, col);
          found = todo ? (min_err <= thr) : true;
          cand_dec = (row << 8) | col;
          cand_uid = todo ? (int)sm.pos_uid[found ? row * UW + lane + W - 1 - col : 0]
                          : (in_row ? (int)sm.pos_uid[sa + lane] : 0);   // already-final blocks of the group
          has_first = ws.first < 0x80000000u;
          best_e = min_err;                                 // only read while !has_first
          ln_valid = ws.lastneg >= 0 && (ws.lastneg >> 7) >= 1;
          const int lrow = ws.lastneg >> 7, lcol = 127 - (ws.lastneg & 127);
          ln_dec = (lrow << 8) | lcol;
          ln_uid = sm.pos_uid[ln_valid ? lrow * UW + lane + W - 1 - lcol : 0];
        };
        derive(0);
        const bool zero_ok = 0 <= thr;
        const int *err_lane = sm.err + lane;
        int stored = 0;           // blocks [0, stored) of the group have their words in global memory
        // Everything of step g after lane g's word id is known.
        auto finish_step = [&](int g, int uid, bool unique) {
          if (lane == g) {   // final from now on
            cand_uid = uid; found = true; cand_dec = unique ? -1 : cand_dec;
            sm.pos_uid[sa + g] = (uint16_t)uid;              // row 0 of the window, for later derive()s
          }
          const int d = lane - g;                            // push to the <= sa targets on the right
          const int e = err_lane[uid * 33];
          const bool acc = d >= 1 && d <= sa && todo && e != kRejectedSmall;
          const bool nonpos = acc && e <= 0;
          const bool better = acc && e > 0 && !has_first && e <= best_e;
          const int c = sa + d - 1;
          cand_uid = nonpos ? (ln_valid ? ln_uid : uid) : (better ? uid : cand_uid);
          cand_dec = nonpos ? (ln_valid ? ln_dec : c) : (better ? c : cand_dec);
          found = nonpos ? zero_ok : (better ? (e <= thr) : found);
          best_e = better ? e : best_e;
          has_first = has_first || nonpos;
          winner_update_fast(ws, acc ? e : kRejectedSmall, (uint32_t)c);   // off the critical chain
          if ((g & (kPublishEvery - 1)) == kPublishEvery - 1 || g == n - 1) {
            // index words of blocks [stored, g]: one coalesced store, then hand over to the publisher
            if (lane >= stored && lane <= g && todo)
              reinterpret_cast<uint32_t *>(cur)[2 * ((size_t)by * v.bw + x0 + lane) + 1] = sm.ulist[cand_uid];
            stored = g + 1;
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *reinterpret_cast<volatile int *>(&s_done) = g + 1;
          }
        };
        // Adds `word` (warp-uniform, known to be absent) to the word table and evaluates it for the
        // 32 targets.  hslot = its free hash slot (lane `leader`'s value counts), HT for 0xFFFFFFFF.
        auto add_word = [&](uint32_t word, int hslot, int leader) {
          const int uid = U++;
#ifdef MPTC_PHASE_TIMING
          if (lane == 0) atomicAdd(&g_phase_cycles[12], 1ull);
#endif
          if (lane == leader) {
            if (word == kEmpty) { s_special = 1; sm.slot_uid[HT] = (uint16_t)uid; }
            else { sm.keys[hslot] = word; sm.slot_uid[hslot] = (uint16_t)uid; }
            sm.ulist[uid] = word;
            word_info(word, sm.info[uid]);
          }
          __syncwarp();
          sm.err[uid * 33 + lane] = eval_uniform(t, word, sm.info[uid], sm.lut5, sm.lut6);
          return uid;
        };
        // Per-lane lookup: uid of `word`, or -1 with hslot = the free slot where it would go.
        auto probe = [&](uint32_t word, int &hslot) -> int {
          if (word == kEmpty) { hslot = HT; return s_special ? (int)sm.slot_uid[HT] : -1; }
          uint32_t h = (word * 0x9E3779B1u) >> hshift, kv;
          while ((kv = sm.keys[h]) != kEmpty && kv != word) h = (h + 1u) & hmask;
          hslot = (int)h;
          return kv == word ? (int)sm.slot_uid[h] : -1;
        };
        // Near rows: merged[r-1] = leading blocks of row by-r whose words this group has consumed.
        // merged[0]: the own row left of the group (all of it is needed before the first decision).
        int merged[kNear + 1];
#pragma unroll
        for (int r = 0; r <= kNear; ++r) merged[r] = s_avail[r];
        auto safe_end_of = [&]() {   // targets g < safe_end have all their near-row candidates
          if (merged[0] < x0) return 0;
          int m = merged[1];
#pragma unroll
          for (int r = 2; r <= kNear; ++r) m = min(m, merged[r]);
          return m >= need ? n : min(n, max(0, m - sa - x0 + 1));
        };
        // Fetches what the near rows have published since, until target g's window is complete.
        // Returns false if the word table is full.
        auto near_update = [&](int g) -> bool {
#ifdef MPTC_PHASE_TIMING
          const long long nu_t0 = clock64();
          struct NuTimer { long long t0; int lane; __device__ ~NuTimer() { if (lane == 0) { atomicAdd(&g_phase_cycles[13], (unsigned long long)(clock64() - t0)); atomicAdd(&g_phase_cycles[15], 1ull); } } } nu_timer{nu_t0, lane};
#endif
#pragma unroll
          for (int r = 0; r <= kNear; ++r) {
            const int need_g = r == 0 ? x0 : min(x0 + g + sa, v.bw);
            if (merged[r] >= need_g) continue;
            int p;
#ifdef MPTC_PHASE_TIMING
            const long long poll_t0 = clock64();
#endif
            for (;;) {
              p = ld_acquire(progress + by - r);
              if (p >= need_g) break;
              __nanosleep(32);
            }
#ifdef MPTC_PHASE_TIMING
            if (lane == 0) atomicAdd(&g_phase_cycles[14], (unsigned long long)(clock64() - poll_t0));
#endif
            p = min(p, r == 0 ? x0 : need);
            const int m0 = max(merged[r], max(x0 - sa, 0));
            const size_t rowbase = (size_t)(by - r) * v.bw;
            for (int c0 = m0; c0 < p; c0 += 32) {
              const int c = c0 + lane;
              const bool valid = c < p;
              const uint32_t word = valid ? ldcg_word(cur, rowbase + c) : 0u;
              int uid = -1, hslot = 0;
              for (;;) {
                if (valid && uid < 0) uid = probe(word, hslot);
                const unsigned newm = __ballot_sync(0xffffffffu, valid && uid < 0);
                if (newm == 0u) break;
                if (U >= kMaxWords) return false;
                const int leader = __ffs(newm) - 1;
                const uint32_t w = __shfl_sync(0xffffffffu, word, leader);
                const int nu = add_word(w, hslot, leader);
                if (valid && word == w) uid = nu;
                __syncwarp();
              }
              if (valid) sm.pos_uid[r * UW + c - (x0 - sa)] = (uint16_t)uid;
            }
            __syncwarp();
            if (todo && lane >= g) {
              const int lo = max(m0, x0 + lane - sa), hi = min(p, x0 + lane + sa);
              for (int c = lo; c < hi; ++c) {
                const int uid = sm.pos_uid[r * UW + c - (x0 - sa)];
                winner_update_fast(ws, err_lane[uid * 33], (uint32_t)((r << 7) | (x0 + lane + sa - 1 - c)));
              }
            }
            merged[r] = p;
          }
          return true;
        };
        int g = 0, safe_end = safe_end_of();
        bool aborted = false;
        while (g < n) {
          if (g >= safe_end) {
            if (!near_update(g)) { aborted = true; break; }
            __syncwarp();
            derive(g);
            safe_end = safe_end_of();
          }
          // Hot loop: a lone warp is bound by instruction latency, so it is kept short and
          // straight-line.  Leaves as soon as a block turns out unique.
          int uid = 0;
          for (; g < safe_end; ++g) {
            uid = __shfl_sync(0xffffffffu, found ? cand_uid : kNeedOwn, g);   // lane g has all its pushes
            if (uid == kNeedOwn) break;
            finish_step(g, uid, false);
          }
          if (g >= safe_end) continue;
          // Rare: block g keeps its own initial word, which later targets may reuse; look it up /
          // add it to the word table and evaluate it for the 32 targets.  Warp-uniform.
          const uint32_t word = __shfl_sync(0xffffffffu, t.own_word, g);
          int hslot;
          uid = probe(word, hslot);
          __syncwarp();                              // every lane has read before the table changes
          if (uid < 0) {
            if (U >= kMaxWords) { aborted = true; break; }
            uid = add_word(word, hslot, 0);
          }
          finish_step(g, uid, true);
          ++g;
        }
        const int g_end = g;       // == n unless the word table overflowed
        if (aborted) {
          // hand blocks [0, g) over, then let every warp finish the group on the direct path
          if (lane >= stored && lane < g && todo)
            reinterpret_cast<uint32_t *>(cur)[2 * ((size_t)by * v.bw + x0 + lane) + 1] = sm.ulist[cand_uid];
          __threadfence_block();
          __syncwarp();
          if (lane == 0) {
            *reinterpret_cast<volatile int *>(&s_done) = g;
            __threadfence_block();
            *reinterpret_cast<volatile int *>(&s_abort) = g;
          }
        }
        PHASE_MARK(7);   // in-row decisions
        if (lane == 0) {   // executed work: every word of the table once per block of the group; rows above + own row scanned
          atomicAdd(v.work + kWorkIntraEvals, (unsigned long long)U * (unsigned long long)n);
          atomicAdd(v.work + kWorkIntraScanned, (unsigned long long)__popc(todo_mask) * (unsigned long long)(min(R - 1, by) * W + sa));
          atomicAdd(v.work + kWorkIntraGroups, 1ull);
        }
        // ---- endpoints + motion for the decided blocks (nobody waits on these inside the kernel) ------
        if (todo && lane < g_end) {
          const size_t b = (size_t)by * v.bw + gx;
          if (cand_dec >= 0) {
            const int row = cand_dec >> 8, col = cand_dec & 0xFF;
            cur[b] = lane_winning_block(t, sm.ulist[cand_uid]);
            motion[2 * b + 0] = (uint8_t)(2 * sa - 1 - col);   // x = (i - bx) + sa
            motion[2 * b + 1] = (uint8_t)(2 * sa - 1 - row);   // y = (j - by) + 2sa - 1
          } else {
            motion[2 * b + 0] = 255;
            motion[2 * b + 1] = 255;
          }
        }
      } else if (wid == 1 && lane == 0) {
        // publisher: turns the decider's block-scope hand-over into device-scope progress
        const int n = x_end - x0;
        for (int last = 0;;) {
          const int d = *reinterpret_cast<volatile int *>(&s_done);
          const int ab = *reinterpret_cast<volatile int *>(&s_abort);
          if (d > last) {
            __threadfence();
            st_release(progress + by, x0 + d);
            last = d;
          } else if (last >= n || (ab >= 0 && last >= ab)) {
            break;
          } else {
            __nanosleep(64);   // do not hammer the shared-memory pipe the decider warp depends on
          }
        }
      }
      __syncthreads();
      if (s_abort >= 0) {          // word table overflow: the rest of the group, one target at a time
        const int ab = s_abort;
        __syncthreads();
        direct_targets(ab, true);
        continue;
      }
      PHASE_MARK(6);   // in-row resolve + write
    }
    if (split == 1 && tid == 0) st_release(progress + by, v.bw);   // covers trailing groups that had nothing to do
  }
}

#ifdef MPTC_PHASE_TIMING
extern "C" void mptc_debug_phase_cycles(unsigned long long *out16, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, g_phase_cycles, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_phase_cycles, z, sizeof z); }
}
#endif

bool launch_intra_wavefront_tiled(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, int *ticket,
                                  int grid_cap, cudaStream_t s) {
  static int max_optin = -1;
  static int max_ctas_dev[kMaxDevices] = {0};
  static size_t configured_dev[kMaxDevices] = {0};   // per device: one context per GPU may live in one process
  static int split_intra = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t bytes = group_smem_bytes(sa, nullptr, nullptr);
  int max_ctas = 0;
  {
    std::lock_guard<std::mutex> lock(launch_cfg_mutex());
    size_t &configured = configured_dev[dev & (kMaxDevices - 1)];
    if (max_optin < 0) cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (bytes + 4096 > (size_t)max_optin) return false;
    if (bytes > configured) {
      if (cudaFuncSetAttribute(k_intra_wavefront_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
        return false;
      configured = bytes;
      int per_sm = 0, sms = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_intra_wavefront_tiled, kThreads, bytes);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      max_ctas_dev[dev & (kMaxDevices - 1)] = (per_sm < 1 ? 1 : per_sm) * sms;
    }
    max_ctas = max_ctas_dev[dev & (kMaxDevices - 1)];
    if (split_intra < 0) {
      const char *e = getenv("MPTC_ROW_SPLIT");
      split_intra = (e && *e) ? atoi(e) : 2;
      if (split_intra < 1) split_intra = 1;
    }
  }
  // Intra frames: `split` CTAs per row (see the kernel).  A CTA of a row waits for its neighbours,
  // whose tickets are adjacent to its own (item = (row * n_gops + gop) * split + part), so `split`
  // resident CTAs are enough whatever n_gops is.
  int split = k_in_gop == 0 ? split_intra : 1;
  if (grid_cap > 0 && grid_cap < split) split = 1;
  const int items = n_gops * v.bh * split;
  int grid = items < max_ctas ? items : max_ctas;
  if (grid_cap > 0 && grid > grid_cap) grid = grid_cap;
  if (grid < split) split = 1;
  k_intra_wavefront_tiled<<<grid, kThreads, bytes, s>>>(v, k_in_gop, n_gops, sa, thr, split, ticket);
  return true;
}

}  // namespace mptc
