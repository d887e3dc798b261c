// mptc_sparse.cu -- K3s: intra search for the blocks of an INTER frame that the inter search
// left over (DXTImage::Reencode, codec/dxt_image.cpp:910-955, reached when :890 fails).
//
// On ordinary content these are a few dozen to a few thousand blocks per frame, scattered over
// the frame, so the row wavefront (mptc_intra_rows.cu) would spend its time handing rows from CTA to
// CTA.  Here every leftover block is one work item and dependencies are tracked per BLOCK:
//   * every CTA builds the frame's raster-ordered directory of leftover blocks itself: the rows
//     flagged by K2 in row_todo and a prefix sum of their leftover counts (item -> row by binary
//     search, -> column by a ballot scan of that row's flags), so there is no cap on the number
//     of items and no list in global memory;
//   * items are handed out in raster order through a ticket; a target's window only contains
//     blocks that precede it in raster order, so a CTA only ever waits for items whose tickets
//     were taken before its own: no deadlock for any grid size;
//   * a window position whose flag is still 0 (leftover, undecided) is polled until its owner
//     publishes flag = 2 behind a fence; everything else is final already.
// Per item the window's index words are de-duplicated in shared memory (as in the tiled kernels)
// and every DISTINCT word is evaluated once -- typically ~60 evaluations instead of (2*sa)^2.
// Windows with more than kMaxPos positions (search_area > 32) are evaluated position by position.
// Frames in which more than max_items blocks are left over (dense dependency chains) go to the
// row wavefront instead; the decision is handed over through n_unique[f] (free until K4 runs).
#include "mptc_kernels.h"
#include "mptc_device.cuh"

namespace mptc {

namespace {

constexpr int kThreads = 256, kWarps = kThreads / 32;
constexpr int kMaxRows = 2048;   // block rows per frame the directory can hold (8192-pixel-high frames)
constexpr int kMaxPos = 4096;    // window positions on the de-duplicating path (search_area <= 32)
constexpr uint32_t kEmpty = 0xFFFFFFFFu;   // empty-slot marker; the real word 0xFFFFFFFF lives in slot kHT
constexpr uint16_t kNone = 0xFFFFu;


// Dynamic shared memory of the de-duplicating path for a window of np positions: hash table of
// ht = 2 * pow2(np) slots, per-position ids, the pending list and the word / err_diff tables.
__host__ __device__ inline int sparse_ht(int np) {
  int p = 1;
  while (p < np) p <<= 1;
  return 2 * p;
}
__host__ __device__ inline size_t sparse_smem_bytes(int np) {
  const int ht = sparse_ht(np);
  return (size_t)(ht + 1) * 4 + (size_t)(2 * np) * 4 * 2 + (size_t)(ht + 2) * 2 + (size_t)np * 2 * 2;
}

__device__ __forceinline__ uint8_t ld_flag(const uint8_t *p) { return *reinterpret_cast<const volatile uint8_t *>(p); }

__device__ __forceinline__ uint16_t hash_insert(uint32_t *keys, int ht, int hshift, int *special, uint32_t word) {
  if (word == kEmpty) {
    *special = 1;
    return (uint16_t)ht;
  }
  uint32_t h = (word * 0x9E3779B1u) >> hshift;
  for (;;) {
    const uint32_t old = atomicCAS(&keys[h], kEmpty, word);
    if (old == kEmpty || old == word) break;
    h = (h + 1u) & (uint32_t)(ht - 1);
  }
  return (uint16_t)h;
}

// Builds the frame's raster-ordered directory of leftover blocks in shared memory (every CTA the
// same one): s_rows[q] = q-th block row with leftovers, s_row_off[q] = leftovers before that row.
// Returns the number of items, or -1 if the frame is left to the row wavefront.
__device__ __forceinline__ int sparse_directory(const SeqView &v, const uint8_t *flags, const uint8_t *row_todo, int max_items,
                                                uint16_t *s_rows, int *s_row_off, int *s_wsum, int *s_n_ptr, int &n_rows_out) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int &s_n = *s_n_ptr;
  // ---- rows with leftovers, in order ------------------------------------------------------------
  int n_rows = 0;
  for (int r0 = 0; r0 < v.bh; r0 += kThreads) {
    const int r = r0 + tid;
    const bool has = r < v.bh && row_todo[r] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, has);
    if (lane == 0) s_wsum[wid] = __popc(m);
    __syncthreads();
    int before = n_rows, total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const int c = s_wsum[w];
      if (w < wid) before += c;
      total += c;
    }
    if (has) {
      const int at = before + __popc(m & ((1u << lane) - 1u));
      if (at < kMaxRows) s_rows[at] = (uint16_t)r;
    }
    n_rows += total;
    __syncthreads();
  }
  bool overflow = n_rows > kMaxRows || v.bh > 65535;

  // ---- leftovers per row (warp = row), then the exclusive prefix over rows ---------------------------
  // The flags of leftover blocks change from 0 to 2 while other CTAs work, so "leftover" means
  // flag != 1: the directory is the same for every CTA whenever it is built.
  if (!overflow) {
    for (int q = wid; q < n_rows; q += kWarps) {
      const uint8_t *fr = flags + (size_t)s_rows[q] * v.bw;
      int c = 0;
      for (int x = lane; x < v.bw; x += 32) c += ld_flag(fr + x) != 1;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
      if (lane == 0) s_row_off[q + 1] = c;
    }
    if (tid == 0) s_row_off[0] = 0;
    __syncthreads();
    if (wid == 0) {   // inclusive scan over rows by one warp
      int carry = 0;
      for (int q0 = 0; q0 < n_rows; q0 += 32) {
        const int q = q0 + lane;
        int x = q < n_rows ? s_row_off[q + 1] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, x, d);
          if (lane >= d) x += y;
        }
        if (q < n_rows) s_row_off[q + 1] = carry + x;
        carry += __shfl_sync(0xffffffffu, x, 31);
      }
      if (lane == 0) s_n = carry;
    }
    __syncthreads();
    overflow = s_n > max_items;
  }
  n_rows_out = n_rows;
  return overflow ? -1 : s_n;
}

// item -> (block row, block column): every warp resolves it redundantly (no barrier needed).
__device__ __forceinline__ void sparse_locate(const SeqView &v, const uint8_t *flags, const uint16_t *s_rows, const int *s_row_off,
                                              int n_rows, int item, int lane, int &by_out, int &bx_out) {
  int lo = 0, hi = n_rows - 1;
  while (lo < hi) {   // largest q with s_row_off[q] <= item
    const int mid = (lo + hi + 1) >> 1;
    if (s_row_off[mid] <= item) lo = mid; else hi = mid - 1;
  }
  const int by = s_rows[lo];
  int nth = item - s_row_off[lo], bx = 0;
  const uint8_t *fr = flags + (size_t)by * v.bw;
  for (int x0 = 0; x0 < v.bw; x0 += 32 * 8) {       // 8 independent loads in flight per lane
    uint8_t fl[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int x = x0 + 32 * q + lane;
      fl[q] = x < v.bw ? ld_flag(fr + x) : (uint8_t)1;
    }
    bool done = false;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const unsigned m = __ballot_sync(0xffffffffu, fl[q] != 1);
      const int c = __popc(m);
      if (!done && nth < c) { bx = x0 + 32 * q + (int)__fns(m, 0, nth + 1); done = true; }
      if (!done) nth -= c;
    }
    if (done) break;
  }
  by_out = by;
  bx_out = bx;
}

}  // namespace

__global__ void __launch_bounds__(kThreads)
k_intra_sparse(SeqView v, int k_in_gop, int sa, int thr, int max_items, int *__restrict__ tickets) {
  __shared__ uint16_t s_rows[kMaxRows];
  __shared__ int s_row_off[kMaxRows + 1];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_wsum[kWarps];
  __shared__ int s_n, s_item, s_count, s_special, s_pending, s_late;
  __shared__ TargetCtx s_t;
  __shared__ WinnerState s_red[kWarps];

  const int f = v.first + blockIdx.y * v.gop + k_in_gop;
  if (f >= v.first + v.count) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint8_t *frame = v.rgb + v.frame_bytes * f;
  uint64_t *cur = v.final_blocks + (size_t)f * v.nb;
  const uint32_t *cur_words = reinterpret_cast<const uint32_t *>(cur);
  const uint64_t *init = v.init_blocks + (size_t)f * v.nb;
  uint8_t *flags = v.flags + (size_t)f * v.nb;
  uint8_t *motion = v.motion + (size_t)f * v.nb * 2;
  const uint8_t *row_todo = v.row_todo + (size_t)f * v.bh;
  int *ticket = tickets + blockIdx.y;

  int n_rows = 0;
  const int dir_items = sparse_directory(v, flags, row_todo, max_items, s_rows, s_row_off, s_wsum, &s_n, n_rows);
  const bool overflow = dir_items < 0;
  const int n_items = overflow ? 0 : s_n;
  if (blockIdx.x == 0 && tid == 0) v.n_unique[f] = overflow ? kSparseNotHandled : (uint32_t)n_items;
  if (overflow || n_items == 0) return;   // the row wavefront takes the frame / nothing to do

  // ---- work items in raster order ---------------------------------------------------------------------
  const int W = 2 * sa, NP = W * W;
  const bool dedup = NP <= kMaxPos;
  const int kHT = sparse_ht(NP), kHashShift = 33 - __ffs(kHT);
  uint32_t *s_keys = reinterpret_cast<uint32_t *>(smem_raw);   // [kHT + 1]
  uint32_t *s_ulist = s_keys + kHT + 1;                         // [2 NP] distinct words: [0, U) from the window, then late arrivals
  int *s_err = reinterpret_cast<int *>(s_ulist + 2 * NP);       // [2 NP]
  uint16_t *s_slot_uid = reinterpret_cast<uint16_t *>(s_err + 2 * NP);   // [kHT + 2]
  uint16_t *s_pos_uid = s_slot_uid + kHT + 2;                   // [NP] per window position: hash slot, then dense word id
  uint16_t *s_pend = s_pos_uid + NP;                            // [NP] positions whose block was undecided at load time
  uint32_t *cur_words_rw = reinterpret_cast<uint32_t *>(cur);
  for (;;) {
    __syncthreads();   // everything of the previous item is consumed
    if (tid == 0) { s_item = atomicAdd(ticket, 1); s_count = 0; s_special = 0; s_pending = 0; s_late = 0; }
    if (dedup)
      for (int s = tid; s < kHT; s += kThreads) s_keys[s] = kEmpty;
    __syncthreads();
    const int item = s_item;
    if (item >= n_items) return;
    int by, bx;
    sparse_locate(v, flags, s_rows, s_row_off, n_rows, item, lane, by, bx);
    const int b = by * v.bw + bx;
    if (tid == 0) build_target(s_t, frame, v.w, bx, by, init[b]);
    WinnerState s;
    winner_init(s);
    if (dedup) {
      // ---- load the window and insert its words.  A position whose block is an earlier leftover
      // that is still undecided goes to the pending list; its INITIAL word is inserted
      // speculatively, because a block's final word is either a word of its own window (mostly
      // shared with this one) or its initial word -- so when the decision arrives, its err_diff is
      // usually in the table already and the dependency chain only pays for the lookup. ---------------
      for (int p = tid; p < NP; p += kThreads) {
        const int row = p / W, col = p - row * W;          // scan order: j downwards, i downwards
        const int j = by - row, i = bx + sa - 1 - col;
        uint16_t slot = kNone;
        if (i >= 0 && j >= 0 && i < v.bw && !(row == 0 && i >= bx)) {
          const size_t idx = (size_t)j * v.bw + i;
          const uint8_t fl = ld_flag(flags + idx);
          const bool undecided = fl == 0;
          if (fl == 2) __threadfence();                     // decided by another CTA of this launch
          const uint32_t word = undecided ? (uint32_t)(init[idx] >> 32) : __ldcg(cur_words + 2 * idx + 1);
          slot = hash_insert(s_keys, kHT, kHashShift, &s_special, word);
          if (undecided) {
            s_pend[atomicAdd(&s_pending, 1)] = (uint16_t)p;
            slot = kNone;                                   // resolved below
          }
        }
        s_pos_uid[p] = slot;
      }
      __syncthreads();
      // ---- dense ids, then every distinct word once -----------------------------------------------------
      for (int q = tid; q <= kHT; q += kThreads) {
        const bool occ = (q < kHT) ? (s_keys[q] != kEmpty) : (s_special != 0);
        if (occ) {
          const int uid = atomicAdd(&s_count, 1);
          s_slot_uid[q] = (uint16_t)uid;
          s_ulist[uid] = (q < kHT) ? s_keys[q] : kEmpty;
        }
      }
      __syncthreads();
      const int U = s_count, n_pend = s_pending;
      for (int p = tid; p < NP; p += kThreads) {
        const uint16_t slot = s_pos_uid[p];
        if (slot != kNone) s_pos_uid[p] = s_slot_uid[slot];
      }
      __syncthreads();   // the pending positions are rewritten below
      for (int u = tid; u < U; u += kThreads) s_err[u] = eval_candidate(s_t, s_ulist[u]);
      // ---- the pending positions: wait for their owners, look their final words up ------------------------
      for (int q = tid; q < n_pend; q += kThreads) {
        const int p = s_pend[q];
        const int row = p / W, col = p - row * W;
        const size_t idx = (size_t)(by - row) * v.bw + (bx + sa - 1 - col);
        while (ld_flag(flags + idx) == 0) __nanosleep(32);
        __threadfence();
        const uint32_t word = __ldcg(cur_words + 2 * idx + 1);
        int uid = -1;
        if (word == kEmpty) {
          if (s_special) uid = s_slot_uid[kHT];
        } else {
          uint32_t h = (word * 0x9E3779B1u) >> kHashShift, kv;
          while ((kv = s_keys[h]) != kEmpty && kv != word) h = (h + 1u) & (kHT - 1);
          if (kv == word) uid = s_slot_uid[h];
        }
        if (uid < 0) {                                      // rare: a word from outside this window
          uid = U + atomicAdd(&s_late, 1);
          s_ulist[uid] = word;
        }
        s_pos_uid[p] = (uint16_t)uid;
      }
      __syncthreads();
      const int n_late = s_late;
      if (n_late > 0) {
        for (int u = tid; u < n_late; u += kThreads) s_err[U + u] = eval_candidate(s_t, s_ulist[U + u]);
        __syncthreads();
      }
      for (int p = tid; p < NP; p += kThreads) {
        const uint16_t uid = s_pos_uid[p];
        if (uid == kNone) continue;
        const int row = p / W;
        winner_update(s, s_err[uid], row, p - row * W, W);
      }
    } else {
      __syncthreads();   // s_t
      for (int p = tid; p < NP; p += kThreads) {
        const int row = p / W, col = p - row * W;
        const int j = by - row, i = bx + sa - 1 - col;
        if (i < 0 || j < 0 || i >= v.bw || (row == 0 && i >= bx)) continue;
        const size_t idx = (size_t)j * v.bw + i;
        if (ld_flag(flags + idx) == 0) {
          while (ld_flag(flags + idx) == 0) __nanosleep(64);
          __threadfence();
        }
        winner_update(s, eval_candidate(s_t, __ldcg(cur_words + 2 * idx + 1)), row, col, W);
      }
    }
    winner_block_reduce<kWarps>(s, s_red);
    if (tid == 0) {
      int row, col;
      const int min_err = winner_resolve(s, W, row, col);
      if (min_err <= thr) {
        const size_t src = (size_t)(by - row) * v.bw + (bx + sa - 1 - col);
        const uint32_t word = __ldcg(cur_words + 2 * src + 1);
        // the index word first: that is all a dependant waits for; endpoints follow off the chain
        cur_words_rw[2 * (size_t)b + 1] = word;
        __threadfence();
        *reinterpret_cast<volatile uint8_t *>(flags + b) = 2;
        cur_words_rw[2 * (size_t)b] = (uint32_t)winning_block(s_t, word);
        motion[2 * b + 0] = (uint8_t)(2 * sa - 1 - col);   // x = (i - bx) + sa
        motion[2 * b + 1] = (uint8_t)(2 * sa - 1 - row);   // y = (j - by) + 2sa - 1
      } else {
        *reinterpret_cast<volatile uint8_t *>(flags + b) = 2;   // keeps its initial block
        motion[2 * b + 0] = 255;
        motion[2 * b + 1] = 255;
      }
    }
  }
}

// Tiny windows (search_area <= 2 as launched; the kernel handles up to 64 positions): one WARP per item.  Small windows leave
// 10-45 % of an inter frame's blocks over and each item is cheap, so what counts is the number of
// items in flight: eight per CTA instead of one.  Same directory, same ticket order, same
// per-block flags as k_intra_sparse; no de-duplication (every lane evaluates its one or two
// positions directly).
__global__ void __launch_bounds__(kThreads)
k_intra_sparse_warp(SeqView v, int k_in_gop, int sa, int thr, int max_items, int *__restrict__ tickets) {
  __shared__ uint16_t s_rows[kMaxRows];
  __shared__ int s_row_off[kMaxRows + 1];
  __shared__ int s_wsum[kWarps];
  __shared__ int s_n;
  __shared__ TargetCtx s_t[kWarps];

  const int f = v.first + blockIdx.y * v.gop + k_in_gop;
  if (f >= v.first + v.count) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint8_t *frame = v.rgb + v.frame_bytes * f;
  uint64_t *cur = v.final_blocks + (size_t)f * v.nb;
  const uint32_t *cur_words = reinterpret_cast<const uint32_t *>(cur);
  uint32_t *cur_words_rw = reinterpret_cast<uint32_t *>(cur);
  const uint64_t *init = v.init_blocks + (size_t)f * v.nb;
  uint8_t *flags = v.flags + (size_t)f * v.nb;
  uint8_t *motion = v.motion + (size_t)f * v.nb * 2;
  const uint8_t *row_todo = v.row_todo + (size_t)f * v.bh;
  int *ticket = tickets + blockIdx.y;

  int n_rows = 0;
  const int dir_items = sparse_directory(v, flags, row_todo, max_items, s_rows, s_row_off, s_wsum, &s_n, n_rows);
  const bool overflow = dir_items < 0;
  const int n_items = overflow ? 0 : s_n;
  if (blockIdx.x == 0 && tid == 0) v.n_unique[f] = overflow ? kSparseNotHandled : (uint32_t)n_items;
  if (overflow || n_items == 0) return;

  const int W = 2 * sa, NP = W * W;   // NP <= 64
  TargetCtx &t = s_t[wid];
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(ticket, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) return;
    int by, bx;
    sparse_locate(v, flags, s_rows, s_row_off, n_rows, item, lane, by, bx);
    const int b = by * v.bw + bx;
    __syncwarp();                      // the previous item's reads of t are done
    if (lane == 0) build_target(t, frame, v.w, bx, by, init[b]);
    __syncwarp();
    WinnerState s;
    winner_init(s);
    for (int p = lane; p < NP; p += 32) {
      const int row = p / W, col = p - row * W;          // scan order: j downwards, i downwards
      const int j = by - row, i = bx + sa - 1 - col;
      if (i < 0 || j < 0 || i >= v.bw || (row == 0 && i >= bx)) continue;
      const size_t idx = (size_t)j * v.bw + i;
      uint8_t fl = ld_flag(flags + idx);
      while (fl == 0) {                                  // an earlier leftover, still undecided
        __nanosleep(32);
        fl = ld_flag(flags + idx);
      }
      if (fl == 2) __threadfence();                      // decided by another warp of this launch
      winner_update(s, eval_candidate(t, __ldcg(cur_words + 2 * idx + 1)), row, col, W);
    }
    winner_warp_reduce(s);
    if (lane == 0) {
      int row, col;
      const int min_err = winner_resolve(s, W, row, col);
      if (min_err <= thr) {
        const size_t src = (size_t)(by - row) * v.bw + (bx + sa - 1 - col);
        const uint32_t word = __ldcg(cur_words + 2 * src + 1);
        cur_words_rw[2 * (size_t)b + 1] = word;          // the index word first: all a dependant waits for
        __threadfence();
        *reinterpret_cast<volatile uint8_t *>(flags + b) = 2;
        cur_words_rw[2 * (size_t)b] = (uint32_t)winning_block(t, word);
        motion[2 * b + 0] = (uint8_t)(2 * sa - 1 - col);   // x = (i - bx) + sa
        motion[2 * b + 1] = (uint8_t)(2 * sa - 1 - row);   // y = (j - by) + 2sa - 1
      } else {
        *reinterpret_cast<volatile uint8_t *>(flags + b) = 2;   // keeps its initial block
        motion[2 * b + 0] = 255;
        motion[2 * b + 1] = 255;
      }
    }
  }
}

bool launch_intra_sparse(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, int *tickets, int ctas_per_frame,
                         int max_items, cudaStream_t s) {
  const int np = 4 * sa * sa;
  const size_t bytes = np <= kMaxPos ? sparse_smem_bytes(np) : 0;
  static size_t configured[kMaxDevices] = {0};   // per device; 0 = nothing beyond what fits without opt-in
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  {
    std::lock_guard<std::mutex> lock(launch_cfg_mutex());
    size_t &conf = configured[cur_dev & (kMaxDevices - 1)];
    if (bytes > 48 * 1024 - 13 * 1024 && bytes > conf) {   // 35 KB fit next to the static arrays without opt-in
      // on failure nothing is launched and the caller reports the error: the frames' n_unique would
      // otherwise keep a stale "handled" value and their leftover blocks would never be searched
      if (cudaFuncSetAttribute(k_intra_sparse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return false;
      conf = bytes;
    }
  }
  dim3 grid(ctas_per_frame, n_gops);
  if (np <= 16) {   // search_area <= 2 (measured: sa 2 +16 % at thr 50, +11 % at thr 0; sa 4 -31 % at thr 0, where
                    // de-duplicating the 64 positions per item pays)
    k_intra_sparse_warp<<<grid, kThreads, 0, s>>>(v, k_in_gop, sa, thr, max_items, tickets);
    return true;
  }
  k_intra_sparse<<<grid, kThreads, bytes, s>>>(v, k_in_gop, sa, thr, max_items, tickets);
  return true;
}

}  // namespace mptc
