// mptc_capi.cu -- the extern "C" boundary declared in include/mptc_gpu.h.
//
// Owns the context (streams, events, device-resident sequence buffers) and schedules the
// kernels of mptc_kernels.cu.  Scheduling mirrors the reference's frame loop
// (codec/codec.cpp:1383-1509) but runs frame k of every GOP in the same launches:
// GOPs are independent, frames inside a GOP are not.
#include "../../include/mptc_gpu.h"
#include "mptc_kernels.h"
#include "mptc_host.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace mptc;

namespace {

struct StageEvent {
  int stage;
  cudaEvent_t a, b;
};

}  // namespace

struct mptc_gpu_ctx {
  int device = 0;
  cudaStream_t s_compute = nullptr, s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_uploaded = nullptr, ev_encoded = nullptr;
  // reserved sequence
  int w = 0, h = 0, bw = 0, bh = 0, nb = 0, pbw = 0, pbh = 0, cap_frames = 0;
  size_t frame_bytes = 0, plane_bytes = 0;  // plane_bytes = 6*pbw*pbh
  uint8_t *d_rgb = nullptr;
  uint64_t *d_init = nullptr, *d_final = nullptr;
  uint8_t *d_motion = nullptr, *d_flags = nullptr, *d_planes = nullptr, *d_row_todo = nullptr;
  uint32_t *d_unique = nullptr, *d_nunique = nullptr, *d_chunks = nullptr;
  int *d_progress = nullptr;
  unsigned long long *d_wordflag = nullptr;
  int8_t *d_pattern = nullptr;   // K2p: pixel offsets of DXTImage::SetPattern for pattern_sa
  int pattern_sa = 0, pattern_n = 0;
  uint32_t epoch = 0;            // one per encode call: validity tag of the wavefront's hand-over entries
  unsigned long long *d_cand = nullptr;
  int max_wave_ctas = 0;
  bool encoded = false;
  // GOP lanes: independent GOP ranges of one encode call run on their own streams, so the
  // latency-bound intra wavefront of one lane overlaps the throughput-bound inter search of
  // the others.  Inside a lane the work is enqueued frame by frame (frame k of every GOP of the
  // lane): search kernels on `s`, compaction + endpoint planes on the side stream `t`, and -- end
  // to end -- the H2D copy of frame k+1 and the D2H copy of frame k-1 overlap the kernels of k.
  struct Lane {
    cudaStream_t s = nullptr, t = nullptr;
    std::vector<cudaEvent_t> ev_up, ev_k, ev_side, ev_down;   // per frame index k inside the GOP
    int *d_tickets = nullptr;
    int tickets_cap = 0;
    int f0 = 0, n = 0, n_gops = 0;   // frame range of the current call
    std::vector<StageEvent> stage_events;
    size_t stage_events_used = 0;
  };
  std::vector<Lane> lanes;
  int lanes_wanted = 0;          // 0 = automatic
  int lanes_used = 0;
  int enc_first = 0, enc_count = 0, enc_gop = 1;   // the last encode call
  bool enc_host_out = false;
  int sparse_ctas = 0;           // K3s CTAs per frame; 0 = automatic
  int sparse_max_pct = 50;       // K3s takes inter frames with up to this share of leftover blocks
  int wave_rows_intra = 0, wave_rows_inter = 0;   // CTAs per frame of the intra wavefront; 0 = default
  uint64_t launches = 0;
  void *pinned[4] = {nullptr, nullptr, nullptr, nullptr};   // result staging of mptc_encode_stream
  size_t pinned_bytes[4] = {0, 0, 0, 0};
  // decoder side (mptc_decode.cu): allocated by the first decode call for the reserved sequence
  uint32_t *d_dec_words = nullptr, *d_dec_chunks = nullptr, *d_dec_uoff = nullptr;
  int *d_dec_link = nullptr, *d_dec_status = nullptr;     // status: [0] errors, then 40 ints per frame
  uint64_t *d_dec_blocks = nullptr;
  uint8_t *d_dec_rgb = nullptr;
  std::vector<uint32_t> h_uoff;   // word offset of every frame's unique words inside d_unique
  cudaEvent_t ev_dec[4] = {nullptr, nullptr, nullptr, nullptr};   // begin, words, planes, rgb/end
  bool decoded = false, dec_pending = false, dec_rgb = false;
  char err[512] = {0};
};

namespace {

int fail(mptc_gpu_ctx *c, int code, const char *fmt, ...) {
  if (c) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c->err, sizeof c->err, fmt, ap);
    va_end(ap);
  }
  return code;
}

#define CU(ctx, call)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return fail(ctx, e__ == cudaErrorMemoryAllocation ? MPTC_E_NOMEM : MPTC_E_CUDA,          \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// stb__PrepareOptTable (Include/stb_dxt.h:111-137): best (max,min) 5/6-bit pair whose 1/3
// interpolant reproduces each 8-bit value, with the 3 % range penalty.
void build_match_table(uint8_t *table, int bits) {
  const int size = 1 << bits;
  for (int target = 0; target < 256; ++target) {
    int best = 256;
    for (int lo = 0; lo < size; ++lo)
      for (int hi = 0; hi < size; ++hi) {
        int e_lo = bits == 5 ? ((lo << 3) | (lo >> 2)) : ((lo << 2) | (lo >> 4));
        int e_hi = bits == 5 ? ((hi << 3) | (hi >> 2)) : ((hi << 2) | (hi >> 4));
        int err = abs((2 * e_hi + e_lo) / 3 - target) + abs(e_hi - e_lo) * 3 / 100;
        if (err < best) {
          table[2 * target + 0] = (uint8_t)hi;
          table[2 * target + 1] = (uint8_t)lo;
          best = err;
        }
      }
  }
}

void free_seq(mptc_gpu_ctx *c) {
  cudaFree(c->d_rgb); cudaFree(c->d_init); cudaFree(c->d_final); cudaFree(c->d_motion);
  cudaFree(c->d_flags); cudaFree(c->d_planes); cudaFree(c->d_unique); cudaFree(c->d_nunique);
  cudaFree(c->d_progress); cudaFree(c->d_row_todo); cudaFree(c->d_chunks); cudaFree(c->d_wordflag);
  c->d_wordflag = nullptr;
  c->d_chunks = nullptr;
  c->d_rgb = nullptr; c->d_init = c->d_final = nullptr; c->d_motion = c->d_flags = c->d_planes = nullptr;
  c->d_unique = c->d_nunique = nullptr; c->d_progress = nullptr; c->d_row_todo = nullptr;
  cudaFree(c->d_dec_words); cudaFree(c->d_dec_chunks); cudaFree(c->d_dec_uoff); cudaFree(c->d_dec_link);
  cudaFree(c->d_dec_status); cudaFree(c->d_dec_blocks); cudaFree(c->d_dec_rgb);
  c->d_dec_words = c->d_dec_chunks = c->d_dec_uoff = nullptr; c->d_dec_link = c->d_dec_status = nullptr;
  c->d_dec_blocks = nullptr; c->d_dec_rgb = nullptr;
  c->cap_frames = 0; c->w = c->h = 0;
  c->encoded = c->decoded = c->dec_pending = false;
}

constexpr int kDecStatusStride = 40;   // pass flags per decode call, indexed by its first frame

// Decoder-side buffers for the reserved sequence (lazy: encode-only users never pay for them).
int ensure_decode(mptc_gpu_ctx *c, bool want_rgb) {
  const size_t F = (size_t)c->cap_frames, nb = (size_t)c->nb;
  if (!c->d_dec_link) {
    CU(c, cudaMalloc(&c->d_dec_words, F * nb * 4));
    CU(c, cudaMalloc(&c->d_dec_link, F * nb * sizeof(int)));
    CU(c, cudaMalloc(&c->d_dec_chunks, F * dec_chunks(c->nb) * 4));
    CU(c, cudaMalloc(&c->d_dec_uoff, F * 4));
    CU(c, cudaMalloc(&c->d_dec_status, (1 + F * kDecStatusStride) * sizeof(int)));
    CU(c, cudaMalloc(&c->d_dec_blocks, F * nb * 8));
    CU(c, cudaMemset(c->d_dec_status, 0, sizeof(int)));
  }
  if (want_rgb && !c->d_dec_rgb) CU(c, cudaMalloc(&c->d_dec_rgb, F * c->frame_bytes));
  for (auto &e : c->ev_dec)
    if (!e) CU(c, cudaEventCreate(&e));
  return MPTC_OK;
}

DecView dec_view_of(const mptc_gpu_ctx *c, int first, int count, int gop, int sa) {
  DecView v;
  v.motion = c->d_motion; v.unique = c->d_unique; v.unique_off = c->d_dec_uoff; v.n_unique = c->d_nunique;
  v.planes = c->d_planes; v.words = c->d_dec_words; v.link = c->d_dec_link; v.chunk_counts = c->d_dec_chunks;
  v.errors = c->d_dec_status; v.status = c->d_dec_status + 1 + (size_t)first * kDecStatusStride;
  v.blocks = c->d_dec_blocks; v.rgb = c->d_dec_rgb;
  v.w = c->w; v.h = c->h; v.bw = c->bw; v.bh = c->bh; v.nb = c->nb; v.pbw = c->pbw; v.pbh = c->pbh;
  v.first = first; v.count = count; v.gop = gop; v.sa = sa;
  return v;
}

SeqView view_of(const mptc_gpu_ctx *c, int first, int count, int gop) {
  SeqView v;
  v.rgb = c->d_rgb; v.init_blocks = c->d_init; v.final_blocks = c->d_final; v.motion = c->d_motion;
  v.flags = c->d_flags; v.row_todo = c->d_row_todo; v.unique = c->d_unique; v.n_unique = c->d_nunique; v.chunk_counts = c->d_chunks; v.planes = c->d_planes;
  v.progress = c->d_progress; v.wordflag = c->d_wordflag; v.epoch = c->epoch; v.work = c->d_cand; v.frame_bytes = c->frame_bytes;
  v.w = c->w; v.h = c->h; v.bw = c->bw; v.bh = c->bh; v.nb = c->nb;
  v.first = first; v.count = count; v.gop = gop;
  return v;
}

typedef mptc_gpu_ctx::Lane Lane;

StageEvent &stage_begin(Lane &L, int stage, cudaStream_t st) {
  if (L.stage_events_used == L.stage_events.size()) {
    StageEvent e;
    e.stage = stage;
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    L.stage_events.push_back(e);
  }
  StageEvent &e = L.stage_events[L.stage_events_used++];
  e.stage = stage;
  cudaEventRecord(e.a, st);
  return e;
}

void stage_end(mptc_gpu_ctx *c, StageEvent &e, cudaStream_t st, int n_launches = 1) {
  cudaEventRecord(e.b, st);
  c->launches += n_launches;
  static const bool debug_sync = getenv("MPTC_DEBUG_SYNC") != nullptr;   // debugging: which stage faults?
  if (debug_sync) {
    const cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) fprintf(stderr, "mptc: stage %d failed: %s\n", e.stage, cudaGetErrorString(err));
  }
}

int check_params(mptc_gpu_ctx *c, int sa, int gop) {
  if (sa < 1 || sa > 63) return fail(c, MPTC_E_ARG, "search_area %d outside 1..63 (uint8 motion bytes)", sa);
  if (gop < 1) return fail(c, MPTC_E_ARG, "gop %d < 1", gop);
  return MPTC_OK;
}

int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

int ensure_lanes(mptc_gpu_ctx *c, int n, int gop) {
  while ((int)c->lanes.size() < n) {
    Lane L;
    CU(c, cudaStreamCreateWithFlags(&L.s, cudaStreamNonBlocking));
    CU(c, cudaStreamCreateWithFlags(&L.t, cudaStreamNonBlocking));
    c->lanes.push_back(L);
  }
  for (int i = 0; i < n; ++i) {
    Lane &L = c->lanes[i];
    while ((int)L.ev_up.size() < gop) {
      cudaEvent_t e[4];
      for (int q = 0; q < 4; ++q) CU(c, cudaEventCreateWithFlags(&e[q], cudaEventDisableTiming));
      L.ev_up.push_back(e[0]); L.ev_k.push_back(e[1]); L.ev_side.push_back(e[2]); L.ev_down.push_back(e[3]);
    }
  }
  return MPTC_OK;
}


// Host buffers of an end-to-end call; all optional.
struct HostIO {
  const uint8_t *frames = nullptr;
  uint64_t *blocks = nullptr;
  uint8_t *motion = nullptr;
  uint32_t *unique = nullptr, *n_unique = nullptr;
  uint8_t *planes = nullptr;
  // device-visible aliases of unique / n_unique when the caller's buffers are page-locked: K4 then
  // writes the n_unique words of a frame straight to the host instead of a D2H copy of nb words
  uint32_t *unique_mapped = nullptr, *n_unique_mapped = nullptr;
  bool any_out() const { return blocks || motion || unique || n_unique || planes; }
};

// The device-visible address of a page-locked host buffer (cudaHostAlloc / cudaHostRegister), or null
// for pageable memory.
template <typename T>
T *mapped_alias(T *host) {
  if (!host) return nullptr;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return a.type == cudaMemoryTypeHost ? static_cast<T *>(a.devicePointer) : nullptr;
}

// Copies `rows` chunks of `width` bytes that lie `pitch` bytes apart in both src and dst (frame k of
// consecutive GOPs).  One 2D copy when the pitch allows it, otherwise one copy per chunk.
cudaError_t copy_strided(void *dst, const void *src, size_t width, size_t pitch, int rows, cudaMemcpyKind kind,
                         cudaStream_t st) {
  if (rows == 1 || width == pitch) return cudaMemcpyAsync(dst, src, width * (size_t)rows, kind, st);
  if (pitch < ((size_t)1 << 31)) return cudaMemcpy2DAsync(dst, pitch, src, pitch, width, (size_t)rows, kind, st);
  for (int r = 0; r < rows; ++r) {
    cudaError_t e = cudaMemcpyAsync(static_cast<char *>(dst) + pitch * r, static_cast<const char *>(src) + pitch * r,
                                    width, kind, st);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// Start of a lane's range [f0, f0 + n): reset the wavefront state of its frames.
int lane_prologue(mptc_gpu_ctx *c, Lane &L, int gop) {
  const int n_tickets = gop * (1 + L.n_gops);
  // tickets: [k] for the row wavefront of frame k of every GOP, then [gop + k*n_gops + g] for K3s
  if (n_tickets > L.tickets_cap) {
    if (L.d_tickets) CU(c, cudaFree(L.d_tickets));   // implicit sync; only on the first calls
    CU(c, cudaMalloc(&L.d_tickets, sizeof(int) * n_tickets));
    L.tickets_cap = n_tickets;
  }
  cudaStream_t s = L.s;
  CU(c, cudaMemsetAsync(L.d_tickets, 0, sizeof(int) * n_tickets, s));
  CU(c, cudaMemsetAsync(c->d_progress + (size_t)L.f0 * c->bh, 0, sizeof(int) * (size_t)L.n * c->bh, s));
  CU(c, cudaMemsetAsync(c->d_flags + (size_t)L.f0 * c->nb, 0, (size_t)L.n * c->nb, s));
  CU(c, cudaMemsetAsync(c->d_row_todo + (size_t)L.f0 * c->bh, 0, (size_t)L.n * c->bh, s));
  return MPTC_OK;
}

// Enqueues frame k of every GOP of lane L: [H2D] -> K1 -> K2 -> K3s/K3 on L.s, then K4 + K5 on the
// side stream L.t, then [D2H] on the download stream.
int enqueue_step(mptc_gpu_ctx *c, Lane &L, int k, int gop, int sa, int thr, bool fit, bool planes, int wave_rows,
                 const HostIO *io, int call_first) {
  const int nf = (L.n - k + gop - 1) / gop;       // GOPs of the lane that have a frame k
  if (nf <= 0) return MPTC_OK;
  const int fk = L.f0 + k;
  const size_t nb = (size_t)c->nb;
  SeqView v = view_of(c, L.f0, L.n, gop);
  // an explicitly requested single lane serialises everything (profiling: no kernel overlaps another)
  cudaStream_t s = L.s, t = c->lanes_wanted == 1 ? L.s : L.t;
  if (io && io->frames) {
    CU(c, copy_strided(c->d_rgb + c->frame_bytes * fk, io->frames + c->frame_bytes * (size_t)(fk - call_first),
                       c->frame_bytes, c->frame_bytes * gop, nf, cudaMemcpyHostToDevice, c->s_h2d));
    CU(c, cudaEventRecord(L.ev_up[k], c->s_h2d));
    CU(c, cudaStreamWaitEvent(s, L.ev_up[k], 0));
  }
  if (fit) {
    StageEvent &e = stage_begin(L, 1, s);
    launch_dxt1_fit(v, fk, gop, nf, s);
    stage_end(c, e, s);
  }
  if (k > 0) {
    // (Measured and dropped, profiles/r2_k2_wide.txt: K2 on a low-priority stream of its own so that the
    // lanes' short kernels overtake the other lanes' queued K2 CTAs -- the GPU no longer drains between
    // rounds of frames, but the K2 CTAs then share their SMs with more of the short kernels: same step time.)
    StageEvent &e = stage_begin(L, 2, s);
    const int n_kernels = launch_inter_search(v, k, L.n_gops, sa, thr, s);
    stage_end(c, e, s, n_kernels);
  }
  {
    StageEvent &e = stage_begin(L, 3, s);
    if (k > 0) {
      int ctas = c->sparse_ctas;
      if (ctas <= 0) {
        // small windows leave more blocks over and make the items cheap: twice the CTAs per frame pay
        // there (sa 2: +18 %), while at the default window they cost 1 % (profiles/r1_v7_sparse_ctas.txt)
        const int cap = sa <= 4 ? 296 : 148;
        ctas = 4 * cap / L.n_gops;
        ctas = ctas < 16 ? 16 : (ctas > cap ? cap : ctas);
      }
      const int max_items = (int)((long long)c->nb * c->sparse_max_pct / 100);
      if (!launch_intra_sparse(v, k, L.n_gops, sa, thr, L.d_tickets + gop + k * L.n_gops, ctas, max_items, s))
        return fail(c, MPTC_E_CUDA, "the sparse intra search kernel could not be configured for search_area %d", sa);
    }
    launch_intra_wavefront(v, k, L.n_gops, sa, thr, L.d_tickets + k, c->max_wave_ctas, wave_rows * L.n_gops, s);
    stage_end(c, e, s, k > 0 ? 2 : 1);
  }
  CU(c, cudaEventRecord(L.ev_k[k], s));
  CU(c, cudaStreamWaitEvent(t, L.ev_k[k], 0));
  {
    StageEvent &e = stage_begin(L, 4, t);
    launch_compact_unique(v, sa, c->d_cand, fk, gop, nf, t, io ? io->unique_mapped : nullptr,
                          io ? io->n_unique_mapped : nullptr, call_first);
    stage_end(c, e, t, 2);
  }
  if (planes) {
    StageEvent &e = stage_begin(L, 5, t);
    launch_endpoint_planes(v, c->pbw, c->pbh, fk, gop, nf, t);
    stage_end(c, e, t);
  }
  CU(c, cudaEventRecord(L.ev_side[k], t));
  if (io && io->any_out()) {
    cudaStream_t d = c->s_d2h;
    const size_t o = (size_t)(fk - call_first), f = (size_t)fk, g = (size_t)gop;
    const cudaMemcpyKind D2H = cudaMemcpyDeviceToHost;
    CU(c, cudaStreamWaitEvent(d, L.ev_side[k], 0));
    if (io->blocks) CU(c, copy_strided(io->blocks + o * nb, c->d_final + f * nb, nb * 8, g * nb * 8, nf, D2H, d));
    if (io->motion) CU(c, copy_strided(io->motion + o * nb * 2, c->d_motion + f * nb * 2, nb * 2, g * nb * 2, nf, D2H, d));
    if (io->unique && !io->unique_mapped) CU(c, copy_strided(io->unique + o * nb, c->d_unique + f * nb, nb * 4, g * nb * 4, nf, D2H, d));
    if (io->n_unique && !io->n_unique_mapped) CU(c, copy_strided(io->n_unique + o, c->d_nunique + f, 4, g * 4, nf, D2H, d));
    if (io->planes) CU(c, copy_strided(io->planes + o * c->plane_bytes, c->d_planes + f * c->plane_bytes, c->plane_bytes,
                                       g * c->plane_bytes, nf, D2H, d));
    CU(c, cudaEventRecord(L.ev_down[k], d));
  }
  return MPTC_OK;
}

// One encode call over frames [first, first+count): the GOPs are split into contiguous ranges,
// one per lane, and enqueued frame-index-major (frame k of every lane, then k+1, ...), which is
// also the order of the H2D and D2H copies.  ev_begin .. ev_end on s_compute bracket everything
// (including the copies when io is given).  k_begin lets single-frame calls start at an inter
// frame whose predecessor's final blocks were supplied by the caller.
int run_encode(mptc_gpu_ctx *c, int first, int count, int gop, int sa, int thr, bool fit, int k_begin,
               bool planes, const HostIO *io = nullptr) {
  const int n_gops = (count + gop - 1) / gop;
  int nl = c->lanes_wanted > 0 ? c->lanes_wanted : 4;
  if (nl > n_gops) nl = n_gops;
  if (k_begin != 0) nl = 1;
  if (nl > 16) nl = 16;
  if (int r = ensure_lanes(c, nl, gop)) return r;
  // CTAs per frame of the intra wavefront: with several lanes a full grid of (mostly waiting)
  // wavefront CTAs would keep the other lanes' kernels off the SMs
  // (37 = a quarter of the SMs per lane with the default four lanes: 18.75 ms against 18.98 with 32 and
  // 19.5 with 48, profiles/r2_sched_sweep.txt)
  const int rows_intra = c->wave_rows_intra > 0 ? c->wave_rows_intra : (nl > 1 ? 37 : 0);
  const int rows_inter = c->wave_rows_inter;
  cudaStream_t s0 = c->s_compute;
  if (++c->epoch == 0) {   // wrapped: entries of 2^32 calls ago would look current
    CU(c, cudaMemsetAsync(c->d_wordflag, 0, (size_t)c->cap_frames * c->nb * sizeof(unsigned long long), s0));
    c->epoch = 1;
  }
  CU(c, cudaMemsetAsync(c->d_cand, 0, kWorkCounters * sizeof(unsigned long long), s0));
  CU(c, cudaEventRecord(c->ev_begin, s0));
  if (io && io->frames) CU(c, cudaStreamWaitEvent(c->s_h2d, c->ev_begin, 0));
  for (int i = 0; i < nl; ++i) {
    Lane &L = c->lanes[i];
    L.stage_events_used = 0;
    const int g0 = (int)((long long)n_gops * i / nl), g1 = (int)((long long)n_gops * (i + 1) / nl);
    L.f0 = first + g0 * gop;
    L.n = (g1 - g0) * gop;
    if (L.f0 + L.n > first + count) L.n = first + count - L.f0;
    L.n_gops = g1 - g0;
    CU(c, cudaStreamWaitEvent(L.s, c->ev_begin, 0));
    if (int r = lane_prologue(c, L, gop)) return r;
  }
  const int k_end = count < gop ? count : gop;
  for (int k = k_begin; k < k_end; ++k)
    for (int i = 0; i < nl; ++i)
      if (int r = enqueue_step(c, c->lanes[i], k, gop, sa, thr, fit, planes, k == 0 ? rows_intra : rows_inter, io, first))
        return r;
  const bool host_out = io && io->any_out();
  for (int i = 0; i < nl; ++i) {
    Lane &L = c->lanes[i];
    const int k_last = (L.n < gop ? L.n : gop) - 1;
    if (k_last < k_begin) continue;
    CU(c, cudaStreamWaitEvent(s0, host_out ? L.ev_down[k_last] : L.ev_side[k_last], 0));
  }
  for (int f = first; f < first + count; ++f) c->h_uoff[f] = (uint32_t)((size_t)f * c->nb);   // K4's unique layout
  c->lanes_used = nl;
  c->enc_first = first; c->enc_count = count; c->enc_gop = gop; c->enc_host_out = host_out;
  CU(c, cudaEventRecord(c->ev_end, s0));
  CU(c, cudaGetLastError());
  c->encoded = true;
  return MPTC_OK;
}

}  // namespace

void *mptc::ctx_pinned(mptc_gpu_ctx *c, int slot, size_t bytes) {
  if (!c || slot < 0 || slot >= 4) return nullptr;
  if (c->pinned_bytes[slot] >= bytes && c->pinned[slot]) return c->pinned[slot];
  cudaSetDevice(c->device);
  if (c->pinned[slot]) {
    cudaDeviceSynchronize();   // nothing may still be copying into the old buffer
    cudaFreeHost(c->pinned[slot]);
  }
  c->pinned[slot] = nullptr;
  c->pinned_bytes[slot] = 0;
  if (cudaHostAlloc(&c->pinned[slot], bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  c->pinned_bytes[slot] = bytes;
  return c->pinned[slot];
}

extern "C" {

int mptc_gpu_create(int device, mptc_gpu_ctx **out) {
  if (!out) return MPTC_E_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return MPTC_E_CUDA;
  mptc_gpu_ctx *c = new mptc_gpu_ctx;
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete c; return MPTC_E_CUDA; }
  bool ok = cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreate(&c->ev_begin) == cudaSuccess && cudaEventCreate(&c->ev_end) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->ev_uploaded, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->ev_encoded, cudaEventDisableTiming) == cudaSuccess &&
            cudaMalloc(&c->d_cand, kWorkCounters * sizeof(unsigned long long)) == cudaSuccess;
  if (ok) {
    uint8_t t5[512], t6[512];
    build_match_table(t5, 5);
    build_match_table(t6, 6);
    ok = upload_tables(t5, t6) == cudaSuccess && decode_kernels_init() == cudaSuccess;
  }
  if (!ok) { mptc_gpu_destroy(c); return MPTC_E_CUDA; }
  c->max_wave_ctas = intra_wavefront_max_ctas(device);
  c->lanes_wanted = env_int("MPTC_LANES", 0);
  c->wave_rows_intra = env_int("MPTC_WAVE_ROWS_INTRA", 0);
  c->wave_rows_inter = env_int("MPTC_WAVE_ROWS_INTER", 0);
  c->sparse_ctas = env_int("MPTC_SPARSE_CTAS", 0);
  c->sparse_max_pct = env_int("MPTC_SPARSE_MAX_PCT", 50);
  if (c->sparse_max_pct < 0) c->sparse_max_pct = 0;
  *out = c;
  return MPTC_OK;
}

void mptc_gpu_destroy(mptc_gpu_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  free_seq(c);
  cudaFree(c->d_cand);
  cudaFree(c->d_pattern);
  for (void *p : c->pinned) if (p) cudaFreeHost(p);
  for (auto &L : c->lanes) {
    for (auto &e : L.stage_events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    cudaFree(L.d_tickets);
    for (auto *ev : {&L.ev_up, &L.ev_k, &L.ev_side, &L.ev_down})
      for (cudaEvent_t e : *ev) cudaEventDestroy(e);
    if (L.s) cudaStreamDestroy(L.s);
    if (L.t) cudaStreamDestroy(L.t);
  }
  for (cudaEvent_t e : c->ev_dec) if (e) cudaEventDestroy(e);
  if (c->ev_begin) cudaEventDestroy(c->ev_begin);
  if (c->ev_end) cudaEventDestroy(c->ev_end);
  if (c->ev_uploaded) cudaEventDestroy(c->ev_uploaded);
  if (c->ev_encoded) cudaEventDestroy(c->ev_encoded);
  if (c->s_compute) cudaStreamDestroy(c->s_compute);
  if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  delete c;
}

const char *mptc_gpu_last_error(const mptc_gpu_ctx *c) { return c ? c->err : "null context"; }
uint64_t mptc_gpu_launch_count(const mptc_gpu_ctx *c) { return c ? c->launches : 0; }

int mptc_gpu_set_schedule(mptc_gpu_ctx *c, int lanes, int wave_rows_intra, int wave_rows_inter) {
  if (!c || lanes < 0 || lanes > 16 || wave_rows_intra < 0 || wave_rows_inter < 0) return MPTC_E_ARG;
  c->lanes_wanted = lanes;
  c->wave_rows_intra = wave_rows_intra;
  c->wave_rows_inter = wave_rows_inter;
  return MPTC_OK;
}

int mptc_gpu_seq_reserve(mptc_gpu_ctx *c, int w, int h, int n_frames) {
  if (!c) return MPTC_E_ARG;
  if (w < 4 || h < 4 || (w & 3) || (h & 3)) return fail(c, MPTC_E_ARG, "frame %dx%d: width/height must be multiples of 4", w, h);
  if (n_frames < 1) return fail(c, MPTC_E_ARG, "n_frames %d < 1", n_frames);
  CU(c, cudaSetDevice(c->device));
  if (c->w == w && c->h == h && c->cap_frames >= n_frames) return MPTC_OK;
  CU(c, cudaDeviceSynchronize());
  free_seq(c);
  c->w = w; c->h = h; c->bw = w / 4; c->bh = h / 4; c->nb = c->bw * c->bh;
  c->pbw = (c->bw + 63) / 64 * 64; c->pbh = (c->bh + 63) / 64 * 64;
  c->frame_bytes = (size_t)w * h * 3;
  c->plane_bytes = (size_t)6 * c->pbw * c->pbh;
  const size_t F = (size_t)n_frames, nb = (size_t)c->nb;
  CU(c, cudaMalloc(&c->d_rgb, F * c->frame_bytes));
  CU(c, cudaMalloc(&c->d_init, F * nb * 8));
  CU(c, cudaMalloc(&c->d_final, F * nb * 8));
  CU(c, cudaMalloc(&c->d_motion, F * nb * 2));
  CU(c, cudaMalloc(&c->d_flags, F * nb));
  CU(c, cudaMalloc(&c->d_row_todo, F * c->bh));
  CU(c, cudaMalloc(&c->d_unique, F * nb * 4));
  CU(c, cudaMalloc(&c->d_nunique, F * 4));
  CU(c, cudaMalloc(&c->d_chunks, F * ((nb + 1023) / 1024) * 4));
  CU(c, cudaMalloc(&c->d_planes, F * c->plane_bytes));
  CU(c, cudaMalloc(&c->d_progress, F * c->bh * sizeof(int)));
  CU(c, cudaMalloc(&c->d_wordflag, F * nb * sizeof(unsigned long long)));
  CU(c, cudaMemset(c->d_wordflag, 0, F * nb * sizeof(unsigned long long)));   // epoch 0 is never used
  c->epoch = 0;
  c->cap_frames = n_frames;
  c->h_uoff.resize(F);
  for (size_t f = 0; f < F; ++f) c->h_uoff[f] = (uint32_t)(f * nb);   // the encoder's layout
  return MPTC_OK;
}

int mptc_gpu_seq_upload(mptc_gpu_ctx *c, const uint8_t *frames, int first, int count) {
  if (!c || !frames) return MPTC_E_ARG;
  if (first < 0 || count < 1 || first + count > c->cap_frames) return fail(c, MPTC_E_STATE, "upload range [%d,%d) outside reserved %d frames", first, first + count, c->cap_frames);
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaMemcpyAsync(c->d_rgb + c->frame_bytes * first, frames, c->frame_bytes * count, cudaMemcpyHostToDevice, c->s_compute));
  return MPTC_OK;
}

int mptc_gpu_seq_encode(mptc_gpu_ctx *c, int first, int count, const mptc_gpu_params *p) {
  if (!c || !p) return MPTC_E_ARG;
  if (int r = check_params(c, p->search_area, p->gop)) return r;
  if (first < 0 || count < 1 || first + count > c->cap_frames) return fail(c, MPTC_E_STATE, "encode range [%d,%d) outside reserved %d frames", first, first + count, c->cap_frames);
  CU(c, cudaSetDevice(c->device));
  return run_encode(c, first, count, p->gop, p->search_area, p->err_threshold, true, 0, true);
}

int mptc_gpu_seq_download(mptc_gpu_ctx *c, int first, int count, uint64_t *blocks, uint64_t *initial,
                          uint8_t *motion, uint32_t *unique, uint32_t *n_unique, uint8_t *planes) {
  if (!c) return MPTC_E_ARG;
  if (first < 0 || count < 1 || first + count > c->cap_frames) return fail(c, MPTC_E_STATE, "download range [%d,%d) outside reserved %d frames", first, first + count, c->cap_frames);
  CU(c, cudaSetDevice(c->device));
  cudaStream_t s = c->s_compute;
  const size_t nb = (size_t)c->nb, f0 = (size_t)first, n = (size_t)count;
  if (blocks) CU(c, cudaMemcpyAsync(blocks, c->d_final + f0 * nb, n * nb * 8, cudaMemcpyDeviceToHost, s));
  if (initial) CU(c, cudaMemcpyAsync(initial, c->d_init + f0 * nb, n * nb * 8, cudaMemcpyDeviceToHost, s));
  if (motion) CU(c, cudaMemcpyAsync(motion, c->d_motion + f0 * nb * 2, n * nb * 2, cudaMemcpyDeviceToHost, s));
  if (unique) CU(c, cudaMemcpyAsync(unique, c->d_unique + f0 * nb, n * nb * 4, cudaMemcpyDeviceToHost, s));
  if (n_unique) CU(c, cudaMemcpyAsync(n_unique, c->d_nunique + f0, n * 4, cudaMemcpyDeviceToHost, s));
  if (planes) CU(c, cudaMemcpyAsync(planes, c->d_planes + f0 * c->plane_bytes, n * c->plane_bytes, cudaMemcpyDeviceToHost, s));
  CU(c, cudaStreamSynchronize(s));
  return MPTC_OK;
}

int mptc_gpu_sync(mptc_gpu_ctx *c) {
  if (!c) return MPTC_E_ARG;
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaStreamSynchronize(c->s_compute));
  CU(c, cudaStreamSynchronize(c->s_h2d));
  CU(c, cudaStreamSynchronize(c->s_d2h));
  return MPTC_OK;
}

int mptc_gpu_last_encode_ms(mptc_gpu_ctx *c, int stage, float *ms) {
  if (!c || !ms) return MPTC_E_ARG;
  if (!c->encoded) return fail(c, MPTC_E_STATE, "no encode has run");
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaEventSynchronize(c->ev_end));
  if (stage == 0) {
    CU(c, cudaEventElapsedTime(ms, c->ev_begin, c->ev_end));
    return MPTC_OK;
  }
  float total = 0.f;
  for (int l = 0; l < c->lanes_used; ++l) {
    const Lane &L = c->lanes[l];
    for (size_t i = 0; i < L.stage_events_used; ++i) {
      if (L.stage_events[i].stage != stage) continue;
      float t = 0.f;
      CU(c, cudaEventElapsedTime(&t, L.stage_events[i].a, L.stage_events[i].b));
      total += t;
    }
  }
  *ms = total;
  return MPTC_OK;
}

int mptc_gpu_last_candidate_count(mptc_gpu_ctx *c, uint64_t *inter, uint64_t *intra) {
  if (!c) return MPTC_E_ARG;
  if (!c->encoded) return fail(c, MPTC_E_STATE, "no encode has run");
  CU(c, cudaSetDevice(c->device));
  unsigned long long h[2];
  CU(c, cudaStreamSynchronize(c->s_compute));
  CU(c, cudaMemcpy(h, c->d_cand, sizeof h, cudaMemcpyDeviceToHost));
  if (inter) *inter = h[0];
  if (intra) *intra = h[1];
  return MPTC_OK;
}

int mptc_gpu_last_work_count(mptc_gpu_ctx *c, uint64_t *counters, int n) {
  if (!c || !counters || n < 1) return MPTC_E_ARG;
  if (!c->encoded) return fail(c, MPTC_E_STATE, "no encode has run");
  CU(c, cudaSetDevice(c->device));
  unsigned long long h[kWorkCounters];
  CU(c, cudaStreamSynchronize(c->s_compute));
  CU(c, cudaMemcpy(h, c->d_cand, sizeof h, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i) counters[i] = i < kWorkCounters ? h[i] : 0;
  return MPTC_OK;
}

int mptc_gpu_dxt1_fit(mptc_gpu_ctx *c, const uint8_t *rgb, int w, int h, uint64_t *blocks_out) {
  if (!c || !rgb || !blocks_out) return MPTC_E_ARG;
  if (int r = mptc_gpu_seq_reserve(c, w, h, 2)) return r;
  if (int r = mptc_gpu_seq_upload(c, rgb, 0, 1)) return r;
  SeqView v = view_of(c, 0, 1, 1);
  launch_dxt1_fit(v, 0, 1, 1, c->s_compute);
  ++c->launches;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(blocks_out, c->d_init, (size_t)c->nb * 8, cudaMemcpyDeviceToHost, c->s_compute));
  CU(c, cudaStreamSynchronize(c->s_compute));
  return MPTC_OK;
}

int mptc_gpu_reencode(mptc_gpu_ctx *c, const uint8_t *rgb, int w, int h, int is_intra, int search_area,
                      int err_threshold, const uint64_t *prev_blocks, uint64_t *initial_out,
                      uint64_t *blocks_out, uint8_t *motion_out, uint32_t *unique_out, uint32_t *n_unique) {
  if (!c || !rgb) return MPTC_E_ARG;
  if (int r = check_params(c, search_area, 1)) return r;
  if (!is_intra && !prev_blocks) return fail(c, MPTC_E_ARG, "inter frame needs prev_blocks (reference->_physical_blocks)");
  if (int r = mptc_gpu_seq_reserve(c, w, h, 2)) return r;
  // slot 0 = the reference frame (only its final blocks matter), slot 1 = this frame
  const int slot = is_intra ? 0 : 1;
  if (int r = mptc_gpu_seq_upload(c, rgb, slot, 1)) return r;
  if (!is_intra)
    CU(c, cudaMemcpyAsync(c->d_final, prev_blocks, (size_t)c->nb * 8, cudaMemcpyHostToDevice, c->s_compute));
  {
    SeqView v = view_of(c, slot, 1, 1);
    launch_dxt1_fit(v, slot, 1, 1, c->s_compute);
    ++c->launches;
  }
  int r = is_intra ? run_encode(c, 0, 1, 1, search_area, err_threshold, false, 0, false)
                   : run_encode(c, 0, 2, 2, search_area, err_threshold, false, 1, false);
  if (r) return r;
  const size_t nb = (size_t)c->nb, off = (size_t)slot * nb;
  cudaStream_t s = c->s_compute;
  if (initial_out) CU(c, cudaMemcpyAsync(initial_out, c->d_init + off, nb * 8, cudaMemcpyDeviceToHost, s));
  if (blocks_out) CU(c, cudaMemcpyAsync(blocks_out, c->d_final + off, nb * 8, cudaMemcpyDeviceToHost, s));
  if (motion_out) CU(c, cudaMemcpyAsync(motion_out, c->d_motion + off * 2, nb * 2, cudaMemcpyDeviceToHost, s));
  uint32_t nu = 0;
  CU(c, cudaMemcpyAsync(&nu, c->d_nunique + slot, 4, cudaMemcpyDeviceToHost, s));
  CU(c, cudaStreamSynchronize(s));
  if (unique_out && nu) CU(c, cudaMemcpy(unique_out, c->d_unique + off, (size_t)nu * 4, cudaMemcpyDeviceToHost));
  if (n_unique) *n_unique = nu;
  return MPTC_OK;
}

int mptc_gpu_inter_pixel_search(mptc_gpu_ctx *c, const uint8_t *rgb, int w, int h, int search_area,
                                const uint64_t *cur_blocks, const uint64_t *prev_blocks, int32_t *min_err_out,
                                uint8_t *motion_out, uint32_t *index_out, uint8_t *reassigned_out) {
  if (!c || !rgb || !prev_blocks) return MPTC_E_ARG;
  if (int r = check_params(c, search_area, 1)) return r;
  if (int r = mptc_gpu_seq_reserve(c, w, h, 2)) return r;
  cudaStream_t s = c->s_compute;
  if (c->pattern_sa != search_area) {   // the ring pattern of DXTImage::SetPattern, once per search area
    std::vector<int8_t> pat((size_t)2 * inter_pixel_pattern(search_area, nullptr));
    c->pattern_n = inter_pixel_pattern(search_area, pat.data());
    CU(c, cudaStreamSynchronize(s));
    cudaFree(c->d_pattern);
    c->d_pattern = nullptr;
    c->pattern_sa = 0;
    CU(c, cudaMalloc(&c->d_pattern, pat.size()));
    CU(c, cudaMemcpy(c->d_pattern, pat.data(), pat.size(), cudaMemcpyHostToDevice));
    c->pattern_sa = search_area;
  }
  // slot 0 = the reference frame (only its final blocks matter), slot 1 = this frame
  const size_t nb = (size_t)c->nb;
  if (int r = mptc_gpu_seq_upload(c, rgb, 1, 1)) return r;
  CU(c, cudaMemcpyAsync(c->d_final, prev_blocks, nb * 8, cudaMemcpyHostToDevice, s));
  if (cur_blocks) {
    CU(c, cudaMemcpyAsync(c->d_init + nb, cur_blocks, nb * 8, cudaMemcpyHostToDevice, s));
  } else {   // the block's state before Reencode: the stb fit
    SeqView v = view_of(c, 1, 1, 1);
    launch_dxt1_fit(v, 1, 1, 1, s);
    ++c->launches;
  }
  int32_t *d_err = reinterpret_cast<int32_t *>(c->d_unique);   // scratch: the sequence's result slots
  uint32_t *d_index = c->d_unique + nb;
  uint8_t *d_mo = c->d_motion + nb * 2, *d_re = c->d_flags + nb;
  launch_inter_pixel_search(c->d_rgb + c->frame_bytes, w, h, search_area, c->pattern_n, c->d_pattern, c->d_init + nb, c->d_final,
                            d_err, d_mo, d_index, d_re, s);
  ++c->launches;
  CU(c, cudaGetLastError());
  if (min_err_out) CU(c, cudaMemcpyAsync(min_err_out, d_err, nb * 4, cudaMemcpyDeviceToHost, s));
  if (motion_out) CU(c, cudaMemcpyAsync(motion_out, d_mo, nb * 2, cudaMemcpyDeviceToHost, s));
  if (index_out) CU(c, cudaMemcpyAsync(index_out, d_index, nb * 4, cudaMemcpyDeviceToHost, s));
  if (reassigned_out) CU(c, cudaMemcpyAsync(reassigned_out, d_re, nb, cudaMemcpyDeviceToHost, s));
  CU(c, cudaStreamSynchronize(s));
  c->encoded = false;   // the scratch slots no longer hold an encode's results
  return MPTC_OK;
}

int mptc_gpu_endpoint_planes(mptc_gpu_ctx *c, const uint64_t *blocks, int bw, int bh, uint8_t *planes_out) {
  if (!c || !blocks || !planes_out) return MPTC_E_ARG;
  if (bw < 1 || bh < 1) return fail(c, MPTC_E_ARG, "bad plane size %dx%d", bw, bh);
  if (int r = mptc_gpu_seq_reserve(c, bw * 4, bh * 4, 2)) return r;
  cudaStream_t s = c->s_compute;
  CU(c, cudaMemcpyAsync(c->d_final, blocks, (size_t)c->nb * 8, cudaMemcpyHostToDevice, s));
  SeqView v = view_of(c, 0, 1, 1);
  launch_endpoint_planes(v, c->pbw, c->pbh, 0, 1, 1, s);
  ++c->launches;
  CU(c, cudaGetLastError());
  CU(c, cudaMemcpyAsync(planes_out, c->d_planes, c->plane_bytes, cudaMemcpyDeviceToHost, s));
  CU(c, cudaStreamSynchronize(s));
  return MPTC_OK;
}

int mptc_gpu_encode_sequence_async(mptc_gpu_ctx *c, const uint8_t *frames, int n_frames, int w, int h,
                                   const mptc_gpu_params *p, uint64_t *blocks, uint8_t *motion, uint32_t *unique,
                                   uint32_t *n_unique, uint8_t *planes) {
  if (!c || !frames || !p) return MPTC_E_ARG;
  if (int r = check_params(c, p->search_area, p->gop)) return r;
  if (int r = mptc_gpu_seq_reserve(c, w, h, n_frames)) return r;
  HostIO io;
  io.frames = frames; io.blocks = blocks; io.motion = motion; io.unique = unique; io.n_unique = n_unique; io.planes = planes;
  if (env_int("MPTC_UNIQUE_COPY", 0) == 0) {   // 1 = always D2H-copy the nb-word slots (A/B measurements)
    io.unique_mapped = mapped_alias(unique);
    io.n_unique_mapped = mapped_alias(n_unique);
    if (!io.unique_mapped || !io.n_unique_mapped) io.unique_mapped = io.n_unique_mapped = nullptr;
  }
  return run_encode(c, 0, n_frames, p->gop, p->search_area, p->err_threshold, true, 0, planes != nullptr, &io);
}

int mptc_gpu_wait_frame(mptc_gpu_ctx *c, int frame) {
  if (!c) return MPTC_E_ARG;
  if (!c->encoded) return fail(c, MPTC_E_STATE, "no encode has run");
  if (frame < c->enc_first || frame >= c->enc_first + c->enc_count)
    return fail(c, MPTC_E_ARG, "frame %d outside the last encode [%d,%d)", frame, c->enc_first, c->enc_first + c->enc_count);
  CU(c, cudaSetDevice(c->device));
  for (int i = 0; i < c->lanes_used; ++i) {
    const Lane &L = c->lanes[i];
    if (frame < L.f0 || frame >= L.f0 + L.n) continue;
    const int k = (frame - L.f0) % c->enc_gop;
    CU(c, cudaEventSynchronize(c->enc_host_out ? L.ev_down[k] : L.ev_side[k]));
    return MPTC_OK;
  }
  return fail(c, MPTC_E_STATE, "frame %d not found in any lane", frame);
}

int mptc_gpu_wait(mptc_gpu_ctx *c) {
  if (!c) return MPTC_E_ARG;
  if (!c->encoded) return fail(c, MPTC_E_STATE, "no encode has run");
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaEventSynchronize(c->ev_end));
  return MPTC_OK;
}

int mptc_gpu_encode_sequence(mptc_gpu_ctx *c, const uint8_t *frames, int n_frames, int w, int h,
                             const mptc_gpu_params *p, uint64_t *blocks, uint8_t *motion, uint32_t *unique,
                             uint32_t *n_unique, uint8_t *planes) {
  if (int r = mptc_gpu_encode_sequence_async(c, frames, n_frames, w, h, p, blocks, motion, unique, n_unique, planes)) return r;
  CU(c, cudaStreamSynchronize(c->s_compute));
  return MPTC_OK;
}

// ---- decoder side -------------------------------------------------------------------------

int mptc_gpu_seq_decode_upload(mptc_gpu_ctx *c, int first, int count, const uint8_t *motion, const uint32_t *unique,
                               const uint32_t *n_unique, size_t unique_stride, const uint8_t *planes) {
  if (!c || !motion || !n_unique || !planes) return MPTC_E_ARG;
  if (first < 0 || count < 1 || first + count > c->cap_frames) return fail(c, MPTC_E_STATE, "decode upload range [%d,%d) outside reserved %d frames", first, first + count, c->cap_frames);
  CU(c, cudaSetDevice(c->device));
  if (int r = ensure_decode(c, false)) return r;
  cudaStream_t s = c->s_compute;
  const size_t nb = (size_t)c->nb, f0 = (size_t)first, n = (size_t)count;
  const cudaMemcpyKind H2D = cudaMemcpyHostToDevice;
  CU(c, cudaMemcpyAsync(c->d_motion + f0 * nb * 2, motion, n * nb * 2, H2D, s));
  CU(c, cudaMemcpyAsync(c->d_planes + f0 * c->plane_bytes, planes, n * c->plane_bytes, H2D, s));
  CU(c, cudaMemcpyAsync(c->d_nunique + f0, n_unique, n * 4, H2D, s));
  size_t total = 0;
  for (size_t i = 0; i < n; ++i) {
    if (n_unique[i] > nb) return fail(c, MPTC_E_DATA, "frame %zu: %u unique words for %zu blocks", f0 + i, n_unique[i], nb);
    c->h_uoff[f0 + i] = (uint32_t)(unique_stride ? (f0 + i) * nb : f0 * nb + total);
    total += n_unique[i];
  }
  if (total && !unique) return MPTC_E_ARG;
  if (unique_stride == 0) {          // packed: the group palettes back to back (codec.cpp:1473-1479)
    if (total) CU(c, cudaMemcpyAsync(c->d_unique + f0 * nb, unique, total * 4, H2D, s));
  } else {                           // the encoder's layout: frame i's words at unique + i*stride
    for (size_t i = 0; i < n; ++i)
      if (n_unique[i]) CU(c, cudaMemcpyAsync(c->d_unique + (f0 + i) * nb, unique + i * unique_stride, (size_t)n_unique[i] * 4, H2D, s));
  }
  return MPTC_OK;
}

int mptc_gpu_seq_decode(mptc_gpu_ctx *c, int first, int count, int search_area, int gop, int want_rgb) {
  if (!c) return MPTC_E_ARG;
  if (int r = check_params(c, search_area, gop)) return r;
  if (first < 0 || count < 1 || first + count > c->cap_frames) return fail(c, MPTC_E_STATE, "decode range [%d,%d) outside reserved %d frames", first, first + count, c->cap_frames);
  CU(c, cudaSetDevice(c->device));
  if (int r = ensure_decode(c, want_rgb != 0)) return r;
  cudaStream_t s = c->s_compute;
  const int passes = dec_jump_passes(gop, c->nb);
  if (passes + 2 > kDecStatusStride) return fail(c, MPTC_E_ARG, "gop %d x %d blocks: chains too long", gop, c->nb);
  DecView v = dec_view_of(c, first, count, gop, search_area);
  CU(c, cudaMemcpyAsync(c->d_dec_uoff + first, c->h_uoff.data() + first, (size_t)count * 4, cudaMemcpyHostToDevice, s));
  if (!c->dec_pending) CU(c, cudaMemsetAsync(c->d_dec_status, 0, sizeof(int), s));
  CU(c, cudaMemsetAsync(v.status, 0, kDecStatusStride * sizeof(int), s));
  CU(c, cudaEventRecord(c->ev_dec[0], s));
  c->launches += launch_decode_words(v, s);
  CU(c, cudaEventRecord(c->ev_dec[1], s));
  c->launches += launch_inverse_planes(v, s);
  CU(c, cudaEventRecord(c->ev_dec[2], s));
  if (want_rgb) c->launches += launch_dxt1_to_rgb(v, s);
  CU(c, cudaEventRecord(c->ev_dec[3], s));
  CU(c, cudaGetLastError());
  c->decoded = c->dec_pending = true;
  c->dec_rgb = want_rgb != 0;
  return MPTC_OK;
}

int mptc_gpu_seq_decode_download(mptc_gpu_ctx *c, int first, int count, uint64_t *blocks, uint8_t *rgb) {
  if (!c) return MPTC_E_ARG;
  if (!c->decoded) return fail(c, MPTC_E_STATE, "no decode has run");
  if (first < 0 || count < 1 || first + count > c->cap_frames) return fail(c, MPTC_E_STATE, "download range [%d,%d) outside reserved %d frames", first, first + count, c->cap_frames);
  if (rgb && !c->d_dec_rgb) return fail(c, MPTC_E_STATE, "the decode did not produce RGB");
  CU(c, cudaSetDevice(c->device));
  cudaStream_t s = c->s_compute;
  const size_t nb = (size_t)c->nb;
  if (blocks) CU(c, cudaMemcpyAsync(blocks, c->d_dec_blocks + (size_t)first * nb, (size_t)count * nb * 8, cudaMemcpyDeviceToHost, s));
  if (rgb) CU(c, cudaMemcpyAsync(rgb, c->d_dec_rgb + (size_t)first * c->frame_bytes, (size_t)count * c->frame_bytes, cudaMemcpyDeviceToHost, s));
  int bad = 0;
  CU(c, cudaMemcpyAsync(&bad, c->d_dec_status, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(c, cudaStreamSynchronize(s));
  c->dec_pending = false;
  if (bad) return fail(c, MPTC_E_DATA, "corrupt stream: %d blocks with a motion vector no encoder emits", bad);
  return MPTC_OK;
}

int mptc_gpu_decode_sequence(mptc_gpu_ctx *c, const uint8_t *motion, const uint32_t *unique, const uint32_t *n_unique,
                             size_t unique_stride, const uint8_t *planes, int n_frames, int w, int h,
                             int search_area, int gop, uint64_t *blocks_out, uint8_t *rgb_out) {
  if (!c) return MPTC_E_ARG;
  if (int r = check_params(c, search_area, gop)) return r;
  if (int r = mptc_gpu_seq_reserve(c, w, h, n_frames)) return r;
  if (int r = mptc_gpu_seq_decode_upload(c, 0, n_frames, motion, unique, n_unique, unique_stride, planes)) return r;
  if (int r = mptc_gpu_seq_decode(c, 0, n_frames, search_area, gop, rgb_out != nullptr)) return r;
  return mptc_gpu_seq_decode_download(c, 0, n_frames, blocks_out, rgb_out);
}

int mptc_gpu_last_decode_ms(mptc_gpu_ctx *c, int stage, float *ms) {
  if (!c || !ms || stage < 0 || stage > 3) return MPTC_E_ARG;
  if (!c->decoded) return fail(c, MPTC_E_STATE, "no decode has run");
  CU(c, cudaSetDevice(c->device));
  CU(c, cudaEventSynchronize(c->ev_dec[3]));
  if (stage == 0) CU(c, cudaEventElapsedTime(ms, c->ev_dec[0], c->ev_dec[3]));
  else CU(c, cudaEventElapsedTime(ms, c->ev_dec[stage - 1], c->ev_dec[stage]));
  return MPTC_OK;
}

void *mptc_gpu_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}

void mptc_gpu_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
