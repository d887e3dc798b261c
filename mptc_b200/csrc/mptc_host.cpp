// mptc_host.cpp -- the host side above the GPU hot path: arithmetic coding and stream assembly
// (include/mptc_codec.h).  The coder is inherently sequential per stream, so it stays on the
// host (BASELINE.json north_star) and is parallelised ACROSS streams: every frame has five
// independent streams and every group one palette stream (codec/codec.cpp:1115-1158, :1479).
#include "../../include/mptc_codec.h"
#include "mptc_host.h"

#include <emmintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <type_traits>
#include <thread>

namespace mptc {

namespace {
constexpr uint32_t kMinLength = 0x01000000u;   // renormalisation threshold (AC__MinLength)
constexpr unsigned kLengthShift = 15;          // DM__LengthShift
}  // namespace

// ---- scratch for the coder's output ----------------------------------------------------------
// A stream's code is at most encode_bound(n) bytes but typically ~0.3 n.  Coding into a fresh
// std::vector of the bound means zero-filling and page-faulting 2n bytes per stream -- with 16
// host threads that (the kernel's page-fault path) cost more than the coding.  The pool below keeps
// raw buffers alive across calls: pages are touched once, workers take a buffer for the duration
// of a call and copy only the bytes actually produced.
namespace {
std::mutex g_scratch_mu;
std::vector<Scratch> g_scratch_free;
}  // namespace

Scratch scratch_acquire(size_t bytes) {
  Scratch s;
  {
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    size_t best = g_scratch_free.size();
    for (size_t i = 0; i < g_scratch_free.size(); ++i)
      if (best == g_scratch_free.size() || g_scratch_free[i].cap > g_scratch_free[best].cap) best = i;
    if (best < g_scratch_free.size()) {
      s = g_scratch_free[best];
      g_scratch_free.erase(g_scratch_free.begin() + (long)best);
    }
  }
  if (s.cap < bytes) {
    free(s.p);
    s.p = static_cast<uint8_t *>(malloc(bytes));
    s.cap = s.p ? bytes : 0;
  }
  return s;
}

void scratch_release(Scratch s) {
  if (!s.p) return;
  std::lock_guard<std::mutex> lock(g_scratch_mu);
  g_scratch_free.push_back(s);
}

// Byte symbols must never be the model's last symbol (step() codes the general branch only).
RangeEncoder::RangeEncoder(unsigned symbols)
    : n_(symbols < 257 ? 257 : symbols), dist_(n_), count_(n_) {}

// Adaptive_Data_Model::reset (arithmetic_codec.cpp:818-829)
void RangeEncoder::reset_model() {
  total_ = 0;
  cycle_ = n_;
  for (unsigned k = 0; k < n_; ++k) count_[k] = 1;
  update_model();
  until_ = cycle_ = (n_ + 6) >> 1;
}

// Adaptive_Data_Model::update, encoder flavour (arithmetic_codec.cpp:781-816)
void RangeEncoder::update_model() {
  if ((total_ += cycle_) > (1u << kLengthShift)) {
    total_ = 0;
    for (unsigned k = 0; k < n_; ++k) total_ += (count_[k] = (count_[k] + 1) >> 1);
  }
  const uint32_t scale = 0x80000000u / total_;
  uint32_t sum = 0;
  for (unsigned k = 0; k < n_; ++k) {
    dist_[k] = (scale * sum) >> (31 - kLengthShift);
    sum += count_[k];
  }
  cycle_ = (5 * cycle_) >> 2;
  const uint32_t max_cycle = (n_ + 6) << 3;
  if (cycle_ > max_cycle) cycle_ = max_cycle;
  until_ = cycle_;
}

// Coder state kept in locals while symbols are coded: the byte stores may alias the members.
struct RangeEncoder::State {
  uint32_t base, length, until;
  uint8_t *p, *buf;
  const uint32_t *dist;
  uint32_t *count;
};

// buf: at least encode_bound(n) bytes.
inline void RangeEncoder::begin(State &e, uint8_t *buf) {
  reset_model();
  e.buf = e.p = buf;
  e.base = 0;
  e.length = 0xFFFFFFFFu;
  e.until = until_;
  e.dist = dist_.data();
  e.count = count_.data();
}

// propagate_carry (arithmetic_codec.cpp:81-86)
static inline void carry(uint8_t *p) {
  uint8_t *q = p - 1;
  while (*q == 0xFFu) *q-- = 0;
  ++*q;
}

// Arithmetic_Codec::encode (arithmetic_codec.cpp:360-387) for one byte symbol (never the model's
// last symbol, 256, so only the general branch of :366-376 is needed).
inline void RangeEncoder::step(State &e, uint32_t s) {
  const uint32_t before = e.base;
  const uint32_t len = e.length >> kLengthShift;
  const uint32_t x = e.dist[s] * len;
  e.base += x;
  e.length = e.dist[s + 1] * len - x;
  if (before > e.base) carry(e.p);
  // renorm_enc_interval (:90-96) without the data-dependent loop: every symbol keeps a range of at
  // least one model unit, so length >= 2^9 here and at most two bytes leave; the byte count comes
  // from the leading zeros, both candidate bytes are stored unconditionally (the buffer has slack
  // and bytes past p are rewritten before they count).
  const unsigned nsh = (unsigned)__builtin_clz(e.length) >> 3;
  e.p[0] = (uint8_t)(e.base >> 24);
  e.p[1] = (uint8_t)(e.base >> 16);
  e.p += nsh;
  e.base <<= 8 * nsh;
  e.length <<= 8 * nsh;
  ++e.count[s];
  if (--e.until == 0) {
    update_model();
    e.until = until_;
  }
}

// stop_encoder (arithmetic_codec.cpp:547-571)
inline size_t RangeEncoder::finish(State &e) {
  until_ = e.until;
  const uint32_t before = e.base;
  if (e.length > 2 * kMinLength) {
    e.base += kMinLength;
    e.length = kMinLength >> 1;
  } else {
    e.base += kMinLength >> 1;
    e.length = kMinLength >> 9;
  }
  if (before > e.base) carry(e.p);
  do {                                      // renorm_enc_interval
    *e.p++ = (uint8_t)(e.base >> 24);
    e.base <<= 8;
  } while ((e.length <<= 8) < kMinLength);
  return (size_t)(e.p - e.buf);
}

size_t RangeEncoder::encode_raw(const uint8_t *sym, size_t n, uint8_t *buf) {
  State e;
  begin(e, buf);
  for (size_t i = 0; i < n; ++i) step(e, sym[i]);
  return finish(e);
}

void RangeEncoder::encode_all(const uint8_t *sym, size_t n, std::vector<uint8_t> &out) {
  Scratch sc = scratch_acquire(encode_bound(n));
  if (!sc.p) throw std::bad_alloc();   // reported as MPTC_E_NOMEM at the C boundary
  const size_t len = encode_raw(sym, n, sc.p);
  out.insert(out.end(), sc.p, sc.p + len);
  scratch_release(sc);
}

// Two independent streams in one loop: the coder is a chain of dependent multiplies and shifts
// (~10 cycles per symbol), so a second chain in flight nearly doubles a thread's throughput.
void RangeEncoder::encode_pair_raw(RangeEncoder &ma, const uint8_t *sa, size_t na, uint8_t *bufa, size_t *lena,
                                   RangeEncoder &mb, const uint8_t *sb, size_t nb, uint8_t *bufb, size_t *lenb) {
  State a, b;
  ma.begin(a, bufa);
  mb.begin(b, bufb);
  const size_t both = na < nb ? na : nb;
  for (size_t i = 0; i < both; ++i) {
    ma.step(a, sa[i]);
    mb.step(b, sb[i]);
  }
  for (size_t i = both; i < na; ++i) ma.step(a, sa[i]);
  for (size_t i = both; i < nb; ++i) mb.step(b, sb[i]);
  *lena = ma.finish(a);
  *lenb = mb.finish(b);
}

void RangeEncoder::encode_pair(RangeEncoder &ma, const uint8_t *sa, size_t na, std::vector<uint8_t> &oa,
                               RangeEncoder &mb, const uint8_t *sb, size_t nb, std::vector<uint8_t> &ob) {
  Scratch sc = scratch_acquire(encode_bound(na) + encode_bound(nb));
  if (!sc.p) throw std::bad_alloc();
  size_t la = 0, lb = 0;
  encode_pair_raw(ma, sa, na, sc.p, &la, mb, sb, nb, sc.p + encode_bound(na), &lb);
  oa.insert(oa.end(), sc.p, sc.p + la);
  ob.insert(ob.end(), sc.p + encode_bound(na), sc.p + encode_bound(na) + lb);
  scratch_release(sc);
}

// ---- decoder ---------------------------------------------------------------------------------
namespace {
constexpr unsigned kTableShift = 8;                            // 128 buckets over the 15-bit range (1024 buckets
                                                               // + branch-free steps measured 28 % slower: rebuilds)
constexpr unsigned kTableSize = (1u << kLengthShift) >> kTableShift;
constexpr unsigned kSearchWidth = 8;                           // cumulative frequencies compared per vector step
}  // namespace

RangeDecoder::RangeDecoder(unsigned symbols)
    : n_(symbols < 257 ? 257 : symbols), dist_(n_ + 1 + kSearchWidth, 0x7FFFFFFFu), count_(n_), start_(kTableSize) {}

void RangeDecoder::reset_model() {     // Adaptive_Data_Model::reset (arithmetic_codec.cpp:818-829)
  total_ = 0;
  cycle_ = n_;
  for (unsigned k = 0; k < n_; ++k) count_[k] = 1;
  update_model();
  until_ = cycle_ = (n_ + 6) >> 1;
}

void RangeDecoder::update_model() {    // Adaptive_Data_Model::update(false) (:781-816)
  if ((total_ += cycle_) > (1u << kLengthShift)) {
    total_ = 0;
    for (unsigned k = 0; k < n_; ++k) total_ += (count_[k] = (count_[k] + 1) >> 1);
  }
  const uint32_t scale = 0x80000000u / total_;
  uint32_t sum = 0;
  unsigned t = 0;
  for (unsigned k = 0; k < n_; ++k) {
    dist_[k] = (scale * sum) >> (31 - kLengthShift);
    sum += count_[k];
    for (; t < kTableSize && (t << kTableShift) < dist_[k]; ++t) start_[t] = (uint16_t)(k - 1);
  }
  for (; t < kTableSize; ++t) start_[t] = (uint16_t)(n_ - 1);
  dist_[n_] = 1u << kLengthShift;      // sentinel: the scan below stops at the last symbol
  cycle_ = (5 * cycle_) >> 2;
  const uint32_t max_cycle = (n_ + 6) << 3;
  if (cycle_ > max_cycle) cycle_ = max_cycle;
  until_ = cycle_;
}

// Decoder state kept in locals while symbols are decoded.
struct RangeDecoder::State {
  uint32_t value, length, until;
  const uint8_t *p, *end;
  size_t overrun;
  bool ok;
  uint32_t next() {
    if (p < end) return *p++;
    ++overrun;
    return 0;
  }
};

inline void RangeDecoder::begin(State &d, const uint8_t *code, size_t nbytes) {
  reset_model();
  d.p = code;
  d.end = code + nbytes;
  d.overrun = 0;
  d.ok = true;
  d.value = 0;
  d.length = 0xFFFFFFFFu;
  for (int i = 0; i < 4; ++i) d.value = (d.value << 8) | d.next();   // start_decoder
  d.until = until_;
}

// Arithmetic_Codec::decode(Adaptive_Data_Model &) (arithmetic_codec.cpp:391-444), without data-dependent
// branches in the common case: the symbol search compares eight cumulative frequencies at once (SSE2; they
// ascend strictly, so the first one above the target ends the search), the renormalisation shifts in zero,
// one or two bytes at once (the interval is at least 2^9 wide after a step, so never three).  The branchy
// form -- `while (dist[s + 1] <= dv) ++s;` and a byte-wise renormalisation loop -- mispredicted about twice
// per symbol, and a misprediction throws away the work of all interleaved streams.
// (Measured and dropped: replacing the 32-bit division by a float-reciprocal estimate, 57 -> 42 Msym/s, and by an
// exact double-precision division, 85 -> 52 Msym/s with four streams on the build container's Xeon.)
inline uint8_t RangeDecoder::step(State &d) {
  const uint32_t *dist = dist_.data();
  const uint32_t len = d.length >> kLengthShift;
  uint32_t dv = d.value / len;
  if (dv >= (1u << kLengthShift)) dv = (1u << kLengthShift) - 1;   // only on corrupt input
  uint32_t s = start_[dv >> kTableShift];
  {
    const __m128i v = _mm_set1_epi32((int)dv);                     // all values < 2^31: signed compares are fine
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(dist + s + 1));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(dist + s + 5));
    const unsigned above = (unsigned)_mm_movemask_ps(_mm_castsi128_ps(_mm_cmpgt_epi32(a, v))) |
                           ((unsigned)_mm_movemask_ps(_mm_castsi128_ps(_mm_cmpgt_epi32(b, v))) << 4);
    if (above) {
      s += (uint32_t)__builtin_ctz(above);
    } else {                                                       // more than eight symbols inside one bucket: rare
      s += kSearchWidth;
      while (dist[s + 1] <= dv) ++s;
    }
  }
  const uint32_t x = dist[s] * len;
  const uint32_t y = s + 1 == n_ ? d.length : dist[s + 1] * len;   // last symbol: y = old length
  uint32_t value = d.value - x, length = y - x;
  if (d.end - d.p >= 4 && length >= (1u << 8)) {                   // renorm_dec_interval, 0 - 2 bytes at once
    const unsigned nb = (unsigned)__builtin_clz(length) >> 3;      // length >= 2^24 -> 0, >= 2^16 -> 1, >= 2^8 -> 2
    uint32_t in;
    memcpy(&in, d.p, 4);
    in = __builtin_bswap32(in);
    value = (uint32_t)(((((uint64_t)value) << 32) | in) >> (32 - 8 * nb));
    length <<= 8 * nb;
    d.p += nb;
  } else {
    while (length < kMinLength) {
      value = (value << 8) | d.next();
      length <<= 8;
    }
  }
  d.value = value;
  d.length = length;
  d.ok &= s < 256;
  ++count_[s];
  if (--d.until == 0) {
    update_model();
    d.until = until_;
  }
  return (uint8_t)s;
}

bool RangeDecoder::decode_all(const uint8_t *code, size_t nbytes, uint8_t *sym, size_t n) {
  State d;
  begin(d, code, nbytes);
  for (size_t i = 0; i < n; ++i) sym[i] = step(d);
  return d.ok && d.overrun <= 4;
}

// Several independent streams in one loop: each symbol costs a 32-bit division and a dependent
// table walk; with two to four chains in flight the divider pipelines and most of that latency
// hides.  Streams may have different lengths (the loop keeps going with the ones that are left).
bool RangeDecoder::decode_multi(RangeDecoder *const *dec, const StreamIO *io, int k) {
  State st[kMaxInterleave];
  size_t longest = 0, shortest = (size_t)-1;
  for (int q = 0; q < k; ++q) {
    dec[q]->begin(st[q], io[q].code, io[q].nbytes);
    longest = io[q].n > longest ? io[q].n : longest;
    shortest = io[q].n < shortest ? io[q].n : shortest;
  }
  size_t i = 0;
  auto run = [&](auto kc) {               // the common prefix of all streams, K chains in flight
    constexpr int K = decltype(kc)::value;
    for (; i < shortest; ++i) {
      uint8_t out[K];
      for (int q = 0; q < K; ++q) out[q] = dec[q]->step(st[q]);
      for (int q = 0; q < K; ++q) io[q].sym[i] = out[q];
    }
  };
  if (k == 8) run(std::integral_constant<int, 8>());
  else if (k == 4) run(std::integral_constant<int, 4>());
  else if (k == 3) run(std::integral_constant<int, 3>());
  else if (k == 2) run(std::integral_constant<int, 2>());
  for (; i < longest; ++i)
    for (int q = 0; q < k; ++q)
      if (i < io[q].n) io[q].sym[i] = dec[q]->step(st[q]);
  bool ok = true;
  for (int q = 0; q < k; ++q) ok = ok && st[q].ok && st[q].overrun <= 4;
  return ok;
}

bool RangeDecoder::decode_pair(RangeDecoder &ma, const uint8_t *ca, size_t ba, uint8_t *sa, size_t na,
                               RangeDecoder &mb, const uint8_t *cb, size_t bb, uint8_t *sb, size_t nb) {
  RangeDecoder *d[2] = {&ma, &mb};
  const StreamIO io[2] = {{ca, ba, sa, na}, {cb, bb, sb, nb}};
  return decode_multi(d, io, 2);
}

namespace {

void put_u32(std::vector<uint8_t> &v, uint32_t x) {
  const size_t o = v.size();
  v.resize(o + 4);
  memcpy(v.data() + o, &x, 4);
}

// Runs `n_tasks` independent tasks on up to `threads` host threads.
template <typename F>
void parallel_for(int n_tasks, int threads, F &&fn) {
  if (threads < 1) threads = 1;
  if (threads > n_tasks) threads = n_tasks;
  if (threads <= 1) {
    for (int i = 0; i < n_tasks; ++i) fn(i);
    return;
  }
  std::atomic<int> next(0);
  std::atomic<bool> threw(false);   // an exception must not leave a std::thread (std::terminate)
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&]() {
      try {
        for (int i = next.fetch_add(1); i < n_tasks; i = next.fetch_add(1)) fn(i);
      } catch (...) {
        threw.store(true);
        next.store(n_tasks);
      }
    });
  for (auto &th : pool) th.join();
  if (threw.load()) throw std::bad_alloc();
}

// Body of an extern "C" entry point: no exception crosses the C ABI (allocation failures of the
// std::vector bookkeeping become MPTC_E_NOMEM).
template <typename F>
int guarded(F &&body) {
  try {
    return body();
  } catch (...) {
    return MPTC_E_NOMEM;
  }
}

struct StreamJob {
  const uint8_t *sym;
  size_t n;
  std::vector<uint8_t> out;
};

int copy_out(const std::vector<uint8_t> &bytes, uint8_t *out, size_t cap, size_t *out_bytes) {
  if (out_bytes) *out_bytes = bytes.size();
  if (bytes.size() > cap || !out) return MPTC_E_SPACE;
  memcpy(out, bytes.data(), bytes.size());
  return MPTC_OK;
}

// The five streams of a frame in the order EntropyEncode/CompressEndpoint emit them.
void frame_jobs(const uint8_t *motion, size_t nb, const uint8_t *planes, size_t ps, StreamJob jobs[5]) {
  jobs[0] = {motion, 2 * nb, {}};              // (x, y) interleaved (codec.cpp:1124-1127)
  jobs[1] = {planes + 0 * ps, ps, {}};         // ep1 Y
  jobs[2] = {planes + 1 * ps, 2 * ps, {}};     // ep1 Co | Cg (codec.cpp:843)
  jobs[3] = {planes + 3 * ps, ps, {}};         // ep2 Y
  jobs[4] = {planes + 4 * ps, 2 * ps, {}};     // ep2 Co | Cg (codec.cpp:873)
}

void append_frame_payload(std::vector<uint8_t> &out, uint32_t n_unique, StreamJob jobs[5]) {
  put_u32(out, n_unique);                      // codec.cpp:1140-1142
  for (int s = 0; s < 5; ++s) {
    put_u32(out, (uint32_t)jobs[s].out.size());
    out.insert(out.end(), jobs[s].out.begin(), jobs[s].out.end());
  }
}

// parallel_for whose tasks are handed out strictly in index order (task i may block until its
// input has arrived; later inputs never arrive before earlier ones).
template <typename F>
void parallel_for_ordered(int n_tasks, int threads, F &&fn) { parallel_for(n_tasks, threads, fn); }

// Everything mptc_assemble_stream / mptc_encode_stream share: the independent arithmetic-coder
// jobs of a sequence (five per frame + one palette per group) and the final byte layout.
struct StreamPlan {
  int n_frames, w, h, n_groups, n_used;
  mptc_gpu_params p;
  size_t nb, ps;
  const uint32_t *unique, *n_unique;
  std::vector<StreamJob> jobs;                  // [n_used * 5] frame streams, then [n_groups] palettes
  std::vector<std::vector<uint8_t>> palettes;   // combined palette per group (codec.cpp:1473-1479)

  StreamPlan(int n_frames_, int w_, int h_, const mptc_gpu_params &p_, const uint8_t *motion, const uint32_t *unique_,
             const uint32_t *n_unique_, const uint8_t *planes)
      : n_frames(n_frames_), w(w_), h(h_), p(p_), unique(unique_), n_unique(n_unique_) {
    nb = (size_t)(w / 4) * (h / 4);
    ps = (size_t)((w / 4 + 63) / 64 * 64) * ((h / 4 + 63) / 64 * 64);
    n_groups = n_frames / p.gop;                // a trailing partial group is never written
    n_used = n_groups * p.gop;
    jobs.resize((size_t)n_used * 5 + n_groups);
    palettes.resize(n_groups);
    for (int f = 0; f < n_used; ++f)
      frame_jobs(motion + (size_t)f * 2 * nb, nb, planes + (size_t)f * 6 * ps, ps, &jobs[(size_t)f * 5]);
  }

  void build_palette(int g) {                   // needs n_unique / unique of the group's frames
    std::vector<uint8_t> &pal = palettes[g];
    pal.clear();
    for (int f = g * p.gop; f < (g + 1) * p.gop; ++f) {
      const uint8_t *src = reinterpret_cast<const uint8_t *>(unique + (size_t)f * nb);
      pal.insert(pal.end(), src, src + (size_t)n_unique[f] * 4);
    }
    jobs[(size_t)n_used * 5 + g] = {pal.data(), pal.size(), {}};
  }

  void encode_job(int i) {
    RangeEncoder enc;
    enc.encode_all(jobs[i].sym, jobs[i].n, jobs[i].out);
  }

  // One task = one or two jobs coded by one thread in an interleaved loop (RangeEncoder::encode_pair).
  struct Task { int a, b; };                    // b < 0: single job
  void encode_task(const Task &t) {
    if (t.b < 0) { encode_job(t.a); return; }
    RangeEncoder ea, eb;
    RangeEncoder::encode_pair(ea, jobs[t.a].sym, jobs[t.a].n, jobs[t.a].out, eb, jobs[t.b].sym, jobs[t.b].n, jobs[t.b].out);
  }
  // Tasks in the order their inputs arrive from the GPU (frame k of every group, then k + 1, ...):
  // per frame the two Y planes and the two Co|Cg planes pair up (equal lengths), motion streams
  // pair across two groups; a group's palette follows its last frame.
  std::vector<Task> tasks_in_arrival_order() const {
    std::vector<Task> t;
    t.reserve(jobs.size());
    for (int k = 0; k < p.gop; ++k) {
      int pending_motion = -1;
      for (int g = 0; g < n_groups; ++g) {
        const int f = g * p.gop + k;
        if (pending_motion < 0) pending_motion = f * 5;
        else { t.push_back({pending_motion, f * 5}); pending_motion = -1; }
        t.push_back({f * 5 + 2, f * 5 + 4});
        t.push_back({f * 5 + 1, f * 5 + 3});
        if (k == p.gop - 1) t.push_back({n_used * 5 + g, -1});
      }
      if (pending_motion >= 0) t.push_back({pending_motion, -1});
    }
    return t;
  }
  // Frames a task's inputs come from: [first, last] of the jobs' frames (palette: the whole group).
  void task_frames(const Task &t, int frames_out[2], int &n) const {
    n = 0;
    for (int j : {t.a, t.b}) {
      if (j < 0) continue;
      if (j < n_used * 5) frames_out[n++] = j / 5;
    }
  }

  // Lays the stream out directly in the caller's buffer (no intermediate copy of the whole stream):
  // the record offsets follow from the code sizes, the copies run on `threads` host threads.
  int write(uint8_t *out, size_t cap, size_t *out_bytes, mptc_stream_stats &st, int threads) {
    memset(&st, 0, sizeof st);
    st.n_groups = (uint32_t)n_groups;
    std::vector<size_t> at(jobs.size());             // where each job's u32 size prefix goes
    std::vector<size_t> group_at(n_groups), frame_at((size_t)n_used);
    size_t total = 34;
    for (int g = 0; g < n_groups; ++g) {
      const size_t pj = (size_t)n_used * 5 + g;
      group_at[g] = total;
      at[pj] = total;
      total += 4 + jobs[pj].out.size() + 4;          // palette record + u32 unique bytes
      for (int f = g * p.gop; f < (g + 1) * p.gop; ++f) {
        frame_at[f] = total;
        total += 4;                                  // u32 n_unique (codec.cpp:1140-1142)
        for (int q = 0; q < 5; ++q) {
          at[(size_t)f * 5 + q] = total;
          total += 4 + jobs[(size_t)f * 5 + q].out.size();
        }
      }
    }
    if (out_bytes) *out_bytes = total;
    if (!out || total > cap) return MPTC_E_SPACE;
    auto put = [&](size_t off, uint32_t x) { memcpy(out + off, &x, 4); };
    put(0, (uint32_t)h);                             // header (codec.cpp:1358-1367)
    put(4, (uint32_t)w);
    out[8] = (uint8_t)p.gop;
    out[9] = (uint8_t)p.search_area;
    put(10, (uint32_t)n_groups);
    for (int g = 0; g < n_groups; ++g) {
      const std::vector<uint8_t> &cpal = jobs[(size_t)n_used * 5 + g].out;
      put(group_at[g] + 4 + cpal.size(), (uint32_t)palettes[g].size());
      if ((uint32_t)cpal.size() > st.max_comp_palette) st.max_comp_palette = (uint32_t)cpal.size();
      if ((uint32_t)palettes[g].size() > st.max_unique_bytes) st.max_unique_bytes = (uint32_t)palettes[g].size();
      for (int f = g * p.gop; f < (g + 1) * p.gop; ++f) {
        const StreamJob *fj = &jobs[(size_t)f * 5];
        put(frame_at[f], n_unique[f]);
        const uint32_t m = (uint32_t)fj[0].out.size();
        if (m > st.max_comp_motion) st.max_comp_motion = m;
        for (int q : {1, 3}) if ((uint32_t)fj[q].out.size() > st.max_comp_ep_y) st.max_comp_ep_y = (uint32_t)fj[q].out.size();
        for (int q : {2, 4}) if ((uint32_t)fj[q].out.size() > st.max_comp_ep_c) st.max_comp_ep_c = (uint32_t)fj[q].out.size();
      }
    }
    const uint32_t patch[5] = {st.max_unique_bytes, st.max_comp_palette, st.max_comp_motion, st.max_comp_ep_y,
                               st.max_comp_ep_c};   // codec.cpp:1514-1520
    memcpy(out + 14, patch, sizeof patch);
    parallel_for((int)jobs.size(), threads, [&](int j) {
      const std::vector<uint8_t> &v = jobs[j].out;
      put(at[j], (uint32_t)v.size());
      if (!v.empty()) memcpy(out + at[j] + 4, v.data(), v.size());
    });
    return MPTC_OK;
  }
};

}  // namespace
}  // namespace mptc

using namespace mptc;

extern "C" {

int mptc_arith_encode(const uint8_t *sym, size_t n, uint8_t *out, size_t cap, size_t *out_bytes) {
  return guarded([&]() -> int {
  if (!sym && n) return MPTC_E_ARG;
  std::vector<uint8_t> bytes;
  RangeEncoder enc;
  enc.encode_all(sym, n, bytes);
  return copy_out(bytes, out, cap, out_bytes);
  });
}

int mptc_frame_payload(const uint8_t *motion, size_t nb, const uint8_t *planes, size_t plane_syms,
                       uint32_t n_unique, int threads, uint8_t *out, size_t cap, size_t *out_bytes,
                       uint32_t *sizes) {
  return guarded([&]() -> int {
  if (!motion || !planes) return MPTC_E_ARG;
  StreamJob jobs[5];
  frame_jobs(motion, nb, planes, plane_syms, jobs);
  parallel_for(3, threads, [&](int t) {   // motion alone; the equal-length plane pairs interleaved
    if (t == 0) {
      RangeEncoder enc;
      enc.encode_all(jobs[0].sym, jobs[0].n, jobs[0].out);
    } else {
      RangeEncoder ea, eb;
      RangeEncoder::encode_pair(ea, jobs[t].sym, jobs[t].n, jobs[t].out, eb, jobs[t + 2].sym, jobs[t + 2].n, jobs[t + 2].out);
    }
  });
  std::vector<uint8_t> bytes;
  append_frame_payload(bytes, n_unique, jobs);
  if (sizes)
    for (int s = 0; s < 5; ++s) sizes[s] = (uint32_t)jobs[s].out.size();
  return copy_out(bytes, out, cap, out_bytes);
  });
}

int mptc_assemble_stream(int n_frames, int w, int h, const mptc_gpu_params *p, const uint8_t *motion,
                         const uint32_t *unique, const uint32_t *n_unique, const uint8_t *planes,
                         int threads, uint8_t *out, size_t cap, size_t *out_bytes,
                         mptc_stream_stats *stats) {
  return guarded([&]() -> int {
  if (!p || !motion || !unique || !n_unique || !planes || n_frames < 1) return MPTC_E_ARG;
  if (p->gop < 1 || p->gop > 255 || p->search_area < 1 || p->search_area > 63) return MPTC_E_ARG;
  const auto t0 = std::chrono::steady_clock::now();
  StreamPlan plan(n_frames, w, h, *p, motion, unique, n_unique, planes);
  for (int g = 0; g < plan.n_groups; ++g) plan.build_palette(g);
  const std::vector<StreamPlan::Task> tasks = plan.tasks_in_arrival_order();
  parallel_for((int)tasks.size(), threads, [&](int i) { plan.encode_task(tasks[i]); });
  mptc_stream_stats st;
  const int wr = plan.write(out, cap, out_bytes, st, threads);
  st.entropy_ms = st.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (stats) {
    st.gpu_ms = stats->gpu_ms;
    *stats = st;
  }
  return wr;
  });
}

// GPU hot path and host arithmetic coding OVERLAPPED (BASELINE.json north_star): the encode is
// enqueued asynchronously, results come back frame by frame in pinned buffers, and a pool of
// host threads codes the five streams of every frame as soon as that frame's D2H copy has
// completed (mptc_gpu_wait_frame), while the GPU is still searching the later frames.
int mptc_encode_stream(mptc_gpu_ctx *ctx, const uint8_t *frames, int n_frames, int w, int h,
                       const mptc_gpu_params *p, int threads, uint8_t *out, size_t cap,
                       size_t *out_bytes, mptc_stream_stats *stats) {
  return guarded([&]() -> int {
  if (!ctx || !frames || !p || n_frames < 1) return MPTC_E_ARG;
  if (w < 4 || h < 4 || (w & 3) || (h & 3)) return MPTC_E_ARG;
  if (p->gop < 1 || p->gop > 255 || p->search_area < 1 || p->search_area > 63) return MPTC_E_ARG;
  const auto t0 = std::chrono::steady_clock::now();
  const size_t nb = (size_t)(w / 4) * (h / 4);
  const size_t ps = (size_t)((w / 4 + 63) / 64 * 64) * ((h / 4 + 63) / 64 * 64);
  const size_t n = (size_t)n_frames;
  // pinned staging (owned by the context, kept across calls) for the results the stream needs;
  // the 8-byte blocks are not part of the stream
  uint8_t *motion = static_cast<uint8_t *>(ctx_pinned(ctx, 0, n * nb * 2));
  uint32_t *unique = static_cast<uint32_t *>(ctx_pinned(ctx, 1, n * nb * 4));
  uint32_t *n_unique = static_cast<uint32_t *>(ctx_pinned(ctx, 2, n * 4));
  uint8_t *planes = static_cast<uint8_t *>(ctx_pinned(ctx, 3, n * 6 * ps));
  int r = MPTC_E_NOMEM;
  if (motion && unique && n_unique && planes) {
    r = mptc_gpu_encode_sequence_async(ctx, frames, n_frames, w, h, p, nullptr, motion, unique, n_unique, planes);
    if (r == MPTC_OK) {
      StreamPlan plan(n_frames, w, h, *p, motion, unique, n_unique, planes);
      const std::vector<StreamPlan::Task> tasks = plan.tasks_in_arrival_order();
      std::atomic<int> failed(MPTC_OK);
      const auto t1 = std::chrono::steady_clock::now();
      parallel_for_ordered((int)tasks.size(), threads, [&](int i) {
        const StreamPlan::Task &t = tasks[i];
        int wr = MPTC_OK;
        if (t.a >= plan.n_used * 5) {          // palette: needs every frame of its group
          const int g = t.a - plan.n_used * 5;
          for (int f = g * p->gop; f < (g + 1) * p->gop && wr == MPTC_OK; ++f) wr = mptc_gpu_wait_frame(ctx, f);
          if (wr == MPTC_OK) plan.build_palette(g);
        } else {
          int fr[2], nf;
          plan.task_frames(t, fr, nf);
          for (int q = 0; q < nf && wr == MPTC_OK; ++q) wr = mptc_gpu_wait_frame(ctx, fr[q]);
        }
        if (wr != MPTC_OK) { failed.store(wr); return; }
        plan.encode_task(t);
      });
      r = failed.load();
      const int wr = mptc_gpu_wait(ctx);   // also frames beyond the last full group
      if (r == MPTC_OK) r = wr;
      if (r == MPTC_OK) {
        mptc_stream_stats st;
        const auto t15 = std::chrono::steady_clock::now();
        r = plan.write(out, cap, out_bytes, st, threads);
        const auto t2 = std::chrono::steady_clock::now();
        st.assemble_ms = std::chrono::duration<double, std::milli>(t2 - t15).count();
        float ms = 0.f;
        if (mptc_gpu_last_encode_ms(ctx, 0, &ms) == MPTC_OK) st.gpu_ms = ms;
        st.entropy_ms = std::chrono::duration<double, std::milli>(t15 - t1).count();
        st.total_ms = std::chrono::duration<double, std::milli>(t2 - t0).count();
        if (stats) *stats = st;
      }
    }
  }
  return r;
  });
}

// ---- decoder side ------------------------------------------------------------------------------

int mptc_arith_decode(const uint8_t *code, size_t nbytes, uint8_t *sym, size_t n) {
  return guarded([&]() -> int {
  if ((!code && nbytes) || (!sym && n)) return MPTC_E_ARG;
  RangeDecoder dec;
  return dec.decode_all(code, nbytes, sym, n) ? MPTC_OK : MPTC_E_DATA;
  });
}

int mptc_arith_decode_multi(int k, const uint8_t *const *code, const size_t *nbytes, uint8_t *const *sym, const size_t *n) {
  return guarded([&]() -> int {
  if (k < 1 || k > RangeDecoder::kMaxInterleave || !code || !nbytes || !sym || !n) return MPTC_E_ARG;
  for (int q = 0; q < k; ++q)
    if ((!code[q] && nbytes[q]) || (!sym[q] && n[q])) return MPTC_E_ARG;
  RangeDecoder pool[RangeDecoder::kMaxInterleave];
  RangeDecoder *d[RangeDecoder::kMaxInterleave];
  RangeDecoder::StreamIO io[RangeDecoder::kMaxInterleave];
  for (int q = 0; q < k; ++q) {
    d[q] = &pool[q];
    io[q] = RangeDecoder::StreamIO{code[q], nbytes[q], sym[q], n[q]};
  }
  const bool ok = k == 1 ? d[0]->decode_all(io[0].code, io[0].nbytes, io[0].sym, io[0].n) : RangeDecoder::decode_multi(d, io, k);
  return ok ? MPTC_OK : MPTC_E_DATA;
  });
}

int mptc_stream_info(const uint8_t *stream, size_t bytes, mptc_stream_header *hdr) {
  if (!stream || !hdr) return MPTC_E_ARG;
  if (bytes < 34) return MPTC_E_DATA;
  uint32_t v[2], g, mx[5];
  memcpy(v, stream, 8);                 // reader: codec.cpp:1172-1184
  memcpy(&g, stream + 10, 4);
  memcpy(mx, stream + 14, 20);
  hdr->height = (int)v[0]; hdr->width = (int)v[1];
  hdr->gop = stream[8]; hdr->search_area = stream[9];
  hdr->n_groups = (int)g;
  hdr->n_frames = 0;
  hdr->max_unique_bytes = mx[0]; hdr->max_comp_palette = mx[1]; hdr->max_comp_motion = mx[2];
  hdr->max_comp_ep_y = mx[3]; hdr->max_comp_ep_c = mx[4];
  if (hdr->width < 4 || hdr->height < 4 || (hdr->width & 3) || (hdr->height & 3) || hdr->width > 65536 ||
      hdr->height > 65536 || hdr->gop < 1 || hdr->search_area < 1 || hdr->search_area > 63 || g < 1 || g > (1u << 24))
    return MPTC_E_DATA;
  // every group holds at least two u32 (palette size, unique bytes) and every frame six (n_unique +
  // five record sizes): a header that promises more than the stream can hold is corrupt, and it is
  // rejected before anything is sized by it
  const uint64_t n_frames = (uint64_t)g * (uint64_t)hdr->gop;
  if (n_frames > (1u << 24) || (uint64_t)bytes < 34 + (uint64_t)g * 8 + n_frames * 24) return MPTC_E_DATA;
  hdr->n_frames = (int)n_frames;
  return MPTC_OK;
}

// DecompressMultiUnique (codec.cpp:1161-1305) with the work re-cut for the machine: the stream is
// walked once to find its records (sizes only), every record is an independent arithmetic-decoder
// job for the host thread pool (a group's palette + five per frame, written straight into pinned
// staging in the layout the kernels read), and as soon as all jobs of a group are done that
// group's symbols go to the GPU and its frames are reconstructed there -- all frames of the group
// in the same launches -- while the pool is already decoding the next groups.
int mptc_decode_stream(mptc_gpu_ctx *ctx, const uint8_t *stream, size_t bytes, int threads, uint64_t *blocks_out,
                       uint8_t *rgb_out, mptc_decode_stats *stats) {
  return guarded([&]() -> int {
  if (!ctx || !stream) return MPTC_E_ARG;
  const auto t0 = std::chrono::steady_clock::now();
  mptc_stream_header H;
  if (int r = mptc_stream_info(stream, bytes, &H)) return r;
  const int gop = H.gop, n_frames = H.n_frames, w = H.width, h = H.height;
  const size_t nb = (size_t)(w / 4) * (h / 4);
  const size_t ps = (size_t)((w / 4 + 63) / 64 * 64) * ((h / 4 + 63) / 64 * 64);
  struct Rec { const uint8_t *code; size_t nbytes; uint8_t *dst; size_t n; int group; };
  std::vector<Rec> recs;
  recs.reserve((size_t)n_frames * 5 + H.n_groups);
  std::vector<size_t> pal_off(H.n_groups + 1, 0);   // byte offset of every group's palette in `pal`
  // pass 1: locate the records
  size_t off = 34, pal_total = 0;
  auto get_u32 = [&](uint32_t &x) {
    if (off + 4 > bytes) return false;
    memcpy(&x, stream + off, 4);
    off += 4;
    return true;
  };
  struct GroupHdr { size_t pal_code_off; uint32_t pal_code_bytes, unique_bytes; };
  std::vector<GroupHdr> groups(H.n_groups);
  std::vector<uint32_t> n_unique_v((size_t)n_frames);
  struct FrameRec { size_t off[5]; uint32_t nbytes[5]; };
  std::vector<FrameRec> frs((size_t)n_frames);
  for (int g = 0; g < H.n_groups; ++g) {
    uint32_t cpal, ub;
    if (!get_u32(cpal) || off + cpal > bytes) return MPTC_E_DATA;
    groups[g].pal_code_off = off; groups[g].pal_code_bytes = cpal;
    off += cpal;
    if (!get_u32(ub) || (ub & 3) || ub > (uint64_t)gop * nb * 4) return MPTC_E_DATA;
    groups[g].unique_bytes = ub;
    pal_off[g] = pal_total;
    pal_total += ub;
    uint64_t group_unique = 0;
    for (int k = 0; k < gop; ++k) {
      const int f = g * gop + k;
      if (!get_u32(n_unique_v[f]) || n_unique_v[f] > nb) return MPTC_E_DATA;
      group_unique += n_unique_v[f];
      for (int q = 0; q < 5; ++q) {
        uint32_t sz;
        if (!get_u32(sz) || off + sz > bytes) return MPTC_E_DATA;
        frs[f].off[q] = off; frs[f].nbytes[q] = sz;
        off += sz;
      }
    }
    if (group_unique * 4 != ub) return MPTC_E_DATA;
  }
  pal_off[H.n_groups] = pal_total;
  // pinned staging in the layout the kernels read
  const size_t n = (size_t)n_frames;
  uint8_t *motion = static_cast<uint8_t *>(ctx_pinned(ctx, 0, n * nb * 2));
  uint8_t *pal = static_cast<uint8_t *>(ctx_pinned(ctx, 1, pal_total + 4));
  uint8_t *planes = static_cast<uint8_t *>(ctx_pinned(ctx, 3, n * 6 * ps));
  if (!motion || !pal || !planes) return MPTC_E_NOMEM;
  if (int r = mptc_gpu_seq_reserve(ctx, w, h, n_frames)) return r;
  for (int g = 0; g < H.n_groups; ++g) {
    recs.push_back({stream + groups[g].pal_code_off, groups[g].pal_code_bytes, pal + pal_off[g], groups[g].unique_bytes, g});
    for (int k = 0; k < gop; ++k) {
      const size_t f = (size_t)g * gop + k;
      uint8_t *pl = planes + f * 6 * ps;
      uint8_t *dst[5] = {motion + f * 2 * nb, pl, pl + ps, pl + 3 * ps, pl + 4 * ps};   // Y1, Co|Cg 1, Y2, Co|Cg 2
      const size_t cnt[5] = {2 * nb, ps, 2 * ps, ps, 2 * ps};
      for (int q = 0; q < 5; ++q) recs.push_back({stream + frs[f].off[q], frs[f].nbytes[q], dst[q], cnt[q], g});
    }
  }
  std::vector<std::atomic<int>> left(H.n_groups);
  for (int g = 0; g < H.n_groups; ++g) left[g].store(1 + 5 * gop);
  // tasks = records decoded by one thread in one interleaved loop (RangeDecoder::decode_multi), streams of equal
  // length together: the motion streams of four frames of a group, the Y planes of two frames, the Co|Cg
  // planes of two frames, a palette alone.  (With the branch-free decoder step four chains in flight are 30 %
  // faster than two -- 148 against 114 Msym/s per thread on the GPU box's Xeon, profiles/micro/decoder_bench.py;
  // the branchy step did not gain beyond two.)  Long tasks first: the threads take tasks in order.
  struct Task { int r[RangeDecoder::kMaxInterleave]; int k; };
  std::vector<Task> tasks;
  tasks.reserve(recs.size());
  auto add_task = [&](std::initializer_list<int> rs) {
    Task t;
    t.k = 0;
    for (int r : rs) t.r[t.k++] = r;
    for (int q = t.k; q < RangeDecoder::kMaxInterleave; ++q) t.r[q] = -1;
    tasks.push_back(t);
  };
  for (int g = 0; g < H.n_groups; ++g) {
    const int r0 = g * (1 + 5 * gop);
    auto fr = [&](int k) { return r0 + 1 + 5 * k; };           // records of frame k: motion, Y1, Co|Cg 1, Y2, Co|Cg 2
    add_task({r0});                                             // palette
    for (int k = 0; k < gop; k += 4) {                          // motion of frames k .. k+3
      const int m = gop - k < 4 ? gop - k : 4;
      if (m == 4) add_task({fr(k), fr(k + 1), fr(k + 2), fr(k + 3)});
      else if (m == 3) add_task({fr(k), fr(k + 1), fr(k + 2)});
      else if (m == 2) add_task({fr(k), fr(k + 1)});
      else add_task({fr(k)});
    }
    for (int k = 0; k < gop; k += 2) {
      if (k + 1 < gop) {
        add_task({fr(k) + 2, fr(k) + 4, fr(k + 1) + 2, fr(k + 1) + 4});   // Co|Cg 1, 2 of two frames (equal lengths)
        add_task({fr(k) + 1, fr(k) + 3, fr(k + 1) + 1, fr(k + 1) + 3});   // Y 1, 2 of two frames
      } else {
        add_task({fr(k) + 2, fr(k) + 4});
        add_task({fr(k) + 1, fr(k) + 3});
      }
    }
  }
  std::stable_sort(tasks.begin(), tasks.end(), [&](const Task &a, const Task &b) {
    size_t na = 0, nb_ = 0;
    for (int q = 0; q < a.k; ++q) na += recs[a.r[q]].n;
    for (int q = 0; q < b.k; ++q) nb_ += recs[b.r[q]].n;
    return na > nb_;
  });
  std::atomic<int> next(0), corrupt(0);
  const int n_tasks = (int)tasks.size();
  auto worker = [&]() {
    try {
      RangeDecoder pool[RangeDecoder::kMaxInterleave];
      RangeDecoder *d[RangeDecoder::kMaxInterleave];
      for (int q = 0; q < RangeDecoder::kMaxInterleave; ++q) d[q] = &pool[q];
      for (int i = next.fetch_add(1); i < n_tasks; i = next.fetch_add(1)) {
        const Task &t = tasks[i];
        RangeDecoder::StreamIO io[RangeDecoder::kMaxInterleave];
        for (int q = 0; q < t.k; ++q) io[q] = {recs[t.r[q]].code, recs[t.r[q]].nbytes, recs[t.r[q]].dst, recs[t.r[q]].n};
        const bool ok = t.k == 1 ? d[0]->decode_all(io[0].code, io[0].nbytes, io[0].sym, io[0].n)
                                 : RangeDecoder::decode_multi(d, io, t.k);
        if (!ok) corrupt.store(1);
        left[recs[t.r[0]].group].fetch_sub(t.k, std::memory_order_release);
      }
    } catch (...) {               // an exception must not leave a std::thread; release the feeder loop
      corrupt.store(2);
      for (int g = 0; g < H.n_groups; ++g) left[g].store(0, std::memory_order_release);
    }
  };
  if (threads < 1) threads = 1;
  if (threads > n_tasks) threads = n_tasks;
  const auto t1 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
  // the calling thread feeds the GPU group by group
  int r = MPTC_OK;
  for (int g = 0; g < H.n_groups && r == MPTC_OK; ++g) {
    while (left[g].load(std::memory_order_acquire) > 0) std::this_thread::yield();
    if (corrupt.load()) break;
    const size_t f0 = (size_t)g * gop;
    r = mptc_gpu_seq_decode_upload(ctx, (int)f0, gop, motion + f0 * 2 * nb, reinterpret_cast<const uint32_t *>(pal + pal_off[g]),
                                   n_unique_v.data() + f0, 0, planes + f0 * 6 * ps);
    if (r == MPTC_OK) r = mptc_gpu_seq_decode(ctx, (int)f0, gop, H.search_area, gop, rgb_out != nullptr);
  }
  for (auto &th : pool) th.join();
  const auto t2 = std::chrono::steady_clock::now();
  if (r == MPTC_OK && corrupt.load()) r = corrupt.load() == 2 ? MPTC_E_NOMEM : MPTC_E_DATA;
  if (r == MPTC_OK) r = mptc_gpu_seq_decode_download(ctx, 0, n_frames, blocks_out, rgb_out);
  const auto t3 = std::chrono::steady_clock::now();
  if (stats) {
    stats->header = H;
    stats->entropy_ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
    stats->total_ms = std::chrono::duration<double, std::milli>(t3 - t0).count();
    stats->symbols = (uint64_t)n * (2 * nb + 6 * ps) + pal_total;
  }
  return r;
  });
}

}  // extern "C"
