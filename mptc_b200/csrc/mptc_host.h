// mptc_host.h -- host-side codec pieces (arithmetic coder, stream assembly).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace mptc {

// Raw scratch buffers that outlive a call (mptc_host.cpp): the coder's worst-case output bound is
// far above what it produces, and fresh zero-filled vectors of that size cost more than coding.
struct Scratch {
  uint8_t *p = nullptr;
  size_t cap = 0;
};
Scratch scratch_acquire(size_t bytes);
void scratch_release(Scratch s);
// An adaptive-model symbol never costs more than 15 bits; +16 for the coder's look-ahead stores.
inline size_t encode_bound(size_t n) { return 2 * n + 16; }

// Adaptive multi-symbol model + 32-bit range coder, bit-compatible with the reference's
// entropy::Adaptive_Data_Model / entropy::Arithmetic_Codec (entropy/arithmetic_codec.cpp).
class RangeEncoder {
 public:
  explicit RangeEncoder(unsigned symbols = 257);
  // Appends the code bytes of `sym[0..n)` (fresh model, start .. stop) to `out`.
  void encode_all(const uint8_t *sym, size_t n, std::vector<uint8_t> &out);
  // The same into a raw buffer of at least encode_bound(n) bytes; returns the code length.
  size_t encode_raw(const uint8_t *sym, size_t n, uint8_t *buf);
  static void encode_pair_raw(RangeEncoder &ma, const uint8_t *sa, size_t na, uint8_t *bufa, size_t *lena,
                              RangeEncoder &mb, const uint8_t *sb, size_t nb, uint8_t *bufb, size_t *lenb);
  // Two independent streams coded in one interleaved loop (each with its own model / encoder).
  static void encode_pair(RangeEncoder &ma, const uint8_t *sa, size_t na, std::vector<uint8_t> &oa,
                          RangeEncoder &mb, const uint8_t *sb, size_t nb, std::vector<uint8_t> &ob);

 private:
  struct State;
  void begin(State &e, uint8_t *buf);
  void step(State &e, uint32_t s);
  size_t finish(State &e);
  void reset_model();
  void update_model();
  unsigned n_;
  std::vector<uint32_t> dist_, count_;
  uint32_t total_ = 0, cycle_ = 0, until_ = 0;
};

// The decoding side of the same coder (Arithmetic_Codec::decode, arithmetic_codec.cpp:391-444,
// start_decoder :511-522, renorm_dec_interval :100-105) with the decoder flavour of the adaptive
// model.  The symbol search is a coarse table plus a short linear scan instead of the reference's
// table plus bisection: any search for the s with dist[s] <= value/length < dist[s+1] decodes the
// same symbol.
class RangeDecoder {
 public:
  explicit RangeDecoder(unsigned symbols = 257);
  // Decodes n byte symbols from code[0..nbytes).  Reads past the end of the code see zeros (a
  // valid stream never needs more than the coder's 4-byte look-ahead).  Returns false if the code
  // ran out more than 4 bytes early or a symbol outside 0..255 appeared (corrupt stream).
  bool decode_all(const uint8_t *code, size_t nbytes, uint8_t *sym, size_t n);
  // Up to kMaxInterleave independent streams decoded in one interleaved loop (each with its own
  // model / decoder).
  static constexpr int kMaxInterleave = 8;
  struct StreamIO { const uint8_t *code; size_t nbytes; uint8_t *sym; size_t n; };
  static bool decode_multi(RangeDecoder *const *dec, const StreamIO *io, int k);
  static bool decode_pair(RangeDecoder &ma, const uint8_t *ca, size_t ba, uint8_t *sa, size_t na,
                          RangeDecoder &mb, const uint8_t *cb, size_t bb, uint8_t *sb, size_t nb);

 private:
  struct State;
  void begin(State &d, const uint8_t *code, size_t nbytes);
  uint8_t step(State &d);
  void reset_model();
  void update_model();
  unsigned n_;
  std::vector<uint32_t> dist_, count_;
  std::vector<uint16_t> start_;   // start_[t] = largest s with dist[s] <= t << kTableShift
  uint32_t total_ = 0, cycle_ = 0, until_ = 0;
};

}  // namespace mptc

struct mptc_gpu_ctx;
namespace mptc {
// Page-locked staging buffer owned by the context (grow-only, freed with it): slot 0..3.
// Internal to the library (not part of the C ABI): mptc_encode_stream keeps its result staging
// across calls instead of pinning ~2 MB per frame every time.
void *ctx_pinned(mptc_gpu_ctx *ctx, int slot, size_t bytes);
}  // namespace mptc
