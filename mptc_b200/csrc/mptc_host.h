// mptc_host.h -- host-side codec pieces (arithmetic coder, stream assembly).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace mptc {

// Adaptive multi-symbol model + 32-bit range coder, bit-compatible with the reference's
// entropy::Adaptive_Data_Model / entropy::Arithmetic_Codec (entropy/arithmetic_codec.cpp).
class RangeEncoder {
 public:
  explicit RangeEncoder(unsigned symbols = 257);
  // Appends the code bytes of `sym[0..n)` (fresh model, start .. stop) to `out`.
  void encode_all(const uint8_t *sym, size_t n, std::vector<uint8_t> &out);
  // Two independent streams coded in one interleaved loop (each with its own model / encoder).
  static void encode_pair(RangeEncoder &ma, const uint8_t *sa, size_t na, std::vector<uint8_t> &oa,
                          RangeEncoder &mb, const uint8_t *sb, size_t nb, std::vector<uint8_t> &ob);

 private:
  struct State;
  void begin(State &e, size_t n, std::vector<uint8_t> &out, size_t &start);
  void step(State &e, uint32_t s);
  void finish(State &e, std::vector<uint8_t> &out, size_t start);
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
  // Two independent streams decoded in one interleaved loop (each with its own model / decoder).
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
