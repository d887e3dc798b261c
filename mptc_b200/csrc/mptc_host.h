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

}  // namespace mptc

struct mptc_gpu_ctx;
namespace mptc {
// Page-locked staging buffer owned by the context (grow-only, freed with it): slot 0..3.
// Internal to the library (not part of the C ABI): mptc_encode_stream keeps its result staging
// across calls instead of pinning ~2 MB per frame every time.
void *ctx_pinned(mptc_gpu_ctx *ctx, int slot, size_t bytes);
}  // namespace mptc
