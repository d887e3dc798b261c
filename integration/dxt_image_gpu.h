// integration/dxt_image_gpu.h -- the reference-side shim: a DXTImage whose hot path ran on the GPU.
//
// This is the file a maintainer of the reference adds to codec/ (INTEGRATION.md section 2).  It is
// written against the reference's own headers (codec/dxt_image.h:46-198) and the C ABI of
// include/mptc_gpu.h; link with -lmptc_b200.  `oracle/Makefile` (target `shim`) compiles it together
// with the UNMODIFIED reference sources, and tests/test_gpu_shim.py feeds the DXTImage it fills to the
// reference's own EntropyEncode (codec/codec.cpp:1115-1158) and compares the payload bytes with the
// golden ones of the CPU path.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <tuple>
#include <vector>

#include "dxt_image.h"
#include "mptc_gpu.h"

namespace MPTC {

class GpuSession {   // one per encoding thread (one context per (host thread, GPU), mptc_gpu.h)
 public:
  explicit GpuSession(int device = 0) {
    if (mptc_gpu_create(device, &ctx_) != MPTC_OK) throw std::runtime_error("mptc_gpu_create failed: no usable GPU");
  }
  ~GpuSession() { mptc_gpu_destroy(ctx_); }
  GpuSession(const GpuSession &) = delete;
  GpuSession &operator=(const GpuSession &) = delete;
  mptc_gpu_ctx *ctx() const { return ctx_; }

 private:
  mptc_gpu_ctx *ctx_ = nullptr;
};

// Replaces   curr_frame.reset(new DXTImage(file, set_intra, search_area, thr));   codec.cpp:1388
//            curr_frame->Reencode(prev_frame, -1);                                 codec.cpp:1394
// Fills the members codec.cpp reads afterwards: _physical_blocks, _motion_indices, _unique_palette,
// _width/_height/_blocks_*/_num_blocks, _search_area and (via SetLogicalBlocks) _logical_blocks.
inline std::unique_ptr<DXTImage> MakeReencodedFrame(GpuSession &gpu, int w, int h, const uint8_t *rgb, bool is_intra,
                                                    int search_area, int32_t err_threshold,
                                                    const std::unique_ptr<DXTImage> &prev) {
  std::unique_ptr<DXTImage> img(new DXTImage(w, h, is_intra, std::vector<uint32_t>()));   // dxt_image.cpp:437
  const size_t nb = img->_physical_blocks.size();
  std::vector<uint8_t> motion(2 * nb);
  std::vector<uint32_t> unique(nb);
  uint32_t n_unique = 0;
  static_assert(sizeof(PhysicalDXTBlock) == 8, "PhysicalDXTBlock is the 8-byte DXT1 block (dxt_image.h:16-23)");
  const uint64_t *prev_blocks =
      (is_intra || !prev) ? nullptr : reinterpret_cast<const uint64_t *>(prev->_physical_blocks.data());
  const int rc = mptc_gpu_reencode(gpu.ctx(), rgb, w, h, is_intra ? 1 : 0, search_area, err_threshold, prev_blocks,
                                   /*initial_out=*/nullptr, reinterpret_cast<uint64_t *>(img->_physical_blocks.data()),
                                   motion.data(), unique.data(), &n_unique);
  if (rc != MPTC_OK) throw std::runtime_error(mptc_gpu_last_error(gpu.ctx()));
  img->_search_area = search_area;
  img->_motion_indices.resize(nb);
  for (size_t b = 0; b < nb; ++b) img->_motion_indices[b] = std::make_tuple(motion[2 * b], motion[2 * b + 1]);
  img->_unique_palette.assign(unique.begin(), unique.begin() + n_unique);
  img->SetLogicalBlocks();   // dxt_image.cpp:452
  return img;
}

}  // namespace MPTC
