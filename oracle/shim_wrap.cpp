// oracle/shim_wrap.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles the reference-side shim (integration/dxt_image_gpu.h) against the UNMODIFIED reference
// headers and sources and drives it the way CompressMultiUnique drives DXTImage
// (codec/codec.cpp:1383-1509): per frame MakeReencodedFrame (GPU, through libmptc_b200.so's C ABI)
// instead of `new DXTImage(file, ...)` + `Reencode(prev, -1)`, then the reference's OWN
// EntropyEncode(curr_frame, bytes) (codec.cpp:1115-1158) and Get8BitPalette() on the GPU-filled
// object.  Built by `make -C oracle shim` into oracle/_ref/libmptc_shim.so.
#include "dxt_image_gpu.h"

#include "codec.h"

#include <cstring>
#include <iostream>

extern "C" {

// frames: n x h x w x 3 RGB8.  For every frame the payload the reference's EntropyEncode appends is
// written to payload_out back to back; payload_sizes[i] receives its size, palette_out / palette_sizes
// the frame's Get8BitPalette() bytes.  Returns 0, -1 if a buffer is too small, -2 on a GPU error.
int mptc_shim_encode_frames(int device, const uint8_t *frames, int n, int w, int h, int search_area, int err_threshold,
                            int gop, uint8_t *payload_out, size_t payload_cap, uint32_t *payload_sizes,
                            uint8_t *palette_out, size_t palette_cap, uint32_t *palette_sizes, uint64_t *blocks_out) {
  std::cout.setstate(std::ios_base::failbit);   // the reference prints per-frame chatter (codec.cpp:846-848)
  try {
    MPTC::GpuSession gpu(device);
    std::unique_ptr<MPTC::DXTImage> prev, curr;
    size_t pay_off = 0, pal_off = 0;
    const size_t frame_bytes = (size_t)w * h * 3;
    for (int i = 0; i < n; ++i) {
      const bool set_intra = (i % gop) == 0;   // intra_interval == unique_interval == gop (codec.cpp:1462-1469)
      curr = MPTC::MakeReencodedFrame(gpu, w, h, frames + frame_bytes * i, set_intra, search_area, err_threshold, prev);
      std::vector<uint8_t> bytes;
      MPTC::EntropyEncode(curr, bytes);                         // the reference's own packaging + coder
      if (pay_off + bytes.size() > payload_cap) return -1;
      memcpy(payload_out + pay_off, bytes.data(), bytes.size());
      payload_sizes[i] = (uint32_t)bytes.size();
      pay_off += bytes.size();
      const std::vector<uint8_t> pal = curr->Get8BitPalette();   // dxt_image.h:117-122
      if (pal_off + pal.size() > palette_cap) return -1;
      if (!pal.empty()) memcpy(palette_out + pal_off, pal.data(), pal.size());
      palette_sizes[i] = (uint32_t)pal.size();
      pal_off += pal.size();
      if (blocks_out)
        memcpy(blocks_out + (size_t)i * curr->_physical_blocks.size(), curr->_physical_blocks.data(),
               curr->_physical_blocks.size() * 8);
      prev = std::move(curr);
    }
  } catch (const std::exception &e) {
    std::cerr << "mptc_shim_encode_frames: " << e.what() << std::endl;
    return -2;
  }
  return 0;
}

}  // extern "C"
