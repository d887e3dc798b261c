// oracle/ref_wrap.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin C-ABI wrapper that lets the tests / the cpu_baseline leg of bench.py call the
// UNMODIFIED reference encoder (compiled where it lies under /root/reference by
// oracle/Makefile into oracle/_ref/libmptc_ref.so).  Nothing under mptc_b200/ may
// link, import or execute this.
//
// The reference's only frame entry point loads a PNG (dxt_image.cpp:385).  A raw-RGB
// constructor `DXTImage(int,int,uint8_t*)` is declared (dxt_image.h:52) but never
// defined, so this TU textually includes dxt_image.cpp (to reach its file-static
// helpers CompressRGB / PhysicalToLogicalBlocks) and supplies that missing
// definition: it performs exactly the steps of the PNG constructor after stbi_load
// (dxt_image.cpp:403-434).  mptc_ref_selfcheck_png() proves both constructors agree.
#include "dxt_image.cpp"   // reference TU, unmodified (found via -I<ref>/codec)

#include "codec.h"
#include "arithmetic_codec.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <unistd.h>

namespace MPTC {
// The declared-but-undefined raw constructor (dxt_image.h:52).
DXTImage::DXTImage(int width, int height, uint8_t *rgb) {
  _width = width;
  _height = height;
  _is_intra = true;
  _search_area = 0;
  _blocks_width = (_width + 3) >> 2;
  _blocks_height = (_height + 3) >> 2;
  const int num_blocks = _blocks_width * _blocks_height;
  _num_blocks = num_blocks;
  _found_at.resize(num_blocks);
  _found.resize(num_blocks);
  std::fill(_found.begin(), _found.end(), 0);
  _index_mask.resize(num_blocks);
  _src_img.assign(rgb, rgb + (size_t)_width * _height * 3);
  _physical_blocks.resize(num_blocks);
  for (int physical_idx = 0; physical_idx < num_blocks; ++physical_idx) {
    int i = physical_idx % _blocks_width;
    int j = physical_idx / _blocks_width;
    const unsigned char *offset_data = _src_img.data() + ((size_t)j * 4 * _width + i * 4) * 3;
    _physical_blocks[physical_idx].dxt_block = CompressRGB(offset_data, _width);
  }
  _logical_blocks = std::move(PhysicalToLogicalBlocks(_physical_blocks));
  MakeDict();
}
}  // namespace MPTC

// Decoder functions the reference defines in codec.cpp (external linkage) but does not declare
// in codec.h; declared here so the wrapper can drive them on a stream held in memory.
namespace MPTC {
void ReconstructDXTData(std::vector<uint32_t> &unique_indices,
                        std::vector<std::tuple<uint8_t, uint8_t> > &motion_indices,
                        std::unique_ptr<DXTImage> &curr_frame, std::unique_ptr<DXTImage> &prev_frame,
                        uint8_t search_area, std::string ep_dir, uint32_t frame_number);
void EntropyDecode(std::vector<uint8_t> &compressed_data, std::vector<uint8_t> &out_symbols, bool is_bit_model);
void ReconstructEndPoints(std::unique_ptr<DXTImage> &dxt_image, std::unique_ptr<std::vector<uint8_t> > &wav_ep1_Y,
                          std::unique_ptr<std::vector<uint8_t> > &wav_ep1_C,
                          std::unique_ptr<std::vector<uint8_t> > &wav_ep2_Y,
                          std::unique_ptr<std::vector<uint8_t> > &wav_ep2_C);
}  // namespace MPTC

using MPTC::DXTImage;

struct RefFrame {
  std::unique_ptr<DXTImage> img;
  double t_fit_s = 0, t_search_s = 0, t_entropy_s = 0;
};

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void decode_stream(const uint8_t *src, uint32_t nbytes, size_t nsym, uint8_t *dst) {
  // The reference's own arithmetic decoder (arithmetic_codec.cpp:391) turns a stream
  // back into the symbol plane the encoder consumed.
  std::vector<uint8_t> buf(src, src + nbytes);
  buf.resize(nbytes + 16, 0);
  entropy::Arithmetic_Codec dec((unsigned)buf.size(), buf.data());
  entropy::Adaptive_Data_Model model(257);
  dec.start_decoder();
  for (size_t i = 0; i < nsym; ++i) dst[i] = (uint8_t)dec.decode(model);
  dec.stop_decoder();
}

extern "C" {

void mptc_ref_quiet(void) {
  // The reference prints per-frame chatter on std::cout (codec.cpp:846-848 etc).
  std::cout.setstate(std::ios_base::failbit);
}

void *mptc_ref_frame_new(int w, int h, const uint8_t *rgb, int is_intra, int search_area,
                         int err_threshold) {
  RefFrame *f = new RefFrame;
  double t0 = now_s();
  f->img.reset(new DXTImage(w, h, const_cast<uint8_t *>(rgb)));
  f->t_fit_s = now_s() - t0;
  f->img->_is_intra = is_intra != 0;
  f->img->_search_area = search_area;
  DXTImage::_err_threshold = err_threshold;  // what the PNG ctor does (dxt_image.cpp:393-394)
  vErrThreshold = err_threshold;
  return f;
}

void mptc_ref_frame_free(void *p) { delete static_cast<RefFrame *>(p); }

int mptc_ref_num_blocks(void *p) { return static_cast<RefFrame *>(p)->img->_num_blocks; }

void mptc_ref_get_blocks(void *p, uint64_t *out) {
  auto &v = static_cast<RefFrame *>(p)->img->_physical_blocks;
  for (size_t i = 0; i < v.size(); ++i) out[i] = v[i].dxt_block;
}

// Reencode(prev, -1) exactly as CompressMultiUnique calls it (codec.cpp:1394).
void mptc_ref_reencode(void *p, void *prev) {
  RefFrame *f = static_cast<RefFrame *>(p);
  std::unique_ptr<DXTImage> null_ref;
  double t0 = now_s();
  if (prev) f->img->Reencode(static_cast<RefFrame *>(prev)->img, -1);
  else f->img->Reencode(null_ref, -1);
  f->t_search_s = now_s() - t0;
}

void mptc_ref_get_motion(void *p, uint8_t *out /* 2*nb */) {
  auto &v = static_cast<RefFrame *>(p)->img->_motion_indices;
  for (size_t i = 0; i < v.size(); ++i) {
    out[2 * i] = std::get<0>(v[i]);
    out[2 * i + 1] = std::get<1>(v[i]);
  }
}

int mptc_ref_num_unique(void *p) {
  return (int)static_cast<RefFrame *>(p)->img->_unique_palette.size();
}

void mptc_ref_get_unique(void *p, uint32_t *out) {
  auto &v = static_cast<RefFrame *>(p)->img->_unique_palette;
  memcpy(out, v.data(), v.size() * 4);
}

// Encoder-internal PSNR (logical blocks; dxt_image.cpp:363).
double mptc_ref_psnr_logical(void *p) { return static_cast<RefFrame *>(p)->img->PSNR(); }

// PSNR of what a decoder reconstructs: rebuild logical blocks from the emitted
// physical blocks with the reference's own PhysicalToLogical, then its own PSNR().
double mptc_ref_psnr_physical(void *p) {
  RefFrame *f = static_cast<RefFrame *>(p);
  std::vector<LogicalDXTBlock> keep = f->img->_logical_blocks;
  f->img->SetLogicalBlocks();
  double r = f->img->PSNR();
  f->img->_logical_blocks = keep;
  return r;
}

// Runs the reference's per-frame EntropyEncode (codec.cpp:1115 -> CompressEndpoint :804)
// and returns the payload bytes it appends.  Only defined by the reference when the
// endpoint planes are multiples of 64 blocks (image_processing.h:293-294).
int mptc_ref_entropy_encode(void *p, uint8_t *out, int out_cap) {
  RefFrame *f = static_cast<RefFrame *>(p);
  std::vector<uint8_t> bytes;
  double t0 = now_s();
  MPTC::EntropyEncode(f->img, bytes);
  f->t_entropy_s = now_s() - t0;
  if ((int)bytes.size() > out_cap) return -(int)bytes.size();
  memcpy(out, bytes.data(), bytes.size());
  return (int)bytes.size();
}

// Splits a frame payload produced above into its five streams and decodes the four
// endpoint streams back to symbol planes with the reference's decoder:
// planes = ep1_Y(nb) | ep1_Co(nb) | ep1_Cg(nb) | ep2_Y | ep2_Co | ep2_Cg.
// sizes[5] = compressed bytes of motion, Y1, C1, Y2, C2.
int mptc_ref_payload_planes(const uint8_t *payload, int nbytes, int nb, uint8_t *planes,
                            uint8_t *motion /* 2*nb */, uint32_t *sizes) {
  const uint8_t *q = payload;
  uint32_t n_unique, msz;
  memcpy(&n_unique, q, 4); q += 4;
  memcpy(&msz, q, 4); q += 4;
  sizes[0] = msz;
  decode_stream(q, msz, (size_t)2 * nb, motion);
  q += msz;
  uint8_t *dst = planes;
  for (int s = 0; s < 4; ++s) {
    uint32_t sz;
    memcpy(&sz, q, 4); q += 4;
    sizes[1 + s] = sz;
    size_t nsym = (s & 1) ? (size_t)2 * nb : (size_t)nb;
    decode_stream(q, sz, nsym, dst);
    dst += nsym;
    q += sz;
  }
  return (int)(q - payload) == nbytes ? (int)n_unique : -1;
}

void mptc_ref_get_times(void *p, double *out3) {
  RefFrame *f = static_cast<RefFrame *>(p);
  out3[0] = f->t_fit_s; out3[1] = f->t_search_s; out3[2] = f->t_entropy_s;
}

// Encode an arbitrary byte vector the way the palette / plane streams are encoded
// (codec.cpp:186-197: Adaptive_Data_Model(257), start..stop).  Used to pin the host coder.
int mptc_ref_arith_encode(const uint8_t *sym, int n, uint8_t *out, int out_cap) {
  std::vector<uint8_t> tmp((size_t)n * 2 + 1024, 0);
  entropy::Arithmetic_Codec enc((unsigned)tmp.size(), tmp.data());
  entropy::Adaptive_Data_Model model(257);
  enc.start_encoder();
  for (int i = 0; i < n; ++i) enc.encode(sym[i], model);
  unsigned nb = enc.stop_encoder();
  if ((int)nb > out_cap) return -(int)nb;
  memcpy(out, tmp.data(), nb);
  return (int)nb;
}

// Whole-sequence encode through the reference's own driver (codec.cpp:1307), from a
// directory of PNGs to a stream file.
void mptc_ref_compress_multi_unique(const char *dir, const char *out_file, unsigned search_area,
                                    int thr, unsigned intra_interval, unsigned unique_interval) {
  MPTC::CompressMultiUnique(dir, out_file, search_area, thr, intra_interval, unique_interval, "");
}

// The reference's decoder on a stream held in memory: the loop of DecompressMultiUnique
// (codec.cpp:1161-1305) with std::ifstream::read replaced by memcpy, every decoding step done by
// the reference's own functions (EntropyDecode :560, the decoder-side DXTImage constructor
// dxt_image.cpp:437, ReconstructDXTData :393, ReconstructEndPoints :697, SetLogicalBlocks,
// DecompressedImage dxt_image.cpp:463).  DecompressMultiUnique itself only writes PNGs (and only
// without NDEBUG); this returns the frames' physical blocks and decoded pixels.  Frame sizes must
// be multiples of 256 (the reference's wavelet, image_processing.h:293-294).  Returns the frame
// count, or -1 if the stream is truncated.
int mptc_ref_decode_stream(const uint8_t *stream, int nbytes, uint64_t *blocks_out, uint8_t *rgb_out) {
  size_t off = 0;
  auto rd = [&](void *dst, size_t n) {
    if (off + n > (size_t)nbytes) return false;
    memcpy(dst, stream + off, n);
    off += n;
    return true;
  };
  uint32_t frame_height, frame_width, total_groups, mx[5];
  uint8_t unique_interval, search_area;
  if (!rd(&frame_height, 4) || !rd(&frame_width, 4) || !rd(&unique_interval, 1) || !rd(&search_area, 1) ||
      !rd(&total_groups, 4) || !rd(mx, 20))
    return -1;
  const uint32_t num_blocks = (frame_height / 4 * frame_width / 4);
  std::unique_ptr<DXTImage> prev_frame(nullptr), curr_frame(nullptr);
  uint32_t frame_number = 0;
  for (uint32_t g = 0; g < total_groups; ++g) {
    uint32_t compressed_palette_size, unique_count, unique_idx_offset = 0;
    if (!rd(&compressed_palette_size, 4)) return -1;
    std::vector<uint8_t> compressed_combined_palette(compressed_palette_size);
    if (!rd(compressed_combined_palette.data(), compressed_palette_size) || !rd(&unique_count, 4)) return -1;
    std::vector<uint8_t> combined_8bit_palette(unique_count);
    MPTC::EntropyDecode(compressed_combined_palette, combined_8bit_palette, false);
    for (uint8_t k = 0; k < unique_interval; ++k) {
      uint32_t num_unique, sz;
      if (!rd(&num_unique, 4) || !rd(&sz, 4)) return -1;
      std::vector<uint32_t> unique_indices(num_unique, 0);
      memcpy(unique_indices.data(), combined_8bit_palette.data() + unique_idx_offset, 4 * (size_t)num_unique);
      unique_idx_offset += 4 * num_unique;
      std::vector<uint8_t> comp(sz), motion(2 * (size_t)num_blocks, 0);
      if (!rd(comp.data(), sz)) return -1;
      MPTC::EntropyDecode(comp, motion, false);
      std::vector<std::tuple<uint8_t, uint8_t> > out_motion_indices;
      for (size_t i = 0; i < motion.size(); i += 2) out_motion_indices.push_back(std::make_tuple(motion[i], motion[i + 1]));
      std::unique_ptr<std::vector<uint8_t> > wav[4];
      for (int q = 0; q < 4; ++q) {
        wav[q].reset(new std::vector<uint8_t>((q & 1) ? 2 * (size_t)num_blocks : num_blocks));
        if (!rd(&sz, 4)) return -1;
        comp.resize(sz);
        if (!rd(comp.data(), sz)) return -1;
        MPTC::EntropyDecode(comp, *wav[q], false);
      }
      curr_frame.reset(new DXTImage(frame_width, frame_height, false, unique_indices));
      MPTC::ReconstructDXTData(unique_indices, out_motion_indices, curr_frame, prev_frame, search_area, "/nonexistent",
                               frame_number + 1);
      MPTC::ReconstructEndPoints(curr_frame, wav[0], wav[1], wav[2], wav[3]);
      curr_frame->SetLogicalBlocks();
      if (blocks_out)
        for (uint32_t i = 0; i < num_blocks; ++i) blocks_out[(size_t)frame_number * num_blocks + i] = curr_frame->_physical_blocks[i].dxt_block;
      if (rgb_out) {
        std::vector<uint8_t> px = curr_frame->DecompressedImage()->Pack();
        memcpy(rgb_out + (size_t)frame_number * frame_width * frame_height * 3, px.data(), (size_t)frame_width * frame_height * 3);
      }
      prev_frame = std::move(curr_frame);
      ++frame_number;
    }
  }
  return (int)frame_number;
}

// DXTImage::InterPixelSearch (dxt_image.cpp:776-832) for every block of `p` against `prev`, with the
// pattern of SetPattern(search_area) (dxt_image.h:135-164).  The function is compiled into the reference
// but its call site in Reencode is commented out (:930-951).
void mptc_ref_inter_pixel_search(void *p, void *prev, int search_area, int32_t *min_err_out, uint8_t *motion_out,
                                 uint32_t *index_out, uint8_t *reassigned_out) {
  RefFrame *f = static_cast<RefFrame *>(p);
  RefFrame *r = static_cast<RefFrame *>(prev);
  DXTImage::SetPattern(search_area);
  const int nb = f->img->_num_blocks;
  for (int b = 0; b < nb; ++b) {
    MPTC::CompressedBlock blk;
    bool re = false;
    int32_t x = 0, y = 0;
    uint32_t index = 0;
    min_err_out[b] = f->img->InterPixelSearch(r->img, b, x, y, -1, index, blk, re);
    motion_out[2 * b] = (uint8_t)x;
    motion_out[2 * b + 1] = (uint8_t)y;
    index_out[b] = index;
    reassigned_out[b] = re ? 1 : 0;
  }
}

// The same search with the reference's undefined behaviour removed.  Get4X4InterpolationBlock
// (dxt_image.cpp:619-634) hands a LogicalDXTBlock whose endpoints and palette were never initialised to
// LogicalToPhysical, which flips the gathered word (^= 0x55555555) or not depending on that stack
// garbage (:150-170): inside InterPixelSearch the garbage is whatever the previous candidate left
// there, so the compiled function's results are not a function of its inputs.  This variant runs the
// loop of InterPixelSearch (:790-829) over the reference's own CompressedBlock methods
// (AssignIndices, operator==, RecalculateEndpoints, Error, LogicalToPhysical) with the candidate word
// built from InterpolationValueAt (:604-608) directly -- the 16 gathered indices, never flipped.
void mptc_ref_inter_pixel_search_defined(void *p, void *prev, int search_area, int32_t *min_err_out,
                                         uint8_t *motion_out, uint32_t *index_out, uint8_t *reassigned_out) {
  RefFrame *f = static_cast<RefFrame *>(p);
  RefFrame *r = static_cast<RefFrame *>(prev);
  DXTImage::SetPattern(search_area);
  DXTImage &img = *f->img;
  for (int b = 0; b < img._num_blocks; ++b) {
    const int block_x = b % img._blocks_width, block_y = b / img._blocks_width;
    MPTC::CompressedBlock blk;
    blk._logical = img._logical_blocks[b];
    blk._uncompressed = img.Get4X4ColorsBlock(4 * block_x, 4 * block_y);
    const int orig_err = static_cast<int>(blk.Error());
    int min_err = std::numeric_limits<int>::max();
    int32_t mx = 0, my = 0;
    uint32_t index = 0;
    bool re = false;
    for (auto val : DXTImage::_search_pattern) {
      const int i = std::get<0>(val), j = std::get<1>(val);
      const int x = 4 * block_x + i, y = 4 * block_y + j;
      if (!(x >= 0 && x <= img._width - 4 && y >= 0 && y <= img._height - 4)) continue;
      uint32_t indices = 0;
      for (int v = 0; v < 4; ++v)
        for (int u = 0; u < 4; ++u) indices |= (uint32_t)r->img->InterpolationValueAt(x + u, y + v) << (2 * (4 * v + u));
      MPTC::CompressedBlock blk2 = blk;
      blk2.AssignIndices(indices);
      bool reset = false;
      if (!(blk2 == blk)) {
        blk2.RecalculateEndpoints();
        reset = true;
        PhysicalDXTBlock maybe = MPTC::LogicalToPhysical(blk2._logical);
        if (!(maybe.interp == indices && blk2._logical.palette[3][3] == 0xFF)) continue;
      }
      const int err_diff = static_cast<int>(blk2.Error()) - orig_err;
      if (err_diff < min_err) {
        min_err = err_diff; mx = i + 64; my = j + 64; index = indices; re = reset;
        if (err_diff <= 0) { min_err = 0; break; }
      }
    }
    min_err_out[b] = min_err;
    motion_out[2 * b] = (uint8_t)mx;
    motion_out[2 * b + 1] = (uint8_t)my;
    index_out[b] = index;
    reassigned_out[b] = re ? 1 : 0;
  }
}

// Proves the supplied raw constructor == the reference's PNG constructor
// (dxt_image.cpp:385): writes the frame as PNG with the reference's bundled
// stb_image_write, loads it through the real constructor, compares every block.
int mptc_ref_selfcheck_png(int w, int h, const uint8_t *rgb) {
  char path[64];
  snprintf(path, sizeof path, "/tmp/mptc_ref_selfcheck_%d.png", (int)getpid());
  if (!stbi_write_png(path, w, h, 3, rgb, 3 * w)) return -1;
  DXTImage a(std::string(path), true, 4, 50);
  DXTImage b(w, h, const_cast<uint8_t *>(rgb));
  unlink(path);
  if (a._physical_blocks.size() != b._physical_blocks.size()) return -2;
  int bad = 0;
  for (size_t i = 0; i < a._physical_blocks.size(); ++i)
    bad += a._physical_blocks[i].dxt_block != b._physical_blocks[i].dxt_block;
  return bad;
}

// Writes a frame as PNG with the reference's bundled stb_image_write (for building the PNG
// directory that CompressMultiUnique reads).
int mptc_ref_write_png(const char *path, int w, int h, const uint8_t *rgb) {
  return stbi_write_png(path, w, h, 3, rgb, 3 * w) ? 0 : -1;
}

}  // extern "C"
