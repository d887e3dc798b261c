/* mptc_codec.h -- host side of the encoder above the GPU hot path: the sequential arithmetic
 * coder and the stream assembly, C-ABI.  Exported by the same libmptc_b200.so.
 *
 * These replace, bit for bit, the reference's host code that consumes DXTImage results:
 *   mptc_arith_encode      EntropyEncode(std::vector<uint8_t>&, ...)  codec/codec.cpp:186-197
 *                          (entropy::Arithmetic_Codec + Adaptive_Data_Model(257),
 *                           entropy/arithmetic_codec.cpp:360-387, :498-571, :749-829)
 *   mptc_frame_payload     EntropyEncode(std::unique_ptr<DXTImage>&, ...) codec.cpp:1115-1158
 *                          + the packaging half of CompressEndpoint      codec.cpp:841-899
 *   mptc_encode_stream     CompressMultiUnique                           codec.cpp:1307-1532
 *                          (frames come from memory instead of a PNG directory)
 * All functions return 0 on success, negative MPTC_E_* codes otherwise (mptc_gpu.h).
 */
#ifndef MPTC_CODEC_H
#define MPTC_CODEC_H

#include <stddef.h>
#include <stdint.h>

#include "mptc_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* MPTC_E_SPACE (mptc_gpu.h): output buffer too small; *out_bytes holds the size needed */

/* Adaptive arithmetic coding of n byte symbols with a fresh 257-symbol model. */
int mptc_arith_encode(const uint8_t *sym, size_t n, uint8_t *out, size_t cap, size_t *out_bytes);

/* One frame's payload: u32 n_unique, u32 motion bytes + stream, then four (u32 size, stream)
 * records: ep1 Y, ep1 Co|Cg, ep2 Y, ep2 Co|Cg.  motion: 2*nb bytes; planes: 6 planes of
 * plane_syms bytes each (plane_syms = pbw*pbh).  sizes[5] receives the compressed sizes
 * (motion, Y1, C1, Y2, C2).  threads > 1 codes the five streams concurrently. */
int mptc_frame_payload(const uint8_t *motion, size_t nb, const uint8_t *planes, size_t plane_syms,
                       uint32_t n_unique, int threads, uint8_t *out, size_t cap, size_t *out_bytes,
                       uint32_t *sizes);

/* Stream header fields that CompressMultiUnique patches at offset 14 (codec.cpp:1514-1520). */
typedef struct mptc_stream_stats {
  uint32_t max_unique_bytes, max_comp_palette, max_comp_motion, max_comp_ep_y, max_comp_ep_c;
  uint32_t n_groups;
  double gpu_ms;      /* device time of the hot path incl. its H2D/D2H copies (CUDA events) */
  double entropy_ms;  /* wall time of arithmetic coding + assembly on `threads` host threads; in
                         mptc_encode_stream this phase starts while the GPU is still running */
  double total_ms;    /* wall time of the whole call */
  double assemble_ms; /* of which: laying the records out in the caller's buffer */
} mptc_stream_stats;

/* Whole-sequence encode to the reference's stream format: 34-byte
 * header, then per group of p->gop frames: u32 palette size, palette stream, u32 unique bytes,
 * frame payloads.  Uses intra_interval == unique_interval == p->gop.  Frames beyond the last
 * full group are encoded but not written, exactly like the reference (codec.cpp:1477-1504).
 * The GPU part runs on `ctx`; the arithmetic coder runs on `threads` host threads, overlapped
 * with it: every frame's five streams are coded as soon as its results have reached the host
 * (mptc_gpu_encode_sequence_async + mptc_gpu_wait_frame). */
int mptc_encode_stream(mptc_gpu_ctx *ctx, const uint8_t *frames, int n_frames, int w, int h,
                       const mptc_gpu_params *p, int threads, uint8_t *out, size_t cap,
                       size_t *out_bytes, mptc_stream_stats *stats);

/* The same assembly from results already on the host (e.g. gathered from several GPUs):
 * motion n*2*nb, unique n*nb u32 (frame f at unique + f*nb), n_unique n, planes n*6*plane_syms. */
int mptc_assemble_stream(int n_frames, int w, int h, const mptc_gpu_params *p, const uint8_t *motion,
                         const uint32_t *unique, const uint32_t *n_unique, const uint8_t *planes,
                         int threads, uint8_t *out, size_t cap, size_t *out_bytes,
                         mptc_stream_stats *stats);

/* ---- decoder side --------------------------------------------------- */

/* EntropyDecode (codec/codec.cpp:560-577): n byte symbols from nbytes of code, fresh
 * Adaptive_Data_Model(257), Arithmetic_Codec::decode (arithmetic_codec.cpp:391-444).
 * MPTC_E_DATA if the code runs out or decodes a symbol outside 0..255. */
int mptc_arith_decode(const uint8_t *code, size_t nbytes, uint8_t *sym, size_t n);
/* k (1..8) independent streams -- e.g. the motion stream and the endpoint-plane streams of a frame, which
 * EntropyDecode handles one after the other -- decoded in one interleaved loop on the calling thread: each symbol
 * costs a 32-bit division and a dependent table walk, and with several chains in flight most of that latency
 * hides.  Same results and errors as k calls of mptc_arith_decode. */
int mptc_arith_decode_multi(int k, const uint8_t *const *code, const size_t *nbytes, uint8_t *const *sym, const size_t *n);

/* The 34-byte stream header (reader: codec.cpp:1172-1184). */
typedef struct mptc_stream_header {
  int width, height, gop /* unique_interval */, search_area, n_groups, n_frames /* n_groups*gop */;
  uint32_t max_unique_bytes, max_comp_palette, max_comp_motion, max_comp_ep_y, max_comp_ep_c;
} mptc_stream_header;
int mptc_stream_info(const uint8_t *stream, size_t bytes, mptc_stream_header *hdr);

typedef struct mptc_decode_stats {
  mptc_stream_header header;
  double entropy_ms;  /* wall time of the arithmetic decoding on `threads` host threads (the GPU
                         reconstructs finished groups meanwhile) */
  double total_ms;    /* wall time of the whole call */
  uint64_t symbols;   /* symbols decoded */
} mptc_decode_stats;

/* Whole-stream decode, replaces DecompressMultiUnique (codec.cpp:1161-1305; the library flavour
 * GetFrame / GetFrameMultiThread, codec/decoder.cpp:256-470): stream bytes in, the DXT1 blocks
 * (PhysicalDXTBlock, ready for texture upload) of all header.n_frames frames out, optionally the
 * decoded RGB frames as well.  Arithmetic decoding on `threads` host threads (one job per stream
 * record), reconstruction on the GPU, group by group as the host finishes them.
 * blocks_out: n_frames*nb u64 (may be NULL); rgb_out: n_frames*w*h*3 bytes (may be NULL). */
int mptc_decode_stream(mptc_gpu_ctx *ctx, const uint8_t *stream, size_t bytes, int threads,
                       uint64_t *blocks_out, uint8_t *rgb_out, mptc_decode_stats *stats);

#ifdef __cplusplus
}
#endif
#endif /* MPTC_CODEC_H */
