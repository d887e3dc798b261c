/* oracle/mptc_oracle.h -- TEST INFRASTRUCTURE (CPU restatement of the reference's
 * encoder hot path).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product (mptc_b200/) never does.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against the
 * unmodified reference compiled into oracle/_ref/libmptc_ref.so (tests/test_oracle_vs_ref.py,
 * where /root/reference exists) and against the committed fixtures in tests/golden/
 * that were generated from it (tests/golden/gen_golden.py).
 */
#ifndef MPTC_ORACLE_H
#define MPTC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* stb_compress_dxt_block(HIGHQUAL) over a whole frame (dxt_image.cpp:419-431). */
void mptc_oracle_dxt1_fit(const uint8_t *rgb, int w, int h, uint64_t *blocks_out);

/* DXTImage::Reencode (dxt_image.cpp:868-957).  blocks: in = initial stb blocks,
 * out = final physical blocks.  prev_blocks = previous frame's FINAL blocks (ignored
 * when is_intra).  motion_out: 2*nb bytes.  unique_out: up to nb words.  Returns the
 * number of unique words. */
int mptc_oracle_reencode(const uint8_t *rgb, int w, int h, int is_intra, int search_area,
                         int err_threshold, uint64_t *blocks, const uint64_t *prev_blocks,
                         uint8_t *motion_out, uint32_t *unique_out);

/* One candidate evaluation (dxt_image.cpp:739-758): returns 1 if accepted and writes
 * err_diff; pixels = 48 bytes RGB of the 4x4 block, own = the block's initial stb block. */
int mptc_oracle_eval_candidate(const uint8_t *pixels48, uint64_t own_block, uint32_t cand_word,
                               int *err_diff, uint64_t *new_block);

/* DXTImage::InterPixelSearch (dxt_image.cpp:776-832) for every block of a frame, each against the
 * previous frame's final blocks: min_err (INT_MAX if nothing was accepted), motion = (i + 64, j + 64)
 * of the winning pixel offset, the winning index word and re_assigned.  cur_blocks = the frame's
 * blocks at the time of the call (the initial stb fit when called before Reencode).  Follows the
 * reference as compiled with the canonical flags (see the UB note in mptc_oracle.c). */
void mptc_oracle_inter_pixel_search(const uint8_t *rgb, int w, int h, int sa, const uint64_t *cur_blocks,
                                    const uint64_t *prev_blocks, int32_t *min_err_out, uint8_t *motion_out,
                                    uint32_t *index_out, uint8_t *reassigned_out);
/* DXTImage::SetPattern (dxt_image.h:135-164): the pixel offsets in search order; returns their count. */
int mptc_oracle_ips_pattern(int sa, int8_t *ij);

/* Decoder-side word reconstruction (ReconstructDXTData, codec.cpp:441-500); returns the
 * number of unique words consumed or -1 if a motion vector is invalid. */
int mptc_oracle_reconstruct_words(const uint8_t *motion, const uint32_t *unique, int n_unique,
                                  const uint32_t *prev_words, int bw, int bh, int sa, uint32_t *out);

/* Per-block inductive check of a frame's results (see mptc_oracle.c); returns #mismatches. */
int mptc_oracle_check_blocks(const uint8_t *rgb, int w, int h, int is_intra, int sa, int thr,
                             const uint64_t *init_blocks, const uint64_t *cur_final,
                             const uint64_t *prev_final, const uint8_t *motion,
                             const int *which, int n_which);

/* Endpoint planes (codec.cpp:804-845 + image_processing.h:292-333 + wavelet.cpp:30-131):
 * planes_out = ep1_Y | ep1_Co | ep1_Cg | ep2_Y | ep2_Co | ep2_Cg, each pbw*pbh symbols,
 * where pbw/pbh = bw/bh rounded up to a multiple of 64.  For bw,bh multiples of 64 this is
 * the reference's behaviour; otherwise the planes are edge-replicated up to the next
 * multiple of 64 first -- an EXTENSION (the reference asserts, image_processing.h:293). */
void mptc_oracle_endpoint_planes(const uint64_t *blocks, int bw, int bh, uint8_t *planes_out);

/* FastAC adaptive-model arithmetic encoder as used by codec.cpp:186-197
 * (Adaptive_Data_Model(257), start_encoder .. stop_encoder).  Returns bytes written. */
int mptc_oracle_arith_encode(const uint8_t *sym, int n, uint8_t *out, int out_cap);

/* Decoder side.  Arithmetic_Codec::decode with Adaptive_Data_Model(257)
 * (arithmetic_codec.cpp:391-444, codec.cpp:560-577): n symbols from nbytes of code (the buffer
 * must be readable 4 bytes past the end).  Returns 0, -1 if the code runs out. */
int mptc_oracle_arith_decode(const uint8_t *code, int nbytes, uint8_t *sym_out, int n);

/* ReconstructEndPoints (codec.cpp:697-800): 6 symbol planes -> ep1 / ep2 (bw*bh RGB565 each). */
void mptc_oracle_inverse_planes(const uint8_t *planes, int bw, int bh, uint16_t *ep1, uint16_t *ep2);

/* DXTImage::DecompressedImage (dxt_image.cpp:463-479) of decoded physical blocks: w*h*3 bytes. */
void mptc_oracle_decode_rgb(const uint64_t *blocks, int w, int h, uint8_t *rgb_out);

/* PSNR of the decoded physical blocks against the source (dxt_image.cpp:363-383 applied
 * to PhysicalToLogical of the emitted blocks). */
double mptc_oracle_psnr(const uint8_t *rgb, int w, int h, const uint64_t *blocks);

/* stb single-colour match tables (stb_dxt.h:111-137), for pinning the product's generator. */
void mptc_oracle_tables(uint8_t *omatch5 /*512*/, uint8_t *omatch6 /*512*/);

/* Multi-threaded GOP encode used as the CPU baseline: frames[n][h][w][3], every `gop`-th
 * frame intra, one GOP per worker thread; returns seconds of wall time. out_blocks may be NULL. */
double mptc_oracle_encode_gops(const uint8_t *frames, int n_frames, int w, int h, int gop,
                               int search_area, int err_threshold, int threads,
                               uint64_t *out_blocks, uint8_t *out_motion);

#ifdef __cplusplus
}
#endif
#endif
