/* mptc_gpu.h -- C-ABI of the B200-native MPTC encoder hot path (libmptc_b200.so).
 *
 * Plain pointers and sizes only; no CUDA or torch types cross this boundary.  Each entry
 * point names the reference interface it replaces (paths relative to the reference tree).
 * The reference has no FFI layer: its seam is the C++ class MPTC::DXTImage
 * (codec/dxt_image.h:46-198) driven by codec/codec.cpp; INTEGRATION.md shows the shim a
 * maintainer would add there.
 *
 * Conventions
 *   frames   RGB8, row-major, stride 3*w, w and h multiples of 4 (dxt_image.cpp:428).
 *   block    uint64 little-endian {u16 ep1; u16 ep2; u32 interp} = PhysicalDXTBlock
 *            (dxt_image.h:16-23); nb = (w/4)*(h/4) blocks per frame, raster order.
 *   motion   2 bytes per block (x, y) as in DXTImage::_motion_indices (dxt_image.h:182):
 *            (255,255) unique; both MSBs set = inter; otherwise intra.
 *   unique   the interp words of unique blocks in raster order (_unique_palette, :181).
 *   planes   6 symbol planes per frame, ep1_Y ep1_Co ep1_Cg ep2_Y ep2_Co ep2_Cg, each
 *            pbw*pbh bytes with pbw/pbh = bw/bh rounded up to a multiple of 64
 *            (codec.cpp:598-614 output, before the arithmetic coder).
 * All functions return 0 on success, a negative MPTC_E_* code otherwise; they never exit().
 * One context per (host thread, GPU); contexts share nothing.
 */
#ifndef MPTC_GPU_H
#define MPTC_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPTC_OK 0
#define MPTC_E_ARG (-1)     /* bad argument (size not multiple of 4, search_area > 63, ...) */
#define MPTC_E_CUDA (-2)    /* CUDA runtime error; see mptc_gpu_last_error */
#define MPTC_E_NOMEM (-3)   /* device or pinned allocation failed */
#define MPTC_E_STATE (-4)   /* call order violated (e.g. encode before reserve/upload) */
#define MPTC_E_SPACE (-5)   /* output buffer too small (mptc_codec.h) */
#define MPTC_E_DATA (-6)    /* corrupt input: a stream / motion vector no encoder emits */

typedef struct mptc_gpu_ctx mptc_gpu_ctx;

typedef struct mptc_gpu_params {
  int search_area;    /* DXTImage::_search_area (dxt_image.h:174), 1..63 */
  int err_threshold;  /* DXTImage::_err_threshold / vErrThreshold (dxt_image.cpp:25,393) */
  int gop;            /* intra_interval == unique_interval (codec.cpp:1462-1505) */
} mptc_gpu_params;

/* ---- lifetime ------------------------------------------------------------------- */
int mptc_gpu_create(int device, mptc_gpu_ctx **out);
void mptc_gpu_destroy(mptc_gpu_ctx *ctx);
const char *mptc_gpu_last_error(const mptc_gpu_ctx *ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
uint64_t mptc_gpu_launch_count(const mptc_gpu_ctx *ctx);

/* ---- single-frame entry points (host buffers, synchronous) ------------------------ */

/* Replaces the block loop of the DXTImage constructor (dxt_image.cpp:419-431 ->
 * stb_compress_dxt_block, Include/stb_dxt.h:613). */
int mptc_gpu_dxt1_fit(mptc_gpu_ctx *ctx, const uint8_t *rgb, int w, int h, uint64_t *blocks_out);

/* Replaces DXTImage constructor + DXTImage::Reencode(reference, -1) for one frame
 * (dxt_image.cpp:385-436, :868-957).  prev_blocks = the previous frame's FINAL blocks
 * (reference->_physical_blocks), ignored / may be NULL when is_intra.  initial_out may be
 * NULL.  *n_unique receives _unique_palette.size(). */
int mptc_gpu_reencode(mptc_gpu_ctx *ctx, const uint8_t *rgb, int w, int h, int is_intra,
                      int search_area, int err_threshold, const uint64_t *prev_blocks,
                      uint64_t *initial_out, uint64_t *blocks_out, uint8_t *motion_out,
                      uint32_t *unique_out, uint32_t *n_unique);

/* Replaces DXTImage::InterPixelSearch (dxt_image.cpp:776-832) for every block of one frame: the
 * pixel-granular inter search over the offsets of DXTImage::SetPattern(search_area) (dxt_image.h:135-164).
 * The reference never calls it (commented out of Reencode, dxt_image.cpp:930-951), so this is a
 * stand-alone analysis call and the sequence encode does not use it.  cur_blocks = the frame's blocks at
 * the time of the call (NULL: the stb fit of `rgb`, i.e. the state before Reencode); prev_blocks = the
 * reference frame's final blocks.  Per block: min_err (0 if a candidate with err_diff <= 0 exists, else
 * the smallest err_diff, INT32_MAX if nothing was accepted), motion = (i + 64, j + 64) of the winning
 * offset, the winning index word, re_assigned.  Any output may be NULL.  The candidate word is the 16
 * gathered indices as they are: the reference's Get4X4InterpolationBlock (:619-634) reads uninitialised
 * memory on the way (see mptc_pixel.cu). */
int mptc_gpu_inter_pixel_search(mptc_gpu_ctx *ctx, const uint8_t *rgb, int w, int h, int search_area,
                                const uint64_t *cur_blocks, const uint64_t *prev_blocks, int32_t *min_err_out,
                                uint8_t *motion_out, uint32_t *index_out, uint8_t *reassigned_out);

/* Replaces CompressEndpoint up to (not including) the arithmetic coder: EndpointOne/TwoValues
 * -> RGB565toYCoCg667 -> FWavelet2D<.,64> -> MakeUnsigned -> Linearize (dxt_image.cpp:496-530,
 * codec.cpp:804-839, :598-614).  planes_out: 6*pbw*pbh bytes. */
int mptc_gpu_endpoint_planes(mptc_gpu_ctx *ctx, const uint64_t *blocks, int bw, int bh,
                             uint8_t *planes_out);

/* ---- sequence entry points (the throughput path) ----------------------------------- */
/* A sequence is n_frames frames; every params.gop-th frame is intra (a GOP).  GOPs are
 * independent and are processed in lock-step on the device: frame k of
 * every GOP in the same launches, which is what fills the wavefront kernel.
 * Replaces the frame loop of CompressMultiUnique / SingleThreadCompressMulti
 * (codec.cpp:1383-1509, :1535-1694) up to the arithmetic coder. */

/* (Re)allocates device + pinned buffers for a w x h x n_frames sequence. */
int mptc_gpu_seq_reserve(mptc_gpu_ctx *ctx, int w, int h, int n_frames);
/* Copies `count` frames starting at `first` from host to the device-resident sequence. */
int mptc_gpu_seq_upload(mptc_gpu_ctx *ctx, const uint8_t *frames, int first, int count);
/* Runs fit + search + compaction + endpoint planes over frames [first, first+count);
 * first must be a GOP boundary.  Asynchronous; CUDA events bracket the kernels. */
int mptc_gpu_seq_encode(mptc_gpu_ctx *ctx, int first, int count, const mptc_gpu_params *p);
/* Copies results of frames [first, first+count) to host.  Any pointer may be NULL.
 * blocks: count*nb u64; initial: count*nb u64; motion: count*2*nb; unique: count*nb u32
 * (frame f's words at unique + f*nb); n_unique: count u32; planes: count*6*pbw*pbh. */
int mptc_gpu_seq_download(mptc_gpu_ctx *ctx, int first, int count, uint64_t *blocks,
                          uint64_t *initial, uint8_t *motion, uint32_t *unique,
                          uint32_t *n_unique, uint8_t *planes);
int mptc_gpu_sync(mptc_gpu_ctx *ctx);

/* Device time in milliseconds of the last mptc_gpu_seq_encode (CUDA events on the compute
 * stream).  stage: 0 whole encode, 1 fit, 2 inter search, 3 intra search, 4 compaction,
 * 5 endpoint planes.  Stages 1..5 are sums over the launches of that kernel. */
int mptc_gpu_last_encode_ms(mptc_gpu_ctx *ctx, int stage, float *ms);

/* Scheduling knobs of the sequence entry points (0 = automatic for each).  The GOPs of one call
 * are split into `lanes` contiguous ranges, each on its own CUDA stream: the reference's
 * ThreadedCompressMultiUnique runs one std::thread per dictionary group the same way
 * (codec/codec.cpp:1781-1793).  wave_rows_* = CTAs per frame of the intra wavefront kernel for
 * intra frames / for the leftovers of inter frames (rows of a frame that are in flight). */
int mptc_gpu_set_schedule(mptc_gpu_ctx *ctx, int lanes, int wave_rows_intra, int wave_rows_inter);

/* End to end from HOST frames to HOST results.  The work is enqueued frame by frame: H2D of
 * frame k of every GOP (copy stream), its kernels (the GOP lane's streams), D2H of its results
 * (a third stream), so the copies of frame k+1 / k-1 overlap the kernels of frame k.  `frames`
 * and the outputs should be page-locked (mptc_gpu_host_alloc) for the copies to overlap.
 * Returns when all results are on the host; mptc_gpu_last_encode_ms(0) then covers copies +
 * kernels. */
int mptc_gpu_encode_sequence(mptc_gpu_ctx *ctx, const uint8_t *frames, int n_frames, int w, int h,
                             const mptc_gpu_params *p, uint64_t *blocks, uint8_t *motion,
                             uint32_t *unique, uint32_t *n_unique, uint8_t *planes);

/* The asynchronous form ("a batched/async encode + wait for stream overlap"):
 * returns as soon as everything is enqueued.  mptc_gpu_wait_frame blocks until the results of
 * frame `frame` are in the host buffers (callable from any host thread), mptc_gpu_wait until the
 * whole call is done.  This is what lets the host's arithmetic coder (codec.cpp:1115-1158) start
 * on the first frames while the GPU is still searching the later ones. */
int mptc_gpu_encode_sequence_async(mptc_gpu_ctx *ctx, const uint8_t *frames, int n_frames, int w, int h,
                                   const mptc_gpu_params *p, uint64_t *blocks, uint8_t *motion,
                                   uint32_t *unique, uint32_t *n_unique, uint8_t *planes);
int mptc_gpu_wait_frame(mptc_gpu_ctx *ctx, int frame);
int mptc_gpu_wait(mptc_gpu_ctx *ctx);

/* ---- decoder side --------------------------------------------------- */
/* From the symbols the arithmetic decoder produced to ready-to-upload DXT1 blocks.  Replaces
 * ReconstructDXTData + ReconstructEndPoints (codec/codec.cpp:393-500, :697-800; the library
 * flavour ReconstructDXTFrame / ReconstructEndpoints, codec/decoder.cpp:68-254) and, with RGB
 * output, DXTImage::DecompressedImage (dxt_image.cpp:463-479).  The reference walks the blocks
 * of every frame in raster order, frame after frame; here all frames of all GOPs are resolved in
 * the same launches (pointer jumping over the copy chains).
 *   motion   n*2*nb bytes;  planes  n*6*pbw*pbh symbols (layout as above)
 *   unique   the index words of the unique blocks: unique_stride == 0 -> packed, frame after frame
 *            (the group palettes of the stream back to back, codec.cpp:1473-1479); otherwise frame
 *            i's words start at unique + i*unique_stride (the encoder's output layout, stride nb)
 * Returns MPTC_E_DATA if a motion vector points outside the frame / forward in raster order /
 * to a previous frame that does not exist (the reference would read out of bounds). */
int mptc_gpu_decode_sequence(mptc_gpu_ctx *ctx, const uint8_t *motion, const uint32_t *unique,
                             const uint32_t *n_unique, size_t unique_stride, const uint8_t *planes,
                             int n_frames, int w, int h, int search_area, int gop,
                             uint64_t *blocks_out, uint8_t *rgb_out /* may be NULL */);

/* The same in three asynchronous steps on the device-resident sequence (mptc_gpu_seq_reserve):
 * upload the symbols of frames [first, first+count), decode them (first = a GOP boundary), fetch
 * the results.  mptc_gpu_seq_decode without an upload decodes what the encoder left on the
 * device (motion / unique / planes of mptc_gpu_seq_encode): the device-resident round trip.
 * Decoded blocks live in their own buffer; the encoder's final blocks stay untouched. */
int mptc_gpu_seq_decode_upload(mptc_gpu_ctx *ctx, int first, int count, const uint8_t *motion,
                               const uint32_t *unique, const uint32_t *n_unique, size_t unique_stride,
                               const uint8_t *planes);
int mptc_gpu_seq_decode(mptc_gpu_ctx *ctx, int first, int count, int search_area, int gop, int want_rgb);
int mptc_gpu_seq_decode_download(mptc_gpu_ctx *ctx, int first, int count, uint64_t *blocks, uint8_t *rgb);
/* Device time (ms) of the last mptc_gpu_seq_decode: stage 0 whole, 1 index words (count, links,
 * pointer jumping), 2 inverse endpoint planes + block assembly, 3 DXT1 -> RGB. */
int mptc_gpu_last_decode_ms(mptc_gpu_ctx *ctx, int stage, float *ms);

/* Page-locked host memory for the buffers above. */
void *mptc_gpu_host_alloc(size_t bytes);
void mptc_gpu_host_free(void *p);

/* Candidate positions of the full search windows of the last mptc_gpu_seq_encode, clipped to
 * the frame (the algorithmic unit count, independent of any early exit). */
int mptc_gpu_last_candidate_count(mptc_gpu_ctx *ctx, uint64_t *inter, uint64_t *intra);

/* Work the search kernels actually EXECUTED in the last encode call (they evaluate each distinct
 * index word of a tile / group once instead of every window position, dxt_image.cpp:731-770):
 * counters[0..1] = the nominal candidate positions above (inter, intra), [2] = (index word, target
 * block) evaluations of the inter search, [3] = window positions its winner scan visited,
 * [4], [5] = the same for the intra wavefront, [6] = inter tiles, [7] = intra groups.  n = how many
 * to copy (<= 8).  This is what bench.py's roofline counts. */
int mptc_gpu_last_work_count(mptc_gpu_ctx *ctx, uint64_t *counters, int n);

#ifdef __cplusplus
}
#endif
#endif /* MPTC_GPU_H */
