// mptc_kernels.h -- launch interface between the C-ABI layer and the kernels.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include <mutex>

namespace mptc {

constexpr int kMaxDevices = 64;   // per-device launch configuration caches (power of two)
// The launchers cache per-device kernel attributes (opt-in shared memory, occupancy) in function-local
// statics; several host threads (one context each, include/mptc_gpu.h) may launch at once, so every
// read-modify-write of those caches happens under this mutex.
std::mutex &launch_cfg_mutex();

// Device-resident sequence: all arrays are [frame][...] over the reserved capacity.
struct SeqView {
  const uint8_t *rgb;      // [F][h][w][3]
  uint64_t *init_blocks;   // [F][nb]   stb fit (DXTImage ctor result)
  uint64_t *final_blocks;  // [F][nb]   after Reencode
  uint8_t *motion;         // [F][nb][2]
  uint8_t *flags;          // [F][nb]   1 = final after the inter search
  uint8_t *row_todo;       // [F][bh]   1 = the row has blocks the inter search left over
  uint32_t *unique;        // [F][nb]
  uint32_t *n_unique;      // [F]
  uint32_t *chunk_counts;  // [F][ceil(nb/1024)]  unique blocks per chunk (K4)
  uint8_t *planes;         // [F][6][pbh][pbw]
  int *progress;           // [F][bh]   wavefront progress counters (direct fallback kernel only)
  unsigned long long *wordflag;  // [F][nb] {index word, epoch}: how the rows of the intra wavefront hand decisions over
  uint32_t epoch;          // tag of this encode call in `wordflag` (entries of earlier calls are stale)
  unsigned long long *work; // executed-work counters of the search kernels (kWork*), one atomicAdd per tile / group
  size_t frame_bytes;
  int w, h, bw, bh, nb;
  int first, count;        // frame range of this encode call
  int gop;
};

// Executed work of the search kernels since the start of the last encode call (bench.py's roofline):
// (index word, target block) evaluations actually run and window positions actually scanned -- the
// de-duplicating kernels evaluate each distinct word of a tile once, so these are far below the
// nominal candidate positions of SURVEY.md 8(d).
enum { kWorkNominalInter = 0, kWorkNominalIntra = 1, kWorkInterEvals = 2, kWorkInterScanned = 3,
       kWorkIntraEvals = 4, kWorkIntraScanned = 5, kWorkInterTiles = 6, kWorkIntraGroups = 7, kWorkCounters = 8 };

// Inter frames: K3s (mptc_sparse.cu) handles frames with at most max_items leftover blocks and
// tells the row wavefront through n_unique[f] (0xFFFFFFFF = not handled, take the frame).
constexpr uint32_t kSparseNotHandled = 0xFFFFFFFFu;
// Returns false if the kernel could not be configured (nothing was launched).
bool launch_intra_sparse(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, int *tickets, int ctas_per_frame,
                         int max_items, cudaStream_t s);

cudaError_t upload_tables(const uint8_t *omatch5, const uint8_t *omatch6);
// K1/K4/K5 run over frames f0, f0 + fstride, ... (nf of them).
void launch_dxt1_fit(const SeqView &v, int f0, int fstride, int nf, cudaStream_t s);
// Returns the number of kernels launched.
int launch_inter_search(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, cudaStream_t s);
bool launch_intra_rows(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, int *ticket, int grid_cap,
                       cudaStream_t s);
bool launch_inter_search_tiled(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, cudaStream_t s);
bool launch_inter_search_wide(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, bool two_per_sm_only, cudaStream_t s);
// grid_cap > 0 limits the number of CTAs (rows in flight): a wavefront only keeps a few rows per
// frame busy, and idle CTAs would block the SMs for kernels of other lanes.
void launch_intra_wavefront(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, int *ticket,
                            int max_ctas, int grid_cap, cudaStream_t s);
// host_unique / host_n_unique: device-visible addresses of page-locked host result buffers (or null):
// the unique words are then also written straight to the host (frame f -> host frame f - host_first).
void launch_compact_unique(const SeqView &v, int sa, unsigned long long *cand_counts, int f0, int fstride, int nf,
                           cudaStream_t s, uint32_t *host_unique = nullptr, uint32_t *host_n_unique = nullptr,
                           int host_first = 0);
void launch_endpoint_planes(const SeqView &v, int pbw, int pbh, int f0, int fstride, int nf, cudaStream_t s);
int intra_wavefront_max_ctas(int device);

// ---- K2p: pixel-granular inter search (mptc_pixel.cu) ------------------------------------------------
int inter_pixel_pattern(int sa, int8_t *ij);   // DXTImage::SetPattern order; returns the number of offsets
void launch_inter_pixel_search(const uint8_t *frame, int w, int h, int sa, int n_pat, const int8_t *pat,
                               const uint64_t *cur_blocks, const uint64_t *prev_blocks, int32_t *min_err,
                               uint8_t *motion, uint32_t *index, uint8_t *reassigned, cudaStream_t s);

// ---- decoder side (mptc_decode.cu) -------------------------------------------------------
// Frames [first, first + count) of the device-resident sequence, every gop-th one intra.
struct DecView {
  const uint8_t *motion;        // [F][nb][2]   decoded motion bytes
  const uint32_t *unique;       // unique index words; frame f's start at unique[unique_off[f]]
  const uint32_t *unique_off;   // [F]
  const uint32_t *n_unique;     // [F]
  const uint8_t *planes;        // [F][6][pbh][pbw] decoded wavelet symbols
  uint32_t *words;              // [F][nb]   index word of the unique blocks
  int *link;                    // [F][nb]   node the block's index word comes from
  uint32_t *chunk_counts;       // [F][dec_chunks(nb)]
  int *errors;                  // [0] invalid motion vectors (cumulative until read)
  int *status;                  // [1 + p] pass p of the pointer jumping left work ([0] unused)
  uint64_t *blocks;             // [F][nb]   out: PhysicalDXTBlock
  uint8_t *rgb;                 // [F][h][w][3] out (optional)
  int w, h, bw, bh, nb, pbw, pbh;
  int first, count, gop, sa;
};
int dec_chunks(int nb);
int dec_jump_passes(int gop, int nb);
cudaError_t decode_kernels_init();
// Each returns the number of kernels it launched.
int launch_decode_words(const DecView &v, cudaStream_t s);
int launch_inverse_planes(const DecView &v, cudaStream_t s);
int launch_dxt1_to_rgb(const DecView &v, cudaStream_t s);

}  // namespace mptc
