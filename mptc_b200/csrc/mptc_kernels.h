// mptc_kernels.h -- launch interface between the C-ABI layer and the kernels.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace mptc {

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
  uint8_t *planes;         // [F][6][pbh][pbw]
  int *progress;           // [F][bh]   wavefront progress counters
  size_t frame_bytes;
  int w, h, bw, bh, nb;
  int first, count;        // frame range of this encode call
  int gop;
};

// Inter frames: K3s (mptc_sparse.cu) handles frames with at most max_items leftover blocks and
// tells the row wavefront through n_unique[f] (0xFFFFFFFF = not handled, take the frame).
constexpr uint32_t kSparseNotHandled = 0xFFFFFFFFu;
void launch_intra_sparse(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, int *tickets, int ctas_per_frame,
                         int max_items, cudaStream_t s);

cudaError_t upload_tables(const uint8_t *omatch5, const uint8_t *omatch6);
// K1/K4/K5 run over frames f0, f0 + fstride, ... (nf of them).
void launch_dxt1_fit(const SeqView &v, int f0, int fstride, int nf, cudaStream_t s);
void launch_inter_search(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, cudaStream_t s);
bool launch_intra_wavefront_tiled(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, int *ticket,
                                  int grid_cap, cudaStream_t s);
bool launch_inter_search_tiled(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, cudaStream_t s);
// grid_cap > 0 limits the number of CTAs (rows in flight): a wavefront only keeps a few rows per
// frame busy, and idle CTAs would block the SMs for kernels of other lanes.
void launch_intra_wavefront(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, int *ticket,
                            int max_ctas, int grid_cap, cudaStream_t s);
void launch_compact_unique(const SeqView &v, int sa, unsigned long long *cand_counts, int f0, int fstride, int nf,
                           cudaStream_t s);
void launch_endpoint_planes(const SeqView &v, int pbw, int pbh, int f0, int fstride, int nf, cudaStream_t s);
int intra_wavefront_max_ctas(int device);

}  // namespace mptc
