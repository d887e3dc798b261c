// mptc_inter.cu -- K2: tiled inter search with per-tile de-duplication of index words.
//
// Restates DXTImage::InterBlockSearch + the winner apply in Reencode
// (codec/dxt_image.cpp:715-774, :885-908).  The reference evaluates every window position;
// the evaluation result depends only on (target block, 32-bit index word), and after
// re-encoding a frame's words are overwhelmingly copies of one another (that is MPTC's
// whole point), so a 32x32 window typically holds well under 100 distinct words.  One CTA
// takes an 8x4 tile of targets, de-duplicates the words of the union of their windows in
// shared memory, evaluates each DISTINCT word once per target (lane = target, word uniform
// across the warp: mptc_uniform_eval.cuh), then every target scans its own window positions
// through the (word id -> err_diff) table with the order-independent form of the reference's
// stateful winner scan (WinnerState).  Results are bit-identical to the
// position-by-position loop.
#include "mptc_inter_tile.cuh"

namespace mptc {

#ifdef MPTC_K2_PHASE_TIMING
extern "C" void mptc_debug_k2_cycles(unsigned long long *out12, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out12, tile32::g_k2_cycles, sizeof(unsigned long long) * 12);
  if (reset) { unsigned long long z[12] = {0}; cudaMemcpyToSymbol(tile32::g_k2_cycles, z, sizeof z); }
}
#endif

using namespace tile32;

__global__ void __launch_bounds__(kThreads, 2)
k_inter_search_tiled(SeqView v, int k_in_gop, int sa, int thr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int f = v.first + blockIdx.y * v.gop + k_in_gop;
  if (f >= v.first + v.count) return;
  const int tiles_x = (v.bw + kTileX - 1) / kTileX;
  search_tile(v, f, (blockIdx.x % tiles_x) * kTileX, (blockIdx.x / tiles_x) * kTileY, sa, thr, smem_raw);
}

// Returns false when the tile's shared memory does not fit (very large search_area): the
// caller then uses the direct kernel.
bool launch_inter_search_tiled(const SeqView &v, int k_in_gop, int n_gops, int sa, int thr, cudaStream_t s) {
  static int max_optin = -1;
  static size_t configured[kMaxDevices] = {0};   // per device: one context per GPU may live in one process
  const size_t bytes = tile_smem_bytes(sa, nullptr, nullptr);
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  {
    std::lock_guard<std::mutex> lock(launch_cfg_mutex());
    if (max_optin < 0) cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cur_dev);
    if (bytes + 1024 > (size_t)max_optin) return false;
    size_t &conf = configured[cur_dev & (kMaxDevices - 1)];
    if (bytes > conf) {
      if (cudaFuncSetAttribute(k_inter_search_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
        return false;
      conf = bytes;
    }
  }
  const int tiles = ((v.bw + kTileX - 1) / kTileX) * ((v.bh + kTileY - 1) / kTileY);
  dim3 grid(tiles, n_gops);
  k_inter_search_tiled<<<grid, kThreads, bytes, s>>>(v, k_in_gop, sa, thr);
  return true;
}

}  // namespace mptc
