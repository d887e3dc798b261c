// Micro-benchmark: cost of __shfl_sync / __syncwarp / predicated STS + LDS round trips for ONE
// warp running alone on an SM (the in-row resolve loop of k_intra_wavefront_tiled).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) k(int iters, long long *out, int mode) {
  __shared__ unsigned short pos[4096];
  __shared__ int err[257 * 33];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < 4096; i += 512) pos[i] = (unsigned short)(i & 255);
  for (int i = tid; i < 257 * 33; i += 512) err[i] = (i * 7) & 1023;
  __syncthreads();
  if (wid == 0) {
    int cand = lane, acc = 0;
    long long t0 = clock64();
    for (int g = 0; g < iters; ++g) {
      if (mode == 0) {            // shfl only
        cand = __shfl_sync(0xffffffffu, cand, g & 31) + 1;
      } else if (mode == 1) {     // shfl + syncwarp
        cand = __shfl_sync(0xffffffffu, cand, g & 31) + 1;
        __syncwarp();
      } else if (mode == 2) {     // shfl + predicated STS + syncwarp + dependent LDS
        int uid = __shfl_sync(0xffffffffu, cand, g & 31) & 255;
        if (lane == (g & 31)) pos[16 + (g & 31)] = (unsigned short)uid;
        __syncwarp();
        int e = err[uid * 33 + lane];
        acc = min(acc + e, 100000);
        cand = pos[(acc + lane) & 4095];
      } else if (mode == 3) {     // same as 2 without syncwarp
        int uid = __shfl_sync(0xffffffffu, cand, g & 31) & 255;
        if (lane == (g & 31)) pos[16 + (g & 31)] = (unsigned short)uid;
        int e = err[uid * 33 + lane];
        acc = min(acc + e, 100000);
        cand = pos[(acc + lane) & 4095];
      } else if (mode == 4) {     // dependent ALU chain of 40 ops, no sync
#pragma unroll
        for (int q = 0; q < 40; ++q) acc = (acc ^ (acc >> 3)) + q;
        cand = acc;
      }
    }
    long long t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
    if (cand == 123456789 || acc == 987654321) out[1] = 1;
  }
  __syncthreads();
}

int main() {
  long long *out;
  cudaMallocManaged(&out, 16);
  const char *names[] = {"shfl", "shfl+syncwarp", "shfl+STS+syncwarp+LDS+LDS", "shfl+STS+LDS+LDS (no syncwarp)", "40 dependent ALU ops"};
  for (int mode = 0; mode < 5; ++mode) {
    k<<<1, 512>>>(100, out, mode);
    cudaDeviceSynchronize();
    k<<<1, 512>>>(4096, out, mode);
    cudaDeviceSynchronize();
    printf("%-34s %7.1f cycles/iteration\n", names[mode], (double)out[0] / 4096);
  }
  return 0;
}
