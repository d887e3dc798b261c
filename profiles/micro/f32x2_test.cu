// Micro-test: (1) are __fmul2_rn + __fadd2_rn bit-identical to scalar __fmul_rn + __fadd_rn
// (i.e. does ptxas keep them un-fused)?  (2) issue throughput of FADD / FMUL / FADD2 / FMUL2.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void check(const float *p, const float *acc0, int n, unsigned *mismatch, unsigned *fused_like) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float w23 = 2.0f / 3.0f, w13 = 1.0f / 3.0f;
  float a = acc0[i], b = acc0[i] * 0.5f;
  float P = p[i];
  float2 acc = make_float2(a, b);
  float2 r2 = __fadd2_rn(acc, __fmul2_rn(make_float2(P, P), make_float2(w23, w13)));
  float s0 = __fadd_rn(a, __fmul_rn(P, w23));
  float s1 = __fadd_rn(b, __fmul_rn(P, w13));
  float f0 = __fmaf_rn(P, w23, a), f1 = __fmaf_rn(P, w13, b);
  if (r2.x != s0 || r2.y != s1) atomicAdd(mismatch, 1u);
  if ((s0 != f0 || s1 != f1) && r2.x == f0 && r2.y == f1) atomicAdd(fused_like, 1u);
}

template <int MODE>
__global__ void tput(float *out, int iters) {
  float2 a0 = make_float2(threadIdx.x, 1.f), a1 = make_float2(2.f, threadIdx.x), a2 = make_float2(3.f, 4.f), a3 = make_float2(5.f, 6.f);
  float2 a4 = a0, a5 = a1, a6 = a2, a7 = a3;
  const float2 w = make_float2(1.0000001f, 0.9999999f);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0) {  // scalar FADD x2 per pair
        a0.x = __fadd_rn(a0.x, w.x); a0.y = __fadd_rn(a0.y, w.y); a1.x = __fadd_rn(a1.x, w.x); a1.y = __fadd_rn(a1.y, w.y);
        a2.x = __fadd_rn(a2.x, w.x); a2.y = __fadd_rn(a2.y, w.y); a3.x = __fadd_rn(a3.x, w.x); a3.y = __fadd_rn(a3.y, w.y);
        a4.x = __fadd_rn(a4.x, w.x); a4.y = __fadd_rn(a4.y, w.y); a5.x = __fadd_rn(a5.x, w.x); a5.y = __fadd_rn(a5.y, w.y);
        a6.x = __fadd_rn(a6.x, w.x); a6.y = __fadd_rn(a6.y, w.y); a7.x = __fadd_rn(a7.x, w.x); a7.y = __fadd_rn(a7.y, w.y);
      } else if (MODE == 1) {  // FADD2
        a0 = __fadd2_rn(a0, w); a1 = __fadd2_rn(a1, w); a2 = __fadd2_rn(a2, w); a3 = __fadd2_rn(a3, w);
        a4 = __fadd2_rn(a4, w); a5 = __fadd2_rn(a5, w); a6 = __fadd2_rn(a6, w); a7 = __fadd2_rn(a7, w);
      } else if (MODE == 2) {  // FMUL2
        a0 = __fmul2_rn(a0, w); a1 = __fmul2_rn(a1, w); a2 = __fmul2_rn(a2, w); a3 = __fmul2_rn(a3, w);
        a4 = __fmul2_rn(a4, w); a5 = __fmul2_rn(a5, w); a6 = __fmul2_rn(a6, w); a7 = __fmul2_rn(a7, w);
      } else if (MODE == 3) {  // scalar FMUL
        a0.x = __fmul_rn(a0.x, w.x); a0.y = __fmul_rn(a0.y, w.y); a1.x = __fmul_rn(a1.x, w.x); a1.y = __fmul_rn(a1.y, w.y);
        a2.x = __fmul_rn(a2.x, w.x); a2.y = __fmul_rn(a2.y, w.y); a3.x = __fmul_rn(a3.x, w.x); a3.y = __fmul_rn(a3.y, w.y);
        a4.x = __fmul_rn(a4.x, w.x); a4.y = __fmul_rn(a4.y, w.y); a5.x = __fmul_rn(a5.x, w.x); a5.y = __fmul_rn(a5.y, w.y);
        a6.x = __fmul_rn(a6.x, w.x); a6.y = __fmul_rn(a6.y, w.y); a7.x = __fmul_rn(a7.x, w.x); a7.y = __fmul_rn(a7.y, w.y);
      } else if (MODE == 4) {  // integer LOP3/IADD mix (ALU pipe)
        unsigned *q = reinterpret_cast<unsigned *>(&a0);
        unsigned x0 = q[0], x1 = q[1];
        x0 = (x0 ^ 0x9E3779B1u) + x1; x1 = (x1 & 0x7fffffffu) | x0; x0 = x0 + x1 * 1u; x1 ^= x0 >> 3;
        x0 = (x0 ^ 0x9E3779B1u) + x1; x1 = (x1 & 0x7fffffffu) | x0; x0 = x0 + x1 * 1u; x1 ^= x0 >> 3;
        q[0] = x0; q[1] = x1;
      } else if (MODE == 5) {  // vabsdiff4 + dp4a
        unsigned *q = reinterpret_cast<unsigned *>(&a0);
        unsigned *r = reinterpret_cast<unsigned *>(&a1);
        unsigned d = __vabsdiffu4(q[0], r[0]); q[1] = __dp4a(d, d, q[1]);
        d = __vabsdiffu4(q[1], r[1]); q[0] = __dp4a(d, d, q[0]);
        d = __vabsdiffu4(r[0], q[0]); r[1] = __dp4a(d, d, r[1]);
        d = __vabsdiffu4(r[1], q[1]); r[0] = __dp4a(d, d, r[0]);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0.x + a0.y + a1.x + a1.y + a2.x + a2.y + a3.x + a3.y + a4.x + a4.y + a5.x + a5.y + a6.x + a6.y + a7.x + a7.y;
}

template <int MODE>
float run(float *out, int iters, int blocks, int threads) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  tput<MODE><<<blocks, threads>>>(out, 10);
  cudaEventRecord(a);
  tput<MODE><<<blocks, threads>>>(out, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  const int n = 1 << 20;
  float *p, *acc; unsigned *cnt;
  cudaMallocManaged(&p, n * 4); cudaMallocManaged(&acc, n * 4); cudaMallocManaged(&cnt, 8);
  unsigned s = 12345;
  for (int i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; p[i] = (float)((s >> 8) & 255); s = s * 1664525u + 1013904223u; acc[i] = (float)(s >> 12) / 4096.0f * 0.37f; }
  cnt[0] = cnt[1] = 0;
  check<<<n / 256, 256>>>(p, acc, n, cnt, cnt + 1);
  cudaDeviceSynchronize();
  printf("f32x2 vs scalar mismatches: %u of %d pairs (fused-like: %u)\n", cnt[0], n, cnt[1]);
  float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000, blocks = 148 * 8, threads = 256;
  const char *names[] = {"FADD  (16 scalar/iter-unroll)", "FADD2 (8 packed)", "FMUL2 (8 packed)", "FMUL  (16 scalar)", "INT mix", "VABSDIFF4+DP4A"};
  float ms[6] = {run<0>(out, iters, blocks, threads), run<1>(out, iters, blocks, threads), run<2>(out, iters, blocks, threads),
                 run<3>(out, iters, blocks, threads), run<4>(out, iters, blocks, threads), run<5>(out, iters, blocks, threads)};
  int prop_clock; cudaDeviceGetAttribute(&prop_clock, cudaDevAttrClockRate, 0);
  for (int m = 0; m < 6; ++m) {
    double flops_pairs = (double)blocks * threads * iters * 8 * 8;  // float2-equivalents (m<4)
    printf("%-32s %8.3f ms   %.1f G pair-ops/s\n", names[m], ms[m], flops_pairs / ms[m] / 1e6);
  }
  return cnt[0] != 0;
}
