// Issue rate of FFMA vs FFMA2 (packed fp32) on one GPU: independent chains, 16 warps per SM (the batch kernel's occupancy)
// and 64 warps per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k_ffma(float* out, float a, float b, int iters) {
  float x[CH];
  for (int i = 0; i < CH; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = fmaf(x[i], a, b);
  float s = 0; for (int i = 0; i < CH; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_ffma2(float* out, float a, float b, int iters) {
  float2 x[CH];
  for (int i = 0; i < CH; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = __ffma2_rn(x[i], a2, b2);
  float s = 0; for (int i = 0; i < CH; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// the batch kernel's inner loop with every operand in registers: 16 sets x (3 dependent FFMA2 + FMUL2), per-set dx, the
// slot's four coefficient pairs shared by the sets -- what the FMA pipe can do for that operand pattern without any
// shared-memory traffic, barrier or branch
__global__ void k_horner2(float* out, float a, float b, int iters) {
  float2 W[16], dx[16];
  for (int i = 0; i < 16; ++i) { W[i] = make_float2(1.f, 1.f); dx[i] = make_float2(threadIdx.x * 1e-6f + i * 1e-4f, threadIdx.x * 2e-6f + i * 1e-4f); }
  float2 y = make_float2(1.f, 1.f), bb = make_float2(a, b), c = make_float2(b, a), d = make_float2(a * b, a - b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float2 t = __ffma2_rn(dx[i], d, c);
      t = __ffma2_rn(dx[i], t, bb);
      t = __ffma2_rn(dx[i], t, y);
      W[i] = __fmul2_rn(W[i], t);
    }
    y.x += 1e-9f; d.y -= 1e-9f;      // a new "slot": the coefficients change, the chains stay
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += W[i].x + W[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_horner1(float* out, float a, float b, int iters) {
  float W0[16], W1[16], dx[16];
  for (int i = 0; i < 16; ++i) { W0[i] = 1.f; W1[i] = 1.f; dx[i] = threadIdx.x * 1e-6f + i * 1e-4f; }
  float y0 = 1.f, b0 = a, c0 = b, d0 = a * b, y1 = 1.f, b1 = b, c1 = a, d1 = a - b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      W0[i] *= fmaf(dx[i], fmaf(dx[i], fmaf(dx[i], d0, c0), b0), y0);
      W1[i] *= fmaf(dx[i], fmaf(dx[i], fmaf(dx[i], d1, c1), b1), y1);
    }
    y0 += 1e-9f; d1 -= 1e-9f;
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += W0[i] + W1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 2048 * 4);
  const int iters = 20000;
  for (int threads : {512, 1024}) {
    for (int bps : {1, 2}) {
      if (threads * bps > 2048) continue;
      const int grid = 148 * bps;
      float m1 = time_ms([&] { k_ffma<8><<<grid, threads>>>(out, 1.0001f, 0.5f, iters); });
      float m2 = time_ms([&] { k_ffma2<8><<<grid, threads>>>(out, 1.0001f, 0.5f, iters); });
      float m3 = time_ms([&] { k_ffma<2><<<grid, threads>>>(out, 1.0001f, 0.5f, iters); });
      float m4 = time_ms([&] { k_ffma2<2><<<grid, threads>>>(out, 1.0001f, 0.5f, iters); });
      const double n = double(grid) * threads * iters;
      printf("threads/SM %4d: FFMA x8 chains %.2f T lane-fma/s | FFMA2 x8 %.2f T lane-fma/s (%.2f T instr-lanes/s) | FFMA x2 %.2f | FFMA2 x2 %.2f T lane-fma/s\n",
             threads * bps, n * 8 / m1 * 1e-9, n * 16 / m2 * 1e-9, n * 8 / m2 * 1e-9, n * 2 / m3 * 1e-9, n * 4 / m4 * 1e-9);
    }
  }
  {
    const int it2 = 4000;
    float m5 = time_ms([&] { k_horner2<<<148, 512>>>(out, 1.0001f, 0.5f, it2); });
    float m6 = time_ms([&] { k_horner1<<<148, 512>>>(out, 1.0001f, 0.5f, it2); });
    const double n = double(148) * 512 * it2 * 16 * 2 * 4;      // lane-ops: 16 sets x 2 events x (3 fma + 1 mul)
    printf("batch inner loop in registers, 16 warps/SM: FFMA2 form %.2f T lane-op/s | scalar FFMA form %.2f T lane-op/s\n", n / m5 * 1e-9, n / m6 * 1e-9);
  }
  return 0;
}
