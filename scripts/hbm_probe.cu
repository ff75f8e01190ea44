// hbm_probe.cu -- what read bandwidth does a B200 give for the fill kernel's access pattern?
// Reads `chunk`-byte contiguous pieces, one out of every `stride_chunks`, through (a) a TMA bulk-copy
// ring (one producer thread per block) or (b) plain 128-bit LDG, and prints GB/s.  Measurement tool
// only (not part of the product); build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/hbm_probe scripts/hbm_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}

// chunk c (0..n_chunks) lives at base + perm(c) * stride_bytes.  Block b takes chunks b, b+grid, ...
// TMA: ring of `slots` buffers of `chunk` bytes; thread 0 waits for slot completion then re-issues.
__global__ void tma_probe(const unsigned char* base, long long n_chunks, long long stride_bytes, int chunk, int slots,
                          int copy_bytes, unsigned long long* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[64];
  if (threadIdx.x == 0) for (int s = 0; s < slots; ++s) mbar_init(&bar[s], 1);
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int per = chunk / copy_bytes;     // copies per chunk (to test copy granularity at fixed locality)
  long long c = blockIdx.x;
  int issued = 0;
  // prologue
  for (int s = 0; s < slots && c < n_chunks; ++s, c += gridDim.x, ++issued) {
    mbar_expect_tx(&bar[s], chunk);
    for (int k = 0; k < per; ++k)
      bulk_g2s(smem + (size_t)s * chunk + (size_t)k * copy_bytes, base + c * stride_bytes + (size_t)k * copy_bytes, copy_bytes, &bar[s]);
  }
  int s = 0; uint32_t phase = 0;
  unsigned long long acc = 0;
  for (int done = 0; done < issued; ) {
    mbar_wait(&bar[s], phase);
    acc += smem[(size_t)s * chunk];
    ++done;
    if (c < n_chunks) {
      mbar_expect_tx(&bar[s], chunk);
      for (int k = 0; k < per; ++k)
        bulk_g2s(smem + (size_t)s * chunk + (size_t)k * copy_bytes, base + c * stride_bytes + (size_t)k * copy_bytes, copy_bytes, &bar[s]);
      c += gridDim.x; ++issued;
    }
    if (++s == slots) { s = 0; phase ^= 1; }
  }
  if (acc == 0xdeadbeefULL) *sink = acc;
}

// LDG: each block handles chunks b, b+grid, ...; threads read float4 contiguous, `unroll` loads in flight
template <int U>
__global__ void ldg_probe(const unsigned char* base, long long n_chunks, long long stride_bytes, int chunk, float* sink) {
  float acc = 0.f;
  const int per_iter = blockDim.x * 16 * U;
  for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const float4* p = reinterpret_cast<const float4*>(base + c * stride_bytes);
    for (int off = 0; off < chunk; off += per_iter) {
      float4 v[U];
      #pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = off / 16 + u * blockDim.x + threadIdx.x;
        v[u] = (idx * 16 < chunk) ? __ldcs(p + idx) : make_float4(0, 0, 0, 0);
      }
      #pragma unroll
      for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
  }
  if (acc == 1234.5f) *sink = acc;
}

int main(int argc, char** argv) {
  const size_t total = (size_t)4 << 30;    // 4 GiB buffer
  unsigned char* d; CK(cudaMalloc(&d, total)); CK(cudaMemset(d, 1, total));
  unsigned long long* sink; CK(cudaMalloc(&sink, 8));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(tma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int sms = 148;
  printf("mode chunk copy stride_chunks slots/unroll blocks GB/s\n");
  const int chunks[] = {2048, 4096, 8192, 16384, 32768};
  for (int stride_chunks : {1, 7}) {
    for (int chunk : chunks) {
      const long long stride = (long long)chunk * stride_chunks;
      const long long n_chunks = (long long)(total / stride);
      const double bytes = (double)n_chunks * chunk;
      // TMA: in-flight = slots*chunk; test 64 KB and 160 KB rings, copy granularity = chunk and 2 KB
      for (int ring_kb : {64, 160}) {
        int slots = ring_kb * 1024 / chunk; if (slots > 64) slots = 64; if (slots < 2) slots = 2;
        for (int copy : {chunk, 2048}) {
          if (copy > chunk) continue;
          for (int bps : {1, 2}) {
            if (bps * slots * chunk > 200 * 1024) continue;
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
              CK(cudaEventRecord(e0));
              tma_probe<<<sms * bps, 32, (size_t)slots * chunk>>>(d, n_chunks, stride, chunk, slots, copy, sink);
              CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
              float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
            }
            printf("tma %d %d %d %d %d %.0f\n", chunk, copy, stride_chunks, slots, sms * bps, bytes / best / 1e6);
          }
        }
      }
      for (int bps : {4, 8}) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
          CK(cudaEventRecord(e0));
          ldg_probe<8><<<sms * bps, 256>>>(d, n_chunks, stride, chunk, reinterpret_cast<float*>(sink));
          CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
          float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        printf("ldg %d - %d 8 %d %.0f\n", chunk, stride_chunks, sms * bps, bytes / best / 1e6);
      }
    }
  }
  // reference: device-to-device copy (read+write bytes), the MEASURED_PEAKS definition
  {
    unsigned char* d2; CK(cudaMalloc(&d2, (size_t)2 << 30));
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(d2, d, (size_t)2 << 30, cudaMemcpyDeviceToDevice));
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    printf("memcpy_d2d 2GiB read+write GB/s %.0f\n", 2.0 * ((size_t)2 << 30) / best / 1e6);
  }
  return 0;
}
