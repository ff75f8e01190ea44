// m3b_device.cuh -- device-side pieces shared by the fill kernels: PTX helpers (mbarrier, bulk/TMA
// copies, streaming loads), the test statistics of SampleHandlerBase::GetTestStatLLH, the block-level
// likelihood reduction and the common block epilogue (flush -> ticket -> publish to peers | fused -lnL).
#pragma once
#include "m3b_internal.h"
#include <cuda_runtime.h>
#include <math.h>

namespace m3b {

// ------------------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + 1-D bulk (TMA) global->shared copy, streaming vector loads
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
// coefficient rows are read once per step: keep them out of L1, default L2 policy (small
// workloads stay L2-resident between steps, large ones stream)
// Stage the per-step table into shared memory: from the kernel parameters (constant bank) when the
// host inlined it, else with one bulk copy from device memory.  Every thread calls it; `bar` must be
// initialised (count 1) and visible.  Returns after the table is readable by the whole block.
__device__ __forceinline__ void stage_step_table(const FillArgs& a, unsigned char* st, uint64_t* bar) {
  if (a.step_inline_bytes > 0) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.step_inline);
    uint32_t* dst = reinterpret_cast<uint32_t*>(st);
    for (int i = threadIdx.x; i < a.step_inline_bytes / 4; i += blockDim.x) dst[i] = src[i];
  } else {
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, static_cast<uint32_t>(a.step.bytes));
      bulk_g2s(st, a.step_table, static_cast<uint32_t>(a.step.bytes), bar);
    }
    mbar_wait(bar, 0);
  }
  __syncthreads();
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// trace slots: 0 block start, 1 tables staged, 2 first stage consumed, 3 producer out of work,
//              4 consumers done, 5 histogram flushed, 6 block end, 7 units processed by the block
__device__ __forceinline__ void trace_mark(const FillArgs& a, int slot) {
  if (a.trace) a.trace[static_cast<size_t>(blockIdx.x) * 8 + slot] = globaltimer_ns();
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ldg_stream(const float2* p) {
  float2 v;
  asm("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}

// ------------------------------------------------------------------------------------------------
// test statistics, SampleHandlerBase::GetTestStatLLH (Samples/SampleHandlerBase.cpp:35-192), f64
// ------------------------------------------------------------------------------------------------
constexpr double kLowMcBound = .00001;   // M3::_LOW_MC_BOUND_ (Manager/Core.h:83)

__device__ __forceinline__ double poisson_llh(double data, double mc) {      // :17-31
  if (data == 0) return mc;
  if (mc < kLowMcBound) {
    if (data > kLowMcBound) return (kLowMcBound - data + data * log(data / kLowMcBound));
    else if (data >= mc) return 0.;
  }
  return (mc - data + data * log(data / mc));
}

// `thrown` is set where the reference throws MaCh3Exception instead of returning a number (the Barlow-Beeston negative
// discriminant, :64-67); the kernels collect it in the handle's status word and m3b_llh turns it into M3B_ERR_MATH.
constexpr int kStatusPeerTimeout = 1, kStatusMathError = 2;
static __device__ double test_stat_llh(int ts, double data, double mc, double w2, bool& thrown) {
  switch (ts) {
    case 1: {   // kBarlowBeeston :46-88
      double newmc = mc;
      if (mc < kLowMcBound) {
        if (data > kLowMcBound) newmc = kLowMcBound;
        else if (data >= mc) return 0.;
      }
      const double fractional = sqrt(w2) / newmc;
      const double fractional2 = fractional * fractional;
      const double temp = newmc * fractional2 - 1;
      const double temp2 = temp * temp + 4 * data * fractional2;
      if (temp2 < 0) { thrown = true; return nan(""); }          // the reference throws here
      const double beta = (-1 * temp + sqrt(temp2)) / 2.;
      double stat = mc * beta;
      if (data > 0) {
        newmc *= beta;
        stat = newmc - data + data * log(data / newmc);
      }
      double penalty = 0;
      if (fractional > 0) penalty = (beta - 1) * (beta - 1) / (2 * fractional2);
      return stat + penalty;
    }
    case 4: {   // kDembinskiAbdelmotteleb :90-126
      if (w2 == 0) return poisson_llh(data, mc);
      double newmc = mc;
      if (mc < kLowMcBound) {
        if (data > kLowMcBound) newmc = kLowMcBound;
        else if (data >= mc) return 0.;
      }
      const double k = newmc * newmc / w2;
      const double beta = (data + k) / (newmc + k);
      newmc *= beta;
      const double penalty = k * beta - k + k * log(k / (k * beta));
      double stat = newmc;
      if (data > 0) stat = newmc - data + data * log(data / newmc);
      return stat + penalty;
    }
    case 2: {   // kIceCube :133-160 (the reference evaluates in long double; f64 here)
      if (w2 == 0) return poisson_llh(data, mc);
      const double b = mc / w2;
      const double a = mc * b + 1;
      const double stat = -1 * (a * log(b) + lgamma(data + a) - lgamma(data + 1) - ((data + a) * log1p(b)) - lgamma(a));
      if (mc <= data) {
        if (data <= kLowMcBound) return 0.;
        const double poisson = poisson_llh(data, kLowMcBound);
        if (stat > poisson) return poisson;
      }
      return stat;
    }
    case 3: {   // kPearson :162-177
      if (data == 0) return mc / 2.;
      if (mc < kLowMcBound) {
        if (data > kLowMcBound) return (data - kLowMcBound) * (data - kLowMcBound) / (2. * kLowMcBound);
        else if (data >= mc) return 0.;
      }
      return (data - mc) * (data - mc) / (2 * mc);
    }
    default:    // kPoisson :178-184
      return poisson_llh(data, mc);
  }
}

__device__ __forceinline__ double warp_sum(double v) {
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Block-level likelihood.  For every sample all threads stride over the sample's bins
// [start,end), then shuffle-tree per warp; thread s finally adds the warps' partial sums of
// sample s in warp order and thread 0 adds the samples in sample order.  The summation shape
// is fixed, so the result is a deterministic function of the histogram.
// scratch: n_samples * 32 doubles of shared memory.
constexpr int kMaxSamples = 64;

static __device__ void block_llh(const double* __restrict__ hist, const double* __restrict__ w2,
                          const double* __restrict__ data, const int32_t* __restrict__ sample_start,
                          int n_samples, int ts, double* llh_dev, double* llh_host, double* scratch, int32_t* status,
                          unsigned long long* seq_host = nullptr, unsigned long long seq = 0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  bool thrown = false;
  for (int s = 0; s < n_samples; ++s) {
    const int b0 = sample_start[s], b1 = sample_start[s + 1];
    double acc = 0.;
    // loads of four strides first, then the arithmetic: one memory latency per batch instead of one
    // per bin (data[] has been flushed out of L2 by the coefficient stream); same summation order
    const int NT = blockDim.x;
    for (int b = b0 + threadIdx.x; b < b1; b += 4 * NT) {
      double d[4], m[4], v[4];
      #pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int bb = b + k * NT;
        const bool ok = bb < b1;
        d[k] = ok ? data[bb] : 0.;
        m[k] = ok ? __ldcg(hist + bb) : 0.;
        v[k] = (ok && w2) ? __ldcg(w2 + bb) : 0.;
      }
      #pragma unroll
      for (int k = 0; k < 4; ++k)
        if (b + k * NT < b1) acc += test_stat_llh(ts, d[k], m[k], v[k], thrown);
    }
    acc = warp_sum(acc);
    if (lane == 0) scratch[s * 32 + warp] = acc;
  }
  if (thrown && status) atomicOr(status, kStatusMathError);
  __syncthreads();
  if (threadIdx.x < n_samples) {
    double tot = 0.;
    for (int w = 0; w < nwarps; ++w) tot += scratch[threadIdx.x * 32 + w];
    scratch[threadIdx.x * 32] = tot;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.;
    for (int s = 0; s < n_samples; ++s) {
      const double v = scratch[s * 32];
      tot += v;
      llh_dev[1 + s] = v;
      if (llh_host) llh_host[1 + s] = v;
    }
    llh_dev[0] = tot;
    if (llh_host) llh_host[0] = tot;     // mapped host memory
    if (llh_host && seq_host) {          // publish: the host polls this word instead of waiting for the kernel to retire
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long*>(seq_host) = seq;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// common epilogue of a fill block: flush the privatised histogram (block -> grid), take a ticket;
// the last block of the grid either pushes the partial histogram to the peers (multi-GPU, own
// exchange) or reduces the likelihood and re-arms the state for the next step.
// Every thread of the block must call it.  `scratch` = n_samples*32 doubles of shared memory that
// no longer hold live data.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void finish_block(const FillArgs& a, const double* s_hist, const double* s_w2,
                                             double* scratch, int* s_last) {
  const int tid = threadIdx.x, NT = blockDim.x;
  const bool w2_live = a.w2 != nullptr;
  // queued steps (programmatic dependent launch): the histogram buffers, the ticket and the scheduler counter were
  // last written by the previous step's final block -- make sure that launch has completed before touching them
  if (a.pdl) grid_dependency_wait();
  if (a.fuse_llh) {
    // the last block will need data[] (and a frozen w2[]) which the coefficient stream has pushed out
    // of L2 by now: every block pulls its slice back in while it flushes
    const int lines = (a.n_bins * 8 + 127) / 128;
    for (int l = blockIdx.x * NT + tid; l < lines; l += gridDim.x * NT) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(a.data) + 128 * l));
      if (a.w2_frozen && a.w2_frozen != a.w2)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(a.w2_frozen) + 128 * l));
    }
  }
  if (a.hist_in_smem && !a.weights_only) {
    __syncthreads();
    for (int i = tid; i < a.n_bins; i += NT) {
      const double v = s_hist[i];
      if (v != 0.) atomicAdd(a.hist + i, v);
    }
    if (w2_live)
      for (int i = tid; i < a.n_bins; i += NT) {
        const double v = s_w2[i];
        if (v != 0.) atomicAdd(a.w2 + i, v);
      }
  }
  if (tid == 0) trace_mark(a, 5);
  // (weights_only: SMonolithGPU::RunGPU_SplineMonolith contract -- the weights are the output)
  if ((a.weights_only || (!a.fuse_llh && a.peer_world == 0)) && !a.tile_counter) return;

  // last-block-done ticket
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int tk = atomicAdd(a.ticket, 1u);
    *s_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (!*s_last) return;
  __threadfence();
  if (tid == 0 && a.tile_counter) *a.tile_counter = 0u;     // dynamic tile scheduler: re-arm

  if (a.weights_only) { if (tid == 0) *a.ticket = 0u; return; }
  if (a.peer_world > 0) {
    // the whole grid has flushed into this rank's exported buffer: publish the epoch; the peers pull
    if (tid == 0) {
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer_flag_own), "r"(a.peer_epoch) : "memory");
      *a.ticket = 0u;
    }
    return;
  }
  if (!a.fuse_llh) { if (tid == 0) *a.ticket = 0u; return; }

  if (tid == 0 && a.trace) a.trace[8 * 4000 + 0] = globaltimer_ns();      // last block: ticket won
  block_llh(a.hist, a.w2_frozen, a.data, a.sample_start_inline, a.n_samples, a.test_stat, a.llh_dev, a.llh_host, scratch, a.status,
            a.llh_seq_host, a.llh_seq);
  if (tid == 0 && a.trace) a.trace[8 * 4000 + 1] = globaltimer_ns();      // last block: -lnL written
  // prepare the next step: zero its histogram(s), re-arm the ticket
  if (a.hist_next) for (int i = tid; i < a.n_bins; i += NT) a.hist_next[i] = 0.;
  if (a.w2_next) for (int i = tid; i < a.n_bins; i += NT) a.w2_next[i] = 0.;
  if (tid == 0) *a.ticket = 0u;
}

}  // namespace m3b
