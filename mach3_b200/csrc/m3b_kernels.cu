// m3b_kernels.cu -- hand-written sm_100a kernels of the MaCh3 per-step likelihood hot path.
//
//   (the fused per-step kernel itself is fill_tma_kernel, m3b_fill_tma.cu; its register-streaming
//    predecessor fill_kernel<> below is compiled only with -DM3B_EXPERIMENTS)
//   llh_kernel    the likelihood reduction alone (after a multi-GPU histogram exchange)
//   bin_kernel    BinningHandler::FindGlobalBin (Samples/BinningHandler.cpp:257-277) per event
//   retile_*      AoS reference monolith -> tiled SoA device layout (setup time)
//
// No tensor cores: the path is a bandwidth-bound gather/evaluate/scatter.
#include "m3b_device.cuh"

namespace m3b {

#ifdef M3B_EXPERIMENTS   // the first design (register-streaming loads), kept out of the product build: A/B experiments only
// ------------------------------------------------------------------------------------------------
// the fused per-step kernel
// ------------------------------------------------------------------------------------------------
// T = events per tile = threads per block; kBatch = coefficient loads in flight per thread;
// kMinBlocks = blocks per SM the register allocation is capped for
template <int T, int kBatch, int kMinBlocks>
__global__ void __launch_bounds__(T, kMinBlocks) fill_kernel(const __grid_constant__ FillArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int s_last;

  const int tid = threadIdx.x;
  unsigned char* st = smem;
  int32_t* s_row = reinterpret_cast<int32_t*>(smem + a.step.bytes);
  float* s_dx = reinterpret_cast<float*>(s_row + a.max_nc);
  float* s_lv = s_dx + a.max_nc;
  double* s_hist = reinterpret_cast<double*>(
      smem + ((a.step.bytes + 4 * (2 * a.max_nc + a.max_nl) + 15) & ~15));
  double* s_w2 = s_hist + a.n_bins;
  const bool w2_live = a.w2 != nullptr;
  const bool smem_hist = a.hist_in_smem != 0;

  // stage the per-step {segment, dx, value, norm} table with one bulk (TMA) copy
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (smem_hist && !a.weights_only) {
    for (int i = tid; i < a.n_bins; i += T) s_hist[i] = 0.;
    if (w2_live) for (int i = tid; i < a.n_bins; i += T) s_w2[i] = 0.;
  }
  stage_step_table(a, st, &bar);

  const int32_t* seg = reinterpret_cast<const int32_t*>(st + a.step.off_seg);
  const float* dxp = reinterpret_cast<const float*>(st + a.step.off_dx);
  const float* val = reinterpret_cast<const float*>(st + a.step.off_val);
  const float* norm = reinterpret_cast<const float*>(st + a.step.off_norm);

  int cur_sig = -1, nc = 0, nl = 0;
  for (int t = a.tile_begin + blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const TileDesc td = a.tiles[t];
    if (td.sig != cur_sig) {            // block-uniform
      __syncthreads();
      const SigDesc sd = a.sigs[td.sig];
      nc = sd.nc; nl = sd.nl;
      const int32_t* pool = a.sig_pool + sd.off;
      for (int s = tid; s < nc; s += T) {
        const int p = pool[s];
        s_row[s] = (pool[nc + s] + seg[p]) * T;
        s_dx[s] = dxp[p];
      }
      for (int s = tid; s < nl; s += T) s_lv[s] = val[pool[2 * nc + s]];
      cur_sig = td.sig;
      __syncthreads();
    }
    const int64_t e = static_cast<int64_t>(t) * T + tid;

    // event-table loads first so they fly together with the coefficient stream
    const int bin = a.bin[e];
    float w_osc = 1.f, w_static = 1.f;
    if (a.osc) {
      const int64_t oi = a.osc_idx ? static_cast<int64_t>(a.osc_idx[e]) : (e < a.n_events ? e : 0);
      w_osc = oi >= 0 ? a.osc[oi] : 1.f;
    }
    if (a.static_w) w_static = a.static_w[e];
    // CalcWeightTotal, first factor: norms (double -> float on the host), reference order
    float w = 1.0f;
    #pragma unroll 4
    for (int j = 0; j < a.norm_slots; ++j) {
      const int i = a.norm_idx[static_cast<int64_t>(j) * a.e_pad + e];
      w *= (i >= 0 ? norm[i] : 1.0f);
    }

    // CalcSplineWeights + CalcTotalEventWeight: cubic responses in slot (= parameter) order, then
    // TF1.  Loads are issued kBatch at a time so every thread keeps kBatch*16 B in flight.
    float w_spl = 1.0f;
    const float4* cub = td.cub + tid;
    for (int s0 = 0; s0 < nc; s0 += kBatch) {
      float4 c[kBatch];
      #pragma unroll
      for (int j = 0; j < kBatch; ++j)
        if (s0 + j < nc) c[j] = ldg_stream(cub + s_row[s0 + j]);
      #pragma unroll
      for (int j = 0; j < kBatch; ++j)
        if (s0 + j < nc) {
          const float dx = s_dx[s0 + j];
          w_spl *= fmaf(dx, fmaf(dx, fmaf(dx, c[j].w, c[j].z), c[j].y), c[j].x);
        }
    }
    const float2* lin = td.lin + tid;
    for (int s0 = 0; s0 < nl; s0 += kBatch) {
      float2 c[kBatch];
      #pragma unroll
      for (int j = 0; j < kBatch; ++j)
        if (s0 + j < nl) c[j] = ldg_stream(lin + (s0 + j) * T);
      #pragma unroll
      for (int j = 0; j < kBatch; ++j)
        if (s0 + j < nl) w_spl *= fmaf(c[j].x, s_lv[s0 + j], c[j].y);
    }

    // then osc, spline, extra -- the push order of total_weight_pointers
    w *= w_osc;
    w *= w_spl;
    w *= w_static;

    if (a.evt_spline_w && e < a.n_events) { a.evt_spline_w[e] = w_spl; a.evt_total_w[e] = w; }

    // FillArray_MP: skip w<=0 and under/overflow; mc += w; w2 += w*w (float product)
    if (w > 0.f && bin >= 0 && !a.weights_only) {
      if (smem_hist) {
        atomicAdd(s_hist + bin, static_cast<double>(w));
        if (w2_live) atomicAdd(s_w2 + bin, static_cast<double>(w * w));
      } else {
        atomicAdd(a.hist + bin, static_cast<double>(w));
        if (w2_live) atomicAdd(a.w2 + bin, static_cast<double>(w * w));
      }
    }
  }

  finish_block(a, s_hist, s_w2, reinterpret_cast<double*>(smem), &s_last);
}

int fill_smem_bytes(const FillArgs& a, bool hist_in_smem, bool w2_live) {
  int b = (a.step.bytes + 4 * (2 * a.max_nc + a.max_nl) + 15) & ~15;
  if (hist_in_smem) b += 8 * a.n_bins * (w2_live ? 2 : 1);
  const int llh_scratch = a.n_samples * 32 * 8;   // block_llh reuses the dynamic region
  return b > llh_scratch ? b : llh_scratch;
}

// kernel variants: {batch, min blocks/SM}; picked per tile size by fill_variant()
#define M3B_FOR_VARIANT(T_, V_, EXPR)                                                       \
  switch (V_) {                                                                             \
    case 0: { constexpr int B = 8, M = 1; auto k = fill_kernel<T_, B, M>; EXPR; } break;    \
    case 1: { constexpr int B = 8, M = (1024 / T_); auto k = fill_kernel<T_, B, M>; EXPR; } break;  \
    case 2: { constexpr int B = 4, M = (1280 / T_); auto k = fill_kernel<T_, B, M>; EXPR; } break;  \
    case 3: { constexpr int B = 4, M = (1536 / T_); auto k = fill_kernel<T_, B, M>; EXPR; } break;  \
    case 4: { constexpr int B = 16, M = (512 / T_); auto k = fill_kernel<T_, B, M>; EXPR; } break;  \
    case 5: { constexpr int B = 12, M = (768 / T_); auto k = fill_kernel<T_, B, M>; EXPR; } break;  \
    default: return cudaErrorInvalidValue;                                                  \
  }
#define M3B_FOR_TILE(T_RUNTIME, V_, EXPR)                                                   \
  switch (T_RUNTIME) {                                                                      \
    case 128: M3B_FOR_VARIANT(128, V_, EXPR) break;                                         \
    case 256: M3B_FOR_VARIANT(256, V_, EXPR) break;                                         \
    case 512: M3B_FOR_VARIANT(512, V_, EXPR) break;                                         \
    default: return cudaErrorInvalidValue;                                                  \
  }

cudaError_t launch_fill(const FillArgs& a, int variant, int grid, int smem, cudaStream_t s) {
  M3B_FOR_TILE(a.T, variant, (k<<<grid, a.T, smem, s>>>(a)))
  return cudaGetLastError();
}
cudaError_t fill_set_smem(int T, int variant, int smem) {
  (void)smem;
  M3B_FOR_TILE(T, variant, return allow_max_dynamic_smem(k))
  return cudaSuccess;
}
cudaError_t fill_occupancy(int T, int variant, int smem, int* bps) {
  M3B_FOR_TILE(T, variant, return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, k, T, smem))
  return cudaSuccess;
}

#endif  // M3B_EXPERIMENTS

// ------------------------------------------------------------------------------------------------
// likelihood alone (after an external all-reduce of the histogram)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) llh_kernel(const __grid_constant__ LlhArgs a) {
  __shared__ double s_part[kMaxSamples * 32];
  block_llh(a.hist, a.w2, a.data, a.sample_start, a.n_samples, a.test_stat, a.llh_dev, a.llh_host, s_part, a.status, a.llh_seq_host, a.llh_seq);
}

// ------------------------------------------------------------------------------------------------
// The library's own histogram exchange + likelihood, fused (multi-GPU, one process per GPU):
// every rank PULLS all ranks' partial histograms over NVLink peer memory (plain P2P loads on CUDA-IPC
// mapped pointers), sums them in rank order -- so all ranks hold bit-identical totals -- and reduces
// -lnL, block-then-grid: block b owns a contiguous slice of the bins, the last block to finish adds the
// per-block per-sample sums in block order.  Replaces ncclAllReduce + a separate likelihood launch.
// Waiting for the peers' epoch flags is bounded (~2 s); a timeout is reported, never a hang.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) llh_pull_kernel(const __grid_constant__ LlhArgs a) {
  __shared__ double s_part[kMaxSamples * 16];
  __shared__ int s_ok, s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) s_ok = 1;
  __syncthreads();
  if (tid < a.peer_world) {
    unsigned int v = 0;
    const long long t0 = clock64();
    while (true) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a.peer_flag[tid]) : "memory");
      if (static_cast<int>(v - a.epoch) >= 0) break;          // a peer may already be one step ahead
      if (clock64() - t0 > 4000000000LL) { s_ok = 0; break; }
      __nanosleep(100);
    }
  }
  __syncthreads();
  if (!s_ok) {
    if (tid == 0 && blockIdx.x == 0) {
      atomicOr(a.status, kStatusPeerTimeout); a.llh_dev[0] = nan("");
      if (a.llh_host) a.llh_host[0] = nan("");
      if (a.llh_host && a.llh_seq_host) { __threadfence_system(); *reinterpret_cast<volatile unsigned long long*>(a.llh_seq_host) = a.llh_seq; }
    }
    return;
  }
  const int per = (a.n_bins + gridDim.x - 1) / gridDim.x;
  const int b0 = blockIdx.x * per, b1 = min(a.n_bins, b0 + per);
  bool thrown = false;
  for (int s = 0; s < a.n_samples; ++s) {
    const int lo = max(b0, a.sample_start_inline[s]), hi = min(b1, a.sample_start_inline[s + 1]);
    double acc = 0.;
    for (int b = lo + tid; b < hi; b += blockDim.x) {
      double mc = 0., w2 = 0.;
      for (int r = 0; r < a.peer_world; ++r) mc += __ldcv(a.peer_hist[r] + b);          // fixed rank order
      a.hist_out[b] = mc;
      if (a.w2_live) {
        for (int r = 0; r < a.peer_world; ++r) w2 += __ldcv(a.peer_hist[r] + a.n_bins + b);
        a.w2_out[b] = w2;
      } else if (a.w2) {
        w2 = a.w2[b];
      }
      acc += test_stat_llh(a.test_stat, a.data[b], mc, w2, thrown);
    }
    acc = warp_sum(acc);
    if (lane == 0) s_part[s * 16 + warp] = acc;
  }
  if (thrown) atomicOr(a.status, kStatusMathError);
  __syncthreads();
  if (tid < a.n_samples) {
    double tot = 0.;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += s_part[tid * 16 + w];
    a.partial[blockIdx.x * a.n_samples + tid] = tot;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < a.n_samples) {
    double tot = 0.;
    for (int b = 0; b < static_cast<int>(gridDim.x); ++b) tot += __ldcg(a.partial + b * a.n_samples + tid);
    s_part[tid] = tot;
  }
  __syncthreads();
  if (tid == 0) {
    double tot = 0.;
    for (int s = 0; s < a.n_samples; ++s) {
      tot += s_part[s];
      a.llh_dev[1 + s] = s_part[s];
      if (a.llh_host) a.llh_host[1 + s] = s_part[s];
    }
    a.llh_dev[0] = tot;
    if (a.llh_host) a.llh_host[0] = tot;
    if (a.llh_host && a.llh_seq_host) {
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long*>(a.llh_seq_host) = a.llh_seq;
    }
    *a.ticket = 0u;
  }
}
cudaError_t launch_llh_pull(const LlhArgs& a, int blocks, cudaStream_t s) {
  llh_pull_kernel<<<blocks, 512, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_llh(const LlhArgs& a, cudaStream_t s) {
  llh_kernel<<<1, 1024, 0, s>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// BinningHandler::FindGlobalBin, uniform arm.  SampleBinningInfo::FindBin's nominal-bin and
// adjacent-bin shortcuts (Samples/SampleStructs.h:588-609) return exactly upper_bound(edges,x)-1
// whenever they fire, so the stateless form below is bit-identical to the reference.
// ------------------------------------------------------------------------------------------------
__global__ void bin_kernel(const __grid_constant__ BinArgs a) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= a.e_pad) return;
  if (e >= a.n_events) { a.bin[e] = -1; return; }
  const int s = a.sample_id[e];
  const int nd = a.n_dim[s];
  int g = 0;
  bool ok = true;
  for (int d = 0; d < nd; ++d) {
    const double x = a.kin[static_cast<int64_t>(d) * a.n_events + e];
    const int nb = a.nbins[s * kMaxDim + d];
    const double* ed = a.edges + a.edge_off[s * kMaxDim + d];
    if (x < ed[0] || x >= ed[nb]) { ok = false; break; }       // SampleStructs.h:584
    int lo = 0, len = nb + 1;                                    // std::upper_bound
    while (len > 0) {
      const int half = len >> 1;
      if (!(x < ed[lo + half])) { lo += half + 1; len -= half + 1; } else len = half;
    }
    g += (lo - 1) * a.stride[s * kMaxDim + d];
  }
  if (ok && a.uniform && !a.uniform[s]) {
    // non-uniform arm of FindGlobalBin (Samples/BinningHandler.cpp:278-290): g is the mega bin; scan its boxes in
    // order, BinInfo::IsEventInside = (lo, hi] in every dimension (Samples/SampleStructs.h:207-219)
    const int g0 = a.grid_off[s] + g;
    int found = -1;
    for (int k = a.grid_start[g0]; k < a.grid_start[g0 + 1] && found < 0; ++k) {
      const int b = a.grid_idx[k];
      const double* ex = a.boxes + (static_cast<int64_t>(a.box_off[s] + b) * nd) * 2;
      bool inside = true;
      for (int d = 0; d < nd; ++d) {
        const double x = a.kin[static_cast<int64_t>(d) * a.n_events + e];
        inside &= (x > ex[2 * d]) & (x <= ex[2 * d + 1]);
      }
      if (inside) found = b;
    }
    a.bin[e] = found >= 0 ? found + a.global_off[s] : -1;
    return;
  }
  a.bin[e] = ok ? g + a.global_off[s] : -1;
}
cudaError_t launch_bins(const BinArgs& a, cudaStream_t s) {
  const int threads = 256;
  const int64_t blocks = (a.e_pad + threads - 1) / threads;
  bin_kernel<<<static_cast<unsigned>(blocks), threads, 0, s>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// SampleHandlerFD::IsEventSelected (Samples/SampleHandlerFD.cpp:281-294): the event fails on the first cut of its
// sample with Val < LowerBound || Val >= UpperBound.  A rejected event never reaches CalcWeightTotal or the
// histogram (:361, :424): the fill kernels see bin -1 for it.
// ------------------------------------------------------------------------------------------------
__global__ void select_kernel(const __grid_constant__ SelectArgs a) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= a.e_pad) return;
  if (e >= a.n_events) { a.bin[e] = -1; return; }
  const int s = a.sample_id[e];
  bool ok = true;
  for (int k = a.cut_start[s]; k < a.cut_start[s + 1] && ok; ++k) {
    const int v = a.cut_var[k];
    const double val = v >= 0 ? a.values[static_cast<int64_t>(v) * a.n_events + e]
                              : a.kin[static_cast<int64_t>(-1 - v) * a.n_events + e];
    if ((val < a.lower[k]) || (val >= a.upper[k])) ok = false;
  }
  a.selected[e] = ok ? 1 : 0;
  a.bin[e] = ok ? a.bin_raw[e] : -1;
}
// A small oscillator table handed over in mapped host memory is fetched by a kernel instead of the copy engine: a DMA
// transfer followed by a dependent kernel costs ~30 us on the stream, a kernel reading 16 KB over PCIe a few.
__global__ void table_copy_kernel(float* __restrict__ dst, const float* __restrict__ src_host, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = src_host[i];
}
cudaError_t launch_table_copy(float* dst, const float* src_host_devptr, int64_t n, cudaStream_t s) {
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 64)));
  table_copy_kernel<<<grid, 256, 0, s>>>(dst, src_host_devptr, n);
  return cudaGetLastError();
}

cudaError_t launch_select(const SelectArgs& a, cudaStream_t s) {
  const int threads = 256;
  select_kernel<<<static_cast<unsigned>((a.e_pad + threads - 1) / threads), threads, 0, s>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// SampleHandlerFD::ApplyShifts (Samples/SampleHandlerFD.cpp:545-564) for linear shifts: ResetShifts (back to the nominal
// values), then x_t += theta * c for the event's entries in funcParsGrid order.  __dmul_rn / __dadd_rn: two roundings,
// no FMA contraction -- the arithmetic of the host code the reference compiles for its functional parameters.
// ------------------------------------------------------------------------------------------------
__global__ void shift_kernel(const __grid_constant__ ShiftArgs a) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= a.n_events) return;
  const int64_t k0 = a.start[e], k1 = a.start[e + 1];
  double x[kMaxDim];
  #pragma unroll
  for (int d = 0; d < kMaxDim; ++d) x[d] = d < a.n_dims ? a.kin_nom[static_cast<int64_t>(d) * a.n_events + e] : 0.;
  for (int v = 0; v < a.n_sel_vars; ++v) a.sel[static_cast<int64_t>(v) * a.n_events + e] = a.sel_nom[static_cast<int64_t>(v) * a.n_events + e];
  for (int64_t k = k0; k < k1; ++k) {
    const int t = a.target[k];
    const double delta = __dmul_rn(a.theta[a.par[k]], a.coef[k]);
    if (t < a.n_dims) {
      #pragma unroll
      for (int d = 0; d < kMaxDim; ++d) if (d == t) x[d] = __dadd_rn(x[d], delta);
    } else {
      double* p = a.sel + static_cast<int64_t>(t - a.n_dims) * a.n_events + e;
      *p = __dadd_rn(*p, delta);
    }
  }
  #pragma unroll
  for (int d = 0; d < kMaxDim; ++d) if (d < a.n_dims) a.kin[static_cast<int64_t>(d) * a.n_events + e] = x[d];
}
cudaError_t launch_shift(const ShiftArgs& a, cudaStream_t s) {
  const int threads = 256;
  shift_kernel<<<static_cast<unsigned>((a.n_events + threads - 1) / threads), threads, 0, s>>>(a);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// setup: AoS monolith -> tiled SoA.  identity rows first ({1,0,0,0} / {0,1}: a response that
// multiplies by exactly 1.0f), then every event scatters its own responses.
// ------------------------------------------------------------------------------------------------
__global__ void identity_kernel(float4* cub, int64_t n_cub, float2* lin, int64_t n_lin) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t k = i; k < n_cub; k += stride) cub[k] = make_float4(1.f, 0.f, 0.f, 0.f);
  for (int64_t k = i; k < n_lin; k += stride) lin[k] = make_float2(0.f, 1.f);
}

__global__ void retile_kernel(const __grid_constant__ RetileArgs a) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const int64_t tile = i / a.T;
  const int lane = static_cast<int>(i - tile * a.T);
  const int sig = a.tile_sig[tile];
  const int16_t* slot_of = a.slot_of_param + static_cast<int64_t>(sig) * a.P;
  const int32_t* segbase_of = a.segbase_of_param + static_cast<int64_t>(sig) * a.P;
  float4* cub = a.cub_pool + a.tile_cub_off[tile] + lane;
  float2* lin = a.lin_pool + a.tile_lin_off[tile] + lane;
  for (uint64_t s = a.start_c[i]; s < a.start_c[i + 1]; ++s) {
    const int p = a.paramNo[s];
    const int nseg = a.nseg[p];
    const float4* src = a.coeff_many + a.knot_off[s];
    float4* dst = cub + static_cast<int64_t>(segbase_of[p]) * a.T;
    for (int k = 0; k < nseg; ++k) dst[static_cast<int64_t>(k) * a.T] = src[k];
  }
  for (uint64_t s = a.start_l[i]; s < a.start_l[i + 1]; ++s) {
    const int p = a.paramNo_l[s];
    lin[static_cast<int64_t>(slot_of[p]) * a.T] = a.coeff_l[s];
  }
}

cudaError_t launch_retile(const RetileArgs& a, int64_t n_identity_cub, int64_t n_identity_lin, cudaStream_t s) {
  if (n_identity_cub > 0 || n_identity_lin > 0) {
    identity_kernel<<<148 * 8, 256, 0, s>>>(a.cub_pool, n_identity_cub, a.lin_pool, n_identity_lin);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  if (a.n > 0) {
    const int threads = 128;
    retile_kernel<<<static_cast<unsigned>((a.n + threads - 1) / threads), threads, 0, s>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace m3b
