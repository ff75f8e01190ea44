// m3b_fill_tma.cu -- the streaming (TMA) form of the fused per-step kernel for sm_100a.
//
// Same arithmetic as fill_kernel (m3b_kernels.cu): SMonolith::CalcSplineWeights + CalcTotalEventWeight
// (Splines/SplineMonolith.cpp:727-830), SampleHandlerFD::CalcWeightTotal + FillArray_MP
// (Samples/SampleHandlerFD.cpp:390-448, 568-594) and the GetLikelihood reduction
// (Samples/SampleHandlerFD.cpp:1284-1300), but the coefficient stream no longer goes through
// registers.  One persistent block per SM:
//
//   producer warp   takes tiles from a global counter (dynamic schedule: no tail imbalance between
//                   SMs) and, for each tile, issues one 1-D bulk copy (cp.async.bulk, SASS UBLKCP)
//                   per active coefficient row -- T*16 B, fully contiguous because the active segment
//                   is uniform per parameter per step -- into a ring of shared-memory stages of G
//                   rows each, completion counted by an mbarrier per stage;
//   T consumer      one event per thread: wait for a stage, read its own float4/float2 from every
//   threads         row (conflict-free LDS.128), Horner fmaf, sequential float product in the
//                   reference's order, release the stage; at the tile's last stage: norms x osc x
//                   spline x static, w<=0 / overflow skip, shared-memory privatised f64 histogram.
//
// Bytes in flight are bounded by the ring (up to ~200 KB per SM), not by registers x occupancy, and
// only T events per SM are in flight, so the end-of-grid tail is ~1/26 of the old kernel's for cfg2.
#include "m3b_device.cuh"

namespace m3b {

constexpr int kMaxStages = 32;
constexpr int kFlagLinear = 1, kFlagFirst = 2, kFlagLast = 4;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync(int n_threads) {     // named barrier 1: consumers only
  asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory");
}

TmaSmem tma_smem_layout(const StepLayout& step, int max_nc, int max_nl, int n_bins, bool hist_in_smem, bool w2_live,
                        int T, int G, int n_stages) {
  TmaSmem L;
  L.off_dx = step.bytes;
  L.off_lv = L.off_dx + 4 * ((max_nc + 3) & ~3);
  L.off_row = (L.off_lv + 4 * max_nl + 15) & ~15;
  L.off_desc = (L.off_row + 4 * max_nc + 15) & ~15;
  L.off_hist = (L.off_desc + 16 * kMaxStages + 15) & ~15;
  const int hist_bytes = hist_in_smem ? 8 * n_bins * (w2_live ? 2 : 1) : 0;
  L.off_ring = (L.off_hist + hist_bytes + 127) & ~127;
  L.stage_bytes = G * T * 16;
  L.total = L.off_ring + n_stages * L.stage_bytes;
  return L;
}

template <int T, int G>
__global__ void __launch_bounds__(T + 32, 1) fill_tma_kernel(const __grid_constant__ FillArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t step_bar;
  __shared__ int s_last;

  constexpr int kConsumerWarps = T / 32;
  constexpr int kStageBytes = G * T * 16;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NS = a.n_stages;
  const bool w2_live = a.w2 != nullptr;
  const bool smem_hist = a.hist_in_smem != 0;

  const TmaSmem L = a.tma;
  unsigned char* st = smem;
  float* s_dx = reinterpret_cast<float*>(smem + L.off_dx);
  float* s_lv = reinterpret_cast<float*>(smem + L.off_lv);
  int32_t* s_row = reinterpret_cast<int32_t*>(smem + L.off_row);     // producer-private
  int4* s_desc = reinterpret_cast<int4*>(smem + L.off_desc);
  double* s_hist = reinterpret_cast<double*>(smem + L.off_hist);
  double* s_w2 = s_hist + a.n_bins;
  unsigned char* ring = smem + L.off_ring;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kConsumerWarps); }
    mbar_init(&step_bar, 1);
  }
  __syncthreads();
  // stage the per-step {segment, dx, value, norm} table with one bulk copy
  if (tid == 0) {
    mbar_expect_tx(&step_bar, static_cast<uint32_t>(a.step.bytes));
    bulk_g2s(st, a.step_table, static_cast<uint32_t>(a.step.bytes), &step_bar);
  }
  if (smem_hist && !a.weights_only) {
    for (int i = tid; i < a.n_bins; i += T + 32) s_hist[i] = 0.;
    if (w2_live) for (int i = tid; i < a.n_bins; i += T + 32) s_w2[i] = 0.;
  }
  mbar_wait(&step_bar, 0);
  __syncthreads();

  const int32_t* seg = reinterpret_cast<const int32_t*>(st + a.step.off_seg);
  const float* dxp = reinterpret_cast<const float*>(st + a.step.off_dx);
  const float* val = reinterpret_cast<const float*>(st + a.step.off_val);
  const float* norm = reinterpret_cast<const float*>(st + a.step.off_norm);

  if (warp == kConsumerWarps) {
    // ------------------------------------------------------------------ producer warp
    int stage = 0, prod_sig = -1;
    uint32_t phase = 1;       // a fresh mbarrier passes a wait on the "previous" phase
    auto advance = [&]() { if (++stage == NS) { stage = 0; phase ^= 1u; } };
    while (true) {
      int t = 0;
      if (lane == 0) t = a.tile_begin + static_cast<int>(atomicAdd(a.tile_counter, 1u));
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= a.n_tiles) break;
      const TileDesc td = a.tiles[t];
      const SigDesc sd = a.sigs[td.sig];
      const int nc = sd.nc, nl = sd.nl;
      if (td.sig != prod_sig) {      // first active row of every cubic slot, this step's segments
        const int32_t* pool = a.sig_pool + sd.off;
        __syncwarp();
        for (int s = lane; s < nc; s += 32) s_row[s] = pool[nc + s] + seg[pool[s]];
        __syncwarp();
        prod_sig = td.sig;
      }
      const int ncs = (nc + G - 1) / G, nls = (nl + 2 * G - 1) / (2 * G);
      const int total = ncs + nls > 0 ? ncs + nls : 1;
      int k = 0;
      for (int s0 = 0; s0 < nc; s0 += G, ++k) {
        const int n = nc - s0 < G ? nc - s0 : G;
        mbar_wait(&empty_bar[stage], phase);
        if (lane == 0) {
          s_desc[stage] = make_int4(t, s0, n, (k == 0 ? kFlagFirst : 0) | (k == total - 1 ? kFlagLast : 0) | (td.sig << 8));
          mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(n) * T * 16u);
        }
        __syncwarp();
        if (lane < n) {
          const int64_t row = static_cast<int64_t>(s_row[s0 + lane]) * T;
          bulk_g2s(ring + static_cast<size_t>(stage) * kStageBytes + static_cast<size_t>(lane) * T * 16, td.cub + row,
                   T * 16u, &full_bar[stage]);
        }
        advance();
      }
      for (int s0 = 0; s0 < nl; s0 += 2 * G, ++k) {
        const int n = nl - s0 < 2 * G ? nl - s0 : 2 * G;
        mbar_wait(&empty_bar[stage], phase);
        if (lane == 0) {
          s_desc[stage] = make_int4(t, s0, n, kFlagLinear | (k == 0 ? kFlagFirst : 0) | (k == total - 1 ? kFlagLast : 0) | (td.sig << 8));
          mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(n) * T * 8u);
          bulk_g2s(ring + static_cast<size_t>(stage) * kStageBytes, td.lin + static_cast<int64_t>(s0) * T,
                   static_cast<uint32_t>(n) * T * 8u, &full_bar[stage]);
        }
        __syncwarp();
        advance();
      }
      if (k == 0) {   // a tile without response functions still has events to weight and fill
        mbar_wait(&empty_bar[stage], phase);
        if (lane == 0) {
          s_desc[stage] = make_int4(t, 0, 0, kFlagFirst | kFlagLast | (td.sig << 8));
          mbar_arrive(&full_bar[stage]);
        }
        __syncwarp();
        advance();
      }
    }
    // terminal stage
    mbar_wait(&empty_bar[stage], phase);
    if (lane == 0) {
      s_desc[stage] = make_int4(-1, 0, 0, 0);
      mbar_arrive(&full_bar[stage]);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ consumers: one event per thread
    int stage = 0;
    uint32_t phase = 0;
    int cur_sig = -1;
    float w_spl = 1.0f, w_osc = 1.f, w_static = 1.f;
    int bin = -1;
    int64_t e = 0;
    while (true) {
      mbar_wait(&full_bar[stage], phase);
      const int4 d = s_desc[stage];
      if (d.x < 0) break;
      const int s0 = d.y, n = d.z, flags = d.w & 0xff;
      if (flags & kFlagFirst) {
        const int sig = d.w >> 8;
        if (sig != cur_sig) {          // uniform over the consumers: they all walk the same stage sequence
          consumer_sync(T);
          const SigDesc sd = a.sigs[sig];
          const int32_t* pool = a.sig_pool + sd.off;
          for (int s = tid; s < sd.nc; s += T) s_dx[s] = dxp[pool[s]];
          for (int s = tid; s < sd.nl; s += T) s_lv[s] = val[pool[2 * sd.nc + s]];
          cur_sig = sig;
          consumer_sync(T);
        }
        e = static_cast<int64_t>(d.x) * T + tid;
        // event-table loads fly while the tile's stages are consumed
        bin = a.bin[e];
        w_osc = 1.f; w_static = 1.f;
        if (a.osc) {
          const int64_t oi = a.osc_idx ? static_cast<int64_t>(a.osc_idx[e]) : (e < a.n_events ? e : 0);
          w_osc = a.osc[oi];
        }
        if (a.static_w) w_static = a.static_w[e];
        w_spl = 1.0f;
      }
      const unsigned char* sb = ring + static_cast<size_t>(stage) * kStageBytes;
      if (!(flags & kFlagLinear)) {
        const float4* rows = reinterpret_cast<const float4*>(sb) + tid;
        if (n == G) {
          float4 c[G];
          #pragma unroll
          for (int j = 0; j < G; ++j) c[j] = rows[j * T];
          #pragma unroll
          for (int j = 0; j < G; ++j) {
            const float dx = s_dx[s0 + j];
            w_spl *= fmaf(dx, fmaf(dx, fmaf(dx, c[j].w, c[j].z), c[j].y), c[j].x);
          }
        } else {
          for (int j = 0; j < n; ++j) {
            const float4 c = rows[j * T];
            const float dx = s_dx[s0 + j];
            w_spl *= fmaf(dx, fmaf(dx, fmaf(dx, c.w, c.z), c.y), c.x);
          }
        }
      } else {
        const float2* rows = reinterpret_cast<const float2*>(sb) + tid;
        #pragma unroll 4
        for (int j = 0; j < n; ++j) {
          const float2 c = rows[j * T];
          w_spl *= fmaf(c.x, s_lv[s0 + j], c.y);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);       // stage may be refilled
      if (++stage == NS) { stage = 0; phase ^= 1u; }

      if (flags & kFlagLast) {
        // CalcWeightTotal: norms (double -> float on the host) in pointer order, then osc, spline, extras
        float w = 1.0f;
        for (int j = 0; j < a.norm_slots; ++j) {
          const int i = a.norm_idx[static_cast<int64_t>(j) * a.e_pad + e];
          w *= (i >= 0 ? norm[i] : 1.0f);
        }
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
    }
  }
  finish_block(a, s_hist, s_w2, reinterpret_cast<double*>(smem), &s_last);
}

#define M3B_TMA_DISPATCH(T_RUNTIME, G_RUNTIME, EXPR)                                          \
  if (T_RUNTIME == 256 && G_RUNTIME == 8) { auto k = fill_tma_kernel<256, 8>; EXPR; }         \
  else if (T_RUNTIME == 256 && G_RUNTIME == 4) { auto k = fill_tma_kernel<256, 4>; EXPR; }    \
  else if (T_RUNTIME == 128 && G_RUNTIME == 8) { auto k = fill_tma_kernel<128, 8>; EXPR; }    \
  else if (T_RUNTIME == 128 && G_RUNTIME == 16) { auto k = fill_tma_kernel<128, 16>; EXPR; }  \
  else if (T_RUNTIME == 512 && G_RUNTIME == 4) { auto k = fill_tma_kernel<512, 4>; EXPR; }    \
  else if (T_RUNTIME == 512 && G_RUNTIME == 8) { auto k = fill_tma_kernel<512, 8>; EXPR; }    \
  else return cudaErrorInvalidValue;

cudaError_t launch_fill_tma(const FillArgs& a, int G, int grid, int smem, cudaStream_t s) {
  M3B_TMA_DISPATCH(a.T, G, (k<<<grid, a.T + 32, smem, s>>>(a)))
  return cudaGetLastError();
}
cudaError_t fill_tma_set_smem(int T, int G, int smem) {
  M3B_TMA_DISPATCH(T, G, return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem))
  return cudaSuccess;
}
cudaError_t fill_tma_occupancy(int T, int G, int smem, int* bps) {
  M3B_TMA_DISPATCH(T, G, return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, k, T + 32, smem))
  return cudaSuccess;
}

}  // namespace m3b
