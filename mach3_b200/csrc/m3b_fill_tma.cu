// m3b_fill_tma.cu -- the streaming (TMA) form of the fused per-step kernel for sm_100a.
//
// Same arithmetic as fill_kernel (m3b_kernels.cu): SMonolith::CalcSplineWeights + CalcTotalEventWeight
// (Splines/SplineMonolith.cpp:727-830), SampleHandlerFD::CalcWeightTotal + FillArray_MP
// (Samples/SampleHandlerFD.cpp:390-448, 568-594) and the GetLikelihood reduction
// (Samples/SampleHandlerFD.cpp:1284-1300), but the coefficient stream no longer goes through
// registers.  One persistent block per SM:
//
//   producer warp   takes work from a global counter (guided self-scheduling: whole tile rows while
//                   plenty is left, single 256-event units at the end, so the SMs finish together) and
//                   issues one 1-D bulk copy (cp.async.bulk, SASS UBLKCP) per active coefficient row
//                   -- up to T*16 B, fully contiguous because the active segment is uniform per
//                   parameter per step -- into a ring of 32 KB shared-memory stages, completion
//                   counted by an mbarrier per stage;
//   256 consumer    g (1,2,4) events per thread: wait for a stage, read their own float4/float2 from
//   threads         every row (conflict-free LDS.128), Horner fmaf, sequential float product in the
//                   reference's order, release the stage; at the grab's last stage: norms x osc x
//                   spline x static, w<=0 / overflow skip, shared-memory privatised f64 histogram.
//
// Bytes in flight are bounded by the ring (up to ~200 KB per SM), not by registers x occupancy.
#include "m3b_device.cuh"

namespace m3b {

constexpr int kMaxStages = 32;
constexpr int kFlagLinear = 1, kFlagFirst = 2, kFlagLast = 4, kFlagOscDirect = 8;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync(int n_threads) {     // named barrier 1: consumers only
  asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory");
}

TmaSmem tma_smem_layout(const StepLayout& step, int max_nc, int max_nl, int n_bins, bool hist_in_smem, bool w2_live,
                        bool osc_slots, int n_stages) {
  TmaSmem L;
  L.off_dx = step.bytes;
  L.off_lv = L.off_dx + 4 * ((max_nc + 3) & ~3);
  L.off_row = (L.off_lv + 4 * max_nl + 15) & ~15;
  L.off_desc = (L.off_row + 4 * max_nc + 15) & ~15;
  L.off_hist = (L.off_desc + 16 * kMaxStages + 15) & ~15;
  const int hist_bytes = hist_in_smem ? 8 * n_bins * (w2_live ? 2 : 1) : 0;
  L.off_osc = (L.off_hist + hist_bytes + 127) & ~127;              // per stage: kMaxG units x 256 floats
  L.off_ring = L.off_osc + (osc_slots ? n_stages * 4 * 256 * 4 : 0);
  L.stage_bytes = 8 * 256 * 16;
  L.total = L.off_ring + n_stages * L.stage_bytes;
  return L;
}

// One work unit = kUnit consecutive lanes of a tile row (kUnit consumer threads, one event each).
// A tile row holds T = q_max*kUnit events; the producer grabs g in {1,2,4,...,q_max} aligned units
// at a time -- whole rows while plenty of work is left (g*kUnit*16 B contiguous per copy: DRAM likes
// long bursts, scripts/hbm_probe.cu), single units near the end of the grid so the SMs finish
// together (guided self-scheduling).  A consumer thread then carries g events.
constexpr int kUnit = 256;
constexpr int kStageBytes = 8 * kUnit * 16;      // 32 KB: 8/g cubic rows or 16/g linear rows of g units
constexpr int kMaxG = 4;

template <int GQ>
__device__ __forceinline__ void consume_cubic(const float4* rows, int n, const float* dx, float (&w)[kMaxG]) {
  constexpr int kRows = 8 / GQ;
  if (n == kRows) {
    float4 c[kRows][GQ];
    #pragma unroll
    for (int j = 0; j < kRows; ++j)
      #pragma unroll
      for (int q = 0; q < GQ; ++q) c[j][q] = rows[(j * GQ + q) * kUnit];
    #pragma unroll
    for (int j = 0; j < kRows; ++j) {
      const float d = dx[j];
      #pragma unroll
      for (int q = 0; q < GQ; ++q) w[q] *= fmaf(d, fmaf(d, fmaf(d, c[j][q].w, c[j][q].z), c[j][q].y), c[j][q].x);
    }
  } else {
    for (int j = 0; j < n; ++j) {
      const float d = dx[j];
      #pragma unroll
      for (int q = 0; q < GQ; ++q) {
        const float4 c = rows[(j * GQ + q) * kUnit];
        w[q] *= fmaf(d, fmaf(d, fmaf(d, c.w, c.z), c.y), c.x);
      }
    }
  }
}
template <int GQ>
__device__ __forceinline__ void consume_linear(const float2* rows, int n, const float* lv, float (&w)[kMaxG]) {
  #pragma unroll 2
  for (int j = 0; j < n; ++j) {
    const float v = lv[j];
    #pragma unroll
    for (int q = 0; q < GQ; ++q) {
      const float2 c = rows[(j * GQ + q) * kUnit];
      w[q] *= fmaf(c.x, v, c.y);
    }
  }
}

__global__ void __launch_bounds__(kUnit + 32, 1) fill_tma_kernel(const __grid_constant__ FillArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t step_bar;
  __shared__ int s_last;

  constexpr int kConsumerWarps = kUnit / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NS = a.n_stages;
  const int T = a.T;                       // events per tile row
  const int qmax = T / kUnit;              // units per tile
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
  float* s_osc = reinterpret_cast<float*>(smem + L.off_osc);

  if (tid == 0) trace_mark(a, 0);
  // let the next queued step's blocks take over SMs as ours retire (it waits before touching anything we write)
  if (a.pdl) grid_launch_dependents();
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kConsumerWarps); }
    mbar_init(&step_bar, 1);
  }
  __syncthreads();
  if (smem_hist && !a.weights_only) {
    for (int i = tid; i < a.n_bins; i += kUnit + 32) s_hist[i] = 0.;
    if (w2_live) for (int i = tid; i < a.n_bins; i += kUnit + 32) s_w2[i] = 0.;
  }
  // Work scheduling.  First round: static (block b owns units [b*w0, (b+1)*w0)), so the producer can fetch its
  // tile descriptor while the tables are staged; afterwards guided grabs from the global counter.
  const unsigned int unit_begin = static_cast<unsigned int>(a.tile_begin) * qmax;
  const unsigned int unit_end = static_cast<unsigned int>(a.n_tiles) * qmax;
  const unsigned int guard = (static_cast<unsigned int>(a.guard_x2) * gridDim.x) >> 1;
  int w0 = qmax;
  while (w0 > 1 && unit_end - unit_begin < guard * w0) w0 >>= 1;
  const unsigned int static_units = gridDim.x * static_cast<unsigned int>(w0);
  const unsigned int u_first = unit_begin + blockIdx.x * static_cast<unsigned int>(w0);
  TileDesc td_pre{};
  if (warp == kConsumerWarps && u_first < unit_end) td_pre = a.tiles[u_first / qmax];     // in flight during the staging
  // stage the per-step {segment, dx, value, norm, per-signature slot} tables
  stage_step_table(a, st, &step_bar);
  if (tid == 0) trace_mark(a, 1);

  const int32_t* seg = reinterpret_cast<const int32_t*>(st + a.step.off_seg);
  const float* dxp = reinterpret_cast<const float*>(st + a.step.off_dx);
  const float* val = reinterpret_cast<const float*>(st + a.step.off_val);
  const float* norm = reinterpret_cast<const float*>(st + a.step.off_norm);

  if (warp == kConsumerWarps) {
    // ------------------------------------------------------------------ producer warp
    int stage = 0, prod_sig = -1;
    uint32_t phase = 1;       // a fresh mbarrier passes a wait on the "previous" phase
    auto advance = [&]() { if (++stage == NS) { stage = 0; phase ^= 1u; } };
    unsigned int seen = 0;          // last counter value this producer saw: estimates the work left
    unsigned int u = u_first;       // units [u, u_end) grabbed and not yet issued
    unsigned int u_end = u + w0 < unit_end ? u + w0 : unit_end;
    unsigned long long n_units = 0;
    const bool expanded = a.step.n_sigs_x > 0;
    const int32_t* rowx = reinterpret_cast<const int32_t*>(st + a.step.off_rowx);
    bool fresh = true, pdl_waited = false;
    while (true) {
      if (fresh) { fresh = false; if (u >= unit_end) break; }
      else if (u >= u_end) {
        // guided grab: one atomicAdd (never retries, no CAS storms); the size comes from the
        // work left as of this producer's previous grab
        unsigned int c = 0; int want = 0;
        if (a.pdl && !pdl_waited) { grid_dependency_wait(); pdl_waited = true; }    // the counter was re-armed by the previous step's last block
        if (lane == 0) {
          const unsigned int done = static_units + seen;
          const unsigned int left = unit_begin + done < unit_end ? unit_end - unit_begin - done : 0u;
          want = qmax;
          while (want > 1 && left < guard * want) want >>= 1;
          c = atomicAdd(a.tile_counter, static_cast<unsigned int>(want));
        }
        want = __shfl_sync(0xffffffffu, want, 0);
        c = __shfl_sync(0xffffffffu, c, 0);
        seen = c + want;
        u = unit_begin + static_units + c;
        if (u >= unit_end) break;
        u_end = u + want < unit_end ? u + want : unit_end;
      }
      // largest aligned power-of-two piece of [u, u_end): stays inside one tile row
      int g = qmax;
      while (g > 1 && ((u & (g - 1)) != 0 || u + g > u_end)) g >>= 1;
      const int t = static_cast<int>(u / qmax);
      const int lane0 = static_cast<int>(u % qmax) * kUnit;       // first lane of the grab within the tile row
      const TileDesc td = (u == u_first) ? td_pre : a.tiles[t];
      const int nc = td.ncnl & 0xffff, nl = td.ncnl >> 16;
      // zero-copy oscillation weights: g*1 KB straight from pinned host memory, riding on the grab's
      // last stage (they are needed only when the event's total weight is formed)
      const int64_t e_first = static_cast<int64_t>(t) * T + lane0;
      // (the grab holding the ragged end of the event list reads its few weights with plain loads)
      const bool osc_tail = a.osc_host && e_first + static_cast<int64_t>(g) * kUnit > a.n_events;
      const uint32_t osc_bytes = (a.osc_host && !osc_tail) ? static_cast<uint32_t>(g) * kUnit * 4u : 0u;
      const int xflags = (td.sig << 8) | (osc_tail ? kFlagOscDirect : 0);
      const int32_t* rowp = rowx + td.sig * a.step.max_nc;     // first active row of every cubic slot
      if (!expanded) {
        if (td.sig != prod_sig) {    // too many signatures for the host-expanded tables: build it here
          const int32_t* pool = a.sig_pool + a.sigs[td.sig].off;
          __syncwarp();
          for (int s = lane; s < nc; s += 32) s_row[s] = pool[nc + s] + seg[pool[s]];
          __syncwarp();
          prod_sig = td.sig;
        }
        rowp = s_row;
      }
      const int rc = 8 / g, rl = 16 / g;                           // rows per stage
      const int ncs = (nc + rc - 1) / rc, nls = (nl + rl - 1) / rl;
      const int total = ncs + nls > 0 ? ncs + nls : 1;
      const int where = t | (lane0 / kUnit) << 24;                 // tile (24 bits) | first unit
      int k = 0;
      for (int s0 = 0; s0 < nc; s0 += rc, ++k) {
        const int n = nc - s0 < rc ? nc - s0 : rc;
        const uint32_t row_bytes = static_cast<uint32_t>(g) * kUnit * 16u;
        mbar_wait(&empty_bar[stage], phase);
        const bool last = k == total - 1;
        if (lane == 0) {
          s_desc[stage] = make_int4(where, s0 | n << 16, g, (k == 0 ? kFlagFirst : 0) | (last ? kFlagLast : 0) | xflags);
          mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(n) * row_bytes + (last ? osc_bytes : 0u));
        }
        __syncwarp();
        if (lane < n)
          bulk_g2s(ring + static_cast<size_t>(stage) * kStageBytes + static_cast<size_t>(lane) * row_bytes,
                   td.cub + static_cast<int64_t>(rowp[s0 + lane]) * T + lane0, row_bytes, &full_bar[stage]);
        if (last && osc_bytes && lane == 31)
          bulk_g2s(s_osc + static_cast<size_t>(stage) * (kMaxG * kUnit), a.osc_host + e_first, osc_bytes, &full_bar[stage]);
        advance();
      }
      for (int s0 = 0; s0 < nl; s0 += rl, ++k) {
        const int n = nl - s0 < rl ? nl - s0 : rl;
        const uint32_t row_bytes = static_cast<uint32_t>(g) * kUnit * 8u;
        mbar_wait(&empty_bar[stage], phase);
        const bool last = k == total - 1;
        if (lane == 0) {
          s_desc[stage] = make_int4(where, s0 | n << 16, g, kFlagLinear | (k == 0 ? kFlagFirst : 0) | (last ? kFlagLast : 0) | xflags);
          mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(n) * row_bytes + (last ? osc_bytes : 0u));
        }
        __syncwarp();
        if (last && osc_bytes && lane == 31)
          bulk_g2s(s_osc + static_cast<size_t>(stage) * (kMaxG * kUnit), a.osc_host + e_first, osc_bytes, &full_bar[stage]);
        if (g == qmax) {       // whole rows: the slots are contiguous, one copy
          if (lane == 0)
            bulk_g2s(ring + static_cast<size_t>(stage) * kStageBytes, td.lin + static_cast<int64_t>(s0) * T,
                     static_cast<uint32_t>(n) * row_bytes, &full_bar[stage]);
        } else if (lane < n) {
          bulk_g2s(ring + static_cast<size_t>(stage) * kStageBytes + static_cast<size_t>(lane) * row_bytes,
                   td.lin + static_cast<int64_t>(s0 + lane) * T + lane0, row_bytes, &full_bar[stage]);
        }
        advance();
      }
      if (k == 0) {   // a tile without response functions still has events to weight and fill
        mbar_wait(&empty_bar[stage], phase);
        if (lane == 0) {
          s_desc[stage] = make_int4(where, 0, g, kFlagFirst | kFlagLast | xflags);
          if (osc_bytes) {
            mbar_expect_tx(&full_bar[stage], osc_bytes);
            bulk_g2s(s_osc + static_cast<size_t>(stage) * (kMaxG * kUnit), a.osc_host + e_first, osc_bytes, &full_bar[stage]);
          } else {
            mbar_arrive(&full_bar[stage]);
          }
        }
        __syncwarp();
        advance();
      }
      u += g;
      n_units += g;
    }
    if (lane == 0) { trace_mark(a, 3); if (a.trace) a.trace[static_cast<size_t>(blockIdx.x) * 8 + 7] = n_units; }
    // terminal stage
    mbar_wait(&empty_bar[stage], phase);
    if (lane == 0) {
      s_desc[stage] = make_int4(-1, 0, 0, 0);
      mbar_arrive(&full_bar[stage]);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ consumers: g events per thread
    int stage = 0;
    uint32_t phase = 0;
    int cur_sig = -1;
    float w_spl[kMaxG], w_osc[kMaxG], w_static[kMaxG];
    int bin[kMaxG];
    #pragma unroll
    for (int q = 0; q < kMaxG; ++q) { w_spl[q] = 1.f; w_osc[q] = 1.f; w_static[q] = 1.f; bin[q] = -1; }
    int64_t e0 = 0;
    const float* c_dx = s_dx;
    const float* c_lv = s_lv;
    bool first = true;
    while (true) {
      mbar_wait(&full_bar[stage], phase);
      const int4 d = s_desc[stage];
      if (d.x < 0) break;
      if (first) { first = false; if (tid == 0) trace_mark(a, 2); }
      const int s0 = d.y & 0xffff, n = d.y >> 16, g = d.z, flags = d.w & 0xff;
      if (flags & kFlagFirst) {
        const int sig = d.w >> 8;
        if (sig != cur_sig) {          // uniform over the consumers: they all walk the same stage sequence
          if (a.step.n_sigs_x > 0) {
            c_dx = reinterpret_cast<const float*>(st + a.step.off_dxx) + sig * a.step.max_nc;
            c_lv = reinterpret_cast<const float*>(st + a.step.off_lvx) + sig * a.step.max_nl;
          } else {
            consumer_sync(kUnit);
            const SigDesc sd = a.sigs[sig];
            const int32_t* pool = a.sig_pool + sd.off;
            for (int s = tid; s < sd.nc; s += kUnit) s_dx[s] = dxp[pool[s]];
            for (int s = tid; s < sd.nl; s += kUnit) s_lv[s] = val[pool[2 * sd.nc + s]];
            consumer_sync(kUnit);
          }
          cur_sig = sig;
        }
        e0 = static_cast<int64_t>(d.x & 0xffffff) * T + (d.x >> 24) * kUnit + tid;
        // event-table loads fly while the tile's stages are consumed
        #pragma unroll
        for (int q = 0; q < kMaxG; ++q) {
          if (q < g) {
            const int64_t e = e0 + q * kUnit;
            bin[q] = a.bin[e];
            w_osc[q] = 1.f; w_static[q] = 1.f;
            if (a.osc && !a.osc_host) {
              const int64_t oi = a.osc_idx ? static_cast<int64_t>(a.osc_idx[e]) : (e < a.n_events ? e : 0);
              w_osc[q] = oi >= 0 ? a.osc[oi] : 1.f;          // osc_idx -1: the event carries no oscillation weight
            }
            if (a.static_w) w_static[q] = a.static_w[e];
            w_spl[q] = 1.0f;
          }
        }
      }
      const unsigned char* sb = ring + static_cast<size_t>(stage) * kStageBytes;
      if (!(flags & kFlagLinear)) {
        const float4* rows = reinterpret_cast<const float4*>(sb) + tid;
        if (g == 4) consume_cubic<4>(rows, n, c_dx + s0, w_spl);
        else if (g == 2) consume_cubic<2>(rows, n, c_dx + s0, w_spl);
        else consume_cubic<1>(rows, n, c_dx + s0, w_spl);
      } else {
        const float2* rows = reinterpret_cast<const float2*>(sb) + tid;
        if (g == 4) consume_linear<4>(rows, n, c_lv + s0, w_spl);
        else if (g == 2) consume_linear<2>(rows, n, c_lv + s0, w_spl);
        else consume_linear<1>(rows, n, c_lv + s0, w_spl);
      }
      if ((flags & kFlagLast) && a.osc_host) {
        #pragma unroll
        for (int q = 0; q < kMaxG; ++q)
          if (q < g) {
            const int64_t e = e0 + q * kUnit;
            if (flags & kFlagOscDirect) w_osc[q] = e < a.n_events ? a.osc_host[e] : 1.f;
            else w_osc[q] = s_osc[stage * (kMaxG * kUnit) + q * kUnit + tid];
            if (e < a.n_events) a.osc_store[e] = w_osc[q];
          }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);       // stage may be refilled
      if (++stage == NS) { stage = 0; phase ^= 1u; }

      if (flags & kFlagLast) {
        #pragma unroll
        for (int q = 0; q < kMaxG; ++q) {
          if (q < g) {
            const int64_t e = e0 + q * kUnit;
            // CalcWeightTotal: norms (double -> float on the host) in pointer order, then osc, spline, extras
            float w = 1.0f;
            for (int j = 0; j < a.norm_slots; ++j) {
              const int i = a.norm_idx[static_cast<int64_t>(j) * a.e_pad + e];
              w *= (i >= 0 ? norm[i] : 1.0f);
            }
            w *= w_osc[q];
            w *= w_spl[q];
            w *= w_static[q];
            if (a.evt_spline_w && e < a.n_events) { a.evt_spline_w[e] = w_spl[q]; a.evt_total_w[e] = w; }
            // FillArray_MP: skip w<=0 and under/overflow; mc += w; w2 += w*w (float product)
            if (w > 0.f && bin[q] >= 0 && !a.weights_only) {
              if (smem_hist) {
                atomicAdd(s_hist + bin[q], static_cast<double>(w));
                if (w2_live) atomicAdd(s_w2 + bin[q], static_cast<double>(w * w));
              } else {
                atomicAdd(a.hist + bin[q], static_cast<double>(w));
                if (w2_live) atomicAdd(a.w2 + bin[q], static_cast<double>(w * w));
              }
            }
          }
        }
      }
    }
  }
  if (tid == 0) trace_mark(a, 4);
  finish_block(a, s_hist, s_w2, reinterpret_cast<double*>(smem), &s_last);
  if (tid == 0) trace_mark(a, 6);
}

cudaError_t launch_fill_tma(const FillArgs& a, int grid, int smem, cudaStream_t s) {
  if (a.T % kUnit != 0 || a.T / kUnit > kMaxG || (a.T / kUnit & (a.T / kUnit - 1)) != 0) return cudaErrorInvalidValue;
  if (a.pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kUnit + 32); cfg.dynamicSmemBytes = static_cast<size_t>(smem); cfg.stream = s;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, fill_tma_kernel, a);
  }
  fill_tma_kernel<<<grid, kUnit + 32, smem, s>>>(a);
  return cudaGetLastError();
}
cudaError_t fill_tma_set_smem(int smem) {
  (void)smem;
  return allow_max_dynamic_smem(fill_tma_kernel);
}
cudaError_t fill_tma_occupancy(int smem, int* bps) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, fill_tma_kernel, kUnit + 32, smem);
}

}  // namespace m3b
