// m3b_binned.cu -- the BinnedSplineHandler form of the hot path (BASELINE config 4).
//
//   binned_eval_kernel   BinnedSplineHandler::CalcSplineWeights (Splines/BinnedSplineHandler.cpp:306-341):
//                        one weight per non-flat binned spline, fmaf Horner on the active segment,
//                        negative weights clamped to 0 (:337); flat splines are never evaluated and
//                        stay at 1.0 (:236,283) -- here they are simply not stored.
//   binned_fill_kernel   SampleHandlerFD::CalcWeightTotal over the event's N weight pointers
//                        (Samples/SampleHandlerFD.cpp:568-594, pointers wired at :1196-1242) + FillArray_MP
//                        (:390-448) + the likelihood (the common epilogue, m3b_device.cuh).
//   host                 m3b_upload_binned_splines / m3b_upload_event_binned_splines / m3b_read_binned_weights
//
// _LOW_MEMORY_STRUCTS_ build of the reference (M3::float_t = float), like the SMonolith path.
#include "m3b_device.cuh"
#include "m3b_handle.h"
#include <type_traits>

namespace m3b {

constexpr int kBTileSplines = 1024;     // splines per BTile: 4 per thread, so one descriptor load feeds 4 x 20 B of stream

__global__ void __launch_bounds__(256) binned_eval_kernel(const __grid_constant__ FillArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  if (a.binned_pdl) grid_launch_dependents();
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  stage_step_table(a, smem, &bar);
  const int32_t* seg = reinterpret_cast<const int32_t*>(smem + a.step.off_seg);
  const float* val = reinterpret_cast<const float*>(smem + a.step.off_val);
  constexpr int U = 2, V = kBTileSplines / 256;      // tiles in flight per block x splines per thread per tile
  for (int t0 = blockIdx.x * U; t0 < a.n_btiles; t0 += gridDim.x * U) {
    float4 c[U][V]; float x[U][V]; float xv[U]; int out[U];
    #pragma unroll
    for (int u = 0; u < U; ++u) {
      out[u] = -1;
      if (t0 + u < a.n_btiles) {
        const BTile bt = a.btiles[t0 + u];
        const int64_t i = bt.coef_off + static_cast<int64_t>(seg[bt.param]) * bt.n_pad + bt.k0 + threadIdx.x;
        #pragma unroll
        for (int v = 0; v < V; ++v) { c[u][v] = ldg_stream(a.bcoef + i + v * 256); x[u][v] = __ldcs(a.bx + i + v * 256); }
        xv[u] = val[bt.param];                 // M3::float_t(*splineParsPointer), :327
        out[u] = bt.out0 + threadIdx.x;
      }
    }
    #pragma unroll
    for (int u = 0; u < U; ++u) {
      if (out[u] >= 0) {
        #pragma unroll
        for (int v = 0; v < V; ++v) {
          const float dx = xv[u] - x[u][v];                                                          // :329
          float w = fmaf(dx, fmaf(dx, fmaf(dx, c[u][v].w, c[u][v].z), c[u][v].y), c[u][v].x);         // :332
          if (w < 0) w = 0.f;                                                                         // :337
          a.bw[out[u] + v * 256] = w;
        }
      }
    }
  }
}

// the same in the reference's default build: M3::float_t = double, fma instead of fmaf (Manager/Core.h:44-51), the
// parameter value read un-narrowed (:327)
__global__ void __launch_bounds__(256) binned_eval_kernel_f64(const __grid_constant__ FillArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  if (a.binned_pdl) grid_launch_dependents();
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  stage_step_table(a, smem, &bar);
  const int32_t* seg = reinterpret_cast<const int32_t*>(smem + a.step.off_seg);
  const double* val = reinterpret_cast<const double*>(smem + a.step.off_val_d);
  constexpr int V = kBTileSplines / 256;
  for (int t = blockIdx.x; t < a.n_btiles; t += gridDim.x) {
    const BTile bt = a.btiles[t];
    const int64_t i = bt.coef_off + static_cast<int64_t>(seg[bt.param]) * bt.n_pad + bt.k0 + threadIdx.x;
    const double xv = val[bt.param];
    double2 lo[V], hi[V]; double x[V];
    #pragma unroll
    for (int v = 0; v < V; ++v) {
      const double2* cp = reinterpret_cast<const double2*>(a.bcoef_d + 4 * (i + v * 256));
      lo[v] = __ldcs(cp); hi[v] = __ldcs(cp + 1); x[v] = __ldcs(a.bx_d + i + v * 256);
    }
    #pragma unroll
    for (int v = 0; v < V; ++v) {
      const double dx = xv - x[v];
      double w = fma(dx, fma(dx, fma(dx, hi[v].y, hi[v].x), lo[v].y), lo[v].x);
      if (w < 0) w = 0.;
      a.bw_d[bt.out0 + threadIdx.x + v * 256] = w;
    }
  }
}

int binned_fill_smem_bytes(const FillArgs& a, bool hist_in_smem, bool w2_live) {
  int b = (a.step.bytes + 15) & ~15;
  if (hist_in_smem) b += 8 * a.n_bins * (w2_live ? 2 : 1);
  const int llh_scratch = a.n_samples * 32 * 8;
  return b > llh_scratch ? b : llh_scratch;
}

// per-event part: CalcWeightTotal's factors other than the binned splines.  The event table is read from the copies laid
// out in walking order (sort_event_table; streaming loads: L1 is for the weight gathers) in two phases, so that the
// prefetch of the next tile only ISSUES loads: binned_event_load = the raw columns, binned_event_weight = what depends on
// them (norm look-ups in the step table, the oscillation weight gather).
constexpr int kBNormFront = 4;
template <class R>
struct BEvent { int32_t e, bin, oi; int16_t ni[kBNormFront]; R w_static; };
template <class R>
__device__ __forceinline__ void binned_event_load(const FillArgs& a, int64_t v, BEvent<R>& o) {
  o.e = __ldcs(a.perm + v);
  o.bin = __ldcs(a.bin_sorted + v);
  o.oi = a.osc_idx_sorted ? __ldcs(a.osc_idx_sorted + v) : -1;
  o.w_static = a.static_sorted ? __ldcs(reinterpret_cast<const R*>(a.static_sorted) + v) : R(1);
  #pragma unroll
  for (int j = 0; j < kBNormFront; ++j) o.ni[j] = j < a.norm_slots ? __ldcs(a.norm_idx_sorted + static_cast<int64_t>(j) * a.e_pad + v) : int16_t(-1);
}
template <class R>
__device__ __forceinline__ R binned_event_weight(const FillArgs& a, const R* norm, const R* oscp, int64_t v, const BEvent<R>& o) {
  // CalcWeightTotal: norms first, then the weight pointers in push order: osc, binned splines, extras
  const R w_osc = (oscp && o.oi >= 0) ? __ldg(oscp + o.oi) : R(1);
  R w = 1;
  #pragma unroll
  for (int j = 0; j < kBNormFront; ++j) w *= (o.ni[j] >= 0 ? norm[o.ni[j]] : R(1));
  for (int j = kBNormFront; j < a.norm_slots; ++j) {
    const int i = __ldcs(a.norm_idx_sorted + static_cast<int64_t>(j) * a.e_pad + v);
    w *= (i >= 0 ? norm[i] : R(1));
  }
  return w * w_osc;
}

// The fill kernel.  One warp = one warp tile of 32 events per iteration, tiles handed out grid-strided, one per warp of the
// block.  The kernel is bound by memory latency, not by bandwidth (three dependent loads per tile: tile descriptor -> ELL
// columns -> weights), so (1) it runs as ONE 1024-thread block per SM at 64 registers: 32 warps share one privatised
// histogram, and whatever shared memory that leaves stays L1 for the gathers (512 x 2 and 256 x 2 were measured slower:
// profiles/r02_binned_fill_ab.txt); (2) it is software-pipelined: while the weights of tile i are gathered, the event table
// and the first kBFront ELL columns of tile i+1 and the descriptor of tile i+2 are already in flight.
template <bool F64, int NT, int kBFront>
__global__ void __launch_bounds__(NT, 1024 / NT >= 4 ? 2 : 1024 / NT) binned_fill_kernel(const __grid_constant__ FillArgs a) {
  using R = typename std::conditional<F64, double, float>::type;      // M3::float_t of the build
  constexpr int kTiles = NT / 32;                    // warp tiles per block iteration
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ int s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double* s_hist = reinterpret_cast<double*>(smem + ((a.step.bytes + 15) & ~15));
  double* s_w2 = s_hist + a.n_bins;
  const bool w2_live = a.w2 != nullptr;
  const bool smem_hist = a.hist_in_smem != 0;
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (smem_hist && !a.weights_only) {
    for (int i = tid; i < a.n_bins; i += NT) s_hist[i] = 0.;
    if (w2_live) for (int i = tid; i < a.n_bins; i += NT) s_w2[i] = 0.;
  }
  stage_step_table(a, smem, &bar);
  const R* norm = reinterpret_cast<const R*>(smem + (F64 ? a.step.off_norm_d : a.step.off_norm));
  const R* bw = F64 ? reinterpret_cast<const R*>(a.bw_d) : reinterpret_cast<const R*>(a.bw);
  const R* oscp = F64 ? reinterpret_cast<const R*>(a.osc_d) : reinterpret_cast<const R*>(a.osc);

  auto load_desc = [&](int64_t t) {                 // max_n = -1: past the end
    WTile d; d.off = 0; d.max_n = -1; d.pad = 0;
    if (t < a.n_wtiles) d = a.wtiles[t];
    return d;
  };
  auto load_front = [&](int64_t t, const WTile& d, BEvent<R>& ev, int (&idx)[kBFront]) {
    if (d.max_n < 0) return;
    // events are processed in the order of their spline-grid cell (a.perm, built at upload): the 32 lanes of a warp --
    // and the warps of the block -- then gather from the same few sectors of every parameter's weight row
    binned_event_load<R>(a, t * 32 + lane, ev);
    const int32_t* col = a.ell + d.off + lane;
    #pragma unroll
    for (int j = 0; j < kBFront; ++j) idx[j] = j < d.max_n ? __ldcs(col + j * 32) : -1;
  };

  // tile of this warp in block iteration `it`: the block takes kTiles consecutive tiles, and the warp-to-tile assignment
  // rotates with the iteration -- the upload orders the tiles of a span from long to short, and a fixed assignment would
  // give the same warps the long tiles every time
  const int64_t n_chunks = (a.n_wtiles + kTiles - 1) / kTiles;
  auto tile_of = [&](int64_t it) -> int64_t {
    const int64_t c = blockIdx.x + it * gridDim.x;
    return c < n_chunks ? c * kTiles + ((warp + static_cast<int>(it % kTiles)) % kTiles) : a.n_wtiles;
  };
  int64_t it = 0;
  int64_t wt = tile_of(0);
  WTile d_nxt = load_desc(wt);
  BEvent<R> ev_nxt; ev_nxt.e = 0; ev_nxt.bin = -1; ev_nxt.oi = -1; ev_nxt.w_static = 0;
  #pragma unroll
  for (int j = 0; j < kBNormFront; ++j) ev_nxt.ni[j] = -1;
  int idx_nxt[kBFront];
  #pragma unroll
  for (int j = 0; j < kBFront; ++j) idx_nxt[j] = -1;
  load_front(wt, d_nxt, ev_nxt, idx_nxt);
  WTile d_nxt2 = load_desc(tile_of(1));
  // everything above is independent of the eval kernel of this step; the weights and the histograms are not
  if (a.binned_pdl) grid_dependency_wait();
  for (; blockIdx.x + it * gridDim.x < n_chunks; ++it, wt = tile_of(it)) {
    const WTile d = d_nxt;
    const BEvent<R> ev = ev_nxt;
    int idx[kBFront]; R gw[kBFront];
    #pragma unroll
    for (int j = 0; j < kBFront; ++j) idx[j] = idx_nxt[j];
    #pragma unroll
    for (int j = 0; j < kBFront; ++j) gw[j] = idx[j] >= 0 ? __ldg(bw + idx[j]) : R(1);
    const R w_pre = binned_event_weight<R>(a, norm, oscp, wt * 32 + lane, ev);
    // next tile's front and the descriptor after it
    d_nxt = d_nxt2;
    load_front(tile_of(it + 1), d_nxt, ev_nxt, idx_nxt);
    d_nxt2 = load_desc(tile_of(it + 2));
    if (d.max_n < 0) continue;                         // the last chunk may be short
    // CalcWeightTotal: norms first, then the weight pointers in push order: osc, binned splines, extras
    R w = w_pre;
    R w_spl = 1;               // product of the binned weights alone (m3b_read_event_weights)
    #pragma unroll
    for (int j = 0; j < kBFront; ++j) if (idx[j] >= 0) { w *= gw[j]; w_spl *= gw[j]; }
    const int32_t* col = a.ell + d.off + lane;
    for (int j0 = kBFront; j0 < d.max_n; j0 += 8) {
      int ix[8]; R gx[8];
      #pragma unroll
      for (int j = 0; j < 8; ++j) ix[j] = (j0 + j < d.max_n) ? __ldcs(col + (j0 + j) * 32) : -1;
      #pragma unroll
      for (int j = 0; j < 8; ++j) gx[j] = ix[j] >= 0 ? __ldg(bw + ix[j]) : R(1);
      #pragma unroll
      for (int j = 0; j < 8; ++j) if (ix[j] >= 0) { w *= gx[j]; w_spl *= gx[j]; }
    }
    w *= ev.w_static;
    const int64_t e = ev.e;
    if (e < a.n_events) {
      if (F64) { if (a.evt_spline_d) { a.evt_spline_d[e] = w_spl; a.evt_total_d[e] = w; } }
      else if (a.evt_spline_w) { a.evt_spline_w[e] = static_cast<float>(w_spl); a.evt_total_w[e] = static_cast<float>(w); }
    }
    if (w > R(0) && ev.bin >= 0 && !a.weights_only) {
      if (smem_hist) {
        atomicAdd(s_hist + ev.bin, static_cast<double>(w));
        if (w2_live) atomicAdd(s_w2 + ev.bin, static_cast<double>(w * w));
      } else {
        atomicAdd(a.hist + ev.bin, static_cast<double>(w));
        if (w2_live) atomicAdd(a.w2 + ev.bin, static_cast<double>(w * w));
      }
    }
  }
  finish_block(a, s_hist, s_w2, reinterpret_cast<double*>(smem), &s_last);
}

template <class T>
__global__ void gather_kernel(T* dst, const T* src, const int32_t* perm, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = src[perm[i]];
}
// dst[i] = src[perm[i]] for 2-, 4- or 8-byte elements
cudaError_t launch_gather(void* dst, const void* src, const int32_t* perm, int64_t n, int elem_bytes, cudaStream_t s) {
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 148 * 8)));
  if (elem_bytes == 2) gather_kernel<<<grid, 256, 0, s>>>(static_cast<int16_t*>(dst), static_cast<const int16_t*>(src), perm, n);
  else if (elem_bytes == 4) gather_kernel<<<grid, 256, 0, s>>>(static_cast<int32_t*>(dst), static_cast<const int32_t*>(src), perm, n);
  else if (elem_bytes == 8) gather_kernel<<<grid, 256, 0, s>>>(static_cast<int64_t*>(dst), static_cast<const int64_t*>(src), perm, n);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_binned_eval(const FillArgs& a, int grid, cudaStream_t s) {
  if (a.real_f64) binned_eval_kernel_f64<<<grid, 256, (a.step.bytes + 15) & ~15, s>>>(a);
  else binned_eval_kernel<<<grid, 256, (a.step.bytes + 15) & ~15, s>>>(a);
  return cudaGetLastError();
}
// 1024 threads x 1 block/SM with an 8-column front is the product configuration; 512 x 2 and 256 x 2 (16-column front)
// remain selectable in the experiments build (M3B_BINNED_THREADS) for the A/B record
template <class F>
static cudaError_t with_fill_kernel(bool f64, int nt, F&& f) {
#ifdef M3B_EXPERIMENTS
  if (nt == 512) return f64 ? f(binned_fill_kernel<true, 512, 4>) : f(binned_fill_kernel<false, 512, 8>);
  if (nt == 256) return f64 ? f(binned_fill_kernel<true, 256, 16>) : f(binned_fill_kernel<false, 256, 16>);
  if (nt == 768) return f64 ? f(binned_fill_kernel<true, 768, 8>) : f(binned_fill_kernel<false, 768, 16>);
#endif
  (void)nt;
  return f64 ? f(binned_fill_kernel<true, 1024, 4>) : f(binned_fill_kernel<false, 1024, 8>);
}
cudaError_t launch_binned_fill(const FillArgs& a, int grid, int nt, int smem, cudaStream_t s) {
  return with_fill_kernel(a.real_f64 != 0, nt, [&](auto* k) {
    if (a.binned_pdl) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(nt); cfg.dynamicSmemBytes = static_cast<size_t>(smem); cfg.stream = s;
      cudaLaunchAttribute at{};
      at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at.val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = &at; cfg.numAttrs = 1;
      return cudaLaunchKernelEx(&cfg, k, a);
    }
    k<<<grid, nt, smem, s>>>(a);
    return cudaGetLastError();
  });
}
cudaError_t binned_fill_prepare(int smem, bool f64, int nt, int* bps) {
  return with_fill_kernel(f64, nt, [&](auto* k) {
    cudaError_t e = allow_max_dynamic_smem(k);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, k, nt, smem);
  });
}

}  // namespace m3b

// ------------------------------------------------------------------------------------------------
// host: uploads
// ------------------------------------------------------------------------------------------------
template <class R>
static int upload_binned_impl(m3b_handle* h, int32_t n_params, int32_t max_knots, const R* knot_x,
                                      const int16_t* n_pts, int64_t n_slots, const int32_t* uniquesplinevec_Monolith,
                                      const int32_t* coeffindexvec, int64_t n_unique, const int32_t* uniquecoeffindices,
                                      int64_t n_coeff, const R* manycoeff_arr, const R* xcoeff_arr) {
  constexpr bool F64 = std::is_same<R, double>::value;
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  REQUIRE(!h->splines_open && !h->splines_done && !h->binned, M3B_ERR_STATE, "m3b_upload_binned_splines: a spline handler is already uploaded");
  REQUIRE(n_params > 0 && n_params <= kMaxParams && max_knots >= 2 && knot_x && n_pts, M3B_ERR_INVALID, "m3b_upload_binned_splines: bad parameter layout");
  REQUIRE(n_slots > 0 && n_slots < (1ll << 31) && uniquesplinevec_Monolith && coeffindexvec, M3B_ERR_INVALID, "m3b_upload_binned_splines: bad slot arrays");
  REQUIRE(n_unique >= 0 && (n_unique == 0 || (uniquecoeffindices && manycoeff_arr && xcoeff_arr)), M3B_ERR_INVALID, "m3b_upload_binned_splines: null coefficient arrays");
  CK(cudaSetDevice(h->device));
  h->P = n_params; h->Kmax = max_knots;
  if constexpr (F64) { h->coeff_x_d.assign(knot_x, knot_x + static_cast<size_t>(n_params) * max_knots); h->f64 = true; }
  else h->coeff_x.assign(knot_x, knot_x + static_cast<size_t>(n_params) * max_knots);
  h->n_pts.assign(n_pts, n_pts + n_params);
  h->nseg.resize(n_params);
  for (int p = 0; p < n_params; ++p) {
    REQUIRE(n_pts[p] >= 0 && n_pts[p] <= max_knots, M3B_ERR_INVALID, "m3b_upload_binned_splines: n_pts[p] > max_knots");
    h->nseg[p] = static_cast<int16_t>(n_pts[p] > 1 ? n_pts[p] - 1 : 0);
  }
  h->curr_segment.assign(n_params, 0);
  h->segments.assign(n_params, 0);
  h->param_values.assign(n_params, -999.f);

  // group the non-flat splines by parameter (stable: ascending slot inside a parameter), pad to 256
  std::vector<int64_t> count(n_params + 1, 0);
  for (int64_t k = 0; k < n_unique; ++k) {
    const int32_t s = uniquecoeffindices[k];
    REQUIRE(s >= 0 && s < n_slots, M3B_ERR_INVALID, "m3b_upload_binned_splines: uniquecoeffindices out of range");
    const int32_t p = uniquesplinevec_Monolith[s];
    REQUIRE(p >= 0 && p < n_params, M3B_ERR_INVALID, "m3b_upload_binned_splines: uniquesplinevec_Monolith out of range");
    REQUIRE(h->nseg[p] > 0, M3B_ERR_KNOTS, "m3b_upload_binned_splines: non-flat spline on a parameter without knots");
    REQUIRE(coeffindexvec[s] >= 0 && static_cast<int64_t>(coeffindexvec[s]) + n_pts[p] <= n_coeff, M3B_ERR_KNOTS,
            "m3b_upload_binned_splines: a spline's knots run past the coefficient arrays (knot count differs from its parameter's)");
    ++count[p + 1];
  }
  std::vector<int64_t> out_base(n_params + 1, 0), coef_base(n_params + 1, 0), npad(n_params, 0);
  for (int p = 0; p < n_params; ++p) {
    npad[p] = (count[p + 1] + kBTileSplines - 1) / kBTileSplines * kBTileSplines;
    out_base[p + 1] = out_base[p] + npad[p];
    coef_base[p + 1] = coef_base[p] + npad[p] * h->nseg[p];
  }
  const int64_t n_act_pad = out_base[n_params], n_coef_dev = coef_base[n_params];
  h->b_slot2compact.assign(static_cast<size_t>(n_slots), -1);
  h->b_compact2slot.assign(static_cast<size_t>(n_act_pad), -1);
  std::vector<R> coef(static_cast<size_t>(n_coef_dev) * 4, R(0));       // {y,b,c,d} per element; padding = the constant 1
  for (int64_t i = 0; i < n_coef_dev; ++i) coef[4 * i] = R(1);
  std::vector<R> xs(static_cast<size_t>(n_coef_dev), R(0));
  std::vector<int64_t> fill(n_params, 0);
  for (int64_t k = 0; k < n_unique; ++k) {
    const int32_t s = uniquecoeffindices[k];
    const int32_t p = uniquesplinevec_Monolith[s];
    const int64_t j = fill[p]++;
    REQUIRE(h->b_slot2compact[s] < 0, M3B_ERR_INVALID, "m3b_upload_binned_splines: slot listed twice in uniquecoeffindices");
    h->b_slot2compact[s] = static_cast<int32_t>(out_base[p] + j);
    h->b_compact2slot[out_base[p] + j] = s;
    const R* src = manycoeff_arr + 4 * static_cast<int64_t>(coeffindexvec[s]);
    for (int g = 0; g < h->nseg[p]; ++g) {
      for (int q = 0; q < 4; ++q) coef[4 * (coef_base[p] + g * npad[p] + j) + q] = src[4 * g + q];
      xs[coef_base[p] + g * npad[p] + j] = xcoeff_arr[coeffindexvec[s] + g];
    }
  }
  std::vector<BTile> tiles;
  for (int p = 0; p < n_params; ++p)
    for (int64_t k0 = 0; k0 < npad[p]; k0 += kBTileSplines)
      tiles.push_back(BTile{coef_base[p], static_cast<int32_t>(npad[p]), p, static_cast<int32_t>(k0), static_cast<int32_t>(out_base[p] + k0)});
  if constexpr (F64) {
    CK(dev_upload(h, &h->d_bcoef_d, coef));
    CK(dev_upload(h, &h->d_bx_d, xs));
    CK(dev_alloc(h, &h->d_bw_d, static_cast<size_t>(n_act_pad)));
  } else {
    float4* dc = nullptr;
    CK(dev_alloc(h, &dc, static_cast<size_t>(n_coef_dev)));
    CK(copy_sync(h, dc, coef.data(), sizeof(float) * coef.size(), cudaMemcpyHostToDevice));
    h->d_bcoef = dc;
    CK(dev_upload(h, &h->d_bx, xs));
    CK(dev_alloc(h, &h->d_bw, static_cast<size_t>(n_act_pad)));
  }
  CK(dev_upload(h, &h->d_btiles, tiles));
  h->n_btiles = static_cast<int32_t>(tiles.size());
  h->b_n_slots = n_slots; h->b_n_act = n_unique; h->b_n_act_pad = n_act_pad;
  h->b_out_base = out_base; h->b_count.assign(count.begin() + 1, count.end());
  h->b_run_start.clear(); h->b_run_super.clear();
  for (int64_t s = 0, super = 0; s < n_slots; ++s) {
    if (s > 0 && uniquesplinevec_Monolith[s] == uniquesplinevec_Monolith[s - 1]) continue;
    if (s > 0 && uniquesplinevec_Monolith[s] < uniquesplinevec_Monolith[s - 1]) ++super;
    h->b_run_start.push_back(s); h->b_run_super.push_back(static_cast<int32_t>(super));
  }
  h->binned = true;
  h->launch_ready = false;
  return M3B_OK;
}


// The event table (bin, static weight, norm indices, oscillation index) once more in the order the fill kernel walks the
// events, so a warp reads 32 consecutive elements of each instead of 32 sectors.  The caller-indexed arrays stay the
// masters (rebinning, selection and re-uploads write those); these copies are refreshed from them.
static int sort_static_weights(m3b_handle* h) {
  const void* src = h->f64 ? static_cast<const void*>(h->d_static_d) : static_cast<const void*>(h->d_static);
  if (!src || !h->d_perm) return M3B_OK;
  const int elem = h->f64 ? 8 : 4;
  if (!h->d_static_sorted) {
    unsigned char* p = nullptr;
    CK(dev_alloc(h, &p, static_cast<size_t>(h->e_pad) * elem));
    h->d_static_sorted = p;
  }
  CK(launch_gather(h->d_static_sorted, src, h->d_perm, h->e_pad, elem, h->stream));
  return M3B_OK;
}
static int sort_event_table(m3b_handle* h) {
  CK(dev_alloc(h, &h->d_bin_sorted, static_cast<size_t>(h->e_pad)));
  CK(launch_gather(h->d_bin_sorted, h->d_bin, h->d_perm, h->e_pad, 4, h->stream));
  if (h->norm_slots > 0) {
    CK(dev_alloc(h, &h->d_norm_idx_sorted, static_cast<size_t>(h->e_pad) * h->norm_slots));
    for (int j = 0; j < h->norm_slots; ++j)
      CK(launch_gather(h->d_norm_idx_sorted + static_cast<int64_t>(j) * h->e_pad, h->d_norm_idx + static_cast<int64_t>(j) * h->e_pad, h->d_perm, h->e_pad, 2, h->stream));
  }
  if (h->use_osc) {
    // without an index array the oscillation weights are per event: the index is the event itself (padding: none)
    CK(dev_alloc(h, &h->d_osc_idx_sorted, static_cast<size_t>(h->e_pad)));
    if (h->d_osc_idx) CK(launch_gather(h->d_osc_idx_sorted, h->d_osc_idx, h->d_perm, h->e_pad, 4, h->stream));
    else {
      std::vector<int32_t> oi(static_cast<size_t>(h->e_pad), -1);
      CK(cudaMemcpy(oi.data(), h->d_perm, sizeof(int32_t) * h->n_events, cudaMemcpyDeviceToHost));
      CK(copy_sync(h, h->d_osc_idx_sorted, oi.data(), sizeof(int32_t) * oi.size(), cudaMemcpyHostToDevice));
    }
  }
  return sort_static_weights(h);
}

extern "C" {

M3B_API int m3b_upload_binned_splines(m3b_handle* h, int32_t n_params, int32_t max_knots, const float* knot_x,
                                      const int16_t* n_pts, int64_t n_slots, const int32_t* uniquesplinevec_Monolith,
                                      const int32_t* coeffindexvec, int64_t n_unique, const int32_t* uniquecoeffindices,
                                      int64_t n_coeff, const float* manycoeff_arr, const float* xcoeff_arr) {
  return upload_binned_impl<float>(h, n_params, max_knots, knot_x, n_pts, n_slots, uniquesplinevec_Monolith, coeffindexvec, n_unique,
                                   uniquecoeffindices, n_coeff, manycoeff_arr, xcoeff_arr);
}
M3B_API int m3b_upload_binned_splines_f64(m3b_handle* h, int32_t n_params, int32_t max_knots, const double* knot_x,
                                          const int16_t* n_pts, int64_t n_slots, const int32_t* uniquesplinevec_Monolith,
                                          const int32_t* coeffindexvec, int64_t n_unique, const int32_t* uniquecoeffindices,
                                          int64_t n_coeff, const double* manycoeff_arr, const double* xcoeff_arr) {
  return upload_binned_impl<double>(h, n_params, max_knots, knot_x, n_pts, n_slots, uniquesplinevec_Monolith, coeffindexvec, n_unique,
                                    uniquecoeffindices, n_coeff, manycoeff_arr, xcoeff_arr);
}

M3B_API int m3b_upload_event_binned_splines(m3b_handle* h, int64_t n_events, const uint32_t* n_per_event,
                                            const int32_t* spline_index) {
  REQUIRE(h && n_per_event, M3B_ERR_INVALID, "m3b_upload_event_binned_splines: null argument");
  REQUIRE(h->binned, M3B_ERR_STATE, "m3b_upload_event_binned_splines: upload the binned splines first");
  REQUIRE(h->n_events > 0 && n_events == h->n_events, M3B_ERR_STATE, "m3b_upload_event_binned_splines: upload the events first (same count)");
  REQUIRE(!h->d_wtiles, M3B_ERR_STATE, "m3b_upload_event_binned_splines: already uploaded");
  CK(cudaSetDevice(h->device));
  const int64_t n_wt = h->e_pad / 32;
  std::vector<WTile> wt(static_cast<size_t>(n_wt));
  // first pass: non-flat pointers per event, and where in its parameter's weight row the event's first one sits
  std::vector<uint32_t> keep(static_cast<size_t>(n_events), 0);
  std::vector<uint64_t> first(static_cast<size_t>(n_events) + 1, 0);
  // Sort key = the spline-grid cell of the event.  Slots are laid out [sample][osc][syst][mode][var1][var2][var3]
  // (Splines/BinnedSplineHandler.h:110) and an event's pointers are pushed syst by syst (SampleHandlerFD.cpp:1196-1242), all
  // into the same (mode, var1, var2, var3) cell of one (sample, osc) super-block: the offset of the event's first pointer
  // (flat or not) inside its systematic's block IS that cell.  With the events in (super-block, cell) order, the pointers of
  // neighbouring events sit next to each other in every systematic's weight row.  (Only the walking order -- and so the
  // speed -- depends on this reading of the layout; the result does not.)
  std::vector<int64_t> cell(static_cast<size_t>(n_events), INT64_MAX);     // INT64_MAX: no pointer at all
  uint64_t off = 0;
  for (int64_t e = 0; e < n_events; ++e) {
    uint32_t k = 0;
    first[e] = off;
    for (uint32_t j = 0; j < n_per_event[e]; ++j) {
      const int32_t s = spline_index[off + j];
      REQUIRE(s >= 0 && s < h->b_n_slots, M3B_ERR_INVALID, "m3b_upload_event_binned_splines: spline_index out of range");
      if (j == 0) {
        const size_t r = static_cast<size_t>(std::upper_bound(h->b_run_start.begin(), h->b_run_start.end(), static_cast<int64_t>(s)) - h->b_run_start.begin()) - 1;
        cell[e] = (static_cast<int64_t>(h->b_run_super[r]) << 40) | (s - h->b_run_start[r]);
      }
      const int32_t c = h->b_slot2compact[s];
      if (c < 0) continue;                       // flat splines hold exactly 1.0f: multiplying by them changes nothing
      ++k;
    }
    keep[e] = k;
    off += n_per_event[e];
  }
  first[n_events] = off;
  // processing order: events sorted by cell (stable).  Only the ORDER in which the fill kernel walks the events changes;
  // every per-event array stays indexed by the caller's event number.
  std::vector<int32_t> perm(static_cast<size_t>(h->e_pad));
  for (int64_t e = 0; e < h->e_pad; ++e) perm[e] = static_cast<int32_t>(e);
  std::stable_sort(perm.begin(), perm.begin() + n_events, [&](int32_t x, int32_t y) { return cell[x] < cell[y]; });
  // inside a span of 8 warp tiles (events that are in flight together anyway and share L1 lines) the order is free: put
  // events with equally many non-flat pointers into the same warp, so a tile's ELL columns (its longest event) are mostly full
  constexpr int kBSortSpan = 8;
  for (int64_t v0 = 0; v0 < n_events; v0 += 32 * kBSortSpan)
    std::stable_sort(perm.begin() + v0, perm.begin() + std::min<int64_t>(n_events, v0 + 32 * kBSortSpan),
                     [&](int32_t x, int32_t y) { return keep[x] > keep[y]; });
  int64_t total = 0;
  for (int64_t t = 0; t < n_wt; ++t) {
    uint32_t mx = 0;
    for (int64_t v = t * 32; v < std::min<int64_t>(n_events, t * 32 + 32); ++v) mx = std::max(mx, keep[perm[v]]);
    wt[t].off = total; wt[t].max_n = static_cast<int32_t>(mx); wt[t].pad = 0;
    total += static_cast<int64_t>(mx) * 32;
  }
  std::vector<int32_t> ell(static_cast<size_t>(std::max<int64_t>(total, 1)), -1);
  for (int64_t v = 0; v < n_events; ++v) {
    const int64_t e = perm[v];
    const WTile& d = wt[v / 32];
    int64_t k = 0;
    for (uint64_t j = first[e]; j < first[e + 1]; ++j) {
      const int32_t c = h->b_slot2compact[spline_index[j]];
      if (c >= 0) ell[d.off + (k++) * 32 + (v & 31)] = c;       // pointer order kept: the product is sequential
    }
  }
  CK(dev_upload(h, &h->d_perm, perm));
  if (experiment_env("M3B_VERBOSE")) fprintf(stderr, "[m3b] binned ELL entries %lld\n", static_cast<long long>(total));
  if (h->f64) {
    // default build: osc / static weights and the per-event outputs are M3::float_t = double.  Until the caller
    // supplies doubles (m3b_upload_event_weights_f64 / m3b_upload_osc_f64) the float uploads are widened (exact).
    if (h->d_static && !h->d_static_d) {
      std::vector<float> f(static_cast<size_t>(h->e_pad));
      CK(cudaMemcpy(f.data(), h->d_static, sizeof(float) * f.size(), cudaMemcpyDeviceToHost));
      std::vector<double> dd(f.begin(), f.end());
      CK(dev_upload(h, &h->d_static_d, dd));
    }
    if (h->use_osc && !h->d_osc_d) {
      std::vector<double> ones(static_cast<size_t>(h->n_osc), 1.0);
      CK(dev_upload(h, &h->d_osc_d, ones));
    }
    if ((h->cfg.flags & M3B_FLAG_KEEP_EVENT_WEIGHTS) && !h->d_evt_spline_d) {
      CK(dev_alloc(h, &h->d_evt_spline_d, static_cast<size_t>(h->e_pad)));
      CK(dev_alloc(h, &h->d_evt_total_d, static_cast<size_t>(h->e_pad)));
    }
  }
  CK(dev_upload(h, &h->d_ell, ell));
  CK(dev_upload(h, &h->d_wtiles, wt));
  if (int rc = sort_event_table(h)) return rc;
  h->n_wtiles = n_wt;
  h->b_gather_per_step = static_cast<uint64_t>(total);
  h->launch_ready = false;
  return M3B_OK;
}

// ---- default (double) build: inputs and mirrors in M3::float_t = double
M3B_API int m3b_upload_event_weights_f64(m3b_handle* h, int64_t n_events, const double* static_w) {
  REQUIRE(h && static_w, M3B_ERR_INVALID, "m3b_upload_event_weights_f64: null argument");
  REQUIRE(h->f64, M3B_ERR_STATE, "m3b_upload_event_weights_f64: the handle does not run the double build (m3b_upload_binned_splines_f64)");
  REQUIRE(h->n_events > 0 && n_events == h->n_events, M3B_ERR_STATE, "m3b_upload_event_weights_f64: upload the events first (same count)");
  CK(cudaSetDevice(h->device));
  std::vector<double> sw(static_cast<size_t>(h->e_pad), 1.0);
  std::copy(static_w, static_w + n_events, sw.begin());
  if (!h->d_static_d) CK(dev_alloc(h, &h->d_static_d, sw.size()));
  CK(copy_sync(h, h->d_static_d, sw.data(), sizeof(double) * sw.size(), cudaMemcpyHostToDevice));
  return sort_static_weights(h);
}
M3B_API int m3b_upload_osc_f64(m3b_handle* h, const double* osc_w, int64_t n) {
  REQUIRE(h && osc_w, M3B_ERR_INVALID, "m3b_upload_osc_f64: null argument");
  REQUIRE(h->f64, M3B_ERR_STATE, "m3b_upload_osc_f64: the handle does not run the double build");
  REQUIRE(h->use_osc && n == h->n_osc, M3B_ERR_INVALID, "m3b_upload_osc_f64: length differs from the oscillation-weight array's");
  CK(cudaSetDevice(h->device));
  if (!h->d_osc_d) CK(dev_alloc(h, &h->d_osc_d, static_cast<size_t>(n)));
  CK(cudaMemcpyAsync(h->d_osc_d, osc_w, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  return M3B_OK;
}
M3B_API int m3b_read_binned_weights_f64(m3b_handle* h, double* weightvec_Monolith) {
  REQUIRE(h && weightvec_Monolith, M3B_ERR_INVALID, "m3b_read_binned_weights_f64: null argument");
  REQUIRE(h->binned && h->f64 && h->steps > 0, M3B_ERR_STATE, "m3b_read_binned_weights_f64: no double-build binned step yet");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  std::vector<double> bw(static_cast<size_t>(h->b_n_act_pad));
  CK(cudaMemcpy(bw.data(), h->d_bw_d, sizeof(double) * bw.size(), cudaMemcpyDeviceToHost));
  for (int64_t s = 0; s < h->b_n_slots; ++s) weightvec_Monolith[s] = 1.0;
  for (int64_t c = 0; c < h->b_n_act_pad; ++c)
    if (h->b_compact2slot[c] >= 0) weightvec_Monolith[h->b_compact2slot[c]] = bw[c];
  return M3B_OK;
}
M3B_API int m3b_read_event_weights_f64(m3b_handle* h, double* spline_w, double* total_w) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  REQUIRE(h->f64 && h->steps > 0 && h->d_evt_spline_d, M3B_ERR_STATE,
          "m3b_read_event_weights_f64: needs the double build, M3B_FLAG_KEEP_EVENT_WEIGHTS and a step");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (spline_w) CK(cudaMemcpy(spline_w, h->d_evt_spline_d, sizeof(double) * h->n_events, cudaMemcpyDeviceToHost));
  if (total_w) CK(cudaMemcpy(total_w, h->d_evt_total_d, sizeof(double) * h->n_events, cudaMemcpyDeviceToHost));
  return M3B_OK;
}

// BinnedSplineHandler::weightvec_Monolith as the host sees it (retPointer targets): 1.0 for flat slots
M3B_API int m3b_read_binned_weights(m3b_handle* h, float* weightvec_Monolith) {
  REQUIRE(h && weightvec_Monolith, M3B_ERR_INVALID, "m3b_read_binned_weights: null argument");
  REQUIRE(h->binned && !h->f64 && h->steps > 0, M3B_ERR_STATE, "m3b_read_binned_weights: no float-build binned step yet");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  std::vector<float> bw(static_cast<size_t>(h->b_n_act_pad));
  CK(cudaMemcpy(bw.data(), h->d_bw, sizeof(float) * bw.size(), cudaMemcpyDeviceToHost));
  for (int64_t s = 0; s < h->b_n_slots; ++s) weightvec_Monolith[s] = 1.0f;
  for (int64_t c = 0; c < h->b_n_act_pad; ++c)
    if (h->b_compact2slot[c] >= 0) weightvec_Monolith[h->b_compact2slot[c]] = bw[c];
  return M3B_OK;
}

}  // extern "C"
