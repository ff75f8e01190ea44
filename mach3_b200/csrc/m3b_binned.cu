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

template <bool F64>
__global__ void __launch_bounds__(256, 2) binned_fill_kernel(const __grid_constant__ FillArgs a) {
  using R = typename std::conditional<F64, double, float>::type;      // M3::float_t of the build
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
    for (int i = tid; i < a.n_bins; i += 256) s_hist[i] = 0.;
    if (w2_live) for (int i = tid; i < a.n_bins; i += 256) s_w2[i] = 0.;
  }
  stage_step_table(a, smem, &bar);
  const R* norm = reinterpret_cast<const R*>(smem + (F64 ? a.step.off_norm_d : a.step.off_norm));
  const R* bw = F64 ? reinterpret_cast<const R*>(a.bw_d) : reinterpret_cast<const R*>(a.bw);
  const R* oscp = F64 ? reinterpret_cast<const R*>(a.osc_d) : reinterpret_cast<const R*>(a.osc);
  const R* statp = F64 ? reinterpret_cast<const R*>(a.static_d) : reinterpret_cast<const R*>(a.static_w);

  // grid-strided walk over the warp tiles (measured against contiguous runs of tiles per block, which would let a block
  // re-use the L1 lines of its previous iteration: 249 vs 299 us per config-4 step -- the kernel is bound by the number
  // of distinct lines a gather instruction touches, not by L1 capacity; profiles/r02_binned_fill_ab.txt)
  const int64_t tiles_per_block = (a.n_wtiles + gridDim.x - 1) / gridDim.x;
  const int64_t wt_begin = a.binned_contiguous ? static_cast<int64_t>(blockIdx.x) * tiles_per_block : static_cast<int64_t>(blockIdx.x) * 8;
  const int64_t wt_end = a.binned_contiguous ? (wt_begin + tiles_per_block < a.n_wtiles ? wt_begin + tiles_per_block : a.n_wtiles) : a.n_wtiles;
  const int64_t wt_step = a.binned_contiguous ? 8 : static_cast<int64_t>(gridDim.x) * 8;
  for (int64_t wt = wt_begin + warp; wt < wt_end; wt += wt_step) {
    const WTile d = a.wtiles[wt];
    // events are processed in the order of their spline-grid cell (a.perm, built at upload): the 32 lanes of a warp --
    // and the 8 warps of the block -- then gather from the same few sectors of every parameter's weight row
    const int64_t e = a.perm ? static_cast<int64_t>(a.perm[wt * 32 + lane]) : wt * 32 + lane;
    const int bin = a.bin[e];
    R w_osc = 1, w_static = 1;
    if (oscp) {
      const int64_t oi = a.osc_idx ? static_cast<int64_t>(a.osc_idx[e]) : (e < a.n_events ? e : 0);
      w_osc = oi >= 0 ? oscp[oi] : R(1);
    }
    if (statp) w_static = statp[e];
    // CalcWeightTotal: norms first, then the weight pointers in push order: osc, binned splines, extras
    R w = 1;
    for (int j = 0; j < a.norm_slots; ++j) {
      const int i = a.norm_idx[static_cast<int64_t>(j) * a.e_pad + e];
      w *= (i >= 0 ? norm[i] : R(1));
    }
    w *= w_osc;
    R w_spl = 1;               // product of the binned weights alone (m3b_read_event_weights)
    const int32_t* col = a.ell + d.off + lane;
    for (int j0 = 0; j0 < d.max_n; j0 += 8) {
      int idx[8]; R g[8];
      #pragma unroll
      for (int j = 0; j < 8; ++j) idx[j] = (j0 + j < d.max_n) ? __ldcs(col + (j0 + j) * 32) : -1;
      #pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = idx[j] >= 0 ? __ldg(bw + idx[j]) : R(1);
      #pragma unroll
      for (int j = 0; j < 8; ++j) if (idx[j] >= 0) { w *= g[j]; w_spl *= g[j]; }
    }
    w *= w_static;
    if (e < a.n_events) {
      if (F64) { if (a.evt_spline_d) { a.evt_spline_d[e] = w_spl; a.evt_total_d[e] = w; } }
      else if (a.evt_spline_w) { a.evt_spline_w[e] = static_cast<float>(w_spl); a.evt_total_w[e] = static_cast<float>(w); }
    }
    if (w > R(0) && bin >= 0 && !a.weights_only) {
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

cudaError_t launch_binned_eval(const FillArgs& a, int grid, cudaStream_t s) {
  if (a.real_f64) binned_eval_kernel_f64<<<grid, 256, (a.step.bytes + 15) & ~15, s>>>(a);
  else binned_eval_kernel<<<grid, 256, (a.step.bytes + 15) & ~15, s>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_binned_fill(const FillArgs& a, int grid, int smem, cudaStream_t s) {
  if (a.real_f64) binned_fill_kernel<true><<<grid, 256, smem, s>>>(a);
  else binned_fill_kernel<false><<<grid, 256, smem, s>>>(a);
  return cudaGetLastError();
}
cudaError_t binned_fill_set_smem(int smem) {
  (void)smem;
  cudaError_t e = allow_max_dynamic_smem(binned_fill_kernel<false>);
  if (e != cudaSuccess) return e;
  return allow_max_dynamic_smem(binned_fill_kernel<true>);
}
cudaError_t binned_fill_occupancy(int smem, bool f64, int* bps) {
  return f64 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, binned_fill_kernel<true>, 256, smem)
             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, binned_fill_kernel<false>, 256, smem);
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
  h->binned = true;
  h->launch_ready = false;
  return M3B_OK;
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
  std::vector<float> cell(static_cast<size_t>(n_events), 2.f);      // position of the event's spline-grid cell, 0..1 (2: no non-flat pointer)
  uint64_t off = 0;
  for (int64_t e = 0; e < n_events; ++e) {
    uint32_t k = 0;
    first[e] = off;
    for (uint32_t j = 0; j < n_per_event[e]; ++j) {
      const int32_t s = spline_index[off + j];
      REQUIRE(s >= 0 && s < h->b_n_slots, M3B_ERR_INVALID, "m3b_upload_event_binned_splines: spline_index out of range");
      const int32_t c = h->b_slot2compact[s];
      if (c < 0) continue;                       // flat splines hold exactly 1.0f: multiplying by them changes nothing
      if (k == 0) {
        // compact indices are grouped by parameter and ascend with the slot inside a parameter, i.e. with the spline-grid
        // cell (Splines/BinnedSplineHandler.h:110: [..][syst][mode][var1][var2][var3]); the rank inside the parameter's
        // row, as a fraction, is comparable between events whose first non-flat pointer belongs to different parameters
        const int p = h->b_param_of_compact_row(c);
        cell[e] = static_cast<float>(static_cast<double>(c - h->b_out_base[p]) / static_cast<double>(std::max<int64_t>(1, h->b_count[p])));
      }
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
  return M3B_OK;
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
