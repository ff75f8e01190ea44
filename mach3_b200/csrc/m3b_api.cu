// m3b_api.cu -- host side of libm3b200.so: the C ABI declared in include/m3b200.h.
// Owns all device state; re-tiles the reference's AoS monolith into the tiled SoA layout;
// mirrors SplineBase::FindSplineSegment on the host (O(nParams), history-dependent); enqueues one
// fused kernel per step.  No CPU fallback: every compute entry point needs the sm_100 device.
#include "m3b_handle.h"

static thread_local std::string g_last_error;
std::string& m3b_last_error_slot() { return g_last_error; }


extern "C" {

M3B_API int m3b_abi_version(void) { return 1; }

M3B_API const char* m3b_last_error(const m3b_handle* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

M3B_API int m3b_create(const m3b_config* cfg, m3b_handle** out) {
  m3b_handle* h = nullptr;
  if (!cfg || !out) return fail(nullptr, M3B_ERR_INVALID, "m3b_create: null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, M3B_ERR_NODEVICE, std::string("m3b_create: no CUDA device (") + cudaGetErrorString(e) +
                                               "); libm3b200 has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, M3B_ERR_INVALID, "m3b_create: bad device ordinal");
  cudaDeviceProp prop{};
  e = cudaGetDeviceProperties(&prop, cfg->device);
  if (e != cudaSuccess) return fail(nullptr, M3B_ERR_CUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, M3B_ERR_NODEVICE, std::string("m3b_create: device '") + prop.name +
                                               "' is not sm_100 (kernels are built for sm_100a only)");
  const int T = cfg->tile_events == 0 ? 1024 : cfg->tile_events;
  if (T != 256 && T != 512 && T != 1024)
    return fail(nullptr, M3B_ERR_INVALID, "m3b_create: tile_events must be 0 (auto), 256, 512 or 1024");
  h = new m3b_handle();
  h->cfg = *cfg;
  h->device = cfg->device;
  h->sm_count = prop.multiProcessorCount;
  h->T = T;
  h->T_auto = cfg->tile_events == 0;
  h->test_stat = cfg->test_statistic;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->own_stream = true;
  CK(dev_alloc(h, &h->d_ticket, 1));
  CK(cudaMemsetAsync(h->d_ticket, 0, sizeof(unsigned int), h->stream));
  CK(dev_alloc(h, &h->d_tile_counter, 1));
  CK(cudaMemsetAsync(h->d_tile_counter, 0, sizeof(unsigned int), h->stream));
  CK(dev_alloc(h, &h->d_status, 1));
  CK(cudaMemsetAsync(h->d_status, 0, sizeof(int32_t), h->stream));
  for (int i = 0; i < m3b_handle::kRing; ++i) CK(cudaEventCreateWithFlags(&h->step_ev[i], cudaEventDisableTiming));
  *out = h;
  return M3B_OK;
}

M3B_API void m3b_destroy(m3b_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (void* p : h->ipc_opened) cudaIpcCloseMemHandle(p);
  for (void* p : h->registered) cudaHostUnregister(p);
  for (void* p : h->host_allocs) cudaFreeHost(p);
  for (void* p : h->allocs) cudaFree(p);
  for (cudaEvent_t e : h->tev) cudaEventDestroy(e);
  for (int i = 0; i < m3b_handle::kRing; ++i) {
    if (h->h_step[i]) cudaFreeHost(h->h_step[i]);
    if (h->step_ev[i]) cudaEventDestroy(h->step_ev[i]);
  }
  if (h->h_llh) cudaFreeHost(h->h_llh);
  if (h->h_seq) cudaFreeHost(h->h_seq);
  if (h->h_batch) cudaFreeHost(h->h_batch);
  for (void* p : {h->bt_dx, h->bt_rowoff, h->bt_val, h->bt_rowlist, h->bt_norm, h->bt_sigs, h->bt_hist, h->bt_llh, h->bt_slot, h->bt_group, h->bt_rank8}) if (p) cudaFree(p);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

M3B_API int m3b_set_stream(m3b_handle* h, void* s) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (h->own_stream && h->stream) { cudaStreamDestroy(h->stream); h->own_stream = false; }
  h->stream = static_cast<cudaStream_t>(s);
  return M3B_OK;
}

// ------------------------------------------------------------------------------------------------
// splines
// ------------------------------------------------------------------------------------------------
M3B_API int m3b_splines_begin(m3b_handle* h, int32_t n_params, int32_t max_knots, const float* coeff_x,
                              const int16_t* n_pts, int64_t n_events_total) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  REQUIRE(!h->splines_open && !h->splines_done, M3B_ERR_STATE, "m3b_splines_begin: monolith already uploaded");
  REQUIRE(n_params > 0 && n_params <= kMaxParams, M3B_ERR_INVALID, "m3b_splines_begin: n_params out of range");
  REQUIRE(max_knots >= 0 && coeff_x && n_pts && n_events_total >= 0, M3B_ERR_INVALID, "m3b_splines_begin: bad argument");
  // tile row length, when the caller left it open: 16 KB rows stream best, but with few tiles per SM
  // the end-of-grid balance matters more and 8 KB rows win (measured at cfg2: 1 M events, 26 units per SM)
  if (h->T_auto && h->n_events == 0) h->T = n_events_total >= 1500000 ? 1024 : 512;
  h->P = n_params; h->Kmax = max_knots;
  h->coeff_x.assign(coeff_x, coeff_x + static_cast<size_t>(n_params) * max_knots);
  h->n_pts.assign(n_pts, n_pts + n_params);
  h->nseg.resize(n_params);
  for (int p = 0; p < n_params; ++p) {
    REQUIRE(n_pts[p] >= 0 && n_pts[p] <= max_knots, M3B_ERR_INVALID, "m3b_splines_begin: n_pts[p] > max_knots");
    h->nseg[p] = static_cast<int16_t>(n_pts[p] > 1 ? n_pts[p] - 1 : (n_pts[p] == 1 ? 1 : 0));
  }
  h->curr_segment.assign(n_params, 0);
  h->segments.assign(n_params, 0);                 // Splines/SplineMonolith.cpp:91-95
  h->param_values.assign(n_params, -999.f);
  h->n_events_total = n_events_total;
  h->n_events_loaded = 0;
  h->splines_open = true;
  return M3B_OK;
}

static int signature_of(m3b_handle* h, const std::vector<int16_t>& cub, const std::vector<int16_t>& lin) {
  std::vector<int16_t> key(cub);
  key.push_back(-1);
  key.insert(key.end(), lin.begin(), lin.end());
  auto it = h->sig_index.find(key);
  if (it != h->sig_index.end()) return it->second;
  const int id = static_cast<int>(h->sigs.size());
  SigDesc sd{};
  sd.nc = static_cast<int32_t>(cub.size());
  sd.nl = static_cast<int32_t>(lin.size());
  sd.off = static_cast<int32_t>(h->sig_pool.size());
  h->sig_slot_of_param.resize(static_cast<size_t>(id + 1) * h->P, -1);
  h->sig_segbase_of_param.resize(static_cast<size_t>(id + 1) * h->P, 0);
  for (int16_t p : cub) h->sig_pool.push_back(p);
  int rows = 0;
  for (size_t s = 0; s < cub.size(); ++s) {
    h->sig_pool.push_back(rows);
    h->sig_slot_of_param[static_cast<size_t>(id) * h->P + cub[s]] = static_cast<int16_t>(s);
    h->sig_segbase_of_param[static_cast<size_t>(id) * h->P + cub[s]] = rows;
    rows += h->nseg[cub[s]];
  }
  for (size_t s = 0; s < lin.size(); ++s) {
    h->sig_pool.push_back(lin[s]);
    h->sig_slot_of_param[static_cast<size_t>(id) * h->P + lin[s]] = static_cast<int16_t>(s);
  }
  sd.rows = rows;
  h->sigs.push_back(sd);
  h->sig_index.emplace(std::move(key), id);
  h->max_nc = std::max(h->max_nc, sd.nc);
  h->max_nl = std::max(h->max_nl, sd.nl);
  return id;
}

// one internal chunk: events [0,n) whose first event sits on a tile boundary
static int append_chunk(m3b_handle* h, int64_t n, const uint32_t* cnt_c, const int16_t* paramNo,
                        const uint64_t* knot_off, uint64_t total_knots, const float* coeff_many,
                        const uint32_t* cnt_l, const int16_t* paramNo_l, const float* coeff_l) {
  const int T = h->T, P = h->P;
  std::vector<uint64_t> start_c(n + 1), start_l(n + 1);
  start_c[0] = 0; start_l[0] = 0;
  for (int64_t i = 0; i < n; ++i) {
    start_c[i + 1] = start_c[i] + cnt_c[2 * i];
    start_l[i + 1] = start_l[i] + cnt_l[2 * i];
  }
  const uint64_t tot_c = start_c[n], tot_l = start_l[n];
  // validate: parameter ids, knot counts (the reference assumes one knot set per parameter)
  for (uint64_t s = 0; s < tot_c; ++s) {
    const int p = paramNo[s];
    REQUIRE(p >= 0 && p < P, M3B_ERR_INVALID, "m3b_splines_append: paramNo_arr out of range");
    uint64_t nk = (s + 1 < tot_c ? knot_off[s + 1] : total_knots) - knot_off[s];
    // the last used response: SMonolith sizes coeff_many from ScanMasterSpline's count, which includes the one-knot
    // splines PrepareForGPU then skips (Splines/SplineMonolith.cpp:370,151), so unused knots may follow it
    if (s + 1 == tot_c && nk > static_cast<uint64_t>(h->n_pts[p]) && knot_off[s] <= total_knots) nk = static_cast<uint64_t>(h->n_pts[p]);
    if (nk != static_cast<uint64_t>(h->n_pts[p])) {
      char b[256];
      snprintf(b, sizeof b, "m3b_splines_append: response %llu of parameter %d has %llu knots, parameter has %d",
               (unsigned long long)s, p, (unsigned long long)nk, (int)h->n_pts[p]);
      return fail(h, M3B_ERR_KNOTS, b);
    }
  }
  for (uint64_t s = 0; s < tot_l; ++s)
    REQUIRE(paramNo_l[s] >= 0 && paramNo_l[s] < P, M3B_ERR_INVALID, "m3b_splines_append: paramNo_tf1 out of range");

  // signatures: union of the parameters of each tile's events
  const int64_t ntile = (n + T - 1) / T;
  std::vector<int32_t> tile_sig(ntile);
  std::vector<uint64_t> tile_cub_off(ntile), tile_lin_off(ntile);
  std::vector<unsigned char> seen(P);
  uint64_t cub_total = 0, lin_total = 0;
  bool all_full = true;
  std::vector<int16_t> cub, lin;
  for (int64_t t = 0; t < ntile; ++t) {
    const int64_t e0 = t * T, e1 = std::min<int64_t>(n, e0 + T);
    std::fill(seen.begin(), seen.end(), 0);
    for (uint64_t s = start_c[e0]; s < start_c[e1]; ++s) seen[paramNo[s]] |= 1;
    for (uint64_t s = start_l[e0]; s < start_l[e1]; ++s) seen[paramNo_l[s]] |= 2;
    cub.clear(); lin.clear();
    for (int p = 0; p < P; ++p) {
      REQUIRE(seen[p] != 3, M3B_ERR_INVALID, "m3b_splines_append: parameter used both as TSpline3 and TF1");
      if (seen[p] & 1) cub.push_back(static_cast<int16_t>(p));
      if (seen[p] & 2) lin.push_back(static_cast<int16_t>(p));
    }
    const int sig = signature_of(h, cub, lin);
    tile_sig[t] = sig;
    tile_cub_off[t] = cub_total;
    tile_lin_off[t] = lin_total;
    cub_total += static_cast<uint64_t>(h->sigs[sig].rows) * T;
    lin_total += static_cast<uint64_t>(h->sigs[sig].nl) * T;
    const uint64_t full_c = static_cast<uint64_t>(cub.size()) * T, full_l = static_cast<uint64_t>(lin.size()) * T;
    if (e1 - e0 != T || start_c[e1] - start_c[e0] != full_c || start_l[e1] - start_l[e0] != full_l) all_full = false;
    h->active_coef_bytes += (16ull * cub.size() + 8ull * lin.size()) * T;
  }

  // device pools for this chunk + staging copies of the AoS arrays
  float4* d_cub = nullptr; float2* d_lin = nullptr;
  CK(dev_alloc(h, &d_cub, cub_total));
  CK(dev_alloc(h, &d_lin, lin_total));
  uint64_t *d_start_c = nullptr, *d_start_l = nullptr, *d_knot_off = nullptr, *d_tco = nullptr, *d_tlo = nullptr;
  int16_t *d_paramNo = nullptr, *d_paramNo_l = nullptr, *d_slot_of = nullptr, *d_nseg = nullptr;
  int32_t *d_tile_sig = nullptr, *d_segbase_of = nullptr;
  float4* d_many = nullptr; float2* d_cl = nullptr;
  std::vector<void*> tmp;
  auto talloc = [&](void** p, size_t bytes) { cudaError_t e = cudaMalloc(p, bytes ? bytes : 16); if (e == cudaSuccess) tmp.push_back(*p); return e; };
  auto tfree = [&]() { for (void* p : tmp) cudaFree(p); tmp.clear(); };
#define TCK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { tfree(); char b__[512]; \
    snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); return fail(h, M3B_ERR_CUDA, b__); } } while (0)
#define TUP(dptr, hptr, count, type) do { TCK(talloc(reinterpret_cast<void**>(&dptr), (count) * sizeof(type))); \
    if ((count) > 0) TCK(cudaMemcpyAsync(dptr, hptr, (count) * sizeof(type), cudaMemcpyHostToDevice, h->stream)); } while (0)
  TUP(d_start_c, start_c.data(), static_cast<size_t>(n + 1), uint64_t);
  TUP(d_start_l, start_l.data(), static_cast<size_t>(n + 1), uint64_t);
  TUP(d_paramNo, paramNo, tot_c, int16_t);
  TUP(d_knot_off, knot_off, tot_c, uint64_t);
  TUP(d_many, coeff_many, total_knots, float4);
  TUP(d_paramNo_l, paramNo_l, tot_l, int16_t);
  TUP(d_cl, coeff_l, tot_l, float2);
  TUP(d_tile_sig, tile_sig.data(), static_cast<size_t>(ntile), int32_t);
  TUP(d_slot_of, h->sig_slot_of_param.data(), h->sig_slot_of_param.size(), int16_t);
  TUP(d_segbase_of, h->sig_segbase_of_param.data(), h->sig_segbase_of_param.size(), int32_t);
  TUP(d_nseg, h->nseg.data(), static_cast<size_t>(P), int16_t);
  TUP(d_tco, tile_cub_off.data(), static_cast<size_t>(ntile), uint64_t);
  TUP(d_tlo, tile_lin_off.data(), static_cast<size_t>(ntile), uint64_t);
  RetileArgs ra{};
  ra.n = n; ra.tile0_event = h->n_events_loaded; ra.T = T; ra.P = P;
  ra.start_c = d_start_c; ra.paramNo = d_paramNo; ra.knot_off = d_knot_off; ra.coeff_many = d_many;
  ra.start_l = d_start_l; ra.paramNo_l = d_paramNo_l; ra.coeff_l = d_cl;
  ra.tile_sig = d_tile_sig; ra.slot_of_param = d_slot_of; ra.segbase_of_param = d_segbase_of; ra.nseg = d_nseg;
  ra.tile_cub_off = d_tco; ra.tile_lin_off = d_tlo; ra.cub_pool = d_cub; ra.lin_pool = d_lin;
  TCK(launch_retile(ra, all_full ? 0 : static_cast<int64_t>(cub_total), all_full ? 0 : static_cast<int64_t>(lin_total), h->stream));
  TCK(cudaStreamSynchronize(h->stream));
  tfree();
#undef TUP
#undef TCK
  for (int64_t t = 0; t < ntile; ++t) {
    TileDesc td{};
    td.cub = d_cub + tile_cub_off[t];
    td.lin = d_lin + tile_lin_off[t];
    td.sig = tile_sig[t];
    td.ncnl = h->sigs[td.sig].nc | (h->sigs[td.sig].nl << 16);
    h->tiles.push_back(td);
  }
  h->n_events_loaded += n;
  h->tiles_dirty = true;
  return M3B_OK;
}

M3B_API int m3b_splines_append(m3b_handle* h, int64_t n, const uint32_t* nParamPerEvent, const int16_t* paramNo_arr,
                               const uint64_t* nKnots_arr, uint64_t total_knots, const float* coeff_many,
                               const uint32_t* nParamPerEvent_tf1, const int16_t* paramNo_tf1, const float* coeff_tf1) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  REQUIRE(h->splines_open, M3B_ERR_STATE, "m3b_splines_append: call m3b_splines_begin first");
  REQUIRE(n >= 0 && nParamPerEvent && nParamPerEvent_tf1, M3B_ERR_INVALID, "m3b_splines_append: bad argument");
  REQUIRE(h->n_events_loaded % h->T == 0, M3B_ERR_STATE,
          "m3b_splines_append: every chunk but the last must hold a multiple of tile_events events");
  REQUIRE(h->n_events_loaded + n <= h->n_events_total, M3B_ERR_INVALID, "m3b_splines_append: more events than announced");
  CK(cudaSetDevice(h->device));
  // internal sub-chunks bound the staging memory
  const int64_t sub = 32768 / h->T * h->T;
  uint64_t tot_c_all = 0;
  for (int64_t i = 0; i < n; ++i) tot_c_all += nParamPerEvent[2 * i];
  REQUIRE(tot_c_all == 0 || (paramNo_arr && nKnots_arr && coeff_many), M3B_ERR_INVALID, "m3b_splines_append: null TSpline3 arrays");
  uint64_t oc = 0, ol = 0;
  std::vector<uint64_t> rel;
  for (int64_t e0 = 0; e0 < n; e0 += sub) {
    const int64_t m = std::min<int64_t>(sub, n - e0);
    uint64_t nc = 0, nl = 0;
    for (int64_t i = 0; i < m; ++i) { nc += nParamPerEvent[2 * (e0 + i)]; nl += nParamPerEvent_tf1[2 * (e0 + i)]; }
    // knots of this sub-chunk's responses [oc, oc+nc): contiguous in the monolith
    const uint64_t k0 = nc ? nKnots_arr[oc] : 0;
    const uint64_t k1 = nc ? (oc + nc < tot_c_all ? nKnots_arr[oc + nc] : total_knots) : 0;
    REQUIRE(k1 >= k0 && k1 <= total_knots, M3B_ERR_INVALID, "m3b_splines_append: nKnots_arr is not increasing");
    rel.resize(nc);
    for (uint64_t s = 0; s < nc; ++s) rel[s] = nKnots_arr[oc + s] - k0;
    int rc = append_chunk(h, m, nParamPerEvent + 2 * e0, nc ? paramNo_arr + oc : nullptr, rel.data(), k1 - k0,
                          nc ? coeff_many + 4 * k0 : nullptr, nParamPerEvent_tf1 + 2 * e0,
                          nl ? paramNo_tf1 + ol : nullptr, nl ? coeff_tf1 + 2 * ol : nullptr);
    if (rc != M3B_OK) return rc;
    oc += nc; ol += nl;
  }
  return M3B_OK;
}

M3B_API int m3b_splines_end(m3b_handle* h) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  REQUIRE(h->splines_open, M3B_ERR_STATE, "m3b_splines_end: no upload in progress");
  REQUIRE(h->n_events_loaded == h->n_events_total, M3B_ERR_STATE, "m3b_splines_end: fewer events appended than announced");
  h->splines_open = false;
  h->splines_done = true;
  h->launch_ready = false;
  return M3B_OK;
}

M3B_API int m3b_upload_spline_monolith(m3b_handle* h, int32_t n_params, int32_t max_knots, const float* coeff_x,
                                       const int16_t* n_pts, int64_t n_events, const uint32_t* nParamPerEvent,
                                       const int16_t* paramNo_arr, const uint32_t* nKnots_arr, uint32_t total_knots,
                                       const float* coeff_many, const uint32_t* nParamPerEvent_tf1,
                                       const int16_t* paramNo_tf1, const float* coeff_tf1) {
  int rc = m3b_splines_begin(h, n_params, max_knots, coeff_x, n_pts, n_events);
  if (rc != M3B_OK) return rc;
  uint64_t tot_c = 0;
  for (int64_t i = 0; i < n_events; ++i) tot_c += nParamPerEvent[2 * i];
  std::vector<uint64_t> k64(tot_c);
  for (uint64_t s = 0; s < tot_c; ++s) k64[s] = nKnots_arr[s];
  rc = m3b_splines_append(h, n_events, nParamPerEvent, paramNo_arr, k64.data(), total_knots, coeff_many,
                          nParamPerEvent_tf1, paramNo_tf1, coeff_tf1);
  if (rc != M3B_OK) return rc;
  return m3b_splines_end(h);
}

// ------------------------------------------------------------------------------------------------
// binning, events, data
// ------------------------------------------------------------------------------------------------
static int upload_binning_body(m3b_handle* h, int32_t n_samples, const int32_t* n_dim, const int32_t* uniform,
                               const int32_t* nbins, const double* edges);
// a rejected upload leaves the handle as it was (no half-initialised binning)
static int upload_binning_impl(m3b_handle* h, int32_t n_samples, const int32_t* n_dim, const int32_t* uniform,
                               const int32_t* nbins, const double* edges) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  REQUIRE(n_samples > 0 && n_samples <= 64 && n_dim && nbins && edges, M3B_ERR_INVALID, "m3b_upload_binning: bad argument (1..64 samples)");
  REQUIRE(h->n_samples == 0, M3B_ERR_STATE, "m3b_upload_binning: binning already uploaded");
  const int rc = upload_binning_body(h, n_samples, n_dim, uniform, nbins, edges);
  if (rc != M3B_OK) {
    h->n_samples = 0; h->n_bins = 0;
    h->b_ndim.clear(); h->b_nbins.clear(); h->b_edge_off.clear(); h->b_stride.clear(); h->b_goff.clear(); h->sample_start.clear();
    h->b_edges.clear(); h->b_uniform.clear(); h->b_box_off.clear(); h->b_grid_off.clear(); h->b_grid_start.clear();
    h->b_grid_idx.clear(); h->b_boxes.clear();
  }
  return rc;
}
static int upload_binning_body(m3b_handle* h, int32_t n_samples, const int32_t* n_dim, const int32_t* uniform,
                               const int32_t* nbins, const double* edges) {
  CK(cudaSetDevice(h->device));
  h->n_samples = n_samples;
  h->b_ndim.assign(n_dim, n_dim + n_samples);
  h->b_nbins.assign(static_cast<size_t>(n_samples) * kMaxDim, 0);
  h->b_edge_off.assign(static_cast<size_t>(n_samples) * kMaxDim, 0);
  h->b_stride.assign(static_cast<size_t>(n_samples) * kMaxDim, 0);
  h->b_goff.assign(n_samples, 0);
  h->sample_start.assign(n_samples + 1, 0);
  h->b_uniform.assign(n_samples, 1); h->b_box_off.assign(n_samples, 0); h->b_grid_off.assign(n_samples, 0);
  h->b_grid_start.assign(1, 0);
  h->b_edges.clear();
  const double* ep = edges;
  int goff = 0;
  bool any_nonuniform = false;
  for (int s = 0; s < n_samples; ++s) {
    REQUIRE(n_dim[s] >= 1 && n_dim[s] <= kMaxDim, M3B_ERR_INVALID, "m3b_upload_binning: 1..4 dimensions per sample");
    const int nd = n_dim[s];
    int stride = 1;                                   // SampleStructs.h:656-664, x fastest
    h->b_goff[s] = goff;                              // BinningHandler.cpp:341-355
    h->sample_start[s] = goff;
    if (uniform && !uniform[s]) {
      // SampleBinningInfo::InitNonUniform + InitialiseGridMapping (Samples/SampleStructs.h:394-528)
      any_nonuniform = true;
      constexpr int kPerDim = 10;                     // BinsPerDimension
      const int nb = nbins[s * kMaxDim];
      REQUIRE(nb >= 1 && nd >= 2, M3B_ERR_INVALID, "m3b_upload_binning: a non-uniform sample needs >= 1 box and >= 2 dimensions");
      h->b_uniform[s] = 0;
      const size_t box0 = h->b_boxes.size();
      h->b_boxes.insert(h->b_boxes.end(), ep, ep + static_cast<size_t>(nb) * nd * 2);
      const double* ex = h->b_boxes.data() + box0;
      std::vector<std::vector<double>> me(nd, std::vector<double>(kPerDim + 1));
      int n_grid = 1;
      for (int d = 0; d < nd; ++d) {
        double mn = 1.7976931348623157e308, mx = -1.7976931348623157e308;
        for (int i = 0; i < nb; ++i) {
          REQUIRE(ex[(static_cast<size_t>(i) * nd + d) * 2] < ex[(static_cast<size_t>(i) * nd + d) * 2 + 1], M3B_ERR_INVALID, "m3b_upload_binning: box with lo >= hi");
          mn = std::min(mn, ex[(static_cast<size_t>(i) * nd + d) * 2]);
          mx = std::max(mx, ex[(static_cast<size_t>(i) * nd + d) * 2 + 1]);
        }
        const double width = (mx - mn) / static_cast<double>(kPerDim);
        for (int e = 0; e <= kPerDim; ++e) me[d][e] = mn + static_cast<double>(e) * width;
        h->b_nbins[s * kMaxDim + d] = kPerDim;
        h->b_edge_off[s * kMaxDim + d] = static_cast<int32_t>(h->b_edges.size());
        h->b_edges.insert(h->b_edges.end(), me[d].begin(), me[d].end());
        h->b_stride[s * kMaxDim + d] = stride;
        stride *= kPerDim; n_grid *= kPerDim;
      }
      h->b_grid_off[s] = static_cast<int32_t>(h->b_grid_start.size()) - 1;
      for (int g = 0; g < n_grid; ++g) {
        int rem = g;
        double cell[kMaxDim][2];
        for (int d = 0; d < nd; ++d) { const int i = rem % kPerDim; rem /= kPerDim; cell[d][0] = me[d][i]; cell[d][1] = me[d][i + 1]; }
        for (int i = 0; i < nb; ++i) {
          bool overlap = true;
          for (int d = 0; d < nd && overlap; ++d)
            overlap = ex[(static_cast<size_t>(i) * nd + d) * 2 + 1] > cell[d][0] && ex[(static_cast<size_t>(i) * nd + d) * 2] < cell[d][1];
          if (overlap) h->b_grid_idx.push_back(i);
        }
        h->b_grid_start.push_back(static_cast<int32_t>(h->b_grid_idx.size()));
      }
      ep += static_cast<size_t>(nb) * nd * 2;
      goff += nb;
      continue;
    }
    for (int d = 0; d < nd; ++d) {
      const int nb = nbins[s * kMaxDim + d];
      REQUIRE(nb >= 1, M3B_ERR_INVALID, "m3b_upload_binning: empty axis");
      for (int i = 0; i < nb; ++i)
        REQUIRE(ep[i] < ep[i + 1], M3B_ERR_INVALID, "m3b_upload_binning: edges must increase strictly");
      h->b_nbins[s * kMaxDim + d] = nb;
      h->b_edge_off[s * kMaxDim + d] = static_cast<int32_t>(h->b_edges.size());
      h->b_edges.insert(h->b_edges.end(), ep, ep + nb + 1);
      h->b_stride[s * kMaxDim + d] = stride;
      stride *= nb;
      ep += nb + 1;
    }
    goff += stride;
  }
  h->sample_start[n_samples] = goff;
  h->n_bins = goff;
  if (any_nonuniform) {
    // first box of every non-uniform sample, in boxes (hence one dimensionality for all of them)
    size_t doubles = 0;
    for (int s = 0; s < n_samples; ++s) {
      if (!h->b_uniform[s]) {
        REQUIRE(doubles % (2 * static_cast<size_t>(n_dim[s])) == 0, M3B_ERR_INVALID, "m3b_upload_binning: non-uniform samples must share one dimensionality");
        h->b_box_off[s] = static_cast<int32_t>(doubles / (2 * static_cast<size_t>(n_dim[s])));
        doubles += static_cast<size_t>(nbins[s * kMaxDim]) * n_dim[s] * 2;
      }
    }
    CK(dev_upload(h, &h->d_uniform, h->b_uniform));
    CK(dev_upload(h, &h->d_box_off, h->b_box_off));
    CK(dev_upload(h, &h->d_grid_off, h->b_grid_off));
    CK(dev_upload(h, &h->d_boxes, h->b_boxes));
    CK(dev_upload(h, &h->d_grid_start, h->b_grid_start));
    if (h->b_grid_idx.empty()) h->b_grid_idx.push_back(0);
    CK(dev_upload(h, &h->d_grid_idx, h->b_grid_idx));
  }
  CK(dev_upload(h, &h->d_ndim, h->b_ndim));
  CK(dev_upload(h, &h->d_nbins, h->b_nbins));
  CK(dev_upload(h, &h->d_edge_off, h->b_edge_off));
  CK(dev_upload(h, &h->d_stride, h->b_stride));
  CK(dev_upload(h, &h->d_goff, h->b_goff));
  CK(dev_upload(h, &h->d_sample_start, h->sample_start));
  CK(dev_upload(h, &h->d_edges, h->b_edges));
  for (int k = 0; k < 2; ++k) {
    CK(dev_alloc(h, &h->d_hw[k], static_cast<size_t>(2) * h->n_bins));
    CK(cudaMemsetAsync(h->d_hw[k], 0, sizeof(double) * 2 * h->n_bins, h->stream));
    h->mc_zero[k] = h->w2_zero[k] = true;
  }
  h->d_w2_frozen = h->d_hw[0] + h->n_bins;
  CK(dev_alloc(h, &h->d_data, static_cast<size_t>(h->n_bins)));
  CK(cudaMemsetAsync(h->d_data, 0, sizeof(double) * h->n_bins, h->stream));
  CK(dev_alloc(h, &h->d_llh, static_cast<size_t>(1 + n_samples)));
  CK(cudaMemsetAsync(h->d_llh, 0, sizeof(double) * (1 + n_samples), h->stream));
  CK(cudaHostAlloc(reinterpret_cast<void**>(&h->h_llh), sizeof(double) * (1 + n_samples), cudaHostAllocMapped));
  memset(h->h_llh, 0, sizeof(double) * (1 + n_samples));
  CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->h_llh_dev), h->h_llh, 0));
  CK(cudaHostAlloc(reinterpret_cast<void**>(&h->h_seq), 64, cudaHostAllocMapped));
  memset(h->h_seq, 0, 64);
  CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->h_seq_dev), h->h_seq, 0));
  h->launch_ready = false;
  return M3B_OK;
}

M3B_API int m3b_upload_binning(m3b_handle* h, int32_t n_samples, const int32_t* n_dim, const int32_t* nbins,
                               const double* edges) {
  return upload_binning_impl(h, n_samples, n_dim, nullptr, nbins, edges);
}

M3B_API int m3b_upload_binning_ex(m3b_handle* h, int32_t n_samples, const int32_t* n_dim, const int32_t* uniform,
                                  const int32_t* nbins, const double* edges) {
  return upload_binning_impl(h, n_samples, n_dim, uniform, nbins, edges);
}

M3B_API int m3b_upload_events(m3b_handle* h, int64_t n_events, const int32_t* sample_id, const double* kin,
                              int32_t n_norm_per_event, const int16_t* norm_idx, int32_t n_norm_values,
                              int32_t use_osc, const int32_t* osc_idx, int64_t n_osc_values, const float* static_w) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  REQUIRE(h->n_samples > 0, M3B_ERR_STATE, "m3b_upload_events: upload the binning first");
  REQUIRE(h->n_events == 0, M3B_ERR_STATE, "m3b_upload_events: events already uploaded");
  REQUIRE(n_events > 0 && sample_id && kin, M3B_ERR_INVALID, "m3b_upload_events: bad argument");
  REQUIRE(n_norm_per_event >= 0 && n_norm_per_event <= kMaxNormSlots, M3B_ERR_INVALID, "m3b_upload_events: at most 16 norm pointers per event");
  REQUIRE(!h->splines_done || h->n_events_total == n_events, M3B_ERR_INVALID, "m3b_upload_events: event count differs from the spline monolith's");
  CK(cudaSetDevice(h->device));
  const int T = h->T;
  h->n_events = n_events;
  h->n_tiles = (n_events + T - 1) / T;
  h->e_pad = h->n_tiles * T;
  const int64_t E = n_events, EP = h->e_pad;
  int max_dim = 0;                 // rows of kin: the largest dimensionality of ANY sample of the binning
  for (int s = 0; s < h->n_samples; ++s) max_dim = std::max(max_dim, h->b_ndim[s]);
  for (int64_t e = 0; e < E; ++e)
    REQUIRE(sample_id[e] >= 0 && sample_id[e] < h->n_samples, M3B_ERR_INVALID, "m3b_upload_events: sample_id out of range");
  // bins on the device
  CK(dev_alloc(h, &h->d_bin, static_cast<size_t>(EP)));
  {
    int32_t* d_sid = nullptr; double* d_kin = nullptr;
    const bool keep = (h->cfg.flags & M3B_FLAG_KEEP_KINEMATICS) != 0;
    CK(dev_alloc(h, &d_sid, static_cast<size_t>(E)));          // kept: a later m3b_upload_selection needs the samples
    if (keep) CK(dev_alloc(h, &d_kin, static_cast<size_t>(E) * max_dim));
    else CK(cudaMalloc(&d_kin, sizeof(double) * E * max_dim));
    CK(cudaMemcpyAsync(d_sid, sample_id, sizeof(int32_t) * E, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d_kin, kin, sizeof(double) * E * max_dim, cudaMemcpyHostToDevice, h->stream));
    BinArgs ba{};
    ba.n_events = E; ba.e_pad = EP; ba.sample_id = d_sid; ba.kin = d_kin; ba.n_samples = h->n_samples;
    ba.n_dim = h->d_ndim; ba.nbins = h->d_nbins; ba.edge_off = h->d_edge_off; ba.stride = h->d_stride;
    ba.global_off = h->d_goff; ba.edges = h->d_edges; ba.bin = h->d_bin;
    ba.uniform = h->d_uniform; ba.box_off = h->d_box_off; ba.grid_off = h->d_grid_off; ba.boxes = h->d_boxes;
    ba.grid_start = h->d_grid_start; ba.grid_idx = h->d_grid_idx;
    CK(launch_bins(ba, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->d_sample_id = d_sid; h->kin_dims = max_dim;
    if (keep) h->d_kin = d_kin; else cudaFree(d_kin);
    h->d_bin_raw = h->d_bin;
  }
  // norm bindings, transposed to [slot][event] and padded
  h->norm_slots = norm_idx ? n_norm_per_event : 0;
  h->n_norm_values = n_norm_values;
  if (h->norm_slots > 0) {
    std::vector<int16_t> tr(static_cast<size_t>(h->norm_slots) * EP, -1);
    for (int64_t e = 0; e < E; ++e)
      for (int j = 0; j < h->norm_slots; ++j) {
        const int16_t v = norm_idx[e * n_norm_per_event + j];
        REQUIRE(v < n_norm_values, M3B_ERR_INVALID, "m3b_upload_events: norm_idx out of range");
        tr[static_cast<size_t>(j) * EP + e] = v;
      }
    CK(dev_upload(h, &h->d_norm_idx, tr));
  }
  h->use_osc = use_osc != 0;
  if (h->use_osc) {
    h->n_osc = osc_idx ? n_osc_values : E;
    REQUIRE(h->n_osc > 0, M3B_ERR_INVALID, "m3b_upload_events: n_osc_values must be > 0 with osc_idx");
    if (osc_idx) {
      std::vector<int32_t> oi(static_cast<size_t>(EP), 0);
      for (int64_t e = 0; e < E; ++e) {
        REQUIRE(osc_idx[e] >= -1 && osc_idx[e] < n_osc_values, M3B_ERR_INVALID, "m3b_upload_events: osc_idx out of range");
        oi[e] = osc_idx[e];
      }
      CK(dev_upload(h, &h->d_osc_idx, oi));
    }
    CK(dev_alloc(h, &h->d_osc, static_cast<size_t>(h->n_osc)));
    std::vector<float> ones(static_cast<size_t>(h->n_osc), 1.f);
    CK(copy_sync(h, h->d_osc, ones.data(), sizeof(float) * h->n_osc, cudaMemcpyHostToDevice));
  }
  if (static_w) {
    std::vector<float> sw(static_cast<size_t>(EP), 1.f);
    std::copy(static_w, static_w + E, sw.begin());
    CK(dev_upload(h, &h->d_static, sw));
  }
  if (h->cfg.flags & M3B_FLAG_KEEP_EVENT_WEIGHTS) {
    CK(dev_alloc(h, &h->d_evt_spline_w, static_cast<size_t>(EP)));
    CK(dev_alloc(h, &h->d_evt_total_w, static_cast<size_t>(EP)));
  }
  h->launch_ready = false;
  return M3B_OK;
}

// the binned fill kernel reads the bins in its own walking order (m3b_upload_event_binned_splines): refresh that copy
static int refresh_sorted_bins(m3b_handle* h) {
  if (!h->d_bin_sorted) return M3B_OK;
  CK(launch_gather(h->d_bin_sorted, h->d_bin, h->d_perm, h->e_pad, 4, h->stream));
  ++h->launches;
  return M3B_OK;
}

// SampleHandlerFD::IsEventSelected (Samples/SampleHandlerFD.cpp:281-294) over the uploaded cuts: bin[] = bin_raw[] or -1
static int run_selection(m3b_handle* h) {
  if (h->n_cuts == 0) return refresh_sorted_bins(h);
  SelectArgs sa{};
  sa.n_events = h->n_events; sa.e_pad = h->e_pad; sa.sample_id = h->d_sample_id;
  sa.cut_start = h->d_cut_start; sa.cut_var = h->d_cut_var; sa.lower = h->d_cut_lo; sa.upper = h->d_cut_hi;
  sa.values = h->d_sel_vals; sa.kin = h->d_kin; sa.bin_raw = h->d_bin_raw; sa.bin = h->d_bin; sa.selected = h->d_selected;
  CK(launch_select(sa, h->stream));
  ++h->launches;
  return refresh_sorted_bins(h);
}

M3B_API int m3b_upload_selection(m3b_handle* h, int32_t n_cuts, const int32_t* cut_sample, const int32_t* cut_var,
                                 const double* lower, const double* upper, int32_t n_vars, const double* values) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  REQUIRE(h->n_events > 0 && h->d_sample_id, M3B_ERR_STATE, "m3b_upload_selection: upload the events first");
  REQUIRE(n_cuts >= 0 && n_vars >= 0, M3B_ERR_INVALID, "m3b_upload_selection: negative count");
  REQUIRE(n_cuts == 0 || (cut_sample && cut_var && lower && upper), M3B_ERR_INVALID, "m3b_upload_selection: null cut arrays");
  REQUIRE(n_vars == 0 || values, M3B_ERR_INVALID, "m3b_upload_selection: null cut-variable table");
  bool uses_kin = false;
  for (int k = 0; k < n_cuts; ++k) {
    REQUIRE(cut_sample[k] >= 0 && cut_sample[k] < h->n_samples, M3B_ERR_INVALID, "m3b_upload_selection: cut_sample out of range");
    REQUIRE(cut_var[k] < n_vars && cut_var[k] >= -h->kin_dims, M3B_ERR_INVALID, "m3b_upload_selection: cut_var out of range");
    uses_kin |= cut_var[k] < 0;
  }
  REQUIRE(!uses_kin || h->d_kin, M3B_ERR_STATE, "m3b_upload_selection: cuts on binning variables (cut_var < 0) need M3B_FLAG_KEEP_KINEMATICS");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (n_cuts == 0) {                        // selection removed: every event is back in
    if (h->d_bin_raw != h->d_bin) CK(copy_sync(h, h->d_bin, h->d_bin_raw, sizeof(int32_t) * h->e_pad, cudaMemcpyDeviceToDevice));
    if (h->d_selected) CK(cudaMemsetAsync(h->d_selected, 1, static_cast<size_t>(h->e_pad), h->stream));
    h->n_cuts = 0;
    return refresh_sorted_bins(h);
  }
  // cuts grouped by sample, StoredSelection order kept inside a sample (the first failing cut decides; any order gives
  // the same answer, but the order of evaluation is the reference's)
  std::vector<int32_t> start(static_cast<size_t>(h->n_samples) + 1, 0), var(n_cuts);
  std::vector<double> lo(n_cuts), hi(n_cuts);
  for (int k = 0; k < n_cuts; ++k) ++start[cut_sample[k] + 1];
  for (int s = 0; s < h->n_samples; ++s) start[s + 1] += start[s];
  std::vector<int32_t> fill(start.begin(), start.end() - 1);
  for (int k = 0; k < n_cuts; ++k) { const int j = fill[cut_sample[k]]++; var[j] = cut_var[k]; lo[j] = lower[k]; hi[j] = upper[k]; }
  if (h->d_bin_raw == h->d_bin) {           // first selection on this handle: keep FindGlobalBin's answer aside
    CK(dev_alloc(h, &h->d_bin_raw, static_cast<size_t>(h->e_pad)));
    CK(copy_sync(h, h->d_bin_raw, h->d_bin, sizeof(int32_t) * h->e_pad, cudaMemcpyDeviceToDevice));
    CK(dev_alloc(h, &h->d_selected, static_cast<size_t>(h->e_pad)));
    CK(cudaMemsetAsync(h->d_selected, 0, static_cast<size_t>(h->e_pad), h->stream));
  }
  CK(dev_upload(h, &h->d_cut_start, start));
  CK(dev_upload(h, &h->d_cut_var, var));
  CK(dev_upload(h, &h->d_cut_lo, lo));
  CK(dev_upload(h, &h->d_cut_hi, hi));
  if (n_vars > 0) {
    if (n_vars != h->n_sel_vars) CK(dev_alloc(h, &h->d_sel_vals, static_cast<size_t>(n_vars) * h->n_events));
    CK(copy_sync(h, h->d_sel_vals, values, sizeof(double) * n_vars * h->n_events, cudaMemcpyHostToDevice));
  }
  REQUIRE(!h->d_sh_start || n_vars == h->n_sel_vars, M3B_ERR_STATE, "m3b_upload_selection: upload the selection before m3b_upload_linear_shifts");
  if (h->d_sel_vals_nom && n_vars > 0) CK(copy_sync(h, h->d_sel_vals_nom, h->d_sel_vals, sizeof(double) * n_vars * h->n_events, cudaMemcpyDeviceToDevice));
  h->n_cuts = n_cuts; h->n_sel_vars = n_vars; h->sel_uses_kin = uses_kin;
  int rc = run_selection(h);
  if (rc != M3B_OK) return rc;
  CK(cudaStreamSynchronize(h->stream));
  return M3B_OK;
}

M3B_API int m3b_update_selection_values(m3b_handle* h, const double* values) {
  REQUIRE(h && values, M3B_ERR_INVALID, "m3b_update_selection_values: null argument");
  REQUIRE(h->n_cuts > 0 && h->n_sel_vars > 0, M3B_ERR_STATE, "m3b_update_selection_values: no selection with cut variables uploaded");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->d_sel_vals, values, sizeof(double) * h->n_sel_vars * h->n_events, cudaMemcpyHostToDevice, h->stream));
  return run_selection(h);
}

M3B_API int m3b_read_event_selected(m3b_handle* h, uint8_t* selected) {
  REQUIRE(h && selected, M3B_ERR_INVALID, "m3b_read_event_selected: null argument");
  REQUIRE(h->n_events > 0, M3B_ERR_STATE, "m3b_read_event_selected: no events");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (!h->d_selected) { memset(selected, 1, static_cast<size_t>(h->n_events)); return M3B_OK; }
  CK(cudaMemcpy(selected, h->d_selected, static_cast<size_t>(h->n_events), cudaMemcpyDeviceToHost));
  return M3B_OK;
}

static int rebin(m3b_handle* h) {
  BinArgs ba{};
  ba.n_events = h->n_events; ba.e_pad = h->e_pad; ba.sample_id = h->d_sample_id; ba.kin = h->d_kin; ba.n_samples = h->n_samples;
  ba.n_dim = h->d_ndim; ba.nbins = h->d_nbins; ba.edge_off = h->d_edge_off; ba.stride = h->d_stride;
  ba.global_off = h->d_goff; ba.edges = h->d_edges; ba.bin = h->d_bin_raw;
  ba.uniform = h->d_uniform; ba.box_off = h->d_box_off; ba.grid_off = h->d_grid_off; ba.boxes = h->d_boxes;
  ba.grid_start = h->d_grid_start; ba.grid_idx = h->d_grid_idx;
  CK(launch_bins(ba, h->stream));
  ++h->launches;
  return run_selection(h);       // ApplyShifts runs before IsEventSelected (Samples/SampleHandlerFD.cpp:359-361)
}

M3B_API int m3b_upload_linear_shifts(m3b_handle* h, int32_t n_shift_pars, int64_t n_events, const uint32_t* n_per_event,
                                     const int32_t* shift_par, const int32_t* target, const double* coef) {
  REQUIRE(h && n_per_event, M3B_ERR_INVALID, "m3b_upload_linear_shifts: null argument");
  REQUIRE(h->n_events > 0 && n_events == h->n_events, M3B_ERR_STATE, "m3b_upload_linear_shifts: upload the events first (same count)");
  REQUIRE(h->d_kin, M3B_ERR_STATE, "m3b_upload_linear_shifts: create the handle with M3B_FLAG_KEEP_KINEMATICS");
  REQUIRE(n_shift_pars > 0 && n_shift_pars <= 4096, M3B_ERR_INVALID, "m3b_upload_linear_shifts: 1..4096 shift parameters");
  REQUIRE(!h->d_sh_start, M3B_ERR_STATE, "m3b_upload_linear_shifts: already uploaded");
  std::vector<int64_t> start(static_cast<size_t>(n_events) + 1, 0);
  for (int64_t e = 0; e < n_events; ++e) start[e + 1] = start[e] + n_per_event[e];
  const int64_t total = start[n_events];
  REQUIRE(total == 0 || (shift_par && target && coef), M3B_ERR_INVALID, "m3b_upload_linear_shifts: null entry arrays");
  for (int64_t k = 0; k < total; ++k) {
    REQUIRE(shift_par[k] >= 0 && shift_par[k] < n_shift_pars, M3B_ERR_INVALID, "m3b_upload_linear_shifts: shift_par out of range");
    REQUIRE(target[k] >= 0 && target[k] < h->kin_dims + h->n_sel_vars, M3B_ERR_INVALID,
            "m3b_upload_linear_shifts: target out of range (binning variables, then the selection's cut variables)");
  }
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  CK(dev_upload(h, &h->d_sh_start, start));
  CK(dev_alloc(h, &h->d_sh_par, static_cast<size_t>(total)));
  CK(dev_alloc(h, &h->d_sh_target, static_cast<size_t>(total)));
  CK(dev_alloc(h, &h->d_sh_coef, static_cast<size_t>(total)));
  if (total > 0) {
    CK(copy_sync(h, h->d_sh_par, shift_par, sizeof(int32_t) * total, cudaMemcpyHostToDevice));
    CK(copy_sync(h, h->d_sh_target, target, sizeof(int32_t) * total, cudaMemcpyHostToDevice));
    CK(copy_sync(h, h->d_sh_coef, coef, sizeof(double) * total, cudaMemcpyHostToDevice));
  }
  // the nominal values every step starts from (ResetShifts)
  CK(dev_alloc(h, &h->d_kin_nom, static_cast<size_t>(h->n_events) * h->kin_dims));
  CK(copy_sync(h, h->d_kin_nom, h->d_kin, sizeof(double) * h->n_events * h->kin_dims, cudaMemcpyDeviceToDevice));
  if (h->n_sel_vars > 0) {
    CK(dev_alloc(h, &h->d_sel_vals_nom, static_cast<size_t>(h->n_events) * h->n_sel_vars));
    CK(copy_sync(h, h->d_sel_vals_nom, h->d_sel_vals, sizeof(double) * h->n_events * h->n_sel_vars, cudaMemcpyDeviceToDevice));
  }
  CK(dev_alloc(h, &h->d_shift_theta, static_cast<size_t>(n_shift_pars)));
  CK(cudaMemsetAsync(h->d_shift_theta, 0, sizeof(double) * n_shift_pars, h->stream));
  h->n_shift_pars = n_shift_pars;
  return M3B_OK;
}

M3B_API int m3b_set_shift_pars(m3b_handle* h, const double* values) {
  REQUIRE(h && values, M3B_ERR_INVALID, "m3b_set_shift_pars: null argument");
  REQUIRE(h->n_shift_pars > 0, M3B_ERR_STATE, "m3b_set_shift_pars: call m3b_upload_linear_shifts first");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->d_shift_theta, values, sizeof(double) * h->n_shift_pars, cudaMemcpyHostToDevice, h->stream));
  ShiftArgs sa{};
  sa.n_events = h->n_events; sa.n_dims = h->kin_dims; sa.n_sel_vars = h->d_sel_vals_nom ? h->n_sel_vars : 0;
  sa.kin_nom = h->d_kin_nom; sa.kin = h->d_kin; sa.sel_nom = h->d_sel_vals_nom; sa.sel = h->d_sel_vals;
  sa.start = h->d_sh_start; sa.par = h->d_sh_par; sa.target = h->d_sh_target; sa.coef = h->d_sh_coef; sa.theta = h->d_shift_theta;
  CK(launch_shift(sa, h->stream));
  ++h->launches;
  return rebin(h);
}

// Functional ("shift") parameters (Samples/SampleHandlerFD.cpp:545-564) call arbitrary std::functions per event, so they
// stay on the host: the caller applies its shifts to the kinematic variables and hands the shifted values over; the
// events are re-binned on the device with the same FindGlobalBin semantics.  Asynchronous on the handle's stream.
M3B_API int m3b_update_kinematics(m3b_handle* h, const double* kin) {
  REQUIRE(h && kin, M3B_ERR_INVALID, "m3b_update_kinematics: null argument");
  REQUIRE(h->d_kin && h->n_events > 0, M3B_ERR_STATE, "m3b_update_kinematics: create the handle with M3B_FLAG_KEEP_KINEMATICS and upload the events first");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->d_kin, kin, sizeof(double) * h->n_events * h->kin_dims, cudaMemcpyHostToDevice, h->stream));
  return rebin(h);
}

M3B_API int m3b_upload_data(m3b_handle* h, const double* data, int32_t n_bins) {
  REQUIRE(h && data, M3B_ERR_INVALID, "m3b_upload_data: null argument");
  REQUIRE(h->n_bins > 0 && n_bins == h->n_bins, M3B_ERR_INVALID, "m3b_upload_data: n_bins differs from the binning's");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->d_data, data, sizeof(double) * n_bins, cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return M3B_OK;
}

M3B_API int m3b_upload_osc(m3b_handle* h, const float* osc_w, int64_t n) {
  REQUIRE(h && osc_w, M3B_ERR_INVALID, "m3b_upload_osc: null argument");
  REQUIRE(h->use_osc && n == h->n_osc, M3B_ERR_INVALID, "m3b_upload_osc: length differs from the oscillation-weight array's");
  CK(cudaSetDevice(h->device));
  CK(cudaMemcpyAsync(h->d_osc, osc_w, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
  return M3B_OK;
}

M3B_API int m3b_register_host_buffer(m3b_handle* h, void* ptr, uint64_t bytes) {
  REQUIRE(h && ptr && bytes, M3B_ERR_INVALID, "m3b_register_host_buffer: bad argument");
  CK(cudaSetDevice(h->device));
  CK(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
  h->registered.push_back(ptr);
  return M3B_OK;
}

M3B_API int m3b_alloc_host(m3b_handle* h, uint64_t bytes, void** ptr) {
  REQUIRE(h && ptr && bytes, M3B_ERR_INVALID, "m3b_alloc_host: bad argument");
  CK(cudaSetDevice(h->device));
  CK(cudaHostAlloc(ptr, bytes, cudaHostAllocMapped | cudaHostAllocPortable));
  h->host_allocs.push_back(*ptr);
  return M3B_OK;
}

M3B_API int m3b_free_host(m3b_handle* h, void* ptr) {
  REQUIRE(h && ptr, M3B_ERR_INVALID, "m3b_free_host: bad argument");
  auto it = std::find(h->host_allocs.begin(), h->host_allocs.end(), ptr);
  REQUIRE(it != h->host_allocs.end(), M3B_ERR_INVALID, "m3b_free_host: not allocated by m3b_alloc_host on this handle");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaFreeHost(ptr));
  h->host_allocs.erase(it);
  return M3B_OK;
}

M3B_API int m3b_set_test_statistic(m3b_handle* h, int32_t ts) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  REQUIRE(ts >= 0 && ts <= 4, M3B_ERR_INVALID, "m3b_set_test_statistic: unknown test statistic");
  h->test_stat = ts;
  return M3B_OK;
}

M3B_API int m3b_set_flags(m3b_handle* h, int32_t set_mask, int32_t clear_mask) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  constexpr int32_t kRuntime = M3B_FLAG_NO_BATCH_KERNEL | M3B_FLAG_BATCH_KERNEL_V1 | M3B_FLAG_NO_SPIN_LLH;
  REQUIRE(((set_mask | clear_mask) & ~kRuntime) == 0, M3B_ERR_INVALID, "m3b_set_flags: only the run-time flags can change after creation");
  h->cfg.flags = (h->cfg.flags | set_mask) & ~clear_mask;
  return M3B_OK;
}

M3B_API int m3b_reset_w2(m3b_handle* h) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  h->first_time_w2 = true;
  return M3B_OK;
}

// ------------------------------------------------------------------------------------------------
// the step
// ------------------------------------------------------------------------------------------------
// SplineBase::FindSplineSegment (Splines/SplineBase.cpp:44-109), same decisions in the same order
extern "C++" {
template <class K>
static void find_segments_t(m3b_handle* h, const double* pars, const K* knots) {
  for (int i = 0; i < h->P; ++i) {
    const int nPoints = h->n_pts[i];
    const K* x = knots + static_cast<size_t>(i) * h->Kmax;      // FastSplineInfo::xPts (M3::float_t)
    const float xvar = static_cast<float>(pars[i]);             // always narrowed to float, :54
    h->param_values[i] = xvar;
    if (nPoints == 0) continue;
    int segment = 0, hi = nPoints - 1;
    const int prev = h->curr_segment[i];
    if (xvar <= x[0]) segment = 0;
    else if (xvar >= x[nPoints - 1]) segment = hi;
    else if (x[prev + 1] > xvar && xvar >= x[prev]) segment = prev;
    else {
      while (hi - segment > 1) {
        const int half = (segment + hi) / 2;
        if (xvar > x[half]) segment = half; else hi = half;
      }
    }
    if (segment >= nPoints - 1 && nPoints > 1) segment = nPoints - 2;
    h->curr_segment[i] = static_cast<int16_t>(segment);
    h->segments[i] = static_cast<int16_t>(segment);
  }
}
}  // extern "C++"
static void find_segments(m3b_handle* h, const double* pars) {
  if (h->f64) find_segments_t(h, pars, h->coeff_x_d.data());
  else if (!h->knots_d.empty()) find_segments_t(h, pars, h->knots_d.data());
  else find_segments_t(h, pars, h->coeff_x.data());
}

M3B_API int m3b_set_spline_knots_f64(m3b_handle* h, const double* x_pts) {
  REQUIRE(h, M3B_ERR_INVALID, "m3b_set_spline_knots_f64: null handle");
  REQUIRE(h->P > 0 && !h->binned, M3B_ERR_STATE, "m3b_set_spline_knots_f64: upload the spline monolith first");
  if (!x_pts) { h->knots_d.clear(); return M3B_OK; }
  h->knots_d.assign(x_pts, x_pts + static_cast<size_t>(h->P) * h->Kmax);
  return M3B_OK;
}

M3B_API int m3b_find_segments(m3b_handle* h, const double* spline_pars, int16_t* segments, float* param_values) {
  REQUIRE(h && spline_pars, M3B_ERR_INVALID, "m3b_find_segments: null argument");
  REQUIRE(h->P > 0, M3B_ERR_STATE, "m3b_find_segments: no spline monolith");
  find_segments(h, spline_pars);
  if (segments) std::copy(h->segments.begin(), h->segments.end(), segments);
  if (param_values) std::copy(h->param_values.begin(), h->param_values.end(), param_values);
  return M3B_OK;
}

static int prepare_launch(m3b_handle* h, bool w2_live) {
  CK(cudaSetDevice(h->device));
  if (h->binned) REQUIRE(h->d_wtiles, M3B_ERR_STATE, "step: call m3b_upload_event_binned_splines first");
  if (!h->binned && (h->tiles_dirty || !h->d_tiles)) {
    REQUIRE(!h->splines_open, M3B_ERR_STATE, "step: spline upload still open (call m3b_splines_end)");
    if (!h->splines_done) {
      // no response functions at all: tiles with an empty signature
      h->P = std::max(h->P, 0);
      std::vector<int16_t> none;
      const int sig = signature_of(h, none, none);
      h->tiles.assign(static_cast<size_t>(h->n_tiles), TileDesc{nullptr, nullptr, sig, 0});
    }
    REQUIRE(static_cast<int64_t>(h->tiles.size()) == h->n_tiles, M3B_ERR_STATE, "step: spline monolith and event table disagree on the number of events");
    CK(dev_upload(h, &h->d_tiles, h->tiles));
    CK(dev_upload(h, &h->d_sigs, h->sigs));
    CK(dev_upload(h, &h->d_sig_pool, h->sig_pool));
    h->tiles_dirty = false;
    h->launch_ready = false;
  }
  if (!h->h_step[0] || h->step.P != h->P || h->step.Nn != h->n_norm_values || h->step_sigs != (h->binned ? 0 : static_cast<int>(h->sigs.size()))) {
    h->step_sigs = h->binned ? 0 : static_cast<int>(h->sigs.size());
    h->step = make_step_layout(h->P, h->n_norm_values, h->step_sigs, h->max_nc, h->max_nl, h->f64);
    for (int i = 0; i < m3b_handle::kRing; ++i) {
      if (h->h_step[i]) cudaFreeHost(h->h_step[i]);
      CK(cudaHostAlloc(reinterpret_cast<void**>(&h->h_step[i]), h->step.bytes, cudaHostAllocDefault));
      CK(dev_alloc(h, &h->d_step[i], static_cast<size_t>(h->step.bytes)));
    }
    h->launch_ready = false;
  }
  if (!h->launch_ready || h->launch_w2_live != w2_live) {
    FillArgs a{};
    a.step = h->step; a.max_nc = h->max_nc; a.max_nl = h->max_nl; a.n_bins = h->n_bins; a.n_samples = h->n_samples;
    h->use_tma = !h->binned;
#ifdef M3B_EXPERIMENTS   // M3B_VARIANT=0..5: the register-streaming predecessor (m3b_kernels.cu), A/B only
    const char* v = getenv("M3B_VARIANT");
    if (v && v[0] >= '0' && v[0] <= '9' && !h->binned) { h->use_tma = false; h->variant = atoi(v); }
#endif
    if (h->binned) {
      // BinnedSplineHandler path: evaluate the non-flat splines, then gather/fill (m3b_binned.cu)
      int smem = binned_fill_smem_bytes(a, true, w2_live);
      const char* hm = experiment_env("M3B_BINNED_SMEM_HIST_MAX_KB");
      h->hist_in_smem = smem <= (hm ? atoi(hm) : 200) * 1024;
      if (!h->hist_in_smem) smem = binned_fill_smem_bytes(a, false, w2_live);
      const char* nt = experiment_env("M3B_BINNED_THREADS");
      h->binned_threads = (nt && (atoi(nt) == 256 || atoi(nt) == 512 || atoi(nt) == 768)) ? atoi(nt) : 1024;
      int bps = 0;
      CK(binned_fill_prepare(smem, h->f64, h->binned_threads, &bps));
      REQUIRE(bps > 0, M3B_ERR_CUDA, "step: binned fill kernel does not fit on an SM");
      h->smem = smem;
      { const char* mb = experiment_env("M3B_BINNED_MAX_BPS"); if (mb && atoi(mb) > 0) bps = std::min(bps, atoi(mb)); }
      const int tpb = h->binned_threads / 32;
      h->grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((h->n_wtiles + tpb - 1) / tpb, static_cast<int64_t>(bps) * h->sm_count)));
      h->binned_eval_grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((h->n_btiles + 1) / 2, 8ll * h->sm_count)));
      h->launch_ready = true;
      h->launch_w2_live = w2_live;
      return M3B_OK;
    }
    const int llh_scratch = h->n_samples * 32 * 8;
    if (h->use_tma) {
      // ring of 32 KB stages in whatever shared memory the fixed tables (and the privatised
      // histogram, if it fits next to >= 3 stages) leave of the 227 KB opt-in limit
      const char* be = experiment_env("M3B_TMA_BLOCKS_PER_SM");
      const int want_bps = be && atoi(be) > 1 ? atoi(be) : 1;
      const int budget = (232448 - 1024 * want_bps) / want_bps - 1024;     // static barriers + per-block reserve
      auto stages_for = [&](bool hist) {
        const TmaSmem L0 = tma_smem_layout(h->step, h->max_nc, h->max_nl, h->n_bins, hist, w2_live, h->zc_slots, 0);
        return std::min(32, (budget - L0.off_ring) / (L0.stage_bytes + (h->zc_slots ? 4096 : 0)));
      };
      h->hist_in_smem = true;
      int ns = stages_for(true);
      if (ns < 3) { h->hist_in_smem = false; ns = stages_for(false); }
      const char* se = experiment_env("M3B_TMA_STAGES");
      if (se && atoi(se) > 0) ns = std::min(ns, atoi(se));
      REQUIRE(ns >= 2, M3B_ERR_NOMEM, "step: per-step tables leave no room for the coefficient ring in shared memory");
      h->tma_stages = ns;
      h->tma = tma_smem_layout(h->step, h->max_nc, h->max_nl, h->n_bins, h->hist_in_smem, w2_live, h->zc_slots, ns);
      int smem = std::max(h->tma.total, llh_scratch);
      CK(fill_tma_set_smem(smem));
      int bps = 0;
      CK(fill_tma_occupancy(smem, &bps));
      REQUIRE(bps > 0, M3B_ERR_CUDA, "step: TMA fill kernel does not fit on an SM");
      h->smem = smem;
      const int64_t units = h->n_tiles * (h->T / 256);
      h->grid = static_cast<int>(std::min<int64_t>(units, static_cast<int64_t>(std::min(bps, want_bps)) * h->sm_count));
    }
#ifdef M3B_EXPERIMENTS
    if (!h->use_tma) {
      int smem = fill_smem_bytes(a, true, w2_live);
      h->hist_in_smem = smem <= 200 * 1024;
      if (!h->hist_in_smem) smem = fill_smem_bytes(a, false, w2_live);
      CK(fill_set_smem(h->T, h->variant, smem));
      int bps = 0;
      CK(fill_occupancy(h->T, h->variant, smem, &bps));
      REQUIRE(bps > 0, M3B_ERR_CUDA, "step: fill kernel does not fit on an SM");
      h->smem = smem;
      h->grid = static_cast<int>(std::min<int64_t>(h->n_tiles, static_cast<int64_t>(bps) * h->sm_count));
      const char* g = getenv("M3B_GRID_BLOCKS_PER_SM");
      if (g && atoi(g) > 0) h->grid = static_cast<int>(std::min<int64_t>(h->n_tiles, static_cast<int64_t>(std::min(atoi(g), bps)) * h->sm_count));
    }
#endif
    h->launch_ready = true;
    h->launch_w2_live = w2_live;
  }
  return M3B_OK;
}

enum StepMode { kFused = 0, kFillOnly = 1, kPeer = 2, kWeightsOnly = 3 };

static int enqueue_step(m3b_handle* h, const float* vals, const int16_t* segs, const double* norm_pars,
                        const float* osc_w, StepMode mode) {
  REQUIRE(h->n_events > 0 && h->n_bins > 0, M3B_ERR_STATE, "step: upload binning and events first");
  REQUIRE(h->n_norm_values == 0 || norm_pars, M3B_ERR_INVALID, "step: norm_pars is NULL but events carry norm pointers");
  const bool w2_live = h->first_time_w2;      // Samples/SampleHandlerFD.cpp:445,460
  // Oscillation weights handed over in pinned (registered) host memory are not copied: the TMA
  // kernel's producers stream them over PCIe while the coefficients stream from HBM.
  const float* osc_zc = nullptr;
  if (osc_w && h->use_osc && !h->d_osc_idx) {
    // asked every step (about a microsecond): the caller may have re-registered or re-allocated the array
    static const bool zc_off = [] { const char* z = experiment_env("M3B_OSC_ZEROCOPY"); return z && z[0] == '0'; }();
    cudaPointerAttributes at{};
    if (!zc_off && cudaPointerGetAttributes(&at, osc_w) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer &&
        (reinterpret_cast<uintptr_t>(at.devicePointer) & 15) == 0)
      osc_zc = static_cast<const float*>(at.devicePointer);
    cudaGetLastError();
    if (osc_zc && !h->zc_slots) { h->zc_slots = true; h->launch_ready = false; }
  }
  int rc = prepare_launch(h, w2_live);
  if (rc != M3B_OK) return rc;
  if (!h->use_tma) osc_zc = nullptr;

  // per-step table {segment, dx, value, norm}
  static const bool no_inline_env = experiment_env("M3B_NO_INLINE_STEP") != nullptr;
  const bool inline_step = h->step.bytes <= kStepInlineMax && !no_inline_env;
  FillArgs a{};
  const int slot = h->ring;
  unsigned char* st = a.step_inline;
  if (!inline_step) {
    h->ring = (h->ring + 1) % m3b_handle::kRing;
    CK(cudaEventSynchronize(h->step_ev[slot]));
    st = h->h_step[slot];
  }
  int32_t* seg = reinterpret_cast<int32_t*>(st + h->step.off_seg);
  float* dx = reinterpret_cast<float*>(st + h->step.off_dx);
  float* val = reinterpret_cast<float*>(st + h->step.off_val);
  float* norm = reinterpret_cast<float*>(st + h->step.off_norm);
  for (int p = 0; p < h->P; ++p) {
    seg[p] = segs[p];
    val[p] = vals[p];
    // dx = ParamValues[Param] - coeff_x[Param*_max_knots+segment]  (Splines/SplineMonolith.cpp:759), in float
    dx[p] = (h->n_pts[p] > 0 && !h->f64) ? vals[p] - h->coeff_x[static_cast<size_t>(p) * h->Kmax + segs[p]] : 0.f;
  }
  for (int j = 0; j < h->n_norm_values; ++j) norm[j] = static_cast<float>(norm_pars[j]);   // SampleHandlerFD.cpp:580
  if (h->f64) {      // default build: the binned eval reads the parameter un-narrowed (BinnedSplineHandler.cpp:327), norms stay double
    double* vd = reinterpret_cast<double*>(st + h->step.off_val_d);
    double* nd = reinterpret_cast<double*>(st + h->step.off_norm_d);
    for (int p = 0; p < h->P; ++p) vd[p] = h->spline_pars_last.empty() ? static_cast<double>(vals[p]) : h->spline_pars_last[p];
    for (int j = 0; j < h->n_norm_values; ++j) nd[j] = norm_pars[j];
  }
  if (h->step.n_sigs_x) {     // per-signature slot tables: no indirection left for the device
    int32_t* rowx = reinterpret_cast<int32_t*>(st + h->step.off_rowx);
    float* dxx = reinterpret_cast<float*>(st + h->step.off_dxx);
    float* lvx = reinterpret_cast<float*>(st + h->step.off_lvx);
    for (size_t g = 0; g < h->sigs.size(); ++g) {
      const SigDesc& sd = h->sigs[g];
      const int32_t* pool = h->sig_pool.data() + sd.off;
      for (int c = 0; c < sd.nc; ++c) {
        rowx[g * h->max_nc + c] = pool[sd.nc + c] + seg[pool[c]];
        dxx[g * h->max_nc + c] = dx[pool[c]];
      }
      for (int l = 0; l < sd.nl; ++l) lvx[g * h->max_nl + l] = val[pool[2 * sd.nc + l]];
    }
  }
  if (!inline_step) {
    CK(cudaMemcpyAsync(h->d_step[slot], st, h->step.bytes, cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(h->step_ev[slot], h->stream));
  }
  if (osc_w) {
    REQUIRE(h->use_osc, M3B_ERR_INVALID, "step: osc_w given but events were uploaded with use_osc=0");
    REQUIRE(!h->f64, M3B_ERR_INVALID, "step: this handle runs the double build; pass oscillation weights with m3b_upload_osc_f64");
    if (!osc_zc) {
      // an indexed oscillator table (NuOscillator's binned mode) is small: from mapped host memory it is fetched by a
      // kernel, which the fill kernel follows without the copy-engine hand-over
      const float* tab = nullptr;
      if (h->d_osc_idx && h->n_osc <= 65536) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, osc_w) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
          tab = static_cast<const float*>(at.devicePointer);
        cudaGetLastError();
      }
      if (tab) { CK(launch_table_copy(h->d_osc, tab, h->n_osc, h->stream)); ++h->launches; }
      else CK(cudaMemcpyAsync(h->d_osc, osc_w, sizeof(float) * h->n_osc, cudaMemcpyHostToDevice, h->stream));
    }
  }

  // fused mode alternates two buffers (the last block zeroes the other one for the next step);
  // the multi-GPU modes keep one buffer so its address is stable for the collective
  const int nxt = mode == kFused ? (h->cur ^ 1) : h->cur;
  double* mc = h->d_hw[nxt];
  double* w2 = h->d_hw[nxt] + h->n_bins;
  if (mode == kPeer) {
    // the partial histogram goes into this rank's exported buffer (two parities); the reduced one is
    // written to d_hw[cur] by the pull kernel
    REQUIRE(h->peer_world > 0, M3B_ERR_STATE, "m3b_step_peer: call m3b_peer_export/import first");
    ++h->peer_epoch;
    double* part = h->d_partial[h->peer_epoch & 1];
    CK(cudaMemsetAsync(part, 0, sizeof(double) * h->n_bins * (w2_live ? 2 : 1), h->stream));
    if (w2_live) h->d_w2_frozen = w2;
  } else if (mode != kWeightsOnly) {
    if (!h->mc_zero[nxt]) CK(cudaMemsetAsync(mc, 0, sizeof(double) * h->n_bins, h->stream));
    if (w2_live && !h->w2_zero[nxt]) CK(cudaMemsetAsync(w2, 0, sizeof(double) * h->n_bins, h->stream));
    h->mc_zero[nxt] = false;
    if (w2_live) { h->w2_zero[nxt] = false; h->d_w2_frozen = w2; }
  }

  a.step_inline_bytes = inline_step ? h->step.bytes : 0;
  a.tiles = h->d_tiles; a.sigs = h->d_sigs; a.sig_pool = h->d_sig_pool;
  a.n_tiles = static_cast<int32_t>(h->n_tiles); a.T = h->T; a.max_nc = h->max_nc; a.max_nl = h->max_nl;
  a.step_table = h->d_step[slot]; a.step = h->step;
  a.bin = h->d_bin; a.osc = h->use_osc ? h->d_osc : nullptr; a.osc_idx = h->d_osc_idx; a.static_w = h->d_static;
  a.norm_idx = h->d_norm_idx; a.norm_slots = h->norm_slots; a.e_pad = h->e_pad; a.n_events = h->n_events;
  a.hist = mc; a.w2 = w2_live ? w2 : nullptr;
  a.n_bins = h->n_bins; a.hist_in_smem = h->hist_in_smem ? 1 : 0;
  a.fuse_llh = (mode == kFused) ? 1 : 0;
  a.weights_only = (mode == kWeightsOnly) ? 1 : 0;
  a.test_stat = h->test_stat; a.n_samples = h->n_samples;
  a.data = h->d_data; a.w2_frozen = h->d_w2_frozen; a.sample_start = h->d_sample_start;
  for (int i = 0; i <= h->n_samples; ++i) a.sample_start_inline[i] = h->sample_start[i];
  a.tile_begin = 0;
  a.osc_host = osc_zc; a.osc_store = h->d_osc;
  if (h->use_tma) {
    a.tile_counter = h->d_tile_counter; a.n_stages = h->tma_stages; a.tma = h->tma;
    // queued fused steps overlap (programmatic dependent launch) unless something reads/writes per-event outputs or
    // brackets the kernel with timing events
    static const bool pdl_off = [] { const char* z = experiment_env("M3B_PDL"); return z && z[0] == '0'; }();
    a.pdl = (!pdl_off && mode == kFused && h->hist_in_smem && !h->timing && !h->d_evt_spline_w && !osc_zc && !h->d_trace) ? 1 : 0;
    static const int guard_env = [] { const char* ge = experiment_env("M3B_GUARD_X2"); return ge && atoi(ge) > 0 ? atoi(ge) : 6; }();
    a.guard_x2 = guard_env;
  }
  a.ticket = h->d_ticket; a.status = h->d_status; a.llh_dev = h->d_llh; a.llh_host = h->llh_host_override ? h->llh_host_override : h->h_llh_dev;
  // the -lnL of this step is announced through the sequence word when it goes to the handle's own mirror
  const bool announce = !h->llh_host_override && !(h->cfg.flags & M3B_FLAG_NO_SPIN_LLH) && mode != kWeightsOnly;
  const unsigned long long seq = announce ? ++h->seq_issued : 0ull;
  a.llh_seq_host = (announce && mode == kFused) ? h->h_seq_dev : nullptr; a.llh_seq = seq;
  if (mode != kWeightsOnly) h->seq_wait = (announce && (mode == kFused || (mode == kPeer && h->peer_pull))) ? seq : 0ull;
  a.evt_spline_w = h->d_evt_spline_w; a.evt_total_w = h->d_evt_total_w;
  a.trace = h->d_trace;
  if (mode == kFused) {
    // the last block zeroes the other buffer for the next step
    const int other = nxt ^ 1;
    a.hist_next = h->d_hw[other];
    h->mc_zero[other] = true;
    const bool next_live = h->cfg.update_w2 != 0;
    if (next_live && (h->d_hw[other] + h->n_bins) != h->d_w2_frozen) { a.w2_next = h->d_hw[other] + h->n_bins; h->w2_zero[other] = true; }
  }
  if (mode == kPeer) {
    double* part = h->d_partial[h->peer_epoch & 1];
    a.hist = part; a.w2 = w2_live ? part + h->n_bins : nullptr;
    a.peer_world = h->peer_world; a.peer_rank = h->peer_rank; a.peer_epoch = h->peer_epoch;
    a.peer_flag_own = h->d_flags[0];
  }
  if (h->timing) {
    if (h->tev_used + 2 > h->tev.size()) {
      for (int i = 0; i < 2; ++i) { cudaEvent_t e; CK(cudaEventCreate(&e)); h->tev.push_back(e); }
    }
    CK(cudaEventRecord(h->tev[h->tev_used], h->stream));
  }
  if (h->binned) {
    a.btiles = h->d_btiles; a.n_btiles = h->n_btiles; a.bcoef = h->d_bcoef; a.bx = h->d_bx; a.bw = h->d_bw;
    a.ell = h->d_ell; a.wtiles = h->d_wtiles; a.n_wtiles = h->n_wtiles; a.perm = h->d_perm;
    a.bin_sorted = h->d_bin_sorted; a.static_sorted = h->d_static_sorted; a.norm_idx_sorted = h->d_norm_idx_sorted; a.osc_idx_sorted = h->d_osc_idx_sorted;
    a.real_f64 = h->f64 ? 1 : 0;
    a.bcoef_d = h->d_bcoef_d; a.bx_d = h->d_bx_d; a.bw_d = h->d_bw_d; a.osc_d = h->d_osc_d; a.static_d = h->d_static_d;
    a.evt_spline_d = h->d_evt_spline_d; a.evt_total_d = h->d_evt_total_d;
    // the fill kernel's set-up overlaps the eval kernel's tail (programmatic dependent launch)
    a.binned_pdl = (h->n_btiles > 0 && !experiment_env("M3B_NO_BINNED_PDL")) ? 1 : 0;
    if (h->n_btiles > 0) { CK(launch_binned_eval(a, h->binned_eval_grid, h->stream)); ++h->launches; }
    CK(launch_binned_fill(a, h->grid, h->binned_threads, h->smem, h->stream));
  } else if (h->use_tma) CK(launch_fill_tma(a, h->grid, h->smem, h->stream));
#ifdef M3B_EXPERIMENTS
  else CK(launch_fill(a, h->variant, h->grid, h->smem, h->stream));
#endif
  if (h->timing) { CK(cudaEventRecord(h->tev[h->tev_used + 1], h->stream)); h->tev_used += 2; }
  ++h->launches;
  if (mode == kPeer && h->peer_pull) {
    const int par = h->peer_epoch & 1;
    LlhArgs l{};
    l.data = h->d_data; l.n_bins = h->n_bins; l.n_samples = h->n_samples;
    l.test_stat = h->test_stat; l.llh_dev = h->d_llh; l.llh_host = h->h_llh_dev;
    l.llh_seq_host = announce ? h->h_seq_dev : nullptr; l.llh_seq = seq;
    l.peer_world = h->peer_world; l.epoch = h->peer_epoch;
    for (int r = 0; r < h->peer_world; ++r) { l.peer_hist[r] = h->peer_partial[par][r]; l.peer_flag[r] = h->peer_flag[0][r]; }
    l.hist_out = mc; l.w2_out = w2; l.w2_live = w2_live ? 1 : 0; l.status = h->d_status;
    l.w2 = h->d_w2_frozen;
    l.partial = h->d_llh_partial; l.ticket = h->d_llh_ticket;
    for (int i = 0; i <= h->n_samples; ++i) l.sample_start_inline[i] = h->sample_start[i];
    const int blocks = std::max(1, std::min(kLlhPullMaxBlocks, (h->n_bins + 511) / 512));
    CK(launch_llh_pull(l, blocks, h->stream));
    ++h->launches;
    h->mc_zero[nxt] = false;
    if (w2_live) h->w2_zero[nxt] = false;
  }
  h->evt_weights_valid = h->d_evt_spline_w != nullptr;
  if (mode == kWeightsOnly) return M3B_OK;             // histograms, W2 state and step count untouched
  h->cur = nxt;
  h->last_w2_live = w2_live;
  if (!h->cfg.update_w2) h->first_time_w2 = false;     // Samples/SampleHandlerFD.cpp:342
  ++h->steps;
  return M3B_OK;
}

static int step_common(m3b_handle* h, const double* spline_pars, const double* norm_pars, const float* osc_w, StepMode mode) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  if (h->P > 0) {
    REQUIRE(spline_pars, M3B_ERR_INVALID, "step: spline_pars is NULL");
    find_segments(h, spline_pars);
    if (h->f64) h->spline_pars_last.assign(spline_pars, spline_pars + h->P);
  }
  return enqueue_step(h, h->param_values.data(), h->segments.data(), norm_pars, osc_w, mode);
}

M3B_API int m3b_step(m3b_handle* h, const double* spline_pars, const double* norm_pars, const float* osc_w) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  return step_common(h, spline_pars, norm_pars, osc_w, (h->cfg.flags & M3B_FLAG_NO_FUSED_LLH) ? kFillOnly : kFused);
}

M3B_API int m3b_step_segments(m3b_handle* h, const float* param_values, const int16_t* segments,
                              const double* norm_pars, const float* osc_w) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  REQUIRE(h->P == 0 || (param_values && segments), M3B_ERR_INVALID, "m3b_step_segments: null argument");
  for (int p = 0; p < h->P; ++p) {
    const int16_t sg = h->nseg[p] > 0 ? segments[p] : 0;
    REQUIRE(sg >= 0 && sg < std::max<int>(1, h->nseg[p]), M3B_ERR_INVALID, "m3b_step_segments: segment out of range");
    h->segments[p] = sg; h->curr_segment[p] = sg; h->param_values[p] = param_values[p];
  }
  h->spline_pars_last.clear();          // only the narrowed values are known on this entry point
  return enqueue_step(h, h->param_values.data(), h->segments.data(), norm_pars, osc_w,
                      (h->cfg.flags & M3B_FLAG_NO_FUSED_LLH) ? kFillOnly : kFused);
}

// SMonolithGPU::RunGPU_SplineMonolith (Splines/gpuSplineUtils.cu:444-512): evaluate every response,
// multiply per event, copy the per-event totals to the caller's (pinned) host array -- asynchronously;
// SynchroniseSplines() = m3b_synchronize.  Works without a sample handler: if no events were uploaded
// the library wires a trivial one-bin sample around the monolith's events.
static int ensure_standalone_events(m3b_handle* h) {
  if (h->n_events > 0) return M3B_OK;
  REQUIRE(h->splines_done, M3B_ERR_STATE, "m3b_eval_weights: upload the spline monolith first");
  if (h->n_samples == 0) {
    const int32_t nd = 1, nb[kMaxDim] = {1, 0, 0, 0};
    const double ed[2] = {0., 1.};
    int rc = m3b_upload_binning(h, 1, &nd, nb, ed);
    if (rc != M3B_OK) return rc;
  }
  CK(cudaSetDevice(h->device));
  const int T = h->T;
  h->n_events = h->n_events_total;
  h->n_tiles = (h->n_events + T - 1) / T;
  h->e_pad = h->n_tiles * T;
  CK(dev_alloc(h, &h->d_bin, static_cast<size_t>(h->e_pad)));
  CK(cudaMemsetAsync(h->d_bin, 0xFF, sizeof(int32_t) * h->e_pad, h->stream));       // bin -1: nothing is ever filled
  h->d_bin_raw = h->d_bin;
  if (!h->d_evt_spline_w) {
    CK(dev_alloc(h, &h->d_evt_spline_w, static_cast<size_t>(h->e_pad)));
    CK(dev_alloc(h, &h->d_evt_total_w, static_cast<size_t>(h->e_pad)));
  }
  h->launch_ready = false;
  return M3B_OK;
}

M3B_API int m3b_eval_weights(m3b_handle* h, const float* param_values, const int16_t* segments,
                             float* host_total_weights) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  REQUIRE(h->P > 0 && param_values && segments, M3B_ERR_INVALID, "m3b_eval_weights: null argument / no monolith");
  REQUIRE(!h->binned, M3B_ERR_STATE, "m3b_eval_weights: event-by-event monolith only (use m3b_step + m3b_read_binned_weights)");
  int rc = ensure_standalone_events(h);
  if (rc != M3B_OK) return rc;
  REQUIRE(h->d_evt_spline_w, M3B_ERR_STATE, "m3b_eval_weights: create the handle with M3B_FLAG_KEEP_EVENT_WEIGHTS");
  for (int p = 0; p < h->P; ++p) {
    // a parameter without any TSpline3 response has no knots here; whatever the caller's segment is, it is unused
    const int16_t sg = h->nseg[p] > 0 ? segments[p] : 0;
    REQUIRE(sg >= 0 && sg < std::max<int>(1, h->nseg[p]), M3B_ERR_INVALID, "m3b_eval_weights: segment out of range");
    h->segments[p] = sg; h->curr_segment[p] = sg; h->param_values[p] = param_values[p];
  }
  std::vector<double> ones(static_cast<size_t>(std::max(h->n_norm_values, 1)), 1.0);
  rc = enqueue_step(h, h->param_values.data(), h->segments.data(), ones.data(), nullptr, kWeightsOnly);
  if (rc != M3B_OK) return rc;
  if (host_total_weights)
    CK(cudaMemcpyAsync(host_total_weights, h->d_evt_spline_w, sizeof(float) * h->n_events, cudaMemcpyDeviceToHost, h->stream));
  return M3B_OK;
}

// Batched proposals (BASELINE config 5: parallel chains, DelayedMR2T2 stages, LLH scans): n_sets parameter
// vectors against the same events.  Sets are evaluated in order with the reference's sequential semantics
// (cached segments of SplineBase::FindSplineSegment, W2 freeze), all launches are enqueued back to back and
// the host synchronises ONCE; every set's -lnL lands in its own slot of a mapped host array.
static int step_batch_impl(m3b_handle* h, int32_t n_sets, const double* spline_pars, const double* norm_pars,
                           const float* osc_w, double* llh_total, double* llh_per_sample, double* mc_out) {
  REQUIRE(h && llh_total, M3B_ERR_INVALID, "m3b_step_batch: null argument");
  REQUIRE(n_sets > 0, M3B_ERR_INVALID, "m3b_step_batch: n_sets must be positive");
  REQUIRE(h->peer_world == 0 && !(h->cfg.flags & M3B_FLAG_NO_FUSED_LLH), M3B_ERR_STATE, "m3b_step_batch: single-GPU fused handles only");
  REQUIRE(h->n_bins > 0, M3B_ERR_STATE, "m3b_step_batch: upload binning and events first");
  REQUIRE(h->P == 0 || spline_pars, M3B_ERR_INVALID, "m3b_step_batch: spline_pars is NULL");
  REQUIRE(h->n_norm_values == 0 || norm_pars, M3B_ERR_INVALID, "m3b_step_batch: norm_pars is NULL but events carry norm pointers");
  CK(cudaSetDevice(h->device));
  const size_t slot = static_cast<size_t>(1 + h->n_samples);
  if (h->batch_cap < static_cast<size_t>(n_sets)) {
    CK(cudaStreamSynchronize(h->stream));
    if (h->h_batch) cudaFreeHost(h->h_batch);
    h->h_batch = nullptr; h->batch_cap = 0;      // (the kernel's own staging buffers are sized by m3b_batch_try)
    CK(cudaHostAlloc(reinterpret_cast<void**>(&h->h_batch), sizeof(double) * slot * n_sets, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->h_batch_dev), h->h_batch, 0));
    h->batch_cap = static_cast<size_t>(n_sets);
  }
  int rc = M3B_OK;
  int32_t i0 = 0;
  // the batched kernel (one pass over the coefficient rows for up to 256 sets) whenever the state allows it
  while (i0 < n_sets) {
    const int32_t n = std::min<int32_t>(256, n_sets - i0);
    int done = 0;
    rc = m3b_batch_try(h, n, h->P > 0 ? spline_pars + static_cast<size_t>(i0) * h->P : nullptr,
                       h->n_norm_values > 0 ? norm_pars + static_cast<size_t>(i0) * h->n_norm_values : nullptr,
                       i0 == 0 ? osc_w : nullptr, h->h_batch_dev + slot * i0, &done);
    if (rc != M3B_OK) return rc;
    if (!done) break;
    if (mc_out) {
      // the chunk's histograms live as [bin][256 sets]; the next chunk reuses the buffer
      std::vector<double> tmp(static_cast<size_t>(h->n_bins) * 256);
      CK(cudaMemcpyAsync(tmp.data(), h->bt_hist, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      for (int32_t k = 0; k < n; ++k)
        for (int b = 0; b < h->n_bins; ++b) mc_out[static_cast<size_t>(i0 + k) * h->n_bins + b] = tmp[static_cast<size_t>(b) * 256 + k];
    }
    i0 += n;
  }
  for (int32_t i = i0; i < n_sets && rc == M3B_OK; ++i) {
    h->llh_host_override = h->h_batch_dev + slot * i;
    rc = step_common(h, h->P > 0 ? spline_pars + static_cast<size_t>(i) * h->P : nullptr,
                     h->n_norm_values > 0 ? norm_pars + static_cast<size_t>(i) * h->n_norm_values : nullptr,
                     (i == 0 && i0 == 0) ? osc_w : nullptr, kFused);
    if (rc == M3B_OK && mc_out)     // stream-ordered after this set's step, before the next one reuses the other buffer
      CK(cudaMemcpyAsync(mc_out + static_cast<size_t>(i) * h->n_bins, h->d_hw[h->cur], sizeof(double) * h->n_bins,
                         cudaMemcpyDeviceToHost, h->stream));
  }
  h->llh_host_override = nullptr;
  if (rc != M3B_OK) return rc;
  CK(cudaStreamSynchronize(h->stream));
  bool any_nan = false;
  for (int32_t i = 0; i < n_sets; ++i) any_nan |= h->h_batch[slot * i] != h->h_batch[slot * i];
  if (any_nan) {
    int32_t st = 0;
    CK(cudaMemcpy(&st, h->d_status, sizeof st, cudaMemcpyDeviceToHost));
    if (st != 0) CK(cudaMemsetAsync(h->d_status, 0, sizeof st, h->stream));
    if (st & 2) return fail(h, M3B_ERR_MATH, "m3b_step_batch: negative square root in the Barlow-Beeston coefficient (the reference throws "
                                             "MaCh3Exception here, Samples/SampleHandlerBase.cpp:64-67)");
  }
  for (int32_t i = 0; i < n_sets; ++i) {
    llh_total[i] = h->h_batch[slot * i];
    if (llh_per_sample) for (int s = 0; s < h->n_samples; ++s) llh_per_sample[static_cast<size_t>(i) * h->n_samples + s] = h->h_batch[slot * i + 1 + s];
  }
  // m3b_llh after a batch returns the last set's value
  for (size_t k = 0; k < slot; ++k) h->h_llh[k] = h->h_batch[slot * (n_sets - 1) + k];
  h->seq_wait = 0;
  return M3B_OK;
}

M3B_API int m3b_step_batch(m3b_handle* h, int32_t n_sets, const double* spline_pars, const double* norm_pars,
                           const float* osc_w, double* llh_total, double* llh_per_sample) {
  return step_batch_impl(h, n_sets, spline_pars, norm_pars, osc_w, llh_total, llh_per_sample, nullptr);
}

M3B_API int m3b_step_batch_hist(m3b_handle* h, int32_t n_sets, const double* spline_pars, const double* norm_pars,
                                const float* osc_w, double* llh_total, double* llh_per_sample, double* mc) {
  REQUIRE(mc, M3B_ERR_INVALID, "m3b_step_batch_hist: null histogram output");
  return step_batch_impl(h, n_sets, spline_pars, norm_pars, osc_w, llh_total, llh_per_sample, mc);
}

M3B_API int m3b_step_fill(m3b_handle* h, const double* spline_pars, const double* norm_pars, const float* osc_w) {
  return step_common(h, spline_pars, norm_pars, osc_w, kFillOnly);
}

M3B_API int m3b_step_peer(m3b_handle* h, const double* spline_pars, const double* norm_pars, const float* osc_w) {
  return step_common(h, spline_pars, norm_pars, osc_w, kPeer);
}

M3B_API int m3b_hist_device_ptr(m3b_handle* h, void** dev_ptr, int32_t* n_bins, int32_t* w2_live) {
  REQUIRE(h && dev_ptr, M3B_ERR_INVALID, "m3b_hist_device_ptr: null argument");
  REQUIRE(h->n_bins > 0, M3B_ERR_STATE, "m3b_hist_device_ptr: no binning");
  *dev_ptr = h->d_hw[h->cur];
  if (n_bins) *n_bins = h->n_bins;
  if (w2_live) *w2_live = h->last_w2_live ? 1 : 0;
  return M3B_OK;
}

M3B_API int m3b_llh_from_hist(m3b_handle* h) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  REQUIRE(h->steps > 0, M3B_ERR_STATE, "m3b_llh_from_hist: no step yet");
  CK(cudaSetDevice(h->device));
  LlhArgs l{};
  l.hist = h->d_hw[h->cur]; l.w2 = h->d_w2_frozen; l.data = h->d_data; l.sample_start = h->d_sample_start;
  l.n_bins = h->n_bins; l.n_samples = h->n_samples; l.test_stat = h->test_stat;
  l.llh_dev = h->d_llh; l.llh_host = h->h_llh_dev; l.status = h->d_status;
  if (!(h->cfg.flags & M3B_FLAG_NO_SPIN_LLH)) { l.llh_seq_host = h->h_seq_dev; l.llh_seq = ++h->seq_issued; h->seq_wait = l.llh_seq; }
  CK(launch_llh(l, h->stream));
  ++h->launches;
  return M3B_OK;
}

M3B_API int m3b_synchronize(m3b_handle* h) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  return M3B_OK;
}

M3B_API int m3b_llh(m3b_handle* h, double* total, double* per_sample) {
  REQUIRE(h && total, M3B_ERR_INVALID, "m3b_llh: null argument");
  REQUIRE(h->steps > 0, M3B_ERR_STATE, "m3b_llh: no step yet");
  bool seen = false;
  if (h->seq_wait != 0) {
    // poll the sequence word the likelihood kernel writes behind -lnL (mapped host memory).  Bounded: after ~2 s fall back
    // to the stream wait, which also surfaces a device fault.
    volatile unsigned long long* sq = h->h_seq;
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spin = 1;; ++spin) {
      if (*sq >= h->seq_wait) { seen = true; break; }
      __builtin_ia32_pause();
      if ((spin & 0xfffu) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(2)) break;
    }
  }
  CK(cudaSetDevice(h->device));
  if (!seen) CK(cudaStreamSynchronize(h->stream));
  if (h->h_llh[0] != h->h_llh[0]) {     // NaN: an exchange time-out, or a case in which the reference throws?
    int32_t st = 0;
    CK(cudaMemcpy(&st, h->d_status, sizeof st, cudaMemcpyDeviceToHost));
    if (st != 0) CK(cudaMemsetAsync(h->d_status, 0, sizeof st, h->stream));
    if (st & 1) return fail(h, M3B_ERR_PEER, "m3b_llh: peer histogram exchange timed out");
    if (st & 2) return fail(h, M3B_ERR_MATH, "m3b_llh: negative square root in the Barlow-Beeston coefficient (the reference throws "
                                             "MaCh3Exception here, Samples/SampleHandlerBase.cpp:64-67)");
  }
  *total = h->h_llh[0];
  if (per_sample) for (int s = 0; s < h->n_samples; ++s) per_sample[s] = h->h_llh[1 + s];
  return M3B_OK;
}

// ------------------------------------------------------------------------------------------------
// read-back
// ------------------------------------------------------------------------------------------------
M3B_API int m3b_read_hist(m3b_handle* h, double* mc, double* w2) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  REQUIRE(h->n_bins > 0, M3B_ERR_STATE, "m3b_read_hist: no binning");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (mc) CK(cudaMemcpy(mc, h->d_hw[h->cur], sizeof(double) * h->n_bins, cudaMemcpyDeviceToHost));
  if (w2) CK(cudaMemcpy(w2, h->d_w2_frozen, sizeof(double) * h->n_bins, cudaMemcpyDeviceToHost));
  return M3B_OK;
}

M3B_API int m3b_read_event_bins(m3b_handle* h, int32_t* bins) {
  REQUIRE(h && bins, M3B_ERR_INVALID, "m3b_read_event_bins: null argument");
  REQUIRE(h->n_events > 0, M3B_ERR_STATE, "m3b_read_event_bins: no events");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(bins, h->d_bin_raw, sizeof(int32_t) * h->n_events, cudaMemcpyDeviceToHost));
  return M3B_OK;
}

M3B_API int m3b_read_event_weights(m3b_handle* h, float* spline_w, float* total_w) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  REQUIRE(h->steps > 0, M3B_ERR_STATE, "m3b_read_event_weights: no step yet");
  REQUIRE(h->d_evt_spline_w, M3B_ERR_STATE, "m3b_read_event_weights: create the handle with M3B_FLAG_KEEP_EVENT_WEIGHTS");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (spline_w) CK(cudaMemcpy(spline_w, h->d_evt_spline_w, sizeof(float) * h->n_events, cudaMemcpyDeviceToHost));
  if (total_w) CK(cudaMemcpy(total_w, h->d_evt_total_w, sizeof(float) * h->n_events, cudaMemcpyDeviceToHost));
  return M3B_OK;
}

// ------------------------------------------------------------------------------------------------
// peer exchange wiring (CUDA IPC between the per-GPU processes)
// ------------------------------------------------------------------------------------------------
}  // extern "C"
int m3b_peer_alloc(m3b_handle* h) {
  REQUIRE(h && h->n_bins > 0, M3B_ERR_STATE, "peer exchange: upload the binning first");
  CK(cudaSetDevice(h->device));
  if (!h->d_partial[0]) {
    // one exported allocation: [2 parities][mc[n_bins] | w2[n_bins]] then this rank's epoch flag
    const size_t part_d = static_cast<size_t>(2) * h->n_bins;
    const size_t total_d = 2 * part_d + 16;
    double* base = nullptr;
    CK(dev_alloc(h, &base, total_d));
    CK(cudaMemsetAsync(base, 0, total_d * sizeof(double), h->stream));
    h->d_partial[0] = base; h->d_partial[1] = base + part_d;
    h->d_flags[0] = reinterpret_cast<unsigned int*>(base + 2 * part_d);
    h->d_flags[1] = h->d_flags[0];
    CK(dev_alloc(h, &h->d_llh_partial, static_cast<size_t>(kLlhPullMaxBlocks) * std::max(1, h->n_samples)));
    CK(dev_alloc(h, &h->d_llh_ticket, 1));
    CK(cudaMemsetAsync(h->d_llh_ticket, 0, sizeof(unsigned int), h->stream));
  }
  return M3B_OK;
}
extern "C" {

M3B_API int m3b_peer_export(m3b_handle* h, int32_t rank, int32_t world, void* ipc_handle_64B) {
  REQUIRE(h && ipc_handle_64B, M3B_ERR_INVALID, "m3b_peer_export: null argument");
  REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, M3B_ERR_INVALID, "m3b_peer_export: 1..8 ranks");
  REQUIRE(h->n_bins > 0, M3B_ERR_STATE, "m3b_peer_export: upload the binning first");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  int rc = m3b_peer_alloc(h);
  if (rc != M3B_OK) return rc;
  h->peer_world = world; h->peer_rank = rank;
  for (int par = 0; par < 2; ++par) { h->peer_partial[par][rank] = h->d_partial[par]; h->peer_flag[par][rank] = h->d_flags[par]; }
  cudaIpcMemHandle_t mh;
  CK(cudaIpcGetMemHandle(&mh, h->d_partial[0]));
  memcpy(ipc_handle_64B, &mh, 64);
  return M3B_OK;
}

M3B_API int m3b_peer_import(m3b_handle* h, int32_t peer_rank, const void* ipc_handle_64B) {
  REQUIRE(h && ipc_handle_64B, M3B_ERR_INVALID, "m3b_peer_import: null argument");
  REQUIRE(h->peer_world > 0 && peer_rank >= 0 && peer_rank < h->peer_world, M3B_ERR_STATE, "m3b_peer_import: export first / bad rank");
  if (peer_rank == h->peer_rank) return M3B_OK;
  CK(cudaSetDevice(h->device));
  cudaIpcMemHandle_t mh;
  memcpy(&mh, ipc_handle_64B, 64);
  void* p = nullptr;
  CK(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
  h->ipc_opened.push_back(p);
  double* base = static_cast<double*>(p);
  const size_t part_d = static_cast<size_t>(2) * h->n_bins;
  h->peer_partial[0][peer_rank] = base; h->peer_partial[1][peer_rank] = base + part_d;
  h->peer_flag[0][peer_rank] = reinterpret_cast<unsigned int*>(base + 2 * part_d);
  h->peer_flag[1][peer_rank] = h->peer_flag[0][peer_rank];
  return M3B_OK;
}

M3B_API int m3b_set_timing(m3b_handle* h, int32_t enabled) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  h->timing = enabled != 0;
  return M3B_OK;
}

M3B_API int m3b_kernel_time(m3b_handle* h, double* total_ms, int64_t* n_launches) {
  REQUIRE(h && total_ms && n_launches, M3B_ERR_INVALID, "m3b_kernel_time: null argument");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  double tot = 0;
  for (size_t i = 0; i + 1 < h->tev_used; i += 2) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->tev[i], h->tev[i + 1]));
    tot += ms;
  }
  *total_ms = tot;
  *n_launches = static_cast<int64_t>(h->tev_used / 2);
  h->tev_used = 0;
  return M3B_OK;
}

// per-block timeline of the NEXT/LAST fill launch (debug aid for the tail/ramp analysis in DESIGN.md):
// enable with out == NULL (allocates), then after a step call again with out = u64[grid*8].
M3B_API int m3b_block_trace(m3b_handle* h, uint64_t* out, int32_t* grid) {
  REQUIRE(h, M3B_ERR_INVALID, "null handle");
  CK(cudaSetDevice(h->device));
  if (!h->d_trace) {
    CK(dev_alloc(h, &h->d_trace, static_cast<size_t>(8) * 4096));
    CK(cudaMemsetAsync(h->d_trace, 0, sizeof(unsigned long long) * 8 * 4096, h->stream));
  }
  if (grid) *grid = h->grid;
  if (out) {
    CK(cudaStreamSynchronize(h->stream));
    REQUIRE(h->grid <= 4096, M3B_ERR_INVALID, "m3b_block_trace: grid too large");
    CK(cudaMemcpy(out, h->d_trace, sizeof(unsigned long long) * 8 * h->grid, cudaMemcpyDeviceToHost));
    if (experiment_env("M3B_TRACE_LAST")) {
      unsigned long long last[2];
      CK(cudaMemcpy(last, h->d_trace + 8 * 4000, sizeof last, cudaMemcpyDeviceToHost));
      unsigned long long t0 = ~0ull;
      for (int b = 0; b < h->grid; ++b) t0 = std::min<unsigned long long>(t0, out[8 * b]);
      fprintf(stderr, "m3b trace: last block won the ticket at %.2f us, wrote -lnL at %.2f us\n", (last[0] - t0) / 1e3, (last[1] - t0) / 1e3);
    }
  }
  return M3B_OK;
}

M3B_API int m3b_get_info(m3b_handle* h, m3b_info* out) {
  REQUIRE(h && out, M3B_ERR_INVALID, "m3b_get_info: null argument");
  memset(out, 0, sizeof *out);
  out->n_events = h->n_events; out->n_tiles = h->n_tiles; out->n_params = h->P; out->n_bins = h->n_bins;
  out->n_samples = h->n_samples; out->n_signatures = static_cast<int32_t>(h->sigs.size());
  out->tile_events = h->T; out->grid_blocks = h->grid; out->smem_bytes = h->smem; out->hist_in_smem = h->hist_in_smem ? 1 : 0;
  out->device_bytes = h->device_bytes;
  uint64_t per_evt = 4;   // bin id
  if (h->use_osc) per_evt += 4 + (h->d_osc_idx ? 4 : 0);
  if (h->d_static) per_evt += 4;
  per_evt += 2ull * h->norm_slots;
  out->active_bytes_per_step = h->active_coef_bytes + per_evt * static_cast<uint64_t>(h->e_pad);
  out->steps = h->steps; out->kernel_launches = h->launches;
  out->kernel_variant = h->binned ? -2 : (h->use_tma ? -1 : h->variant); out->tma_stages = h->use_tma ? h->tma_stages : 0;
  return M3B_OK;
}

}  // extern "C"
