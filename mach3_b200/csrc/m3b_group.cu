// m3b_group.cu -- ONE sample handler over several B200s of a box, driven by ONE process and ONE calling thread.
//
// The reference's fitters are single-process, single-threaded callers: per MCMC step they call, for every sample
// handler, Reweight() and then GetLikelihood() (Fitters/MR2T2.cpp:62-74; FitterBase::DragRace, Fitters/FitterBase.cpp:
// 461-520).  m3b_group_step / m3b_group_llh keep exactly that surface.  Inside:
//
//   members      one m3b_handle per device, each holding a contiguous, tile-aligned shard of the events (all members hold
//                the full binning; only the lead needs the data histogram);
//   workers      one host thread per non-lead member, bound to its device, parked on an atomic step counter (spin for a
//                short while after a step, then sleep on a condition variable).  m3b_group_step copies the step's
//                parameters, bumps the counter, enqueues the lead's launches itself and returns once every worker has
//                enqueued: the n fill kernels start within a few microseconds of each other instead of one launch
//                latency apart;
//   exchange     PEER: every member's fill kernel flushes into a partial-histogram buffer and publishes an epoch flag;
//                the lead's exchange+likelihood kernel (llh_pull_kernel) reads all partials with peer-to-peer loads over
//                NVLink, sums them in member order and reduces -lnL in the same launch;
//                NCCL: every member issues ONE ncclAllReduce (in place, f64 sum) on its stream behind its fill, the lead
//                then launches the likelihood reduction.  libnccl.so.2 is opened at run time: no link dependency.
#include "m3b_handle.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <dlfcn.h>
#include <mutex>
#include <thread>

namespace {

// ---- the four NCCL entry points the NCCL arm needs, resolved from libnccl.so.2 at run time ------------------------
struct NcclApi {
  void* lib = nullptr;
  int (*CommInitAll)(void** comms, int ndev, const int* devlist) = nullptr;
  int (*CommDestroy)(void* comm) = nullptr;
  int (*AllReduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t s) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string& err) {
    if (lib) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) { lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) { err = std::string("cannot open libnccl.so.2: ") + dlerror(); return false; }
    CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
    CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
    AllReduce = reinterpret_cast<decltype(AllReduce)>(dlsym(lib, "ncclAllReduce"));
    GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
    if (!CommInitAll || !CommDestroy || !AllReduce) { err = "libnccl.so.2 lacks ncclCommInitAll / ncclAllReduce"; return false; }
    return true;
  }
};
constexpr int kNcclFloat64 = 8, kNcclSum = 0;     // ncclDataType_t::ncclFloat64, ncclRedOp_t::ncclSum (nccl.h)

}  // namespace

struct m3b_group {
  m3b_config cfg{};
  std::vector<m3b_handle*> members;
  std::vector<int> devices;
  std::vector<int64_t> event_begin;          // first event of every member (+ total at the end), set at connect
  int exchange = M3B_EXCHANGE_PEER;
  bool connected = false;
  std::string err;
  std::vector<void*> host_allocs;
  // NCCL arm
  NcclApi nccl;
  std::vector<void*> comms;
  // workers
  std::vector<std::thread> workers;
  std::atomic<uint64_t> seq{0};              // steps published by the caller
  std::atomic<int> enqueued{0};              // workers that have enqueued step `seq`
  std::atomic<bool> quit{false};
  std::atomic<int> sleepers{0};
  std::mutex mu;
  std::condition_variable cv;
  // the step being published (group-owned copies: the caller's arrays may change as soon as m3b_group_step returns)
  std::vector<double> pars, norms;
  const float* osc = nullptr;
  std::vector<int> rc;
  std::vector<std::string> rc_msg;
};

static int gfail(m3b_group* g, int code, const std::string& msg) {
  m3b_last_error_slot() = msg;
  if (g) g->err = msg;
  return code;
}
#define GREQUIRE(cond, code, msg) do { if (!(cond)) return gfail(g, code, std::string(msg)); } while (0)
#define GMEMBER(call) do { int rc__ = (call); if (rc__ != M3B_OK) return gfail(g, rc__, m3b_last_error(nullptr)); } while (0)

// one member's part of a step, on the calling thread (the device must be current for NCCL; the handle sets it itself)
static int member_step(m3b_group* g, int i) {
  m3b_handle* h = g->members[i];
  const double* sp = g->pars.empty() ? nullptr : g->pars.data();
  const double* nm = g->norms.empty() ? nullptr : g->norms.data();
  const float* osc = g->osc;
  if (osc && !h->d_osc_idx) osc += g->event_begin[i];          // one array over all events: this member's range
  if (h->n_events == 0) return M3B_OK;                          // an empty shard (more devices than tiles) has nothing to add
  if (g->exchange == M3B_EXCHANGE_PEER) return m3b_step_peer(h, sp, nm, osc);
  int rc = m3b_step_fill(h, sp, nm, osc);
  if (rc != M3B_OK) return rc;
  void* hist = nullptr; int32_t nb = 0, live = 0;
  rc = m3b_hist_device_ptr(h, &hist, &nb, &live);
  if (rc != M3B_OK) return rc;
  const int nrc = g->nccl.AllReduce(hist, hist, static_cast<size_t>(nb) * (live ? 2 : 1), kNcclFloat64, kNcclSum, g->comms[i], h->stream);
  if (nrc != 0) return fail(h, M3B_ERR_PEER, std::string("ncclAllReduce: ") + (g->nccl.GetErrorString ? g->nccl.GetErrorString(nrc) : "error"));
  if (i == 0) rc = m3b_llh_from_hist(h);
  return rc;
}

static void worker_main(m3b_group* g, int i) {
  cudaSetDevice(g->devices[i]);
  uint64_t seen = 0;
  while (true) {
    // park: spin ~200 us (MCMC steps follow each other within that), then sleep
    const auto t0 = std::chrono::steady_clock::now();
    int spins = 0;
    while (g->seq.load(std::memory_order_acquire) == seen && !g->quit.load(std::memory_order_acquire)) {
      if (++spins < 64) continue;
      spins = 0;
      if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200)) {
        std::unique_lock<std::mutex> lk(g->mu);
        g->sleepers.fetch_add(1);
        // (sequentially consistent accesses on both sides: either this predicate sees the new step, or the caller sees
        //  the sleeper and notifies under the mutex)
        g->cv.wait(lk, [&] { return g->seq.load() != seen || g->quit.load(); });
        g->sleepers.fetch_sub(1);
      }
    }
    if (g->quit.load(std::memory_order_acquire)) return;
    seen = g->seq.load(std::memory_order_acquire);
    g->rc[i] = member_step(g, i);
    if (g->rc[i] != M3B_OK) g->rc_msg[i] = m3b_last_error(g->members[i]);
    g->enqueued.fetch_add(1, std::memory_order_release);
  }
}

extern "C" {

M3B_API int m3b_group_create(const m3b_config* cfg, const int32_t* devices, int32_t n_devices, m3b_group** out) {
  m3b_group* g = nullptr;
  GREQUIRE(cfg && devices && out, M3B_ERR_INVALID, "m3b_group_create: null argument");
  GREQUIRE(n_devices >= 1 && n_devices <= 8, M3B_ERR_INVALID, "m3b_group_create: 1..8 devices");
  g = new m3b_group();
  g->cfg = *cfg;
  for (int i = 0; i < n_devices; ++i) {
    m3b_config c = *cfg;
    c.device = devices[i];
    c.flags |= M3B_FLAG_NO_FUSED_LLH;
    m3b_handle* h = nullptr;
    const int rc = m3b_create(&c, &h);
    if (rc != M3B_OK) {
      const std::string msg = m3b_last_error(nullptr);
      for (m3b_handle* m : g->members) m3b_destroy(m);
      delete g;
      return gfail(nullptr, rc, msg);
    }
    g->members.push_back(h);
    g->devices.push_back(devices[i]);
  }
  g->rc.assign(n_devices, M3B_OK);
  g->rc_msg.assign(n_devices, std::string());
  *out = g;
  return M3B_OK;
}

M3B_API void m3b_group_destroy(m3b_group* g) {
  if (!g) return;
  g->quit.store(true);
  { std::lock_guard<std::mutex> lk(g->mu); g->cv.notify_all(); }
  for (std::thread& t : g->workers) if (t.joinable()) t.join();
  for (m3b_handle* h : g->members) { cudaSetDevice(h->device); cudaStreamSynchronize(h->stream); }
  for (size_t i = 0; i < g->comms.size(); ++i) if (g->comms[i]) g->nccl.CommDestroy(g->comms[i]);
  for (void* p : g->host_allocs) cudaFreeHost(p);
  for (m3b_handle* h : g->members) m3b_destroy(h);
  delete g;
}

M3B_API const char* m3b_group_last_error(const m3b_group* g) { return g ? g->err.c_str() : m3b_last_error(nullptr); }
M3B_API int32_t m3b_group_size(const m3b_group* g) { return g ? static_cast<int32_t>(g->members.size()) : 0; }
M3B_API m3b_handle* m3b_group_member(m3b_group* g, int32_t i) {
  return (g && i >= 0 && i < static_cast<int32_t>(g->members.size())) ? g->members[i] : nullptr;
}

M3B_API int m3b_group_shard(const m3b_group* cg, int64_t n_events, int32_t i, int64_t* e0, int64_t* e1) {
  m3b_group* g = const_cast<m3b_group*>(cg);
  GREQUIRE(g && e0 && e1 && n_events >= 0, M3B_ERR_INVALID, "m3b_group_shard: bad argument");
  const int64_t n = static_cast<int64_t>(g->members.size());
  GREQUIRE(i >= 0 && i < n, M3B_ERR_INVALID, "m3b_group_shard: member out of range");
  int64_t per = (n_events + n - 1) / n;
  per = (per + 1023) / 1024 * 1024;                     // never split a tile row (largest tile: 1024 events)
  *e0 = std::min<int64_t>(n_events, i * per);
  *e1 = std::min<int64_t>(n_events, (i + 1) * per);
  return M3B_OK;
}

M3B_API int m3b_group_upload_spline_monolith(m3b_group* g, int32_t n_params, int32_t max_knots, const float* coeff_x,
                                             const int16_t* n_pts, int64_t n_events, const uint32_t* nParamPerEvent,
                                             const int16_t* paramNo_arr, const uint32_t* nKnots_arr, uint32_t total_knots,
                                             const float* coeff_many, const uint32_t* nParamPerEvent_tf1,
                                             const int16_t* paramNo_tf1, const float* coeff_tf1) {
  GREQUIRE(g && nParamPerEvent && nParamPerEvent_tf1 && n_events >= 0, M3B_ERR_INVALID, "m3b_group_upload_spline_monolith: null argument");
  // response offsets of every event: {count,start} pairs carry 32-bit starts in the reference; recompute in 64 bits
  std::vector<uint64_t> sc(static_cast<size_t>(n_events) + 1, 0), sl(static_cast<size_t>(n_events) + 1, 0);
  for (int64_t e = 0; e < n_events; ++e) { sc[e + 1] = sc[e] + nParamPerEvent[2 * e]; sl[e + 1] = sl[e] + nParamPerEvent_tf1[2 * e]; }
  std::vector<uint64_t> rel;
  for (size_t i = 0; i < g->members.size(); ++i) {
    int64_t e0 = 0, e1 = 0;
    m3b_group_shard(g, n_events, static_cast<int32_t>(i), &e0, &e1);
    m3b_handle* h = g->members[i];
    GMEMBER(m3b_splines_begin(h, n_params, max_knots, coeff_x, n_pts, e1 - e0));
    if (e1 > e0) {
      const uint64_t oc = sc[e0], nc = sc[e1] - sc[e0], ol = sl[e0];
      const uint64_t k0 = nc ? nKnots_arr[oc] : 0;
      const uint64_t k1 = nc ? (sc[e1] < sc[n_events] ? nKnots_arr[sc[e1]] : total_knots) : 0;
      rel.resize(nc);
      for (uint64_t s = 0; s < nc; ++s) rel[s] = nKnots_arr[oc + s] - k0;
      GMEMBER(m3b_splines_append(h, e1 - e0, nParamPerEvent + 2 * e0, paramNo_arr ? paramNo_arr + oc : nullptr, rel.data(), k1 - k0,
                                 coeff_many ? coeff_many + 4 * k0 : nullptr, nParamPerEvent_tf1 + 2 * e0,
                                 paramNo_tf1 ? paramNo_tf1 + ol : nullptr, coeff_tf1 ? coeff_tf1 + 2 * ol : nullptr));
    }
    GMEMBER(m3b_splines_end(h));
  }
  return M3B_OK;
}

M3B_API int m3b_group_upload_binning_ex(m3b_group* g, int32_t n_samples, const int32_t* n_dim, const int32_t* uniform,
                                        const int32_t* nbins, const double* edges) {
  GREQUIRE(g, M3B_ERR_INVALID, "null group");
  for (m3b_handle* h : g->members) GMEMBER(m3b_upload_binning_ex(h, n_samples, n_dim, uniform, nbins, edges));
  return M3B_OK;
}

M3B_API int m3b_group_upload_events(m3b_group* g, int64_t n_events, const int32_t* sample_id, const double* kin,
                                    int32_t n_norm_per_event, const int16_t* norm_idx, int32_t n_norm_values,
                                    int32_t use_osc, const int32_t* osc_idx, int64_t n_osc_values, const float* static_w) {
  GREQUIRE(g && sample_id && kin && n_events > 0, M3B_ERR_INVALID, "m3b_group_upload_events: bad argument");
  m3b_handle* lead = g->members[0];
  GREQUIRE(lead->n_samples > 0, M3B_ERR_STATE, "m3b_group_upload_events: upload the binning first");
  int max_dim = 0;
  for (int s = 0; s < lead->n_samples; ++s) max_dim = std::max(max_dim, lead->b_ndim[s]);
  std::vector<double> k;
  for (size_t i = 0; i < g->members.size(); ++i) {
    int64_t e0 = 0, e1 = 0;
    m3b_group_shard(g, n_events, static_cast<int32_t>(i), &e0, &e1);
    const int64_t n = e1 - e0;
    if (n == 0) continue;                               // more devices than tile rows: this member stays empty
    k.resize(static_cast<size_t>(max_dim) * n);         // kin is dimension-major over ALL events: re-pack the member's columns
    for (int d = 0; d < max_dim; ++d) std::copy(kin + d * n_events + e0, kin + d * n_events + e1, k.begin() + static_cast<size_t>(d) * n);
    GMEMBER(m3b_upload_events(g->members[i], n, sample_id + e0, k.data(), n_norm_per_event,
                              norm_idx ? norm_idx + e0 * n_norm_per_event : nullptr, n_norm_values, use_osc,
                              osc_idx ? osc_idx + e0 : nullptr, n_osc_values, static_w ? static_w + e0 : nullptr));
  }
  return M3B_OK;
}

M3B_API int m3b_group_upload_selection(m3b_group* g, int32_t n_cuts, const int32_t* cut_sample, const int32_t* cut_var,
                                       const double* lower, const double* upper, int32_t n_vars, const double* values) {
  GREQUIRE(g, M3B_ERR_INVALID, "null group");
  int64_t n_events = 0;
  for (m3b_handle* h : g->members) n_events += h->n_events;
  std::vector<double> v;
  int64_t e0 = 0;
  for (m3b_handle* h : g->members) {
    const int64_t n = h->n_events;
    if (n == 0) continue;
    v.resize(static_cast<size_t>(std::max(n_vars, 0)) * n);
    for (int j = 0; j < n_vars; ++j) std::copy(values + j * n_events + e0, values + j * n_events + e0 + n, v.begin() + static_cast<size_t>(j) * n);
    GMEMBER(m3b_upload_selection(h, n_cuts, cut_sample, cut_var, lower, upper, n_vars, n_vars > 0 ? v.data() : nullptr));
    e0 += n;
  }
  return M3B_OK;
}

M3B_API int m3b_group_upload_linear_shifts(m3b_group* g, int32_t n_shift_pars, int64_t n_events, const uint32_t* n_per_event,
                                           const int32_t* shift_par, const int32_t* target, const double* coef) {
  GREQUIRE(g && n_per_event, M3B_ERR_INVALID, "m3b_group_upload_linear_shifts: null argument");
  int64_t e0 = 0, k0 = 0;
  for (m3b_handle* h : g->members) {
    const int64_t n = h->n_events;
    if (n == 0) continue;
    GREQUIRE(e0 + n <= n_events, M3B_ERR_INVALID, "m3b_group_upload_linear_shifts: event count differs from the members'");
    int64_t k1 = k0;
    for (int64_t e = e0; e < e0 + n; ++e) k1 += n_per_event[e];
    GMEMBER(m3b_upload_linear_shifts(h, n_shift_pars, n, n_per_event + e0, shift_par ? shift_par + k0 : nullptr, target ? target + k0 : nullptr,
                                     coef ? coef + k0 : nullptr));
    e0 += n; k0 = k1;
  }
  return M3B_OK;
}

M3B_API int m3b_group_set_shift_pars(m3b_group* g, const double* values) {
  GREQUIRE(g && values, M3B_ERR_INVALID, "m3b_group_set_shift_pars: null argument");
  for (m3b_handle* h : g->members) if (h->n_events > 0) GMEMBER(m3b_set_shift_pars(h, values));
  return M3B_OK;
}

M3B_API int m3b_group_upload_data(m3b_group* g, const double* data, int32_t n_bins) {
  GREQUIRE(g, M3B_ERR_INVALID, "null group");
  for (m3b_handle* h : g->members) GMEMBER(m3b_upload_data(h, data, n_bins));
  return M3B_OK;
}

M3B_API int m3b_group_upload_osc(m3b_group* g, const float* osc_w, int64_t n) {
  GREQUIRE(g && osc_w, M3B_ERR_INVALID, "m3b_group_upload_osc: null argument");
  int64_t e0 = 0;
  for (m3b_handle* h : g->members) {
    if (h->n_events == 0 || !h->use_osc) continue;
    if (h->d_osc_idx) GMEMBER(m3b_upload_osc(h, osc_w, n));
    else { GREQUIRE(e0 + h->n_events <= n, M3B_ERR_INVALID, "m3b_group_upload_osc: array shorter than the events"); GMEMBER(m3b_upload_osc(h, osc_w + e0, h->n_events)); }
    e0 += h->n_events;
  }
  return M3B_OK;
}

M3B_API int m3b_group_alloc_host(m3b_group* g, uint64_t bytes, void** ptr) {
  GREQUIRE(g && ptr && bytes, M3B_ERR_INVALID, "m3b_group_alloc_host: bad argument");
  cudaSetDevice(g->devices[0]);
  cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e != cudaSuccess) return gfail(g, M3B_ERR_NOMEM, std::string("m3b_group_alloc_host: ") + cudaGetErrorString(e));
  g->host_allocs.push_back(*ptr);
  return M3B_OK;
}

M3B_API int m3b_group_connect(m3b_group* g, int32_t exchange) {
  GREQUIRE(g, M3B_ERR_INVALID, "null group");
  GREQUIRE(!g->connected, M3B_ERR_STATE, "m3b_group_connect: already connected");
  GREQUIRE(exchange == M3B_EXCHANGE_PEER || exchange == M3B_EXCHANGE_NCCL, M3B_ERR_INVALID, "m3b_group_connect: unknown exchange");
  const int n = static_cast<int>(g->members.size());
  m3b_handle* lead = g->members[0];
  GREQUIRE(lead->n_bins > 0 && lead->n_events > 0, M3B_ERR_STATE, "m3b_group_connect: upload binning and events first (member 0 must hold events)");
  g->event_begin.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) {
    GREQUIRE(g->members[i]->n_bins == lead->n_bins, M3B_ERR_STATE, "m3b_group_connect: members disagree on the binning");
    g->event_begin[i + 1] = g->event_begin[i] + g->members[i]->n_events;
  }
  g->exchange = exchange;
  if (exchange == M3B_EXCHANGE_PEER) {
    // members that hold events take part; the lead reads their partial histograms with peer-to-peer loads
    std::vector<int> act;
    for (int i = 0; i < n; ++i) if (g->members[i]->n_events > 0) act.push_back(i);
    const int world = static_cast<int>(act.size());
    for (int r = 0; r < world; ++r) {
      m3b_handle* h = g->members[act[r]];
      GMEMBER(m3b_peer_alloc(h));
      h->peer_world = world; h->peer_rank = r; h->peer_pull = (r == 0);
      for (int par = 0; par < 2; ++par) { h->peer_partial[par][r] = h->d_partial[par]; h->peer_flag[par][r] = h->d_flags[par]; }
    }
    cudaSetDevice(lead->device);
    for (int r = 1; r < world; ++r) {
      m3b_handle* h = g->members[act[r]];
      if (h->device != lead->device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, lead->device, h->device);
        GREQUIRE(can, M3B_ERR_PEER, "m3b_group_connect: no peer access between the lead device and a member's device (use M3B_EXCHANGE_NCCL)");
        const cudaError_t e = cudaDeviceEnablePeerAccess(h->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return gfail(g, M3B_ERR_PEER, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
      }
      for (int par = 0; par < 2; ++par) { lead->peer_partial[par][r] = h->d_partial[par]; lead->peer_flag[par][r] = h->d_flags[par]; }
    }
  } else {
    for (int i = 0; i < n; ++i) {
      GREQUIRE(g->members[i]->n_events > 0, M3B_ERR_STATE, "m3b_group_connect: the NCCL exchange needs events on every member");
      for (int j = 0; j < i; ++j) GREQUIRE(g->devices[i] != g->devices[j], M3B_ERR_INVALID, "m3b_group_connect: the NCCL exchange needs distinct devices");
    }
    std::string why;
    GREQUIRE(g->nccl.load(why), M3B_ERR_PEER, "m3b_group_connect: " + why);
    g->comms.assign(n, nullptr);
    const int nrc = g->nccl.CommInitAll(g->comms.data(), n, g->devices.data());
    GREQUIRE(nrc == 0, M3B_ERR_PEER, std::string("ncclCommInitAll: ") + (g->nccl.GetErrorString ? g->nccl.GetErrorString(nrc) : "error"));
  }
  for (int i = 1; i < n; ++i) g->workers.emplace_back(worker_main, g, i);
  g->connected = true;
  return M3B_OK;
}

M3B_API int m3b_group_step(m3b_group* g, const double* spline_pars, const double* norm_pars, const float* osc_w) {
  GREQUIRE(g, M3B_ERR_INVALID, "null group");
  GREQUIRE(g->connected, M3B_ERR_STATE, "m3b_group_step: call m3b_group_connect first");
  m3b_handle* lead = g->members[0];
  GREQUIRE(lead->P == 0 || spline_pars, M3B_ERR_INVALID, "m3b_group_step: spline_pars is NULL");
  GREQUIRE(lead->n_norm_values == 0 || norm_pars, M3B_ERR_INVALID, "m3b_group_step: norm_pars is NULL but events carry norm pointers");
  const int n = static_cast<int>(g->members.size());
  if (lead->P > 0) g->pars.assign(spline_pars, spline_pars + lead->P); else g->pars.clear();
  if (lead->n_norm_values > 0) g->norms.assign(norm_pars, norm_pars + lead->n_norm_values); else g->norms.clear();
  g->osc = osc_w;
  g->enqueued.store(0, std::memory_order_relaxed);
  g->seq.fetch_add(1);
  if (g->sleepers.load() > 0) { std::lock_guard<std::mutex> lk(g->mu); g->cv.notify_all(); }
  g->rc[0] = member_step(g, 0);
  if (g->rc[0] != M3B_OK) g->rc_msg[0] = m3b_last_error(lead);
  while (g->enqueued.load(std::memory_order_acquire) < n - 1) { /* spin: every worker is enqueueing right now */ }
  for (int i = 0; i < n; ++i)
    if (g->rc[i] != M3B_OK) return gfail(g, g->rc[i], "m3b_group_step: member " + std::to_string(i) + ": " + g->rc_msg[i]);
  return M3B_OK;
}

M3B_API int m3b_group_llh(m3b_group* g, double* total, double* per_sample) {
  GREQUIRE(g && total, M3B_ERR_INVALID, "m3b_group_llh: null argument");
  GMEMBER(m3b_llh(g->members[0], total, per_sample));
  return M3B_OK;
}

M3B_API int m3b_group_read_hist(m3b_group* g, double* mc, double* w2) {
  GREQUIRE(g, M3B_ERR_INVALID, "null group");
  GMEMBER(m3b_read_hist(g->members[0], mc, w2));         // the lead holds the reduced histograms (both exchanges)
  return M3B_OK;
}

M3B_API int m3b_group_synchronize(m3b_group* g) {
  GREQUIRE(g, M3B_ERR_INVALID, "null group");
  for (m3b_handle* h : g->members) GMEMBER(m3b_synchronize(h));
  return M3B_OK;
}

}  // extern "C"
