// m3b_handle.h -- the handle behind the C ABI and the small host helpers shared by the .cu files
// that implement it (m3b_api.cu: monolith path; m3b_binned.cu: BinnedSplineHandler path).
#pragma once
#include "m3b200.h"
#include "m3b_internal.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace m3b;

int m3b_batch_try(m3b_handle* h, int32_t n_sets, const double* spline_pars, const double* norm_pars, const float* osc_w,
                  double* host_slots_dev, int* done);      // m3b_batch.cu
int m3b_batch2_try(m3b_handle* h, int32_t n_sets, const double* spline_pars, const double* norm_pars, const float* osc_w,
                   double* host_slots_dev, int* done);     // m3b_batch2.cuh (same translation unit)
std::string& m3b_last_error_slot();     // thread-local last error (defined in m3b_api.cu)
int m3b_peer_alloc(m3b_handle* h);      // the exported partial-histogram buffers + epoch flag of the peer exchange (m3b_api.cu)

// Tuning knobs read from the environment exist only in -DM3B_EXPERIMENTS builds (A/B measurements); the product
// library takes its configuration from m3b_config alone.
static inline const char* experiment_env(const char* name) {
#ifdef M3B_EXPERIMENTS
  return getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

struct m3b_handle {
  m3b_config cfg{};
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;

  // ---- spline parameters (FastSplineInfo, Splines/SplineStructs.h:21-44)
  int P = 0, Kmax = 0;
  std::vector<float> coeff_x;
  std::vector<double> knots_d;   // optional FastSplineInfo::xPts in double (m3b_set_spline_knots_f64); empty: coeff_x
  std::vector<int16_t> n_pts;
  std::vector<int16_t> curr_segment;     // FastSplineInfo::CurrSegment
  std::vector<int16_t> segments;         // SplineBase::SplineSegments
  std::vector<float> param_values;       // SplineBase::ParamValues
  std::vector<int16_t> nseg;             // stored segments per parameter (n_pts-1)
  bool splines_open = false, splines_done = false;
  int64_t n_events_total = 0, n_events_loaded = 0;
  int T = 256;
  bool T_auto = false;                   // cfg.tile_events == 0: chosen from the event count at m3b_splines_begin

  // ---- signatures and tiles
  std::map<std::vector<int16_t>, int> sig_index;   // key: cubic params, -1, linear params
  std::vector<SigDesc> sigs;
  std::vector<int32_t> sig_pool;
  std::vector<int16_t> sig_slot_of_param;          // [n_sigs*P]
  std::vector<int32_t> sig_segbase_of_param;       // [n_sigs*P]
  std::vector<TileDesc> tiles;
  std::vector<void*> allocs;
  uint64_t device_bytes = 0;
  uint64_t active_coef_bytes = 0;
  int max_nc = 0, max_nl = 0;
  TileDesc* d_tiles = nullptr;
  SigDesc* d_sigs = nullptr;
  int32_t* d_sig_pool = nullptr;
  bool tiles_dirty = true;

  // ---- binning
  int n_samples = 0, n_bins = 0;
  std::vector<int32_t> b_ndim, b_nbins, b_edge_off, b_stride, b_goff, sample_start;
  std::vector<double> b_edges;
  int32_t *d_ndim = nullptr, *d_nbins = nullptr, *d_edge_off = nullptr, *d_stride = nullptr, *d_goff = nullptr,
          *d_sample_start = nullptr;
  double* d_edges = nullptr;
  // non-uniform samples
  std::vector<int32_t> b_uniform, b_box_off, b_grid_off, b_grid_start, b_grid_idx;
  std::vector<double> b_boxes;
  int32_t *d_uniform = nullptr, *d_box_off = nullptr, *d_grid_off = nullptr, *d_grid_start = nullptr, *d_grid_idx = nullptr;
  double* d_boxes = nullptr;

  // ---- events
  int64_t n_events = 0, e_pad = 0, n_tiles = 0;
  int32_t* d_bin = nullptr;              // what the fill kernels read (bin_raw, or -1 for events the selection drops)
  int32_t* d_bin_raw = nullptr;          // FindGlobalBin per event; == d_bin while no selection is uploaded
  // selection (SampleHandlerFD::IsEventSelected): cuts CSR by sample + the caller's cut-variable table
  int n_cuts = 0, n_sel_vars = 0;
  bool sel_uses_kin = false;
  int32_t *d_cut_start = nullptr, *d_cut_var = nullptr;
  double *d_cut_lo = nullptr, *d_cut_hi = nullptr, *d_sel_vals = nullptr;
  uint8_t* d_selected = nullptr;
  // linear functional shifts (SampleHandlerFD::ApplyShifts on the device): nominal copies + per-event entry lists
  int n_shift_pars = 0;
  double *d_kin_nom = nullptr, *d_sel_vals_nom = nullptr, *d_shift_theta = nullptr, *d_sh_coef = nullptr;
  int64_t* d_sh_start = nullptr;
  int32_t *d_sh_par = nullptr, *d_sh_target = nullptr;
  int32_t* d_osc_idx = nullptr;
  float* d_osc = nullptr;
  int64_t n_osc = 0;
  bool use_osc = false;
  float* d_static = nullptr;
  int16_t* d_norm_idx = nullptr;
  int norm_slots = 0, n_norm_values = 0;
  double* d_kin = nullptr;
  int kin_dims = 0;
  int32_t* d_sample_id = nullptr;
  float *d_evt_spline_w = nullptr, *d_evt_total_w = nullptr;
  bool evt_weights_valid = false;

  // ---- histograms, likelihood
  double* d_hw[2] = {nullptr, nullptr};   // each {mc[n_bins], w2[n_bins]}
  bool mc_zero[2] = {false, false}, w2_zero[2] = {false, false};
  int cur = 0;                            // buffer of the last step
  double* d_w2_frozen = nullptr;
  double* d_data = nullptr;
  unsigned int* d_ticket = nullptr;
  double* d_llh = nullptr;
  double* h_llh = nullptr;                // mapped pinned
  double* h_llh_dev = nullptr;            // device alias of h_llh
  // the step's sequence word (mapped pinned): the kernel that writes -lnL sets it to seq_issued afterwards; m3b_llh polls
  // it instead of waiting for the stream (the likelihood kernel's retirement and the driver's wake-up are off the path)
  unsigned long long* h_seq = nullptr; unsigned long long* h_seq_dev = nullptr;
  unsigned long long seq_issued = 0, seq_wait = 0;      // seq_wait == 0: the last step published no sequence word
  int test_stat = 0;
  bool first_time_w2 = true;
  bool last_w2_live = false;

  // ---- per-step staging
  StepLayout step{};
  int step_sigs = -1;
  static constexpr int kRing = 4;
  unsigned char* h_step[kRing] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t step_ev[kRing] = {nullptr, nullptr, nullptr, nullptr};
  unsigned char* d_step[kRing] = {nullptr, nullptr, nullptr, nullptr};
  int ring = 0;
  std::vector<unsigned char> last_step_table;
  bool have_step = false;

  // ---- launch configuration
  int grid = 0, smem = 0, variant = 1;   // LDG kernel variant (m3b_kernels.cu), used when use_tma is false
  bool use_tma = true;                   // streaming TMA kernel (m3b_fill_tma.cu), the default
  int tma_stages = 0;
  TmaSmem tma{};
  unsigned int* d_tile_counter = nullptr;
  bool zc_slots = false;                   // shared-memory slots for zero-copy oscillation weights
  unsigned long long* d_trace = nullptr;   // m3b_block_trace
  int trace_grid = 0;
  bool hist_in_smem = true;
  bool launch_ready = false;
  bool launch_w2_live = false;

  // ---- peer exchange
  int peer_world = 0, peer_rank = 0;
  bool peer_pull = true;                         // false: a non-lead member of a single-process group (m3b_group.cu) only publishes
  double* d_partial[2] = {nullptr, nullptr};     // this rank's exported partial histograms (two epochs' parity), {mc,w2}[n_bins]
  unsigned int* d_flags[2] = {nullptr, nullptr}; // [world]
  double* peer_partial[2][8] = {};
  unsigned int* peer_flag[2][8] = {};
  unsigned int peer_epoch = 0;
  double* d_llh_partial = nullptr; unsigned int* d_llh_ticket = nullptr;
  int32_t* d_status = nullptr;
  std::vector<void*> ipc_opened;
  std::vector<void*> registered;
  std::vector<void*> host_allocs;        // m3b_alloc_host

  // ---- BinnedSplineHandler path (m3b_binned.cu)
  bool binned = false;
  bool f64 = false;                         // binned path in the reference's default build (M3::float_t = double)
  std::vector<double> coeff_x_d;            // knots as the default build holds them
  std::vector<double> spline_pars_last;     // the un-narrowed parameter values of the step being enqueued
  double *d_bcoef_d = nullptr, *d_bx_d = nullptr, *d_bw_d = nullptr, *d_osc_d = nullptr, *d_static_d = nullptr,
         *d_evt_spline_d = nullptr, *d_evt_total_d = nullptr;
  int64_t b_n_slots = 0, b_n_act = 0, b_n_act_pad = 0, n_wtiles = 0;
  int32_t n_btiles = 0;
  std::vector<int32_t> b_slot2compact, b_compact2slot;
  std::vector<int64_t> b_out_base, b_count;      // per parameter: first compact index of its (padded) weight row / non-flat splines in it
  int b_param_of_compact_row(int64_t c) const {  // parameter whose weight row holds compact index c
    return static_cast<int>(std::upper_bound(b_out_base.begin(), b_out_base.end(), c) - b_out_base.begin()) - 1;
  }
  // slot layout as the upload found it: maximal runs of consecutive slots belonging to one parameter ("blocks": one
  // systematic's (mode, var1, var2, var3) grid), and for each the number of times the parameter index went down before it
  // ("super-block": one (sample, oscillation channel) in the reference's [sample][osc][syst][mode][var1][var2][var3] order)
  std::vector<int64_t> b_run_start; std::vector<int32_t> b_run_super;
  int32_t* d_perm = nullptr;                     // order in which the binned fill kernel walks the events (sorted by spline-grid cell)
  int32_t* d_bin_sorted = nullptr; void* d_static_sorted = nullptr; int16_t* d_norm_idx_sorted = nullptr; int32_t* d_osc_idx_sorted = nullptr;   // event table in walking order
  float4* d_bcoef = nullptr; float* d_bx = nullptr; float* d_bw = nullptr;
  BTile* d_btiles = nullptr; WTile* d_wtiles = nullptr; int32_t* d_ell = nullptr;
  uint64_t b_gather_per_step = 0;
  int binned_eval_grid = 0, binned_threads = 256;

  // ---- batched proposals (m3b_step_batch): one -lnL slot per set in mapped host memory
  double* h_batch = nullptr; double* h_batch_dev = nullptr; size_t batch_cap = 0;
  double* llh_host_override = nullptr;
  // staging of the batched kernel (m3b_batch.cu), grown on demand
  void *bt_dx = nullptr, *bt_rowoff = nullptr, *bt_val = nullptr, *bt_rowlist = nullptr, *bt_norm = nullptr, *bt_sigs = nullptr,
       *bt_hist = nullptr, *bt_llh = nullptr, *bt_slot = nullptr, *bt_group = nullptr, *bt_rank8 = nullptr;
  size_t bt_dx_cap = 0, bt_rowoff_cap = 0, bt_val_cap = 0, bt_rowlist_cap = 0, bt_norm_cap = 0, bt_sigs_cap = 0, bt_hist_cap = 0,
         bt_llh_cap = 0, bt_slot_cap = 0, bt_group_cap = 0, bt_rank8_cap = 0;

  uint64_t steps = 0, launches = 0;

  // ---- optional kernel timing
  bool timing = false;
  std::vector<cudaEvent_t> tev;   // pairs
  size_t tev_used = 0;
};

// ------------------------------------------------------------------------------------------------
static inline int fail(m3b_handle* h, int code, const std::string& msg) {
  m3b_last_error_slot() = msg;
  if (h) h->err = msg;
  return code;
}
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      char b__[512];                                                                               \
      snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return fail(h, M3B_ERR_CUDA, b__);                                                           \
    }                                                                                              \
  } while (0)
#define REQUIRE(cond, code, msg)                                                                   \
  do { if (!(cond)) return fail(h, code, std::string(msg)); } while (0)

template <class Tp>
static inline cudaError_t dev_alloc(m3b_handle* h, Tp** p, size_t n) {
  if (n == 0) n = 1;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(Tp));
  if (e == cudaSuccess) { h->allocs.push_back(*p); h->device_bytes += n * sizeof(Tp); }
  return e;
}
// Set-up copies go through the HANDLE'S stream and are waited for: the handle's kernels run on a non-blocking stream,
// which is not ordered with the legacy default stream a plain cudaMemcpy uses (and a pageable host-to-device cudaMemcpy
// may return before its DMA has landed, a device-to-device one before it has run at all).
static inline cudaError_t copy_sync(m3b_handle* h, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
  if (bytes == 0) return cudaSuccess;
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, h->stream);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(h->stream);
}
template <class Tp>
static inline cudaError_t dev_upload(m3b_handle* h, Tp** p, const std::vector<Tp>& v) {
  cudaError_t e = dev_alloc(h, p, v.size());
  if (e != cudaSuccess) return e;
  return copy_sync(h, *p, v.data(), v.size() * sizeof(Tp), cudaMemcpyHostToDevice);
}

