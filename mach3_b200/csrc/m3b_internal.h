// m3b_internal.h -- structures shared by the host API (m3b_api.cu) and the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace m3b {

constexpr int kMaxDim = 4;          // kinematic dimensions per sample (matches the C ABI's nbins[s*4+d])
constexpr int kMaxParams = 2048;    // spline parameters per handle (step table lives in shared memory)
constexpr int kMaxNormSlots = 16;   // norm pointers per event

// One tile = T consecutive events.  Cubic coefficients of the tile are stored
//   cub[(segbase(slot) + segment) * T + lane]   float4 {y,b,c,d}
// so that, the active segment being uniform per parameter per step, every warp reads 512
// contiguous bytes per response.  Linear coefficients: lin[slot * T + lane] float2 {a,b}.
struct TileDesc {
  const float4* cub;
  const float2* lin;
  int32_t sig;       // signature: which parameters the tile's slots refer to
  int32_t ncnl;      // the signature's slot counts, nc | nl << 16 (saves the producer a dependent load)
};

// Signature = the (sorted) union of parameters of the tile's events.  pool layout at `off`:
//   cparam[nc] | segbase[nc] | lparam[nl]       (int32 each)
struct SigDesc {
  int32_t nc, nl, off, rows;   // rows = total segment rows of the tile's cubic block
};

// Per-step table, built on the host (SplineBase::FindSplineSegment), copied H2D once per step and
// staged into shared memory by every block with one bulk (TMA) copy:
//   int32 seg[P] | float dx[P] | float val[P] | float norm[Nn]
// and, when the signatures are few enough (n_sigs_x > 0), the same expanded per signature slot so
// the kernels index by slot with no indirection through the signature pool:
//   int32 rowx[n_sigs][max_nc]  first active coefficient row of the slot = segbase + segment
//   float dxx [n_sigs][max_nc]  dx of the slot's parameter
//   float lvx [n_sigs][max_nl]  value of the TF1 slot's parameter
// (padded to 16 B)
struct StepLayout {
  int32_t P, Nn;
  int32_t off_seg, off_dx, off_val, off_norm, bytes;
  int32_t n_sigs_x, max_nc, max_nl, off_rowx, off_dxx, off_lvx;
  int32_t off_val_d, off_norm_d;   // f64 copies of val[P] / norm[Nn] (the reference's default build, M3::float_t = double)
};
constexpr int kMaxExpandedStepBytes = 16384;
constexpr int kStepInlineMax = 3072;      // kernel parameters stay below the classic 4 KB limit
inline StepLayout make_step_layout(int P, int Nn, int n_sigs, int max_nc, int max_nl, bool f64 = false) {
  StepLayout L;
  L.P = P; L.Nn = Nn;
  L.off_seg = 0;
  L.off_dx = L.off_seg + 4 * P;
  L.off_val = L.off_dx + 4 * P;
  L.off_norm = L.off_val + 4 * P;
  int end = L.off_norm + 4 * Nn;
  L.max_nc = max_nc; L.max_nl = max_nl;
  const long long x = 4ll * n_sigs * (2 * max_nc + max_nl);
  L.n_sigs_x = (n_sigs > 0 && x > 0 && x <= kMaxExpandedStepBytes) ? n_sigs : 0;
  L.off_rowx = L.off_dxx = L.off_lvx = 0;
  if (L.n_sigs_x) {
    L.off_rowx = (end + 15) & ~15;
    L.off_dxx = L.off_rowx + 4 * n_sigs * max_nc;
    L.off_lvx = L.off_dxx + 4 * n_sigs * max_nc;
    end = L.off_lvx + 4 * n_sigs * max_nl;
  }
  L.off_val_d = L.off_norm_d = 0;
  if (f64) {
    L.off_val_d = (end + 7) & ~7;
    L.off_norm_d = L.off_val_d + 8 * P;
    end = L.off_norm_d + 8 * Nn;
  }
  L.bytes = (end + 15) & ~15;
  if (L.bytes == 0) L.bytes = 16;
  return L;
}

// ---- BinnedSplineHandler path (m3b_binned.cu) ------------------------------------------------------
// Active (non-flat) binned splines are grouped by parameter and padded to 1024 per parameter:
//   bcoef[coef_off + segment * n_pad + k]  float4 {y,b,c,d}     bx[same index]  knot x of that segment
// so one step reads, per parameter, ONE contiguous row of coefficients and one of x.
struct BTile {                 // 1024 consecutive active splines of one parameter
  int64_t coef_off;            // first element of the parameter's [nseg][n_pad] block
  int32_t n_pad, param, k0, out0;   // row length, parameter, first spline of the tile in the row, first compact weight
};
struct WTile {                 // 32 events (consecutive in the kernel's walking order): their binned-spline pointers as ELL columns
  int64_t off;                 // ell[off + j*32 + lane] = the lane's j-th pointer, = compact weight index, -1 = none
  int32_t max_n, pad;
};

// Shared-memory map of the TMA fill kernel (m3b_fill_tma.cu); computed on the host, passed by value.
//   [step table][dx[max_nc]][lv[max_nl]][row[max_nc]][stage descriptors][hist (+w2)][ring: n_stages x stage_bytes]
struct TmaSmem {
  int32_t off_dx, off_lv, off_row, off_desc, off_hist, off_osc, off_ring, stage_bytes, total;
};

struct FillArgs {
  // tiles: this launch covers [tile_begin, n_tiles)
  const TileDesc* tiles;
  const SigDesc* sigs;
  const int32_t* sig_pool;
  int32_t tile_begin, n_tiles, T, max_nc, max_nl;
  TmaSmem tma;
  // per-step table: in device memory (step_table, one H2D copy per step) or, when it is small
  // enough (the usual case), inside the kernel parameters themselves (step_inline): no copy at all
  const unsigned char* step_table;
  StepLayout step;
  int32_t step_inline_bytes;
  // event table (flat, padded to n_tiles*T)
  const int32_t* bin;
  const float* osc;
  const float* osc_host;       // TMA kernel, zero-copy: this step's weights in pinned host memory, streamed
                               // over PCIe by the producer's bulk copies (no separate H2D pass); the
  float* osc_store;            // consumers also store them here so later steps find them on the device
  const int32_t* osc_idx;
  const float* static_w;
  const int16_t* norm_idx;     // [slot * e_pad + e]
  int32_t norm_slots;
  int64_t e_pad, n_events;
  // histograms
  double* hist;                // [n_bins] mc, this step
  double* w2;                  // [n_bins] or nullptr when W2 is frozen
  double* hist_next;           // zeroed by the last block for the next step (nullptr = caller memsets)
  double* w2_next;
  int32_t n_bins, hist_in_smem;
  int32_t weights_only;        // 1: stop after the per-event weights (no fill, no likelihood)
  // fused likelihood (last block)
  int32_t fuse_llh, test_stat, n_samples;
  const double* data;
  const double* w2_frozen;     // w2 histogram to use in the LLH (== w2 when live)
  const int32_t* sample_start; // [n_samples+1] global bin offsets
  int32_t sample_start_inline[65];   // the same, in the kernel parameters (saves the last block a DRAM round trip)
  unsigned int* ticket;
  int32_t* status;             // handle's status word: bit 1 = a test statistic hit a case where the reference throws
  unsigned int* tile_counter;  // dynamic tile scheduler of the TMA kernel (nullptr: static interleave)
  int32_t n_stages;            // TMA kernel: stages of the shared-memory coefficient ring
  int32_t pdl;                 // TMA kernel launched with programmatic stream serialization: the next step's ramp may
                               // overlap this step's tail; everything the previous launch wrote is touched only after
                               // griddepcontrol.wait
  int32_t binned_pdl;          // binned path: the fill kernel is launched with programmatic stream serialization behind the
                               // eval kernel -- its set-up and first prefetch overlap the eval kernel's tail, the weights
                               // are touched only after griddepcontrol.wait
  int32_t guard_x2;            // TMA kernel: whole tile rows are grabbed while more than guard_x2/2 * grid * g units are left
  double* llh_dev;             // [1+n_samples]
  double* llh_host;            // mapped pinned mirror (nullptr = none)
  unsigned long long* llh_seq_host;   // mapped host word: set to llh_seq after llh_host is written (m3b_llh polls it), or nullptr
  unsigned long long llh_seq;
  // optional per-event outputs
  float* evt_spline_w;
  float* evt_total_w;
  // BinnedSplineHandler path
  const BTile* btiles; int32_t n_btiles;
  const float4* bcoef; const float* bx; float* bw;
  // ... in the reference's default build (M3::float_t = double): coefficients, weights, osc/static weights and the
  // per-event product are double
  int32_t real_f64;
  const double* bcoef_d;       // 4 doubles {y,b,c,d} per element
  const double* bx_d; double* bw_d;
  const double* osc_d; const double* static_d;
  double* evt_spline_d; double* evt_total_d;
  const int32_t* ell; const WTile* wtiles; int64_t n_wtiles;
  const int32_t* perm;         // binned fill: event handled by (warp tile, lane), nullptr = identity
  const int32_t* bin_sorted;   // binned fill: event table in walking order (bin, static weight, norm indices), so the
  const void* static_sorted;   // 32 lanes of a warp read consecutive elements; float or double like static_w / static_d
  const int16_t* norm_idx_sorted; const int32_t* osc_idx_sorted;
  // optional per-block timeline (m3b_block_trace): 8 x u64 globaltimer ns per block
  unsigned long long* trace;
  alignas(16) unsigned char step_inline[kStepInlineMax];
  // peer exchange (multi-GPU, own collective): in peer mode `hist`/`w2` point into this rank's exported
  // partial-histogram buffer; when the whole grid has flushed, the last block publishes `peer_epoch` in
  // this rank's flag and the peers PULL the buffer over NVLink (llh_pull_kernel)
  int32_t peer_world, peer_rank;
  unsigned int* peer_flag_own;
  unsigned int peer_epoch;
};

struct LlhArgs {
  const double* hist; const double* w2; const double* data;
  const int32_t* sample_start;
  int32_t n_bins, n_samples, test_stat;
  double* llh_dev; double* llh_host;
  unsigned long long* llh_seq_host; unsigned long long llh_seq;      // see FillArgs
  // peer (pull) mode: every rank reads all ranks' partial histograms over NVLink peer memory, in rank order
  int32_t peer_world; unsigned int epoch;
  const double* peer_hist[8];        // rank r's exported partial {mc[n_bins], w2[n_bins]} of this epoch's parity
  const unsigned int* peer_flag[8];  // rank r's epoch flag
  double* hist_out; double* w2_out; int32_t w2_live;
  double* partial;                   // [blocks * n_samples] per-block per-sample sums
  unsigned int* ticket;
  int32_t sample_start_inline[65];
  int32_t* status;   // set to 1 on peer timeout
};

struct BinArgs {
  int64_t n_events, e_pad;
  const int32_t* sample_id;
  const double* kin;           // [d * n_events + e]
  int32_t n_samples;
  const int32_t* n_dim;        // [n_samples]
  const int32_t* nbins;        // [n_samples*kMaxDim]
  const int32_t* edge_off;     // [n_samples*kMaxDim] offset into edges
  const int32_t* stride;       // [n_samples*kMaxDim]
  const int32_t* global_off;   // [n_samples]
  const double* edges;
  int32_t* bin;                // [e_pad]
  // non-uniform samples (boxes behind a 10-per-dimension mega-bin grid, Samples/SampleStructs.h:394-528):
  // for them nbins/edge_off/stride describe the mega grid
  const int32_t* uniform;      // [n_samples] (nullptr = all uniform)
  const int32_t* box_off;      // [n_samples] first box of the sample in `boxes` (in boxes)
  const int32_t* grid_off;     // [n_samples] first mega bin of the sample in grid_start
  const double* boxes;         // [box][dim][2] {lo, hi}
  const int32_t* grid_start;   // CSR: boxes overlapping mega bin g are grid_idx[grid_start[g] .. grid_start[g+1])
  const int32_t* grid_idx;
};

// SampleHandlerFD::IsEventSelected on the device (m3b_upload_selection)
struct SelectArgs {
  int64_t n_events, e_pad;
  const int32_t* sample_id;
  const int32_t* cut_start;    // [n_samples+1] CSR over the cut arrays below
  const int32_t* cut_var;      // >= 0: row of `values`; -1-d: row d of `kin`
  const double* lower; const double* upper;
  const double* values;        // [v * n_events + e]
  const double* kin;           // [d * n_events + e] (may be null when no cut refers to it)
  const int32_t* bin_raw;      // FindGlobalBin per event
  int32_t* bin;                // what the fill kernels read: bin_raw, or -1 for a rejected event
  uint8_t* selected;
};

// SampleHandlerFD::ApplyShifts for linear functional parameters (m3b_upload_linear_shifts)
struct ShiftArgs {
  int64_t n_events;
  int32_t n_dims, n_sel_vars;
  const double* kin_nom; double* kin;              // [d * n_events + e]
  const double* sel_nom; double* sel;              // [v * n_events + e]  (may be null)
  const int64_t* start;                            // [n_events + 1] CSR over the entries below
  const int32_t* par; const int32_t* target; const double* coef;
  const double* theta;                             // [n_shift_pars] this step's parameter values
};

struct RetileArgs {
  int64_t n;                   // events in chunk
  int64_t tile0_event;         // chunk's first event is the first lane of a tile
  int32_t T, P;
  const uint64_t* start_c;     // [n+1]
  const int16_t* paramNo;
  const uint64_t* knot_off;
  const float4* coeff_many;
  const uint64_t* start_l;     // [n+1]
  const int16_t* paramNo_l;
  const float2* coeff_l;
  const int32_t* tile_sig;     // [tiles in chunk]
  const int16_t* slot_of_param;// [n_sigs * P]  slot (cubic or linear numbering) or -1
  const int32_t* segbase_of_param; // [n_sigs * P] first segment row of the parameter's slot
  const int16_t* nseg;         // [P] segments stored per parameter (n_pts-1)
  const uint64_t* tile_cub_off;// [tiles] float4 offset of tile's cubic block in chunk pool
  const uint64_t* tile_lin_off;// [tiles] float2 offset
  float4* cub_pool;
  float2* lin_pool;
};

// Opt-in dynamic shared memory of a kernel: always raised to the LARGEST size the device allows for it, never to the
// size one handle happens to need.  The attribute is per function and per device, i.e. shared by every handle (and every
// worker thread of a group) on that device: setting it per handle would let a later, smaller handle pull it below what an
// earlier one launches with.
template <class Kernel>
inline cudaError_t allow_max_dynamic_smem(Kernel k) {
  cudaFuncAttributes fa{};
  cudaError_t e = cudaFuncGetAttributes(&fa, k);
  if (e != cudaSuccess) return e;
  int dev = 0, optin = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - static_cast<int>(fa.sharedSizeBytes));
}

// launchers (m3b_kernels.cu)
cudaError_t launch_fill(const FillArgs& a, int variant, int grid, int smem_bytes, cudaStream_t s);
cudaError_t launch_llh(const LlhArgs& a, cudaStream_t s);
cudaError_t launch_llh_pull(const LlhArgs& a, int blocks, cudaStream_t s);
constexpr int kLlhPullMaxBlocks = 32;
cudaError_t launch_bins(const BinArgs& a, cudaStream_t s);
cudaError_t launch_select(const SelectArgs& a, cudaStream_t s);
cudaError_t launch_table_copy(float* dst, const float* src_host_devptr, int64_t n, cudaStream_t s);
cudaError_t launch_shift(const ShiftArgs& a, cudaStream_t s);
cudaError_t launch_retile(const RetileArgs& a, int64_t n_identity_cub, int64_t n_identity_lin, cudaStream_t s);
cudaError_t fill_occupancy(int T, int variant, int smem_bytes, int* blocks_per_sm);
cudaError_t fill_set_smem(int T, int variant, int smem_bytes);
// TMA streaming kernel (m3b_fill_tma.cu): tile rows of T = 256, 512 or 1024 events
TmaSmem tma_smem_layout(const StepLayout& step, int max_nc, int max_nl, int n_bins, bool hist_in_smem, bool w2_live,
                        bool osc_slots, int n_stages);
cudaError_t launch_fill_tma(const FillArgs& a, int grid, int smem_bytes, cudaStream_t s);
cudaError_t fill_tma_set_smem(int smem_bytes);
cudaError_t fill_tma_occupancy(int smem_bytes, int* blocks_per_sm);
int fill_smem_bytes(const FillArgs& a, bool hist_in_smem, bool w2_live);
// BinnedSplineHandler path (m3b_binned.cu)
cudaError_t launch_binned_eval(const FillArgs& a, int grid, cudaStream_t s);
cudaError_t launch_binned_fill(const FillArgs& a, int grid, int threads, int smem_bytes, cudaStream_t s);
cudaError_t binned_fill_prepare(int smem_bytes, bool f64, int threads, int* blocks_per_sm);   // opt-in smem + occupancy
int binned_fill_smem_bytes(const FillArgs& a, bool hist_in_smem, bool w2_live);
cudaError_t launch_gather(void* dst, const void* src, const int32_t* perm, int64_t n, int elem_bytes, cudaStream_t s);

}  // namespace m3b
