/*
 * m3b200.h -- C ABI of libm3b200.so: the B200-native (sm_100a) per-MCMC-step likelihood hot path
 * of MaCh3:  event-by-event TSpline3 / TF1 response evaluation -> per-event total weight ->
 * histogram fill -> Poisson / Barlow-Beeston -lnL, fused on the device.
 *
 * The reference (mach3-software/MaCh3 v2.4.2) has no FFI for this path; its seam is C++ virtual
 * dispatch plus ONE class that isolates CUDA from host code, `SMonolithGPU`
 * (Splines/gpuSplineUtils.cuh:63-218).  Every entry point below names the reference interface it
 * replaces (file:line relative to the MaCh3 tree).  INTEGRATION.md shows the reference-side
 * bindings (a drop-in SMonolithGPU, a SampleHandlerFD::Reweight/GetLikelihood override, ctypes).
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host buffer; the library copies on
 *     upload and never returns pointers into its own host memory (device pointers are returned
 *     only by the explicit *_device_ptr calls used to wire collectives).
 *   - every function returns an m3b_status; m3b_last_error() gives the message.  Unlike the
 *     reference's CudaCheckError (a no-op in release builds, Manager/gpuUtils.cu:18-35) every CUDA
 *     call is checked.  There is NO CPU fallback: without a usable sm_100 device m3b_create fails.
 *   - one handle = one SampleHandlerFD (+ its SMonolith).  Not thread-safe per handle, like the
 *     reference (one host thread drives the chain).
 *   - m3b_step* are asynchronous on the handle's stream; m3b_llh / m3b_read_* synchronise, which is
 *     the contract of SplineBase::Evaluate + SynchroniseMemTransfer (Splines/SplineBase.h:35,50).
 *   - types follow the reference's _LOW_MEMORY_STRUCTS_ build (float weights/coefficients, short
 *     segments, double histograms and likelihood; Manager/Core.h:27-35).
 */
#ifndef M3B200_H
#define M3B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define M3B_API __attribute__((visibility("default")))
#else
#define M3B_API
#endif

typedef struct m3b_handle m3b_handle;

typedef enum {
  M3B_OK = 0,
  M3B_ERR_INVALID = 1,   /* bad argument / inconsistent arrays                                     */
  M3B_ERR_CUDA = 2,      /* a CUDA runtime call or kernel failed                                   */
  M3B_ERR_STATE = 3,     /* call order violated (e.g. step before upload)                          */
  M3B_ERR_KNOTS = 4,     /* a spline's knot count differs from its parameter's (reference assumes  */
                         /* identical knots per parameter, Splines/SplineMonolith.cpp:102-104)     */
  M3B_ERR_NOMEM = 5,
  M3B_ERR_NODEVICE = 6,  /* no sm_100 device: the product path has no CPU fallback                 */
  M3B_ERR_PEER = 7,      /* peer (multi-GPU) exchange failed or timed out                          */
  M3B_ERR_MATH = 8       /* a test statistic hit a case in which the reference throws MaCh3Exception  */
                         /* (Barlow-Beeston negative discriminant, Samples/SampleHandlerBase.cpp:64-67) */
} m3b_status;

/* Samples/SampleStructs.h:105-112 (same numeric values as enum TestStatistic) */
typedef enum {
  M3B_POISSON = 0, M3B_BARLOW_BEESTON = 1, M3B_ICECUBE = 2, M3B_PEARSON = 3, M3B_DEMBINSKI_ABDELMOTTELEB = 4
} m3b_test_statistic;

enum {
  M3B_FLAG_KEEP_EVENT_WEIGHTS = 1, /* write per-event spline + total weights every step (+8 B/event)   */
  M3B_FLAG_KEEP_KINEMATICS    = 2, /* keep kinematic variables on the device so bins can be recomputed */
  M3B_FLAG_NO_FUSED_LLH       = 4, /* never fuse the LLH into the fill kernel (multi-GPU callers)      */
  M3B_FLAG_NO_BATCH_KERNEL    = 8, /* m3b_step_batch always runs sequential single-set launches         */
  M3B_FLAG_BATCH_KERNEL_V1    = 16, /* m3b_step_batch uses the first-generation batched kernel (A/B)     */
  M3B_FLAG_NO_SPIN_LLH        = 32  /* m3b_llh waits with cudaStreamSynchronize instead of polling the     */
                                    /* step's sequence word in mapped host memory                          */
};

typedef struct {
  int32_t device;          /* CUDA ordinal                                                          */
  int32_t test_statistic;  /* m3b_test_statistic; LikelihoodOptions:TestStatistic (Manager.cpp:98-127)*/
  int32_t update_w2;       /* LikelihoodOptions:UpdateW2 (Samples/SampleHandlerFD.cpp:64)            */
  int32_t tile_events;     /* events per tile row of the device layout: 0 (auto: 512 below 1.5M events, else 1024), 256, 512, 1024 */
  int32_t flags;           /* M3B_FLAG_*                                                            */
  int32_t reserved[11];
} m3b_config;

/* ---- lifetime --------------------------------------------------------------------------------
 * replaces SMonolithGPU::SMonolithGPU / InitGPU_* / Cleanup* (Splines/gpuSplineUtils.cu:80-190,
 * 520-557) and the device-side half of SampleHandlerFD's arrays (SampleHandlerFD.cpp:749-754).   */
M3B_API int m3b_create(const m3b_config* cfg, m3b_handle** out);
M3B_API void m3b_destroy(m3b_handle* h);
M3B_API const char* m3b_last_error(const m3b_handle* h);   /* h may be NULL: last global error      */
M3B_API int m3b_abi_version(void);
/* run everything on this cudaStream_t (e.g. the caller's current stream); default: own stream    */
M3B_API int m3b_set_stream(m3b_handle* h, void* cuda_stream);

/* ---- spline monolith -------------------------------------------------------------------------
 * replaces SMonolithGPU::InitGPU_SplineMonolith + CopyToGPU_SplineMonolith
 * (Splines/gpuSplineUtils.cu:103-168, 193-330).  Arrays are the reference's own
 * (SplineMonoStruct, Splines/SplineCommon.h:30-50; SMonolith members, SplineMonolith.h:96-137):
 *   coeff_x[n_params*max_knots]  knot x per parameter          n_pts[n_params] knots per parameter
 *                                                              (FastSplineInfo::nPts, 0 = none)
 * then, per chunk of events (any chunking; offsets relative to the chunk):
 *   nParamPerEvent[2n]     {count,start} of TSpline3 responses (start is recomputed from counts)
 *   paramNo_arr[tot_c]     parameter of each response          nKnots_arr[tot_c] first-knot offset
 *   coeff_many[total_knots*4]  AoS {y,b,c,d}
 *   nParamPerEvent_tf1[2n], paramNo_tf1[tot_l], coeff_tf1[tot_l*2] {a,b}   (TF1: a*x+b)
 * The library re-tiles into its own SoA layout on the device.                                      */
M3B_API int m3b_splines_begin(m3b_handle* h, int32_t n_params, int32_t max_knots,
                              const float* coeff_x, const int16_t* n_pts, int64_t n_events_total);
M3B_API int m3b_splines_append(m3b_handle* h, int64_t n_events,
                               const uint32_t* nParamPerEvent, const int16_t* paramNo_arr,
                               const uint64_t* nKnots_arr, uint64_t total_knots, const float* coeff_many,
                               const uint32_t* nParamPerEvent_tf1, const int16_t* paramNo_tf1,
                               const float* coeff_tf1);
M3B_API int m3b_splines_end(m3b_handle* h);
/* one-shot form with exactly the reference's types (unsigned int knot offsets) */
M3B_API int m3b_upload_spline_monolith(m3b_handle* h, int32_t n_params, int32_t max_knots,
                                       const float* coeff_x, const int16_t* n_pts, int64_t n_events,
                                       const uint32_t* nParamPerEvent, const int16_t* paramNo_arr,
                                       const uint32_t* nKnots_arr, uint32_t total_knots, const float* coeff_many,
                                       const uint32_t* nParamPerEvent_tf1, const int16_t* paramNo_tf1,
                                       const float* coeff_tf1);

/* ---- the monolith on disk, without ROOT -----------------------------------------------------------------------
 * The reference caches a built SMonolith in a ROOT file (SMonolith::PrepareSplineFile / LoadSplineFile,
 * Splines/SplineMonolith.cpp:454-614; FastSplineInfo directory, Splines/SplineBase.cpp:139-191).  The same arrays as ONE
 * flat little-endian file ("M3BMONO1": header with the scalars of the "Settings" tree, then named, 64-byte aligned
 * sections carrying the reference's member names -- layout in mach3_b200/csrc/m3b_file.cu):
 * m3b_write_monolith_file   the writer a MaCh3 maintainer calls next to PrepareSplineFile (adapters/M3BMonolithFile.h);
 *                           x_pts_f64 (FastSplineInfo::xPts as doubles, rows padded to max_knots) may be NULL; needs no GPU
 * m3b_monolith_file_info    the header scalars (needs no GPU)
 * m3b_upload_from_file      = LoadSplineFile + MoveToGPU: streams the file to the device in chunks of chunk_events events
 *                           (0 = 131072) so host memory stays bounded however large the monolith is
 * m3b_group_upload_from_file every member of a single-process group reads its own event range                        */
M3B_API int m3b_write_monolith_file(const char* path, int32_t n_params, int32_t max_knots, const float* coeff_x,
                                    const int16_t* n_pts, const double* x_pts_f64, int64_t n_events,
                                    const uint32_t* nParamPerEvent, const int16_t* paramNo_arr, const uint32_t* nKnots_arr,
                                    uint32_t total_knots, const float* coeff_many, const uint32_t* nParamPerEvent_tf1,
                                    const int16_t* paramNo_tf1, const float* coeff_tf1);
M3B_API int m3b_monolith_file_info(const char* path, int64_t* n_events, int32_t* n_params, int32_t* max_knots, uint64_t* total_knots);
M3B_API int m3b_upload_from_file(m3b_handle* h, const char* path, int64_t chunk_events);

/* FastSplineInfo::xPts in the default build's M3::float_t = double (Splines/SplineStructs.h:21-44).  The monolith
 * arrays only carry the knots as floats (coeff_x); SplineBase::FindSplineSegment (Splines/SplineBase.cpp:44-111)
 * compares the float-narrowed parameter against xPts, which hold the splines' double knots when SMonolith was
 * built in-process (Splines/SplineMonolith.cpp:393-404) and float-rounded ones when it was reloaded from a spline
 * file (Splines/SplineBase.cpp:166-190) or in a _LOW_MEMORY_STRUCTS_ build.  The two differ only when a proposed
 * value lands between a knot and its float rounding.  Default: coeff_x.  x_pts[n_params*max_knots] (rows padded,
 * only n_pts[p] entries of row p are read); NULL restores the default.  dx still uses coeff_x, as the reference. */
M3B_API int m3b_set_spline_knots_f64(m3b_handle* h, const double* x_pts);

/* ---- binned splines --------------------------------------------------------------------------
 * The other SplineBase implementation, BinnedSplineHandler (Splines/BinnedSplineHandler.h:110-135,
 * Evaluate/CalcSplineWeights .cpp:295-341), _LOW_MEMORY_STRUCTS_ build (M3::float_t = float).  Either this
 * or the event-by-event monolith above, per handle.  Arrays are the reference's own:
 *   knot_x[n_params*max_knots], n_pts[n_params]   FastSplineInfo::xPts / nPts per spline parameter
 *   uniquesplinevec_Monolith[n_slots]             parameter of every weightvec_Monolith slot
 *   coeffindexvec[n_slots]                        first knot of the slot's spline in the coefficient arrays
 *   uniquecoeffindices[n_unique]                  the non-flat slots (the only ones evaluated, .cpp:311)
 *   manycoeff_arr[n_coeff*4] {y,b,c,d}, xcoeff_arr[n_coeff]   per knot per spline (x is per spline, .cpp:329)
 * m3b_upload_event_binned_splines (after m3b_upload_events): the event's pointers into weightvec_Monolith
 *   (BinnedSplineHandler::retPointer, wired at Samples/SampleHandlerFD.cpp:1196-1242) as slot indices, in
 *   pointer order; they are multiplied after the oscillation weight and before the static weight.
 * m3b_read_binned_weights: lazy host mirror of weightvec_Monolith (1.0 for flat slots).                   */
M3B_API int m3b_upload_binned_splines(m3b_handle* h, int32_t n_params, int32_t max_knots, const float* knot_x,
                                      const int16_t* n_pts, int64_t n_slots, const int32_t* uniquesplinevec_Monolith,
                                      const int32_t* coeffindexvec, int64_t n_unique, const int32_t* uniquecoeffindices,
                                      int64_t n_coeff, const float* manycoeff_arr, const float* xcoeff_arr);
M3B_API int m3b_upload_event_binned_splines(m3b_handle* h, int64_t n_events, const uint32_t* n_per_event,
                                            const int32_t* spline_index);
M3B_API int m3b_read_binned_weights(m3b_handle* h, float* weightvec_Monolith /* [n_slots] */);
/* The same path in the reference's DEFAULT build (M3::float_t = double, Manager/Core.h:27-51 -- the build in which
 * BinnedSplineHandler is normally used): coefficients, knots, binned weights, oscillation and static weights and the
 * per-event product are double, fma instead of fmaf, the parameter value is read un-narrowed by the evaluation
 * (Splines/BinnedSplineHandler.cpp:327) but narrowed to float by FindSplineSegment (Splines/SplineBase.cpp:54).
 * m3b_upload_binned_splines_f64 switches the handle to that build; oscillation weights then come through
 * m3b_upload_osc_f64 (m3b_step's float osc_w must be NULL), static weights through m3b_upload_event_weights_f64
 * (or widened from m3b_upload_events' floats), mirrors through the *_f64 readers.                                 */
M3B_API int m3b_upload_binned_splines_f64(m3b_handle* h, int32_t n_params, int32_t max_knots, const double* knot_x,
                                          const int16_t* n_pts, int64_t n_slots, const int32_t* uniquesplinevec_Monolith,
                                          const int32_t* coeffindexvec, int64_t n_unique, const int32_t* uniquecoeffindices,
                                          int64_t n_coeff, const double* manycoeff_arr, const double* xcoeff_arr);
M3B_API int m3b_upload_event_weights_f64(m3b_handle* h, int64_t n_events, const double* static_w);
M3B_API int m3b_upload_osc_f64(m3b_handle* h, const double* osc_w, int64_t n);
M3B_API int m3b_read_binned_weights_f64(m3b_handle* h, double* weightvec_Monolith /* [n_slots] */);
M3B_API int m3b_read_event_weights_f64(m3b_handle* h, double* spline_w, double* total_w);

/* ---- binning, events, data -------------------------------------------------------------------
 * m3b_upload_binning: BinningHandler's uniform binning (Samples/BinningHandler.cpp:341-355,
 *   SampleBinningInfo Samples/SampleStructs.h:232-676): per sample n_dim[s] axes, nbins[s*4+d],
 *   edges concatenated sample-major, dim-major.  Global bin = sum bin_d*stride_d + GlobalOffset.
 * m3b_upload_events: what SampleHandlerFD::Initialise wires per event (SampleHandlerFD.cpp:169-202)
 *   sample_id[E]          EventInfo::NominalSample
 *   kin[d*E+e]            *EventInfo::KinVar[d], d < the largest dimensionality among the samples of the binning (rows a
 *                         lower-dimensional sample does not use are ignored; bins are found on the device with
 *                         FindGlobalBin semantics, BinningHandler.cpp:257-277 -> SampleStructs.h:577-613)
 *   norm_idx[e*npe+j]     index into the per-step norm value array (EventInfo::norm_pointers as
 *                         offsets from ParameterHandlerBase::_fPropVal), <0 = none
 *   osc_idx[E]            index into the per-step oscillation weight array (osc_w_pointer as an
 *                         offset; -1 = this event has none, e.g. &M3::Unity for NC events), NULL = event e reads
 *                         osc[e]; use_osc=0: no osc weight at all
 *   static_w[E]           product of the event's constant extra weights, NULL = none
 * Events must come in the same order as the spline monolith's events.                              */
M3B_API int m3b_upload_binning(m3b_handle* h, int32_t n_samples, const int32_t* n_dim,
                               const int32_t* nbins /*[n_samples*4]*/, const double* edges);
/* m3b_upload_binning_ex: like m3b_upload_binning, but sample s may use the reference's NON-UNIFORM binning
 *   (uniform[s] == 0; SampleBinningInfo::InitNonUniform, Samples/SampleStructs.h:468-528): its bins are boxes,
 *   nbins[s*4+0] = number of boxes, and its part of `edges` holds boxes*n_dim[s] {lo,hi} pairs (BinInfo::Extent).
 *   The library builds the same 10-per-dimension "mega bin" grid and box lists as InitialiseGridMapping
 *   (:394-466); an event is in the first listed box with lo < x <= hi in every dimension
 *   (BinInfo::IsEventInside :207-219; FindGlobalBin's non-uniform arm, Samples/BinningHandler.cpp:278-290).
 *   All non-uniform samples of one handle must have the same dimensionality.                               */
M3B_API int m3b_upload_binning_ex(m3b_handle* h, int32_t n_samples, const int32_t* n_dim, const int32_t* uniform,
                                  const int32_t* nbins /*[n_samples*4]*/, const double* edges);
M3B_API int m3b_upload_events(m3b_handle* h, int64_t n_events, const int32_t* sample_id, const double* kin,
                              int32_t n_norm_per_event, const int16_t* norm_idx, int32_t n_norm_values,
                              int32_t use_osc, const int32_t* osc_idx, int64_t n_osc_values,
                              const float* static_w);
/* m3b_update_kinematics: functional ("shift") parameters (Samples/SampleHandlerFD.cpp:545-564, ApplyShifts) call
 *   arbitrary std::functions per event, so they stay on the host; the caller hands over the shifted kinematic
 *   variables (same layout as m3b_upload_events' kin) and the events are re-binned on the device.  Needs
 *   M3B_FLAG_KEEP_KINEMATICS.  Asynchronous; takes effect for the following steps.                             */
M3B_API int m3b_update_kinematics(m3b_handle* h, const double* kin);
/* m3b_upload_selection: SampleHandlerFD::IsEventSelected (Samples/SampleHandlerFD.cpp:281-294), applied to every
 *   event before its weight is formed (FillArray :361, FillArray_MP :424): the per-sample lists of KinematicCut
 *   {ParamToCutOnIt, LowerBound, UpperBound} (Samples/SampleStructs.h:149-157; StoredSelection, filled at
 *   SampleHandlerFD.cpp:162 and copied into Selection at the top of every fill, :355/:393).  An event of sample s is
 *   dropped when, for any cut k of that sample, Val < lower[k] || Val >= upper[k] (so Val == lower passes, Val ==
 *   upper fails, NaN passes -- the reference's comparison).  Val = ReturnKinematicParameter(ParamToCutOnIt, event)
 *   is experiment code, so the caller evaluates it once per distinct cut variable and hands the table over:
 *     cut_sample[n_cuts], cut_var[n_cuts], lower[n_cuts], upper[n_cuts]      cuts in StoredSelection order
 *     values[v*n_events + e], v < n_vars                                      the v-th cut variable of event e
 *   cut_var[k] >= 0 indexes `values`; cut_var[k] = -1-d means "the event's d-th binning variable" (kin of
 *   m3b_upload_events; needs M3B_FLAG_KEEP_KINEMATICS) and follows m3b_update_kinematics automatically.
 *   After m3b_upload_events; replaces any earlier selection (n_cuts = 0 removes it).  Dropped events keep their
 *   FindGlobalBin result in m3b_read_event_bins but never reach the histogram; m3b_read_event_selected gives the mask.
 * m3b_update_selection_values: the cut variables after functional shifts (ApplyShifts runs before IsEventSelected,
 *   SampleHandlerFD.cpp:359-361): same layout as `values`; the selection is re-evaluated on the device.          */
M3B_API int m3b_upload_selection(m3b_handle* h, int32_t n_cuts, const int32_t* cut_sample, const int32_t* cut_var,
                                 const double* lower, const double* upper, int32_t n_vars, const double* values);
M3B_API int m3b_update_selection_values(m3b_handle* h, const double* values);
M3B_API int m3b_read_event_selected(m3b_handle* h, uint8_t* selected /* [n_events] 1 = passes every cut */);
/* m3b_upload_linear_shifts / m3b_set_shift_pars: functional ("shift") parameters ON THE DEVICE.
 *   SampleHandlerFD::ApplyShifts (Samples/SampleHandlerFD.cpp:545-564) runs, for every event and every step,
 *   ResetShifts(event) -> (*funcPtr)(valuePtr, event) for each FunctionalShifter of funcParsGrid[event], in order ->
 *   FinaliseShifts(event), before IsEventSelected and FindGlobalBin see the event.  The std::functions are experiment
 *   code; the family covered here is the linear one,  x_t += (*valuePtr) * c  with a per-event constant c (energy-scale,
 *   bias and resolution shifts are of this form: c = the energy deposit the parameter scales).  Per event, in
 *   funcParsGrid order: n_per_event[e] entries {shift_par (index into the per-step shift-parameter array), target, coef};
 *   target t < n_dims: the event's t-th binning variable; t >= n_dims: row t - n_dims of the selection's cut-variable
 *   table (m3b_upload_selection, which must come first if cut variables are shifted).  Arithmetic: double, one
 *   multiplication and one addition per entry, each rounded (what `x += par * c` compiles to without FMA contraction),
 *   starting from the nominal values every step: bit-identical shifted variables, hence bit-identical bins and cuts.
 *   Needs M3B_FLAG_KEEP_KINEMATICS.  m3b_set_shift_pars is asynchronous and takes effect for the following steps
 *   (bins and selection are recomputed on the device: no per-event data crosses PCIe).
 *   Shifts that are not of this form stay with the caller: m3b_update_kinematics / m3b_update_selection_values.   */
M3B_API int m3b_upload_linear_shifts(m3b_handle* h, int32_t n_shift_pars, int64_t n_events, const uint32_t* n_per_event,
                                     const int32_t* shift_par, const int32_t* target, const double* coef);
M3B_API int m3b_set_shift_pars(m3b_handle* h, const double* values /* [n_shift_pars] */);
/* SampleHandlerFD::AddData (Samples/SampleHandlerFD.cpp:955-1044), array form */
M3B_API int m3b_upload_data(m3b_handle* h, const double* data, int32_t n_bins);
/* oscillation weights computed elsewhere (NuOscillator) and already valid for the next steps    */
M3B_API int m3b_upload_osc(m3b_handle* h, const float* osc_w, int64_t n);
/* pin the caller's persistent oscillation-weight array so per-step copies are true DMA           */
M3B_API int m3b_register_host_buffer(m3b_handle* h, void* ptr, uint64_t bytes);
/* ... or, better, let the library allocate it: pinned + mapped host memory from the CUDA allocator (like the
 * reference's cudaMallocHost of cpu_total_weights, Splines/gpuSplineUtils.cu:139).  Measured on this pool: 50 GB/s
 * H2D against ~20 GB/s for registered malloc memory.  Freed by m3b_free_host or with the handle.              */
M3B_API int m3b_alloc_host(m3b_handle* h, uint64_t bytes, void** ptr);
M3B_API int m3b_free_host(m3b_handle* h, void* ptr);
M3B_API int m3b_set_test_statistic(m3b_handle* h, int32_t test_statistic);   /* SampleHandlerBase.h:185 */
/* switch run-time flags of m3b_config::flags on (set_mask) / off (clear_mask): M3B_FLAG_NO_BATCH_KERNEL,
 * M3B_FLAG_BATCH_KERNEL_V1, M3B_FLAG_NO_SPIN_LLH (the other flags fix allocations at creation and cannot change) */
M3B_API int m3b_set_flags(m3b_handle* h, int32_t set_mask, int32_t clear_mask);
M3B_API int m3b_reset_w2(m3b_handle* h);   /* FirstTimeW2 = true again                              */

/* ---- the step --------------------------------------------------------------------------------
 * m3b_step = SampleHandlerFD::Reweight (Samples/SampleHandlerFD.cpp:316-343):
 *   ResetHistograms -> SplineBase::FindSplineSegment (Splines/SplineBase.cpp:44-109, on the host,
 *   with the reference's cached-segment history) -> CalcSplineWeights + CalcTotalEventWeight
 *   (Splines/SplineMonolith.cpp:727-830) -> FillArray_MP (SampleHandlerFD.cpp:390-448) -> the
 *   GetLikelihood reduction (SampleHandlerFD.cpp:1284-1300), all in one device pass.
 *   spline_pars[n_params]   the doubles behind FastSplineInfo::splineParsPointer
 *   norm_pars[n_norm_values] the doubles behind EventInfo::norm_pointers
 *   osc_w                   host array of this step's oscillation weights (copied H2D inside the
 *                           call), or NULL to keep the weights already on the device
 * m3b_step_segments = SMonolithGPU::RunGPU_SplineMonolith's contract
 *   (Splines/gpuSplineUtils.cu:444-512): the caller already ran FindSplineSegment and passes
 *   ParamValues (float) and SplineSegments (short).
 * m3b_llh = SampleHandlerFD::GetLikelihood / GetSampleLikelihood: blocks, returns -lnL (NOT -2lnL).  */
M3B_API int m3b_step(m3b_handle* h, const double* spline_pars, const double* norm_pars, const float* osc_w);
M3B_API int m3b_step_segments(m3b_handle* h, const float* param_values, const int16_t* segments,
                              const double* norm_pars, const float* osc_w);
M3B_API int m3b_llh(m3b_handle* h, double* total, double* per_sample /* [n_samples] or NULL */);
/* m3b_step_batch: n_sets proposals against the same events (parallel chains, DelayedMR2T2 stages
 *   Fitters/DelayedMR2T2.cpp:110-157, RunLLHScan Fitters/FitterBase.cpp:742-798).  spline_pars[n_sets*n_params],
 *   norm_pars[n_sets*n_norm_values] row-major; sets are evaluated in order with the reference's sequential
 *   semantics (cached segments, W2 freeze); ONE host synchronisation; llh_total[n_sets] (-lnL each),
 *   llh_per_sample[n_sets*n_samples] or NULL.  osc_w (or NULL) applies to the whole batch.                 */
M3B_API int m3b_step_batch(m3b_handle* h, int32_t n_sets, const double* spline_pars, const double* norm_pars,
                           const float* osc_w, double* llh_total, double* llh_per_sample);
/* m3b_step_batch_hist: the same, and every set's MC histogram comes back too: mc[n_sets*n_bins] (what
 *   PredictiveThrower writes per toy after samples[i]->Reweight(), Fitters/PredictiveThrower.cpp:507-563).      */
M3B_API int m3b_step_batch_hist(m3b_handle* h, int32_t n_sets, const double* spline_pars, const double* norm_pars,
                                const float* osc_w, double* llh_total, double* llh_per_sample /* or NULL */,
                                double* mc);
/* m3b_eval_weights = SMonolithGPU::RunGPU_SplineMonolith exactly (Splines/gpuSplineUtils.cu:444-512):
 *   evaluate all responses, multiply per event, and enqueue the copy of the per-event totals into the
 *   caller's host array (cpu_total_weights, pinned by InitGPU_SplineMonolith :139) -- asynchronous until
 *   m3b_synchronize (= SynchroniseSplines, :515-518).  No fill, no likelihood; needs no sample handler
 *   (if no events were uploaded the library wires a one-bin sample around the monolith's events).
 *   Handle must carry M3B_FLAG_KEEP_EVENT_WEIGHTS.  This is what adapters/SMonolithGPU_m3b200.cu calls.  */
M3B_API int m3b_eval_weights(m3b_handle* h, const float* param_values, const int16_t* segments,
                             float* host_total_weights);
/* SplineBase::FindSplineSegment alone (host, history-dependent); outputs are optional            */
M3B_API int m3b_find_segments(m3b_handle* h, const double* spline_pars, int16_t* segments, float* param_values);
M3B_API int m3b_synchronize(m3b_handle* h);   /* SplineBase::SynchroniseMemTransfer                 */

/* ---- read-back (lazy host mirrors) -----------------------------------------------------------
 * m3b_read_hist           SampleHandlerFD_array / _array_w2 (SampleHandlerFD.h:337-341)
 * m3b_read_event_weights  spline_w[e] = *SMonolith::retPointer(e) (Splines/SplineMonolith.h:40);
 *                         total_w[e] = CalcWeightTotal (SampleHandlerFD.cpp:568-594); either may be NULL
 * m3b_read_event_bins     FindGlobalBin per event (-1 = under/overflow)                              */
M3B_API int m3b_read_hist(m3b_handle* h, double* mc, double* w2);
M3B_API int m3b_read_event_weights(m3b_handle* h, float* spline_w, float* total_w);
M3B_API int m3b_read_event_bins(m3b_handle* h, int32_t* bins);

/* ---- multi-GPU: one process per GPU, events sharded, partial histograms summed ---------------
 * (new capability: the reference supports one GPU, Manager/gpuUtils.cu:71.)
 * m3b_step_fill           like m3b_step but stops after the partial histogram (no LLH)
 * m3b_hist_device_ptr     device address of {mc[n_bins], w2[n_bins]} (contiguous, 2*n_bins doubles)
 *                         so the caller can all-reduce it in place (NCCL through torch.distributed)
 * m3b_llh_from_hist       enqueue the LLH reduction over the (now global) histogram
 * m3b_peer_*, m3b_step_peer  the library's own exchange over NVLink peer memory, fused with the likelihood: every
 *                         rank's partial histogram lives in a buffer exported through CUDA IPC (m3b_peer_export gives
 *                         the 64-byte handle, m3b_peer_import maps a peer's); the fill kernel publishes an epoch flag,
 *                         the exchange+likelihood kernel on every rank pulls all partials over NVLink, sums them in
 *                         rank order (bit-identical totals on all ranks) and reduces -lnL in the same launch.  The
 *                         wait for the peers' flags is bounded: a dead peer gives M3B_ERR_PEER, never a hang.          */
M3B_API int m3b_step_fill(m3b_handle* h, const double* spline_pars, const double* norm_pars, const float* osc_w);
M3B_API int m3b_hist_device_ptr(m3b_handle* h, void** dev_ptr, int32_t* n_bins, int32_t* w2_live);
M3B_API int m3b_llh_from_hist(m3b_handle* h);
M3B_API int m3b_peer_export(m3b_handle* h, int32_t rank, int32_t world, void* ipc_handle_64B);
M3B_API int m3b_peer_import(m3b_handle* h, int32_t peer_rank, const void* ipc_handle_64B);
M3B_API int m3b_step_peer(m3b_handle* h, const double* spline_pars, const double* norm_pars, const float* osc_w);

/* ---- multi-GPU from ONE process, ONE calling thread ---------------------------------------------------------------
 * The reference's fitters are single-process, single-threaded callers (Fitters/MR2T2.cpp:62-74: per sample handler
 * Reweight(), then GetLikelihood(); Fitters/FitterBase.cpp:461-520).  A group keeps that call surface and spreads ONE
 * sample handler over n devices of the box: member i owns a contiguous, tile-aligned shard of the events on devices[i]
 * (the same device may be listed more than once), the library launches the n fills from its own per-device worker
 * threads, and member 0 -- the lead -- sums the partial histograms and reduces -lnL in one launch:
 *   M3B_EXCHANGE_PEER  the lead pulls the members' partial histograms through peer-to-peer loads (cudaDeviceEnablePeerAccess
 *                      pointers; no IPC, no second process) behind their epoch flags -- exchange fused with the likelihood;
 *   M3B_EXCHANGE_NCCL  every member all-reduces its histogram in place with ONE ncclAllReduce over NVLink (libnccl.so.2
 *                      is loaded at run time, communicators from ncclCommInitAll), the lead then reduces -lnL.
 * Set-up: m3b_group_create -> uploads (either per member through m3b_group_member() with the ordinary m3b_upload_* calls
 * on the member's shard [m3b_group_shard], or for the whole workload at once through the m3b_group_upload_* calls below,
 * which slice the reference's arrays by event range) -> m3b_group_connect.  Per step: m3b_group_step (asynchronous) +
 * m3b_group_llh (blocks).  Not thread-safe per group; one group = one SampleHandlerFD.                              */
typedef struct m3b_group m3b_group;
enum { M3B_EXCHANGE_PEER = 0, M3B_EXCHANGE_NCCL = 1 };
M3B_API int m3b_group_create(const m3b_config* cfg /* .device ignored */, const int32_t* devices, int32_t n_devices, m3b_group** out);
M3B_API void m3b_group_destroy(m3b_group* g);
M3B_API const char* m3b_group_last_error(const m3b_group* g);
M3B_API int32_t m3b_group_size(const m3b_group* g);
M3B_API m3b_handle* m3b_group_member(m3b_group* g, int32_t i);
/* events [*e0, *e1) of member i when n_events are spread over the group (equal contiguous shards, multiples of 1024) */
M3B_API int m3b_group_shard(const m3b_group* g, int64_t n_events, int32_t i, int64_t* e0, int64_t* e1);
/* whole-workload uploads: same arguments as the per-handle calls, sliced by the members' event ranges */
M3B_API int m3b_group_upload_spline_monolith(m3b_group* g, int32_t n_params, int32_t max_knots, const float* coeff_x,
                                             const int16_t* n_pts, int64_t n_events, const uint32_t* nParamPerEvent,
                                             const int16_t* paramNo_arr, const uint32_t* nKnots_arr, uint32_t total_knots,
                                             const float* coeff_many, const uint32_t* nParamPerEvent_tf1,
                                             const int16_t* paramNo_tf1, const float* coeff_tf1);
M3B_API int m3b_group_upload_from_file(m3b_group* g, const char* path, int64_t chunk_events);
M3B_API int m3b_group_upload_binning_ex(m3b_group* g, int32_t n_samples, const int32_t* n_dim, const int32_t* uniform,
                                        const int32_t* nbins, const double* edges);
M3B_API int m3b_group_upload_events(m3b_group* g, int64_t n_events, const int32_t* sample_id, const double* kin,
                                    int32_t n_norm_per_event, const int16_t* norm_idx, int32_t n_norm_values,
                                    int32_t use_osc, const int32_t* osc_idx, int64_t n_osc_values, const float* static_w);
M3B_API int m3b_group_upload_selection(m3b_group* g, int32_t n_cuts, const int32_t* cut_sample, const int32_t* cut_var,
                                       const double* lower, const double* upper, int32_t n_vars, const double* values);
M3B_API int m3b_group_upload_linear_shifts(m3b_group* g, int32_t n_shift_pars, int64_t n_events, const uint32_t* n_per_event,
                                           const int32_t* shift_par, const int32_t* target, const double* coef);
M3B_API int m3b_group_set_shift_pars(m3b_group* g, const double* values);
M3B_API int m3b_group_upload_data(m3b_group* g, const double* data, int32_t n_bins);
M3B_API int m3b_group_upload_osc(m3b_group* g, const float* osc_w, int64_t n);
/* wires the exchange (after every member has its binning and events); exchange = M3B_EXCHANGE_* */
M3B_API int m3b_group_connect(m3b_group* g, int32_t exchange);
/* pinned + mapped host memory every device of the group can stream from (the oscillator's weight array) */
M3B_API int m3b_group_alloc_host(m3b_group* g, uint64_t bytes, void** ptr);
/* = SampleHandlerFD::Reweight on all shards.  osc_w: NULL, or this step's oscillation weights in host memory -- one
 * array over ALL events in event order when the events were uploaded without osc_idx (member i reads its own range),
 * else the shared array every osc_idx points into.  Asynchronous; the parameter arrays are copied before it returns. */
M3B_API int m3b_group_step(m3b_group* g, const double* spline_pars, const double* norm_pars, const float* osc_w);
M3B_API int m3b_group_llh(m3b_group* g, double* total, double* per_sample);
M3B_API int m3b_group_read_hist(m3b_group* g, double* mc, double* w2);
M3B_API int m3b_group_synchronize(m3b_group* g);

/* ---- introspection ----------------------------------------------------------------------------- */
typedef struct {
  int64_t n_events, n_tiles;
  int32_t n_params, n_bins, n_samples, n_signatures;
  int32_t tile_events, grid_blocks, smem_bytes, hist_in_smem;
  uint64_t device_bytes;          /* HBM held by the handle                                       */
  uint64_t active_bytes_per_step; /* coefficient + event-table bytes one step actually loads       */
  uint64_t steps, kernel_launches;
  int32_t kernel_variant;         /* -1: streaming TMA kernel (default); 0..5: register-streaming variants;  */
                                  /* -2: binned-spline kernels                                               */
  int32_t tma_stages;             /* 32 KB shared-memory stages of the coefficient ring                    */
} m3b_info;
M3B_API int m3b_get_info(m3b_handle* h, m3b_info* out);
/* CUDA-event timing of the fill kernel alone, on the handle's stream (the reference has only a
 * TStopwatch around the whole step, Fitters/MCMCBase.cpp:93).  m3b_kernel_time synchronises and
 * returns the summed duration and the number of launches timed since the last call.              */
M3B_API int m3b_set_timing(m3b_handle* h, int32_t enabled);
M3B_API int m3b_kernel_time(m3b_handle* h, double* total_ms, int64_t* n_launches);
/* per-block timeline of the fill kernel (globaltimer ns; 8 u64 per block: start, tables staged, first
 * stage consumed, producer out of work, consumers done, histogram flushed, block end, units done).
 * Call once with out == NULL to switch tracing on, then after a step with out = u64[grid*8].       */
M3B_API int m3b_block_trace(m3b_handle* h, uint64_t* out, int32_t* grid);

#ifdef __cplusplus
}
#endif
#endif
