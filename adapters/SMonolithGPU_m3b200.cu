// SMonolithGPU_m3b200.cu -- drop-in replacement for MaCh3's Splines/gpuSplineUtils.cu.
//
// The reference isolates CUDA from host code behind ONE class, SMonolithGPU
// (Splines/gpuSplineUtils.cuh:63-218); SMonolith is its only caller (Splines/SplineMonolith.cpp:254-313
// MoveToGPU, :695-708 Evaluate, :619-637 dtor) and SynchroniseSplines() its only fence (:851-856).
// This file defines exactly those member functions -- same names, arguments and ownership rules -- on
// top of libm3b200's C ABI (include/m3b200.h), so a MaCh3 build picks the B200 path by compiling this
// file INSTEAD of Splines/gpuSplineUtils.cu (one line in Splines/CMakeLists.txt).
// The class declaration is the reference's own, unmodified header: nothing of MaCh3 is copied here.
//
// Differences a maintainer should know about:
//   * no compile-time NSplines_GPU limit (gpuSplineUtils.cu:211-216): any number of parameters;
//   * every CUDA call is checked and reported (the reference's CudaCheckError is a no-op in release);
//   * the knot count of a parameter is taken from its first response; a response whose knot count
//     differs is an error (the reference silently assumes identical knots, SplineMonolith.cpp:102-104);
//   * this seam still copies NEvents x 4 B to the host every step, as the reference does
//     (gpuSplineUtils.cu:505).  The fused path (adapters/SampleHandlerB200.h) does not.
#include "Splines/gpuSplineUtils.cuh"   // the reference's header, found through -I<MaCh3 root>
#include "m3b200.h"

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace {
struct State {
  m3b_handle* h = nullptr;
  int n_events = 0;
};
// SMonolithGPU's data members are fixed by the reference header, so the handle lives beside the
// object.  (InitGPU_Segments/Vals are called on a null object, SplineMonolith.cpp:82-83: they must
// not touch `this`, and do not.)
std::mutex g_mu;
std::unordered_map<const SMonolithGPU*, State> g_state;
int g_sync_device = 0;

State& state_of(const SMonolithGPU* self) {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_state[self];
}
[[noreturn]] void die(const char* where, m3b_handle* h) {
  // MaCh3 convention at this seam: report, then terminate (Manager/gpuUtils.cu:11-35)
  std::fprintf(stderr, "SMonolithGPU[m3b200] %s: %s\n", where, m3b_last_error(h));
  std::abort();
}
void cuda_or_die(cudaError_t e, const char* where) {
  if (e != cudaSuccess) { std::fprintf(stderr, "SMonolithGPU[m3b200] %s: %s\n", where, cudaGetErrorString(e)); std::abort(); }
}
}  // namespace

// Splines/gpuSplineUtils.cu:515-518
__host__ void SynchroniseSplines() { cuda_or_die(cudaDeviceSynchronize(), "SynchroniseSplines"); }

SMonolithGPU::SMonolithGPU() {
  h_n_params = -1; h_n_events = -1;
  gpu_nParamPerEvent = nullptr; gpu_nParamPerEvent_TF1 = nullptr; gpu_coeff_x = nullptr; gpu_coeff_many = nullptr;
  gpu_nKnots_arr = nullptr; gpu_paramNo_arr = nullptr; gpu_coeff_TF1_many = nullptr; gpu_nPoints_arr = nullptr;
  gpu_paramNo_TF1_arr = nullptr; gpu_total_weights = nullptr; gpu_weights = nullptr; gpu_weights_tf1 = nullptr;
}

SMonolithGPU::~SMonolithGPU() {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_state.find(this);
  if (it != g_state.end()) { if (it->second.h) m3b_destroy(it->second.h); g_state.erase(it); }
}

// Splines/gpuSplineUtils.cu:103-168 -- device allocations happen inside libm3b200 at copy time; what
// the caller needs back from this call is the pinned host array for the per-event total weights.
__host__ void SMonolithGPU::InitGPU_SplineMonolith(float** cpu_total_weights, int n_events, unsigned int /*total_nknots*/,
                                                   unsigned int /*n_splines*/, unsigned int /*n_tf1*/, int /*Eve_size*/) {
  State& st = state_of(this);
  if (!st.h) {
    m3b_config cfg{};
    cuda_or_die(cudaGetDevice(&cfg.device), "cudaGetDevice");
    g_sync_device = cfg.device;
    cfg.flags = M3B_FLAG_KEEP_EVENT_WEIGHTS;
    if (m3b_create(&cfg, &st.h) != M3B_OK) die("m3b_create", nullptr);
  }
  st.n_events = n_events;
  h_n_events = n_events;
  cuda_or_die(cudaMallocHost(reinterpret_cast<void**>(cpu_total_weights), sizeof(float) * static_cast<size_t>(n_events)),
              "cudaMallocHost(cpu_total_weights)");   // :139
}

// Splines/gpuSplineUtils.cu:193-330
__host__ void SMonolithGPU::CopyToGPU_SplineMonolith(SplineMonoStruct* cpu_spline_handler, std::vector<float> cpu_many_array_TF1,
                                                     std::vector<short int> cpu_paramNo_arr_TF1, int n_events,
                                                     std::vector<unsigned int> cpu_nParamPerEvent,
                                                     std::vector<unsigned int> cpu_nParamPerEvent_TF1, int n_params,
                                                     unsigned int n_splines, short int spline_size, unsigned int total_nknots,
                                                     unsigned int /*n_tf1*/) {
  State& st = state_of(this);
  if (!st.h) die("CopyToGPU_SplineMonolith before InitGPU_SplineMonolith", nullptr);
  h_n_params = n_params;
  // knots per parameter = knot count of the parameter's first response (FastSplineInfo::nPts,
  // Splines/SplineMonolith.cpp:386-400); parameters that only carry TF1s have none.
  std::vector<int16_t> n_pts(static_cast<size_t>(n_params), 0);
  const std::vector<unsigned int>& kn = cpu_spline_handler->nKnots_arr;
  const std::vector<short int>& pn = cpu_spline_handler->paramNo_arr;
  for (unsigned int s = 0; s < n_splines; ++s) {
    const int p = pn[s];
    if (n_pts[p] == 0) n_pts[p] = static_cast<int16_t>((s + 1 < n_splines ? kn[s + 1] : total_nknots) - kn[s]);
  }
  static_assert(sizeof(short int) == sizeof(int16_t) && sizeof(unsigned int) == sizeof(uint32_t), "reference types");
  const int rc = m3b_upload_spline_monolith(
      st.h, n_params, spline_size, cpu_spline_handler->coeff_x.data(), n_pts.data(), n_events, cpu_nParamPerEvent.data(),
      reinterpret_cast<const int16_t*>(pn.data()), kn.data(), total_nknots, cpu_spline_handler->coeff_many.data(),
      cpu_nParamPerEvent_TF1.data(), reinterpret_cast<const int16_t*>(cpu_paramNo_arr_TF1.data()), cpu_many_array_TF1.data());
  if (rc != M3B_OK) die("m3b_upload_spline_monolith", st.h);
}

// Splines/gpuSplineUtils.cu:171-190 -- pinned staging for the per-step inputs.  Called on a null
// object by the reference (SplineMonolith.cpp:82-83): no member access.  The parameter count is not
// known yet at that point in the reference either, which sizes these by its compile-time
// NSplines_GPU; 2048 is libm3b200's own bound (kMaxParams).
__host__ void SMonolithGPU::InitGPU_Segments(short int** segment) {
  cuda_or_die(cudaMallocHost(reinterpret_cast<void**>(segment), 2048 * sizeof(short int)), "cudaMallocHost(segment)");
}
__host__ void SMonolithGPU::InitGPU_Vals(float** vals) {
  cuda_or_die(cudaMallocHost(reinterpret_cast<void**>(vals), 2048 * sizeof(float)), "cudaMallocHost(vals)");
}

// Splines/gpuSplineUtils.cu:444-512: asynchronous; cpu_total_weights is valid after SynchroniseSplines()
__host__ void SMonolithGPU::RunGPU_SplineMonolith(float* cpu_total_weights, float* vals, short int* segment,
                                                  const unsigned int /*h_n_splines*/, const unsigned int /*h_n_tf1*/) {
  State& st = state_of(this);
  if (m3b_eval_weights(st.h, vals, reinterpret_cast<const int16_t*>(segment), cpu_total_weights) != M3B_OK)
    die("m3b_eval_weights", st.h);
}

// Splines/gpuSplineUtils.cu:520-557
__host__ void SMonolithGPU::CleanupGPU_SplineMonolith(float* cpu_total_weights) {
  State& st = state_of(this);
  if (st.h) { m3b_destroy(st.h); st.h = nullptr; }
  if (cpu_total_weights) cudaFreeHost(cpu_total_weights);
}
__host__ void SMonolithGPU::CleanupGPU_Segments(short int* segment, float* vals) {
  if (segment) cudaFreeHost(segment);
  if (vals) cudaFreeHost(vals);
}
