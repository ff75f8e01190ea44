// harness.cu -- drives the REFERENCE's own CUDA spline kernels (class SMonolithGPU,
// /root/reference/Splines/gpuSplineUtils.cu) exactly the way SMonolith does
// (Splines/SplineMonolith.cpp:254-313 MoveToGPU, :695-708 Evaluate), behind a C ABI for the tests.
//
// TEST INFRASTRUCTURE (part of oracle/): the reference sources are compiled where they lie under
// /root/reference by oracle/ref_gpu/Makefile; nothing of the reference is copied into this repo.
// It gives (a) a second oracle for the per-event spline weights -- the reference itself, run on the
// B200 -- used to pin oracle/m3_oracle.c, and (b) the incumbent-GPU timing (including its
// mandatory E x 4 B device->host copy per step, gpuSplineUtils.cu:505).
#include "Splines/gpuSplineUtils.cuh"
#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

struct RefMono {
  SMonolithGPU* gpu = nullptr;
  SplineMonoStruct* cpu = nullptr;
  float* cpu_total_weights = nullptr;
  short* segments = nullptr;
  float* vals = nullptr;
  int n_params = 0;
  unsigned n_events = 0, n_splines = 0, n_tf1 = 0;
};

extern "C" {

#ifdef M3REF_ANY_NPARAMS   /* harness linked against adapters/SMonolithGPU_m3b200.cu: no compile-time limit */
__attribute__((visibility("default"))) int m3ref_compiled_nparams(void) { return -1; }
#else
__attribute__((visibility("default"))) int m3ref_compiled_nparams(void) { return NSplines_GPU; }
#endif

__attribute__((visibility("default"))) void* m3ref_create(
    int n_params, int max_knots, const float* coeff_x, unsigned n_events, const unsigned* nParamPerEvent,
    const short* paramNo_arr, const unsigned* nKnots_arr, unsigned total_knots, const float* coeff_many,
    const unsigned* nParamPerEvent_tf1, const short* paramNo_tf1, const float* coeff_tf1) {
#ifndef M3REF_ANY_NPARAMS
  if (n_params != NSplines_GPU) return nullptr;      // the reference would `throw;` (gpuSplineUtils.cu:211-216)
#endif
  RefMono* r = new RefMono();
  r->n_params = n_params; r->n_events = n_events;
  unsigned ns = 0, nt = 0;
  for (unsigned e = 0; e < n_events; ++e) { ns += nParamPerEvent[2 * e]; nt += nParamPerEvent_tf1[2 * e]; }
  r->n_splines = ns; r->n_tf1 = nt;
  r->cpu = new SplineMonoStruct();
  r->cpu->coeff_x.assign(coeff_x, coeff_x + size_t(n_params) * max_knots);
  r->cpu->coeff_many.assign(coeff_many, coeff_many + size_t(total_knots) * 4);
  r->cpu->nKnots_arr.assign(nKnots_arr, nKnots_arr + ns);
  r->cpu->paramNo_arr.assign(paramNo_arr, paramNo_arr + ns);
  std::vector<float> tf1(coeff_tf1, coeff_tf1 + size_t(nt) * 2);
  std::vector<short> ptf1(paramNo_tf1, paramNo_tf1 + nt);
  std::vector<unsigned> npe(nParamPerEvent, nParamPerEvent + size_t(n_events) * 2);
  std::vector<unsigned> npe1(nParamPerEvent_tf1, nParamPerEvent_tf1 + size_t(n_events) * 2);
  r->gpu = new SMonolithGPU();
  r->gpu->InitGPU_Segments(&r->segments);
  r->gpu->InitGPU_Vals(&r->vals);
  r->gpu->InitGPU_SplineMonolith(&r->cpu_total_weights, int(n_events), total_knots, ns, nt, max_knots * n_params);
  r->gpu->CopyToGPU_SplineMonolith(r->cpu, tf1, ptf1, int(n_events), npe, npe1, n_params, ns, short(max_knots),
                                   total_knots, nt);
  cudaDeviceSynchronize();
  if (cudaGetLastError() != cudaSuccess) return nullptr;
  return r;
}

// SMonolith::Evaluate (CUDA build) after FindSplineSegment + SynchroniseMemTransfer
__attribute__((visibility("default"))) int m3ref_run(void* h, const float* vals, const short* segments) {
  RefMono* r = static_cast<RefMono*>(h);
  memcpy(r->vals, vals, sizeof(float) * r->n_params);
  memcpy(r->segments, segments, sizeof(short) * r->n_params);
  r->gpu->RunGPU_SplineMonolith(r->cpu_total_weights, r->vals, r->segments, r->n_splines, r->n_tf1);
  SynchroniseSplines();
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

__attribute__((visibility("default"))) const float* m3ref_total_weights(void* h) {
  return static_cast<RefMono*>(h)->cpu_total_weights;
}

// wall-clock per step of Evaluate + SynchroniseMemTransfer (what SampleHandlerFD::Reweight waits for)
__attribute__((visibility("default"))) double m3ref_time_ms(void* h, const float* vals, const short* segments, int laps) {
  RefMono* r = static_cast<RefMono*>(h);
  m3ref_run(h, vals, segments);
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < laps; ++i) {
    r->gpu->RunGPU_SplineMonolith(r->cpu_total_weights, r->vals, r->segments, r->n_splines, r->n_tf1);
    SynchroniseSplines();
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double, std::milli>(t1 - t0).count() / laps;
}

__attribute__((visibility("default"))) void m3ref_destroy(void* h) {
  RefMono* r = static_cast<RefMono*>(h);
  if (!r) return;
  r->gpu->CleanupGPU_SplineMonolith(r->cpu_total_weights);
  r->gpu->CleanupGPU_Segments(r->segments, r->vals);
  delete r->gpu; delete r->cpu; delete r;
}

}  // extern "C"
