/*
 * m3b_synth.h -- deterministic synthetic workload generator.
 *
 * Produces, for any event range [e0,e1), exactly the arrays the reference holds after
 * SMonolith::PrepareForGPU (Splines/SplineMonolith.cpp:53-250, layout in
 * Splines/SplineCommon.h:30-50) and after SampleHandlerFD::Initialise
 * (Samples/SampleHandlerFD.cpp:169-202), so that the CPU oracle and the B200 library
 * consume the same bytes.  Every value is a pure function of (seed, event, param, ...)
 * through a counter-based hash, so chunks can be generated independently and in parallel.
 *
 * This is workload generation, not the oracle and not the product: plain C, host only.
 */
#ifndef M3B_SYNTH_H
#define M3B_SYNTH_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M3S_MAX_SAMPLES 16

typedef struct {
  uint64_t seed;
  int64_t  n_events;
  int32_t  n_params;          /* P = cubic + linear spline parameters                       */
  int32_t  n_linear;          /* how many of the P are TF1 (linear) responses                */
  int32_t  n_knots;           /* K knots of every cubic response                            */
  int32_t  n_modes;           /* interaction-mode classes used for structured sparsity      */
  float    density;           /* P(mode,param has a response); 1.0 = dense                  */
  int32_t  n_samples;
  int32_t  n_dims;            /* 1 or 2 kinematic dimensions                                */
  int32_t  nbins_x, nbins_y;  /* per-sample binning                                         */
  int32_t  n_norm_params;     /* normalisation parameters                                   */
  int32_t  n_norm_per_event;  /* each event is bound to this many of them                   */
  int64_t  sample_start[M3S_MAX_SAMPLES + 1]; /* event index where each sample starts       */
  int32_t  mode_block;        /* events come in runs of this many sharing one mode (>=1)    */
  int32_t  reserved;
} m3s_config;

/* per-parameter layout: type[p] (0 = TSpline3, 1 = TF1), n_pts[p], coeff_x[p*K+j]          */
void m3s_param_layout(const m3s_config* c, int8_t* type, int16_t* n_pts, float* coeff_x);

/* number of cubic / linear responses of each event in [e0,e1); returns totals via out ptrs  */
void m3s_count(const m3s_config* c, int64_t e0, int64_t e1,
               uint32_t* n_cubic, uint32_t* n_linear, uint64_t* tot_cubic, uint64_t* tot_linear);

/* reference monolith arrays for events [e0,e1).  {count,start} pairs and knot offsets are
 * relative to the chunk (start of chunk = 0).  knot_off is 64-bit (the reference's is u32,
 * Splines/SplineMonolith.h:104-114; SURVEY §7 "32-bit indexing limit").                     */
void m3s_fill_splines(const m3s_config* c, int64_t e0, int64_t e1,
                      uint32_t* nParamPerEvent,     /* [2n] {count,start}  cubic            */
                      int16_t*  paramNo_arr,        /* [tot_cubic]                          */
                      uint64_t* knot_off,           /* [tot_cubic] first knot of spline     */
                      float*    coeff_many,         /* [tot_cubic*K*4] AoS {y,b,c,d}        */
                      uint32_t* nParamPerEvent_tf1, /* [2n] {count,start}  linear           */
                      int16_t*  paramNo_tf1,        /* [tot_linear]                         */
                      float*    coeff_tf1);         /* [tot_linear*2] {a,b}                 */

/* event table for [e0,e1): sample id, kinematic variables (dim-major: kin[d*n+i]), norm
 * bindings (n_norm_per_event indices per event), per-event static weight                    */
void m3s_fill_events(const m3s_config* c, int64_t e0, int64_t e1,
                     int32_t* sample_id, double* kin, int16_t* norm_idx, float* static_w);

/* osc-weight input array for events [e0,e1) at a given step (U(0,1), a few exact zeros)      */
void m3s_fill_osc(const m3s_config* c, int64_t e0, int64_t e1, int64_t step, float* osc_w);

/* bin edges of one dimension of one sample; n_edges = nbins+1                               */
void m3s_bin_edges(const m3s_config* c, int sample, int dim, double* edges);

/* proposal for a step: spline parameter values (double, sigma units) and norm values.
 * step < 0 requests special parity proposals: -1 all zero (nominal, on a knot), -2 every
 * parameter exactly on a knot, -3 below first knot, -4 above last knot.                     */
void m3s_proposal(const m3s_config* c, int64_t step, double* spline_pars, double* norm_pars);

#ifdef __cplusplus
}
#endif
#endif
