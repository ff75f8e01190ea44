/*
 * m3b_synth.c -- deterministic synthetic workload generator (see m3b_synth.h).
 * Host-only plain C; OpenMP over events.  Not the oracle, not the product path.
 */
#include "m3b_synth.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>

#define M3S_MAX_KNOTS 64

enum { TAG_MODE = 1, TAG_ACT, TAG_CUB_A, TAG_CUB_B, TAG_LIN_A, TAG_KIN_X, TAG_KIN_Y,
       TAG_NORM, TAG_STATIC, TAG_OSC, TAG_PROP, TAG_PROPN };

static inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline uint64_t h4(uint64_t seed, uint64_t tag, uint64_t a, uint64_t b) {
  uint64_t h = mix64(seed ^ (tag * 0xD6E8FEB86659FD93ull));
  h = mix64(h ^ a);
  h = mix64(h ^ (b * 0xA24BAED4963EE407ull + 0x9FB21C651E98DF25ull));
  return h;
}
static inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }
/* Box-Muller from two hashes */
static inline double gaus(uint64_t h) {
  double u1 = u01(h); if (u1 < 1e-300) u1 = 1e-300;
  double u2 = u01(mix64(h));
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

static inline int param_is_linear(const m3s_config* c, int p) {
  /* spread the n_linear TF1 parameters evenly among the P parameters */
  int64_t P = c->n_params, L = c->n_linear;
  return ((int64_t)(p + 1) * L / P) > ((int64_t)p * L / P);
}
static inline int event_mode(const m3s_config* c, int64_t e) {
  const int64_t blk = c->mode_block > 1 ? c->mode_block : 1;
  return c->n_modes <= 1 ? 0 : (int)(h4(c->seed, TAG_MODE, (uint64_t)(e / blk), 0) % (uint64_t)c->n_modes);
}
static inline int has_response(const m3s_config* c, int p, int mode) {
  if (c->density >= 1.0f) return 1;
  return u01(h4(c->seed, TAG_ACT, (uint64_t)p, (uint64_t)mode)) < (double)c->density;
}
static inline double knot_x(const m3s_config* c, int k) {
  return -3.0 + 6.0 * (double)k / (double)(c->n_knots - 1);
}

void m3s_param_layout(const m3s_config* c, int8_t* type, int16_t* n_pts, float* coeff_x) {
  const int K = c->n_knots;
  for (int p = 0; p < c->n_params; ++p) {
    const int lin = param_is_linear(c, p);
    type[p] = (int8_t)lin;
    n_pts[p] = (int16_t)(lin ? 0 : K);
    for (int k = 0; k < K; ++k) {
      /* parameters with no cubic response keep the reference's "-999" marker
       * (Splines/SplineMonolith.cpp:104) */
      coeff_x[p * K + k] = lin ? -999.0f : (float)knot_x(c, k);
    }
  }
}

void m3s_count(const m3s_config* c, int64_t e0, int64_t e1,
               uint32_t* n_cubic, uint32_t* n_linear, uint64_t* tot_cubic, uint64_t* tot_linear) {
  uint64_t tc = 0, tl = 0;
  for (int64_t e = e0; e < e1; ++e) {
    const int mode = event_mode(c, e);
    uint32_t nc = 0, nl = 0;
    for (int p = 0; p < c->n_params; ++p) {
      if (!has_response(c, p, mode)) continue;
      if (param_is_linear(c, p)) ++nl; else ++nc;
    }
    if (n_cubic)  n_cubic[e - e0] = nc;
    if (n_linear) n_linear[e - e0] = nl;
    tc += nc; tl += nl;
  }
  if (tot_cubic)  *tot_cubic = tc;
  if (tot_linear) *tot_linear = tl;
}

/* natural cubic spline through (x_k, y_k): segment k uses
 *   y(x) = y_k + b_k dx + c_k dx^2 + d_k dx^3, dx = x - x_k   (the TSpline3 convention the
 * reference stores, Splines/SplineMonolith.cpp:673-677).  Second derivatives by Thomas. */
static void natural_spline(int K, const double* x, const double* y, double* b, double* cc, double* d) {
  double M[M3S_MAX_KNOTS], cp[M3S_MAX_KNOTS], dp[M3S_MAX_KNOTS];
  M[0] = 0.0; M[K - 1] = 0.0;
  if (K > 2) {
    /* rows i = 1..K-2: h_{i-1} M_{i-1} + 2(h_{i-1}+h_i) M_i + h_i M_{i+1} = 6((y_{i+1}-y_i)/h_i - (y_i-y_{i-1})/h_{i-1}) */
    for (int i = 1; i <= K - 2; ++i) {
      const double hl = x[i] - x[i - 1], hr = x[i + 1] - x[i];
      const double diag = 2.0 * (hl + hr);
      const double rhs = 6.0 * ((y[i + 1] - y[i]) / hr - (y[i] - y[i - 1]) / hl);
      if (i == 1) { cp[i] = hr / diag; dp[i] = rhs / diag; }
      else {
        const double m = diag - hl * cp[i - 1];
        cp[i] = hr / m; dp[i] = (rhs - hl * dp[i - 1]) / m;
      }
    }
    for (int i = K - 2; i >= 1; --i) M[i] = dp[i] - (i == K - 2 ? 0.0 : cp[i] * M[i + 1]);
  }
  for (int k = 0; k < K - 1; ++k) {
    const double h = x[k + 1] - x[k];
    b[k] = (y[k + 1] - y[k]) / h - h * (2.0 * M[k] + M[k + 1]) / 6.0;
    cc[k] = 0.5 * M[k];
    d[k] = (M[k + 1] - M[k]) / (6.0 * h);
  }
  b[K - 1] = 0.0; cc[K - 1] = 0.0; d[K - 1] = 0.0; /* last knot row is never used (SplineBase.cpp:97) */
}

void m3s_fill_splines(const m3s_config* c, int64_t e0, int64_t e1,
                      uint32_t* nParamPerEvent, int16_t* paramNo_arr, uint64_t* knot_off,
                      float* coeff_many, uint32_t* nParamPerEvent_tf1, int16_t* paramNo_tf1,
                      float* coeff_tf1) {
  const int64_t n = e1 - e0;
  const int K = c->n_knots;
  /* pass 1: {count,start} prefix sums (serial, cheap) */
  uint64_t sc = 0, sl = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int mode = event_mode(c, e0 + i);
    uint32_t nc = 0, nl = 0;
    for (int p = 0; p < c->n_params; ++p) {
      if (!has_response(c, p, mode)) continue;
      if (param_is_linear(c, p)) ++nl; else ++nc;
    }
    nParamPerEvent[2 * i] = nc;      nParamPerEvent[2 * i + 1] = (uint32_t)sc;
    nParamPerEvent_tf1[2 * i] = nl;  nParamPerEvent_tf1[2 * i + 1] = (uint32_t)sl;
    sc += nc; sl += nl;
  }
  double xk[M3S_MAX_KNOTS];
  for (int k = 0; k < K; ++k) xk[k] = knot_x(c, k);
  /* pass 2: coefficients (parallel over events); 64-bit running offsets recomputed per event */
  uint64_t* start_c = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n + 1));
  uint64_t* start_l = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n + 1));
  start_c[0] = 0; start_l[0] = 0;
  for (int64_t i = 0; i < n; ++i) {
    start_c[i + 1] = start_c[i] + nParamPerEvent[2 * i];
    start_l[i + 1] = start_l[i] + nParamPerEvent_tf1[2 * i];
  }
  #pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const int64_t e = e0 + i;
    const int mode = event_mode(c, e);
    uint64_t ic = start_c[i], il = start_l[i];
    double y[M3S_MAX_KNOTS], b[M3S_MAX_KNOTS], cc[M3S_MAX_KNOTS], d[M3S_MAX_KNOTS];
    for (int p = 0; p < c->n_params; ++p) {
      if (!has_response(c, p, mode)) continue;
      if (param_is_linear(c, p)) {
        const double a = -0.1 + 0.2 * u01(h4(c->seed, TAG_LIN_A, (uint64_t)e, (uint64_t)p));
        paramNo_tf1[il] = (int16_t)p;
        coeff_tf1[2 * il] = (float)a;
        coeff_tf1[2 * il + 1] = 1.0f;
        ++il;
      } else {
        const double a  = -0.1  + 0.2  * u01(h4(c->seed, TAG_CUB_A, (uint64_t)e, (uint64_t)p));
        const double bq = -0.02 + 0.04 * u01(h4(c->seed, TAG_CUB_B, (uint64_t)e, (uint64_t)p));
        for (int k = 0; k < K; ++k) {
          double v = 1.0 + a * xk[k] + bq * xk[k] * xk[k];
          y[k] = v < 0.05 ? 0.05 : v;
        }
        natural_spline(K, xk, y, b, cc, d);
        paramNo_arr[ic] = (int16_t)p;
        knot_off[ic] = ic * (uint64_t)K;
        float* out = coeff_many + ic * (uint64_t)K * 4u;
        for (int k = 0; k < K; ++k) {
          out[4 * k + 0] = (float)y[k];
          out[4 * k + 1] = (float)b[k];
          out[4 * k + 2] = (float)cc[k];
          out[4 * k + 3] = (float)d[k];
        }
        ++ic;
      }
    }
  }
  free(start_c); free(start_l);
}

static inline int sample_of(const m3s_config* c, int64_t e) {
  int s = 0;
  while (s + 1 < c->n_samples && e >= c->sample_start[s + 1]) ++s;
  return s;
}

void m3s_fill_events(const m3s_config* c, int64_t e0, int64_t e1,
                     int32_t* sample_id, double* kin, int16_t* norm_idx, float* static_w) {
  const int64_t n = e1 - e0;
  #pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const int64_t e = e0 + i;
    if (sample_id) sample_id[i] = sample_of(c, e);
    if (kin) {
      /* Erec ~ Gamma(k=3, theta=0.3) GeV as a sum of three exponentials */
      uint64_t h = h4(c->seed, TAG_KIN_X, (uint64_t)e, 0);
      double u = u01(h); h = mix64(h); u *= u01(h); h = mix64(h); u *= u01(h);
      if (u < 1e-300) u = 1e-300;
      kin[i] = -0.3 * log(u);
      if (c->n_dims > 1)
        kin[n + i] = 3.141592653589793 * u01(h4(c->seed, TAG_KIN_Y, (uint64_t)e, 0));
    }
    if (norm_idx) {
      const int npe = c->n_norm_per_event, N = c->n_norm_params;
      if (N > 0) {
        /* npe distinct norm parameters: start + j*stride pattern */
        const int first = (int)(h4(c->seed, TAG_NORM, (uint64_t)e, 0) % (uint64_t)N);
        for (int j = 0; j < npe; ++j) norm_idx[i * npe + j] = (int16_t)((first + j) % N);
      }
    }
    if (static_w)
      static_w[i] = (float)(0.5 + u01(h4(c->seed, TAG_STATIC, (uint64_t)e, 0)));
  }
}

void m3s_fill_osc(const m3s_config* c, int64_t e0, int64_t e1, int64_t step, float* osc_w) {
  const int64_t n = e1 - e0;
  #pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const uint64_t h = h4(c->seed, TAG_OSC, (uint64_t)(e0 + i), (uint64_t)step);
    /* 1 in 64 events carry an exactly-zero oscillation weight, like the reference's
     * &M3::Zero for NC events with a flavour change (Samples/SampleHandlerFD.cpp:1128-1131) */
    osc_w[i] = ((h & 63u) == 0u) ? 0.0f : (float)u01(mix64(h));
  }
}

void m3s_bin_edges(const m3s_config* c, int sample, int dim, double* edges) {
  const int nb = dim == 0 ? c->nbins_x : c->nbins_y;
  if (dim == 0) {
    /* variable widths on [0, 3 + 0.25*sample] GeV: narrow at low energy */
    const double xmax = 3.0 + 0.25 * (double)sample;
    for (int i = 0; i <= nb; ++i) edges[i] = xmax * pow((double)i / (double)nb, 1.3);
  } else {
    for (int i = 0; i <= nb; ++i) edges[i] = 3.141592653589793 * (double)i / (double)nb;
    edges[nb] = 3.2; /* keep theta < pi strictly inside */
  }
}

void m3s_proposal(const m3s_config* c, int64_t step, double* spline_pars, double* norm_pars) {
  const int K = c->n_knots;
  for (int p = 0; p < c->n_params; ++p) {
    double v;
    if (step == -1) v = 0.0;
    else if (step == -2) v = knot_x(c, 1 + (p % (K > 2 ? K - 2 : 1)));   /* exactly on an interior knot */
    else if (step == -3) v = -3.5;
    else if (step == -4) v = 3.25;
    else {
      v = gaus(h4(c->seed, TAG_PROP, (uint64_t)step, (uint64_t)p));
      if (v > 2.9) v = 2.9;
      if (v < -2.9) v = -2.9;
    }
    spline_pars[p] = v;
  }
  for (int j = 0; j < c->n_norm_params; ++j) {
    double v = step == -1 ? 1.0 : 1.0 + 0.1 * gaus(h4(c->seed, TAG_PROPN, (uint64_t)step, (uint64_t)j));
    if (v < 0.0) v = 0.0; /* norm parameters are bounded below at 0 (ParameterHandlerGeneric.cpp:166-171) */
    norm_pars[j] = v;
  }
}
