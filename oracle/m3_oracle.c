/*
 * m3_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's per-MCMC-step likelihood path, function by
 * function, each citing the reference file:line it follows (paths relative to the
 * mach3-software/MaCh3 tree, v2.4.2).  Types are those of the reference's
 * _LOW_MEMORY_STRUCTS_ build (M3::float_t = float, M3::int_t = short; Manager/Core.h:27-35),
 * the only build in which SMonolith is wired into SampleHandlerFD
 * (Samples/SampleHandlerFD.cpp:1244-1254).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker or the timed CPU baseline.  The product
 * (mach3_b200/) never links, imports or calls it.
 *
 * PARITY PIN STATUS: pinned.  The reference ships no tests, fixtures or golden vectors for this path
 *  and its own build system cannot run here, but its translation units of the path
 * (Splines/SplineMonolith.cpp, SplineBase.cpp, BinnedSplineHandler.cpp, Samples/SampleHandlerFD.cpp,
 * SampleHandlerBase.cpp, BinningHandler.cpp; Samples/SampleStructs.h; Splines/gpuSplineUtils.cu) compile from
 * /root/reference with compile-only stand-ins for the absent ROOT / spdlog / yaml-cpp headers
 * (oracle/ref_host, oracle/ref_gpu -> oracle/_ref/).  Their outputs on seeded inputs are committed under
 * tests/golden/ (ref_host_path.npz, ref_host_fd.npz, ref_host_binning.npz, ref_gpu_weights.npz) with the
 * generating scripts; this file reproduces them bit for bit in the serial build, in both M3::float_t builds
 * (tests/test_reference_path.py, test_reference_host.py, test_reference_gpu.py).
 *
 * The data layout deliberately mirrors the reference (AoS {y,b,c,d} knots, {count,start}
 * CSR, one heap-allocated pointer vector per event) so that timing it is a fair stand-in
 * for the reference's multithreaded CPU path.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define M3O_API __attribute__((visibility("default")))

/* The reference has two builds of this path: MULTITHREAD (OpenMP pragmas; `omp simd
 * reduction(*)` lets the compiler reassociate the per-event products) and the serial build
 * (no pragmas; strictly left-to-right products, SampleHandlerFD::FillArray instead of
 * FillArray_MP).  Both are restated; tests use the serial one where bit-exact weights matter. */
static int g_multithread = 1;
M3O_API void m3o_set_multithread(int on) { g_multithread = on; }
M3O_API int m3o_get_multithread(void) { return g_multithread; }

/* Manager/Core.h:83,91 */
static const double LOW_MC_BOUND = .00001;
enum { UnderOverFlowBin = -1 };
/* Splines/SplineCommon.h:13,15 */
enum { nCoeff = 4, nTF1Coeff = 2 };
/* Samples/SampleStructs.h:105-112 */
enum TestStatistic { kPoisson, kBarlowBeeston, kIceCube, kPearson, kDembinskiAbdelmotteleb, kNTestStatistics };

/* ============================================================================================
 * Splines: FastSplineInfo (Splines/SplineStructs.h:21-44) + SplineBase + SMonolith state
 * ========================================================================================== */
typedef struct {
  short nPts;
  const float* xPts;              /* knot positions (first spline seen for the parameter) */
  const double* xPts_d;           /* the same in M3::float_t = double (default build, splines built in-process:
                                     Splines/SplineMonolith.cpp:393-404); NULL: the float ones (values reloaded
                                     from a spline file, Splines/SplineBase.cpp:166-190, or _LOW_MEMORY_STRUCTS_) */
  int   n_x;                      /* xPts.size() */
  short CurrSegment;
  const double* splineParsPointer;
} FastSplineInfo;

typedef struct {
  /* SplineBase (Splines/SplineBase.h:77-81) */
  short nParams;
  FastSplineInfo* SplineInfoArray;
  short* SplineSegments;
  float* ParamValues;
  /* SplineMonoStruct (Splines/SplineCommon.h:30-50); knot offsets widened to 64 bit */
  const float* coeff_x;
  const float* coeff_many;
  const uint64_t* nKnots_arr;
  const short* paramNo_arr;
  /* SMonolith (Splines/SplineMonolith.h:96-137) */
  unsigned int NEvents;
  short _max_knots;
  uint64_t NSplines_valid, NTF1_valid;
  const uint32_t* cpu_nParamPerEvent;
  const uint32_t* cpu_nParamPerEvent_tf1;
  const float* cpu_coeff_TF1_many;
  const short* cpu_paramNo_TF1_arr;
  float* cpu_weights_spline_var;
  float* cpu_weights_tf1_var;
  float* cpu_total_weights;
  /* 64-bit {start} of every event: the reference's unsigned start index overflows beyond
   * 2^32 responses; recomputed here from the counts */
  uint64_t* start_c;
  uint64_t* start_l;
} SMonolith;

M3O_API SMonolith* m3o_monolith_create(int nParams, int max_knots, const float* coeff_x, const short* n_pts,
                                       unsigned int NEvents,
                                       const uint32_t* nParamPerEvent, const short* paramNo_arr,
                                       const uint64_t* nKnots_arr, const float* coeff_many,
                                       const uint32_t* nParamPerEvent_tf1, const short* paramNo_tf1,
                                       const float* coeff_tf1) {
  SMonolith* m = (SMonolith*)calloc(1, sizeof(SMonolith));
  m->nParams = (short)nParams;
  m->_max_knots = (short)max_knots;
  m->coeff_x = coeff_x; m->coeff_many = coeff_many; m->nKnots_arr = nKnots_arr; m->paramNo_arr = paramNo_arr;
  m->NEvents = NEvents;
  m->cpu_nParamPerEvent = nParamPerEvent; m->cpu_nParamPerEvent_tf1 = nParamPerEvent_tf1;
  m->cpu_coeff_TF1_many = coeff_tf1; m->cpu_paramNo_TF1_arr = paramNo_tf1;
  m->SplineInfoArray = (FastSplineInfo*)calloc((size_t)nParams, sizeof(FastSplineInfo));
  /* Splines/SplineMonolith.cpp:85-95: segments 0, values -999 */
  m->SplineSegments = (short*)calloc((size_t)nParams, sizeof(short));
  m->ParamValues = (float*)malloc(sizeof(float) * (size_t)nParams);
  for (int j = 0; j < nParams; ++j) {
    m->ParamValues[j] = -999;
    m->SplineInfoArray[j].nPts = n_pts[j];
    m->SplineInfoArray[j].n_x = n_pts[j] > 0 ? n_pts[j] : 0;
    m->SplineInfoArray[j].xPts = coeff_x + (size_t)j * (size_t)max_knots;
    m->SplineInfoArray[j].CurrSegment = 0;
    m->SplineInfoArray[j].splineParsPointer = NULL;
  }
  m->start_c = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)NEvents + 1));
  m->start_l = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)NEvents + 1));
  m->start_c[0] = 0; m->start_l[0] = 0;
  for (unsigned int e = 0; e < NEvents; ++e) {
    m->start_c[e + 1] = m->start_c[e] + nParamPerEvent[2 * e];
    m->start_l[e + 1] = m->start_l[e] + nParamPerEvent_tf1[2 * e];
  }
  m->NSplines_valid = m->start_c[NEvents];
  m->NTF1_valid = m->start_l[NEvents];
  /* Splines/SplineMonolith.cpp:241-245 */
  m->cpu_total_weights = (float*)calloc((size_t)NEvents + 1, sizeof(float));
  m->cpu_weights_spline_var = (float*)calloc((size_t)m->NSplines_valid + 1, sizeof(float));
  m->cpu_weights_tf1_var = (float*)calloc((size_t)m->NTF1_valid + 1, sizeof(float));
  return m;
}

M3O_API void m3o_monolith_destroy(SMonolith* m) {
  if (!m) return;
  free(m->SplineInfoArray); free(m->SplineSegments); free(m->ParamValues);
  free(m->cpu_total_weights); free(m->cpu_weights_spline_var); free(m->cpu_weights_tf1_var);
  free(m->start_c); free(m->start_l);
  free(m);
}

/* SMonolith::setSplinePointers (Splines/SplineMonolith.h:44-46); here param i -> &base[i] */
M3O_API void m3o_set_spline_pointers(SMonolith* m, const double* base) {
  for (short i = 0; i < m->nParams; ++i) m->SplineInfoArray[i].splineParsPointer = base + i;
}

/* SplineBase::FindSplineSegment (Splines/SplineBase.cpp:44-109) */
#define XARR(k) (xd ? xd[k] : (double)xf[k])      /* float xvar against M3::float_t knots: the comparison is in double */
M3O_API void m3o_find_spline_segment(SMonolith* m) {
  for (short i = 0; i < m->nParams; ++i) {
    const short nPoints = m->SplineInfoArray[i].nPts;
    const float* xf = m->SplineInfoArray[i].xPts;
    const double* xd = m->SplineInfoArray[i].xPts_d;

    /* :54-55 the variation is narrowed to float, stored for every parameter */
    const float xvar = (float)(*m->SplineInfoArray[i].splineParsPointer);
    m->ParamValues[i] = xvar;

    /* :60 parameters without any spline are skipped */
    if (m->SplineInfoArray[i].n_x == 0) continue;

    short segment = 0;
    short kHigh = (short)(nPoints - 1);
    const short PreviousSegment = m->SplineInfoArray[i].CurrSegment;

    if (xvar <= XARR(0)) {                                   /* :69 */
      segment = 0;
    } else if (xvar >= XARR(nPoints - 1)) {                  /* :72 */
      segment = kHigh;
    } else if (XARR(PreviousSegment + 1) > xvar && xvar >= XARR(PreviousSegment)) { /* :76 */
      segment = PreviousSegment;
    } else {                                                   /* :79-95 binary search */
      short kHalf = 0;
      while (kHigh - segment > 1) {
        kHalf = (short)((segment + kHigh) / 2);
        if (xvar > XARR(kHalf)) segment = kHalf;
        else kHigh = kHalf;
      }
    }
    if (segment >= nPoints - 1 && nPoints > 1) segment = (short)(nPoints - 2);   /* :97 */

    m->SplineInfoArray[i].CurrSegment = segment;               /* :102-103 */
    m->SplineSegments[i] = (short)m->SplineInfoArray[i].CurrSegment;
  }
}
#undef XARR

/* knots as doubles (the reference's default build when the monolith is built in-process); xPts[nParams*max_knots] */
M3O_API void m3o_set_knots_f64(SMonolith* m, const double* xPts) {
  for (short i = 0; i < m->nParams; ++i) m->SplineInfoArray[i].xPts_d = xPts ? xPts + (size_t)i * (size_t)m->_max_knots : NULL;
}

/* SMonolith::CalcSplineWeights (Splines/SplineMonolith.cpp:727-788), serial build */
static void calc_spline_weights_serial(SMonolith* m) {
  for (uint64_t splineNum = 0; splineNum < m->NSplines_valid; ++splineNum) {
    const short Param = m->paramNo_arr[splineNum];
    const short segment = m->SplineSegments[Param];
    const short segment_X = (short)(Param * m->_max_knots + segment);
    const uint64_t CurrentKnotPos = m->nKnots_arr[splineNum] * nCoeff + (uint64_t)(segment * nCoeff);
    const float fY = m->coeff_many[CurrentKnotPos];
    const float fB = m->coeff_many[CurrentKnotPos + 1];
    const float fC = m->coeff_many[CurrentKnotPos + 2];
    const float fD = m->coeff_many[CurrentKnotPos + 3];
    const float dx = m->ParamValues[Param] - m->coeff_x[segment_X];
    m->cpu_weights_spline_var[splineNum] = fmaf(dx, fmaf(dx, fmaf(dx, fD, fC), fB), fY);
  }
  for (uint64_t tf1Num = 0; tf1Num < m->NTF1_valid; ++tf1Num) {
    const float x = m->ParamValues[m->cpu_paramNo_TF1_arr[tf1Num]];
    const uint64_t TF1_Index = tf1Num * nTF1Coeff;
    m->cpu_weights_tf1_var[tf1Num] = fmaf(m->cpu_coeff_TF1_many[TF1_Index], x, m->cpu_coeff_TF1_many[TF1_Index + 1]);
  }
}

/* SMonolith::CalcSplineWeights (Splines/SplineMonolith.cpp:727-788), MULTITHREAD build */
M3O_API void m3o_calc_spline_weights(SMonolith* m) {
  if (!g_multithread) { calc_spline_weights_serial(m); return; }
  #pragma omp parallel
  {
    #pragma omp for simd nowait
    for (uint64_t splineNum = 0; splineNum < m->NSplines_valid; ++splineNum) {
      const short Param = m->paramNo_arr[splineNum];                              /* :741 */
      const short segment = m->SplineSegments[Param];                             /* :744 */
      const short segment_X = (short)(Param * m->_max_knots + segment);           /* :747 */
      const uint64_t CurrentKnotPos = m->nKnots_arr[splineNum] * nCoeff + (uint64_t)(segment * nCoeff); /* :750 */
      const float fY = m->coeff_many[CurrentKnotPos];
      const float fB = m->coeff_many[CurrentKnotPos + 1];
      const float fC = m->coeff_many[CurrentKnotPos + 2];
      const float fD = m->coeff_many[CurrentKnotPos + 3];
      const float dx = m->ParamValues[Param] - m->coeff_x[segment_X];             /* :759 */
      m->cpu_weights_spline_var[splineNum] = fmaf(dx, fmaf(dx, fmaf(dx, fD, fC), fB), fY); /* :762 */
    }
    #pragma omp for simd
    for (uint64_t tf1Num = 0; tf1Num < m->NTF1_valid; ++tf1Num) {
      const float x = m->ParamValues[m->cpu_paramNo_TF1_arr[tf1Num]];             /* :773 */
      const uint64_t TF1_Index = tf1Num * nTF1Coeff;
      const float a = m->cpu_coeff_TF1_many[TF1_Index];
      const float b = m->cpu_coeff_TF1_many[TF1_Index + 1];
      m->cpu_weights_tf1_var[tf1Num] = fmaf(a, x, b);                             /* :780 */
    }
  }
}

/* SMonolith::CalcTotalEventWeight (Splines/SplineMonolith.cpp:792-830), serial build */
static void calc_total_event_weight_serial(SMonolith* m) {
  for (unsigned int EventNum = 0; EventNum < m->NEvents; ++EventNum) {
    float totalWeight = 1.0f;
    const uint64_t startIndex = m->start_c[EventNum];
    const unsigned int numParams = m->cpu_nParamPerEvent[2 * EventNum];
    for (unsigned int id = 0; id < numParams; ++id) totalWeight *= m->cpu_weights_spline_var[startIndex + id];
    const uint64_t startIndex_tf1 = m->start_l[EventNum];
    const unsigned int numParams_tf1 = m->cpu_nParamPerEvent_tf1[2 * EventNum];
    for (unsigned int id = 0; id < numParams_tf1; ++id) totalWeight *= m->cpu_weights_tf1_var[startIndex_tf1 + id];
    m->cpu_total_weights[EventNum] = totalWeight;
  }
}

/* SMonolith::CalcTotalEventWeight (Splines/SplineMonolith.cpp:792-830), MULTITHREAD build */
M3O_API void m3o_calc_total_event_weight(SMonolith* m) {
  if (!g_multithread) { calc_total_event_weight_serial(m); return; }
  #pragma omp parallel for
  for (unsigned int EventNum = 0; EventNum < m->NEvents; ++EventNum) {
    float totalWeight = 1.0f;
    const unsigned int Offset = 2 * EventNum;
    const uint64_t startIndex = m->start_c[EventNum];
    const unsigned int numParams = m->cpu_nParamPerEvent[Offset];
    #pragma omp simd reduction(*:totalWeight)
    for (unsigned int id = 0; id < numParams; ++id) totalWeight *= m->cpu_weights_spline_var[startIndex + id];
    const uint64_t startIndex_tf1 = m->start_l[EventNum];
    const unsigned int numParams_tf1 = m->cpu_nParamPerEvent_tf1[Offset];
    #pragma omp simd reduction(*:totalWeight)
    for (unsigned int id = 0; id < numParams_tf1; ++id) totalWeight *= m->cpu_weights_tf1_var[startIndex_tf1 + id];
    m->cpu_total_weights[EventNum] = totalWeight;
  }
}

/* SMonolith::Evaluate, CPU build (Splines/SplineMonolith.cpp:712-723) */
M3O_API void m3o_evaluate(SMonolith* m) {
  m3o_find_spline_segment(m);
  m3o_calc_spline_weights(m);
  m3o_calc_total_event_weight(m);
}

M3O_API const short* m3o_segments(const SMonolith* m) { return m->SplineSegments; }
M3O_API const float* m3o_param_values(const SMonolith* m) { return m->ParamValues; }
M3O_API const float* m3o_total_weights(const SMonolith* m) { return m->cpu_total_weights; }       /* retPointer(0) */
M3O_API const float* m3o_spline_weights(const SMonolith* m) { return m->cpu_weights_spline_var; }
M3O_API const float* m3o_tf1_weights(const SMonolith* m) { return m->cpu_weights_tf1_var; }
M3O_API uint64_t m3o_n_splines_valid(const SMonolith* m) { return m->NSplines_valid; }
M3O_API uint64_t m3o_n_tf1_valid(const SMonolith* m) { return m->NTF1_valid; }
/* let tests force the cached segment (history dependence of SplineBase.cpp:76) */
M3O_API void m3o_set_curr_segment(SMonolith* m, int p, int seg) { m->SplineInfoArray[p].CurrSegment = (short)seg; }

/* ============================================================================================
 * Binning: SampleBinningInfo (Samples/SampleStructs.h:232-676) + BinningHandler
 * ========================================================================================== */
typedef struct {        /* Samples/SampleStructs.h:172-183 */
  double lower_binedge, upper_binedge, lower_lower_binedge, upper_upper_binedge;
} BinShiftLookup;

#define M3O_MAX_DIM 4
typedef struct {
  int nDim;
  int AxisNBins[M3O_MAX_DIM];
  double* BinEdges[M3O_MAX_DIM];
  BinShiftLookup* BinLookup[M3O_MAX_DIM];
  int Strides[M3O_MAX_DIM];
  int nBins;
  int GlobalOffset;
  /* non-uniform binning (Samples/SampleStructs.h:186-220,249-252): boxes + the "mega bin" grid that maps to them */
  int Uniform;
  int nBoxes;
  double* Extent;            /* [box][dim][2] = BinInfo::Extent */
  int* GridStart;            /* CSR over mega bins: BinGridMapping[mega] = GridIdx[GridStart[mega] .. GridStart[mega+1]) */
  int* GridIdx;
} SampleBinningInfo;

/* SampleBinningInfo::InitialiseLookUpSingleDimension (Samples/SampleStructs.h:618-647) */
static void InitialiseLookUpSingleDimension(BinShiftLookup* Bin_Lookup, const double* Bin_Edges, const int TotBins) {
  for (int bin_i = 0; bin_i < TotBins; bin_i++) {
    double low_lower_edge = -999999.123456;      /* M3::_DEFAULT_RETURN_VAL_ */
    double low_edge = Bin_Edges[bin_i];
    double upper_edge = Bin_Edges[bin_i + 1];
    double upper_upper_edge = -999999.123456;
    if (bin_i == 0) low_lower_edge = Bin_Edges[0];
    else low_lower_edge = Bin_Edges[bin_i - 1];
    if (bin_i + 2 < TotBins) upper_upper_edge = Bin_Edges[bin_i + 2];
    else if (bin_i + 1 < TotBins) upper_upper_edge = Bin_Edges[bin_i + 1];
    Bin_Lookup[bin_i].lower_binedge = low_edge;
    Bin_Lookup[bin_i].upper_binedge = upper_edge;
    Bin_Lookup[bin_i].lower_lower_binedge = low_lower_edge;
    Bin_Lookup[bin_i].upper_upper_binedge = upper_upper_edge;
  }
}

/* std::upper_bound over doubles */
static int upper_bound_d(const double* a, int n, double v) {
  int lo = 0, len = n;
  while (len > 0) {
    int half = len >> 1;
    if (!(v < a[lo + half])) { lo += half + 1; len -= half + 1; }
    else len = half;
  }
  return lo;
}

/* SampleBinningInfo::FindBin (Samples/SampleStructs.h:577-613) */
static int FindBin(const double KinVar, const int NomBin, const int N_Bins,
                   const double* Bin_Edges, const BinShiftLookup* Bin_Lookup) {
  if (KinVar < Bin_Edges[0] || KinVar >= Bin_Edges[N_Bins]) return UnderOverFlowBin;   /* :584 */
  if (NomBin > UnderOverFlowBin) {                                                     /* :588 */
    const BinShiftLookup* Bin = &Bin_Lookup[NomBin];
    const double lower = Bin->lower_binedge, upper = Bin->upper_binedge;
    const double lower_lower = Bin->lower_lower_binedge, upper_upper = Bin->upper_upper_binedge;
    if (KinVar < upper && KinVar >= lower) return NomBin;                              /* :597 */
    if (KinVar < lower && KinVar >= lower_lower) return NomBin - 1;                    /* :602 */
    if (KinVar < upper_upper && KinVar >= upper) return NomBin + 1;                    /* :606 */
  }
  return upper_bound_d(Bin_Edges, N_Bins + 1, KinVar) - 1;                             /* :612 */
}

/* BinningHandler::FindNominalBin (Samples/BinningHandler.cpp:294-307) */
static int FindNominalBin(const SampleBinningInfo* info, const int iDim, const double Var) {
  const double* edges = info->BinEdges[iDim];
  const int ne = info->AxisNBins[iDim] + 1;
  if (Var < edges[0] || Var >= edges[ne - 1]) return UnderOverFlowBin;
  return upper_bound_d(edges, ne, Var) - 1;
}


/* ============================================================================================
 * BinnedSplineHandler (Splines/BinnedSplineHandler.h:110-135, .cpp:295-341), _LOW_MEMORY_STRUCTS_
 * build (M3::float_t = float, like the SMonolith path above; the default build holds the same
 * arrays in double -- see m3o_binned_calc_spline_weights at the end of this file).
 *   weightvec_Monolith[n_slots]         one weight per (sample,osc,syst,mode,bin) slot, 1.0 for flat
 *   uniquesplinevec_Monolith[iSpline]   spline parameter of the slot
 *   coeffindexvec[iSpline]              first knot of the slot's spline in manycoeff_arr / xcoeff_arr
 *   uniquecoeffindices[]                the non-flat slots: only these are evaluated (:311-340)
 * Segments come from SplineBase::FindSplineSegment on the parameter's xPts (first spline seen).
 * ========================================================================================== */
typedef struct BinnedSplineHandler_ {
  SMonolith base;                       /* SplineBase part only: nParams, SplineInfoArray, SplineSegments, ParamValues */
  int64_t n_slots, n_unique;
  const int* uniquesplinevec_Monolith;
  const int* coeffindexvec;
  const int* uniquecoeffindices;
  const float* manycoeff_arr;
  const float* xcoeff_arr;
  float* weightvec_Monolith;
} BinnedSplineHandler;

M3O_API BinnedSplineHandler* m3o_binned_create(int nParams, int max_knots, const float* knot_x, const short* n_pts,
                                               int64_t n_slots, const int* uniquesplinevec_Monolith, const int* coeffindexvec,
                                               int64_t n_unique, const int* uniquecoeffindices,
                                               const float* manycoeff_arr, const float* xcoeff_arr) {
  BinnedSplineHandler* b = (BinnedSplineHandler*)calloc(1, sizeof(BinnedSplineHandler));
  b->base.nParams = (short)nParams;
  b->base.SplineInfoArray = (FastSplineInfo*)calloc((size_t)nParams, sizeof(FastSplineInfo));
  b->base.SplineSegments = (short*)calloc((size_t)nParams, sizeof(short));
  b->base.ParamValues = (float*)calloc((size_t)nParams, sizeof(float));
  for (int j = 0; j < nParams; ++j) {
    b->base.SplineInfoArray[j].nPts = n_pts[j];
    b->base.SplineInfoArray[j].n_x = n_pts[j] > 0 ? n_pts[j] : 0;
    b->base.SplineInfoArray[j].xPts = knot_x + (size_t)j * (size_t)max_knots;
  }
  b->n_slots = n_slots; b->n_unique = n_unique;
  b->uniquesplinevec_Monolith = uniquesplinevec_Monolith; b->coeffindexvec = coeffindexvec;
  b->uniquecoeffindices = uniquecoeffindices; b->manycoeff_arr = manycoeff_arr; b->xcoeff_arr = xcoeff_arr;
  b->weightvec_Monolith = (float*)malloc(sizeof(float) * (size_t)(n_slots > 0 ? n_slots : 1));
  for (int64_t i = 0; i < n_slots; ++i) b->weightvec_Monolith[i] = 1.0f;     /* flat splines stay at 1 (.cpp:236,283) */
  return b;
}
M3O_API void m3o_binned_destroy(BinnedSplineHandler* b) {
  if (!b) return;
  free(b->base.SplineInfoArray); free(b->base.SplineSegments); free(b->base.ParamValues); free(b->weightvec_Monolith);
  free(b);
}
M3O_API void m3o_binned_set_pointers(BinnedSplineHandler* b, const double* pars) { m3o_set_spline_pointers(&b->base, pars); }

/* BinnedSplineHandler::CalcSplineWeights (Splines/BinnedSplineHandler.cpp:306-341) */
static void binned_calc_spline_weights(BinnedSplineHandler* h) {
  const int64_t n = h->n_unique;
  #pragma omp parallel for simd if (g_multithread)
  for (int64_t iCoeff = 0; iCoeff < n; ++iCoeff) {
    const int iSpline = h->uniquecoeffindices[iCoeff];
    const short uniqueIndex = (short)h->uniquesplinevec_Monolith[iSpline];
    const short currentsegment = (short)h->base.SplineSegments[uniqueIndex];
    const int segCoeff = h->coeffindexvec[iSpline] + currentsegment;
    const int coeffOffset = segCoeff * nCoeff;
    const float y = h->manycoeff_arr[coeffOffset + 0];
    const float b = h->manycoeff_arr[coeffOffset + 1];
    const float c = h->manycoeff_arr[coeffOffset + 2];
    const float d = h->manycoeff_arr[coeffOffset + 3];
    const float xvar = (float)(*h->base.SplineInfoArray[uniqueIndex].splineParsPointer);   /* :327, M3::float_t */
    const float dx = xvar - h->xcoeff_arr[segCoeff];                                       /* :329 */
    float weight = fmaf(dx, fmaf(dx, fmaf(dx, d, c), b), y);                               /* :332 */
    if (weight < 0) weight = 0.;                                                           /* :337 */
    h->weightvec_Monolith[iSpline] = weight;
  }
}
/* BinnedSplineHandler::Evaluate (:295-303) */
M3O_API void m3o_binned_evaluate(BinnedSplineHandler* b) {
  m3o_find_spline_segment(&b->base);
  binned_calc_spline_weights(b);
}
M3O_API const float* m3o_binned_weights(const BinnedSplineHandler* b) { return b->weightvec_Monolith; }
M3O_API const short* m3o_binned_segments(const BinnedSplineHandler* b) { return b->base.SplineSegments; }


/* ============================================================================================
 * The binned-spline path in the reference's DEFAULT build (M3::float_t = double, Manager/Core.h:27-51):
 * FastSplineInfo::xPts, the coefficient arrays, weightvec_Monolith, the oscillation/extra weights and
 * CalcWeightTotal's product are double; FindSplineSegment still narrows the parameter to float
 * (Splines/SplineBase.cpp:54) while CalcSplineWeights reads it un-narrowed (BinnedSplineHandler.cpp:327);
 * M3::fmaf_t is std::fma.
 * ========================================================================================== */
typedef struct BinnedSplineHandlerD_ {
  int nParams, max_knots;
  const double* knot_x; const short* n_pts;
  short* CurrSegment; short* SplineSegments; float* ParamValues;
  const double* pars;
  int64_t n_slots, n_unique;
  const int* uniquesplinevec_Monolith; const int* coeffindexvec; const int* uniquecoeffindices;
  const double* manycoeff_arr; const double* xcoeff_arr;
  double* weightvec_Monolith;
} BinnedSplineHandlerD;

M3O_API BinnedSplineHandlerD* m3o_binnedd_create(int nParams, int max_knots, const double* knot_x, const short* n_pts,
                                                 int64_t n_slots, const int* uniquesplinevec_Monolith, const int* coeffindexvec,
                                                 int64_t n_unique, const int* uniquecoeffindices,
                                                 const double* manycoeff_arr, const double* xcoeff_arr) {
  BinnedSplineHandlerD* b = (BinnedSplineHandlerD*)calloc(1, sizeof(BinnedSplineHandlerD));
  b->nParams = nParams; b->max_knots = max_knots; b->knot_x = knot_x; b->n_pts = n_pts;
  b->CurrSegment = (short*)calloc((size_t)nParams, sizeof(short));
  b->SplineSegments = (short*)calloc((size_t)nParams, sizeof(short));
  b->ParamValues = (float*)calloc((size_t)nParams, sizeof(float));
  b->n_slots = n_slots; b->n_unique = n_unique;
  b->uniquesplinevec_Monolith = uniquesplinevec_Monolith; b->coeffindexvec = coeffindexvec; b->uniquecoeffindices = uniquecoeffindices;
  b->manycoeff_arr = manycoeff_arr; b->xcoeff_arr = xcoeff_arr;
  b->weightvec_Monolith = (double*)malloc(sizeof(double) * (size_t)(n_slots > 0 ? n_slots : 1));
  for (int64_t i = 0; i < n_slots; ++i) b->weightvec_Monolith[i] = 1.0;
  return b;
}
M3O_API void m3o_binnedd_destroy(BinnedSplineHandlerD* b) {
  if (!b) return;
  free(b->CurrSegment); free(b->SplineSegments); free(b->ParamValues); free(b->weightvec_Monolith); free(b);
}
M3O_API void m3o_binnedd_set_pointers(BinnedSplineHandlerD* b, const double* pars) { b->pars = pars; }

/* SplineBase::FindSplineSegment (Splines/SplineBase.cpp:44-109) with xPts in double */
static void find_spline_segment_d(BinnedSplineHandlerD* m) {
  for (int i = 0; i < m->nParams; ++i) {
    const short nPoints = m->n_pts[i];
    const double* xArray = m->knot_x + (size_t)i * (size_t)m->max_knots;
    const float xvar = (float)(m->pars[i]);                    /* :54 */
    m->ParamValues[i] = xvar;
    if (nPoints == 0) continue;
    short segment = 0;
    short kHigh = (short)(nPoints - 1);
    const short PreviousSegment = m->CurrSegment[i];
    if (xvar <= xArray[0]) segment = 0;
    else if (xvar >= xArray[nPoints - 1]) segment = kHigh;
    else if (xArray[PreviousSegment + 1] > xvar && xvar >= xArray[PreviousSegment]) segment = PreviousSegment;
    else {
      short kHalf = 0;
      while (kHigh - segment > 1) {
        kHalf = (short)((segment + kHigh) / 2);
        if (xvar > xArray[kHalf]) segment = kHalf; else kHigh = kHalf;
      }
    }
    if (segment >= nPoints - 1 && nPoints > 1) segment = (short)(nPoints - 2);
    m->CurrSegment[i] = segment;
    m->SplineSegments[i] = segment;
  }
}
/* BinnedSplineHandler::Evaluate (Splines/BinnedSplineHandler.cpp:295-341), default build */
M3O_API void m3o_binnedd_evaluate(BinnedSplineHandlerD* h) {
  find_spline_segment_d(h);
  const int64_t n = h->n_unique;
  #pragma omp parallel for simd if (g_multithread)
  for (int64_t iCoeff = 0; iCoeff < n; ++iCoeff) {
    const int iSpline = h->uniquecoeffindices[iCoeff];
    const short uniqueIndex = (short)h->uniquesplinevec_Monolith[iSpline];
    const short currentsegment = (short)h->SplineSegments[uniqueIndex];
    const int segCoeff = h->coeffindexvec[iSpline] + currentsegment;
    const int coeffOffset = segCoeff * nCoeff;
    const double y = h->manycoeff_arr[coeffOffset + 0], b = h->manycoeff_arr[coeffOffset + 1];
    const double c = h->manycoeff_arr[coeffOffset + 2], d = h->manycoeff_arr[coeffOffset + 3];
    const double xvar = h->pars[uniqueIndex];                                  /* :327 un-narrowed */
    const double dx = xvar - h->xcoeff_arr[segCoeff];
    double weight = fma(dx, fma(dx, fma(dx, d, c), b), y);
    if (weight < 0) weight = 0.;
    h->weightvec_Monolith[iSpline] = weight;
  }
}
M3O_API const double* m3o_binnedd_weights(const BinnedSplineHandlerD* b) { return b->weightvec_Monolith; }
M3O_API const short* m3o_binnedd_segments(const BinnedSplineHandlerD* b) { return b->SplineSegments; }

/* ============================================================================================
 * Samples: EventInfo (Samples/FarDetectorCoreInfoStruct.h:82-126) + SampleHandlerFD state
 * ========================================================================================== */
typedef struct {          /* Samples/SampleStructs.h:149-157 */
  int ParamToCutOnIt;
  double LowerBound, UpperBound;
} KinematicCut;

typedef struct {
  const double** norm_pointers;         int n_norm;      /* std::vector<const double*>        */
  const float**  total_weight_pointers; int n_tw;        /* std::vector<const M3::float_t*>   */
  const double** KinVar;                                 /* std::vector<const double*>        */
  int*           NomBin;                int n_dim;       /* std::vector<int>                  */
  int NominalSample;
} EventInfo;

typedef struct {
  unsigned int nEvents;
  int nSamples;
  SampleBinningInfo* SampleBinning;
  int TotalBins;
  EventInfo* MCSamples;
  double* SampleHandlerFD_array;
  double* SampleHandlerFD_array_w2;
  double* SampleHandlerFD_data;
  int fTestStatistic;
  int FirstTimeW2;   /* Samples/SampleHandlerFD.h: FirstTimeW2 = true initially */
  int UpdateW2;      /* LikelihoodOptions:UpdateW2 (Samples/SampleHandlerFD.cpp:64) */
  SMonolith* SplineHandler;
  struct BinnedSplineHandler_* BinnedHandler;   /* the other SplineBase implementation (either/or) */
  /* default build (M3::float_t = double) of the binned path: per-event pointer vectors in double */
  struct BinnedSplineHandlerD_* BinnedHandlerD;
  const double*** tw_d;   /* [event] -> std::vector<const M3::float_t*> total_weight_pointers */
  int* n_tw_d;
  /* std::vector<std::vector<KinematicCut>> StoredSelection / Selection (Samples/SampleHandlerFD.h:364-371), flattened:
   * cuts of sample s are Selection[SelStart[s] .. SelStart[s+1]).  ReturnKinematicParameter(var, event) -- pure
   * virtual in the reference, experiment code -- is a table look-up here: CutValues[var*nEvents + event]. */
  int* SelStart; KinematicCut* StoredSelection; KinematicCut* Selection; int nCuts;
  const double* CutValues;
  /* std::vector<std::vector<FunctionalShifter*>> funcParsGrid (Samples/SampleHandlerFD.h:257), flattened: the shifters
   * of event e are entries ShiftStart[e] .. ShiftStart[e+1]); FunctionalShifter = {valuePtr, funcPtr}
   * (Samples/SampleStructs.h:161-168).  The functions are experiment code; the ones restated here are the linear family
   * x_target += (*valuePtr) * coef.  ResetShifts() restores the nominal values kept in KinNominal / CutNominal. */
  int64_t* ShiftStart; int* ShiftPar; int* ShiftTarget; double* ShiftCoef; const double* ShiftValues;
  double* KinShifted; double* KinNominal; int nKinRows; double* CutShifted; double* CutNominal; int nCutRows;
} SampleHandlerFD;


/* SampleBinningInfo::InitNonUniform + InitialiseGridMapping (Samples/SampleStructs.h:394-528): the boxes are
 * mapped through a regular grid of BinsPerDimension = 10 "mega bins" per dimension spanning their bounding
 * box; every mega bin lists, in box order, the boxes that overlap it (a_hi > b_lo && a_lo < b_hi). */
static void InitNonUniform(SampleBinningInfo* b, int nDim, int nBoxes, const double* extent) {
  enum { BinsPerDimension = 10 };
  b->Uniform = 0; b->nDim = nDim; b->nBoxes = nBoxes; b->nBins = nBoxes;
  b->Extent = (double*)malloc(sizeof(double) * (size_t)nBoxes * (size_t)nDim * 2);
  memcpy(b->Extent, extent, sizeof(double) * (size_t)nBoxes * (size_t)nDim * 2);
  int NGridBins = 1, stride = 1;
  for (int d = 0; d < nDim; ++d) {
    double MinVal = 1.7976931348623157e308, MaxVal = -1.7976931348623157e308;
    for (int i = 0; i < nBoxes; ++i) {
      const double lo = extent[((size_t)i * nDim + d) * 2], hi = extent[((size_t)i * nDim + d) * 2 + 1];
      if (lo < MinVal) MinVal = lo;
      if (hi > MaxVal) MaxVal = hi;
    }
    b->AxisNBins[d] = BinsPerDimension;
    b->BinEdges[d] = (double*)malloc(sizeof(double) * (BinsPerDimension + 1));
    const double BinWidth = (MaxVal - MinVal) / (double)BinsPerDimension;
    for (int e = 0; e <= BinsPerDimension; ++e) b->BinEdges[d][e] = MinVal + (double)e * BinWidth;     /* :518-521 */
    b->BinLookup[d] = (BinShiftLookup*)malloc(sizeof(BinShiftLookup) * BinsPerDimension);
    InitialiseLookUpSingleDimension(b->BinLookup[d], b->BinEdges[d], BinsPerDimension);
    b->Strides[d] = stride; stride *= BinsPerDimension; NGridBins *= BinsPerDimension;
  }
  b->GridStart = (int*)calloc((size_t)NGridBins + 1, sizeof(int));
  for (int pass = 0; pass < 2; ++pass) {
    int total = 0;
    for (int g = 0; g < NGridBins; ++g) {
      if (pass == 1) b->GridStart[g] = total;
      int rem = g;
      double cell[M3O_MAX_DIM][2];
      for (int d = 0; d < nDim; ++d) { const int i = rem % BinsPerDimension; rem /= BinsPerDimension; cell[d][0] = b->BinEdges[d][i]; cell[d][1] = b->BinEdges[d][i + 1]; }
      for (int i = 0; i < nBoxes; ++i) {
        int overlap = 1;
        for (int d = 0; d < nDim; ++d) {
          const double a_lo = extent[((size_t)i * nDim + d) * 2], a_hi = extent[((size_t)i * nDim + d) * 2 + 1];
          if (!(a_hi > cell[d][0] && a_lo < cell[d][1])) { overlap = 0; break; }                       /* :425 */
        }
        if (overlap) { if (pass == 1) b->GridIdx[total] = i; ++total; }
      }
    }
    if (pass == 0) b->GridIdx = (int*)malloc(sizeof(int) * (size_t)(total > 0 ? total : 1));
    else b->GridStart[NGridBins] = total;
  }
}

/* like m3o_sample_create, but sample i may be non-uniform (uniform[i] == 0): then nbins[i*M3O_MAX_DIM] is its
 * number of boxes and its part of `edges` holds nBoxes*nDim {lo,hi} pairs */
M3O_API void* m3o_sample_create_ex(unsigned int nEvents, int nSamples, const int* nDim, const int* uniform,
                                   const int* nbins, const double* edges, int test_statistic, int update_w2);

/* binning description: for each sample, nDim then per dim nbins; edges concatenated */
static SampleHandlerFD* sample_create(unsigned int nEvents, int nSamples, const int* nDim, const int* uniform,
                                      const int* nbins, const double* edges, int test_statistic, int update_w2);
M3O_API SampleHandlerFD* m3o_sample_create(unsigned int nEvents, int nSamples, const int* nDim,
                                           const int* nbins /*[nSamples*M3O_MAX_DIM]*/,
                                           const double* edges /*concatenated per sample per dim*/,
                                           int test_statistic, int update_w2) {
  return sample_create(nEvents, nSamples, nDim, NULL, nbins, edges, test_statistic, update_w2);
}
M3O_API void* m3o_sample_create_ex(unsigned int nEvents, int nSamples, const int* nDim, const int* uniform,
                                   const int* nbins, const double* edges, int test_statistic, int update_w2) {
  return sample_create(nEvents, nSamples, nDim, uniform, nbins, edges, test_statistic, update_w2);
}
static SampleHandlerFD* sample_create(unsigned int nEvents, int nSamples, const int* nDim, const int* uniform,
                                      const int* nbins, const double* edges, int test_statistic, int update_w2) {
  SampleHandlerFD* s = (SampleHandlerFD*)calloc(1, sizeof(SampleHandlerFD));
  s->nEvents = nEvents; s->nSamples = nSamples;
  s->SampleBinning = (SampleBinningInfo*)calloc((size_t)nSamples, sizeof(SampleBinningInfo));
  const double* ep = edges;
  int GlobalOffsetCounter = 0;                         /* BinningHandler::SetGlobalBinNumbers (:341-355) */
  for (int i = 0; i < nSamples; ++i) {
    SampleBinningInfo* b = &s->SampleBinning[i];
    b->nDim = nDim[i];
    b->Uniform = 1;
    if (uniform && !uniform[i]) {
      const int nBoxes = nbins[i * M3O_MAX_DIM];
      InitNonUniform(b, nDim[i], nBoxes, ep);
      ep += (size_t)nBoxes * (size_t)nDim[i] * 2;
      b->GlobalOffset = GlobalOffsetCounter;
      GlobalOffsetCounter += nBoxes;
      continue;
    }
    int stride = 1, tot = 1;
    for (int d = 0; d < b->nDim; ++d) {
      const int nb = nbins[i * M3O_MAX_DIM + d];
      b->AxisNBins[d] = nb;
      b->BinEdges[d] = (double*)malloc(sizeof(double) * (size_t)(nb + 1));
      memcpy(b->BinEdges[d], ep, sizeof(double) * (size_t)(nb + 1));
      ep += nb + 1;
      b->BinLookup[d] = (BinShiftLookup*)malloc(sizeof(BinShiftLookup) * (size_t)nb);
      InitialiseLookUpSingleDimension(b->BinLookup[d], b->BinEdges[d], nb);
      b->Strides[d] = stride;                          /* InitialiseStrides (SampleStructs.h:656-664) */
      stride *= nb; tot *= nb;
    }
    b->nBins = tot;
    b->GlobalOffset = GlobalOffsetCounter;
    GlobalOffsetCounter += tot;
  }
  s->TotalBins = GlobalOffsetCounter;
  /* SampleHandlerFD::SetupReweightArrays (Samples/SampleHandlerFD.cpp:749-754) */
  s->SampleHandlerFD_array = (double*)calloc((size_t)s->TotalBins, sizeof(double));
  s->SampleHandlerFD_array_w2 = (double*)calloc((size_t)s->TotalBins, sizeof(double));
  s->SampleHandlerFD_data = (double*)calloc((size_t)s->TotalBins, sizeof(double));
  s->MCSamples = (EventInfo*)calloc((size_t)nEvents, sizeof(EventInfo));
  s->fTestStatistic = test_statistic;
  s->FirstTimeW2 = 1;
  s->UpdateW2 = update_w2;
  return s;
}

M3O_API void m3o_sample_destroy(SampleHandlerFD* s) {
  if (!s) return;
  for (unsigned int e = 0; e < s->nEvents; ++e) {
    free(s->MCSamples[e].norm_pointers); free(s->MCSamples[e].total_weight_pointers);
    free(s->MCSamples[e].KinVar); free(s->MCSamples[e].NomBin);
  }
  free(s->MCSamples);
  if (s->tw_d) { for (unsigned int e = 0; e < s->nEvents; ++e) free((void*)s->tw_d[e]); free((void*)s->tw_d); free(s->n_tw_d); }
  for (int i = 0; i < s->nSamples; ++i)
  {
    for (int d = 0; d < s->SampleBinning[i].nDim; ++d) { free(s->SampleBinning[i].BinEdges[d]); free(s->SampleBinning[i].BinLookup[d]); }
    free(s->SampleBinning[i].Extent); free(s->SampleBinning[i].GridStart); free(s->SampleBinning[i].GridIdx);
  }
  free(s->SampleBinning);
  free(s->SampleHandlerFD_array); free(s->SampleHandlerFD_array_w2); free(s->SampleHandlerFD_data);
  free(s->SelStart); free(s->StoredSelection); free(s->Selection);
  free(s->ShiftStart); free(s->ShiftPar); free(s->ShiftTarget); free(s->ShiftCoef); free(s->KinNominal); free(s->CutNominal);
  free(s);
}

/* Wires the per-event pointer vectors the way SampleHandlerFD::Initialise does
 * (Samples/SampleHandlerFD.cpp:169-202):
 *   KinVar/NomBin    FindNominalBinAndEdges (:858-887)
 *   norm_pointers    SetupNormParameters    (:637-663)   -> &norm_base[idx]
 *   total_weight_pointers, in push order:
 *       oscillation weight  SetupNuOscillatorPointers (:1108-1122) -> &osc_base[osc_idx or e]
 *       spline weight       SetSplinePointers         (:1244-1249) -> SMonolith::retPointer(e)
 *       extra weight        AddAdditionalWeightPointers (experiment) -> &static_w[e]
 * kin is dim-major: kin[d*nEvents + e].  norm_idx holds n_norm_per_event entries per event,
 * a negative entry = no pointer.  Any base may be NULL = that weight is absent. */
M3O_API void m3o_sample_set_events(SampleHandlerFD* s, const int* sample_id, const double* kin,
                                   int n_norm_per_event, const short* norm_idx, const double* norm_base,
                                   const float* osc_base, const int* osc_idx,
                                   SMonolith* spline, const float* static_w) {
  s->SplineHandler = spline;
  for (unsigned int e = 0; e < s->nEvents; ++e) {
    EventInfo* ev = &s->MCSamples[e];
    ev->NominalSample = sample_id[e];
    const SampleBinningInfo* b = &s->SampleBinning[ev->NominalSample];
    ev->n_dim = b->nDim;
    ev->KinVar = (const double**)malloc(sizeof(double*) * (size_t)b->nDim);
    ev->NomBin = (int*)malloc(sizeof(int) * (size_t)b->nDim);
    for (int d = 0; d < b->nDim; ++d) {
      ev->KinVar[d] = &kin[(size_t)d * s->nEvents + e];
      const int bin = FindNominalBin(b, d, *ev->KinVar[d]);
      ev->NomBin[d] = (bin >= 0 && bin < b->AxisNBins[d]) ? bin : UnderOverFlowBin;
    }
    int nn = 0;
    for (int j = 0; j < n_norm_per_event; ++j) if (norm_base && norm_idx[(size_t)e * n_norm_per_event + j] >= 0) ++nn;
    ev->n_norm = nn;
    ev->norm_pointers = (const double**)malloc(sizeof(double*) * (size_t)(nn > 0 ? nn : 1));
    nn = 0;
    for (int j = 0; j < n_norm_per_event; ++j) {
      if (!norm_base) break;
      const short idx = norm_idx[(size_t)e * n_norm_per_event + j];
      if (idx >= 0) ev->norm_pointers[nn++] = &norm_base[idx];
    }
    int nt = (osc_base ? 1 : 0) + (spline ? 1 : 0) + (static_w ? 1 : 0);
    ev->n_tw = nt;
    ev->total_weight_pointers = (const float**)malloc(sizeof(float*) * (size_t)(nt > 0 ? nt : 1));
    nt = 0;
    if (osc_base) ev->total_weight_pointers[nt++] = &osc_base[osc_idx ? (size_t)osc_idx[e] : (size_t)e];
    if (spline)   ev->total_weight_pointers[nt++] = &spline->cpu_total_weights[e];
    if (static_w) ev->total_weight_pointers[nt++] = &static_w[e];
  }
}


/* The same wiring with a BinnedSplineHandler (Samples/SampleHandlerFD.cpp:1196-1242): after the
 * oscillation pointer every event gets one pointer per binned spline that applies to it --
 * BinnedSplineHandler::retPointer(...) = &weightvec_Monolith[index] -- in the order GetEventSplines
 * returned them; then the experiment's extra weights.  n_per_event[e] pointers, indices concatenated. */
M3O_API void m3o_sample_set_events_binned(SampleHandlerFD* s, const int* sample_id, const double* kin,
                                          int n_norm_per_event, const short* norm_idx, const double* norm_base,
                                          const float* osc_base, const int* osc_idx,
                                          BinnedSplineHandler* binned, const uint32_t* n_per_event, const int* spline_index,
                                          const float* static_w) {
  m3o_sample_set_events(s, sample_id, kin, n_norm_per_event, norm_idx, norm_base, osc_base, osc_idx, NULL, static_w);
  s->BinnedHandler = binned;
  uint64_t off = 0;
  for (unsigned int e = 0; e < s->nEvents; ++e) {
    EventInfo* ev = &s->MCSamples[e];
    const int n_old = ev->n_tw, n_b = (int)n_per_event[e];
    const float** tw = (const float**)malloc(sizeof(float*) * (size_t)(n_old + n_b > 0 ? n_old + n_b : 1));
    int nt = 0, k = 0;
    if (osc_base) tw[nt++] = ev->total_weight_pointers[k++];
    for (int j = 0; j < n_b; ++j) tw[nt++] = &binned->weightvec_Monolith[spline_index[off + (uint64_t)j]];
    while (k < n_old) tw[nt++] = ev->total_weight_pointers[k++];
    free(ev->total_weight_pointers);
    ev->total_weight_pointers = tw;
    ev->n_tw = nt;
    off += (uint64_t)n_b;
  }
}

/* StoredSelection as the experiment's YAML fills it (Samples/SampleHandlerFD.cpp:140-165): n_cuts cuts
 * {sample, ParamToCutOnIt, LowerBound, UpperBound}, kept in the given order inside each sample.  values[var*nEvents+e]
 * stands in for ReturnKinematicParameter(var, e); the caller owns it (and may rewrite it: functional shifts). */
M3O_API void m3o_sample_set_selection(SampleHandlerFD* s, int n_cuts, const int* cut_sample, const int* cut_var,
                                      const double* lower, const double* upper, const double* values) {
  free(s->SelStart); free(s->StoredSelection); free(s->Selection);
  s->SelStart = (int*)calloc((size_t)s->nSamples + 1, sizeof(int));
  s->StoredSelection = (KinematicCut*)calloc((size_t)(n_cuts > 0 ? n_cuts : 1), sizeof(KinematicCut));
  s->Selection = (KinematicCut*)calloc((size_t)(n_cuts > 0 ? n_cuts : 1), sizeof(KinematicCut));
  for (int k = 0; k < n_cuts; ++k) ++s->SelStart[cut_sample[k] + 1];
  for (int i = 0; i < s->nSamples; ++i) s->SelStart[i + 1] += s->SelStart[i];
  int* fill = (int*)malloc(sizeof(int) * (size_t)s->nSamples);
  for (int i = 0; i < s->nSamples; ++i) fill[i] = s->SelStart[i];
  for (int k = 0; k < n_cuts; ++k) {
    KinematicCut* c = &s->StoredSelection[fill[cut_sample[k]]++];
    c->ParamToCutOnIt = cut_var[k]; c->LowerBound = lower[k]; c->UpperBound = upper[k];
  }
  free(fill);
  s->nCuts = n_cuts;
  s->CutValues = values;
}

/* SampleHandlerFD::IsEventSelected (Samples/SampleHandlerFD.cpp:281-294) */
static inline int IsEventSelected(const SampleHandlerFD* s, const int iSample, const unsigned int iEvent) {
  if (!s->SelStart) return 1;
  for (int iSelection = s->SelStart[iSample]; iSelection < s->SelStart[iSample + 1]; ++iSelection) {
    const KinematicCut* Cut = &s->Selection[iSelection];
    const double Val = s->CutValues[(size_t)Cut->ParamToCutOnIt * s->nEvents + iEvent];   /* ReturnKinematicParameter */
    if ((Val < Cut->LowerBound) || (Val >= Cut->UpperBound)) return 0;
  }
  return 1;
}
M3O_API void m3o_event_selected(const SampleHandlerFD* s, unsigned char* out) {
  if (s->SelStart) memcpy(s->Selection, s->StoredSelection, sizeof(KinematicCut) * (size_t)s->nCuts);
  for (unsigned int e = 0; e < s->nEvents; ++e) out[e] = (unsigned char)IsEventSelected(s, s->MCSamples[e].NominalSample, e);
}

/* funcParsGrid for linear shifts.  kin[n_kin_rows*nEvents] is the live array the KinVar pointers look into (the one
 * handed to m3o_sample_set_events), cut_values the live table of m3o_sample_set_selection (or NULL); both are read now
 * as the nominal values.  Entry k of event e: *values[par[k]] * coef[k] is added to target[k] (< n_kin_rows: that
 * kinematic row; else cut-variable row target - n_kin_rows). */
M3O_API void m3o_sample_set_linear_shifts(SampleHandlerFD* s, const uint32_t* n_per_event, const int* par, const int* target,
                                          const double* coef, const double* values, double* kin, int n_kin_rows,
                                          double* cut_values, int n_cut_rows) {
  free(s->ShiftStart); free(s->ShiftPar); free(s->ShiftTarget); free(s->ShiftCoef); free(s->KinNominal); free(s->CutNominal);
  s->ShiftStart = (int64_t*)calloc((size_t)s->nEvents + 1, sizeof(int64_t));
  for (unsigned int e = 0; e < s->nEvents; ++e) s->ShiftStart[e + 1] = s->ShiftStart[e] + n_per_event[e];
  const int64_t tot = s->ShiftStart[s->nEvents];
  s->ShiftPar = (int*)malloc(sizeof(int) * (size_t)(tot > 0 ? tot : 1));
  s->ShiftTarget = (int*)malloc(sizeof(int) * (size_t)(tot > 0 ? tot : 1));
  s->ShiftCoef = (double*)malloc(sizeof(double) * (size_t)(tot > 0 ? tot : 1));
  memcpy(s->ShiftPar, par, sizeof(int) * (size_t)tot); memcpy(s->ShiftTarget, target, sizeof(int) * (size_t)tot);
  memcpy(s->ShiftCoef, coef, sizeof(double) * (size_t)tot);
  s->ShiftValues = values;
  s->KinShifted = kin; s->nKinRows = n_kin_rows;
  s->KinNominal = (double*)malloc(sizeof(double) * (size_t)n_kin_rows * s->nEvents);
  memcpy(s->KinNominal, kin, sizeof(double) * (size_t)n_kin_rows * s->nEvents);
  s->CutShifted = cut_values; s->nCutRows = cut_values ? n_cut_rows : 0;
  s->CutNominal = NULL;
  if (cut_values) {
    s->CutNominal = (double*)malloc(sizeof(double) * (size_t)n_cut_rows * s->nEvents);
    memcpy(s->CutNominal, cut_values, sizeof(double) * (size_t)n_cut_rows * s->nEvents);
  }
}

/* SampleHandlerFD::ApplyShifts (Samples/SampleHandlerFD.cpp:545-564) */
static inline void ApplyShifts(SampleHandlerFD* s, const unsigned int iEvent) {
  if (!s->ShiftStart) return;
  const int64_t k0 = s->ShiftStart[iEvent], nShifts = s->ShiftStart[iEvent + 1] - k0;
  if (nShifts == 0) return;                                       /* :549-552: nothing to reset either */
  /* ResetShifts(iEvent): back to the nominal values */
  for (int d = 0; d < s->nKinRows; ++d) s->KinShifted[(size_t)d * s->nEvents + iEvent] = s->KinNominal[(size_t)d * s->nEvents + iEvent];
  for (int v = 0; v < s->nCutRows; ++v) s->CutShifted[(size_t)v * s->nEvents + iEvent] = s->CutNominal[(size_t)v * s->nEvents + iEvent];
  for (int64_t iShift = 0; iShift < nShifts; ++iShift) {           /* (*fp->funcPtr)(fp->valuePtr, iEvent) */
    const int t = s->ShiftTarget[k0 + iShift];
    double* x = t < s->nKinRows ? &s->KinShifted[(size_t)t * s->nEvents + iEvent] : &s->CutShifted[(size_t)(t - s->nKinRows) * s->nEvents + iEvent];
    const double delta = s->ShiftValues[s->ShiftPar[k0 + iShift]] * s->ShiftCoef[k0 + iShift];
    *x = *x + delta;
  }
  /* FinaliseShifts(iEvent): nothing to do for this family */
}

/* SampleHandlerFD::CalcWeightTotal (Samples/SampleHandlerFD.cpp:568-594) */
static inline float CalcWeightTotal(const EventInfo* restrict MCEvent) {
  float TotalWeight = 1.0;
  const int nNorms = MCEvent->n_norm;
  #pragma omp simd reduction(*:TotalWeight)
  for (int iParam = 0; iParam < nNorms; ++iParam) TotalWeight *= (float)(*(MCEvent->norm_pointers[iParam]));
  const int TotalWeights = MCEvent->n_tw;
  #pragma omp simd reduction(*:TotalWeight)
  for (int iWeight = 0; iWeight < TotalWeights; ++iWeight) TotalWeight *= *(MCEvent->total_weight_pointers[iWeight]);
  return TotalWeight;
}

/* SampleHandlerFD::CalcWeightTotal, serial build (no omp simd) */
static float CalcWeightTotal_serial(const EventInfo* restrict MCEvent) {
  float TotalWeight = 1.0;
  for (int iParam = 0; iParam < MCEvent->n_norm; ++iParam) TotalWeight *= (float)(*(MCEvent->norm_pointers[iParam]));
  for (int iWeight = 0; iWeight < MCEvent->n_tw; ++iWeight) TotalWeight *= *(MCEvent->total_weight_pointers[iWeight]);
  return TotalWeight;
}

/* BinningHandler::FindGlobalBin, uniform binning arm (Samples/BinningHandler.cpp:257-277) */
static inline int FindGlobalBin(const SampleHandlerFD* s, const int NomSample, const double* const* KinVar, const int* NomBin, int Dim) {
  const SampleBinningInfo* restrict Binning = &s->SampleBinning[NomSample];
  int GlobalBin = 0;
  for (int i = 0; i < Dim; ++i) {
    const double Var = *KinVar[i];
    const int Bin = FindBin(Var, NomBin[i], Binning->AxisNBins[i], Binning->BinEdges[i], Binning->BinLookup[i]);
    if (Bin < 0) return UnderOverFlowBin;
    GlobalBin += Bin * Binning->Strides[i];
  }
  if (Binning->Uniform) {
    GlobalBin += Binning->GlobalOffset;
    return GlobalBin;
  }
  /* non-uniform arm (Samples/BinningHandler.cpp:278-290): scan the mega bin's boxes in order; BinInfo::IsEventInside
   * tests (lo, hi] in every dimension (Samples/SampleStructs.h:207-219) */
  for (int k = Binning->GridStart[GlobalBin]; k < Binning->GridStart[GlobalBin + 1]; ++k) {
    const int BinNumber = Binning->GridIdx[k];
    int inside = 1;
    for (int i = 0; i < Dim; ++i) {
      const double Var = *KinVar[i];
      const double lo = Binning->Extent[((size_t)BinNumber * Dim + i) * 2], hi = Binning->Extent[((size_t)BinNumber * Dim + i) * 2 + 1];
      inside &= (Var > lo) & (Var <= hi);
    }
    if (inside) return BinNumber + Binning->GlobalOffset;
  }
  return UnderOverFlowBin;
}

/* SampleHandlerFD::ResetHistograms (Samples/SampleHandlerFD.cpp:454-463) */
static void ResetHistograms(SampleHandlerFD* s) {
  for (int i = 0; i < s->TotalBins; ++i) s->SampleHandlerFD_array[i] = 0.0;
  if (s->FirstTimeW2) for (int i = 0; i < s->TotalBins; ++i) s->SampleHandlerFD_array_w2[i] = 0.0;
}

/* SampleHandlerFD::FillArray_MP (Samples/SampleHandlerFD.cpp:390-448); functional shifts are the caller's (it
 * rewrites the kinematic / cut-variable arrays the pointers look at), no CalcWeightFunc (default: does nothing,
 * Samples/SampleHandlerFD.h:241,289) */
static void FillArray_MP(SampleHandlerFD* s) {
  if (s->SelStart) memcpy(s->Selection, s->StoredSelection, sizeof(KinematicCut) * (size_t)s->nCuts);   /* Selection = StoredSelection, :393 */
  const int TotalBins = s->TotalBins;
  const unsigned int NumberOfEvents = s->nEvents;
  double* MC_Array_for_reduction = s->SampleHandlerFD_array;
  double* W2_array_for_reduction = s->SampleHandlerFD_array_w2;
  const int FirstTimeW2 = s->FirstTimeW2;
  #pragma omp parallel for reduction(+:MC_Array_for_reduction[:TotalBins], W2_array_for_reduction[:TotalBins])
  for (unsigned int iEvent = 0; iEvent < NumberOfEvents; ++iEvent) {
    ApplyShifts(s, iEvent);                                                               /* :420 */
    const EventInfo* restrict MCEvent = &s->MCSamples[iEvent];
    if (!IsEventSelected(s, MCEvent->NominalSample, iEvent)) continue;                    /* :424 */
    const float totalweight = CalcWeightTotal(MCEvent);
    if (totalweight <= 0.) continue;                                                      /* :432 */
    const int GlobalBin = FindGlobalBin(s, MCEvent->NominalSample, MCEvent->KinVar, MCEvent->NomBin, MCEvent->n_dim);
    if (GlobalBin > UnderOverFlowBin) {                                                   /* :443 */
      MC_Array_for_reduction[GlobalBin] += totalweight;
      if (FirstTimeW2) W2_array_for_reduction[GlobalBin] += totalweight * totalweight;    /* float product, :445 */
    }
  }
}

/* SampleHandlerFD::FillArray (Samples/SampleHandlerFD.cpp:352-383), the serial build's fill */
static void FillArray(SampleHandlerFD* s) {
  if (s->SelStart) memcpy(s->Selection, s->StoredSelection, sizeof(KinematicCut) * (size_t)s->nCuts);   /* :355 */
  for (unsigned int iEvent = 0; iEvent < s->nEvents; iEvent++) {
    ApplyShifts(s, iEvent);                                                               /* :358 */
    const EventInfo* restrict MCEvent = &s->MCSamples[iEvent];
    if (!IsEventSelected(s, MCEvent->NominalSample, iEvent)) continue;                    /* :361 */
    const float totalweight = CalcWeightTotal_serial(MCEvent);
    if (totalweight <= 0.) continue;
    const int GlobalBin = FindGlobalBin(s, MCEvent->NominalSample, MCEvent->KinVar, MCEvent->NomBin, MCEvent->n_dim);
    if (GlobalBin > UnderOverFlowBin) {
      s->SampleHandlerFD_array[GlobalBin] += totalweight;
      if (s->FirstTimeW2) s->SampleHandlerFD_array_w2[GlobalBin] += totalweight * totalweight;
    }
  }
}

/* SampleHandlerFD::Reweight (Samples/SampleHandlerFD.cpp:316-343).  The oscillator is an
 * input array here (north_star), so Oscillator->Evaluate() is the caller's job. */
M3O_API void m3o_reweight(SampleHandlerFD* s) {
  ResetHistograms(s);
  if (s->SplineHandler) m3o_evaluate(s->SplineHandler);
  if (s->BinnedHandler) m3o_binned_evaluate(s->BinnedHandler);
  if (g_multithread) FillArray_MP(s); else FillArray(s);
  if (!s->UpdateW2) s->FirstTimeW2 = 0;
}
/* FillArray part alone (spline weights already evaluated) -- for DragRace-style split timing */
M3O_API void m3o_fill_only(SampleHandlerFD* s) {
  ResetHistograms(s);
  if (g_multithread) FillArray_MP(s); else FillArray(s);
  if (!s->UpdateW2) s->FirstTimeW2 = 0;
}


/* Default-build wiring: total_weight_pointers are const double* -- oscillation weight, the binned-spline weights
 * (BinnedSplineHandler::retPointer), the extra weight, in that order (see m3o_sample_set_events_binned). */
M3O_API void m3o_sample_set_events_binned_d(SampleHandlerFD* s, const int* sample_id, const double* kin,
                                            int n_norm_per_event, const short* norm_idx, const double* norm_base,
                                            const double* osc_base, struct BinnedSplineHandlerD_* binned,
                                            const uint32_t* n_per_event, const int* spline_index, const double* static_w) {
  m3o_sample_set_events(s, sample_id, kin, n_norm_per_event, norm_idx, norm_base, NULL, NULL, NULL, NULL);
  s->BinnedHandlerD = binned;
  s->tw_d = (const double***)calloc((size_t)s->nEvents, sizeof(double**));
  s->n_tw_d = (int*)calloc((size_t)s->nEvents, sizeof(int));
  uint64_t off = 0;
  for (unsigned int e = 0; e < s->nEvents; ++e) {
    const int n_b = (int)n_per_event[e];
    const double** tw = (const double**)malloc(sizeof(double*) * (size_t)(n_b + 2));
    int nt = 0;
    if (osc_base) tw[nt++] = &osc_base[e];
    for (int j = 0; j < n_b; ++j) tw[nt++] = &binned->weightvec_Monolith[spline_index[off + (uint64_t)j]];
    if (static_w) tw[nt++] = &static_w[e];
    s->tw_d[e] = tw; s->n_tw_d[e] = nt;
    off += (uint64_t)n_b;
  }
}

/* SampleHandlerFD::CalcWeightTotal (Samples/SampleHandlerFD.cpp:568-594) with M3::float_t = double */
static inline double CalcWeightTotal_d(const SampleHandlerFD* s, unsigned int e, int simd) {
  const EventInfo* MCEvent = &s->MCSamples[e];
  double TotalWeight = 1.0;
  if (simd) {
    #pragma omp simd reduction(*:TotalWeight)
    for (int iParam = 0; iParam < MCEvent->n_norm; ++iParam) TotalWeight *= (double)(*(MCEvent->norm_pointers[iParam]));
    #pragma omp simd reduction(*:TotalWeight)
    for (int iWeight = 0; iWeight < s->n_tw_d[e]; ++iWeight) TotalWeight *= *(s->tw_d[e][iWeight]);
  } else {
    for (int iParam = 0; iParam < MCEvent->n_norm; ++iParam) TotalWeight *= (double)(*(MCEvent->norm_pointers[iParam]));
    for (int iWeight = 0; iWeight < s->n_tw_d[e]; ++iWeight) TotalWeight *= *(s->tw_d[e][iWeight]);
  }
  return TotalWeight;
}

/* SampleHandlerFD::Reweight, default build, binned splines */
M3O_API void m3o_reweight_d(SampleHandlerFD* s) {
  ResetHistograms(s);
  if (s->BinnedHandlerD) m3o_binnedd_evaluate(s->BinnedHandlerD);
  if (s->SelStart) memcpy(s->Selection, s->StoredSelection, sizeof(KinematicCut) * (size_t)s->nCuts);
  const int FirstTimeW2 = s->FirstTimeW2;
  if (g_multithread) {
    const int TotalBins = s->TotalBins;
    double* MC = s->SampleHandlerFD_array; double* W2 = s->SampleHandlerFD_array_w2;
    #pragma omp parallel for reduction(+:MC[:TotalBins], W2[:TotalBins])
    for (unsigned int iEvent = 0; iEvent < s->nEvents; ++iEvent) {
      const EventInfo* MCEvent = &s->MCSamples[iEvent];
      if (!IsEventSelected(s, MCEvent->NominalSample, iEvent)) continue;
      const double totalweight = CalcWeightTotal_d(s, iEvent, 1);
      if (totalweight <= 0.) continue;
      const int GlobalBin = FindGlobalBin(s, MCEvent->NominalSample, MCEvent->KinVar, MCEvent->NomBin, MCEvent->n_dim);
      if (GlobalBin > UnderOverFlowBin) { MC[GlobalBin] += totalweight; if (FirstTimeW2) W2[GlobalBin] += totalweight * totalweight; }
    }
  } else {
    for (unsigned int iEvent = 0; iEvent < s->nEvents; ++iEvent) {
      const EventInfo* MCEvent = &s->MCSamples[iEvent];
      if (!IsEventSelected(s, MCEvent->NominalSample, iEvent)) continue;
      const double totalweight = CalcWeightTotal_d(s, iEvent, 0);
      if (totalweight <= 0.) continue;
      const int GlobalBin = FindGlobalBin(s, MCEvent->NominalSample, MCEvent->KinVar, MCEvent->NomBin, MCEvent->n_dim);
      if (GlobalBin > UnderOverFlowBin) {
        s->SampleHandlerFD_array[GlobalBin] += totalweight;
        if (FirstTimeW2) s->SampleHandlerFD_array_w2[GlobalBin] += totalweight * totalweight;
      }
    }
  }
  if (!s->UpdateW2) s->FirstTimeW2 = 0;
}
M3O_API void m3o_event_weights_d(const SampleHandlerFD* s, double* out) {
  for (unsigned int e = 0; e < s->nEvents; ++e) out[e] = CalcWeightTotal_d(s, e, 0);
}

/* SampleHandlerBase::GetPoissonLLH (Samples/SampleHandlerBase.cpp:17-31) */
static double GetPoissonLLH(const double data, const double mc) {
  if (data == 0) return mc;
  if (mc < LOW_MC_BOUND) {
    if (data > LOW_MC_BOUND) return (LOW_MC_BOUND - data + data * log(data / LOW_MC_BOUND));
    else if (data >= mc) return 0.;
  }
  return (mc - data + data * log(data / mc));
}

/* SampleHandlerBase::GetTestStatLLH (Samples/SampleHandlerBase.cpp:35-192) */
static double GetTestStatLLH(const int fTestStatistic, const double data, const double mc, const double w2) {
  switch (fTestStatistic) {
    case kBarlowBeeston: {                                                  /* :46-88 */
      double newmc = mc;
      if (mc < LOW_MC_BOUND) {
        if (data > LOW_MC_BOUND) newmc = LOW_MC_BOUND;
        else if (data >= mc) return 0.;
      }
      const double fractional = sqrt(w2) / newmc;
      const double fractional2 = fractional * fractional;
      const double temp = newmc * fractional2 - 1;
      const double temp2 = temp * temp + 4 * data * fractional2;
      if (temp2 < 0) return NAN;            /* the reference throws here (:65-68) */
      const double beta = (-1 * temp + sqrt(temp2)) / 2.;
      double stat = mc * beta;
      if (data > 0) {
        newmc *= beta;
        stat = newmc - data + data * log(data / newmc);
      }
      double penalty = 0;
      if (fractional > 0) penalty = (beta - 1) * (beta - 1) / (2 * fractional2);
      return stat + penalty;
    }
    case kDembinskiAbdelmotteleb: {                                         /* :90-126 */
      if (w2 == 0) return GetPoissonLLH(data, mc);
      double newmc = mc;
      if (mc < LOW_MC_BOUND) {
        if (data > LOW_MC_BOUND) newmc = LOW_MC_BOUND;
        else if (data >= mc) return 0.;
      }
      const double k = newmc * newmc / w2;
      const double beta = (data + k) / (newmc + k);
      newmc *= beta;
      const double penalty = k * beta - k + k * log(k / (k * beta));
      double stat = newmc;
      if (data > 0) stat = newmc - data + data * log(data / newmc);
      return stat + penalty;
    }
    case kIceCube: {                                                        /* :133-160 */
      if (w2 == 0) return GetPoissonLLH(data, mc);
      const long double b = mc / w2;
      const long double a = mc * b + 1;
      const double stat = (double)(-1 * (a * logl(b) + lgammal(data + a) - lgammal(data + 1) - ((data + a) * log1pl(b)) - lgammal(a)));
      if (mc <= data) {
        if (data <= LOW_MC_BOUND) return 0.;
        const double poisson = GetPoissonLLH(data, LOW_MC_BOUND);
        if (stat > poisson) return poisson;
      }
      return stat;
    }
    case kPearson: {                                                        /* :162-177 */
      if (data == 0) return mc / 2.;
      if (mc < LOW_MC_BOUND) {
        if (data > LOW_MC_BOUND) return (data - LOW_MC_BOUND) * (data - LOW_MC_BOUND) / (2. * LOW_MC_BOUND);
        else if (data >= mc) return 0.;
      }
      return (data - mc) * (data - mc) / (2 * mc);
    }
    case kPoisson:                                                          /* :178-184 */
      return GetPoissonLLH(data, mc);
    default:
      return NAN;
  }
}
M3O_API double m3o_test_stat_llh(int test_statistic, double data, double mc, double w2) {
  return GetTestStatLLH(test_statistic, data, mc, w2);
}

/* SampleHandlerFD::GetLikelihood (Samples/SampleHandlerFD.cpp:1284-1300) */
M3O_API double m3o_get_likelihood(const SampleHandlerFD* s) {
  double negLogL = 0.;
  #pragma omp parallel for reduction(+:negLogL)
  for (int idx = 0; idx < s->TotalBins; ++idx)
    negLogL += GetTestStatLLH(s->fTestStatistic, s->SampleHandlerFD_data[idx], s->SampleHandlerFD_array[idx], s->SampleHandlerFD_array_w2[idx]);
  return negLogL;
}
/* SampleHandlerFD::GetSampleLikelihood (Samples/SampleHandlerFD.cpp:1262-1281) */
M3O_API double m3o_get_sample_likelihood(const SampleHandlerFD* s, int isample) {
  const int Start = s->SampleBinning[isample].GlobalOffset;
  const int End = isample + 1 < s->nSamples ? s->SampleBinning[isample + 1].GlobalOffset : s->TotalBins;
  double negLogL = 0.;
  #pragma omp parallel for reduction(+:negLogL)
  for (int idx = Start; idx < End; ++idx)
    negLogL += GetTestStatLLH(s->fTestStatistic, s->SampleHandlerFD_data[idx], s->SampleHandlerFD_array[idx], s->SampleHandlerFD_array_w2[idx]);
  return negLogL;
}

M3O_API int m3o_total_bins(const SampleHandlerFD* s) { return s->TotalBins; }
M3O_API double* m3o_mc_array(SampleHandlerFD* s) { return s->SampleHandlerFD_array; }
M3O_API double* m3o_w2_array(SampleHandlerFD* s) { return s->SampleHandlerFD_array_w2; }
M3O_API double* m3o_data_array(SampleHandlerFD* s) { return s->SampleHandlerFD_data; }
M3O_API void m3o_set_test_statistic(SampleHandlerFD* s, int t) { s->fTestStatistic = t; }
M3O_API void m3o_set_first_time_w2(SampleHandlerFD* s, int v) { s->FirstTimeW2 = v; }
/* SampleHandlerFD::AddData (Samples/SampleHandlerFD.cpp:955-1044), array form */
M3O_API void m3o_add_data(SampleHandlerFD* s, const double* data) { memcpy(s->SampleHandlerFD_data, data, sizeof(double) * (size_t)s->TotalBins); }

/* per-event views used by the bit-exact index parity tests */
M3O_API void m3o_event_bins(const SampleHandlerFD* s, int* out) {
  #pragma omp parallel for
  for (unsigned int e = 0; e < s->nEvents; ++e) {
    const EventInfo* ev = &s->MCSamples[e];
    out[e] = FindGlobalBin(s, ev->NominalSample, ev->KinVar, ev->NomBin, ev->n_dim);
  }
}
M3O_API void m3o_event_weights(const SampleHandlerFD* s, float* out) {
  #pragma omp parallel for
  for (unsigned int e = 0; e < s->nEvents; ++e)
    out[e] = g_multithread ? CalcWeightTotal(&s->MCSamples[e]) : CalcWeightTotal_serial(&s->MCSamples[e]);
}
/* single-value bin lookup for the FindBin known-answer tests */
/* accessors used to pin the binning restatement against the reference's own SampleBinningInfo (oracle/ref_host) */
M3O_API double m3o_bin_edge(const SampleHandlerFD* s, int sample, int dim, int i) { return s->SampleBinning[sample].BinEdges[dim][i]; }
M3O_API int m3o_axis_nbins(const SampleHandlerFD* s, int sample, int dim) { return s->SampleBinning[sample].AxisNBins[dim]; }
M3O_API int m3o_grid_size(const SampleHandlerFD* s, int sample, int mega) {
  const SampleBinningInfo* b = &s->SampleBinning[sample];
  return b->Uniform ? 0 : b->GridStart[mega + 1] - b->GridStart[mega];
}
M3O_API int m3o_grid_entry(const SampleHandlerFD* s, int sample, int mega, int k) {
  const SampleBinningInfo* b = &s->SampleBinning[sample];
  return b->GridIdx[b->GridStart[mega] + k];
}
M3O_API int m3o_find_bin(const SampleHandlerFD* s, int sample, int dim, double var, int nom_bin) {
  const SampleBinningInfo* b = &s->SampleBinning[sample];
  return FindBin(var, nom_bin, b->AxisNBins[dim], b->BinEdges[dim], b->BinLookup[dim]);
}
M3O_API int m3o_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ============================================================================================
 * BinnedSplineHandler::CalcSplineWeights (Splines/BinnedSplineHandler.cpp:306-341), in the
 * reference's default build (M3::float_t = double).  xvar is read un-narrowed through the
 * parameter pointer (:327), x is stored per spline (:329), negative weights clamp to 0 (:337).
 * ========================================================================================== */
M3O_API void m3o_binned_calc_spline_weights(int64_t n_unique, const int* uniquecoeffindices,
                                            const short* uniquesplinevec_Monolith, const short* SplineSegments,
                                            const int* coeffindexvec, const double* manycoeff_arr,
                                            const double* xcoeff_arr, const double* const* splineParsPointer,
                                            double* weightvec_Monolith) {
  #pragma omp parallel for simd
  for (int64_t iCoeff = 0; iCoeff < n_unique; ++iCoeff) {
    const int iSpline = uniquecoeffindices[iCoeff];
    const short uniqueIndex = (short)uniquesplinevec_Monolith[iSpline];
    const short currentsegment = (short)SplineSegments[uniqueIndex];
    const int segCoeff = coeffindexvec[iSpline] + currentsegment;
    const int coeffOffset = segCoeff * nCoeff;
    const double y = manycoeff_arr[coeffOffset + 0];
    const double b = manycoeff_arr[coeffOffset + 1];
    const double c = manycoeff_arr[coeffOffset + 2];
    const double d = manycoeff_arr[coeffOffset + 3];
    const double xvar = *splineParsPointer[uniqueIndex];
    const double dx = xvar - xcoeff_arr[segCoeff];
    double weight = fma(dx, fma(dx, fma(dx, d, c), b), y);
    if (weight < 0) weight = 0.;
    weightvec_Monolith[iSpline] = weight;
  }
}
