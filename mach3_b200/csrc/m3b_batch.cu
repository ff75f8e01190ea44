// m3b_batch.cu -- batched proposals (BASELINE config 5): up to 256 parameter sets against the same events
// in ONE pass over the coefficient rows.
//
// The single-set kernel is HBM-bound; a batch re-uses every coefficient row for all sets, so the work flips
// to FP32 issue.  The mapping is transposed with respect to fill_tma_kernel:
//
//     tile    = 32 consecutive events, whose coefficient rows -- for every segment that ANY set of the batch
//               selects (1-3 per parameter for proposals around one point) -- a producer warp stages in
//               shared memory with 1-D bulk copies (multi-buffered);
//     thread  = one EVENT of the tile (lane) x 16 parameter SETS (warp w owns sets 16w..16w+15), so a
//               coefficient row is read from shared memory ONCE per warp (conflict-free LDS.128) and re-used
//               from registers for the warp's 16 sets; per (slot,set) a broadcast LDS.64 brings {dx, which
//               staged row}.  Slots whose sets use 1/2/3 distinct segments evaluate those 1/2/3 polynomials and
//               select the result (warp-uniform); more fall back to a per-set LDS.128.  16 running products in
//               registers, cubic slots in order, then TF1 slots: the reference's order
//               (Splines/SplineMonolith.cpp:799-828), so every (event,set) weight is bit-identical to the
//               single-set path;
//     fill    = CalcWeightTotal per (event,set) (Samples/SampleHandlerFD.cpp:568-594) and one f64 atomic into
//               the set's own histogram [n_sets][n_bins] (L2-resident);
//     -lnL    = llh_batch_kernel, one block per set.
//
// Requires a frozen W2 (the usual state after the first Reweight, SampleHandlerFD.cpp:342); otherwise
// m3b_step_batch falls back to sequential single-set launches.
#include "m3b_device.cuh"
#include "m3b_handle.h"

namespace m3b {

constexpr int kBT = 32;            // events per batch tile
constexpr int kBSets = 256;        // sets per launch
constexpr int kSW = 16;            // sets per consumer warp: 16 warps x 16 sets (more warps per scheduler than 8 x 32)
constexpr int kCW = kBSets / kSW;  // consumer warps
constexpr int kCT = kCW * 32;      // consumer threads
constexpr int kRowF4 = kBT + 1;    // staged cubic row stride in float4: +16 B of padding so that lanes whose sets select
                                   // different segments of a slot (rows r, r') hit different banks at the same event

struct BatchSig {                  // per signature: where its tables start (element offsets) and its row count
  int32_t nc, nl, n_rows, pad;
  int64_t off_dx;                  // float  [nc][kBSets]
  int64_t off_rowoff;              // uint16 [nc][kBSets]   buffer row of (slot, set)
  int64_t off_val;                 // float  [nl][kBSets]
  int64_t off_rowlist;             // int32  [n_rows]       layout row (segbase + segment) of every buffer row
  int64_t off_slot;                // int32  [nc][2]        {first buffer row, number of distinct segments} of the slot
};

struct BatchArgs {
  const TileDesc* tiles; int32_t n_units, units_per_tile, T;
  const BatchSig* sigs;
  const float* t_dx; const uint16_t* t_rowoff; const float* t_val; const int32_t* t_rowlist; const int32_t* t_slot;
  const float* t_norm;             // [n_norm][kBSets]
  int32_t n_norm, n_sets, max_nc, max_nl, max_rows, n_buf;
  const int32_t* bin; const float* osc; const int32_t* osc_idx; const float* static_w;
  const int16_t* norm_idx; int32_t norm_slots; int64_t e_pad, n_events;
  double* hist;                    // [n_bins][kBSets]: a warp's 32 sets update 256 contiguous bytes per event
  int32_t n_bins;
  unsigned int* counter;
};

__device__ __forceinline__ void mbar_arrive_b(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kCT + 32, 1) fill_batch_kernel(const __grid_constant__ BatchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[4], empty_bar[4];
  __shared__ int4 s_desc[4];       // {unit, sig, 0, 0}; unit < 0 = no more work
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // shared-memory map: [tab nc*256 float2 {dx, row rank}][val nl*256 f32][norm 256*nnp f32][slot nc int2][buffers]
  const int nnp = a.n_norm | 1;                    // odd stride: lanes with different norm indices hit different banks
  float2* s_tab = reinterpret_cast<float2*>(smem);
  float* s_val = reinterpret_cast<float*>(s_tab + a.max_nc * kBSets);
  float* s_norm = s_val + a.max_nl * kBSets;
  int2* s_slot = reinterpret_cast<int2*>(s_norm + kBSets * nnp);
  const int tables_bytes = (a.max_nc * kBSets * 8 + a.max_nl * kBSets * 4 + kBSets * nnp * 4 + a.max_nc * 8 + 127) & ~127;
  const int buf_bytes = a.max_rows * kRowF4 * 16 + a.max_nl * kBT * 8;
  unsigned char* bufs = smem + tables_bytes;

  if (tid == 0) for (int b = 0; b < a.n_buf; ++b) { mbar_init(&full_bar[b], 1); mbar_init(&empty_bar[b], kCW); }
  for (int i = tid; i < a.n_norm * kBSets; i += kCT + 32) s_norm[(i % kBSets) * nnp + i / kBSets] = a.t_norm[i];
  __syncthreads();

  if (warp == kCW) {
    // ---------------------------------------------------------------- producer warp
    int buf = 0; uint32_t phase = 1;
    while (true) {
      int u = 0;
      if (lane == 0) u = static_cast<int>(atomicAdd(a.counter, 1u));
      u = __shfl_sync(0xffffffffu, u, 0);
      mbar_wait(&empty_bar[buf], phase);
      if (u >= a.n_units) {
        if (lane == 0) { s_desc[buf] = make_int4(-1, 0, 0, 0); mbar_arrive_b(&full_bar[buf]); }
        break;
      }
      const int t = u / a.units_per_tile, lane0 = (u % a.units_per_tile) * kBT;
      const TileDesc td = a.tiles[t];
      const BatchSig sg = a.sigs[td.sig];
      unsigned char* dst = bufs + static_cast<size_t>(buf) * buf_bytes;
      if (lane == 0) {
        s_desc[buf] = make_int4(u, td.sig, 0, 0);
        const uint32_t bytes = static_cast<uint32_t>(sg.n_rows) * kBT * 16u + static_cast<uint32_t>(sg.nl) * kBT * 8u;
        if (bytes) mbar_expect_tx(&full_bar[buf], bytes); else mbar_arrive_b(&full_bar[buf]);
      }
      __syncwarp();
      const int32_t* rl = a.t_rowlist + sg.off_rowlist;
      for (int r = lane; r < sg.n_rows; r += 32)
        bulk_g2s(dst + static_cast<size_t>(r) * kRowF4 * 16, td.cub + static_cast<int64_t>(rl[r]) * a.T + lane0, kBT * 16u, &full_bar[buf]);
      unsigned char* dl = dst + static_cast<size_t>(a.max_rows) * kRowF4 * 16;
      for (int l = lane; l < sg.nl; l += 32)
        bulk_g2s(dl + static_cast<size_t>(l) * kBT * 8, td.lin + static_cast<int64_t>(l) * a.T + lane0, kBT * 8u, &full_bar[buf]);
      if (++buf == a.n_buf) { buf = 0; phase ^= 1u; }
    }
  } else {
    // ---------------------------------------------------------------- consumers: lane = event, warp = 32 sets
    const int set0 = warp * kSW;
    int buf = 0; uint32_t phase = 0; int cur_sig = -1, nc = 0, nl = 0;
    auto H = [](const float4& k, float dx) { return fmaf(dx, fmaf(dx, fmaf(dx, k.w, k.z), k.y), k.x); };
    while (true) {
      mbar_wait(&full_bar[buf], phase);
      const int4 d = s_desc[buf];
      if (d.x < 0) break;
      if (d.y != cur_sig) {            // block-uniform: reload this signature's per-set tables
        asm volatile("bar.sync 1, %0;" ::"r"(kCT) : "memory");
        const BatchSig sg = a.sigs[d.y];
        nc = sg.nc; nl = sg.nl;
        for (int i = tid; i < nc * kBSets; i += kCT)
          s_tab[i] = make_float2(a.t_dx[sg.off_dx + i], __int_as_float(static_cast<int>(a.t_rowoff[sg.off_rowoff + i])));
        for (int i = tid; i < nl * kBSets; i += kCT) s_val[i] = a.t_val[sg.off_val + i];
        for (int c = tid; c < nc; c += kCT) s_slot[c] = make_int2(a.t_slot[sg.off_slot + 2 * c], a.t_slot[sg.off_slot + 2 * c + 1]);
        cur_sig = d.y;
        asm volatile("bar.sync 1, %0;" ::"r"(kCT) : "memory");
      }
      const unsigned char* src = bufs + static_cast<size_t>(buf) * buf_bytes;
      const float4* cub = reinterpret_cast<const float4*>(src) + lane;
      const float2* lin = reinterpret_cast<const float2*>(src + static_cast<size_t>(a.max_rows) * kRowF4 * 16) + lane;
      // event-table loads of this lane's event fly while the products are formed
      const int t = d.x / a.units_per_tile, lane0 = (d.x % a.units_per_tile) * kBT;
      const int64_t ev = static_cast<int64_t>(t) * a.T + lane0 + lane;
      const int bin = a.bin[ev];
      float w_osc = 1.f, w_static = 1.f;
      if (a.osc) {
        const int64_t oi = a.osc_idx ? static_cast<int64_t>(a.osc_idx[ev]) : (ev < a.n_events ? ev : 0);
        w_osc = oi >= 0 ? a.osc[oi] : 1.f;
      }
      if (a.static_w) w_static = a.static_w[ev];
      int ni[4] = {-1, -1, -1, -1};
      #pragma unroll
      for (int j = 0; j < 4; ++j) if (j < a.norm_slots) ni[j] = a.norm_idx[static_cast<int64_t>(j) * a.e_pad + ev];

      float W[kSW];
      #pragma unroll
      for (int q = 0; q < kSW; ++q) W[q] = 1.0f;
      const bool active = set0 < a.n_sets;              // warps whose 32 sets are all padding only keep the ring moving
      for (int c = 0; active && c < nc; ++c) {
        const int2 si = s_slot[c];                      // {first staged row, distinct segments}: block-uniform
        const float2* tab = s_tab + c * kBSets + set0;
        const float4* rp = cub + si.x * kRowF4;
        if (si.y == 1) {
          const float4 k0 = rp[0];
          #pragma unroll
          for (int q = 0; q < kSW; ++q) W[q] *= H(k0, tab[q].x);
        } else if (si.y <= 3) {
          // staged row 0 is the segment most sets of the batch select: evaluate it for every set and redo the
          // few sets that sit in a neighbouring segment (warp-uniform: the row rank depends on the set only)
          const float4 k0 = rp[0], k1 = rp[kRowF4], k2 = rp[(si.y - 1) * kRowF4];
          #pragma unroll
          for (int q = 0; q < kSW; ++q) {
            const float2 tq = tab[q];
            const int r = __float_as_int(tq.y);
            float h = H(k0, tq.x);
            if (r != 0) h = H(r == 1 ? k1 : k2, tq.x);
            W[q] *= h;
          }
        } else {
          #pragma unroll
          for (int q = 0; q < kSW; ++q) {
            const float2 tq = tab[q];
            W[q] *= H(rp[__float_as_int(tq.y) * kRowF4], tq.x);
          }
        }
      }
      for (int l = 0; active && l < nl; ++l) {
        const float2 k = lin[l * kBT];
        const float* vp = s_val + l * kBSets + set0;
        #pragma unroll
        for (int q = 0; q < kSW; ++q) W[q] *= fmaf(k.x, vp[q], k.y);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_b(&empty_bar[buf]);
      if (++buf == a.n_buf) { buf = 0; phase ^= 1u; }

      // CalcWeightTotal + fill, per (event, set): norms (reference order), osc, spline, static
      if (!active) continue;
      #pragma unroll
      for (int q = 0; q < kSW; ++q) {
        const int set = set0 + q;
        const float* nv = s_norm + set * nnp;
        float w = 1.0f;
        #pragma unroll
        for (int j = 0; j < 4; ++j) if (j < a.norm_slots) w *= (ni[j] >= 0 ? nv[ni[j]] : 1.0f);
        for (int j = 4; j < a.norm_slots; ++j) {
          const int i = a.norm_idx[static_cast<int64_t>(j) * a.e_pad + ev];
          w *= (i >= 0 ? nv[i] : 1.0f);
        }
        w *= w_osc;
        w *= W[q];
        w *= w_static;
        if (set < a.n_sets && w > 0.f && bin >= 0) atomicAdd(a.hist + static_cast<int64_t>(bin) * kBSets + set, static_cast<double>(w));
      }
    }
  }
}

// one block per set: -lnL of the set's histogram column hist[bin*kBSets + set] (frozen W2), into its slot of
// the mapped host array; also un-transposes the LAST set's histogram into `last_out` (the handle's current one)
__global__ void __launch_bounds__(256) llh_batch_kernel(const double* hist, const double* w2, const double* data,
                                                        const int32_t* sample_start, int n_bins, int n_samples, int test_stat,
                                                        double* llh_dev, double* llh_host, double* last_out, int32_t* status) {
  __shared__ double s_part[kMaxSamples * 8];
  bool thrown = false;
  const int set = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const double* col = hist + set;
  for (int s = 0; s < n_samples; ++s) {
    const int b0 = sample_start[s], b1 = sample_start[s + 1];
    double acc = 0.;
    for (int b = b0 + tid; b < b1; b += 256) {
      const double mc = col[static_cast<int64_t>(b) * kBSets];
      if (set == static_cast<int>(gridDim.x) - 1) last_out[b] = mc;
      acc += test_stat_llh(test_stat, data[b], mc, w2 ? w2[b] : 0., thrown);
    }
    acc = warp_sum(acc);
    if (lane == 0) s_part[s * 8 + warp] = acc;
  }
  if (thrown) atomicOr(status, kStatusMathError);
  __syncthreads();
  if (tid < n_samples) {
    double tot = 0.;
    for (int w = 0; w < 8; ++w) tot += s_part[tid * 8 + w];
    s_part[tid * 8] = tot;
  }
  __syncthreads();
  if (tid == 0) {
    double* od = llh_dev + static_cast<int64_t>(set) * (1 + n_samples);
    double* oh = llh_host + static_cast<int64_t>(set) * (1 + n_samples);
    double tot = 0.;
    for (int s = 0; s < n_samples; ++s) { const double v = s_part[s * 8]; tot += v; od[1 + s] = v; oh[1 + s] = v; }
    od[0] = tot; oh[0] = tot;
  }
}

}  // namespace m3b

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
// returns M3B_OK and *done = 1 when the batch ran on the batched kernel; *done = 0 = caller falls back
int m3b_batch_try(m3b_handle* h, int32_t n_sets, const double* spline_pars, const double* norm_pars, const float* osc_w,
                  double* host_slots_dev, int* done) {
  *done = 0;
  if (h->binned || !h->splines_done || h->first_time_w2 || h->cfg.update_w2 || n_sets > kBSets || h->T % kBT != 0) return M3B_OK;
  if (h->cfg.flags & M3B_FLAG_NO_BATCH_KERNEL) return M3B_OK;
  if (!(h->cfg.flags & M3B_FLAG_BATCH_KERNEL_V1)) {      // the second-generation kernel first (m3b_batch2.cuh)
    const int rc2 = m3b_batch2_try(h, n_sets, spline_pars, norm_pars, osc_w, host_slots_dev, done);
    if (rc2 != M3B_OK || *done) return rc2;
  }
  if (h->tiles_dirty || !h->d_tiles) return M3B_OK;     // first step has not run yet
  CK(cudaSetDevice(h->device));
  if (h->Kmax > 64) return M3B_OK;
  const int P = h->P, S = kBSets, n_sigs = static_cast<int>(h->sigs.size());
  // 1. segments of every set, in order (SplineBase::FindSplineSegment keeps its cached-segment history)
  std::vector<int16_t> seg(static_cast<size_t>(n_sets) * P);
  std::vector<float> val(static_cast<size_t>(n_sets) * P);
  for (int s = 0; s < n_sets; ++s) {
    int rc = m3b_find_segments(h, spline_pars + static_cast<size_t>(s) * P, seg.data() + static_cast<size_t>(s) * P,
                               val.data() + static_cast<size_t>(s) * P);
    if (rc != M3B_OK) return rc;
  }
  // 2. per signature and slot: distinct segments -> buffer rows; per set: dx and its row
  std::vector<BatchSig> bs(n_sigs);
  std::vector<float> t_dx, t_val; std::vector<uint16_t> t_rowoff; std::vector<int32_t> t_rowlist, t_slot;
  int max_rows = 0;
  for (int g = 0; g < n_sigs; ++g) {
    const SigDesc& sd = h->sigs[g];
    const int32_t* pool = h->sig_pool.data() + sd.off;
    BatchSig& b = bs[g];
    b.nc = sd.nc; b.nl = sd.nl; b.pad = 0;
    b.off_dx = static_cast<int64_t>(t_dx.size()); b.off_rowoff = static_cast<int64_t>(t_rowoff.size());
    b.off_val = static_cast<int64_t>(t_val.size()); b.off_rowlist = static_cast<int64_t>(t_rowlist.size());
    b.off_slot = static_cast<int64_t>(t_slot.size());
    t_dx.resize(t_dx.size() + static_cast<size_t>(sd.nc) * S, 0.f);
    t_rowoff.resize(t_rowoff.size() + static_cast<size_t>(sd.nc) * S, 0);
    t_val.resize(t_val.size() + static_cast<size_t>(sd.nl) * S, 0.f);
    int rows = 0;
    for (int c = 0; c < sd.nc; ++c) {
      const int p = pool[c], segbase = pool[sd.nc + c];
      int rank_of[64]; for (int k = 0; k < 64; ++k) rank_of[k] = -1;
      int used[64] = {0};
      for (int s = 0; s < n_sets; ++s) ++used[seg[static_cast<size_t>(s) * P + p]];
      const int row0 = rows;
      while (true) {                       // most popular segment first (rank 0 = the kernel's fast path)
        int best = -1;
        for (int k = 0; k < 64; ++k) if (used[k] > 0 && (best < 0 || used[k] > used[best])) best = k;
        if (best < 0) break;
        rank_of[best] = rows++ - row0; t_rowlist.push_back(segbase + best); used[best] = 0;
      }
      t_slot.push_back(row0); t_slot.push_back(rows - row0);
      for (int s = 0; s < S; ++s) {
        const int ss = s < n_sets ? s : n_sets - 1;           // padding lanes repeat the last set (never filled)
        const int sg = seg[static_cast<size_t>(ss) * P + p];
        t_rowoff[b.off_rowoff + static_cast<size_t>(c) * S + s] = static_cast<uint16_t>(rank_of[sg]);
        // dx = ParamValues[Param] - coeff_x[Param*_max_knots+segment]   (Splines/SplineMonolith.cpp:759), float
        t_dx[b.off_dx + static_cast<size_t>(c) * S + s] = val[static_cast<size_t>(ss) * P + p] - h->coeff_x[static_cast<size_t>(p) * h->Kmax + sg];
      }
    }
    for (int l = 0; l < sd.nl; ++l) {
      const int p = pool[2 * sd.nc + l];
      for (int s = 0; s < S; ++s) t_val[b.off_val + static_cast<size_t>(l) * S + s] = val[static_cast<size_t>(s < n_sets ? s : n_sets - 1) * P + p];
    }
    b.n_rows = rows;
    max_rows = std::max(max_rows, rows);
  }
  const int Nn = h->n_norm_values;
  std::vector<float> t_norm(static_cast<size_t>(std::max(Nn, 1)) * S, 1.f);
  for (int n = 0; n < Nn; ++n)
    for (int s = 0; s < S; ++s) t_norm[static_cast<size_t>(n) * S + s] = static_cast<float>(norm_pars[static_cast<size_t>(s < n_sets ? s : n_sets - 1) * Nn + n]);
  // 3. shared-memory budget: tables + n_buf tile buffers
  const int tables_bytes = (h->max_nc * S * 8 + h->max_nl * S * 4 + S * (Nn | 1) * 4 + h->max_nc * 8 + 127) & ~127;
  const int buf_bytes = max_rows * kRowF4 * 16 + h->max_nl * kBT * 8;
  const int budget = 232448 - 2048;
  int n_buf = buf_bytes > 0 ? std::min(3, (budget - tables_bytes) / buf_bytes) : 2;
  if (n_buf < 1) return M3B_OK;                 // too many distinct segments for shared memory: sequential path
  const int smem = tables_bytes + n_buf * std::max(buf_bytes, 16);

  // 4. device staging (grown on demand, kept in the handle)
  auto grow = [&](void** p, size_t& cap, size_t bytes) -> cudaError_t {
    if (cap >= bytes && *p) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr; cap = 0;
    const cudaError_t e = cudaMalloc(p, bytes + bytes / 4 + 256);
    if (e == cudaSuccess) cap = bytes + bytes / 4 + 256;
    return e;
  };
  const size_t slot = static_cast<size_t>(1 + h->n_samples);
  CK(grow(&h->bt_dx, h->bt_dx_cap, t_dx.size() * 4 + 16));
  CK(grow(&h->bt_rowoff, h->bt_rowoff_cap, t_rowoff.size() * 2 + 16));
  CK(grow(&h->bt_val, h->bt_val_cap, t_val.size() * 4 + 16));
  CK(grow(&h->bt_rowlist, h->bt_rowlist_cap, t_rowlist.size() * 4 + 16));
  CK(grow(&h->bt_slot, h->bt_slot_cap, t_slot.size() * 4 + 16));
  CK(grow(&h->bt_norm, h->bt_norm_cap, t_norm.size() * 4));
  CK(grow(&h->bt_sigs, h->bt_sigs_cap, bs.size() * sizeof(BatchSig)));
  CK(grow(&h->bt_hist, h->bt_hist_cap, static_cast<size_t>(S) * h->n_bins * 8));
  CK(grow(&h->bt_llh, h->bt_llh_cap, static_cast<size_t>(S) * slot * 8));
  CK(cudaStreamSynchronize(h->stream));          // previous batch may still read the staging vectors' device copies
  if (!t_dx.empty()) CK(cudaMemcpyAsync(h->bt_dx, t_dx.data(), t_dx.size() * 4, cudaMemcpyHostToDevice, h->stream));
  if (!t_rowoff.empty()) CK(cudaMemcpyAsync(h->bt_rowoff, t_rowoff.data(), t_rowoff.size() * 2, cudaMemcpyHostToDevice, h->stream));
  if (!t_val.empty()) CK(cudaMemcpyAsync(h->bt_val, t_val.data(), t_val.size() * 4, cudaMemcpyHostToDevice, h->stream));
  if (!t_rowlist.empty()) CK(cudaMemcpyAsync(h->bt_rowlist, t_rowlist.data(), t_rowlist.size() * 4, cudaMemcpyHostToDevice, h->stream));
  if (!t_slot.empty()) CK(cudaMemcpyAsync(h->bt_slot, t_slot.data(), t_slot.size() * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->bt_norm, t_norm.data(), t_norm.size() * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->bt_sigs, bs.data(), bs.size() * sizeof(BatchSig), cudaMemcpyHostToDevice, h->stream));
  if (osc_w) {
    REQUIRE(h->use_osc, M3B_ERR_INVALID, "step: osc_w given but events were uploaded with use_osc=0");
    CK(cudaMemcpyAsync(h->d_osc, osc_w, sizeof(float) * h->n_osc, cudaMemcpyHostToDevice, h->stream));
  }
  CK(cudaMemsetAsync(h->bt_hist, 0, static_cast<size_t>(S) * h->n_bins * 8, h->stream));
  CK(cudaMemsetAsync(h->d_tile_counter, 0, sizeof(unsigned int), h->stream));

  BatchArgs a{};
  a.tiles = h->d_tiles; a.units_per_tile = h->T / kBT; a.n_units = static_cast<int32_t>(h->n_tiles) * a.units_per_tile; a.T = h->T;
  a.sigs = static_cast<const BatchSig*>(h->bt_sigs);
  a.t_dx = static_cast<const float*>(h->bt_dx); a.t_rowoff = static_cast<const uint16_t*>(h->bt_rowoff);
  a.t_val = static_cast<const float*>(h->bt_val); a.t_rowlist = static_cast<const int32_t*>(h->bt_rowlist);
  a.t_norm = static_cast<const float*>(h->bt_norm); a.t_slot = static_cast<const int32_t*>(h->bt_slot);
  a.n_norm = Nn; a.n_sets = n_sets; a.max_nc = h->max_nc; a.max_nl = h->max_nl; a.max_rows = max_rows; a.n_buf = n_buf;
  a.bin = h->d_bin; a.osc = h->use_osc ? h->d_osc : nullptr; a.osc_idx = h->d_osc_idx; a.static_w = h->d_static;
  a.norm_idx = h->d_norm_idx; a.norm_slots = h->norm_slots; a.e_pad = h->e_pad; a.n_events = h->n_events;
  a.hist = static_cast<double*>(h->bt_hist); a.n_bins = h->n_bins; a.counter = h->d_tile_counter;
  CK(allow_max_dynamic_smem(fill_batch_kernel));
  const int grid = static_cast<int>(std::min<int64_t>(a.n_units, h->sm_count));
  if (h->timing) {
    if (h->tev_used + 2 > h->tev.size()) for (int i = 0; i < 2; ++i) { cudaEvent_t e; CK(cudaEventCreate(&e)); h->tev.push_back(e); }
    CK(cudaEventRecord(h->tev[h->tev_used], h->stream));
  }
  fill_batch_kernel<<<grid, kCT + 32, smem, h->stream>>>(a);
  CK(cudaGetLastError());
  if (h->timing) { CK(cudaEventRecord(h->tev[h->tev_used + 1], h->stream)); h->tev_used += 2; }
  const int nxt = h->cur ^ 1;       // leave the handle as after the last set's step: its histogram becomes the current one
  llh_batch_kernel<<<n_sets, 256, 0, h->stream>>>(static_cast<const double*>(h->bt_hist), h->d_w2_frozen, h->d_data, h->d_sample_start,
                                                  h->n_bins, h->n_samples, h->test_stat, static_cast<double*>(h->bt_llh), host_slots_dev,
                                                  h->d_hw[nxt], h->d_status);
  CK(cudaGetLastError());
  CK(cudaMemsetAsync(h->d_tile_counter, 0, sizeof(unsigned int), h->stream));
  h->mc_zero[nxt] = false;
  h->cur = nxt;
  h->launches += 2; h->steps += static_cast<uint64_t>(n_sets);
  h->evt_weights_valid = false;
  *done = 1;
  return M3B_OK;
}

#include "m3b_batch2.cuh"
