// m3b_batch2.cuh -- batched proposals, second generation (BASELINE config 5).  Included at the end of m3b_batch.cu (one
// translation unit: both generations share llh_batch_kernel and the handle's staging buffers).
//
// Same contract as m3b_batch.cu (up to 256 parameter sets against the same events in ONE pass over the coefficient
// rows, every (event,set) weight formed in the reference's order, Splines/SplineMonolith.cpp:738-830 and
// Samples/SampleHandlerFD.cpp:568-594, so bit-identical to the single-set path) with the three things the first kernel's
// profile asked for (profiles/r01_ncu_full_fill_batch_cfg5_600k.json: issue-active 48 %, 3.6e8 shared wavefronts):
//
//   two events per thread   a consumer lane carries events lane and lane+32 of a 64-event unit, so every per-(slot,set)
//                           control instruction -- the table load, the rank test, the branch -- is paid once for two
//                           polynomial evaluations instead of once for one;
//   a ring of slot groups   the unit's coefficient rows no longer sit in one whole-unit buffer: they are streamed in groups
//                           of slots (<= 32 rows, 1 KB bulk copies) through a 4-stage ring while the consumers keep their
//                           2 x 16 running products in registers across the groups, exactly like fill_tma_kernel -- the
//                           next unit's rows are in flight while this one is evaluated.  Warp 0 doubles as the producer
//                           (16 warps = 4 per scheduler = 128 registers per thread; a 17th warp would cap all at 96);
//   wide table loads        dx of a warp's 16 sets comes in as 4 broadcast LDS.128, the ranks (which of the <= 3 staged
//                           segments a set selects) as ONE 2-bit-per-set word per (slot, warp): 5 shared-memory
//                           wavefronts per (slot, warp) instead of 16; the rank branch is warp-uniform and three-way, so
//                           a set outside the majority segment costs one extra test, not a second polynomial.
//
//   packed arithmetic       a lane's two events ride in the two halves of FFMA2 / FMUL2 (sm_100 packed fp32: two independent
//                           IEEE fmaf per instruction, so the weights stay bit-identical): 3 + 1 instructions per
//                           (slot, set) for both events.  The packed operands must sit in aligned register pairs, so one
//                           warp per staged row first re-packs it in shared memory from {y,b,c,d} per event to
//                           {y0,y1,b0,b1} | {c0,c1,d0,d1} per event PAIR (each lane rewrites only its own two float4: no
//                           hazard inside the row) and a second mbarrier per stage tells the block the rows are ready
//                           (packing in registers instead -- four pair moves per row -- was measured slower: 36.1 vs
//                           31.5 ms on config 5); dx is duplicated into its pair in registers;
//   a cheap epilogue        the per-(event,set) norm look-ups walk pre-computed shared-memory offsets (absent slots point
//                           at a 1.0 entry): 8 LDS + 8 FMUL per set instead of an address computation and a branch each.
//
//   uniform and wide slots  a slot at which every set has the same segment and dx (an LLH scan of another parameter) evaluates
//                           one polynomial and multiplies it into the 16 products; a slot whose sets spread over 4..8
//                           segments (the scanned parameter) takes its row per set from a rank byte.
//
// Slots with more than eight live segments, more than eight wide slots per signature, more than four norm slots, and batches
// on handles without a frozen W2, take the first kernel / the sequential path (m3b_batch_try, m3b_step_batch).
#pragma once

namespace m3b {

constexpr int kB2E = 64;            // events per unit (two per consumer lane)
constexpr int kB2Sets = 256;        // sets per launch
constexpr int kB2SW = 16;           // sets per consumer warp
constexpr int kB2CW = kB2Sets / kB2SW;
constexpr int kB2CT = kB2CW * 32;
constexpr int kB2RowF4 = kB2E + 1;  // staged row stride in float4 (+16 B: rows r, r' at one event land in different banks)
constexpr int kB2MaxStageRows = 32; // cubic rows per ring stage (the host picks 32, 24 or 16 and 2..4 stages: what fits next to the tables)
constexpr int kB2MaxStages = 4;
constexpr int kB2MaxGroups = 96;
constexpr int kB2MaxRanks = 8;      // staged segments of one slot: up to 3 go through the 2-bit code word, 4..8 ("wide" slots: a
constexpr int kB2MaxWide = 8;       // parameter scanned over its whole range) through a byte per set, at most kB2MaxWide such slots
__host__ __device__ inline int batch2_norm_stride(int n_norm) { return (n_norm + 1) | 1; }     // + the 1.0 entry; odd: lanes with different indices hit different banks
__host__ __device__ inline int batch2_tables_bytes(int max_nc, int max_nl, int n_norm) {
  return (max_nc * kB2Sets * 4 + max_nc * kB2CW * 4 + max_nl * kB2Sets * 4 + kB2Sets * batch2_norm_stride(n_norm) * 4 + max_nc * 16 + kB2MaxWide * kB2Sets + 127) & ~127;
}

struct Batch2Group { int32_t c0, c1, n_rows, row0; };   // slots [c0,c1) of the signature; rows [row0,row0+n_rows) of its row list
struct Batch2Sig {
  int32_t nc, nl, n_groups, n_wide;
  int64_t off_dx;        // float    [nc][256]
  int64_t off_code;      // uint32   [nc][16]     2 bits per set of the warp: rank of the set's segment among the slot's staged rows
  int64_t off_val;       // float    [nl][256]
  int64_t off_rowlist;   // int32    [rows]       layout row (segbase + segment) of every staged row, slot-major, rank-minor
  int64_t off_slot;      // int32    [nc][4]      {row offset inside its group's stage, distinct segments, wide index or -1,
                         //                        1 = every set has the same segment and dx: one polynomial serves all sets}
  int64_t off_rank8;     // uint8    [n_wide][256] rank of every set's segment, for the slots with more than three
  uint64_t lin_uniform;  // bit l: every set has the same value of TF1 slot l's parameter (l < 64)
  int64_t off_group;     // Batch2Group [n_groups]
};

struct Batch2Args {
  const TileDesc* tiles; int32_t n_units, units_per_tile, T;
  const Batch2Sig* sigs;
  const float* t_dx; const uint32_t* t_code; const float* t_val; const int32_t* t_rowlist; const int32_t* t_slot; const Batch2Group* t_group; const uint8_t* t_rank8;
  const float* t_norm;             // [n_norm][256]
  int32_t n_norm, n_sets, max_nc, max_nl;
  int32_t n_stages, stage_rows;
  int32_t norm_uniform;            // every set has the same normalisation parameters: the norm product is formed once per event
  int32_t dbg;                     // experiments build only (timing ablations, scripts/batch_ablation.py; results are wrong
                                   // when set): 1 no atomics, 2 no epilogue, 4 no spline arithmetic, 8 no re-pack / second barrier
  const int32_t* bin; const float* osc; const int32_t* osc_idx; const float* static_w;
  const int16_t* norm_idx; int32_t norm_slots; int64_t e_pad, n_events;
  double* hist;                    // [n_bins][256]
  int32_t n_bins;
  unsigned int* counter;
};

#ifdef M3B_EXPERIMENTS
#define B2DBG(bit) ((a.dbg & (bit)) != 0)
#else
#define B2DBG(bit) false
#endif
__device__ __forceinline__ void mbar_arrive2(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Packed fp32 pairs live in 64-bit registers (lo = event lane, hi = event lane + 32).  fma.rn.f32x2 / mul.rn.f32x2 are two
// independent IEEE operations per instruction, so every weight equals the scalar fmaf path's bit for bit.
typedef unsigned long long pk2;
__device__ __forceinline__ pk2 pk2_ffma(pk2 a, pk2 b, pk2 c) { pk2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ pk2 pk2_mul(pk2 a, pk2 b) { pk2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ pk2 pk2_make(float lo, float hi) { pk2 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ float pk2_lo(pk2 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return lo; }
__device__ __forceinline__ float pk2_hi(pk2 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return hi; }
// one staged row for the lane's two events, as re-packed in shared memory: {y pair, b pair} | {c pair, d pair}, two LDS.128
struct B2Row { pk2 y, b, c, d; };
__device__ __forceinline__ B2Row b2_row(const float4* rp) {
  const ulonglong2 yb = *reinterpret_cast<const ulonglong2*>(rp), cd = *reinterpret_cast<const ulonglong2*>(rp + 32);
  B2Row k; k.y = yb.x; k.b = yb.y; k.c = cd.x; k.d = cd.y;
  return k;
}
// fmaf(dx, fmaf(dx, fmaf(dx, d, c), b), y)          (Splines/SplineMonolith.cpp:765), both events
__device__ __forceinline__ pk2 horner2(const B2Row& k, pk2 dx) {
  return pk2_ffma(dx, pk2_ffma(dx, pk2_ffma(dx, k.d, k.c), k.b), k.y);
}
// The producer's state: which unit / slot group goes into the ring next.  The producer role is taken by consumer warp 0
// (a 17th warp would put five warps on one scheduler's register file and cap every thread at 96 registers; 16 warps get
// 128): before warp 0 waits for a stage it makes sure that stage has been issued, and it issues further ahead whenever a
// ring slot is free -- a few dozen instructions per stage next to ~2500 of evaluation.  The state lives in shared memory
// and is loaded into registers only inside pump(): every thread runs the same code, and 20 registers of producer state
// held across the evaluation loops would be spilled by all 512.
struct Batch2Producer {
  int stage = 0; uint32_t phase = 1;       // a fresh mbarrier passes a wait on the "previous" phase
  int u = -2, g = 0, n_groups = 0, nl = 0; // u == -2: fetch the next unit
  int64_t off_group = 0, off_rowlist = 0;
  TileDesc td{};
  int lane0 = 0;
  bool done = false;
};
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(kB2CT, 1) fill_batch2_kernel(const __grid_constant__ Batch2Args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[kB2MaxStages], ready_bar[kB2MaxStages], empty_bar[kB2MaxStages];
  __shared__ int4 s_desc[kB2MaxStages];   // {unit, sig, group (-1: the TF1 group), 0}; unit < 0 = no more work
  __shared__ Batch2Producer s_pr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // shared-memory map: [dx nc*256 f32][code nc*16 u32][val nl*256 f32][norm 256*nnp f32][slot nc int2][ring]
  const int nnp = batch2_norm_stride(a.n_norm);
  float* s_dx = reinterpret_cast<float*>(smem);
  uint32_t* s_code = reinterpret_cast<uint32_t*>(s_dx + a.max_nc * kB2Sets);
  float* s_val = reinterpret_cast<float*>(s_code + a.max_nc * kB2CW);
  float* s_norm = s_val + a.max_nl * kB2Sets;
  int4* s_slot = reinterpret_cast<int4*>((reinterpret_cast<uintptr_t>(s_norm + kB2Sets * nnp) + 15) & ~static_cast<uintptr_t>(15));
  uint8_t* s_rank8 = reinterpret_cast<uint8_t*>(s_slot + a.max_nc);
  unsigned char* ring = smem + batch2_tables_bytes(a.max_nc, a.max_nl, a.n_norm);
  const int stage_bytes = a.stage_rows * kB2RowF4 * 16;
  const int n_stages = a.n_stages;

  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&ready_bar[s], kB2CW); mbar_init(&empty_bar[s], kB2CW); }
    s_pr = Batch2Producer();
  }
  // norm table [set][nnp]; entry n_norm of every set = 1.0 (where absent norm slots point)
  for (int i = tid; i < a.n_norm * kB2Sets; i += kB2CT) s_norm[(i % kB2Sets) * nnp + i / kB2Sets] = a.t_norm[i];
  for (int i = tid; i < kB2Sets; i += kB2CT) s_norm[i * nnp + a.n_norm] = 1.0f;
  __syncthreads();

  // ---------------------------------------------------------------- producer (warp 0 only; all 32 lanes, convergent)
  // issues ONE stage (a slot group, a TF1 group or the terminal marker); block = wait for the ring slot, else give up
  auto pump = [&](bool block) -> bool {
    Batch2Producer pr = s_pr;
    __syncwarp();
    if (pr.done) return false;
    if (pr.u == -2) {
      int u = 0;
      if (lane == 0) u = static_cast<int>(atomicAdd(a.counter, 1u));
      u = __shfl_sync(0xffffffffu, u, 0);
      pr.u = u; pr.g = 0;
      if (u < a.n_units) {
        pr.td = a.tiles[u / a.units_per_tile];
        pr.lane0 = (u % a.units_per_tile) * kB2E;
        const Batch2Sig sg = a.sigs[pr.td.sig];
        pr.n_groups = sg.n_groups; pr.nl = sg.nl; pr.off_group = sg.off_group; pr.off_rowlist = sg.off_rowlist;
      }
    }
    if (!mbar_test(&empty_bar[pr.stage], pr.phase)) {
      if (!block) { if (lane == 0) s_pr = pr; __syncwarp(); return false; }      // (the unit fetched above is kept)
      mbar_wait(&empty_bar[pr.stage], pr.phase);
    }
    unsigned char* dst = ring + static_cast<size_t>(pr.stage) * stage_bytes;
    if (pr.u >= a.n_units) {
      if (lane == 0) { s_desc[pr.stage] = make_int4(-1, 0, 0, 0); mbar_arrive2(&full_bar[pr.stage]); }
      pr.done = true;
    } else if (pr.g < pr.n_groups) {
      const Batch2Group gr = a.t_group[pr.off_group + pr.g];
      const int32_t* rl = a.t_rowlist + pr.off_rowlist;
      if (lane == 0) {
        s_desc[pr.stage] = make_int4(pr.u, pr.td.sig, pr.g, 0);
        mbar_expect_tx(&full_bar[pr.stage], static_cast<uint32_t>(gr.n_rows) * kB2E * 16u);
      }
      __syncwarp();
      for (int r = lane; r < gr.n_rows; r += 32)
        bulk_g2s(dst + static_cast<size_t>(r) * kB2RowF4 * 16, pr.td.cub + static_cast<int64_t>(rl[gr.row0 + r]) * a.T + pr.lane0, kB2E * 16u, &full_bar[pr.stage]);
      ++pr.g;
    } else {
      // the TF1 group: nl rows of {a,b}; also the stage at which the unit's events are filled (always present)
      if (lane == 0) {
        s_desc[pr.stage] = make_int4(pr.u, pr.td.sig, -1, 0);
        if (pr.nl > 0) mbar_expect_tx(&full_bar[pr.stage], static_cast<uint32_t>(pr.nl) * kB2E * 8u); else mbar_arrive2(&full_bar[pr.stage]);
      }
      __syncwarp();
      for (int l = lane; l < pr.nl; l += 32)
        bulk_g2s(dst + static_cast<size_t>(l) * kB2E * 8, pr.td.lin + static_cast<int64_t>(l) * a.T + pr.lane0, kB2E * 8u, &full_bar[pr.stage]);
      pr.u = -2;
    }
    if (++pr.stage == n_stages) { pr.stage = 0; pr.phase ^= 1u; }
    __syncwarp();
    if (lane == 0) s_pr = pr;
    __syncwarp();
    return true;
  };

  {
    // ---------------------------------------------------------------- consumers: lane = 2 events, warp = 16 sets
    const int set0 = warp * kB2SW;
    const bool active = set0 < a.n_sets;                // warps whose sets are all padding only keep the ring moving
    int stage = 0; uint32_t phase = 0; int cur_sig = -1, cur_nl = 0; int64_t off_group = 0; uint64_t lin_uniform = 0;
    pk2 W[kB2SW];                                       // running products: lo = event lane, hi = event lane + 32
    const pk2 one2 = pk2_make(1.0f, 1.0f);
    #pragma unroll
    for (int q = 0; q < kB2SW; ++q) W[q] = one2;
    unsigned consumed = 0, issued = 0;
    while (true) {
      if (warp == 0) {
        while (issued <= consumed && pump(true)) ++issued;                    // the stage about to be waited for must be in flight
        while (issued < consumed + n_stages && pump(false)) ++issued;         // and as many further ones as the ring has room for
      }
      mbar_wait(&full_bar[stage], phase);
      ++consumed;
      const int4 d = s_desc[stage];
      if (d.x < 0) break;
      if (d.y != cur_sig) {            // block-uniform: reload this signature's per-set tables
        asm volatile("bar.sync 1, %0;" ::"r"(kB2CT) : "memory");
        const Batch2Sig sg = a.sigs[d.y];
        for (int i = tid; i < sg.nc * kB2Sets; i += kB2CT) s_dx[i] = a.t_dx[sg.off_dx + i];
        for (int i = tid; i < sg.nc * kB2CW; i += kB2CT) s_code[i] = a.t_code[sg.off_code + i];
        for (int i = tid; i < sg.nl * kB2Sets; i += kB2CT) s_val[i] = a.t_val[sg.off_val + i];
        for (int c = tid; c < sg.nc; c += kB2CT) s_slot[c] = reinterpret_cast<const int4*>(a.t_slot + sg.off_slot)[c];
        for (int i = tid; i < sg.n_wide * (kB2Sets / 4); i += kB2CT) reinterpret_cast<uint32_t*>(s_rank8)[i] = reinterpret_cast<const uint32_t*>(a.t_rank8 + sg.off_rank8)[i];
        cur_sig = d.y; off_group = sg.off_group; cur_nl = sg.nl; lin_uniform = sg.lin_uniform;
        asm volatile("bar.sync 1, %0;" ::"r"(kB2CT) : "memory");
      }
      unsigned char* src = ring + static_cast<size_t>(stage) * stage_bytes;
      if (d.z >= 0) {
        // ---- a group of TSpline3 slots
        const Batch2Group gr = a.t_group[off_group + d.z];
        // re-pack this warp's share of the staged rows for packed arithmetic: lane l owns float4 l and l+32 of a row
        if (!B2DBG(8)) {
          float4* all = reinterpret_cast<float4*>(src) + lane;
          for (int r = warp; r < gr.n_rows; r += kB2CW) {
            float4* rp = all + r * kB2RowF4;
            const float4 e0 = rp[0], e1 = rp[32];
            rp[0] = make_float4(e0.x, e1.x, e0.y, e1.y);
            rp[32] = make_float4(e0.z, e1.z, e0.w, e1.w);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the ring slot is re-filled by bulk copies later
          __syncwarp();
          if (lane == 0) mbar_arrive2(&ready_bar[stage]);
          mbar_wait(&ready_bar[stage], phase);
        }
        // fmaf Horner on the set's segment, running products in the reference's slot order
        const float4* rows = reinterpret_cast<const float4*>(src) + lane;
        for (int c = gr.c0; active && !B2DBG(4) && c < gr.c1; ++c) {
          const int4 si = s_slot[c];                    // {first staged row of the slot in this stage, distinct segments, wide, uniform}
          const float4* rp = rows + si.x * kB2RowF4;
          const float4* dxv = reinterpret_cast<const float4*>(s_dx + c * kB2Sets + set0);
          const B2Row k0 = b2_row(rp);
          if (si.w) {
            // every set has this slot's parameter at the same value (an LLH scan of another parameter, sets that differ
            // in a few parameters only): one polynomial, then the 16 products -- each set's product order is unchanged
            const float d = *reinterpret_cast<const float*>(dxv);
            const pk2 t = horner2(k0, pk2_make(d, d));
            #pragma unroll
            for (int q = 0; q < kB2SW; ++q) W[q] = pk2_mul(W[q], t);
            continue;
          }
          if (si.y > 3) {
            // a slot whose sets spread over more than three segments (the scanned parameter): row per set from its rank byte
            const uint4 rk = *reinterpret_cast<const uint4*>(s_rank8 + si.z * kB2Sets + set0);
            const uint32_t rw[4] = {rk.x, rk.y, rk.z, rk.w};
            #pragma unroll
            for (int q = 0; q < kB2SW; ++q) {
              const uint32_t r = (rw[q >> 2] >> (8 * (q & 3))) & 0xffu;
              const float d = reinterpret_cast<const float*>(dxv)[q];
              W[q] = pk2_mul(W[q], horner2(b2_row(rp + r * kB2RowF4), pk2_make(d, d)));
            }
            continue;
          }
          // 2 bits per set, warp-uniform; rank 0 = the slot's most popular segment
          const uint32_t code = si.y == 1 ? 0u : s_code[c * kB2CW + warp];
          if (code == 0u) {                             // all 16 sets of this warp on the slot's first row
            #pragma unroll
            for (int hq = 0; hq < kB2SW; hq += 8) {     // dx in two halves: 16 registers live instead of 32
              pk2 dx[8];
              #pragma unroll
              for (int j = 0; j < 2; ++j) {             // (dx,dx): the same set's dx for both events of the lane
                const float4 v = dxv[hq / 4 + j];
                dx[4 * j] = pk2_make(v.x, v.x); dx[4 * j + 1] = pk2_make(v.y, v.y); dx[4 * j + 2] = pk2_make(v.z, v.z); dx[4 * j + 3] = pk2_make(v.w, v.w);
              }
              #pragma unroll
              for (int q = 0; q < 8; ++q) W[hq + q] = pk2_mul(W[hq + q], horner2(k0, dx[q]));
            }
          } else {
            // mixed ranks: a warp-uniform branch per set.  (Evaluating both candidate rows and selecting keeps the 16
            // chains independent but doubles the FMA-pipe work -- FFMA2 halves the issue slots, not the pipe time:
            // scripts/microbench/ffma2_rate.cu -- and measured slower: 34.8 vs 32.1 ms on config 5.)
            const B2Row k1 = b2_row(rp + kB2RowF4), k2 = b2_row(rp + (si.y - 1) * kB2RowF4);
            #pragma unroll
            for (int hq = 0; hq < kB2SW; hq += 8) {
              pk2 dx[8];
              #pragma unroll
              for (int j = 0; j < 2; ++j) {             // (dx,dx): the same set's dx for both events of the lane
                const float4 v = dxv[hq / 4 + j];
                dx[4 * j] = pk2_make(v.x, v.x); dx[4 * j + 1] = pk2_make(v.y, v.y); dx[4 * j + 2] = pk2_make(v.z, v.z); dx[4 * j + 3] = pk2_make(v.w, v.w);
              }
              #pragma unroll
              for (int q = 0; q < 8; ++q) {
                pk2 t;
                if ((code & (3u << (2 * (hq + q)))) == 0u) t = horner2(k0, dx[q]);
                else if ((code & (2u << (2 * (hq + q)))) == 0u) t = horner2(k1, dx[q]);
                else t = horner2(k2, dx[q]);
                W[hq + q] = pk2_mul(W[hq + q], t);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive2(&empty_bar[stage]);
        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        continue;
      }
      // ---- the unit's last stage: TF1 slots, then CalcWeightTotal + fill for both events and all 16 sets
      const int t = d.x / a.units_per_tile, lane0 = (d.x % a.units_per_tile) * kB2E;
      const int64_t ev0 = static_cast<int64_t>(t) * a.T + lane0 + lane, ev1 = ev0 + 32;
      const int bin0 = a.bin[ev0], bin1 = a.bin[ev1];
      float osc0 = 1.f, osc1 = 1.f, st0 = 1.f, st1 = 1.f;
      if (a.osc) {
        const int64_t o0 = a.osc_idx ? static_cast<int64_t>(a.osc_idx[ev0]) : (ev0 < a.n_events ? ev0 : 0);
        const int64_t o1 = a.osc_idx ? static_cast<int64_t>(a.osc_idx[ev1]) : (ev1 < a.n_events ? ev1 : 0);
        osc0 = o0 >= 0 ? a.osc[o0] : 1.f; osc1 = o1 >= 0 ? a.osc[o1] : 1.f;
      }
      if (a.static_w) { st0 = a.static_w[ev0]; st1 = a.static_w[ev1]; }
      // shared-memory offsets of the (at most four: m3b_batch2_try) norm slots' values for set0 (absent: the 1.0 entry)
      int np0[4], np1[4];
      #pragma unroll
      for (int j = 0; j < 4; ++j) {
        int i0 = -1, i1 = -1;
        if (j < a.norm_slots) { i0 = a.norm_idx[static_cast<int64_t>(j) * a.e_pad + ev0]; i1 = a.norm_idx[static_cast<int64_t>(j) * a.e_pad + ev1]; }
        np0[j] = set0 * nnp + (i0 >= 0 ? i0 : a.n_norm);
        np1[j] = set0 * nnp + (i1 >= 0 ? i1 : a.n_norm);
      }
      // every warp arrives on the stage's second barrier once per use, whether or not the stage had rows to re-pack:
      // its phase must advance in step with the ring's
      if (lane == 0 && !B2DBG(8)) mbar_arrive2(&ready_bar[stage]);
      if (active) {
        const float2* lin = reinterpret_cast<const float2*>(src) + lane;
        for (int l = 0; l < cur_nl; ++l) {
          const float2 c0 = lin[l * kB2E], c1 = lin[l * kB2E + 32];
          const float4* vv = reinterpret_cast<const float4*>(s_val + l * kB2Sets + set0);
          // fmaf(a, x, b) (Splines/SplineMonolith.cpp:800) for both events: a, b packed over the events, x per set
          const pk2 a2 = pk2_make(c0.x, c1.x), b2 = pk2_make(c0.y, c1.y);
          if (l < 64 && ((lin_uniform >> l) & 1ull)) {          // same parameter value in every set: one evaluation
            const float x = *reinterpret_cast<const float*>(vv);
            const pk2 t = pk2_ffma(a2, pk2_make(x, x), b2);
            #pragma unroll
            for (int q = 0; q < kB2SW; ++q) W[q] = pk2_mul(W[q], t);
            continue;
          }
          #pragma unroll
          for (int j = 0; j < kB2SW / 4; ++j) {
            const float4 v = vv[j];
            W[4 * j] = pk2_mul(W[4 * j], pk2_ffma(a2, pk2_make(v.x, v.x), b2));
            W[4 * j + 1] = pk2_mul(W[4 * j + 1], pk2_ffma(a2, pk2_make(v.y, v.y), b2));
            W[4 * j + 2] = pk2_mul(W[4 * j + 2], pk2_ffma(a2, pk2_make(v.z, v.z), b2));
            W[4 * j + 3] = pk2_mul(W[4 * j + 3], pk2_ffma(a2, pk2_make(v.w, v.w), b2));
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive2(&empty_bar[stage]);
      if (++stage == n_stages) { stage = 0; phase ^= 1u; }
      if (active && !B2DBG(2)) {
        // CalcWeightTotal + fill, per (event, set): norms (reference order), osc, spline, static
        double* h0 = a.hist + static_cast<int64_t>(bin0 >= 0 ? bin0 : 0) * kB2Sets + set0;
        double* h1 = a.hist + static_cast<int64_t>(bin1 >= 0 ? bin1 : 0) * kB2Sets + set0;
        // two passes, so the 16 sets' shared-memory look-ups and products are independent of the atomics' ordering
        if (a.norm_uniform) {                              // the norm product is the same for every set: formed once
          float wn0 = 1.0f, wn1 = 1.0f;
          #pragma unroll
          for (int j = 0; j < 4; ++j) { wn0 *= s_norm[np0[j]]; wn1 *= s_norm[np1[j]]; }
          #pragma unroll
          for (int q = 0; q < kB2SW; ++q) {
            float w0 = wn0, w1 = wn1;
            w0 *= osc0; w0 *= pk2_lo(W[q]); w0 *= st0;
            w1 *= osc1; w1 *= pk2_hi(W[q]); w1 *= st1;
            W[q] = pk2_make(w0, w1);
          }
        } else {
          #pragma unroll
          for (int q = 0; q < kB2SW; ++q) {
            float w0 = 1.0f, w1 = 1.0f;
            #pragma unroll
            for (int j = 0; j < 4; ++j) { w0 *= s_norm[np0[j] + q * nnp]; w1 *= s_norm[np1[j] + q * nnp]; }
            w0 *= osc0; w0 *= pk2_lo(W[q]); w0 *= st0;
            w1 *= osc1; w1 *= pk2_hi(W[q]); w1 *= st1;
            W[q] = pk2_make(w0, w1);
          }
        }
        #pragma unroll
        for (int q = 0; q < kB2SW; ++q) {
          const float w0 = pk2_lo(W[q]), w1 = pk2_hi(W[q]);
          if (set0 + q < a.n_sets && !B2DBG(1)) {
            if (w0 > 0.f && bin0 >= 0) atomicAdd(h0 + q, static_cast<double>(w0));
            if (w1 > 0.f && bin1 >= 0) atomicAdd(h1 + q, static_cast<double>(w1));
          }
          W[q] = B2DBG(1) ? pk2_make(w0 * 0.f + 1.f, w1 * 0.f + 1.f) : one2;
        }
      }
    }
  }
}

// (-lnL per set: llh_batch_kernel of m3b_batch.cu, same [bin][256] histogram layout)

}  // namespace m3b

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
// returns M3B_OK and *done = 1 when the batch ran on this kernel; *done = 0 = the caller tries the first-generation kernel
int m3b_batch2_try(m3b_handle* h, int32_t n_sets, const double* spline_pars, const double* norm_pars, const float* osc_w,
                   double* host_slots_dev, int* done) {
  using namespace m3b;
  *done = 0;
  if (h->binned || !h->splines_done || h->first_time_w2 || h->cfg.update_w2 || n_sets > kB2Sets || h->T % kB2E != 0) return M3B_OK;
  if (h->cfg.flags & (M3B_FLAG_NO_BATCH_KERNEL | M3B_FLAG_BATCH_KERNEL_V1)) return M3B_OK;
  if (h->tiles_dirty || !h->d_tiles) return M3B_OK;     // first step has not run yet
  if (h->Kmax > 64 || h->norm_slots > 4) return M3B_OK;        // more norm parameters per event: first-generation kernel
  CK(cudaSetDevice(h->device));
  const int P = h->P, S = kB2Sets, n_sigs = static_cast<int>(h->sigs.size());
  // 1. segments of every set, in order (SplineBase::FindSplineSegment keeps its cached-segment history); the handle's
  //    cached segments are restored if this kernel declines, so the fallback starts from the same history
  const std::vector<int16_t> curr_save = h->curr_segment, seg_save = h->segments;
  const std::vector<float> val_save = h->param_values;
  std::vector<int16_t> seg(static_cast<size_t>(n_sets) * P);
  std::vector<float> val(static_cast<size_t>(n_sets) * P);
  for (int s = 0; s < n_sets; ++s) {
    int rc = m3b_find_segments(h, spline_pars + static_cast<size_t>(s) * P, seg.data() + static_cast<size_t>(s) * P, val.data() + static_cast<size_t>(s) * P);
    if (rc != M3B_OK) return rc;
  }
  auto decline = [&]() { h->curr_segment = curr_save; h->segments = seg_save; h->param_values = val_save; return M3B_OK; };
  // 2. shared-memory budget: the per-set tables + a ring of 2..4 stages of 32, 24 or
  //    16 rows, whatever fits (more stages first: the ring hides the bulk-copy latency, the row count only sets how many
  //    slots share one barrier round)
  const int Nn = h->n_norm_values;
  const int tables_bytes = batch2_tables_bytes(h->max_nc, h->max_nl, Nn);
  int n_stages = 0, stage_rows = 0;
  {
    const int budget = 232448 - 1024 - tables_bytes;
    const int cand[][2] = {{4, 32}, {3, 32}, {4, 24}, {3, 24}, {4, 16}, {2, 32}, {3, 16}, {2, 24}, {2, 16}};
    for (const auto& c : cand)
      if (c[0] * c[1] * kB2RowF4 * 16 <= budget && h->max_nl * kB2E * 8 <= c[1] * kB2RowF4 * 16) { n_stages = c[0]; stage_rows = c[1]; break; }
    if (n_stages == 0) return decline();                // too many parameters for the tables: first-generation kernel
  }
  // 3. per signature and slot: distinct segments -> staged rows (most popular first); per set: dx and its rank
  std::vector<Batch2Sig> bs(n_sigs);
  std::vector<float> t_dx, t_val; std::vector<uint32_t> t_code; std::vector<int32_t> t_rowlist, t_slot; std::vector<Batch2Group> t_group;
  std::vector<uint8_t> t_rank8;
  for (int g = 0; g < n_sigs; ++g) {
    const SigDesc& sd = h->sigs[g];
    const int32_t* pool = h->sig_pool.data() + sd.off;
    Batch2Sig& b = bs[g];
    b.nc = sd.nc; b.nl = sd.nl; b.n_wide = 0;
    b.off_rank8 = static_cast<int64_t>(t_rank8.size());
    b.off_dx = static_cast<int64_t>(t_dx.size()); b.off_code = static_cast<int64_t>(t_code.size());
    b.off_val = static_cast<int64_t>(t_val.size()); b.off_rowlist = static_cast<int64_t>(t_rowlist.size());
    b.off_slot = static_cast<int64_t>(t_slot.size()); b.off_group = static_cast<int64_t>(t_group.size());
    t_dx.resize(t_dx.size() + static_cast<size_t>(sd.nc) * S, 0.f);
    t_code.resize(t_code.size() + static_cast<size_t>(sd.nc) * kB2CW, 0u);
    t_val.resize(t_val.size() + static_cast<size_t>(sd.nl) * S, 0.f);
    int rows = 0;
    Batch2Group cur{0, 0, 0, 0};
    for (int c = 0; c < sd.nc; ++c) {
      const int p = pool[c], segbase = pool[sd.nc + c];
      int rank_of[64]; for (int k = 0; k < 64; ++k) rank_of[k] = -1;
      int used[64] = {0};
      for (int s = 0; s < n_sets; ++s) ++used[seg[static_cast<size_t>(s) * P + p]];
      int n_rank = 0;
      std::vector<int32_t> slot_rows;
      while (true) {                       // most popular segment first (rank 0 = the kernel's fall-through path)
        int best = -1;
        for (int k = 0; k < 64; ++k) if (used[k] > 0 && (best < 0 || used[k] > used[best])) best = k;
        if (best < 0) break;
        rank_of[best] = n_rank++; slot_rows.push_back(segbase + best); used[best] = 0;
      }
      if (n_rank > kB2MaxRanks || n_rank > stage_rows) return decline();
      int wide = -1;
      if (n_rank > 3) {                                  // more than the 2-bit code word holds: a rank byte per set
        if (b.n_wide == kB2MaxWide) return decline();
        wide = b.n_wide++;
        t_rank8.resize(t_rank8.size() + S, 0);
      }
      if (cur.n_rows + n_rank > stage_rows) {            // close the group: its rows fill one ring stage
        cur.c1 = c; t_group.push_back(cur);
        cur = Batch2Group{c, c, 0, rows};
      }
      const int slot_row = cur.n_rows;
      for (int32_t r : slot_rows) t_rowlist.push_back(r);
      cur.n_rows += n_rank; rows += n_rank;
      bool uniform = n_rank == 1;
      for (int s = 0; s < S; ++s) {
        const int ss = s < n_sets ? s : n_sets - 1;           // padding sets repeat the last set (never filled)
        const int sg = seg[static_cast<size_t>(ss) * P + p];
        if (wide >= 0) t_rank8[static_cast<size_t>(b.off_rank8) + static_cast<size_t>(wide) * S + s] = static_cast<uint8_t>(rank_of[sg]);
        else t_code[b.off_code + static_cast<size_t>(c) * kB2CW + s / kB2SW] |= static_cast<uint32_t>(rank_of[sg]) << (2 * (s % kB2SW));
        // dx = ParamValues[Param] - coeff_x[Param*_max_knots+segment]   (Splines/SplineMonolith.cpp:759), float
        const float dxs = val[static_cast<size_t>(ss) * P + p] - h->coeff_x[static_cast<size_t>(p) * h->Kmax + sg];
        t_dx[b.off_dx + static_cast<size_t>(c) * S + s] = dxs;
        uniform = uniform && std::memcmp(&dxs, &t_dx[b.off_dx + static_cast<size_t>(c) * S], sizeof(float)) == 0;
      }
      t_slot.push_back(slot_row); t_slot.push_back(n_rank); t_slot.push_back(wide); t_slot.push_back(uniform ? 1 : 0);
    }
    if (sd.nc > 0) { cur.c1 = sd.nc; t_group.push_back(cur); }
    b.n_groups = static_cast<int32_t>(t_group.size() - static_cast<size_t>(b.off_group));
    if (b.n_groups > kB2MaxGroups) return decline();
    b.lin_uniform = 0;
    for (int l = 0; l < sd.nl; ++l) {
      const int p = pool[2 * sd.nc + l];
      bool uniform = true;
      for (int s = 0; s < S; ++s) {
        const float v = val[static_cast<size_t>(s < n_sets ? s : n_sets - 1) * P + p];
        t_val[b.off_val + static_cast<size_t>(l) * S + s] = v;
        uniform = uniform && std::memcmp(&v, &t_val[b.off_val + static_cast<size_t>(l) * S], sizeof(float)) == 0;
      }
      if (uniform && l < 64) b.lin_uniform |= 1ull << l;
    }
  }
  std::vector<float> t_norm(static_cast<size_t>(std::max(Nn, 1)) * S, 1.f);
  bool norm_uniform = true;
  for (int n = 0; n < Nn; ++n)
    for (int s = 0; s < S; ++s) {
      const float v = static_cast<float>(norm_pars[static_cast<size_t>(s < n_sets ? s : n_sets - 1) * Nn + n]);
      t_norm[static_cast<size_t>(n) * S + s] = v;
      norm_uniform = norm_uniform && std::memcmp(&v, &t_norm[static_cast<size_t>(n) * S], sizeof(float)) == 0;
    }
  const int smem = tables_bytes + n_stages * stage_rows * kB2RowF4 * 16;
  // 4. device staging (grown on demand, kept in the handle; shared with the first kernel's buffers)
  auto grow = [&](void** p, size_t& cap, size_t bytes) -> cudaError_t {
    if (cap >= bytes && *p) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr; cap = 0;
    const cudaError_t e = cudaMalloc(p, bytes + bytes / 4 + 256);
    if (e == cudaSuccess) cap = bytes + bytes / 4 + 256;
    return e;
  };
  const size_t slot = static_cast<size_t>(1 + h->n_samples);
  CK(cudaStreamSynchronize(h->stream));          // a previous batch may still read the staging buffers
  CK(grow(&h->bt_dx, h->bt_dx_cap, t_dx.size() * 4 + 16));
  CK(grow(&h->bt_rowoff, h->bt_rowoff_cap, t_code.size() * 4 + 16));
  CK(grow(&h->bt_val, h->bt_val_cap, t_val.size() * 4 + 16));
  CK(grow(&h->bt_rowlist, h->bt_rowlist_cap, t_rowlist.size() * 4 + 16));
  CK(grow(&h->bt_slot, h->bt_slot_cap, t_slot.size() * 4 + 16));
  CK(grow(&h->bt_rank8, h->bt_rank8_cap, t_rank8.size() + 256));
  CK(grow(&h->bt_group, h->bt_group_cap, t_group.size() * sizeof(Batch2Group) + 16));
  CK(grow(&h->bt_norm, h->bt_norm_cap, t_norm.size() * 4));
  CK(grow(&h->bt_sigs, h->bt_sigs_cap, bs.size() * sizeof(Batch2Sig)));
  CK(grow(&h->bt_hist, h->bt_hist_cap, static_cast<size_t>(S) * h->n_bins * 8));
  CK(grow(&h->bt_llh, h->bt_llh_cap, static_cast<size_t>(S) * slot * 8));
  if (!t_dx.empty()) CK(cudaMemcpyAsync(h->bt_dx, t_dx.data(), t_dx.size() * 4, cudaMemcpyHostToDevice, h->stream));
  if (!t_code.empty()) CK(cudaMemcpyAsync(h->bt_rowoff, t_code.data(), t_code.size() * 4, cudaMemcpyHostToDevice, h->stream));
  if (!t_val.empty()) CK(cudaMemcpyAsync(h->bt_val, t_val.data(), t_val.size() * 4, cudaMemcpyHostToDevice, h->stream));
  if (!t_rowlist.empty()) CK(cudaMemcpyAsync(h->bt_rowlist, t_rowlist.data(), t_rowlist.size() * 4, cudaMemcpyHostToDevice, h->stream));
  if (!t_slot.empty()) CK(cudaMemcpyAsync(h->bt_slot, t_slot.data(), t_slot.size() * 4, cudaMemcpyHostToDevice, h->stream));
  if (!t_rank8.empty()) CK(cudaMemcpyAsync(h->bt_rank8, t_rank8.data(), t_rank8.size(), cudaMemcpyHostToDevice, h->stream));
  if (!t_group.empty()) CK(cudaMemcpyAsync(h->bt_group, t_group.data(), t_group.size() * sizeof(Batch2Group), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->bt_norm, t_norm.data(), t_norm.size() * 4, cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(h->bt_sigs, bs.data(), bs.size() * sizeof(Batch2Sig), cudaMemcpyHostToDevice, h->stream));
  if (osc_w) {
    REQUIRE(h->use_osc, M3B_ERR_INVALID, "step: osc_w given but events were uploaded with use_osc=0");
    CK(cudaMemcpyAsync(h->d_osc, osc_w, sizeof(float) * h->n_osc, cudaMemcpyHostToDevice, h->stream));
  }
  CK(cudaMemsetAsync(h->bt_hist, 0, static_cast<size_t>(S) * h->n_bins * 8, h->stream));
  CK(cudaMemsetAsync(h->d_tile_counter, 0, sizeof(unsigned int), h->stream));
  CK(cudaStreamSynchronize(h->stream));          // the host staging vectors go out of scope below

  Batch2Args a{};
  a.tiles = h->d_tiles; a.units_per_tile = h->T / kB2E; a.n_units = static_cast<int32_t>(h->n_tiles) * a.units_per_tile; a.T = h->T;
  a.sigs = static_cast<const Batch2Sig*>(h->bt_sigs);
  a.t_dx = static_cast<const float*>(h->bt_dx); a.t_code = static_cast<const uint32_t*>(h->bt_rowoff);
  a.t_val = static_cast<const float*>(h->bt_val); a.t_rowlist = static_cast<const int32_t*>(h->bt_rowlist);
  a.t_norm = static_cast<const float*>(h->bt_norm); a.t_slot = static_cast<const int32_t*>(h->bt_slot);
  a.t_group = static_cast<const Batch2Group*>(h->bt_group); a.t_rank8 = static_cast<const uint8_t*>(h->bt_rank8);
  a.n_norm = Nn; a.n_sets = n_sets; a.max_nc = h->max_nc; a.max_nl = h->max_nl;
  a.n_stages = n_stages; a.stage_rows = stage_rows; a.norm_uniform = norm_uniform ? 1 : 0;
  { const char* dbg = experiment_env("M3B_BATCH_DBG"); a.dbg = dbg ? atoi(dbg) : 0; }
  a.bin = h->d_bin; a.osc = h->use_osc ? h->d_osc : nullptr; a.osc_idx = h->d_osc_idx; a.static_w = h->d_static;
  a.norm_idx = h->d_norm_idx; a.norm_slots = h->norm_slots; a.e_pad = h->e_pad; a.n_events = h->n_events;
  a.hist = static_cast<double*>(h->bt_hist); a.n_bins = h->n_bins; a.counter = h->d_tile_counter;
  CK(allow_max_dynamic_smem(fill_batch2_kernel));
  const int grid = static_cast<int>(std::min<int64_t>(a.n_units, h->sm_count));
  if (h->timing) {
    if (h->tev_used + 2 > h->tev.size()) for (int i = 0; i < 2; ++i) { cudaEvent_t e; CK(cudaEventCreate(&e)); h->tev.push_back(e); }
    CK(cudaEventRecord(h->tev[h->tev_used], h->stream));
  }
  fill_batch2_kernel<<<grid, kB2CT, smem, h->stream>>>(a);
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
