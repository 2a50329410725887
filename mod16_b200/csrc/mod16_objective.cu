// K2 -- objective batch: D parameter draws x E = T x N elements, reduced per draw to
// sum(((pred - obs) w)^2) and the number of valid pairs (include/mod16_b200.h).
//
// float32 entry (static semantics, D >= 32) -- five launches on the caller's stream:
//   1. obj_classify_kernel  one thread per element: valid (obs, w not NaN)?  inside the
//                           domain of the rearranged float algebra for EVERY draw?
//   2. obj_scan_kernel      exclusive scan of the per-block counts (deterministic, stable
//                           compaction -- the summation order never depends on timing)
//   3. obj_pack_kernel      writes the draw-independent coefficients of every fast element
//                           (obj::ElemRec, 160 B) and the index list of the exact elements
//   4. objective_fast_kernel  the hot loop.  Thread = two draws (the two lanes of the
//                           packed-FP32 instructions), block = 512 draws x one chunk of the
//                           element stream.  A producer warp streams the records through a
//                           four-stage shared-memory ring with one bulk copy per tile
//                           (cp.async.bulk + mbarrier); the eight consumer warps read each
//                           record with warp-uniform 128-bit loads, so an element's
//                           coefficients are scalar operands and element properties are
//                           uniform branches.  Bound: FP32 issue (~65 lane-operations and 6
//                           MUFU per draw-element; nothing is re-read from HBM but 160 B per
//                           512 draw-elements).
//      objective_exact_kernel  the exact elements x all draws, op-for-op in double
//   5. objective_reduce_kernel  per draw: sums the chunk partials in a fixed order, applies
//                           the draw's transpiration gate (any g_surf > 0 over ITS elements,
//                           :343), lists the draws that are outside the fast domain
//   6. objective_exact_kernel  those draws x all elements in double (normally none).
// float64 entry, instance semantics, D < 32: objective_exact_kernel over everything
// (block = one draw x one element chunk, warp-shuffle + shared-memory block reduction).
#include "mod16_common.cuh"
#include "mod16_fast.cuh"
#include "mod16_objective.cuh"
#include "mod16_pipe.cuh"

namespace mod16 {

using fast::F2;
using obj::ElemRec;

constexpr int OBJ_CONSUMER_WARPS = 8;
constexpr int OBJ_CONSUMERS = OBJ_CONSUMER_WARPS * 32;
constexpr int OBJ_THREADS = OBJ_CONSUMERS + 32;
constexpr int OBJ_DRAWS_PER_BLOCK = OBJ_CONSUMERS * 2;
constexpr int OBJ_TILE = 32;    // elements per stage (5 KB)
constexpr int OBJ_STAGES = 4;
constexpr int OBJ_CLS_THREADS = 256;
constexpr int OBJ_EX_THREADS = 128;
constexpr int OBJ_MIN_FAST_DRAWS = 32;

template <typename R>
struct ObjArgs {
  Arr<R> drv[MOD16_N_DRIVERS];
  Arr<R> obs, w;
  const double* draws;
  int64_t T, N, E, D;
  int gate_mode;  // 0: per-draw global gate, 1: open, 2: closed
  // workspace
  uint8_t* cls;            // per element: 0 invalid, 1 fast, 2 exact
  int32_t* blk_fast;       // per classify block: count, then exclusive offset
  int32_t* blk_exact;
  int64_t* counts;         // [0] n_fast, [1] n_exact, [2] n_bad
  ElemRec* recs;
  uint32_t* exact_list;
  double* part_fast;       // [nf_chunks][D][3]: sse open, sse closed, max gate quantity
  double* part_exact;      // [ne_slots][D][5]: sse open, n open, sse closed, n closed, any
  int32_t* bad_list;
  int nf_chunks, ne_slots;
  // outputs
  double* sse;
  double* count;
  int32_t* gate_out;
  double* raw;             // (D, 5) un-gated statistics, or NULL
};

template <typename R>
__device__ __forceinline__ void load_element(const ObjArgs<R>& a, int64_t e, Drivers<double>& d,
                                             double& ob, double& w) {
  const int64_t t = e / a.N, n = e - t * a.N;
  d.lw_d = (double)a.drv[MOD16_D_LW_NET_DAY].at(t, n);
  d.lw_n = (double)a.drv[MOD16_D_LW_NET_NIGHT].at(t, n);
  d.sw_d = (double)a.drv[MOD16_D_SW_RAD_DAY].at(t, n);
  d.sw_n = (double)a.drv[MOD16_D_SW_RAD_NIGHT].at(t, n);
  d.alb = (double)a.drv[MOD16_D_SW_ALBEDO].at(t, n);
  d.t_d = (double)a.drv[MOD16_D_TEMP_DAY].at(t, n);
  d.t_n = (double)a.drv[MOD16_D_TEMP_NIGHT].at(t, n);
  d.t_ann = (double)a.drv[MOD16_D_TEMP_ANNUAL].at(t, n);
  d.tmin = (double)a.drv[MOD16_D_TMIN].at(t, n);
  d.vpd_d = (double)a.drv[MOD16_D_VPD_DAY].at(t, n);
  d.vpd_n = (double)a.drv[MOD16_D_VPD_NIGHT].at(t, n);
  d.p = (double)a.drv[MOD16_D_PRESSURE].at(t, n);
  d.fpar = (double)a.drv[MOD16_D_FPAR].at(t, n);
  d.lai = (double)a.drv[MOD16_D_LAI].at(t, n);
  d.rc_d = 0.0; d.rc_n = 0.0;
  ob = (double)a.obs.at(t, n);
  w = (double)a.w.at(t, n);
}

__device__ __forceinline__ fast::StepIn to_step_in(const Drivers<double>& d) {
  fast::StepIn in;
  in.lw_d = (float)d.lw_d; in.lw_n = (float)d.lw_n; in.sw_d = (float)d.sw_d;
  in.sw_n = (float)d.sw_n; in.alb = (float)d.alb; in.t_d = (float)d.t_d; in.t_n = (float)d.t_n;
  in.t_ann = (float)d.t_ann; in.tmin = (float)d.tmin; in.vpd_d = (float)d.vpd_d;
  in.vpd_n = (float)d.vpd_n; in.p = (float)d.p; in.fpar = (float)d.fpar; in.lai = (float)d.lai;
  in.rc_d = 0.0f; in.rc_n = 0.0f;
  return in;
}

// ---- 1-3: classification and stable compaction (float32 entry) ---------------------------
__global__ void __launch_bounds__(OBJ_CLS_THREADS)
obj_classify_kernel(const __grid_constant__ ObjArgs<float> a) {
  const int64_t e = (int64_t)blockIdx.x * OBJ_CLS_THREADS + threadIdx.x;
  int c = 0;
  if (e < a.E) {
    Drivers<double> d;
    double ob, w;
    load_element<float>(a, e, d, ob, w);
    if (ob == ob && w == w) {
      ElemRec r;
      const fast::StepIn in = to_step_in(d);
      c = obj::make_elem_rec(in, (float)ob, (float)w, r) ? 1 : 2;
      if (c == 2) {
        // masked (NaN) drivers: the prediction is NaN for every draw, so the pair is not
        // valid for any of them -- same as a missing observation.  (Elements dropped here
        // still belong to each draw's `_et` call and so to its gate, :343: the exact
        // kernel's passes over ALL elements evaluate their g_surf, see objective_exact_kernel.)
        fast::PftConstV<float> sane;
        sane.tmin_bias = -26.5f; sane.inv_dtmin = 0.1f; sane.vpd_off = 3.0f;
        sane.inv_dvpd = 1e-3f; sane.csl = 0.0f; sane.ok = 1.0f;
        bool gp;
        if (fast::nan_halves<SEM_STATIC>(in, sane, false, gp) != 0) c = 0;
      }
    }
    a.cls[e] = (uint8_t)c;
  }
  const int nf = __syncthreads_count(c == 1);
  const int ne = __syncthreads_count(c == 2);
  if (threadIdx.x == 0) {
    a.blk_fast[blockIdx.x] = nf;
    a.blk_exact[blockIdx.x] = ne;
  }
}

// One block: in-place exclusive scan of both count arrays, totals to counts[0..1].
__global__ void __launch_bounds__(1024)
obj_scan_kernel(int32_t* blk_fast, int32_t* blk_exact, int64_t nblk, int64_t* counts) {
  __shared__ long long warp_tot[2][32];
  __shared__ long long carry[2];
  if (threadIdx.x < 2) carry[threadIdx.x] = 0;
  if (threadIdx.x == 0) counts[2] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t base = 0; base < nblk; base += 1024) {
    const int64_t i = base + threadIdx.x;
    long long v[2] = {i < nblk ? blk_fast[i] : 0, i < nblk ? blk_exact[i] : 0};
    long long inc[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      long long x = v[k];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      inc[k] = x;
      if (lane == 31) warp_tot[k][wid] = x;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        long long x = warp_tot[k][lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const long long y = __shfl_up_sync(0xffffffffu, x, o);
          if (lane >= o) x += y;
        }
        warp_tot[k][lane] = x;  // inclusive over warps
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const long long before = carry[k] + (wid ? warp_tot[k][wid - 1] : 0) + inc[k] - v[k];
      if (i < nblk) (k == 0 ? blk_fast : blk_exact)[i] = (int32_t)before;
    }
    __syncthreads();
    if (threadIdx.x < 2) carry[threadIdx.x] += warp_tot[threadIdx.x][31];
    __syncthreads();
  }
  if (threadIdx.x < 2) counts[threadIdx.x] = carry[threadIdx.x];
}

__global__ void __launch_bounds__(OBJ_CLS_THREADS)
obj_pack_kernel(const __grid_constant__ ObjArgs<float> a) {
  __shared__ int warp_off[2][OBJ_CLS_THREADS / 32];
  const int64_t e = (int64_t)blockIdx.x * OBJ_CLS_THREADS + threadIdx.x;
  const int c = e < a.E ? (int)a.cls[e] : 0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned mf = __ballot_sync(0xffffffffu, c == 1);
  const unsigned me = __ballot_sync(0xffffffffu, c == 2);
  if (lane == 0) { warp_off[0][wid] = __popc(mf); warp_off[1][wid] = __popc(me); }
  __syncthreads();
  if (threadIdx.x < 2) {  // exclusive scan over the eight warps
    int run = 0;
    for (int k = 0; k < OBJ_CLS_THREADS / 32; ++k) {
      const int v = warp_off[threadIdx.x][k];
      warp_off[threadIdx.x][k] = run;
      run += v;
    }
  }
  __syncthreads();
  const unsigned below = (1u << lane) - 1u;
  if (c == 1) {
    const int64_t pos = (int64_t)a.blk_fast[blockIdx.x] + warp_off[0][wid] + __popc(mf & below);
    Drivers<double> d;
    double ob, w;
    load_element<float>(a, e, d, ob, w);
    ElemRec r;
    obj::make_elem_rec(to_step_in(d), (float)ob, (float)w, r);
    float4* dst = reinterpret_cast<float4*>(a.recs + pos);
    const float4* src = reinterpret_cast<const float4*>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(ElemRec) / 16); ++k) dst[k] = src[k];
  } else if (c == 2) {
    const int64_t pos = (int64_t)a.blk_exact[blockIdx.x] + warp_off[1][wid] + __popc(me & below);
    a.exact_list[pos] = (uint32_t)e;
  }
}

// ---- 4: the hot loop ----------------------------------------------------------------------
#ifndef M16_OBJ_BLOCKS
#define M16_OBJ_BLOCKS 4
#endif
#ifndef M16_OBJ_UNROLL
#define M16_OBJ_UNROLL 2
#endif
constexpr int OBJ_UNROLL = M16_OBJ_UNROLL;
// MODE 0: gate-open sums only (a draw whose gate turns out closed is redone by the exact
// kernel: it needs csl == 0 or a Tmin ramp above every Tmin); 1: gate-closed sums only
// (MOD16_GATE_CLOSED); 2: both (the un-gated statistics of a site-sharded batch).
template <int MODE>
__global__ void __launch_bounds__(OBJ_THREADS, M16_OBJ_BLOCKS)
objective_fast_kernel(const __grid_constant__ ObjArgs<float> a) {
  __shared__ __align__(128) ElemRec ring[OBJ_STAGES][OBJ_TILE];
  __shared__ __align__(8) uint64_t full[OBJ_STAGES], empty[OBJ_STAGES];
  const int tid = threadIdx.x;
  const int64_t n_fast = a.counts[0];
  const int64_t n_tiles = (n_fast + OBJ_TILE - 1) / OBJ_TILE;
  const int64_t per_chunk = (n_tiles + gridDim.y - 1) / gridDim.y;
  const int64_t tile_lo = (int64_t)blockIdx.y * per_chunk;
  int64_t tile_hi = tile_lo + per_chunk;
  if (tile_hi > n_tiles) tile_hi = n_tiles;
  const int niter = tile_hi > tile_lo ? (int)(tile_hi - tile_lo) : 0;
  const uint32_t full0 = smem_u32(&full[0]), empty0 = smem_u32(&empty[0]);
  const uint32_t data0 = smem_u32(&ring[0][0]);
  constexpr uint32_t STAGEB = OBJ_TILE * sizeof(ElemRec);
  const bool producer = tid >= OBJ_CONSUMERS;

  if (tid == OBJ_CONSUMERS) {
    for (int s = 0; s < OBJ_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, OBJ_CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (producer) {
    int stage = 0;
    uint32_t phase = 0;
#pragma unroll 1
    for (int i = 0; i < niter; ++i) {
      if (i >= OBJ_STAGES) mbar_wait(empty0 + 8 * stage, phase ^ 1u);
      if (elect_one()) {
        const int64_t e0 = (tile_lo + i) * OBJ_TILE;
        const int64_t left = n_fast - e0;
        const uint32_t bytes = (uint32_t)((left < OBJ_TILE ? left : OBJ_TILE) * sizeof(ElemRec));
        mbar_expect_tx(full0 + 8 * stage, bytes);
        bulk_g2s(data0 + (uint32_t)stage * STAGEB, a.recs + e0, bytes, full0 + 8 * stage);
      }
      __syncwarp();
      if (++stage == OBJ_STAGES) { stage = 0; phase ^= 1u; }
    }
    return;
  }

  // this thread's two draws
  const int64_t d0 = ((int64_t)blockIdx.x * OBJ_CONSUMERS + tid) * 2;
  obj::DrawConstV<F2> dc;
  {
    obj::DrawConstV<float> c[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      int64_t d = d0 + l;
      if (d >= a.D) d = a.D - 1;  // idle lanes repeat the last draw and never write
      double v[MOD16_N_PARAMS];
#pragma unroll
      for (int k = 0; k < MOD16_N_PARAMS; ++k) v[k] = __ldg(a.draws + d * MOD16_N_PARAMS + k);
      obj::make_draw_const(v, c[l]);
    }
    dc = obj::pair_draws(c[0], c[1]);
  }
  double sum_o[2] = {0.0, 0.0}, sum_c[2] = {0.0, 0.0};
  F2 gmax = F2{0.0f, 0.0f};
  int stage = 0;
  uint32_t phase = 0;
#pragma unroll 1
  for (int i = 0; i < niter; ++i) {
    const int64_t left = n_fast - (tile_lo + i) * OBJ_TILE;
    const int cnt = left < OBJ_TILE ? (int)left : OBJ_TILE;
    mbar_wait(full0 + 8 * stage, phase);
    F2 acc_o = F2{0.0f, 0.0f}, acc_c = F2{0.0f, 0.0f};
#pragma unroll OBJ_UNROLL
    for (int j = 0; j < cnt; ++j) {
      const ElemRec& e = ring[stage][j];
      F2 tr, gs;
      const F2 tot = obj::objective_step<F2>(e, dc, tr, gs);
      gmax = fast::vmax(gmax, gs);
      const float w = e.w;
      const F2 eo = fast::vfma(tot, w, -e.obs_w);  // (pred - obs) w, gate open
      if (MODE != 1) acc_o = fast::vfma(eo, eo, acc_o);
      if (MODE != 0) {
        const F2 ec = fast::vfma(-tr, w, eo);      // the same without transpiration
        acc_c = fast::vfma(ec, ec, acc_c);
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(empty0 + 8 * stage);
    sum_o[0] += (double)acc_o.x; sum_o[1] += (double)acc_o.y;
    sum_c[0] += (double)acc_c.x; sum_c[1] += (double)acc_c.y;
    if (++stage == OBJ_STAGES) { stage = 0; phase ^= 1u; }
  }
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    const int64_t d = d0 + l;
    if (d < a.D) {
      double* p = a.part_fast + ((int64_t)blockIdx.y * a.D + d) * 3;
      p[0] = sum_o[l];
      p[1] = sum_c[l];
      p[2] = (double)(l == 0 ? gmax.x : gmax.y);
    }
  }
}

// ---- exact evaluation: one block = one draw x one element chunk ---------------------------
// mode 0: all elements, chunked over gridDim.x; partials to part_exact[blockIdx.x][d]
// mode 1: the exact-element list (counts[1] entries), gridDim.x == 1; partials to slot 0
// mode 2: the listed bad draws x all elements, gridDim.x == 1; writes the final outputs
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void finish_draw(int gate_mode, bool is_static, double sse_o,
                                            double n_o, double sse_c, double n_c, bool any,
                                            double& sse, double& cnt, int& gate,
                                            double* raw = nullptr) {
  if (raw) {
    raw[0] = sse_o; raw[1] = n_o;
    raw[2] = is_static ? sse_c : sse_o; raw[3] = is_static ? n_c : n_o;
    raw[4] = any ? 1.0 : 0.0;
  }
  bool open = true;
  if (is_static) open = gate_mode == 1 || (gate_mode == 0 && any);
  sse = open ? sse_o : sse_c;
  cnt = open ? n_o : n_c;
  gate = any ? 1 : 0;
}

template <typename R, int SEM>
__global__ void __launch_bounds__(OBJ_EX_THREADS)
objective_exact_kernel(const __grid_constant__ ObjArgs<R> a, const int mode) {
  __shared__ double red[5][OBJ_EX_THREADS / 32];
  int64_t n_items, lo, hi;
  if (mode == 1) {
    n_items = a.counts[1];
    lo = 0; hi = n_items;
  } else {
    n_items = a.E;
    const int64_t per = (n_items + gridDim.x - 1) / gridDim.x;
    lo = (int64_t)blockIdx.x * per;
    hi = lo + per < n_items ? lo + per : n_items;
  }
  const int64_t n_draw_items = mode == 2 ? a.counts[2] : a.D;
  for (int64_t dj = blockIdx.y; dj < n_draw_items; dj += gridDim.y) {
    const int64_t d = mode == 2 ? (int64_t)a.bad_list[dj] : dj;
    Params<double> q;
    {
      const double* v = a.draws + d * MOD16_N_PARAMS;
      q.tmin_close = v[0]; q.tmin_open = v[1]; q.vpd_open = v[2]; q.vpd_close = v[3];
      q.gl_sh = v[4]; q.gl_wv = v[5]; q.g_cuticular = v[6]; q.csl = v[7];
      q.rbl_min = v[8]; q.rbl_max = v[9]; q.beta = v[10];
      q.tann_lo = 273.15 + (SEM == SEM_STATIC ? v[1] : v[0]);
    }
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int64_t i = lo + threadIdx.x; i < hi; i += OBJ_EX_THREADS) {
      const int64_t e = mode == 1 ? (int64_t)a.exact_list[i] : i;
      Drivers<double> dr;
      double ob, w;
      load_element<R>(a, e, dr, ob, w);
      Half<double> day, night;
      bool g_pos;
      if (!(ob == ob) || !(w == w)) {
        // No valid pair -- but the element is part of the draw's `_et` call, so its g_surf
        // takes part in the batch-global gate `np.any(g_surf > 0)` (:343) like any other.
        // Only the passes over ALL elements see it (mode 1 walks the compacted list of
        // valid elements; a draw no valid element opens is redone by mode 2, see the
        // reduce kernel), and only a thread that has not seen an open gate yet needs it.
        if (SEM == SEM_STATIC && mode != 1 && a.gate_mode == 0 && acc[4] == 0.0) {
          pixel_step<double, SEM>(dr, q, false, GATE_OPEN, day, night, g_pos);
          if (g_pos) acc[4] = 1.0;
        }
        continue;
      }
      pixel_step<double, SEM>(dr, q, false, GATE_OPEN, day, night, g_pos);
      const double tn = half_total<double, SEM>(night);
      const double x_o = ((half_total<double, SEM>(day) + tn) - ob) * w;   // :193
      if (x_o == x_o) { acc[0] += x_o * x_o; acc[1] += 1.0; }
      if (SEM == SEM_STATIC) {
        Half<double> dc = day;
        dc.trans = 0.0;                                                     // :348
        const double x_c = ((half_total<double, SEM>(dc) + tn) - ob) * w;
        if (x_c == x_c) { acc[2] += x_c * x_c; acc[3] += 1.0; }
        if (g_pos) acc[4] = 1.0;
      }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const double s = warp_sum_d(acc[k]);
      if (lane == 0) red[k][wid] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot[5];
      for (int k = 0; k < 5; ++k) {
        double s = 0.0;
        for (int wv = 0; wv < OBJ_EX_THREADS / 32; ++wv) s += red[k][wv];
        tot[k] = s;
      }
      if (mode == 2) {
        double sse, cnt;
        int gate;
        finish_draw(a.gate_mode, SEM == SEM_STATIC, tot[0], tot[1], tot[2], tot[3], tot[4] > 0.0,
                    sse, cnt, gate, a.raw ? a.raw + d * 5 : nullptr);
        a.sse[d] = sse;
        a.count[d] = cnt;
        if (a.gate_out) a.gate_out[d] = gate;
      } else {
        double* p = a.part_exact + ((int64_t)(mode == 1 ? 0 : blockIdx.x) * a.D + d) * 5;
        for (int k = 0; k < 5; ++k) p[k] = tot[k];
      }
    }
    __syncthreads();
  }
}

// ---- 5: per-draw reduction of the partials -------------------------------------------------
__global__ void __launch_bounds__(256)
objective_reduce_kernel(const __grid_constant__ ObjArgs<float> a, const int is_static,
                        const int use_fast, const int fast_mode) {
  const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= a.D) return;
  double sse_o = 0.0, sse_c = 0.0, n_o = 0.0, n_c = 0.0;
  bool any = false, bad = false;
  if (use_fast) {
    double gq = 0.0;
    for (int c = 0; c < a.nf_chunks; ++c) {
      const double* p = a.part_fast + ((int64_t)c * a.D + d) * 3;
      sse_o += p[0]; sse_c += p[1];
      gq = p[2] > gq ? p[2] : gq;
    }
    any = gq > 0.0;
    n_o = n_c = (double)a.counts[0];
    double v[MOD16_N_PARAMS];
    for (int k = 0; k < MOD16_N_PARAMS; ++k) v[k] = a.draws[d * MOD16_N_PARAMS + k];
    obj::DrawConstV<float> dc;
    bad = !obj::make_draw_const(v, dc);
    // a non-finite float sum: let the double code decide what the value is
    bad = bad || !(fabs(sse_o) <= 1.7e308) || !(fabs(sse_c) <= 1.7e308);
  }
  for (int s = 0; s < a.ne_slots; ++s) {
    const double* p = a.part_exact + ((int64_t)s * a.D + d) * 5;
    sse_o += p[0]; n_o += p[1]; sse_c += p[2]; n_c += p[3];
    any = any || p[4] > 0.0;
  }
  // No valid element opened this draw's gate.  Two reasons to hand it to the exact kernel's
  // pass over ALL elements (mode 2): in MODE 0 the fast kernel did not accumulate the sums a
  // closed gate selects; and in every mode the elements dropped at classification (missing
  // observation or weight, masked drivers) are part of the draw's `_et` call and may still
  // open its gate (:343) -- only that pass sees them.  Such draws are rare (csl == 0, or a
  // Tmin ramp above every Tmin), so the pass costs nothing measurable.
  if (use_fast && is_static && a.gate_mode == 0 && !any) bad = true;
  double sse, cnt;
  int gate;
  finish_draw(a.gate_mode, is_static != 0, sse_o, n_o, sse_c, n_c, any, sse, cnt, gate,
              a.raw ? a.raw + d * 5 : nullptr);
  a.sse[d] = sse;
  a.count[d] = cnt;
  if (a.gate_out) a.gate_out[d] = gate;
  if (bad) {
    const unsigned long long slot = atomicAdd((unsigned long long*)(a.counts + 2), 1ULL);
    a.bad_list[slot] = (int32_t)d;  // order is irrelevant: each bad draw is redone on its own
  }
}

// ---- host side -----------------------------------------------------------------------------
struct ObjPlan {
  int64_t E, nblk;
  int nf_chunks, ne_chunks;
  bool fast;
  int64_t off_cls, off_blk_fast, off_blk_exact, off_counts, off_recs, off_list, off_part_fast,
      off_part_exact, off_bad, total;
};

static int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

static ObjPlan make_plan(int64_t T, int64_t N, int64_t D, bool want_fast) {
  ObjPlan p;
  p.E = T * N;
  p.fast = want_fast && D >= OBJ_MIN_FAST_DRAWS;
  p.nblk = (p.E + OBJ_CLS_THREADS - 1) / OBJ_CLS_THREADS;
  const int64_t draw_blocks = (D + OBJ_DRAWS_PER_BLOCK - 1) / OBJ_DRAWS_PER_BLOCK;
  const int64_t tiles = (p.E + OBJ_TILE - 1) / OBJ_TILE;
  int64_t nf = (148 * M16_OBJ_BLOCKS * 2 + draw_blocks - 1) / (draw_blocks > 0 ? draw_blocks : 1);
  if (nf > tiles) nf = tiles;
  if (nf > 65535) nf = 65535;
  if (nf < 1) nf = 1;
  p.nf_chunks = (int)nf;
  int64_t ne = (148 * 16 + D - 1) / (D > 0 ? D : 1);
  const int64_t max_ne = (p.E + 4 * OBJ_EX_THREADS - 1) / (4 * OBJ_EX_THREADS);
  if (ne > max_ne) ne = max_ne;
  if (ne < 1) ne = 1;
  p.ne_chunks = (int)ne;
  int64_t o = 0;
  p.off_counts = o; o = align256(o + 4 * 8);
  p.off_bad = o; o = align256(o + 4 * D);
  p.off_part_exact = o; o = align256(o + 8 * 5 * D * (p.fast ? 1 : p.ne_chunks));
  p.off_cls = p.off_blk_fast = p.off_blk_exact = p.off_recs = p.off_list = p.off_part_fast = o;
  if (p.fast) {
    p.off_cls = o; o = align256(o + p.E);
    p.off_blk_fast = o; o = align256(o + 4 * p.nblk);
    p.off_blk_exact = o; o = align256(o + 4 * p.nblk);
    p.off_list = o; o = align256(o + 4 * p.E);
    p.off_part_fast = o; o = align256(o + 8 * 3 * D * p.nf_chunks);
    p.off_recs = o; o = align256(o + (int64_t)sizeof(ElemRec) * p.E);
  }
  p.total = o;
  return p;
}

template <typename R>
static int fill_obj_args(const mod16_drivers_t* drv, const mod16_objective_t* ob, int64_t T,
                         int64_t N, int64_t D, uint32_t flags, const ObjPlan& p, ObjArgs<R>& a) {
  for (int k = 0; k < MOD16_N_DRIVERS; ++k) {
    if (drv->drv[k].t_stride < 0 || drv->drv[k].n_stride < 0) return MOD16_ERR_SHAPE;
    fill_arr<R>(a.drv[k], drv->drv[k]);
  }
  if (ob->observed.t_stride < 0 || ob->observed.n_stride < 0 || ob->weights.t_stride < 0 ||
      ob->weights.n_stride < 0)
    return MOD16_ERR_SHAPE;
  fill_arr<R>(a.obs, ob->observed);
  fill_arr<R>(a.w, ob->weights);
  a.draws = ob->draws;
  a.T = T; a.N = N; a.E = p.E; a.D = D;
  const uint32_t gm = flags & MOD16_GATE_MASK;
  a.gate_mode = gm == MOD16_GATE_GLOBAL ? 0 : (gm == MOD16_GATE_OPEN ? 1 : 2);
  unsigned char* ws = (unsigned char*)ob->workspace;
  a.cls = ws + p.off_cls;
  a.blk_fast = (int32_t*)(ws + p.off_blk_fast);
  a.blk_exact = (int32_t*)(ws + p.off_blk_exact);
  a.counts = (int64_t*)(ws + p.off_counts);
  a.recs = (ElemRec*)(ws + p.off_recs);
  a.exact_list = (uint32_t*)(ws + p.off_list);
  a.part_fast = (double*)(ws + p.off_part_fast);
  a.part_exact = (double*)(ws + p.off_part_exact);
  a.bad_list = (int32_t*)(ws + p.off_bad);
  a.nf_chunks = p.nf_chunks;
  a.ne_slots = p.fast ? 1 : p.ne_chunks;
  a.sse = ob->sse; a.count = ob->count; a.gate_out = ob->gate; a.raw = ob->raw;
  return MOD16_OK;
}

template <typename R>
static int run_objective(const mod16_drivers_t* drv, const mod16_objective_t* ob, int64_t T,
                         int64_t N, int64_t D, uint32_t flags, cudaStream_t st) {
  if (!drv || !ob) return MOD16_ERR_NULL;
  if (T < 0 || N < 0 || D < 0) return MOD16_ERR_SHAPE;
  const uint32_t sem = flags & MOD16_SEM_MASK, gm = flags & MOD16_GATE_MASK;
  if (sem > MOD16_SEM_INSTANCE) return MOD16_ERR_FLAGS;
  if (gm != MOD16_GATE_GLOBAL && gm != MOD16_GATE_OPEN && gm != MOD16_GATE_CLOSED)
    return MOD16_ERR_FLAGS;
  if (flags & ~(MOD16_SEM_MASK | MOD16_GATE_MASK | MOD16_FORCE_MIRROR)) return MOD16_ERR_FLAGS;
  if (drv->has_r_corr) return MOD16_ERR_FLAGS;
  if (D == 0) return MOD16_OK;
  if (!ob->draws || !ob->sse || !ob->count) return MOD16_ERR_NULL;
  if (T * N >= 4294967295LL) return MOD16_ERR_SHAPE;
  const bool is_static = sem == MOD16_SEM_STATIC;
  const bool want_fast = sizeof(R) == 4 && is_static && !(flags & MOD16_FORCE_MIRROR);
  const ObjPlan p = make_plan(T, N, D, want_fast);
  if (!ob->workspace || (reinterpret_cast<uintptr_t>(ob->workspace) & 255)) return MOD16_ERR_NULL;
  if (ob->workspace_bytes < p.total) return MOD16_ERR_WORKSPACE;
  ObjArgs<R> a;
  int rc = fill_obj_args<R>(drv, ob, T, N, D, flags, p, a);
  if (rc != MOD16_OK) return rc;
  M16_CUDA_CHECK(cudaMemsetAsync(a.counts, 0, 4 * 8, st));
  // the reduce kernel only reads ObjArgs fields that do not depend on R
  ObjArgs<float> ar;
  {
    mod16_drivers_t none = *drv;
    rc = fill_obj_args<float>(&none, ob, T, N, D, flags, p, ar);
    if (rc != MOD16_OK) return rc;
  }
  const unsigned draw_rows = (unsigned)(D < 32768 ? D : 32768);
  const int fast_mode = ob->raw ? 2 : (a.gate_mode == 2 ? 1 : 0);
  if (p.E == 0) {
    M16_CUDA_CHECK(cudaMemsetAsync(ob->sse, 0, 8 * D, st));
    M16_CUDA_CHECK(cudaMemsetAsync(ob->count, 0, 8 * D, st));
    if (ob->gate) M16_CUDA_CHECK(cudaMemsetAsync(ob->gate, 0, 4 * D, st));
    if (ob->raw) M16_CUDA_CHECK(cudaMemsetAsync(ob->raw, 0, 8 * 5 * D, st));
    return MOD16_OK;
  }
#define M16_EXACT(MODE, GRID)                                                              \
  do {                                                                                     \
    if (is_static) objective_exact_kernel<R, SEM_STATIC><<<GRID, OBJ_EX_THREADS, 0, st>>>(a, MODE); \
    else objective_exact_kernel<R, SEM_INSTANCE><<<GRID, OBJ_EX_THREADS, 0, st>>>(a, MODE);        \
    M16_CUDA_CHECK(cudaGetLastError());                                                    \
  } while (0)
  if (p.fast) {
    if constexpr (sizeof(R) == 4) {
      obj_classify_kernel<<<(unsigned)p.nblk, OBJ_CLS_THREADS, 0, st>>>(a);
      M16_CUDA_CHECK(cudaGetLastError());
      obj_scan_kernel<<<1, 1024, 0, st>>>(a.blk_fast, a.blk_exact, p.nblk, a.counts);
      M16_CUDA_CHECK(cudaGetLastError());
      obj_pack_kernel<<<(unsigned)p.nblk, OBJ_CLS_THREADS, 0, st>>>(a);
      M16_CUDA_CHECK(cudaGetLastError());
      const dim3 grid((unsigned)((D + OBJ_DRAWS_PER_BLOCK - 1) / OBJ_DRAWS_PER_BLOCK),
                      (unsigned)p.nf_chunks);
      if (fast_mode == 0) objective_fast_kernel<0><<<grid, OBJ_THREADS, 0, st>>>(a);
      else if (fast_mode == 1) objective_fast_kernel<1><<<grid, OBJ_THREADS, 0, st>>>(a);
      else objective_fast_kernel<2><<<grid, OBJ_THREADS, 0, st>>>(a);
      M16_CUDA_CHECK(cudaGetLastError());
    }
    M16_EXACT(1, dim3(1, draw_rows));
  } else {
    M16_EXACT(0, dim3((unsigned)p.ne_chunks, draw_rows));
  }
  objective_reduce_kernel<<<(unsigned)((D + 255) / 256), 256, 0, st>>>(
      ar, is_static ? 1 : 0, p.fast ? 1 : 0, fast_mode);
  M16_CUDA_CHECK(cudaGetLastError());
  if (p.fast) M16_EXACT(2, dim3(1, (unsigned)(D < 1024 ? D : 1024)));
#undef M16_EXACT
  return MOD16_OK;
}

}  // namespace mod16

extern "C" {
int64_t mod16_objective_workspace_bytes(int64_t T, int64_t N, int64_t D, int is_f32) {
  if (T < 0 || N < 0 || D < 0) return -1;
  // (the float32 entry may take either route: instance semantics and tiny batches go
  // through the exact kernel)
  const int64_t exact = mod16::make_plan(T, N, D, false).total;
  const int64_t fast = is_f32 ? mod16::make_plan(T, N, D, true).total : 0;
  return exact > fast ? exact : fast;
}
int mod16_objective_batch_f32(const mod16_drivers_t* drivers, const mod16_objective_t* obj,
                              int64_t T, int64_t N, int64_t D, uint32_t flags, void* stream) {
  return mod16::run_objective<float>(drivers, obj, T, N, D, flags, (cudaStream_t)stream);
}
int mod16_objective_batch_f64(const mod16_drivers_t* drivers, const mod16_objective_t* obj,
                              int64_t T, int64_t N, int64_t D, uint32_t flags, void* stream) {
  return mod16::run_objective<double>(drivers, obj, T, N, D, flags, (cudaStream_t)stream);
}
}
