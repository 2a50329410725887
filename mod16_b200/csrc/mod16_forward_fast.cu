// K1, float32 fast path: host-side dispatch and the small kernels around the headline one.
//
// forward_pipe_kernel (mod16_forward_pipe.cuh, instantiated in mod16_forward_pipe_{static,
//   instance}.cu) -- the headline kernel: dense (T, N) structure-of-arrays drivers streamed
//   through a shared-memory ring by bulk asynchronous copies, packed-FP32 math.
//
// forward_strided_fast_kernel -- same math, one pixel-timestep at a time, for layouts the
//   bulk copies cannot express (broadcast strides, N % 4 != 0, unaligned bases, r_corr_list).
//
// redo_queue_kernel / redo_sweep_kernel -- the double-precision evaluation of the elements
//   the fast math flagged.
#include "mod16_forward_pipe.cuh"

namespace mod16 {

#define M16_DECL(name)                                                                  \
  int name(const ForwardArgs<float>& a, const PipeSrc& src, int out_mode, int extra, int fmt, \
           dim3 grid, cudaStream_t st);
M16_DECL(launch_pipe_static) M16_DECL(launch_pipe_static_b)
M16_DECL(launch_pipe_instance) M16_DECL(launch_pipe_instance_b)
M16_DECL(launch_pipe_static_eg)
#undef M16_DECL

template <int SEM, int OUT>
__global__ void __launch_bounds__(256, 4)
forward_strided_fast_kernel(const __grid_constant__ ForwardArgs<float> a) {
  if (a.skip_if_flag && *((volatile int*)a.gate_flag) != 0) return;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t t0 = (int64_t)blockIdx.y * a.t_chunk;
  int64_t t1 = t0 + a.t_chunk;
  if (t1 > a.T) t1 = a.T;
  bool any_pos = false;
  if (n < a.N) {
    PftConst pc;
    {
      double col[MOD16_N_PARAMS];
      load_column(a, n, col);
      pc = fast::make_pft_const(col, SEM);
    }
    const bool pix_static = a.drv[MOD16_D_PRESSURE].ts == 0 &&
                            a.drv[MOD16_D_TEMP_ANNUAL].ts == 0;
    PixConst px = fast::make_pix_const(a.drv[MOD16_D_PRESSURE].at(t0, n),
                                       a.drv[MOD16_D_TEMP_ANNUAL].at(t0, n), pc, SEM);
    const bool has_rc = a.has_rc != 0;
    for (int64_t t = t0; t < t1; ++t) {
      const StepIn in = load_step(a, t, n);
      if (!pix_static) px = fast::make_pix_const(in.p, in.t_ann, pc, SEM);
      fast::StepOutV<float> r;
      bool g_pos;
      const fast::StepInV<float> v = fast::to_v(in);
      const bool need = has_rc ? fast::step_fast<float, SEM, true>(v, pc, px, a.gate, r, g_pos)
                               : fast::step_fast<float, SEM, false>(v, pc, px, a.gate, r, g_pos);
      if (need) {
        any_pos |= rare_one<SEM, OUT>(a, t, n, false) > 0;
      } else {
        any_pos |= g_pos;
        const int64_t o = t * a.N + n;
        if (OUT == 0) {
          a.out[0][o] = r.tot;
        } else if (OUT == 1) {
          a.out[0][o] = r.tot_day;
          a.out[1][o] = r.tot_night;
        } else {
          a.out[0][o] = r.day.canopy;   a.out[1][o] = r.day.soil;   a.out[2][o] = r.day.trans;
          a.out[3][o] = r.night.canopy; a.out[4][o] = r.night.soil; a.out[5][o] = r.night.trans;
        }
      }
    }
  }
  if (a.track_gate) {
    if (__any_sync(0xffffffffu, any_pos) && (threadIdx.x & 31) == 0) *a.gate_flag = 1;
  }
}

// Second, tiny launch of the float32 path: evaluates the queued elements in double.
template <int SEM, int OUT>
__global__ void __launch_bounds__(128)
redo_queue_kernel(const __grid_constant__ ForwardArgs<float> a) {
  if (a.skip_if_flag && *((volatile int*)a.gate_flag) != 0) return;
  unsigned long long cnt = a.redo_ws[0];
  if (cnt == 0) return;
  if (cnt > (unsigned long long)a.redo_cap) cnt = (unsigned long long)a.redo_cap;
  bool any_pos = false;
  for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < cnt;
       e += (unsigned long long)gridDim.x * blockDim.x) {
    const long long idx = (long long)a.redo_ws[1 + e];
    any_pos |= fix_exact_one<SEM, OUT>(a, idx / a.N, idx % a.N);
  }
  if (a.track_gate) {
    if (__any_sync(0xffffffffu, any_pos) && (threadIdx.x & 31) == 0) *a.gate_flag = 1;
  }
}

// Third launch of the float32 path; returns at once unless the queue overflowed.  Scans the
// first output for the marker the main kernel left and evaluates those elements exactly.
// A block collects the marked indices of a chunk into shared memory first and then works
// through them one element per thread: the double code runs on full warps however sparse the
// marks are.
constexpr int SWEEP_THREADS = 128;
constexpr int SWEEP_CHUNK = SWEEP_THREADS * 16;

template <int SEM, int OUT>
__global__ void __launch_bounds__(SWEEP_THREADS)
redo_sweep_kernel(const __grid_constant__ ForwardArgs<float> a) {
  __shared__ long long found[SWEEP_CHUNK];
  __shared__ int n_found;
  if (a.skip_if_flag && *((volatile int*)a.gate_flag) != 0) return;
  if (a.redo_ws[0] <= (unsigned long long)a.redo_cap) return;
  bool any_pos = false;
  const long long total = (long long)a.T * a.N;
  for (long long base = (long long)blockIdx.x * SWEEP_CHUNK; base < total;
       base += (long long)gridDim.x * SWEEP_CHUNK) {
    if (threadIdx.x == 0) n_found = 0;
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < SWEEP_CHUNK / SWEEP_THREADS; ++k) {
      const long long e = base + k * SWEEP_THREADS + threadIdx.x;
      if (e < total && __float_as_int(__ldg(a.out[0] + e)) == REDO_MARK)
        found[atomicAdd(&n_found, 1)] = e;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_found; k += SWEEP_THREADS) {
      const long long e = found[k];
      any_pos |= fix_exact_one<SEM, OUT>(a, e / a.N, e % a.N);
    }
    __syncthreads();
  }
  if (a.track_gate) {
    if (__any_sync(0xffffffffu, any_pos) && (threadIdx.x & 31) == 0) *a.gate_flag = 1;
  }
}

// MOD16_OUT_ANNUAL, fix-up pass of the call-exact gate: forget the open pass's sums
__global__ void annual_clear_kernel(unsigned long long* acc, int64_t n, const int* gate_flag) {
  if (*((volatile const int*)gate_flag) != 0) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    acc[i] = 0ULL;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// a (T, N) array the bulk copies can stream: unit pixel stride, every row 16-byte aligned
// (rows need not be adjacent: a pixel range of a wider array qualifies)
static bool dense_tn(const Arr<float>& d, int64_t T, int64_t N, int esize = 4) {
  return d.ptr && d.ns == 1 && aligned16(d.ptr) &&
         (T == 1 || (d.ts >= N && (d.ts * esize) % 16 == 0));
}

// The pipelined kernel needs: N % 4 == 0 and 16-byte aligned bases (every bulk copy is a
// multiple of 16 bytes at a 16-byte aligned address); the eleven per-step drivers (T, N)
// with unit pixel stride and 16-byte aligned rows; sw_rad_night a scalar or dense; pressure and temp_annual both per pixel / scalar,
// or both dense (T, N) (the calibration's annual means, mod16/calibration.py:556-560);
// parameters from the table (uniform or gathered by land cover) or per-pixel arrays (the
// reference callers' `params[pft_map].reshape((1, N))`); no r_corr_list.
// Returns the EXTRA variant (mod16_forward_pipe.cuh), or -1.
static int pipe_variant(const ForwardArgs<float>& a, uint32_t flags) {
  if (a.N % 4 != 0 || a.has_rc) return -1;
  // the per-element gate of the static path needs the guard at the zeros of the Tmin / VPD
  // ramps (fast::ramp_zero_need): the pipelined kernel has it in separate instantiations, which
  // exist for the plain float32 layouts and the three classic output modes only
  if ((flags & MOD16_SEM_MASK) == MOD16_SEM_STATIC && a.gate == GATE_ELEMENT &&
      (a.raw_fmt != 0 || (flags & MOD16_OUT_MASK) > MOD16_OUT_COMPONENTS))
    return -1;
  if (a.raw_fmt != 0) {  // storage formats: the eleven per-step arrays in the format's types
    static const int order[] = {
        MOD16_D_LW_NET_DAY, MOD16_D_LW_NET_NIGHT, MOD16_D_SW_RAD_DAY, MOD16_D_SW_ALBEDO,
        MOD16_D_TEMP_DAY,   MOD16_D_TEMP_NIGHT,   MOD16_D_TMIN,       MOD16_D_VPD_DAY,
        MOD16_D_VPD_NIGHT,  MOD16_D_FPAR,         MOD16_D_LAI};
    if (a.raw_fmt >= FMT_MODIS && a.N % 16 != 0) return -1;  // uint8 rows: 16-byte multiples
    for (int k = 0; k < 11; ++k) {
      const int es = pipe_esize(a.raw_fmt, k);
      const int want = es == 4 ? MOD16_T_F32 : (es == 2 ? MOD16_T_I16 : MOD16_T_U8);
      if (a.raw[order[k]].dtype != want || !dense_tn(a.drv[order[k]], a.T, a.N, es)) return -1;
    }
    if (a.drv[MOD16_D_SW_RAD_NIGHT].ptr) return -1;
    for (int k : {MOD16_D_PRESSURE, MOD16_D_TEMP_ANNUAL}) {
      const Arr<float>& d = a.drv[k];
      if (a.raw[k].dtype != MOD16_T_F32) return -1;
      if (d.ptr && !(d.ts == 0 && (d.ns == 1 || d.ns == 0))) return -1;
    }
    const uint32_t om = flags & MOD16_OUT_MASK;
    if (om != MOD16_OUT_SUM && om != MOD16_OUT_DAY_NIGHT) return -1;
    if (a.prep & MOD16_PREP_VPD_FROM_QV) {   // + the two surface-pressure arrays
      if (a.raw_fmt != FMT_F32_PREP || !dense_tn(a.ps[0], a.T, a.N) ||
          !dense_tn(a.ps[1], a.T, a.N))
        return -1;
      return 4;
    }
    return 0;
  }
  static const int per_step[] = {
      MOD16_D_LW_NET_DAY, MOD16_D_LW_NET_NIGHT, MOD16_D_SW_RAD_DAY, MOD16_D_SW_ALBEDO,
      MOD16_D_TEMP_DAY,   MOD16_D_TEMP_NIGHT,   MOD16_D_TMIN,       MOD16_D_VPD_DAY,
      MOD16_D_VPD_NIGHT,  MOD16_D_FPAR,         MOD16_D_LAI};
  for (int k : per_step)
    if (!dense_tn(a.drv[k], a.T, a.N)) return -1;
  int extra = 0;
  const Arr<float>& sn = a.drv[MOD16_D_SW_RAD_NIGHT];
  if (sn.ptr) {
    if (!dense_tn(sn, a.T, a.N)) return -1;
    extra |= 1;
  }
  const Arr<float>& pr = a.drv[MOD16_D_PRESSURE];
  const Arr<float>& ta = a.drv[MOD16_D_TEMP_ANNUAL];
  const bool pr_pix = !pr.ptr || (pr.ts == 0 && pr.ns == 1) || (pr.ts == 0 && pr.ns == 0);
  const bool ta_pix = !ta.ptr || (ta.ts == 0 && ta.ns == 1) || (ta.ts == 0 && ta.ns == 0);
  if (!(pr_pix && ta_pix)) {
    // time-varying: both streamed (T > 1 here: a T == 1 array counts as per pixel)
    if (a.T == 1 || !dense_tn(pr, a.T, a.N) || !dense_tn(ta, a.T, a.N)) return -1;
    extra |= 2;
  }
  const uint32_t om = flags & MOD16_OUT_MASK;
  const int out_mode = om == MOD16_OUT_SUM ? 0 : (om == MOD16_OUT_DAY_NIGHT ? 1
                       : (om == MOD16_OUT_ANNUAL ? 3 : (om == MOD16_OUT_POTENTIAL ? 4
                       : (om == MOD16_OUT_COMPOSITE ? 5 : (om == MOD16_OUT_SUM_I16 ? 6 : 2)))));
  if (!pipe_has_variant(out_mode, extra)) return -1;
  return extra;
}

// Grid of the pipelined kernel: one block per 480-pixel tile and time chunk; the chunk is
// as long as possible (the per-block set-up is amortised over it) while the grid still
// holds several waves of 148 SMs x 1 resident block.
static int launch_pipe(ForwardArgs<float>& a, uint32_t flags, int extra, cudaStream_t st) {
  static const int order[] = {
      MOD16_D_LW_NET_DAY, MOD16_D_LW_NET_NIGHT, MOD16_D_SW_RAD_DAY, MOD16_D_SW_ALBEDO,
      MOD16_D_TEMP_DAY,   MOD16_D_TEMP_NIGHT,   MOD16_D_TMIN,       MOD16_D_VPD_DAY,
      MOD16_D_VPD_NIGHT,  MOD16_D_FPAR,         MOD16_D_LAI};
  PipeSrc src;
  for (int k = 0; k < PIPE_MAX_ARR; ++k) { src.ptr[k] = nullptr; src.ts[k] = 0; }
  int n_arr = 0;
  const int fmt = a.raw_fmt;
  auto add = [&](int d) {   // (strides of the copy engine are in bytes)
    src.ptr[n_arr] = (const char*)a.drv[d].ptr;
    src.ts[n_arr] = a.drv[d].ts * pipe_esize(fmt, n_arr);
    ++n_arr;
  };
  for (int k = 0; k < 11; ++k) add(order[k]);
  if (extra & 1) add(MOD16_D_SW_RAD_NIGHT);
  if (extra & 2) {
    add(MOD16_D_TEMP_ANNUAL);
    add(MOD16_D_PRESSURE);
  }
  if (extra & 4) {
    for (int h = 0; h < 2; ++h) {
      src.ptr[n_arr] = (const char*)a.ps[h].ptr;
      src.ts[n_arr] = a.ps[h].ts * 4;
      ++n_arr;
    }
  }
  const int64_t bx = (a.N + PIPE_TILE - 1) / PIPE_TILE;
  if (bx > 2147483647LL) return MOD16_ERR_SHAPE;
  int64_t want_rows = (148 * 6 + bx - 1) / bx;
  if (want_rows < 1) want_rows = 1;
  int64_t chunk = (a.T + want_rows - 1) / want_rows;
  if (chunk < 8) chunk = a.T < 8 ? a.T : 8;
  chunk += chunk & 1;  // even chunks: only the last one may end on a single step
  int64_t by = (a.T + chunk - 1) / chunk;
  while (by > 65535) { chunk *= 2; by = (a.T + chunk - 1) / chunk; }
  a.t_chunk = (int)chunk;
  dim3 grid((unsigned)bx, (unsigned)by);
  const uint32_t om = flags & MOD16_OUT_MASK;
  const int out_mode = om == MOD16_OUT_SUM ? 0 : (om == MOD16_OUT_DAY_NIGHT ? 1
                       : (om == MOD16_OUT_ANNUAL ? 3 : (om == MOD16_OUT_POTENTIAL ? 4
                       : (om == MOD16_OUT_COMPOSITE ? 5 : (om == MOD16_OUT_SUM_I16 ? 6 : 2)))));
  const bool is_static = (flags & MOD16_SEM_MASK) == MOD16_SEM_STATIC;
  if (a.redo_cap > 0) M16_CUDA_CHECK(cudaMemsetAsync(a.redo_ws, 0, 8, st));
  if (out_mode == 3 && a.skip_if_flag) {
    // the fix-up pass of the call-exact gate recomputes everything: the sums of the open
    // pass go (only if the fix-up will run at all, i.e. the flag stayed 0)
    annual_clear_kernel<<<148 * 8, 256, 0, st>>>((unsigned long long*)a.out[0],
                                                 (int64_t)a.n_groups * a.N, a.gate_flag);
    M16_CUDA_CHECK(cudaGetLastError());
  }
  const bool part_b = fmt != 0 || out_mode >= 3;
  // (the per-element gate of the static path: the instantiations with the ramp-zero guard)
  const bool eg = is_static && a.gate == GATE_ELEMENT;
  const int rc = is_static
      ? (eg ? launch_pipe_static_eg : (part_b ? launch_pipe_static_b : launch_pipe_static))(
            a, src, out_mode, extra, fmt, grid, st)
      : (part_b ? launch_pipe_instance_b : launch_pipe_instance)(a, src, out_mode, extra, fmt, grid, st);
  if (rc != MOD16_OK) return rc;
  if (a.redo_cap > 0) {
#define M16_REDO(S, O)                                                                  \
  do {                                                                                  \
    redo_queue_kernel<S, O><<<148 * 4, 128, 0, st>>>(a);                                \
    M16_CUDA_CHECK(cudaGetLastError());                                                 \
    redo_sweep_kernel<S, O><<<148 * 4, SWEEP_THREADS, 0, st>>>(a);                      \
    M16_CUDA_CHECK(cudaGetLastError());                                                 \
  } while (0)
    if (is_static && out_mode == 3) {   // (no marker sweep: nothing per step to mark)
      redo_queue_kernel<SEM_STATIC, 3><<<148 * 4, 128, 0, st>>>(a);
      M16_CUDA_CHECK(cudaGetLastError());
    } else if (out_mode == 6) {
      if (is_static) redo_queue_kernel<SEM_STATIC, 6><<<148 * 4, 128, 0, st>>>(a);
      else redo_queue_kernel<SEM_INSTANCE, 6><<<148 * 4, 128, 0, st>>>(a);
      M16_CUDA_CHECK(cudaGetLastError());
    } else if (is_static) {
      if (out_mode == 0) M16_REDO(SEM_STATIC, 0);
      else if (out_mode == 1) M16_REDO(SEM_STATIC, 1);
      else if (out_mode == 4) M16_REDO(SEM_STATIC, 4);
      else if (out_mode == 5) M16_REDO(SEM_STATIC, 5);
      else M16_REDO(SEM_STATIC, 2);
    } else {
      if (out_mode == 0) M16_REDO(SEM_INSTANCE, 0);
      else if (out_mode == 1) M16_REDO(SEM_INSTANCE, 1);
      else if (out_mode == 4) M16_REDO(SEM_INSTANCE, 4);
      else if (out_mode == 5) M16_REDO(SEM_INSTANCE, 5);
      else M16_REDO(SEM_INSTANCE, 2);
    }
#undef M16_REDO
  }
  return MOD16_OK;
}

template <int SEM>
static int launch_strided_sem(const ForwardArgs<float>& a, uint32_t om, dim3 grid, dim3 block,
                              cudaStream_t st) {
  if (om == MOD16_OUT_SUM) forward_strided_fast_kernel<SEM, 0><<<grid, block, 0, st>>>(a);
  else if (om == MOD16_OUT_DAY_NIGHT) forward_strided_fast_kernel<SEM, 1><<<grid, block, 0, st>>>(a);
  else forward_strided_fast_kernel<SEM, 2><<<grid, block, 0, st>>>(a);
  M16_CUDA_CHECK(cudaGetLastError());
  return MOD16_OK;
}

int launch_fast_f32(ForwardArgs<float>& a, uint32_t flags, cudaStream_t st) {
  if (a.T == 0 || a.N == 0) return MOD16_OK;
  if (!(flags & MOD16_FORCE_GENERIC)) {
    const int extra = pipe_variant(a, flags);
    if (extra >= 0) {
      g_last_forward_kernel = MOD16_KERNEL_PIPE | (extra << 8) | (a.raw_fmt << 12);
      return launch_pipe(a, flags, extra, st);
    }
  }
  // (no strided kernel for storage formats, nor for the annual sums)
  if (a.raw_fmt != 0 || (flags & MOD16_OUT_MASK) == MOD16_OUT_ANNUAL ||
      (flags & MOD16_OUT_MASK) == MOD16_OUT_POTENTIAL ||
      (flags & MOD16_OUT_MASK) == MOD16_OUT_COMPOSITE ||
      (flags & MOD16_OUT_MASK) == MOD16_OUT_SUM_I16)
    return MOD16_ERR_LAYOUT;
  g_last_forward_kernel = MOD16_KERNEL_STRIDED;
  const int threads = 256;
  const int64_t bx = (a.N + threads - 1) / threads;
  if (bx > 2147483647LL) return MOD16_ERR_SHAPE;
  // Block rows: split time so that the grid is several waves of 148 SMs x resident blocks
  // while each thread amortises its per-pixel setup over a run of time steps.
  int64_t want_rows = (148 * 2 * 8 + bx - 1) / bx;
  if (want_rows < 1) want_rows = 1;
  int64_t chunk = (a.T + want_rows - 1) / want_rows;
  if (chunk < 1) chunk = 1;
  if (chunk > 32) chunk = 32;
  int64_t by = (a.T + chunk - 1) / chunk;
  while (by > 65535) { chunk *= 2; by = (a.T + chunk - 1) / chunk; }
  a.t_chunk = (int)chunk;
  dim3 grid((unsigned)bx, (unsigned)by), block(threads);
  const uint32_t om = flags & MOD16_OUT_MASK;
  if ((flags & MOD16_SEM_MASK) == MOD16_SEM_STATIC)
    return launch_strided_sem<SEM_STATIC>(a, om, grid, block, st);
  return launch_strided_sem<SEM_INSTANCE>(a, om, grid, block, st);
}

}  // namespace mod16
