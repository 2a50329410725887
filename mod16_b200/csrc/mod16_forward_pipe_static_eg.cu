// forward_pipe_kernel, static semantics, part 2: the float32 layouts of part 0 instantiated with
// the guard at the zeros of the Tmin / VPD ramps (EG) -- what a launch with the per-element
// gate (MOD16_GATE_ELEMENT) runs on.
#include "mod16_forward_pipe.cuh"

namespace mod16 {
int launch_pipe_static_eg(const ForwardArgs<float>& a, const PipeSrc& src, int out_mode,
                          int extra, int fmt, dim3 grid, cudaStream_t st) {
  return launch_pipe_sem<SEM_STATIC, 2>(a, src, out_mode, extra, fmt, grid, st);
}
}  // namespace mod16
