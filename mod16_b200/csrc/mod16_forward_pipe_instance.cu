// forward_pipe_kernel, instance semantics (`evapotranspiration()`): instantiations, part 0
// (float32 layouts, sum / day-night / components).
#include "mod16_forward_pipe.cuh"

namespace mod16 {
int launch_pipe_instance(const ForwardArgs<float>& a, const PipeSrc& src, int out_mode, int extra,
                               int fmt, dim3 grid, cudaStream_t st) {
  return launch_pipe_sem<SEM_INSTANCE, 0>(a, src, out_mode, extra, fmt, grid, st);
}
}  // namespace mod16
