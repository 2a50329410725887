// placeholder until the vectorised f32 kernel lands
#include "mod16_common.cuh"
namespace mod16 {
bool fast_f32_applicable(const ForwardArgs<float>&, uint32_t) { return false; }
int launch_fast_f32(ForwardArgs<float>&, uint32_t, cudaStream_t) { return MOD16_ERR_FLAGS; }
}  // namespace mod16
