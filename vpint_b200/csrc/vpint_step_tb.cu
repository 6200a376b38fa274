// Temporally blocked 2-D kernel (placeholder until the TMA kernel lands).
#include "vpint_internal.cuh"

namespace vpint {

void tb_tile_shape(int k_block, int *th, int *tw) { *th = 64 - 2 * k_block; *tw = 64 - 2 * k_block; }
bool tb_supported(const Geom &, int) { return false; }
cudaError_t launch_step2d_tb(const StepArgs &, int, int, int64_t, cudaStream_t, int64_t *) {
    return cudaErrorNotSupported;
}

}  // namespace vpint
