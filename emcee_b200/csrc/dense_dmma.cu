// Stretch move + dense Gaussian on FP64 tensor cores -- placeholder until the
// DMMA kernel lands (the generic kernel serves every shape meanwhile).
#include "engine.cuh"

namespace eb {
bool dense_dmma_supported(int) { return false; }
size_t dense_dmma_factor_doubles(int) { return 0; }
void dense_dmma_pack_factor(const double*, int, double*) {}
cudaError_t launch_half_step_dense_dmma(const HalfStepArgs&, int, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace eb
