// Batch normalisation to affine, both groups.
#include "launch.h"
#include "normalize.cuh"

namespace b200msm {
void launch_normalize_batch(int g2, const uint32_t *proj, size_t n, uint32_t *aff, int sm_count, cudaStream_t st) {
    count_launch();
    // ~2 warps per scheduler keeps the fma pipe busy while the per-thread batches stay long
    size_t want = (size_t)sm_count * 4 * 2 * 32;
    size_t threads = n < want ? n : want;
    unsigned blocks = blocks_for(threads, 128);
    if (g2) k_normalize_batch<fp2><<<blocks, 128, 0, st>>>(proj, n, aff);
    else k_normalize_batch<fp><<<blocks, 128, 0, st>>>(proj, n, aff);
}
}  // namespace b200msm
