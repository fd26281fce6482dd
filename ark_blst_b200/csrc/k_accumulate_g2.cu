// Bucket accumulation, G2 instantiation.
#include "accumulate.cuh"
#include "launch.h"

namespace b200msm {
void launch_accumulate_g2(const uint32_t *bases, const uint32_t *vals, const uint32_t *start, const uint32_t *order,
                          uint32_t nb, uint32_t *buckets, cudaStream_t st) {
    k_accumulate<fp2><<<blocks_for(nb, 128), 128, 0, st>>>(bases, vals, start, order, nb, buckets);
}
}  // namespace b200msm
