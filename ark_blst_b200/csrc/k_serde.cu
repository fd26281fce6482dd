// Point (de)serialisation kernels, both groups.
#include "launch.h"
#include "serde.cuh"

namespace b200msm {
void launch_deserialize(int g2, const uint8_t *in, size_t n, int compressed, int validate, uint32_t *aff, uint8_t *status,
                        cudaStream_t st) {
    count_launch();
    if (g2) k_deserialize<fp2><<<blocks_for(n, 64), 64, 0, st>>>(in, n, compressed, validate, aff, status);
    else k_deserialize<fp><<<blocks_for(n, 64), 64, 0, st>>>(in, n, compressed, validate, aff, status);
}
void launch_serialize(int g2, const uint32_t *aff, size_t n, int compressed, uint8_t *out, cudaStream_t st) {
    count_launch();
    if (g2) k_serialize<fp2><<<blocks_for(n, 64), 64, 0, st>>>(aff, n, compressed, out);
    else k_serialize<fp><<<blocks_for(n, 64), 64, 0, st>>>(aff, n, compressed, out);
}
}  // namespace b200msm
