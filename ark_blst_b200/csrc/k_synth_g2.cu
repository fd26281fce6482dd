// Synthetic bases and point-operation unit hooks, G2 instantiation (bench / tests only).
#include "launch.h"
#include "synth.cuh"

namespace b200msm {
void launch_synth_bases_g2(uint64_t seed, size_t n, uint32_t *out, cudaStream_t st) {
    count_launch();
    k_synth_bases<fp2><<<blocks_for(n, 128), 128, 0, st>>>(seed, n, out);
}
void launch_dbg_point_op_g2(int op, const uint32_t *acc, const uint32_t *q, uint32_t *out, size_t n) {
    count_launch();
    if (op >= 3) k_dbg_point_op_quad<fp2><<<blocks_for(n * 4, 128), 128>>>(op, acc, q, out, n);
    else k_dbg_point_op<fp2><<<blocks_for(n, 64), 64>>>(op, acc, q, out, n);
}
}  // namespace b200msm
