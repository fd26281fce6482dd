// Synthetic scalars, the integer-pipe peak microbenchmark and the field-operation unit hooks.
#include "launch.h"
#include "synth.cuh"

namespace b200msm {
void launch_synth_scalars(uint64_t seed, size_t n, int mont, uint32_t *out, cudaStream_t st) {
    count_launch();
    k_synth_scalars<<<blocks_for(n, 256), 256, 0, st>>>(seed, n, mont, out);
}
void launch_imad_peak(int mode, int blocks, int threads, uint32_t *buf, int iters, cudaStream_t st) {
    count_launch();
    if (mode == 0) k_imad_peak<0><<<blocks, threads, 0, st>>>(buf, 12345, iters);
    else k_imad_peak<1><<<blocks, threads, 0, st>>>(buf, 12345, iters);
}
void launch_dbg_field_op(int is_fp2, int op, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n) {
    count_launch();
    if (is_fp2) k_dbg_field_op<fp2><<<blocks_for(n, 64), 64>>>(op, a, b, out, n);
    else k_dbg_field_op<fp><<<blocks_for(n, 64), 64>>>(op, a, b, out, n);
}
}  // namespace b200msm
