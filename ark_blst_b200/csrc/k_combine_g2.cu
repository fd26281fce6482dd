// Window combination and partial sums, G2 instantiation (split from k_reduce_g2.cu so the two compile in parallel).
#include "launch.h"
#include "reduce.cuh"

namespace b200msm {
void launch_combine_g2(const uint32_t *Sroot, const uint32_t *V, const uint32_t *Croot, size_t stride, int logS, int log2M,
                       int nwin, int c, int split_top, uint32_t *wsum, uint32_t *out, cudaStream_t st) {
    count_launch();
    k_combine<fp2><<<1, 128, 0, st>>>(Sroot, V, Croot, stride, logS, log2M, nwin, c, split_top, wsum, out);
}
void launch_sum_partials_g2(const uint32_t *partials, int count, uint32_t *out, cudaStream_t st) {
    count_launch();
    k_sum_partials<fp2><<<1, 32, 0, st>>>(partials, count, out);
}
}  // namespace b200msm
