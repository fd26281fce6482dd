// Bucket reduction / window combination / partial sums, G1 instantiation.
#include "launch.h"
#include "reduce.cuh"

namespace b200msm {
void launch_wsum_level_g1(const uint32_t *X, const uint32_t *Cin, uint32_t len, uint32_t m, int log2M, uint32_t nwin,
                          uint32_t *Rout, uint32_t *Cout, cudaStream_t st) {
    count_launch();
    uint32_t nseg = len / m;
    k_wsum_level<fp><<<blocks_for((size_t)nseg * nwin * 4, 128), 128, 0, st>>>(X, Cin, len, m, log2M, nwin, Rout, Cout);
}
void launch_combine_g1(const uint32_t *C, const uint32_t *R, int nwin, int c, uint32_t *out, cudaStream_t st) {
    count_launch();
    k_combine<fp><<<1, 32, 0, st>>>(C, R, nwin, c, out);
}
void launch_sum_partials_g1(const uint32_t *partials, int count, uint32_t *out, cudaStream_t st) {
    count_launch();
    k_sum_partials<fp><<<1, 32, 0, st>>>(partials, count, out);
}
}  // namespace b200msm
