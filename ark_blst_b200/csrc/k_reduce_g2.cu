// Bucket reduction / window combination / partial sums, G2 instantiation.
#include "launch.h"
#include "reduce.cuh"

namespace b200msm {
void launch_wsum_level_g2(const uint32_t *X, const uint32_t *Cin, uint32_t len, uint32_t m, int log2M, uint32_t nwin,
                          uint32_t *Rout, uint32_t *Cout, cudaStream_t st) {
    count_launch();
    uint32_t nseg = len / m;
    // (G2 stays in quad form at every size: three XYZZ values over Fp2 are 288 registers — the thread form spills 1.5 KB)
    k_wsum_level<fp2><<<blocks_for((size_t)nseg * nwin * 4, 128), 128, 0, st>>>(X, Cin, len, m, log2M, nwin, Rout, Cout);
}
void launch_tree_level_g2(const uint32_t *Sin, size_t sin_stride, const uint32_t *Vin, const uint32_t *Cin, size_t cin_stride,
                          uint32_t *Sout, uint32_t *Vout, uint32_t *Cout, size_t out_stride, uint32_t S, int j, uint32_t nwin,
                          cudaStream_t st) {
    count_launch();
    size_t quads = (size_t)(S >> (j + 1)) * (j + 1 + (Cin ? 1 : 0)) * nwin;
    k_tree_level<fp2><<<blocks_for(quads * 4, 128), 128, 0, st>>>(Sin, sin_stride, Vin, Cin, cin_stride, Sout, Vout, Cout,
                                                                 out_stride, S, j, nwin);
}
}  // namespace b200msm
