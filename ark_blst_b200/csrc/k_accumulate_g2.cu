// Bucket accumulation, G2 instantiation.
#include "accumulate.cuh"
#include "launch.h"

namespace b200msm {
void launch_accumulate_g2(const uint32_t *bases, const uint32_t *vals, const uint32_t *start, const uint32_t *order,
                          uint32_t nb, uint32_t heavy_thr, const uint32_t *endo_x, uint32_t n_pts, int into, uint32_t *buckets,
                          cudaStream_t st, int img_full) {
    count_launch();
    k_accumulate<fp2><<<blocks_for(nb, 128), 128, 0, st>>>(bases, vals, start, order, nb, heavy_thr, endo_x, n_pts, img_full, into, buckets);
}
void launch_endo_table_g2(const uint32_t *bases, size_t n, uint32_t *endo_x, cudaStream_t st) {
    count_launch();
    k_endo_table<fp2><<<blocks_for(n, 256), 256, 0, st>>>(bases, n, endo_x);
}
void launch_psi_tables_g2(const uint32_t *bases, size_t n, uint32_t *img, cudaStream_t st) {
    count_launch();
    k_psi_tables<<<blocks_for(n, 128), 128, 0, st>>>(bases, n, img);
}
void launch_table_shift_g2(const uint32_t *prev, size_t n, int c, uint32_t *jac_out, cudaStream_t st) {
    count_launch();
    k_table_shift<fp2><<<blocks_for(n, 128), 128, 0, st>>>(prev, n, c, jac_out);
}
}  // namespace b200msm
