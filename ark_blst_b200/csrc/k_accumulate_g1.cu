// Bucket accumulation, G1 instantiation.
#include "accumulate.cuh"
#include "launch.h"

namespace b200msm {
void launch_accumulate_g1(const uint32_t *bases, const uint32_t *vals, const uint32_t *start, const uint32_t *order,
                          uint32_t nb, uint32_t heavy_thr, const uint32_t *endo_x, uint32_t n_pts, int into, uint32_t *buckets,
                          cudaStream_t st, int img_full) {
    count_launch();
    k_accumulate<fp><<<blocks_for(nb, 128), 128, 0, st>>>(bases, vals, start, order, nb, heavy_thr, endo_x, n_pts, img_full, into, buckets);
}
void launch_heavy_g1(const uint32_t *bases, const uint32_t *vals, const uint32_t *start, const uint32_t *order,
                     uint32_t nb, uint32_t heavy_thr, const uint32_t *endo_x, uint32_t n_pts, void *hdr, void *hb, void *tasks,
                     uint32_t *partials, int into, uint32_t *buckets, int grid, cudaStream_t st, int shift, int img_full) {
    count_launch();
    count_launch();
    count_launch();
    k_plan_heavy<<<blocks_for(nb, 256), 256, 0, st>>>(start, order, nb, heavy_thr, HEAVY_CHUNK, shift, (HeavyHeader *)hdr,
                                                      (HeavyBucket *)hb, (HeavyTask *)tasks);
    k_heavy_tasks<fp, 128><<<grid, 128, 0, st>>>(bases, vals, (const HeavyHeader *)hdr, (const HeavyTask *)tasks, endo_x, n_pts, img_full, partials);
    k_heavy_final<fp, 128><<<grid, 128, 0, st>>>((const HeavyHeader *)hdr, (const HeavyBucket *)hb, partials, into, buckets);
}
void launch_endo_table_g1(const uint32_t *bases, size_t n, uint32_t *endo_x, cudaStream_t st) {
    count_launch();
    k_endo_table<fp><<<blocks_for(n, 256), 256, 0, st>>>(bases, n, endo_x);
}
void launch_table_shift_g1(const uint32_t *prev, size_t n, int c, uint32_t *jac_out, cudaStream_t st) {
    count_launch();
    k_table_shift<fp><<<blocks_for(n, 128), 128, 0, st>>>(prev, n, c, jac_out);
}
}  // namespace b200msm
