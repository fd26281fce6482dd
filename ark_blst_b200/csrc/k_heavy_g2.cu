// Heavy-bucket path, G2 instantiation (its own translation unit: with the Fp2 additions inlined it is one of the two
// slowest files of the build, and the TUs compile in parallel).
#include "accumulate.cuh"
#include "launch.h"

namespace b200msm {
void launch_heavy_g2(const uint32_t *bases, const uint32_t *vals, const uint32_t *start, const uint32_t *order,
                     uint32_t nb, uint32_t heavy_thr, const uint32_t *endo_x, uint32_t n_pts, void *hdr, void *hb, void *tasks,
                     uint32_t *partials, int into, uint32_t *buckets, int grid, cudaStream_t st, int shift, int img_full) {
    count_launch();
    count_launch();
    count_launch();
    k_plan_heavy<<<blocks_for(nb, 256), 256, 0, st>>>(start, order, nb, heavy_thr, HEAVY_CHUNK, shift, (HeavyHeader *)hdr,
                                                      (HeavyBucket *)hb, (HeavyTask *)tasks);
    k_heavy_tasks<fp2, 128><<<grid, 128, 0, st>>>(bases, vals, (const HeavyHeader *)hdr, (const HeavyTask *)tasks, endo_x, n_pts, img_full, partials);
    k_heavy_final<fp2, 128><<<grid, 128, 0, st>>>((const HeavyHeader *)hdr, (const HeavyBucket *)hb, partials, into, buckets);
}
}  // namespace b200msm
