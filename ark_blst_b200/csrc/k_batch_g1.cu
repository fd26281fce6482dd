// Batched-affine pairing rounds, G1 instantiation (batch_affine.cuh).
#include "batch_affine.cuh"
#include "launch.h"

namespace b200msm {
void launch_ba_round_g1(int first, const uint32_t *src, const uint32_t *vals, const uint32_t *endo_x, uint32_t n_pts,
                        const uint32_t *total_ptr, int round, const BaPlan &bp, uint32_t *prefix, uint32_t *T, uint32_t *prefix2,
                        uint32_t *U, uint32_t *out, cudaStream_t st, int part, int split_align_log, int img_full) {
    for (int k = 0; k < 5; k++) count_launch();
    BaSrc s{src, vals, endo_x, n_pts, img_full};
    const int shift = round + 1;
    const unsigned blocks = blocks_for(bp.NT, 128);
    if (first) k_ba_fwd<fp, true><<<blocks, 128, 0, st>>>(s, total_ptr, shift, part, split_align_log, bp.NT, bp.K, prefix, T);
    else k_ba_fwd<fp, false><<<blocks, 128, 0, st>>>(s, total_ptr, shift, part, split_align_log, bp.NT, bp.K, prefix, T);
    k_ba_prod_fwd<fp><<<blocks_for(bp.NU, 128), 128, 0, st>>>(T, bp.NT, bp.NU, bp.K2, prefix2, U);
    k_ba_invert<fp><<<blocks_for(bp.NU, 64), 64, 0, st>>>(U, bp.NU);
    k_ba_prod_bwd<fp><<<blocks_for(bp.NU, 128), 128, 0, st>>>(T, bp.NT, bp.NU, bp.K2, prefix2, U);
    if (first) k_ba_bwd<fp, true><<<blocks, 128, 0, st>>>(s, total_ptr, shift, part, split_align_log, bp.NT, bp.K, prefix, T, out);
    else k_ba_bwd<fp, false><<<blocks, 128, 0, st>>>(s, total_ptr, shift, part, split_align_log, bp.NT, bp.K, prefix, T, out);
}
void launch_accumulate_direct_g1(const uint32_t *pts, const uint32_t *start, const uint32_t *order, uint32_t nb, uint32_t heavy_thr,
                                 int shift, int into, uint32_t *buckets, cudaStream_t st) {
    count_launch();
    k_accumulate_direct<fp><<<blocks_for(nb, 128), 128, 0, st>>>(pts, start, order, nb, heavy_thr, shift, into, buckets);
}
}  // namespace b200msm

namespace b200msm {
static __global__ void k_dbg_inv_sg(int is_fp2, const uint32_t *in, uint32_t *out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (is_fp2) {
        fp2 a;
        f_load(a, in + i * 24);
        f_inv_sg(a, a);
        f_store(out + i * 24, a);
    } else {
        fp a;
        f_load(a, in + i * 12);
        f_inv_sg(a, a);
        f_store(out + i * 12, a);
    }
}
void launch_dbg_inv_sg(int is_fp2, const uint32_t *in, uint32_t *out, size_t n, cudaStream_t st) {
    count_launch();
    k_dbg_inv_sg<<<blocks_for(n, 64), 64, 0, st>>>(is_fp2, in, out, n);
}
}  // namespace b200msm
