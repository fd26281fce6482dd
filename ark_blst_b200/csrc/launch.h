// Host-callable launchers, one translation unit per kernel family × group so the heavy field code
// compiles in parallel (and only once per change).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "plan.h"

namespace b200msm {

// k_prep.cu
void launch_digits_dbg(const uint32_t *scalars, size_t n, int mont, int c, int nwin, int *out, cudaStream_t st);
// hand-written grouping by bucket (counting sort on the bucket id, built from the scalars):
// glv: 1 = split every scalar in two 128-bit halves (entries per window double), 2 = the same with the top c bits taken
// unsigned and spread over the last two windows (c | 128, nwin = 128/c + 1; see k_hist); 3 / 4 = four 64-bit parts (G2,
// gls4.cuh) with a carry window / with the unsigned top digit (entries per window ×4); dig: entries·nwin u32, count: nb u32, start: nb+2 u32, tile_sums: nb/2048+2 u32, vals: ≥ n·nwin u32
// tbl_stride > 0 (fixed-base window table, `tbl_stride` points per window): all windows share one bucket set — nb = 2^(c−1),
// entries are grouped by the digit alone and carry the table index w·tbl_stride + i (must stay below 2^31)
// pad_log > 0 (batched-affine rounds follow): every bucket's segment is padded to a multiple of 2^pad_log entries and
// the padding slots of vals hold 0xffffffff (vals must then hold n·nwin + nb·(2^pad_log − 1) entries)
void launch_group_by_bucket(const uint32_t *scalars, size_t n, int mont, int glv, int c, int nwin, uint32_t nb, uint32_t *dig,
                            uint32_t *count, uint32_t *start, uint32_t *tile_sums, uint32_t *vals, cudaStream_t st,
                            size_t tbl_stride = 0, int pad_log = 0);
// bucket ids in decreasing-size order (counting sort on the clamped size); hist: 8192 u32 scratch
void launch_order_by_size(const uint32_t *start, uint32_t nb, uint32_t *hist, uint32_t *order, cudaStream_t st);
// k_accumulate_g{1,2}.cu
constexpr uint32_t HEAVY_CHUNK = 4096;  // entries per block task of a heavy bucket
// endo_x / n_pts: GLV: value indices ≥ n_pts name φ(P) = (β·x, y) of point index − n_pts, x read from
// the β·x table; pass nullptr / 0xffffffff when unused
void launch_accumulate_g1(const uint32_t *bases, const uint32_t *vals, const uint32_t *start, const uint32_t *order,
                          uint32_t nb, uint32_t heavy_thr, const uint32_t *endo_x, uint32_t n_pts, int into, uint32_t *buckets,
                          cudaStream_t st, int img_full = 0);
void launch_accumulate_g2(const uint32_t *bases, const uint32_t *vals, const uint32_t *start, const uint32_t *order,
                          uint32_t nb, uint32_t heavy_thr, const uint32_t *endo_x, uint32_t n_pts, int into, uint32_t *buckets,
                          cudaStream_t st, int img_full = 0);
void launch_endo_table_g1(const uint32_t *bases, size_t n, uint32_t *endo_x, cudaStream_t st);
void launch_endo_table_g2(const uint32_t *bases, size_t n, uint32_t *endo_x, cudaStream_t st);
// G2, four-part decomposition: img ← the image tables −ψ(Q), ψ²(Q), −ψ³(Q) (3·n full points); kernels then take img as
// `endo_x` with img_full = 1 (value indices n_pts + j·n_pts + i name image j+1 of point i)
void launch_psi_tables_g2(const uint32_t *bases, size_t n, uint32_t *img, cudaStream_t st);
// plan + block tasks + per-bucket fold for buckets above heavy_thr; hdr must be zeroed (8 bytes)
void launch_heavy_g1(const uint32_t *bases, const uint32_t *vals, const uint32_t *start, const uint32_t *order,
                     uint32_t nb, uint32_t heavy_thr, const uint32_t *endo_x, uint32_t n_pts, void *hdr, void *hb, void *tasks,
                     uint32_t *partials, int into, uint32_t *buckets, int grid, cudaStream_t st, int shift = 0, int img_full = 0);
void launch_heavy_g2(const uint32_t *bases, const uint32_t *vals, const uint32_t *start, const uint32_t *order,
                     uint32_t nb, uint32_t heavy_thr, const uint32_t *endo_x, uint32_t n_pts, void *hdr, void *hb, void *tasks,
                     uint32_t *partials, int into, uint32_t *buckets, int grid, cudaStream_t st, int shift = 0, int img_full = 0);
// (vals == nullptr with shift = R: direct mode over the array the batched-affine rounds left, see k_batch_g{1,2}.cu)

// k_batch_g{1,2}.cu — batched-affine pairing rounds (batch_affine.cuh).  One round halves the (padded) entry array:
// first = 1 reads entries through vals from the bases (sign / GLV image applied), else from `src` (the previous
// round's output).  total_ptr → the scan's grand total (slots of the entry array); round = 0-based; s_out_max bounds
// the output slots on the host.  Scratch: prefix ≥ NT·K elements, T and prefix2 ≥ NT, U ≥ NU (sizes from plan.h: ba_plan / ba_layout).
// split_align_log > 0: the launch covers part 0 / 1 of the slot range only (batch_affine.cuh: ba_range); the caller
// gives each part its own scratch and stream.
void launch_ba_round_g1(int first, const uint32_t *src, const uint32_t *vals, const uint32_t *endo_x, uint32_t n_pts,
                        const uint32_t *total_ptr, int round, const BaPlan &bp, uint32_t *prefix, uint32_t *T, uint32_t *prefix2,
                        uint32_t *U, uint32_t *out, cudaStream_t st, int part = 0, int split_align_log = 0, int img_full = 0);
void launch_ba_round_g2(int first, const uint32_t *src, const uint32_t *vals, const uint32_t *endo_x, uint32_t n_pts,
                        const uint32_t *total_ptr, int round, const BaPlan &bp, uint32_t *prefix, uint32_t *T, uint32_t *prefix2,
                        uint32_t *U, uint32_t *out, cudaStream_t st, int part = 0, int split_align_log = 0, int img_full = 0);
void launch_accumulate_direct_g1(const uint32_t *pts, const uint32_t *start, const uint32_t *order, uint32_t nb, uint32_t heavy_thr,
                                 int shift, int into, uint32_t *buckets, cudaStream_t st);
void launch_accumulate_direct_g2(const uint32_t *pts, const uint32_t *start, const uint32_t *order, uint32_t nb, uint32_t heavy_thr,
                                 int shift, int into, uint32_t *buckets, cudaStream_t st);
// unit hooks: out[i] = 1/in[i] by divsteps (Montgomery form in and out); one standalone round over explicit affine pairs
void launch_dbg_inv_sg(int is_fp2, const uint32_t *in, uint32_t *out, size_t n, cudaStream_t st);

// jac_out[i] = 2^c · prev[i] (affine in, Jacobian out): one window step of the table build
void launch_table_shift_g1(const uint32_t *prev, size_t n, int c, uint32_t *jac_out, cudaStream_t st);
void launch_table_shift_g2(const uint32_t *prev, size_t n, int c, uint32_t *jac_out, cudaStream_t st);

// k_reduce_g{1,2}.cu
constexpr size_t WSUM_THREAD_FORM_MIN = 131072;   // chains from which a running-sum level runs one thread per chain instead of a quad
void launch_wsum_level_g1(const uint32_t *X, const uint32_t *Cin, uint32_t len, uint32_t m, int log2M, uint32_t nwin,
                          uint32_t *Rout, uint32_t *Cout, cudaStream_t st);
void launch_wsum_level_g2(const uint32_t *X, const uint32_t *Cin, uint32_t len, uint32_t m, int log2M, uint32_t nwin,
                          uint32_t *Rout, uint32_t *Cout, cudaStream_t st);
// one level of the log-depth tree over the per-window arrays (see reduce.cuh)
void launch_tree_level_g1(const uint32_t *Sin, size_t sin_stride, const uint32_t *Vin, const uint32_t *Cin, size_t cin_stride,
                          uint32_t *Sout, uint32_t *Vout, uint32_t *Cout, size_t out_stride, uint32_t S, int j, uint32_t nwin,
                          cudaStream_t st);
void launch_tree_level_g2(const uint32_t *Sin, size_t sin_stride, const uint32_t *Vin, const uint32_t *Cin, size_t cin_stride,
                          uint32_t *Sout, uint32_t *Vout, uint32_t *Cout, size_t out_stride, uint32_t S, int j, uint32_t nwin,
                          cudaStream_t st);
void launch_combine_g1(const uint32_t *Sroot, const uint32_t *V, const uint32_t *Croot, size_t stride, int logS, int log2M,
                       int nwin, int c, int split_top, uint32_t *wsum, uint32_t *out, cudaStream_t st);
void launch_combine_g2(const uint32_t *Sroot, const uint32_t *V, const uint32_t *Croot, size_t stride, int logS, int log2M,
                       int nwin, int c, int split_top, uint32_t *wsum, uint32_t *out, cudaStream_t st);
void launch_sum_partials_g1(const uint32_t *partials, int count, uint32_t *out, cudaStream_t st);
void launch_sum_partials_g2(const uint32_t *partials, int count, uint32_t *out, cudaStream_t st);

// k_normalize.cu
void launch_normalize_batch(int g2, const uint32_t *proj, size_t n, uint32_t *aff, int sm_count, cudaStream_t st);

// k_serde.cu
void launch_deserialize(int g2, const uint8_t *in, size_t n, int compressed, int validate, uint32_t *aff, uint8_t *status,
                        cudaStream_t st);
void launch_serialize(int g2, const uint32_t *aff, size_t n, int compressed, uint8_t *out, cudaStream_t st);

// k_synth_g{1,2}.cu / k_util.cu
void launch_synth_bases_g1(uint64_t seed, size_t n, uint32_t *out, cudaStream_t st);
void launch_synth_bases_g2(uint64_t seed, size_t n, uint32_t *out, cudaStream_t st);
void launch_synth_scalars(uint64_t seed, size_t n, int mont, uint32_t *out, cudaStream_t st);
void launch_imad_peak(int mode, int blocks, int threads, uint32_t *buf, int iters, cudaStream_t st);
void launch_dbg_field_op(int is_fp2, int op, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n);
void launch_dbg_point_op_g1(int op, const uint32_t *acc, const uint32_t *q, uint32_t *out, size_t n);
void launch_dbg_point_op_g2(int op, const uint32_t *acc, const uint32_t *q, uint32_t *out, size_t n);

// number of this library's own kernel launches since load (bench.py reports the per-step count)
extern unsigned long long g_own_launches;
extern thread_local unsigned long long t_own_launches;   // the calling thread's share (graph recording subtracts it again)
inline void count_launch() {
    __atomic_fetch_add(&g_own_launches, 1ull, __ATOMIC_RELAXED);
    t_own_launches++;
}

inline unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace b200msm
