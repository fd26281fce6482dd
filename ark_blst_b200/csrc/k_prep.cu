// Scalar preparation kernels: window digits, grouping of the point indices by bucket (a counting
// sort on the bucket id — the pipeline's "radix sort by bucket", hand-written: no library kernel
// runs on the MSM path) and the ordering of the buckets by size.

#include "launch.h"
#include "gls4.cuh"
#include "scalar.cuh"

namespace b200msm {

__global__ void k_digits_dbg(const uint32_t *__restrict__ scalars, size_t n, int mont, int c, int nwin,
                             int *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    load_scalar(s, scalars, i, mont);
    for (int w = 0; w < nwin; w++) out[(size_t)w * n + i] = booth_digit(s, w, c);
}

// ---- bucket grouping without a general-purpose sort -------------------------------------------
// Only the grouping by bucket matters (order inside a bucket is irrelevant to the sum), so the
// "radix sort by bucket" is a counting sort on the full bucket id in two passes: (1) canonicalise
// the scalar, write its window digits (window-major) and histogram them (one RED per non-zero
// digit), exclusive scan, (2) scatter point indices to start[b] + cursor[b]++, window by window.
// Zero digits are dropped instead of being carried to a sentinel bucket, and the scan directly
// yields the bucket offsets.
// `glv` (G1 only): every scalar is split into (k1, k2) with k = k1 + k2·λ; entry i of a window is
// k1's digit for point i, entry n + i is k2's digit for the endomorphism image φ(P_i).
// glv: 0 none; 1 / 2 two parts (k = k1 + k2·λ over P, φ(P)) with a carry window / with the unsigned top digit;
// 3 / 4 four parts (G2: base-|z| digits over Q, −ψ(Q), ψ²(Q), −ψ³(Q), gls4.cuh) with a carry window / unsigned top digit.
// Entry part·n + i of a window is part `part`'s digit for point i.
__global__ void __launch_bounds__(256)
k_hist(const uint32_t *__restrict__ scalars, size_t n, int mont, int glv, int c, int nwin, int shared_buckets,
       uint32_t *__restrict__ dig, uint32_t *__restrict__ count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[4][8];
    load_scalar(s[0], scalars, i, mont);
    const int parts = glv == 0 ? 1 : (glv <= 2 ? 2 : 4);
    if (parts == 2) {
        uint32_t k[8];
#pragma unroll
        for (int j = 0; j < 8; j++) k[j] = s[0][j];
        glv_decompose(k, s[0], s[1]);
    } else if (parts == 4) {
        uint32_t k[8];
#pragma unroll
        for (int j = 0; j < 8; j++) k[j] = s[0][j];
        gls4_decompose(k, s);
    }
    const uint32_t nbw = 1u << (c - 1);
    const size_t per_window = (size_t)parts * n;
    // split (c divides the parts' width): the top c bits of a part are taken UNSIGNED, so no carry
    // window follows them; their digit u ∈ [0, 2^c] goes to window nwin−2 when u ≤ 2^(c−1) and, as
    // u − 2^(c−1), to window nwin−1 otherwise (k_combine adds 2^(c−1)·Σ buckets for that window
    // and gives both the weight of window nwin−2)
    const int split = glv == 2 || glv == 4;
#pragma unroll
    for (int part = 0; part < 4; part++) {         // (unrolled: s[part] stays in registers)
        if (part >= parts) break;
        const uint32_t *sc = s[part];
        uint32_t utop = 0;
        if (split) utop = scalar_bits(sc, (nwin - 2) * c - 1, c + 1), utop = (utop >> 1) + (utop & 1);
        for (int w = 0; w < nwin; w++) {
            int d;
            if (split && w >= nwin - 2) {
                const uint32_t hi = utop > nbw;
                d = w == nwin - 2 ? (hi ? 0 : (int)utop) : (hi ? (int)(utop - nbw) : 0);
            } else d = booth_digit(sc, w, c);
            uint32_t neg = d < 0;
            uint32_t mag = neg ? (uint32_t)(-d) : (uint32_t)d;
            dig[(size_t)w * per_window + (size_t)part * n + i] = (mag << 1) | neg;   // window-major; 0 = zero digit
            if (mag) atomicAdd(&count[(shared_buckets ? 0u : (uint32_t)w * nbw) + (mag - 1)], 1u);
        }
    }
}
// grid (point blocks, windows): blocks are dispatched window after window, so at any time the
// scattered stores fall into one window's slice of vals (n·4 bytes) and its cursors — a working
// set that stays in the 126 MB L2 up to n = 2^24 instead of spraying the whole n·W·4-byte array
__global__ void __launch_bounds__(256)
k_scatter(const uint32_t *__restrict__ dig, size_t n, int c, size_t tbl_stride, const uint32_t *__restrict__ start,
          uint32_t *__restrict__ cursor, uint32_t *__restrict__ vals) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t w = blockIdx.y;
    uint32_t d = dig[(size_t)w * n + i];
    if (!d) return;
    // table mode: one bucket set for all windows, the entry names table[w][i]. A bucket's cursor
    // only moves forward, so the stores of one window pass still fall into one open sector per bucket.
    uint32_t bk = (tbl_stride ? 0u : w << (c - 1)) + ((d >> 1) - 1);
    uint32_t pos = start[bk] + atomicAdd(&cursor[bk], 1u);
    vals[pos] = (uint32_t)(tbl_stride ? (size_t)w * tbl_stride + i : i) | (d << 31);
}

// exclusive scan of n u32 (n up to 2^27): 2048 elements per block, block sums scanned by one block
constexpr int SCAN_ITEMS = 8, SCAN_THREADS = 256, SCAN_TILE = SCAN_ITEMS * SCAN_THREADS;
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < SCAN_THREADS / 32 ? wsum[lane] : 0;
#pragma unroll
        for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        if (lane < SCAN_THREADS / 32) wsum[lane] = w;
    }
    __syncthreads();
    uint32_t base = wid ? wsum[wid - 1] : 0;
    *total = wsum[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return base + x - v;
}
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_tiles(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, size_t n, uint32_t *__restrict__ tile_sums, uint32_t pad_mask) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < n ? (in[base + k] + pad_mask) & ~pad_mask : 0;   // bucket sizes rounded up to the padding unit
        sum += v[k];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(sum, &total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_sums(uint32_t *__restrict__ tile_sums, size_t ntiles, uint32_t *__restrict__ grand_total) {
    uint32_t carry = 0;
    for (size_t base = 0; base < ntiles; base += SCAN_THREADS) {
        size_t i = base + threadIdx.x;
        uint32_t v = i < ntiles ? tile_sums[i] : 0, total;
        uint32_t ex = block_exclusive_scan(v, &total);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *grand_total = carry;
}
// out[i] += tile offset; also out[n] = out[n+1] = grand total (start[nb] and start[nb+1])
__global__ void __launch_bounds__(256)
k_scan_add(uint32_t *__restrict__ out, size_t n, const uint32_t *__restrict__ tile_sums, const uint32_t *__restrict__ grand_total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tile_sums[i / SCAN_TILE];
    if (i == 0) { out[n] = *grand_total; out[n + 1] = *grand_total; }
}

// ---- buckets ordered by decreasing size: one counting-sort pass on min(count, SIZE_BINS-1) ------
constexpr int SIZE_BINS = 4096, SIZE_THREADS = 256, SIZE_ITEMS = 4;
__global__ void __launch_bounds__(SIZE_THREADS)
k_size_hist(const uint32_t *__restrict__ start, uint32_t nb, uint32_t *__restrict__ hist) {
    __shared__ uint32_t lh[SIZE_BINS];
    for (int k = threadIdx.x; k < SIZE_BINS; k += SIZE_THREADS) lh[k] = 0;
    __syncthreads();
    uint32_t b0 = (blockIdx.x * SIZE_THREADS + threadIdx.x) * SIZE_ITEMS;
#pragma unroll
    for (int k = 0; k < SIZE_ITEMS; k++) {
        uint32_t b = b0 + k;
        if (b < nb) {
            uint32_t cn = start[b + 1] - start[b];
            atomicAdd(&lh[cn < SIZE_BINS - 1 ? cn : SIZE_BINS - 1], 1u);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < SIZE_BINS; k += SIZE_THREADS)
        if (lh[k]) atomicAdd(&hist[k], lh[k]);
}
// descending exclusive scan of the size histogram: binstart[s] = #buckets with size bin > s
__global__ void __launch_bounds__(SCAN_THREADS)
k_size_scan(const uint32_t *__restrict__ hist, uint32_t *__restrict__ binstart) {
    uint32_t carry = 0;
    for (int base = 0; base < SIZE_BINS; base += SCAN_THREADS) {
        int r = base + threadIdx.x;                // rank in descending order
        uint32_t v = hist[SIZE_BINS - 1 - r], total;
        uint32_t ex = block_exclusive_scan(v, &total);
        binstart[SIZE_BINS - 1 - r] = carry + ex;
        carry += total;
    }
}
__global__ void __launch_bounds__(SIZE_THREADS)
k_size_scatter(const uint32_t *__restrict__ start, uint32_t nb, uint32_t *__restrict__ bincursor,
               uint32_t *__restrict__ order) {
    __shared__ uint32_t lh[SIZE_BINS];   // local count, then global base of this block's run per bin
    for (int k = threadIdx.x; k < SIZE_BINS; k += SIZE_THREADS) lh[k] = 0;
    __syncthreads();
    uint32_t b0 = (blockIdx.x * SIZE_THREADS + threadIdx.x) * SIZE_ITEMS;
    uint32_t bin[SIZE_ITEMS], rank[SIZE_ITEMS];
#pragma unroll
    for (int k = 0; k < SIZE_ITEMS; k++) {
        uint32_t b = b0 + k;
        bin[k] = 0xffffffffu;
        if (b < nb) {
            uint32_t cn = start[b + 1] - start[b];
            bin[k] = cn < SIZE_BINS - 1 ? cn : SIZE_BINS - 1;
            rank[k] = atomicAdd(&lh[bin[k]], 1u);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < SIZE_BINS; k += SIZE_THREADS)
        if (lh[k]) lh[k] = atomicAdd(&bincursor[k], lh[k]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SIZE_ITEMS; k++)
        if (bin[k] != 0xffffffffu) order[lh[bin[k]] + rank[k]] = b0 + k;
}

void launch_group_by_bucket(const uint32_t *scalars, size_t n, int mont, int glv, int c, int nwin, uint32_t nb, uint32_t *dig,
                            uint32_t *count, uint32_t *start, uint32_t *tile_sums, uint32_t *vals, cudaStream_t st,
                            size_t tbl_stride, int pad_log) {
    for (int k = 0; k < 7; k++) count_launch();
    const uint32_t pad_mask = (1u << pad_log) - 1;
    const size_t parts = glv == 0 ? 1 : (glv <= 2 ? 2 : 4);
    if (pad_log) cudaMemsetAsync(vals, 0xff, (parts * n * (size_t)nwin + (size_t)nb * pad_mask) * 4, st);
    cudaMemsetAsync(count, 0, (size_t)nb * 4, st);
    k_hist<<<blocks_for(n, 256), 256, 0, st>>>(scalars, n, mont, glv, c, nwin, tbl_stride ? 1 : 0, dig, count);
    size_t ntiles = (nb + SCAN_TILE - 1) / SCAN_TILE;
    k_scan_tiles<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(count, start, nb, tile_sums, pad_mask);
    k_scan_sums<<<1, SCAN_THREADS, 0, st>>>(tile_sums, ntiles, tile_sums + ntiles);
    k_scan_add<<<blocks_for(nb, 256), 256, 0, st>>>(start, nb, tile_sums, tile_sums + ntiles);
    cudaMemsetAsync(count, 0, (size_t)nb * 4, st);   // reused as the scatter cursors
    const size_t entries = parts * n;  // per window
    k_scatter<<<dim3(blocks_for(entries, 256), (unsigned)nwin), 256, 0, st>>>(dig, entries, c, tbl_stride, start, count, vals);
}
// hist: 2·SIZE_BINS u32 of scratch
void launch_order_by_size(const uint32_t *start, uint32_t nb, uint32_t *hist, uint32_t *order, cudaStream_t st) {
    for (int k = 0; k < 3; k++) count_launch();
    cudaMemsetAsync(hist, 0, SIZE_BINS * 4, st);
    unsigned blocks = blocks_for(nb, SIZE_THREADS * SIZE_ITEMS);
    k_size_hist<<<blocks, SIZE_THREADS, 0, st>>>(start, nb, hist);
    k_size_scan<<<1, SCAN_THREADS, 0, st>>>(hist, hist + SIZE_BINS);
    k_size_scatter<<<blocks, SIZE_THREADS, 0, st>>>(start, nb, hist + SIZE_BINS, order);
}

void launch_digits_dbg(const uint32_t *scalars, size_t n, int mont, int c, int nwin, int *out, cudaStream_t st) {
    count_launch();
    k_digits_dbg<<<blocks_for(n, 256), 256, 0, st>>>(scalars, n, mont, c, nwin, out);
}
}  // namespace b200msm
