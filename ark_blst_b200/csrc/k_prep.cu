// Scalar preparation kernels: window digits, bucket offsets, bucket-size keys, and the device
// radix sorts (CUB DeviceRadixSort on sm_100a; see DESIGN.md for why the sort is not the bottleneck).
#include <cub/device/device_radix_sort.cuh>

#include "launch.h"
#include "scalar.cuh"

namespace b200msm {

// keys/vals are window-major: entry (w, i) at w·n + i, so every store is coalesced.
__global__ void __launch_bounds__(256)
k_digits(const uint32_t *__restrict__ scalars, size_t n, int mont, int c, int nwin,
         uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    load_scalar(s, scalars, i, mont);
    const uint32_t nbw = 1u << (c - 1);
    const uint32_t sentinel = nbw * (uint32_t)nwin;  // zero digits sort past every real bucket
    for (int w = 0; w < nwin; w++) {
        int d = booth_digit(s, w, c);
        uint32_t neg = d < 0;
        uint32_t mag = neg ? (uint32_t)(-d) : (uint32_t)d;
        keys[(size_t)w * n + i] = mag ? (uint32_t)w * nbw + (mag - 1) : sentinel;
        vals[(size_t)w * n + i] = (uint32_t)i | (neg << 31);
    }
}
__global__ void k_digits_dbg(const uint32_t *__restrict__ scalars, size_t n, int mont, int c, int nwin,
                             int *__restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8];
    load_scalar(s, scalars, i, mont);
    for (int w = 0; w < nwin; w++) out[(size_t)w * n + i] = booth_digit(s, w, c);
}

// start[b] = first sorted position whose key ≥ b, for b in [0, nb+1]; m = number of entries.
// Bucket b owns [start[b], start[b+1]); bucket nb is the sentinel (zero digits), start[nb+1] = m.
// Position j opens every bucket in (keys[j-1], keys[j]]; each b is written by exactly one j.
__global__ void __launch_bounds__(256)
k_bounds(const uint32_t *__restrict__ keys, size_t m, uint32_t nb, uint32_t *__restrict__ start) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j > m) return;
    uint32_t lo = j ? keys[j - 1] + 1 : 0;
    uint32_t hi = j < m ? keys[j] : nb + 1;
    for (uint32_t b = lo; b <= hi; b++) start[b] = (uint32_t)j;
}
// sort key for the size ordering: min(count, 2^bits - 1)
__global__ void __launch_bounds__(256)
k_counts(const uint32_t *__restrict__ start, uint32_t nb, uint32_t clampv,
         uint32_t *__restrict__ cnt, uint32_t *__restrict__ ids) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    uint32_t c = start[b + 1] - start[b];
    cnt[b] = c < clampv ? c : clampv;
    ids[b] = b;
}


void launch_digits(const uint32_t *scalars, size_t n, int mont, int c, int nwin, uint32_t *keys, uint32_t *vals,
                   cudaStream_t st) {
    count_launch();
    k_digits<<<blocks_for(n, 256), 256, 0, st>>>(scalars, n, mont, c, nwin, keys, vals);
}
void launch_digits_dbg(const uint32_t *scalars, size_t n, int mont, int c, int nwin, int *out, cudaStream_t st) {
    count_launch();
    k_digits_dbg<<<blocks_for(n, 256), 256, 0, st>>>(scalars, n, mont, c, nwin, out);
}
void launch_bounds(const uint32_t *keys, size_t m, uint32_t nb, uint32_t *start, cudaStream_t st) {
    count_launch();
    k_bounds<<<blocks_for(m + 1, 256), 256, 0, st>>>(keys, m, nb, start);
}
void launch_counts(const uint32_t *start, uint32_t nb, uint32_t clampv, uint32_t *cnt, uint32_t *ids, cudaStream_t st) {
    count_launch();
    k_counts<<<blocks_for(nb, 256), 256, 0, st>>>(start, nb, clampv, cnt, ids);
}
cudaError_t sort_pairs(void *tmp, size_t *tmp_bytes, uint32_t *k0, uint32_t *k1, uint32_t *v0, uint32_t *v1, size_t m,
                       int end_bit, bool descending, int *sel, cudaStream_t st) {
    cub::DoubleBuffer<uint32_t> dk(k0, k1), dv(v0, v1);
    cudaError_t e = descending ? cub::DeviceRadixSort::SortPairsDescending(tmp, *tmp_bytes, dk, dv, m, 0, end_bit, st)
                               : cub::DeviceRadixSort::SortPairs(tmp, *tmp_bytes, dk, dv, m, 0, end_bit, st);
    if (sel) *sel = dk.selector;
    return e;
}

}  // namespace b200msm
