// Integer-pipe microbenchmark for B200 (sm_100a): the roofline denominator of the MSM engine.
// MEASURED_PEAKS.json carries HBM and bf16 figures only; this measures what the Fp kernels are
// bound by: issue rate of 32-bit IMAD, IMAD.HI, IMAD.WIDE (64-bit accumulate) and of the
// carry-chained IMAD.WIDE.X form that ptxas emits for mad.lo.cc/madc.hi.cc pairs.
// Prints one JSON object. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_peak
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define INNER 64

template <int MODE>
__global__ void __launch_bounds__(256) bench(uint32_t *out, uint32_t seed, int iters, long long *cyc) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + 1 + threadIdx.x * 2;
    uint32_t r[CHAINS * 2];
#pragma unroll
    for (int i = 0; i < CHAINS * 2; i++) r[i] = seed + i + threadIdx.x * 7;
    uint64_t w[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) w[i] = ((uint64_t)r[i] << 32) | r[i + 1];
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < INNER; k++) {
            if (MODE == 0) {  // IMAD lo
#pragma unroll
                for (int c = 0; c < CHAINS; c++)
                    asm volatile("mad.lo.u32 %0, %0, %2, %1;" : "+r"(r[c]) : "r"(a), "r"(b));
            } else if (MODE == 1) {  // IMAD.HI
#pragma unroll
                for (int c = 0; c < CHAINS; c++)
                    asm volatile("mad.hi.u32 %0, %0, %2, %1;" : "+r"(r[c]) : "r"(a), "r"(b));
            } else if (MODE == 2) {  // IMAD.WIDE, independent 64-bit accumulators
#pragma unroll
                for (int c = 0; c < CHAINS; c++)
                    w[c] += (uint64_t)(uint32_t)(w[(c + 1) % CHAINS] >> 32) * b;
            } else if (MODE == 3) {  // carry chain: one row of the Montgomery product (lo/hi pairs)
                asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(r[0]) : "r"(a), "r"(r[15]));
                asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(r[1]) : "r"(a), "r"(b));
#pragma unroll
                for (int c = 1; c < CHAINS; c++) {
                    asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(r[2 * c]) : "r"(a), "r"(b));
                    asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(r[2 * c + 1]) : "r"(a), "r"(b));
                }
            } else if (MODE == 4) {  // LOP3 (alu pipe)
#pragma unroll
                for (int c = 0; c < CHAINS; c++)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(a), "r"(r[(c + 1) % CHAINS]));
            } else if (MODE == 5) {  // IMAD.WIDE + LOP3 mixed 1:1 (do the pipes overlap?)
#pragma unroll
                for (int c = 0; c < CHAINS / 2; c++) {
                    w[c] += (uint64_t)(uint32_t)(w[(c + 1) % (CHAINS / 2)] >> 32) * b;
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;"
                                 : "+r"(r[CHAINS + c])
                                 : "r"(a), "r"(r[CHAINS + ((c + 1) % (CHAINS / 2))]));
                }
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS * 2; i++) s ^= r[i];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s ^= (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char *name, double ops_per_inner, int nsm, int blocks_per_sm, bool last) {
    int blocks = nsm * blocks_per_sm, threads = 256, iters = 2000;
    uint32_t *out;
    long long *cyc;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaMalloc(&cyc, blocks * sizeof(long long));
    bench<MODE><<<blocks, threads>>>(out, 12345, 10, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    long long c0 = 0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        bench<MODE><<<blocks, threads>>>(out, 12345, iters, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) {
            best = ms;
            cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
        }
    }
    double ops = (double)blocks * threads * iters * INNER * ops_per_inner;
    double per_clk_sm = (double)threads * blocks_per_sm * iters * INNER * ops_per_inner / (double)c0;
    printf("  \"%s\": {\"ms\": %.4f, \"Tops_per_s\": %.4f, \"ops_per_clk_per_sm\": %.2f, \"cycles\": %lld, "
           "\"eff_mhz\": %.1f}%s\n",
           name, best, ops / (best * 1e-3) / 1e12, per_clk_sm, c0, (double)c0 / (best * 1e-3) / 1e6,
           last ? "" : ",");
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d,\n", p.name, nsm, p.clockRate);
    const int bps = 4;  // 1024 threads/SM = 8 warps/SMSP
    run<0>("imad_lo", CHAINS, nsm, bps, false);
    run<1>("imad_hi", CHAINS, nsm, bps, false);
    run<2>("imad_wide", CHAINS, nsm, bps, false);
    run<3>("imad_wide_carry_chain", CHAINS, nsm, bps, false);  // counted as WIDE ops (lo+hi pair = 1)
    run<4>("lop3", CHAINS, nsm, bps, false);
    run<5>("imad_wide_plus_lop3", CHAINS, nsm, bps, true);  // total instructions
    printf("}\n");
    return 0;
}
