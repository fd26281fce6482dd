// Planning of one MSM pass: window width, window count, GLV on/off — pure host arithmetic, no
// CUDA types, so it is unit-tested without a device (tests/cpp/test_plan.cpp, tests/test_abi.py).
// Replaces SingleMultiexpKernel::calc_window_size (reference src/gpu.rs:218-223: w = ⌈log2 n⌉ − 3
// clamped to 10 because its kernel keeps the buckets of every thread in global memory).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>

namespace b200msm {

struct Plan {
    int c = 0, nwin = 0;
    bool glv = false;          // split every scalar into two 128-bit halves (k1 + k2·λ) over P and φ(P) = (β·x, y) …
    int parts = 1;             // … or, on G2, into four 64-bit digits base |z| over Q, −ψ(Q), ψ²(Q), −ψ³(Q) (gls4.cuh): 1 / 2 / 4
    bool split = false;        // GLV with c | part bits: unsigned top digit spread over the last two windows (k_hist)
    uint32_t nbw = 0, nb = 0;  // buckets per window, total
};

// Windows of a plan. Plain: c·W ≥ 256 so the top Booth carry stays inside. The parts of a decomposed
// scalar are below 2^128 (two) or 2^64 (four): when c divides that width their top c bits are taken
// unsigned (digit ≤ 2^c → two windows' worth of buckets, no carry window); otherwise c·W ≥ width + 1.
inline int plan_windows(int parts, int c, bool *split) {
    const int bits = 256 / parts;
    *split = parts > 1 && bits % c == 0;
    if (parts == 1) return (256 + c - 1) / c;
    return *split ? bits / c + 1 : (bits + 1 + c - 1) / c;
}
inline int plan_windows(bool glv, int c, bool *split) { return plan_windows(glv ? 2 : 1, c, split); }

// Window width and GLV choice from a time model fitted to the measured phases on B200 (µs):
// accumulation at 88 % (G1) / 76 % (G2) of the 18.5 T IMAD/s pipe, the fan-in-32 reduction level
// at 55 %, 3.4 µs (G1) / 11 µs (G2) per dependent doubling of the Horner chain, 23 ps per sorted
// entry.  Without GLV this lands on the work-minimising c* of SURVEY §8(d) (13/16/18/20 at
// 2^16/20/22/24) — the width the roofline numerator assumes.
// Batched-affine pairing rounds in front of the XYZZ accumulation (batch_affine.cuh): how many, from the mean
// bucket occupancy (measured on B200, profiles/r02_experiments.md: three rounds pay from ≈24 entries per bucket,
// one from ≈12; below that the padding of every bucket to 2^R entries costs more than the affine additions save),
// and what they make of the accumulation time (measured 0.74–0.80 with three rounds, 0.93 with one).
inline int ba_rounds_for(double avg_occupancy) { return avg_occupancy >= 24 ? 3 : (avg_occupancy >= 12 ? 1 : 0); }
inline double ba_time_factor(double avg_occupancy, bool g2, int ba_mode) {
    const int R = ba_mode < 0 ? ba_rounds_for(avg_occupancy) : ba_mode;
    if (R == 0) return 1.0;
    if (R == 1) return 0.93;
    return g2 ? 0.75 : 0.79;
}
inline double plan_time_us(size_t n, bool g2, int parts, int c, int ba_mode = -1) {
    bool split;
    const bool glv = parts > 1;
    const double W = plan_windows(parts, c, &split), entries = (double)parts * (double)n;
    const double Wacc = split ? W - 1 : W;  // an entry lands in one of the two top windows
    const double madd = g2 ? 28 : 10, add = g2 ? 40 : 14, pipe = 18.5e6;  // IMAD per µs
    // one thread per bucket: below ≈2.5 warps per scheduler (4 × 148 of them) the dependent-issue latency of the
    // product chain shows (measured with 2^15 buckets: 0.6 of the pipe)
    // (GLV, round 1: the φ(P) entries read x from the β·x table and y from the base record — 2.5 % / 6 % slower in the XYZZ-only kernel;
    //  with the batched-affine rounds in front the difference is gone)
    const double wps = W * std::pow(2.0, c - 1) / 32 / 592;
    // (four parts on G2: the images come from three full-point tables, 768 B of gather per base instead of 288 — measured
    // with the batched-affine rounds: accumulate 15.15 → 15.5 ms at 2^20, 57.6 → 59.6 at 2^22; two parts cost nothing there)
    const double eff = std::min(g2 ? 0.76 : 0.88, 0.35 * wps) * (glv && parts == 4 ? 0.97 : 1.0);   // (G1 2^22: accumulate 18.47 ms with and without the two-part split)
    double t = entries * Wacc * madd * 588 / (pipe * eff) * ba_time_factor(entries / std::pow(2.0, c - 1), g2, ba_mode);
    t += W * std::pow(2.0, c - 1) * 2 * add * 588 / (pipe * 0.55);
    t += (split ? (W - 2) * c + c - 1 : (W - 1) * c) * (g2 ? 11.0 : 3.4) + 250;
    t += entries * Wacc * 2.3e-5;   // grouping (an entry of a split plan lands in ONE of the two top windows)
    // a plain plan whose top window keeps only a few bits (c = 18 at 2^23: 3 bits) piles its n entries onto a handful
    // of bucket counters: the grouping's atomics serialise (measured: digits 4.8 ms instead of 2.4 at 2^23)
    if (!glv && 255 - ((int)W - 1) * c < c - 5) t += (double)n * 3e-4;
    if (glv) {
        t += (double)n * (parts == 4 ? 12 : (g2 ? 2 : 1)) * 588 / (pipe * 0.8) + (double)n * (parts == 4 ? 1.2e-4 : 2e-5);  // image table(s), decomposition (k_endo_table: 44 µs at G1 2^20)
        // a top window with few bits piles its entries into few buckets: block-cooperative path, ≈3.5× the cost
        // The halves are below λ ≈ 0.673·2^128, so the top window only uses ⌊λ / 2^(c(W−1))⌋ + 1 of its buckets.
        // When that is few, its entries pile up past the heavy-bucket threshold and go down the block-cooperative
        // path (and serialise the grouping's atomics): never pick such a width.
        // (four parts: digits below |z| ≈ 0.82·2^64)
        const int shift = c * ((int)W - 1);
        const double top_buckets = std::floor((parts == 4 ? 0.82 : 0.673) * std::pow(2.0, 256 / parts - shift)) + 1;
        const double thr = std::max(std::max(32.0, 3 * entries / std::pow(2.0, c - 1)), entries * W / 175000);
        if (!split && entries / top_buckets > 0.7 * thr) t += 1e5;
    }
    return t;
}
inline void auto_plan(size_t n, bool g2, int glv_mode, int c_override, Plan &pl, int ba_mode = -1) {
    // glv_mode: -1 automatic, 0 never, 1 two parts (φ), 2 four parts on G2 (ψ; G1 has no such endomorphism: two)
    double best = 1e300;
    for (int parts = 1; parts <= (g2 ? 4 : 2); parts *= 2) {
        if (glv_mode == 0 && parts > 1) continue;
        if (glv_mode == 1 && parts != 2) continue;
        if (glv_mode == 2 && parts != (g2 ? 4 : 2)) continue;
        if (parts > 1 && glv_mode < 0 && n > (1u << 23)) continue;  // automatic choice only where it was measured to pay (2^24: 77.4 ms plain, 80.0 split)
        if (parts > 1 && (uint64_t)parts * n >= (1ull << 31)) continue;
        for (int c = 2; c <= 22; c++) {
            if (c_override > 0 && c != c_override) continue;
            double t = plan_time_us(n, g2, parts, c, ba_mode);
            if (t < best) { best = t; pl.c = c; pl.parts = parts; }
        }
    }
    pl.glv = pl.parts > 1;
    pl.nwin = plan_windows(pl.parts, pl.c, &pl.split);
    pl.nbw = 1u << (pl.c - 1);
    pl.nb = pl.nbw * (uint32_t)pl.nwin;
}
inline int auto_window(size_t n, bool g2) {  // non-GLV width (scratch estimates)
    Plan pl;
    auto_plan(n, g2, 0, 0, pl);
    return pl.c;
}

// ---- shapes of the batched-affine pairing rounds (batch_affine.cuh) — host arithmetic, unit-tested without a device ----
// One round over at most s_out_max output slots: NT threads of K slots each, then NU second-level threads of K2 totals
// each.  The second level and the inversions are latency-bound (measured at G1 2^20: ≈1.5 µs per dependent product,
// ≈50 µs per divsteps inversion however few there are), so K is as large as two waves of 3 blocks × 128 threads per SM
// allow (≤ 32) and K2 keeps the inversions near one warp per scheduler.
struct BaPlan { uint32_t NT, K, NU, K2; };
inline BaPlan ba_plan(size_t s_out_max, int sm_count) {
    BaPlan bp;
    const size_t wave = (size_t)sm_count * 3 * 128;
    size_t K = (s_out_max + 2 * wave - 1) / (2 * wave);
    K = K < 4 ? 4 : (K > 32 ? 32 : K);
    size_t NT = (s_out_max + K - 1) / K;
    NT = (NT + 127) / 128 * 128;
    if (NT == 0) NT = 128;
    size_t K2 = (NT + 16383) / 16384;
    K2 = K2 < 8 ? 8 : (K2 > 64 ? 64 : K2);
    bp.NT = (uint32_t)NT;
    bp.K = (uint32_t)K;
    bp.K2 = (uint32_t)K2;
    bp.NU = (uint32_t)((NT + K2 - 1) / K2);
    return bp;
}
// The rounds of a pass: per round the launch plan of one pipeline (the whole slot range, or one of its halves when
// the rounds run as two pipelines on two streams — from 2^20 output slots), and the scratch ONE pipeline needs over
// ALL rounds: the maxima, because the two pipelines may be a round apart and the per-round sizes are not monotone
// (K shrinks with the slot count, so NT = ⌈slots/K⌉ can grow from one round to the next).  A first version placed
// part 1's scratch by the current round's sizes, a second by the first round's: both lost results (tests/cpp/test_plan.cpp
// checks that no round of either pipeline ever leaves its part).
struct BaLayout {
    bool split = false;
    BaPlan bp[3];
    size_t s_part[3] = {0, 0, 0};             // slots one pipeline covers at most, per round
    size_t pre_el = 0, t_el = 0, u_el = 0;    // elements per pipeline: prefix, T / prefix2, U
};
inline BaLayout ba_layout(size_t s1, int R, int sm_count, bool split_ok = true) {
    BaLayout L;
    L.split = split_ok && s1 >= ((size_t)1 << 20);
    size_t s_out = s1;
    for (int r = 0; r < R && r < 3; r++) {
        // a half is bounded by half the round's slots plus the boundary's rounding (2^(6+R) entries ≫ r+1 ≤ 2^(5+R))
        L.s_part[r] = L.split ? s_out / 2 + ((size_t)1 << (5 + R)) + 1 : s_out;
        L.bp[r] = ba_plan(L.s_part[r], sm_count);
        L.pre_el = std::max(L.pre_el, (size_t)L.bp[r].NT * L.bp[r].K);
        L.t_el = std::max<size_t>(L.t_el, L.bp[r].NT);
        L.u_el = std::max<size_t>(L.u_el, L.bp[r].NU);
        s_out = (s_out + 1) / 2;
    }
    return L;
}

// Fixed-base window table (table[w][i] = 2^(c·w)·P_i): one bucket set for all windows, no Horner
// chain, so the width only trades the n·W additions of the accumulation against one window's
// bucket reduction (+ a log-depth tree) — wider than the plain plan's at every n.
inline int table_plan(size_t n, bool g2) {
    const double madd = g2 ? 28 : 10, add = g2 ? 40 : 14, pipe = 18.5e6;
    double best = 1e300;
    int bc = 10;
    for (int c = 10; c <= 23; c++) {           // W ≤ 26: the table stays within 26× the bases
        const double W = std::ceil(256.0 / c);
        // one thread per bucket and only 2^(c−1) buckets in all: below ≈2.5 warps per scheduler (4 × 148 of them)
        // the dependent-issue latency of the product chain shows (measured: c = 16 → 0.6 of the pipe)
        const double wps = std::pow(2.0, c - 1) / 32 / 592, eff = std::min(g2 ? 0.76 : 0.88, 0.35 * wps);
        double t = (double)n * W * madd * 588 / (pipe * eff) + std::pow(2.0, c - 1) * 2 * add * 588 / (pipe * 0.55) +
                   (double)n * W * 2.3e-5 + c * (g2 ? 40.0 : 16.0);
        // the top window only holds 255 − (W−1)·c bits: when that is few, its n entries pile into a
        // handful of buckets and go down the block-cooperative path (≈3.5× the cost per entry)
        const int topbits = std::max(0, 255 - ((int)W - 1) * c);
        const double m = (double)n * W, thr = std::max(std::max(32.0, 4 * m / std::pow(2.0, c - 1)), m / 175000);
        if (topbits < c - 1 && (double)n / std::pow(2.0, topbits) > thr) t += (double)n * madd * 588 / pipe * 3.5;
        if (topbits < c - 5) t += 1e4;  // … and their counters serialise the grouping's atomics: never pick such a width
        if (t < best) { best = t; bc = c; }
    }
    return bc;
}

}  // namespace b200msm
