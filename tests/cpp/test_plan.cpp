// Host-only unit test of the planner (ark_blst_b200/csrc/plan.h): compiled with g++, no CUDA.
// The plain plan must keep every scalar bit plus the Booth carry inside its windows and land on
// the work-minimising c* of SURVEY §8(d) at the BASELINE sizes; GLV plans must cover 128-bit
// halves; table plans must only use widths whose top window is not degenerate.
#include <cstdio>

#include "../../ark_blst_b200/csrc/plan.h"

using namespace b200msm;
static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); failures++; } } while (0)

int main() {
    for (int g2 = 0; g2 < 2; g2++) {
        for (int logn = 1; logn <= 27; logn++) {
            for (int mode = -1; mode <= 2; mode++) {
                Plan p;
                auto_plan((size_t)1 << logn, g2, mode, 0, p);
                CHECK(p.c >= 2 && p.c <= 22 && p.nwin >= 1);
                CHECK(p.nbw == 1u << (p.c - 1) && p.nb == p.nbw * (uint32_t)p.nwin);
                CHECK(p.parts == 1 || p.parts == 2 || (p.parts == 4 && g2));
                CHECK(p.glv == (p.parts > 1));
                const int bits = 256 / p.parts;       // width of a part of the decomposed scalar
                if (!p.glv) CHECK(p.c * p.nwin >= 256 && (p.nwin - 1) * p.c < 256);
                else if (p.split) CHECK(bits % p.c == 0 && p.nwin == bits / p.c + 1);
                else CHECK(p.c * p.nwin >= bits + 1 && (p.nwin - 1) * p.c < bits + 1);
                if (mode == 0) CHECK(!p.glv);
                if (mode == 1) CHECK(p.parts == 2);
                if (mode == 2) CHECK(p.parts == (g2 ? 4 : 2));
                if (mode == -1 && logn > 23) CHECK(!p.glv);
            }
            for (int c = 2; c <= 22; c++) {  // an explicit width is honoured
                Plan p;
                auto_plan((size_t)1 << logn, g2, 0, c, p);
                CHECK(p.c == c && !p.glv);
            }
            int tc = table_plan((size_t)1 << logn, g2);
            int tw = (256 + tc - 1) / tc;
            CHECK((tc == 10 || tc == 13 || tc == 16 || tc == 20) && 255 - (tw - 1) * tc >= tc - 5);
        }
        Plan p;
        auto_plan(1u << 16, g2, 0, 0, p, 0); CHECK(p.c == 13 && p.nwin == 20);   // (batched-affine rounds off: the plain work model)
        auto_plan(1u << 20, g2, 0, 0, p, 0); CHECK(p.c == 16 && p.nwin == 16);
        auto_plan(1u << 24, g2, 0, 0, p, 0); CHECK(p.c == 20 && p.nwin == 13);
        CHECK(ba_rounds_for(64) == 3 && ba_rounds_for(16) == 1 && ba_rounds_for(4) == 0);
        auto_plan(1u << 20, g2, -1, 0, p); CHECK(p.glv && p.split && p.c == 16 && (p.parts == 2 ? p.nwin == 9 : (g2 && p.parts == 4 && p.nwin == 5)));
        auto_plan(1u << 20, g2, 2, 0, p); CHECK(p.glv && p.split && p.c == 16 && p.nwin == (g2 ? 5 : 9) && p.parts == (g2 ? 4 : 2));
    }
    // batched-affine rounds: one pipeline's scratch part holds every round of that pipeline, whatever the size —
    // NT·K covers the slots, and no round needs more than the per-pipeline maxima the layout reserves and offsets by
    for (int R = 1; R <= 3; R++)
        for (size_t s1 = 1; s1 < ((size_t)1 << 29); s1 = s1 * 3 / 2 + 977) {
            for (int split_ok = 0; split_ok <= 1; split_ok++) {
                BaLayout L = ba_layout(s1, R, 148, split_ok);
                CHECK(L.split == (split_ok && s1 >= ((size_t)1 << 20)));
                size_t s_out = s1;
                for (int r = 0; r < R; r++) {
                    const BaPlan &b = L.bp[r];
                    CHECK((size_t)b.NT * b.K >= L.s_part[r] && b.NT % 128 == 0 && b.K >= 4 && b.K <= 32);
                    CHECK((size_t)b.NU * b.K2 >= b.NT && b.K2 >= 8 && b.K2 <= 64);
                    CHECK((size_t)b.NT * b.K <= L.pre_el && b.NT <= L.t_el && b.NU <= L.u_el);
                    // the two halves of a round cover its slots: part 0 ≤ half, part 1 ≤ half + the boundary's rounding
                    if (L.split) CHECK(2 * L.s_part[r] >= s_out && L.s_part[r] >= s_out / 2 + ((size_t)1 << (5 + R - r - 1)));
                    else CHECK(L.s_part[r] == s_out);
                    s_out = (s_out + 1) / 2;
                }
            }
        }
    {   // the sizes that broke the earlier layouts: NT grows from round 0 to round 1 when K halves with the slot count
        BaLayout L = ba_layout(5750000, 3, 148, true);
        CHECK(L.split && L.bp[1].NT > L.bp[0].NT && L.t_el >= L.bp[1].NT);
    }
    std::printf(failures ? "FAILED (%d)\n" : "ok\n", failures);
    return failures ? 1 : 0;
}
