// Host side of the engine and the C-ABI of include/b200msm.h.
// Replaces reference src/gpu.rs (SingleMultiexpKernel + msm, :45-241): planning (window width,
// window count), device buffers, launches, and — unlike the reference, which downloads 18 944
// partials and folds them on the host (:185-207) — the whole reduction stays on the device and
// only the 144/288-byte result crosses PCIe.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200msm.h"
#include "launch.h"
#include "plan.h"

using namespace b200msm;

namespace b200msm {
unsigned long long g_own_launches = 0;
thread_local unsigned long long t_own_launches = 0;
}

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) cudaGetLastError(); /* do not leave it for the next call's check */ \
        if (e__ != cudaSuccess)                                                              \
            return fail(e__ == cudaErrorMemoryAllocation ? B200MSM_ENOMEM : B200MSM_ECUDA,   \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                \
    } while (0)

// (re)allocations of any device buffer of this library: cached graphs hold raw pointers into the arenas
std::atomic<unsigned> g_arena_epoch{0};

// grow-only device buffer
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        g_arena_epoch++;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            return fail(B200MSM_ENOMEM, std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e));
        }
        cap = want;
        return 0;
    }
    void release() {
        if (p) { cudaFree(p); g_arena_epoch++; }
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// Everything the launches of one pass depend on — also the key under which its CUDA graph is cached.
struct PassArgs {
    int group = 0, mont = 0;
    const uint32_t *d_bases = nullptr, *d_scalars = nullptr;
    uint32_t *d_out = nullptr;
    size_t n = 0;
    int c = 0, nwin = 0, glv = 0;          // plan; glv: 0 off; 1 / 2 two parts (φ) with a carry window / unsigned top digit; 3 / 4 four parts (ψ, G2)
    uint32_t nbw = 0, nb = 0;
    const void *tbl_p = nullptr;           // fixed-base table (or null)
    size_t tbl_stride = 0;
    int into = 0, finish = 1;
    uint32_t heavy_thr = 0;
    int R = 0;                             // batched-affine pairing rounds before the XYZZ accumulation
    unsigned arena = 0;                    // DeviceCtx::arena_epoch the scratch pointers belong to
};
bool same_pass(const PassArgs &a, const PassArgs &b) { return memcmp(&a, &b, sizeof a) == 0; }

// A pass seen for the second time with the same arguments is recorded as two CUDA graphs (run_pass)
struct PassGraph {
    PassArgs key;
    cudaGraphExec_t prep = nullptr, main = nullptr;
    unsigned long long last_use = 0;
    unsigned long long launches_prep = 0, launches_main = 0;   // kernels inside, for b200msm_launch_count
    int plan[4] = {0, 0, 0, 0};
};
struct TableRef {
    const void *p;     // window-major affine table on the device
    size_t stride;     // points per window
    int c, nwin;
};
struct DeviceCtx {
    int dev = -1, lane = 0;
    cudaStream_t tail_stream = nullptr;            // high priority: the latency-bound reduction + combination
    cudaEvent_t ev_tail_fork = nullptr, ev_tail_join = nullptr;
    std::mutex mu;
    cudaStream_t stream = nullptr, copy_stream = nullptr, aux_stream = nullptr;
    cudaEvent_t ev_scalars = nullptr, ev_bases = nullptr, ev_busy = nullptr, ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_share = nullptr;  // this device's share of a sharded MSM is complete (partial delivered to device 0)
    bool busy_valid = false;
    unsigned fits_epoch = 0;    // Tun::epoch the fit caches below were recorded under
    size_t fits_n[2] = {0, 0};  // largest n per group that already ran as a single pass (arena is big enough)
    size_t fits_tbl_n[2] = {0, 0};  // same for the table path, valid for window width fits_tbl_c
    int fits_tbl_c[2] = {0, 0};
    DevBuf bases, scalars, digits, vals, start, cnt, ord, buckets, lvlR[2], lvlC[2], out;
    DevBuf hvy_hdr, hvy_buckets, hvy_tasks, hvy_partials, treeS[2], treeV[2], treeC[2], wsum, chunk_partials, norm_in, norm_out, tile_sums, size_hist, endo, tbl_tmp;
    DevBuf ba_prefix, ba_T, ba_prefix2, ba_U, ba_pts[3];   // (one output array per round: the two half-range pipelines may be a round apart)   // batched-affine rounds (batch_affine.cuh)
    std::vector<PassGraph> graphs;   // CUDA graphs of passes seen before (run_pass)
    std::vector<PassArgs> seen;
    unsigned long long graph_clock = 0;
    cudaStream_t cap_stream = nullptr;      // graphs are recorded on this stream (never executed on it)
    unsigned arena_epoch() const { return g_arena_epoch.load(); }
    cudaEvent_t ev[8] = {};
    cudaEvent_t ev_slice[16] = {};  // per slice of a streamed MSM: scalars ready, bases ready
    double phase_ms[8] = {};
    int last_plan[4] = {0, 0, 0, 0};  // c, windows, GLV (0/1/2 = off / on / on with the unsigned top digit), table
    bool phase_pending = false;
    int sm_count = 0;
    std::vector<DevBuf *> scratch() {
        return {&digits, &vals, &start, &cnt, &ord, &buckets,
                &lvlR[0], &lvlR[1], &lvlC[0], &lvlC[1], &hvy_hdr, &hvy_buckets, &hvy_tasks, &hvy_partials, &treeS[0],
                &treeS[1], &treeV[0], &treeV[1], &treeC[0], &treeC[1], &wsum, &ba_prefix, &ba_T, &ba_prefix2, &ba_U, &ba_pts[0], &ba_pts[1], &ba_pts[2]};
    }
    size_t scratch_bytes() {
        size_t s = 0;
        for (DevBuf *b : scratch()) s += b->cap;
        return s;
    }
    void release_all() {
        for (DevBuf *b : {&bases, &scalars, &digits, &vals, &start, &cnt, &ord, &buckets, &lvlR[0], &lvlR[1], &lvlC[0], &lvlC[1], &out, &hvy_hdr, &hvy_buckets,
                          &hvy_tasks, &hvy_partials, &treeS[0], &treeS[1], &treeV[0], &treeV[1], &treeC[0], &treeC[1], &wsum, &chunk_partials, &norm_in, &norm_out, &tile_sums, &size_hist, &endo, &tbl_tmp, &ba_prefix, &ba_T, &ba_prefix2, &ba_U, &ba_pts[0], &ba_pts[1], &ba_pts[2]})
            b->release();
    }
};

// Tunables (the b200msm_set_* calls).  A call takes ONE snapshot when it starts and plans every
// pass, chunk and slice from it, so a setter racing with running MSMs on other host threads never
// changes a plan half way through (tests/test_gpu_threads.py flips them under load).
struct Tun {
    int window_override = 0;
    size_t max_chunk_override = 0;
    int glv_mode = -1;  // -1 automatic (time model; default), 0 never, 1 two parts (φ), 2 four parts on G2 (ψ; G1: two)
    bool profiling = false;
    int stream_slices = 8;          // (at most) host-buffer MSMs of ≥ stream_min points per device are uploaded and accumulated in slices
    size_t stream_min = 1u << 18;
    int heavy_factor = 0;  // a bucket is heavy above heavy_factor × the mean occupancy (see run_pass); 0 = automatic
    int batch_affine = -1; // -1 automatic, 0 never, 1..3 = pairwise batched-affine rounds before the XYZZ accumulation
    bool graphs = !(getenv("B200MSM_GRAPHS") && atoi(getenv("B200MSM_GRAPHS")) == 0);   // record a pass seen twice as CUDA graphs and replay it (env: debugging)
    unsigned epoch = 0;    // bumped by every setter that changes what a pass allocates (fit caches are keyed on it)
};
struct Share;
// One persistent host thread per bound device (only when more than one is bound); see msm_host.
struct DeviceWorker {
    DeviceCtx *cx = nullptr;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    const Share *job = nullptr;   // set by the caller, cleared by the worker
    bool quit = false, done = false;
    int rc = 0;
    std::string err;
    void *gather_dst = nullptr;   // device 0's slot for this device's partial
    int gather_dev = 0;
    void loop();
};
struct Engine {
    std::mutex mu;
    bool inited = false;
    int first = 0;                    // bound device range [first, first + ctx.size())
    unsigned generation = 0;          // bumped by every (re-)initialisation; resident handles remember theirs
    std::vector<std::unique_ptr<DeviceCtx>> ctx;
    // extra lanes (b200msm_set_lane): further contexts on the same devices with their own scratch, so that MSMs
    // issued on different streams overlap — one's latency-bound tail under another's accumulation
    std::vector<std::unique_ptr<DeviceCtx>> lane_ctx;
    // one persistent host thread per bound device when more than one is bound: a sharded host-buffer MSM is
    // issued on all devices at once instead of device after device from the calling thread
    std::vector<std::unique_ptr<DeviceWorker>> workers;
    std::mutex multi_mu;              // one sharded (multi-device) call at a time: they share the gather slots
    std::mutex tun_mu;
    Tun tun;
    // A process that bound several devices and exits without b200msm_shutdown must not die in ~std::thread of a
    // worker that is still waiting for work: stop and join them (they touch no CUDA state on the way out).
    ~Engine() { stop_workers(); }
    void stop_workers() {
        for (auto &w : workers) {
            {
                std::lock_guard<std::mutex> wl(w->mu);
                w->quit = true;
                w->cv.notify_all();
            }
            if (w->th.joinable()) w->th.join();
        }
        workers.clear();
    }
};
Engine g_eng;
Tun tun_snapshot() {
    std::lock_guard<std::mutex> lk(g_eng.tun_mu);
    return g_eng.tun;
}
template <class Fn> void tun_update(Fn fn, bool replans = true) {
    std::lock_guard<std::mutex> lk(g_eng.tun_mu);
    fn(g_eng.tun);
    if (replans) g_eng.tun.epoch++;
}

int make_ctx(int d, int lane, int sm_count, std::unique_ptr<DeviceCtx> &out) {
    auto c = std::make_unique<DeviceCtx>();
    c->dev = d;
    c->lane = lane;
    c->sm_count = sm_count;
    CUDA_TRY(cudaSetDevice(d));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_scalars, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_bases, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_busy, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_share, cudaEventDisableTiming));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking));
    int lo_pri = 0, hi_pri = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
    CUDA_TRY(cudaStreamCreateWithPriority(&c->tail_stream, cudaStreamNonBlocking, hi_pri));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_tail_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_tail_join, cudaEventDisableTiming));
    for (auto &ev : c->ev) CUDA_TRY(cudaEventCreate(&ev));
    for (auto &ev : c->ev_slice) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    out = std::move(c);
    return 0;
}
void destroy_ctx(DeviceCtx &c) {
    std::lock_guard<std::mutex> l2(c.mu);
    cudaSetDevice(c.dev);
    cudaStreamSynchronize(c.stream);
    cudaStreamSynchronize(c.tail_stream);
    for (auto &g : c.graphs) { cudaGraphExecDestroy(g.prep); cudaGraphExecDestroy(g.main); }
    c.graphs.clear();
    c.seen.clear();
    cudaStreamDestroy(c.cap_stream);
    c.release_all();
    for (auto &ev : c.ev) cudaEventDestroy(ev);
    for (auto &ev : c.ev_slice) cudaEventDestroy(ev);
    cudaStreamDestroy(c.stream);
    cudaStreamDestroy(c.copy_stream);
    cudaStreamDestroy(c.aux_stream);
    cudaStreamDestroy(c.tail_stream);
    for (cudaEvent_t e : {c.ev_scalars, c.ev_bases, c.ev_busy, c.ev_fork, c.ev_join, c.ev_share, c.ev_tail_fork, c.ev_tail_join}) cudaEventDestroy(e);
}

int engine_init_locked(int first, int ndev) {
    if (g_eng.inited) {
        // lazy initialisation (first < 0: "whatever is bound") never re-binds; an explicit request for a
        // different range is an error rather than a silent no-op
        if (first < 0) return 0;
        int count = 0;
        cudaGetDeviceCount(&count);
        const int want = ndev <= 0 ? count - first : ndev;
        if (first == g_eng.first && want == (int)g_eng.ctx.size()) return 0;
        return fail(B200MSM_EINVAL, "engine already bound to devices [" + std::to_string(g_eng.first) + ", " +
                                        std::to_string(g_eng.first + (int)g_eng.ctx.size()) + "): call b200msm_shutdown before binding another range");
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(B200MSM_ENODEV, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count=0"));
    if (first < 0) {  // bind to whatever device is current (torch's, under torchrun)
        if (cudaGetDevice(&first) != cudaSuccess) first = 0;
        ndev = 1;
    }
    if (ndev <= 0) ndev = count - first;
    if (first >= count || first + ndev > count) return fail(B200MSM_EINVAL, "device range out of bounds");
    int prev = 0;
    cudaGetDevice(&prev);
    auto undo = [&](int rc) {
        for (auto &c : g_eng.ctx) destroy_ctx(*c);
        g_eng.ctx.clear();
        cudaSetDevice(prev);
        return rc;
    };
    for (int d = first; d < first + ndev; d++) {
        cudaDeviceProp p;
        cudaError_t pe = cudaGetDeviceProperties(&p, d);
        if (pe != cudaSuccess) return undo(fail(B200MSM_ECUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(pe)));
        if (p.major != 10)
            return undo(fail(B200MSM_ENODEV, std::string("device ") + std::to_string(d) + " (" + p.name + ") is sm_" +
                                                 std::to_string(p.major * 10 + p.minor) + "; this build is sm_100a only"));
        std::unique_ptr<DeviceCtx> c;
        if (int rc = make_ctx(d, 0, p.multiProcessorCount, c)) return undo(rc);
        g_eng.ctx.push_back(std::move(c));
    }
    if (ndev > 1) {
        // partials travel to the first device by peer copy: enable direct access where the topology has it
        // (NVLink / NVSwitch on an 8×B200 box); without it the copy is staged through the host and still correct
        for (int d = 1; d < ndev; d++) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, first + d, first) == cudaSuccess && can) {
                cudaSetDevice(first + d);
                cudaError_t pe = cudaDeviceEnablePeerAccess(first, 0);
                if (pe != cudaSuccess) cudaGetLastError();  // already enabled (by torch, say) is fine
            }
        }
        for (int d = 0; d < ndev; d++) {
            auto w = std::make_unique<DeviceWorker>();
            w->cx = g_eng.ctx[d].get();
            DeviceWorker *wp = w.get();
            w->th = std::thread([wp] { wp->loop(); });
            g_eng.workers.push_back(std::move(w));
        }
    }
    cudaSetDevice(prev);
    g_eng.first = first;
    g_eng.generation++;
    g_eng.inited = true;
    return 0;
}
int engine_init(int first, int ndev) {
    std::lock_guard<std::mutex> lk(g_eng.mu);
    return engine_init_locked(first, ndev);
}
thread_local int t_lane = 0;
// context of the calling thread's current device and lane (engine lock held while the lists are walked)
DeviceCtx *ctx_for_current_device() {
    int d = 0;
    cudaGetDevice(&d);
    std::lock_guard<std::mutex> lk(g_eng.mu);
    DeviceCtx *base = nullptr;
    for (auto &c : g_eng.ctx)
        if (c->dev == d) base = c.get();
    if (!base || t_lane == 0) return base;
    for (auto &c : g_eng.lane_ctx)
        if (c->dev == d && c->lane == t_lane) return c.get();
    std::unique_ptr<DeviceCtx> c;
    if (make_ctx(d, t_lane, base->sm_count, c)) return nullptr;
    g_eng.lane_ctx.push_back(std::move(c));
    return g_eng.lane_ctx.back().get();
}

// The pipeline on one device. Inputs already in device memory; d_out receives 3 field elements.
// Must be called with ctx.mu held and ctx.dev current. Asynchronous on `st`.
// `bases_ready` (optional): an event after which d_bases is valid — the scalar-side phases
// (digits, sort, bucket offsets) do not read the bases, so they overlap the bases' H2D copy.
// Streamed MSM (host bases arriving in slices): every slice is grouped and accumulated into the
// SAME buckets under one plan — `plan` fixed by the caller, `into` for every slice but the first —
// and only the last one (`finish`) runs the reduction and the combination.
struct PassOpts {
    const Plan *plan = nullptr;
    bool into = false, finish = true;
};
// (environment switch, read once: the rounds as ONE pipeline instead of two interleaved half-range pipelines)
bool ba_split_ok() {
    static const bool ok = !(getenv("B200MSM_BA_SPLIT") && atoi(getenv("B200MSM_BA_SPLIT")) == 0);
    return ok;
}

// Plan + scratch reservation of one pass (may allocate: never inside a graph capture).
int pass_prepare(int group, DeviceCtx &cx, const Tun &tn, const void *d_bases_v, const void *d_scalars_v, size_t n, int mont,
                 void *d_out_v, const TableRef *tbl, const PassOpts &po, PassArgs &pa) {
    const bool g2 = group == B200MSM_G2;
    const int W = g2 ? 24 : 12;                      // u32 words per field element
    const size_t PB = 4 * (size_t)W * sizeof(uint32_t);  // bytes per XYZZ point
    memset(&pa, 0, sizeof pa);                       // (padding bytes too: the struct is compared bytewise)
    pa.group = group; pa.mont = mont; pa.n = n;
    pa.d_bases = (const uint32_t *)d_bases_v; pa.d_scalars = (const uint32_t *)d_scalars_v; pa.d_out = (uint32_t *)d_out_v;
    pa.into = po.into; pa.finish = po.finish;
    Plan pl;
    if (po.plan) pl = *po.plan;
    else if (tbl) {  // the table fixes the width
        pl.c = tbl->c;
        pl.nwin = tbl->nwin;
        pl.glv = pl.split = false;
        pl.parts = 1;
        pl.nbw = 1u << (pl.c - 1);
        pl.nb = pl.nbw;                       // one bucket set shared by all windows
    } else auto_plan(n, g2, tn.glv_mode, tn.window_override, pl, tn.batch_affine);
    if (tbl) { pa.d_bases = (const uint32_t *)tbl->p; pa.tbl_p = tbl->p; pa.tbl_stride = tbl->stride; }
    pa.c = pl.c; pa.nwin = pl.nwin; pa.glv = pl.glv ? (pl.parts == 4 ? 3 : 1) + (pl.split ? 1 : 0) : 0; pa.nbw = pl.nbw; pa.nb = pl.nb;
    const int rwin = tbl ? 1 : pl.nwin;       // windows the reduction sees
    cx.last_plan[0] = pl.c; cx.last_plan[1] = pl.nwin; cx.last_plan[2] = pa.glv; cx.last_plan[3] = tbl ? 1 : 0;
    const size_t entries = (size_t)pl.parts * n;  // per window
    const size_t m = entries * (size_t)pl.nwin;
    // heavy buckets: one thread per bucket is the efficient shape (≈0.31 product-times per entry
    // per warp vs ≈1 for the block-cooperative path), so a bucket only counts as heavy when its
    // serial chain would be a visible fraction (≈8 %) of the whole accumulation — the kernel lasts
    // ≈ m/175k bucket-entry times — or when it exceeds 3× the mean occupancy, whichever is larger.
    // Buckets are taken in decreasing-size order, so the long chains start first.
    // Worst-case list sizes follow from Σ counts = m.
    // (table mode: the buckets the top window also feeds hold up to ≈2.3× the mean, so the factor is 4 there)
    const uint32_t avg = (uint32_t)(((tbl ? m : entries) + pl.nbw - 1) / pl.nbw);
    // batched-affine pairing rounds: they pay when a bucket holds enough entries for the pairs to be real additions
    // (padding to 2^R entries per bucket is wasted slots) — measured crossover, see profiles/r02_experiments.md
    int R = tn.batch_affine;
    if (R < 0) R = ba_rounds_for((double)avg);
    if (n == 0) R = 0;
    pa.R = R;
    const int hfac = tn.heavy_factor ? tn.heavy_factor : (tbl ? 4 : 3);
    // after R rounds a bucket's run is 2^R times shorter and so is the whole accumulation: the threshold scales with both
    pa.heavy_thr = std::max<uint32_t>(std::max<uint32_t>(32 >> R, ((uint32_t)hfac * avg) >> R), (uint32_t)((m >> R) / 175000));
    if (pa.heavy_thr < 4) pa.heavy_thr = 4;
    const size_t slots_max = m + (size_t)pl.nb * ((1u << R) - 1);   // padded entry slots (upper bound)

    if (int rc = cx.digits.reserve(m * 4)) return rc;
    if (int rc = cx.vals.reserve(slots_max * 4)) return rc;
    if (int rc = cx.cnt.reserve((size_t)pl.nb * 4)) return rc;
    if (int rc = cx.ord.reserve((size_t)pl.nb * 4)) return rc;
    for (int i = 0; i < 2; i++) {
        size_t lvl = (size_t)rwin * std::max<size_t>(1, pl.nbw / 32) * PB;
        if (int rc = cx.lvlR[i].reserve(lvl)) return rc;
        if (int rc = cx.lvlC[i].reserve(lvl)) return rc;
    }
    if (int rc = cx.start.reserve(((size_t)pl.nb + 2) * 4)) return rc;
    if (int rc = cx.buckets.reserve((size_t)pl.nb * PB)) return rc;
    const size_t m_acc = slots_max >> R;      // entries the XYZZ accumulation (and its heavy path) sees
    const size_t max_heavy = m_acc / (pa.heavy_thr + 1) + 1, max_tasks = m_acc / HEAVY_CHUNK + max_heavy + 1;
    if (int rc = cx.hvy_hdr.reserve(16)) return rc;
    if (int rc = cx.hvy_buckets.reserve(max_heavy * 12)) return rc;
    if (int rc = cx.hvy_tasks.reserve(max_tasks * 8)) return rc;
    if (int rc = cx.hvy_partials.reserve(max_tasks * PB)) return rc;
    if (int rc = cx.tile_sums.reserve(((size_t)pl.nb / 2048 + 4) * 4)) return rc;
    if (int rc = cx.size_hist.reserve(2 * 4096 * 4)) return rc;
    if (pl.glv)   // β·x table (two parts), or the three full-point image tables −ψ, ψ², −ψ³ (four parts, G2)
        if (int rc = cx.endo.reserve(n * (size_t)W * 4 * (pl.parts == 4 ? 6 : 1))) return rc;
    if (R > 0) {
        const size_t FB = (size_t)W * 4, s1 = (slots_max + 1) / 2;
        const BaLayout bl = ba_layout(s1, R, cx.sm_count, ba_split_ok());
        const size_t np = bl.split ? 2 : 1;
        if (int rc = cx.ba_prefix.reserve((np * bl.pre_el + 256) * FB)) return rc;
        if (int rc = cx.ba_T.reserve((np * bl.t_el + 256) * FB)) return rc;
        if (int rc = cx.ba_prefix2.reserve((np * bl.t_el + 256) * FB)) return rc;
        if (int rc = cx.ba_U.reserve((np * bl.u_el + 256) * FB)) return rc;
        // one output array per round (NOT a ping-pong pair: with the rounds running as two half-range pipelines on two
        // streams, the pipeline that is a round ahead would overwrite what the other has not read yet)
        size_t sr = s1;
        for (int r = 0; r < R; r++) {
            if (int rc = cx.ba_pts[r].reserve(sr * 2 * FB)) return rc;
            sr = (sr + 1) / 2;
        }
    }
    if (pa.finish) {
        uint32_t len = pl.nbw;
        while (len > 2048 && (!tbl || len / 32 >= 4096)) len /= 32;
        const size_t tstride = std::max<uint32_t>(1, len / 2);
        for (int i = 0; i < 2; i++) {
            if (int rc = cx.treeS[i].reserve((size_t)rwin * tstride * PB)) return rc;
            if (int rc = cx.treeV[i].reserve((size_t)rwin * tstride * PB)) return rc;
            if (int rc = cx.treeC[i].reserve((size_t)rwin * tstride * PB)) return rc;
        }
        if (int rc = cx.wsum.reserve(((size_t)rwin + 1) * PB)) return rc;   // (+1: k_combine parks the split top digit's term there)
    }
    pa.arena = cx.arena_epoch();
    return 0;
}

// Stage 1 of a pass — the scalar side: canonical scalars → window digits, per-bucket histogram, scan → bucket
// offsets, scatter of the point indices (a counting sort on the bucket id; zero digits are dropped), buckets in
// decreasing-size order.  Does not read the bases, so it runs while they are still crossing PCIe.
int pass_issue_prep(DeviceCtx &cx, const PassArgs &pa, cudaStream_t st, bool prof, int &evi) {
    auto mark = [&]() { if (prof) cudaEventRecord(cx.ev[evi++], st); };
    mark();
    launch_group_by_bucket(pa.d_scalars, pa.n, pa.mont, pa.glv, pa.c, pa.nwin, pa.nb, cx.digits.as<uint32_t>(), cx.cnt.as<uint32_t>(),
                           cx.start.as<uint32_t>(), cx.tile_sums.as<uint32_t>(), cx.vals.as<uint32_t>(), st, pa.tbl_stride, pa.R);
    mark();
    mark();
    launch_order_by_size(cx.start.as<uint32_t>(), pa.nb, cx.size_hist.as<uint32_t>(), cx.ord.as<uint32_t>(), st);
    mark();
    CUDA_TRY(cudaMemsetAsync(cx.hvy_hdr.p, 0, 16, st));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Stage 2 — everything that touches points: (GLV) β·x table, batched-affine pairing rounds, XYZZ bucket
// accumulation with the heavy buckets on a side stream, and — when `finish` — bucket reduction and window combination.
int pass_issue_main(DeviceCtx &cx, const PassArgs &pa, cudaStream_t st, bool prof, int &evi) {
    const bool g2 = pa.group == B200MSM_G2;
    const int W = g2 ? 24 : 12;
    const bool tbl = pa.tbl_p != nullptr;
    const int rwin = tbl ? 1 : pa.nwin;
    auto mark = [&]() { if (prof) cudaEventRecord(cx.ev[evi++], st); };
    uint32_t *vals = cx.vals.as<uint32_t>(), *ord = cx.ord.as<uint32_t>(), *start = cx.start.as<uint32_t>();
    const uint32_t *endo_x = nullptr;
    uint32_t n_pts = 0xffffffffu;
    const int parts = pa.glv == 0 ? 1 : (pa.glv <= 2 ? 2 : 4), img_full = parts == 4;
    if (pa.glv) {  // β·x table for the endomorphism images (one product per base), or the three ψ image tables
        if (img_full) launch_psi_tables_g2(pa.d_bases, pa.n, cx.endo.as<uint32_t>(), st);
        else (g2 ? launch_endo_table_g2 : launch_endo_table_g1)(pa.d_bases, pa.n, cx.endo.as<uint32_t>(), st);
        endo_x = cx.endo.as<uint32_t>();
        n_pts = (uint32_t)pa.n;
    }
    // 4a. batched-affine pairing rounds (batch_affine.cuh): the (padded) entry array is halved R times by affine
    //     additions that share one inversion per ≈10^5 pairs; what is left is accumulated in XYZZ form below
    const uint32_t *acc_pts = pa.d_bases, *acc_vals = vals;
    if (pa.R > 0) {
        const size_t entries = (size_t)parts * pa.n, m = entries * (size_t)pa.nwin;
        size_t s_out = (m + (size_t)pa.nb * ((1u << pa.R) - 1) + 1) / 2;
        const uint32_t *src = pa.d_bases;
        // Large rounds run as TWO interleaved pipelines over the two halves of the slot range, on two streams: the
        // second level + inversion of a round are three small latency-bound launches (≈0.1–0.2 ms in all) during
        // which a single pipeline leaves the GPU almost idle; with two, the other half's large kernels fill it.
        // Each pipeline owns a fixed part of every scratch array for all rounds (ba_layout).
        const BaLayout bl = ba_layout(s_out, pa.R, cx.sm_count, ba_split_ok());
        const bool split = bl.split;
        const int align_log = split ? 6 + pa.R : 0;
        const size_t FW = (size_t)W;   // u32 words per field element
        if (split) {
            CUDA_TRY(cudaEventRecord(cx.ev_fork, st));
            CUDA_TRY(cudaStreamWaitEvent(cx.aux_stream, cx.ev_fork, 0));
        }
        for (int r = 0; r < pa.R; r++) {
            uint32_t *out = cx.ba_pts[r].as<uint32_t>();
            for (int part = 0; part < (split ? 2 : 1); part++) {
                (g2 ? launch_ba_round_g2 : launch_ba_round_g1)(r == 0, src, vals, endo_x, n_pts, start + pa.nb, r, bl.bp[r],
                                                               cx.ba_prefix.as<uint32_t>() + (part ? bl.pre_el : 0) * FW,
                                                               cx.ba_T.as<uint32_t>() + (part ? bl.t_el : 0) * FW,
                                                               cx.ba_prefix2.as<uint32_t>() + (part ? bl.t_el : 0) * FW,
                                                               cx.ba_U.as<uint32_t>() + (part ? bl.u_el : 0) * FW, out,
                                                               part ? cx.aux_stream : st, part, align_log, img_full);
            }
            src = out;
        }
        if (split) {
            CUDA_TRY(cudaEventRecord(cx.ev_join, cx.aux_stream));
            CUDA_TRY(cudaStreamWaitEvent(st, cx.ev_join, 0));
        }
        acc_pts = src;
        acc_vals = nullptr;
    }
    // 4b. heavy buckets run on a side stream next to the light kernel (disjoint outputs): each fills
    //     the SMs the other leaves idle at its tail
    CUDA_TRY(cudaEventRecord(cx.ev_fork, st));
    CUDA_TRY(cudaStreamWaitEvent(cx.aux_stream, cx.ev_fork, 0));
    (g2 ? launch_heavy_g2 : launch_heavy_g1)(acc_pts, acc_vals, start, ord, pa.nb, pa.heavy_thr, endo_x, n_pts, cx.hvy_hdr.p,
                                             cx.hvy_buckets.p, cx.hvy_tasks.p, cx.hvy_partials.as<uint32_t>(), pa.into,
                                             cx.buckets.as<uint32_t>(), cx.sm_count * 4, cx.aux_stream, pa.R, img_full);
    CUDA_TRY(cudaEventRecord(cx.ev_join, cx.aux_stream));
    if (pa.R > 0)
        (g2 ? launch_accumulate_direct_g2 : launch_accumulate_direct_g1)(acc_pts, start, ord, pa.nb, pa.heavy_thr, pa.R, pa.into,
                                                                         cx.buckets.as<uint32_t>(), st);
    else
        (g2 ? launch_accumulate_g2 : launch_accumulate_g1)(pa.d_bases, vals, start, ord, pa.nb, pa.heavy_thr, endo_x, n_pts, pa.into,
                                                           cx.buckets.as<uint32_t>(), st, img_full);
    CUDA_TRY(cudaStreamWaitEvent(st, cx.ev_join, 0));
    mark();
    if (!pa.finish) {
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    // 5. per-window weighted bucket sums: running-sum levels of fan-in 32 while the arrays are long
    //    (throughput-bound), then a log-depth tree (latency-bound part).  From here on the pass is a
    //    chain of small dependent launches: it moves to the context's high-priority stream, so that
    //    when another lane's accumulation fills the GPU these blocks take the next free slots.
    const cudaStream_t caller_st = st;
    if (!prof) {
        CUDA_TRY(cudaEventRecord(cx.ev_tail_fork, st));
        CUDA_TRY(cudaStreamWaitEvent(cx.tail_stream, cx.ev_tail_fork, 0));
        st = cx.tail_stream;
    }
    const uint32_t *X = cx.buckets.as<uint32_t>();
    const uint32_t *Cin = nullptr;
    uint32_t len = pa.nbw;
    int log2M = 0, pp = 0;
    // (a single window — table mode — goes to the tree as soon as a level would leave fewer than 4096 chains)
    constexpr int fan_log = 5;                // fan-in 32: 16 and 8 measured equal or slower (profiles/r01_experiments.md)
    const uint32_t fan = 1u << fan_log;
    while (len > 2048 && (!tbl || len / fan >= 4096)) {
        (g2 ? launch_wsum_level_g2 : launch_wsum_level_g1)(X, Cin, len, fan, log2M, (uint32_t)rwin,
                                                           cx.lvlR[pp].as<uint32_t>(), cx.lvlC[pp].as<uint32_t>(), st);
        X = cx.lvlR[pp].as<uint32_t>();
        Cin = cx.lvlC[pp].as<uint32_t>();
        log2M += fan_log;
        len /= fan;
        pp ^= 1;
    }
    const uint32_t S = len;
    int logS = 0;
    while ((1u << logS) < S) logS++;
    const size_t tstride = std::max<uint32_t>(1, S / 2);  // points per window in every tree array
    const uint32_t *Sin = X, *Ccur = Cin;
    size_t sin_stride = S, cin_stride = S;
    int cur = 0;
    for (int j = 0; j < logS; j++) {
        (g2 ? launch_tree_level_g2 : launch_tree_level_g1)(Sin, sin_stride, cx.treeV[cur].as<uint32_t>(), Ccur, cin_stride,
                                                           cx.treeS[cur ^ 1].as<uint32_t>(), cx.treeV[cur ^ 1].as<uint32_t>(),
                                                           cx.treeC[cur ^ 1].as<uint32_t>(), tstride, S, j, (uint32_t)rwin, st);
        cur ^= 1;
        Sin = cx.treeS[cur].as<uint32_t>();
        sin_stride = tstride;
        if (Cin) { Ccur = cx.treeC[cur].as<uint32_t>(); cin_stride = tstride; }
    }
    mark();
    // 6. window values and Horner over the windows → one Jacobian point
    (g2 ? launch_combine_g2 : launch_combine_g1)(Sin, cx.treeV[cur].as<uint32_t>(), Cin ? Ccur : nullptr, tstride, logS, log2M,
                                                 rwin, pa.c, (pa.glv == 2 || pa.glv == 4) ? 1 : 0, cx.wsum.as<uint32_t>(), pa.d_out, st);
    mark();
    if (st != caller_st) {
        CUDA_TRY(cudaEventRecord(cx.ev_tail_join, st));
        CUDA_TRY(cudaStreamWaitEvent(caller_st, cx.ev_tail_join, 0));
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// A pass seen for the second time with the same arguments (a prover repeats its sizes; every slice
// of a streamed MSM; every step of a benchmark) is captured into two CUDA graphs — scalar side,
// point side — and replayed from then on: two launches instead of ≈45, which is what keeps eight
// host threads driving eight GPUs from one process off the driver's locks.  The wait for the bases'
// upload stays between the two as an ordinary stream wait.
int capture_stage(DeviceCtx &cx, const PassArgs &pa, cudaStream_t st, bool main_stage, cudaGraphExec_t *exec, unsigned long long *launches) {
    (void)st;
    const unsigned long long l0 = t_own_launches;
    CUDA_TRY(cudaStreamBeginCapture(cx.cap_stream, cudaStreamCaptureModeThreadLocal));
    int evi = 0;
    int rc = main_stage ? pass_issue_main(cx, pa, cx.cap_stream, false, evi) : pass_issue_prep(cx, pa, cx.cap_stream, false, evi);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(cx.cap_stream, &g);
    *launches = t_own_launches - l0;
    __atomic_fetch_sub(&g_own_launches, *launches, __ATOMIC_RELAXED);    // counted when the graph is launched, not when it is recorded
    if (rc) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        return rc;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B200MSM_ECUDA, std::string("graph capture: ") + cudaGetErrorString(e));
    }
    e = cudaGraphInstantiate(exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B200MSM_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    }
    return 0;
}

// The pipeline on one device. Inputs already in device memory; d_out receives 3 field elements.
// Must be called with ctx.mu held and ctx.dev current. Asynchronous on `st`.
// `bases_ready` (optional): an event after which d_bases is valid — the scalar-side phases
// (digits, sort, bucket offsets) do not read the bases, so they overlap the bases' H2D copy.
int run_pass(int group, DeviceCtx &cx, const Tun &tn, const void *d_bases_v, const void *d_scalars_v, size_t n, int mont, void *d_out_v,
              cudaStream_t st, cudaEvent_t bases_ready = nullptr, const TableRef *tbl = nullptr, const PassOpts &po = PassOpts()) {
    if (group != B200MSM_G1 && group != B200MSM_G2) return fail(B200MSM_EINVAL, "group must be B200MSM_G1 or B200MSM_G2");
    const int W = group == B200MSM_G2 ? 24 : 12;
    if (n == 0 && !po.plan) {
        CUDA_TRY(cudaMemsetAsync(d_out_v, 0, 3 * W * 4, st));
        return 0;
    }
    PassArgs pa;
    if (int rc = pass_prepare(group, cx, tn, d_bases_v, d_scalars_v, n, mont, d_out_v, tbl, po, pa)) return rc;
    const bool prof = tn.profiling;
    int evi = 0;
    PassGraph *pg = nullptr;
    if (tn.graphs && !prof) {
        for (auto &g : cx.graphs)
            if (same_pass(g.key, pa)) { pg = &g; break; }
        if (!pg) {
            // first sight: remember the arguments and run directly; second sight: record
            bool seen = false;
            for (auto &k : cx.seen)
                if (same_pass(k, pa)) { seen = true; break; }
            if (!seen) {
                if (cx.seen.size() >= 64) cx.seen.erase(cx.seen.begin());
                cx.seen.push_back(pa);
            } else {
                if (cx.graphs.size() >= 48) {           // evict the least recently used
                    size_t lru = 0;
                    for (size_t i = 1; i < cx.graphs.size(); i++)
                        if (cx.graphs[i].last_use < cx.graphs[lru].last_use) lru = i;
                    cudaGraphExecDestroy(cx.graphs[lru].prep);
                    cudaGraphExecDestroy(cx.graphs[lru].main);
                    cx.graphs.erase(cx.graphs.begin() + lru);
                }
                PassGraph ng;
                ng.key = pa;
                memcpy(ng.plan, cx.last_plan, sizeof ng.plan);
                if (capture_stage(cx, pa, st, false, &ng.prep, &ng.launches_prep) == 0) {
                    if (capture_stage(cx, pa, st, true, &ng.main, &ng.launches_main) == 0) {
                        cx.graphs.push_back(ng);
                        pg = &cx.graphs.back();
                    } else cudaGraphExecDestroy(ng.prep);
                }
                // (a failed capture falls back to direct issue below; the error text stays in last_error)
            }
        }
    }
    if (pg) {
        pg->last_use = ++cx.graph_clock;
        CUDA_TRY(cudaGraphLaunch(pg->prep, st));
        if (bases_ready) CUDA_TRY(cudaStreamWaitEvent(st, bases_ready, 0));
        CUDA_TRY(cudaGraphLaunch(pg->main, st));
        __atomic_fetch_add(&g_own_launches, pg->launches_prep + pg->launches_main, __ATOMIC_RELAXED);
        return 0;
    }
    if (int rc = pass_issue_prep(cx, pa, st, prof, evi)) return rc;
    if (bases_ready) CUDA_TRY(cudaStreamWaitEvent(st, bases_ready, 0));
    if (int rc = pass_issue_main(cx, pa, st, prof, evi)) return rc;
    cx.phase_pending = prof && pa.finish;  // elapsed times are read lazily by b200msm_last_phase_ms (no sync here)
    return 0;
}

// Scratch bytes one pass over n points needs (sort double buffers dominate).
size_t pass_scratch_bytes(size_t n, bool g2, int c_override, bool table = false) {
    int c = c_override > 0 ? c_override : auto_window(n, g2);
    c = std::max(2, std::min(c, 22));
    size_t nwin = (256 + c - 1) / c, nb = nwin << (c - 1), m = n * nwin;  // the GLV plan needs about the same
    size_t PB = g2 ? 384 : 192;
    const size_t ba = m * (PB * 9 / 16) + nb * 16;  // batched-affine rounds: prefix products (m/2 elements) + one output array per round (m/2 + m/4 + m/8 points)
    if (table) return m * 8 + (nb / nwin) * (PB + PB / 8 + 20) + m / 96 * (PB + 20) + ba + (64u << 20);
    return m * 8 + nb * (PB + PB / 8 + 20) + m / 96 * (PB + 20) + n * (PB / 4) + ba + (64u << 20);
}

// The whole MSM on one device: one pass when it fits, otherwise the chunking the reference left
// as a TODO (src/gpu.rs:238-239; its calc_chunk_size result is never used, :64-85,114): the points
// are cut into equal chunks whose sort arrays stay below 2^32 entries and within the free HBM,
// each chunk yields a Jacobian partial, and the partials are added on the device.
int run_group(int group, DeviceCtx &cx, const Tun &tn, const void *d_bases, const void *d_scalars, size_t n, int mont, void *d_out,
              cudaStream_t st, cudaEvent_t bases_ready = nullptr, const TableRef *tbl = nullptr) {
    if (group != B200MSM_G1 && group != B200MSM_G2) return fail(B200MSM_EINVAL, "group must be B200MSM_G1 or B200MSM_G2");
    const bool g2 = group == B200MSM_G2;
    // GPU-side serialisation of the context's scratch arena across streams
    if (cx.busy_valid) CUDA_TRY(cudaStreamWaitEvent(st, cx.ev_busy, 0));
    // cudaMemGetInfo is a slow synchronous driver call (≈0.3–1 ms): only ask for sizes that have
    // not already run as a single pass on this context (the grow-only arena then fits)
    size_t budget = ~(size_t)0;
    const int gi = g2 ? 1 : 0;
    if (cx.fits_epoch != tn.epoch) {  // a setter changed what a pass of a given size allocates: forget what is known to fit
        cx.fits_epoch = tn.epoch;
        cx.fits_n[0] = cx.fits_n[1] = cx.fits_tbl_n[0] = cx.fits_tbl_n[1] = 0;
    }
    const bool known_fit = tbl ? (tbl->c == cx.fits_tbl_c[gi] && n <= cx.fits_tbl_n[gi]) : n <= cx.fits_n[gi];
    if (!tn.max_chunk_override && !known_fit) {
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        budget = (size_t)((double)(free_b + cx.scratch_bytes()) * 0.85);
    }
    size_t chunks = 1;
    auto too_big = [&](size_t cn) {
        if (tn.max_chunk_override) return cn > tn.max_chunk_override;
        // (positions in the sorted entry array are 32-bit; the batched-affine padding adds up to 7 slots per bucket)
        if (tbl) return cn * tbl->nwin + ((size_t)7 << (tbl->c - 1)) >= 0xfff00000ull || pass_scratch_bytes(cn, g2, tbl->c, true) > budget;
        Plan p;
        auto_plan(cn, g2, tn.glv_mode, tn.window_override, p, tn.batch_affine);
        size_t ent = (size_t)p.parts * cn;
        return ent * p.nwin + (size_t)p.nb * 7 >= 0xfff00000ull || cn >= (1ull << 30) || pass_scratch_bytes(cn, g2, tn.window_override) > budget;
    };
    while (too_big((n + chunks - 1) / chunks)) {
        chunks *= 2;
        if (chunks > (1u << 20)) return fail(B200MSM_ENOMEM, "cannot fit even a tiny chunk of this MSM in device memory");
    }
    int rc = 0;
    if (chunks == 1) {
        rc = run_pass(group, cx, tn, d_bases, d_scalars, n, mont, d_out, st, bases_ready, tbl);
        if (!rc && !tn.max_chunk_override) {
            if (tbl) {
                if (tbl->c != cx.fits_tbl_c[gi]) { cx.fits_tbl_c[gi] = tbl->c; cx.fits_tbl_n[gi] = 0; }
                cx.fits_tbl_n[gi] = std::max(cx.fits_tbl_n[gi], n);
            } else cx.fits_n[gi] = std::max(cx.fits_n[gi], n);
        }
    } else {
        const size_t AB = g2 ? 192 : 96, JB = g2 ? 288 : 144;
        if ((rc = cx.chunk_partials.reserve(chunks * JB))) return rc;
        for (size_t k = 0; k < chunks && !rc; k++) {
            size_t lo = n * k / chunks, hi = n * (k + 1) / chunks;
            TableRef sub;
            if (tbl) { sub = *tbl; sub.p = (const char *)tbl->p + lo * AB; }  // same stride, shifted origin
            rc = run_pass(group, cx, tn, (const char *)d_bases + lo * AB, (const char *)d_scalars + lo * 32, hi - lo, mont,
                          (char *)cx.chunk_partials.p + k * JB, st, k == 0 ? bases_ready : nullptr, tbl ? &sub : nullptr);
        }
        if (!rc) (g2 ? launch_sum_partials_g2 : launch_sum_partials_g1)((const uint32_t *)cx.chunk_partials.p, (int)chunks, (uint32_t *)d_out, st);
    }
    if (!rc) {
        CUDA_TRY(cudaEventRecord(cx.ev_busy, st));
        cx.busy_valid = true;
    }
    return rc;
}

size_t aff_bytes(int group) { return group == B200MSM_G2 ? 192 : 96; }
size_t jac_bytes(int group) { return group == B200MSM_G2 ? 288 : 144; }

// Windows 1..nwin−1 of a fixed-base table whose window 0 (the bases themselves) is already in
// place: table[w] = 2^c·table[w−1], c doublings per point then a batch normalisation back to
// affine, in slices that bound the Jacobian scratch.  ctx.mu held, ctx.dev current, async on st.
int build_table(int group, DeviceCtx &cx, void *d_table, size_t n, size_t stride, int c, int nwin, cudaStream_t st) {
    const bool g2 = group == B200MSM_G2;
    const size_t AB = aff_bytes(group), JB = jac_bytes(group);
    const size_t slice = std::min<size_t>(n, (size_t)1 << 22);
    if (cx.busy_valid) CUDA_TRY(cudaStreamWaitEvent(st, cx.ev_busy, 0));
    if (int rc = cx.tbl_tmp.reserve(slice * JB)) return rc;
    for (int w = 1; w < nwin; w++)
        for (size_t lo = 0; lo < n; lo += slice) {
            const size_t cn = std::min(slice, n - lo);
            const uint32_t *prev = (const uint32_t *)((const char *)d_table + ((size_t)(w - 1) * stride + lo) * AB);
            uint32_t *next = (uint32_t *)((char *)d_table + ((size_t)w * stride + lo) * AB);
            (g2 ? launch_table_shift_g2 : launch_table_shift_g1)(prev, cn, c, cx.tbl_tmp.as<uint32_t>(), st);
            launch_normalize_batch(g2, cx.tbl_tmp.as<uint32_t>(), cn, next, cx.sm_count, st);
        }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(cx.ev_busy, st));
    cx.busy_valid = true;
    return 0;
}

// host-buffer MSM over the bound devices: shard by index range, one partial per device, final
// addition on the first device.
int msm_host(int group, const uint64_t *bases, const uint64_t *scalars, size_t n, int mont, uint64_t *out,
             const struct b200msm_bases *resident);

}  // namespace

struct b200msm_bases {
    int group;
    size_t n;
    unsigned generation = 0;        // engine binding the shards were uploaded under (b200msm_run checks it)
    std::vector<int> dev;           // CUDA device id per shard: memory is released by these, whatever the engine is bound to by then
    std::vector<DevBuf> shard;      // one per bound device
    std::vector<size_t> lo, cnt;    // index range per device
    // after b200msm_bases_precompute: window-major table per device (window 0 = the shard, which is then released)
    std::vector<DevBuf> table;
    int tbl_c = 0, tbl_nwin = 0;
    // b200msm_run holds it shared-style (one at a time is enough: the devices serialise anyway),
    // b200msm_bases_precompute exclusively while it swaps shards for tables
    mutable std::mutex mu;
};

namespace {

// One device's share of a host-buffer MSM, streamed: the points are cut into slices whose uploads
// (scalars, then bases) queue back to back on the copy stream while the compute stream groups
// and accumulates slice k into the shared buckets as soon as it has landed — the PCIe transfer
// of slices 1.. hides behind the accumulation of the slices before them, and the reduction and
// combination run once.  Result in cx.out; asynchronous.  ctx.mu held, ctx.dev current.
int msm_streamed(int group, DeviceCtx &cx, const Tun &tn, const void *h_bases, const void *d_resident, const TableRef *tbl,
                 const uint64_t *h_scalars, size_t n, int mont) {
    const bool g2 = group == B200MSM_G2;
    const size_t AB = aff_bytes(group);
    // Slice boundaries.  From page-locked memory the copies are asynchronous and faster per point than the
    // accumulation (128 B at ≈55 GB/s = 2.3 ns against ≈7 ns for G1, ≈26 ns for G2 when sliced), so the slices GROW by that
    // ratio: a small first slice gets the GPU going after ≈0.2 ms, each next one has landed when the previous is
    // done, and the late, large slices hold enough entries per bucket for the batched-affine rounds (which need
    // ≈24).  From pageable memory (a Rust Vec) the driver stages the copy at a fifth of that rate and blocks
    // the caller — the transfer is the critical path — so the slices stay equal and small (≈2^17 points: the
    // accumulation left after the last copy is short).  `stream_min` below 2^17 forces small equal slices (tests).
    size_t bounds[9];
    int K;
    {
        cudaPointerAttributes at;
        const bool pinned = cudaPointerGetAttributes(&at, h_scalars) == cudaSuccess && at.type == cudaMemoryTypeHost;
        cudaGetLastError();
        const size_t per = std::min<size_t>((size_t)1 << 17, std::max<size_t>(tn.stream_min, 1));
        const int kmax = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(tn.stream_slices, 8), n / per));
        if (!pinned || tn.stream_min < ((size_t)1 << 17) || kmax < 2) {
            K = kmax;
            for (int k = 0; k <= K; k++) bounds[k] = n * k / K;
        } else {
            const double h2d_ns = (double)(32 + (h_bases ? AB : 0)) / 55.0, acc_ns = g2 ? 26.0 : 7.0;   // (sliced accumulation: ≈7 ns per G1 point; measured best ratio 3 one-shot, 7–8 with resident bases)
            static const double env_ratio = getenv("B200MSM_SLICE_RATIO") ? atof(getenv("B200MSM_SLICE_RATIO")) : 0;   // (sweeps)
            static const int env_k = getenv("B200MSM_SLICE_K") ? atoi(getenv("B200MSM_SLICE_K")) : 0;
            const double ratio = env_ratio > 1 ? env_ratio : std::min(8.0, std::max(1.5, acc_ns / h2d_ns));
            K = 1;
            for (int k = 2; k <= (env_k ? std::min(env_k, kmax) : kmax); k++) {   // as many slices as leave the first one ≥ 2^16 points
                const double first = (double)n * (ratio - 1) / (std::pow(ratio, k) - 1);
                if (first >= 65536.0) K = k;
            }
            double acc = 0, tot = (std::pow(ratio, K) - 1) / (ratio - 1);
            bounds[0] = 0;
            for (int k = 0; k < K; k++) {
                acc += std::pow(ratio, k);
                bounds[k + 1] = k == K - 1 ? n : ((size_t)((double)n * acc / tot) + 127) / 128 * 128;
            }
        }
    }
    Plan pl;
    if (tbl) {  // resident fixed-base table: the table fixes the width, one bucket set
        pl.c = tbl->c;
        pl.nwin = tbl->nwin;
        pl.nbw = pl.nb = 1u << (pl.c - 1);
    } else auto_plan(n, g2, tn.glv_mode, tn.window_override, pl, tn.batch_affine);
    if (h_bases)
        if (int rc = cx.bases.reserve(n * AB)) return rc;
    if (cx.busy_valid) CUDA_TRY(cudaStreamWaitEvent(cx.stream, cx.ev_busy, 0));
    // the previous call's kernels may still read cx.scalars / cx.bases: copies wait for them too
    if (cx.busy_valid) CUDA_TRY(cudaStreamWaitEvent(cx.copy_stream, cx.ev_busy, 0));
    const char *dev_bases = h_bases ? (const char *)cx.bases.p : (const char *)d_resident;
    // Copies and kernels of slice k are issued together, slice after slice: with pinned host memory
    // everything is asynchronous anyway; with ordinary pageable memory (a Rust Vec) cudaMemcpyAsync
    // stages through the driver and blocks the host, and this order keeps the GPU accumulating
    // slice k while the host is stuck copying slice k+1 (19.5 → 13.7 ms at G1 2^20 one-shot).
    for (int k = 0; k < K; k++) {
        const size_t lo = bounds[k], hi = bounds[k + 1];
        CUDA_TRY(cudaMemcpyAsync((char *)cx.scalars.p + lo * 32, h_scalars + 4 * lo, (hi - lo) * 32, cudaMemcpyHostToDevice, cx.copy_stream));
        CUDA_TRY(cudaEventRecord(cx.ev_slice[2 * k], cx.copy_stream));
        if (h_bases) {
            CUDA_TRY(cudaMemcpyAsync((char *)cx.bases.p + lo * AB, (const char *)h_bases + lo * AB, (hi - lo) * AB, cudaMemcpyHostToDevice, cx.copy_stream));
            CUDA_TRY(cudaEventRecord(cx.ev_slice[2 * k + 1], cx.copy_stream));
        }
        PassOpts po;
        po.plan = &pl;
        po.into = k > 0;
        po.finish = k == K - 1;
        TableRef sub;
        if (tbl) { sub = *tbl; sub.p = (const char *)tbl->p + lo * AB; }  // same stride, shifted origin
        CUDA_TRY(cudaStreamWaitEvent(cx.stream, cx.ev_slice[2 * k], 0));
        if (int rc = run_pass(group, cx, tn, dev_bases ? dev_bases + lo * AB : nullptr, (const char *)cx.scalars.p + lo * 32, hi - lo, mont, cx.out.p,
                              cx.stream, h_bases ? cx.ev_slice[2 * k + 1] : nullptr, tbl ? &sub : nullptr, po))
            return rc;
    }
    CUDA_TRY(cudaEventRecord(cx.ev_busy, cx.stream));
    cx.busy_valid = true;
    return 0;
}

// Wait for everything a context may still have in flight.  Error paths call this before
// returning: the caller owns `bases` / `scalars` again the moment the call returns, so no copy
// may still be reading them (with page-locked memory the H2D copies are truly asynchronous), and
// the next call must not find work left over on the side streams.
void drain(DeviceCtx &cx) {
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(cx.dev);
    for (cudaStream_t st : {cx.copy_stream, cx.stream, cx.aux_stream, cx.tail_stream}) cudaStreamSynchronize(st);
    cudaGetLastError();
    cx.busy_valid = false;
    cudaSetDevice(prev);
}

// One device's share of a host-buffer MSM: uploads (streamed in slices when large), the whole
// pipeline, result in the first Jacobian slot of cx.out.  Asynchronous; ctx.mu held, ctx.dev current.
struct Share {
    int group = 0, mont = 0, d = 0;
    const uint64_t *bases = nullptr, *scalars = nullptr;  // already offset to this device's range
    size_t cnt = 0;
    const b200msm_bases *resident = nullptr;
    Tun tn;
};
int device_share(DeviceCtx &cx, const Share &sh) {
    const size_t AB = aff_bytes(sh.group), JB = jac_bytes(sh.group);
    const b200msm_bases *resident = sh.resident;
    const int d = sh.d;
    if (sh.cnt == 0) {
        if (cx.busy_valid) CUDA_TRY(cudaStreamWaitEvent(cx.stream, cx.ev_busy, 0));
        CUDA_TRY(cudaMemsetAsync(cx.out.p, 0, JB, cx.stream));
        return 0;
    }
    if (int rc = cx.scalars.reserve(sh.cnt * 32)) return rc;
    // large enough to be worth slicing (and, for resident bases, small enough for one pass per slice set)
    if (sh.cnt >= sh.tn.stream_min && sh.tn.stream_slices > 1 && sh.cnt <= ((size_t)1 << 24) && !sh.tn.max_chunk_override) {
        if (!resident) return msm_streamed(sh.group, cx, sh.tn, sh.bases, nullptr, nullptr, sh.scalars, sh.cnt, sh.mont);
        if (resident->tbl_c) {
            TableRef tr{resident->table[d].p, resident->cnt[d], resident->tbl_c, resident->tbl_nwin};
            return msm_streamed(sh.group, cx, sh.tn, nullptr, nullptr, &tr, sh.scalars, sh.cnt, sh.mont);
        }
        return msm_streamed(sh.group, cx, sh.tn, nullptr, resident->shard[d].p, nullptr, sh.scalars, sh.cnt, sh.mont);
    }
    // the previous call's kernels may still read cx.scalars / cx.bases
    if (cx.busy_valid) CUDA_TRY(cudaStreamWaitEvent(cx.copy_stream, cx.ev_busy, 0));
    // scalars first (the digit kernel needs them), then the bases behind them on the same copy
    // stream; the compute stream only waits for the bases right before the accumulation
    CUDA_TRY(cudaMemcpyAsync(cx.scalars.p, sh.scalars, sh.cnt * 32, cudaMemcpyHostToDevice, cx.copy_stream));
    CUDA_TRY(cudaEventRecord(cx.ev_scalars, cx.copy_stream));
    CUDA_TRY(cudaStreamWaitEvent(cx.stream, cx.ev_scalars, 0));
    const void *db;
    cudaEvent_t ready = nullptr;
    TableRef tr{nullptr, 0, 0, 0};
    if (resident && resident->tbl_c) {
        tr = TableRef{resident->table[d].p, resident->cnt[d], resident->tbl_c, resident->tbl_nwin};
        db = tr.p;
    } else if (resident) db = resident->shard[d].p;
    else {
        if (int rc = cx.bases.reserve(sh.cnt * AB)) return rc;
        CUDA_TRY(cudaMemcpyAsync(cx.bases.p, sh.bases, sh.cnt * AB, cudaMemcpyHostToDevice, cx.copy_stream));
        CUDA_TRY(cudaEventRecord(cx.ev_bases, cx.copy_stream));
        db = cx.bases.p;
        ready = cx.ev_bases;
    }
    return run_group(sh.group, cx, sh.tn, db, cx.scalars.p, sh.cnt, sh.mont, cx.out.p, cx.stream, ready, tr.c ? &tr : nullptr);
}

// One persistent host thread per bound device (only when more than one is bound).  The calling
// thread hands every device its share and sleeps; the workers issue copies and launches in
// parallel (a streamed share is ≈ 300 launches and, from pageable memory, blocking staged
// copies — issued device after device from one thread they would serialise), each sends its
// partial to device 0 over NVLink, and the caller adds the partials there.
static const bool g_trace = getenv("B200MSM_TRACE") != nullptr;   // host-side timeline of a sharded call on stderr (development aid)
static double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
void DeviceWorker::loop() {
        cudaSetDevice(cx->dev);
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return quit || job; });
            if (quit) return;
            const Share *sh = job;
            lk.unlock();
            int r;
            const double tw0 = now_ms();
            {
                std::lock_guard<std::mutex> cl(cx->mu);
                r = device_share(*cx, *sh);
                const size_t JB = jac_bytes(sh->group);
                if (!r && gather_dst) {  // partial → device 0 (peer copy over NVLink), then mark the share complete
                    cudaError_t e = cudaMemcpyPeerAsync(gather_dst, gather_dev, cx->out.p, cx->dev, JB, cx->stream);
                    if (e != cudaSuccess) r = fail(B200MSM_ECUDA, std::string("peer copy of a partial: ") + cudaGetErrorString(e));
                }
                if (!r) {
                    cudaError_t e = cudaEventRecord(cx->ev_share, cx->stream);
                    if (e != cudaSuccess) r = fail(B200MSM_ECUDA, std::string("cudaEventRecord: ") + cudaGetErrorString(e));
                }
                if (r) drain(*cx);
            }
            if (g_trace) {
                const double tw1 = now_ms();
                cudaStreamSynchronize(cx->stream);
                fprintf(stderr, "[b200msm trace] dev %d: issue %.3f ms (from %.3f), device done +%.3f ms\n", cx->dev, tw1 - tw0, tw0, now_ms() - tw1);
            }
            lk.lock();
            rc = r;
            err = r ? g_err : std::string();
            job = nullptr;
            done = true;
            cv.notify_all();
        }
    
}

int msm_host(int group, const uint64_t *bases, const uint64_t *scalars, size_t n, int mont, uint64_t *out,
             const b200msm_bases *resident) {
    if (group != B200MSM_G1 && group != B200MSM_G2) return fail(B200MSM_EINVAL, "bad group");
    if (!out || (n && (!scalars || (!bases && !resident)))) return fail(B200MSM_EINVAL, "null pointer");
    if (int rc = engine_init(-1, 1)) return rc;
    const size_t AB = aff_bytes(group), JB = jac_bytes(group);
    const int ndev = (int)g_eng.ctx.size();
    if (resident) {
        if (resident->generation != g_eng.generation || (int)resident->shard.size() != ndev)
            return fail(B200MSM_EINVAL, "resident bases were uploaded under another engine binding (b200msm_shutdown / b200msm_init since)");
    }
    if (n == 0) {
        memset(out, 0, JB);
        return 0;
    }
    const Tun tn = tun_snapshot();
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    // index ranges: the upload's own split when resident, else an even split
    std::vector<Share> shares(ndev);
    for (int d = 0; d < ndev; d++) {
        size_t lo, cnt;
        if (resident) {
            lo = resident->lo[d];
            cnt = lo >= n ? 0 : std::min(resident->cnt[d], n - lo);
        } else {
            lo = n * d / ndev;
            cnt = n * (d + 1) / ndev - lo;
        }
        Share &sh = shares[d];
        sh.group = group; sh.mont = mont; sh.d = d; sh.cnt = cnt; sh.resident = resident; sh.tn = tn;
        sh.bases = bases ? (const uint64_t *)((const char *)bases + lo * AB) : nullptr;
        sh.scalars = scalars + 4 * lo;
    }
    int rc = 0;
    DeviceCtx &c0 = *g_eng.ctx[0];
    if (ndev == 1) {
        std::lock_guard<std::mutex> lk(c0.mu);
        cudaSetDevice(c0.dev);
        if (!(rc = c0.out.reserve(JB * 2)) && !(rc = device_share(c0, shares[0]))) {
            cudaError_t e = cudaMemcpyAsync(out, c0.out.p, JB, cudaMemcpyDeviceToHost, c0.stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c0.stream);
            if (e != cudaSuccess) rc = fail(B200MSM_ECUDA, std::string("msm: ") + cudaGetErrorString(e));
        }
        if (rc) drain(c0);
    } else {
        std::lock_guard<std::mutex> ml(g_eng.multi_mu);
        // result slots: every device's own partial in its slot 0; device 0 also holds the gathered partials 1..ndev−1 and the sum
        for (int d = 0; d < ndev && !rc; d++) {
            DeviceCtx &cx = *g_eng.ctx[d];
            std::lock_guard<std::mutex> lk(cx.mu);
            cudaSetDevice(cx.dev);
            rc = cx.out.reserve(JB * (size_t)(ndev + 1));
        }
        const double tm0 = now_ms();
        if (!rc) {
            for (int d = 0; d < ndev; d++) {
                DeviceWorker &w = *g_eng.workers[d];
                std::lock_guard<std::mutex> lk(w.mu);
                w.gather_dst = d ? (char *)c0.out.p + JB * (size_t)d : nullptr;
                w.gather_dev = c0.dev;
                w.done = false;
                w.job = &shares[d];
                w.cv.notify_all();
            }
            for (int d = 0; d < ndev; d++) {       // every share issued (not yet finished)
                DeviceWorker &w = *g_eng.workers[d];
                std::unique_lock<std::mutex> lk(w.mu);
                w.cv.wait(lk, [&] { return w.done; });
                if (w.rc && !rc) rc = fail(w.rc, w.err);
            }
            const double tm1 = now_ms();
            std::lock_guard<std::mutex> lk(c0.mu);
            cudaSetDevice(c0.dev);
            if (g_trace) fprintf(stderr, "[b200msm trace] posted at %.3f, all shares issued +%.3f ms\n", tm0, tm1 - tm0);
            if (!rc) {
                // device 0 waits for every share (its own included), adds the partials, downloads
                cudaError_t e = cudaSuccess;
                for (int d = 1; d < ndev && e == cudaSuccess; d++) e = cudaStreamWaitEvent(c0.stream, g_eng.ctx[d]->ev_share, 0);
                void *sum = (char *)c0.out.p + JB * (size_t)ndev;
                if (e == cudaSuccess) {
                    (group == B200MSM_G1 ? launch_sum_partials_g1 : launch_sum_partials_g2)((const uint32_t *)c0.out.p, ndev, (uint32_t *)sum, c0.stream);
                    e = cudaMemcpyAsync(out, sum, JB, cudaMemcpyDeviceToHost, c0.stream);
                }
                if (e == cudaSuccess) e = cudaStreamSynchronize(c0.stream);
                if (g_trace) fprintf(stderr, "[b200msm trace] combined on device 0 +%.3f ms after issue\n", now_ms() - tm1);
                if (e != cudaSuccess) rc = fail(B200MSM_ECUDA, std::string("msm combine: ") + cudaGetErrorString(e));
                // the other devices are idle now (device 0 waited for them), but say so to their contexts:
                // the caller's buffers must not be in use once we return
                for (int d = 1; d < ndev; d++) {
                    cudaSetDevice(g_eng.ctx[d]->dev);
                    cudaError_t e2 = cudaStreamSynchronize(g_eng.ctx[d]->stream);
                    if (e2 != cudaSuccess && !rc) rc = fail(B200MSM_ECUDA, std::string("msm shard: ") + cudaGetErrorString(e2));
                }
            }
        }
        if (rc) {
            const std::string keep = g_err;
            for (int d = 0; d < ndev; d++) drain(*g_eng.ctx[d]);
            g_err = keep;
        }
    }
    cudaSetDevice(prev_dev);
    return rc;
}

}  // namespace

// =============================================================================================
extern "C" {

int b200msm_init(int first_device, int n_devices) { return engine_init(first_device, n_devices); }

void b200msm_shutdown(void) {
    std::lock_guard<std::mutex> ml(g_eng.multi_mu);
    std::lock_guard<std::mutex> lk(g_eng.mu);
    int prev = 0;
    cudaGetDevice(&prev);
    g_eng.stop_workers();
    for (auto &c : g_eng.ctx) destroy_ctx(*c);
    for (auto &c : g_eng.lane_ctx) destroy_ctx(*c);
    g_eng.lane_ctx.clear();
    g_eng.ctx.clear();
    g_eng.inited = false;
    cudaSetDevice(prev);
}
int b200msm_device_count(void) {
    std::lock_guard<std::mutex> lk(g_eng.mu);
    return (int)g_eng.ctx.size();
}
const char *b200msm_last_error(void) { return g_err.c_str(); }
const char *b200msm_version(void) { return "b200msm 0.1 (sm_100a)"; }

int b200msm_g1(const uint64_t *bases, const uint64_t *scalars, size_t n, int mont, uint64_t out[18]) {
    return msm_host(B200MSM_G1, bases, scalars, n, mont, out, nullptr);
}
int b200msm_g2(const uint64_t *bases, const uint64_t *scalars, size_t n, int mont, uint64_t out[36]) {
    return msm_host(B200MSM_G2, bases, scalars, n, mont, out, nullptr);
}

int b200msm_bases_upload(int group, const uint64_t *bases, size_t n, b200msm_bases **handle) {
    if (!handle || (n && !bases)) return fail(B200MSM_EINVAL, "null pointer");
    if (group != B200MSM_G1 && group != B200MSM_G2) return fail(B200MSM_EINVAL, "bad group");
    if (int rc = engine_init(-1, 1)) return rc;
    auto h = std::make_unique<b200msm_bases>();
    h->group = group;
    h->n = n;
    const int ndev = (int)g_eng.ctx.size();
    const size_t AB = aff_bytes(group);
    h->generation = g_eng.generation;
    for (auto &c : g_eng.ctx) h->dev.push_back(c->dev);
    h->shard.resize(ndev);
    h->lo.resize(ndev);
    h->cnt.resize(ndev);
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = 0;
    for (int d = 0; d < ndev && !rc; d++) {
        h->lo[d] = n * d / ndev;
        h->cnt[d] = n * (d + 1) / ndev - h->lo[d];
        cudaSetDevice(g_eng.ctx[d]->dev);
        if (h->cnt[d] == 0) continue;
        if ((rc = h->shard[d].reserve(h->cnt[d] * AB))) break;
        cudaError_t e = cudaMemcpy(h->shard[d].p, (const char *)bases + h->lo[d] * AB, h->cnt[d] * AB, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = fail(B200MSM_ECUDA, std::string("upload: ") + cudaGetErrorString(e));
    }
    cudaSetDevice(prev);
    if (rc) {
        for (auto &s : h->shard) s.release();
        return rc;
    }
    *handle = h.release();
    return 0;
}
int b200msm_bases_free(b200msm_bases *h) {
    if (!h) return 0;
    int prev = 0;
    cudaGetDevice(&prev);
    {
        std::lock_guard<std::mutex> hl(h->mu);   // a b200msm_run still using the handle finishes first
        for (size_t d = 0; d < h->shard.size(); d++) {   // by the device ids of the upload: valid after b200msm_shutdown too
            cudaSetDevice(h->dev[d]);
            h->shard[d].release();
            if (d < h->table.size()) h->table[d].release();
        }
    }
    cudaSetDevice(prev);
    delete h;
    return 0;
}
int b200msm_run(const b200msm_bases *h, const uint64_t *scalars, size_t n, int mont, uint64_t *out) {
    if (!h) return fail(B200MSM_EINVAL, "null handle");
    if (n > h->n) return fail(B200MSM_EINVAL, "n exceeds the uploaded base count");
    std::lock_guard<std::mutex> hl(h->mu);       // not while b200msm_bases_precompute swaps the shards for tables
    return msm_host(h->group, nullptr, scalars, n, mont, out, h);
}

int b200msm_table_plan(int group, size_t n, int *window_bits, int *windows) {
    if (group != B200MSM_G1 && group != B200MSM_G2) return fail(B200MSM_EINVAL, "bad group");
    if (!window_bits || !windows) return fail(B200MSM_EINVAL, "null pointer");
    int c = *window_bits;
    if (c == 0) c = table_plan(std::max<size_t>(n, 1), group == B200MSM_G2);
    if (c < 2 || c > 23) return fail(B200MSM_EINVAL, "table window bits must be 0 (auto) or 2..23");
    if ((uint64_t)n * (uint64_t)((256 + c - 1) / c) >= (1ull << 31))
        return fail(B200MSM_EINVAL, "table too large: windows × points per device must stay below 2^31 (shard the bases)");
    *window_bits = c;
    *windows = (256 + c - 1) / c;
    return 0;
}
int b200msm_bases_precompute(b200msm_bases *h, int window_bits) {
    if (!h) return fail(B200MSM_EINVAL, "null handle");
    std::lock_guard<std::mutex> hl(h->mu);
    if (h->tbl_c) return 0;  // already a table
    if (int rc = engine_init(-1, 1)) return rc;
    if (h->generation != g_eng.generation || h->shard.size() != g_eng.ctx.size())
        return fail(B200MSM_EINVAL, "resident bases were uploaded under another engine binding (b200msm_shutdown / b200msm_init since)");
    const int ndev = (int)h->shard.size();
    size_t maxcnt = 0;
    for (int d = 0; d < ndev; d++) maxcnt = std::max(maxcnt, h->cnt[d]);
    int c = window_bits, nwin = 0;
    if (int rc = b200msm_table_plan(h->group, maxcnt, &c, &nwin)) return rc;
    const size_t AB = aff_bytes(h->group);
    int prev = 0;
    cudaGetDevice(&prev);
    h->table.resize(h->shard.size());
    int rc = 0;
    for (int d = 0; d < ndev && !rc; d++) {
        if (h->cnt[d] == 0) continue;
        DeviceCtx &cx = *g_eng.ctx[d];
        std::lock_guard<std::mutex> lk(cx.mu);
        cudaSetDevice(cx.dev);
        if ((rc = h->table[d].reserve((size_t)nwin * h->cnt[d] * AB))) break;
        cudaMemcpyAsync(h->table[d].p, h->shard[d].p, h->cnt[d] * AB, cudaMemcpyDeviceToDevice, cx.stream);
        rc = build_table(h->group, cx, h->table[d].p, h->cnt[d], h->cnt[d], c, nwin, cx.stream);
    }
    for (int d = 0; d < ndev; d++) {
        if (h->cnt[d] == 0) continue;
        DeviceCtx &cx = *g_eng.ctx[d];
        cudaSetDevice(cx.dev);
        cudaError_t e = cudaStreamSynchronize(cx.stream);
        if (e != cudaSuccess && !rc) rc = fail(B200MSM_ECUDA, std::string("precompute: ") + cudaGetErrorString(e));
    }
    if (rc) {
        for (int d = 0; d < ndev; d++) { cudaSetDevice(g_eng.ctx[d]->dev); h->table[d].release(); }
    } else {
        for (int d = 0; d < ndev; d++) { cudaSetDevice(g_eng.ctx[d]->dev); h->shard[d].release(); }
        h->tbl_c = c;
        h->tbl_nwin = nwin;
    }
    cudaSetDevice(prev);
    return rc;
}
int b200msm_bases_table_info(const b200msm_bases *h, int *window_bits, int *windows, size_t *device_bytes) {
    if (!h) return fail(B200MSM_EINVAL, "null handle");
    if (window_bits) *window_bits = h->tbl_c;
    if (windows) *windows = h->tbl_nwin;
    if (device_bytes) {
        size_t b = 0;
        for (auto &t : h->table) b += t.cap;
        for (auto &t : h->shard) b += t.cap;
        *device_bytes = b;
    }
    return 0;
}
int b200msm_table_build_device(int group, const void *d_bases, size_t n, int window_bits, void *d_table, void *stream) {
    if (group != B200MSM_G1 && group != B200MSM_G2) return fail(B200MSM_EINVAL, "bad group");
    if (n == 0) return 0;
    if (!d_bases || !d_table) return fail(B200MSM_EINVAL, "null pointer");
    if (window_bits < 2 || window_bits > 23) return fail(B200MSM_EINVAL, "table window bits must be 2..23 (see b200msm_table_plan)");
    if (int rc = engine_init(-1, 1)) return rc;
    DeviceCtx *cx = ctx_for_current_device();
    if (!cx) return fail(B200MSM_EINVAL, "current device is not bound to the engine");
    std::lock_guard<std::mutex> lk(cx->mu);
    if (d_table != d_bases)
        CUDA_TRY(cudaMemcpyAsync(d_table, d_bases, n * aff_bytes(group), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return build_table(group, *cx, d_table, n, n, window_bits, (256 + window_bits - 1) / window_bits, (cudaStream_t)stream);
}
int b200msm_run_table_device(int group, const void *d_table, size_t stride, int window_bits, const void *d_scalars, size_t n,
                             int mont, void *d_out, void *stream) {
    if (!d_out || (n && (!d_table || !d_scalars))) return fail(B200MSM_EINVAL, "null pointer");
    if (window_bits < 2 || window_bits > 23 || n > stride || (uint64_t)stride * ((256 + window_bits - 1) / window_bits) >= (1ull << 31))
        return fail(B200MSM_EINVAL, "bad table description");
    if (int rc = engine_init(-1, 1)) return rc;
    DeviceCtx *cx = ctx_for_current_device();
    if (!cx) return fail(B200MSM_EINVAL, "current device is not bound to the engine");
    std::lock_guard<std::mutex> lk(cx->mu);
    TableRef tr{d_table, stride, window_bits, (256 + window_bits - 1) / window_bits};
    return run_group(group, *cx, tun_snapshot(), d_table, d_scalars, n, mont, d_out, (cudaStream_t)stream, nullptr, &tr);
}

int b200msm_run_device(int group, const void *d_bases, const void *d_scalars, size_t n, int mont, void *d_out,
                       void *stream) {
    if (!d_out || (n && (!d_bases || !d_scalars))) return fail(B200MSM_EINVAL, "null pointer");
    if (int rc = engine_init(-1, 1)) return rc;
    DeviceCtx *cx = ctx_for_current_device();
    if (!cx) return fail(B200MSM_EINVAL, "current device is not bound to the engine");
    std::lock_guard<std::mutex> lk(cx->mu);
    return run_group(group, *cx, tun_snapshot(), d_bases, d_scalars, n, mont, d_out, (cudaStream_t)stream);
}
int b200msm_sum_partials_device(int group, const void *d_partials, int count, void *d_out, void *stream) {
    if (!d_partials || !d_out || count < 0) return fail(B200MSM_EINVAL, "bad argument");
    if (group == B200MSM_G1) launch_sum_partials_g1((const uint32_t *)d_partials, count, (uint32_t *)d_out, (cudaStream_t)stream);
    else if (group == B200MSM_G2) launch_sum_partials_g2((const uint32_t *)d_partials, count, (uint32_t *)d_out, (cudaStream_t)stream);
    else return fail(B200MSM_EINVAL, "bad group");
    CUDA_TRY(cudaGetLastError());
    return 0;
}

unsigned long long b200msm_launch_count(void) { return __atomic_load_n(&g_own_launches, __ATOMIC_RELAXED); }

int b200msm_normalize_batch_device(int group, const void *d_proj, size_t n, void *d_affine, void *stream) {
    if (group != B200MSM_G1 && group != B200MSM_G2) return fail(B200MSM_EINVAL, "bad group");
    if (n == 0) return 0;
    if (!d_proj || !d_affine) return fail(B200MSM_EINVAL, "null pointer");
    if (int rc = engine_init(-1, 1)) return rc;
    DeviceCtx *cx = ctx_for_current_device();
    if (!cx) return fail(B200MSM_EINVAL, "current device is not bound to the engine");
    launch_normalize_batch(group == B200MSM_G2, (const uint32_t *)d_proj, n, (uint32_t *)d_affine, cx->sm_count, (cudaStream_t)stream);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int b200msm_normalize_batch(int group, const uint64_t *proj, size_t n, uint64_t *affine_out) {
    if (group != B200MSM_G1 && group != B200MSM_G2) return fail(B200MSM_EINVAL, "bad group");
    if (n == 0) return 0;
    if (!proj || !affine_out) return fail(B200MSM_EINVAL, "null pointer");
    if (int rc = engine_init(-1, 1)) return rc;
    DeviceCtx &cx = *g_eng.ctx[0];
    std::lock_guard<std::mutex> lk(cx.mu);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(cx.dev);
    const size_t JB = jac_bytes(group), AB = aff_bytes(group);
    int rc = 0;
    if (!(rc = cx.norm_in.reserve(n * JB)) && !(rc = cx.norm_out.reserve(n * AB))) {
        cudaMemcpyAsync(cx.norm_in.p, proj, n * JB, cudaMemcpyHostToDevice, cx.stream);
        launch_normalize_batch(group == B200MSM_G2, cx.norm_in.as<uint32_t>(), n, cx.norm_out.as<uint32_t>(), cx.sm_count, cx.stream);
        cudaMemcpyAsync(affine_out, cx.norm_out.p, n * AB, cudaMemcpyDeviceToHost, cx.stream);
        cudaError_t e = cudaStreamSynchronize(cx.stream);
        if (e != cudaSuccess) rc = fail(B200MSM_ECUDA, std::string("normalize_batch: ") + cudaGetErrorString(e));
    }
    cudaSetDevice(prev);
    return rc;
}

int b200msm_deserialize(int group, const uint8_t *in, size_t n, int compressed, int validate, uint64_t *affine_out,
                        uint8_t *status_out) {
    if (group != B200MSM_G1 && group != B200MSM_G2) return fail(B200MSM_EINVAL, "bad group");
    if (n == 0) return 0;
    if (!in || !affine_out || !status_out) return fail(B200MSM_EINVAL, "null pointer");
    if (int rc = engine_init(-1, 1)) return rc;
    DeviceCtx &cx = *g_eng.ctx[0];
    std::lock_guard<std::mutex> lk(cx.mu);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(cx.dev);
    const size_t AB = aff_bytes(group), EB = compressed ? AB / 2 : AB;
    int rc = 0;
    if (!(rc = cx.norm_in.reserve(n * EB + n)) && !(rc = cx.norm_out.reserve(n * AB))) {
        uint8_t *d_in = cx.norm_in.as<uint8_t>(), *d_st = d_in + n * EB;
        cudaMemcpyAsync(d_in, in, n * EB, cudaMemcpyHostToDevice, cx.stream);
        launch_deserialize(group == B200MSM_G2, d_in, n, compressed, validate, cx.norm_out.as<uint32_t>(), d_st, cx.stream);
        cudaMemcpyAsync(affine_out, cx.norm_out.p, n * AB, cudaMemcpyDeviceToHost, cx.stream);
        cudaMemcpyAsync(status_out, d_st, n, cudaMemcpyDeviceToHost, cx.stream);
        cudaError_t e = cudaStreamSynchronize(cx.stream);
        if (e != cudaSuccess) rc = fail(B200MSM_ECUDA, std::string("deserialize: ") + cudaGetErrorString(e));
    }
    cudaSetDevice(prev);
    return rc;
}
int b200msm_serialize(int group, const uint64_t *affine, size_t n, int compressed, uint8_t *out) {
    if (group != B200MSM_G1 && group != B200MSM_G2) return fail(B200MSM_EINVAL, "bad group");
    if (n == 0) return 0;
    if (!affine || !out) return fail(B200MSM_EINVAL, "null pointer");
    if (int rc = engine_init(-1, 1)) return rc;
    DeviceCtx &cx = *g_eng.ctx[0];
    std::lock_guard<std::mutex> lk(cx.mu);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(cx.dev);
    const size_t AB = aff_bytes(group), EB = compressed ? AB / 2 : AB;
    int rc = 0;
    if (!(rc = cx.norm_in.reserve(n * AB)) && !(rc = cx.norm_out.reserve(n * EB))) {
        cudaMemcpyAsync(cx.norm_in.p, affine, n * AB, cudaMemcpyHostToDevice, cx.stream);
        launch_serialize(group == B200MSM_G2, cx.norm_in.as<uint32_t>(), n, compressed, cx.norm_out.as<uint8_t>(), cx.stream);
        cudaMemcpyAsync(out, cx.norm_out.p, n * EB, cudaMemcpyDeviceToHost, cx.stream);
        cudaError_t e = cudaStreamSynchronize(cx.stream);
        if (e != cudaSuccess) rc = fail(B200MSM_ECUDA, std::string("serialize: ") + cudaGetErrorString(e));
    }
    cudaSetDevice(prev);
    return rc;
}

int b200msm_set_window_bits(int c) {
    if (c < 0 || c == 1 || c > 22) return fail(B200MSM_EINVAL, "window bits must be 0 (auto) or 2..22");
    tun_update([&](Tun &t) { t.window_override = c; });
    return 0;
}
int b200msm_set_stream_slices(int slices, size_t min_points) {
    if (slices < 1 || slices > 8) return fail(B200MSM_EINVAL, "stream slices must be 1 (off) .. 8");
    tun_update([&](Tun &t) {
        t.stream_slices = slices;
        t.stream_min = min_points ? min_points : (size_t)1 << 18;
    });
    return 0;
}
int b200msm_host_register(const void *ptr, size_t bytes) {
    if (!ptr || !bytes) return fail(B200MSM_EINVAL, "null pointer");
    CUDA_TRY(cudaHostRegister(const_cast<void *>(ptr), bytes, cudaHostRegisterPortable));
    return 0;
}
int b200msm_host_unregister(const void *ptr) {
    if (!ptr) return fail(B200MSM_EINVAL, "null pointer");
    CUDA_TRY(cudaHostUnregister(const_cast<void *>(ptr)));
    return 0;
}
int b200msm_set_lane(int lane) {
    if (lane < 0 || lane > 7) return fail(B200MSM_EINVAL, "lane must be 0..7");
    t_lane = lane;
    return 0;
}
int b200msm_set_heavy_factor(int f) {
    if (f < 0 || f > 1 << 20) return fail(B200MSM_EINVAL, "heavy factor must be 0 (automatic) or 1..2^20");
    tun_update([&](Tun &t) { t.heavy_factor = f; });
    return 0;
}
int b200msm_set_glv(int mode) {
    if (mode < -1 || mode > 2) return fail(B200MSM_EINVAL, "glv mode must be -1 (auto), 0 (off), 1 (two parts) or 2 (four parts on G2)");
    tun_update([&](Tun &t) { t.glv_mode = mode; });
    return 0;
}
int b200msm_set_batch_affine(int rounds) {
    if (rounds < -1 || rounds > 3) return fail(B200MSM_EINVAL, "batched-affine rounds must be -1 (auto), 0 (off) or 1..3");
    tun_update([&](Tun &t) { t.batch_affine = rounds; });
    return 0;
}
int b200msm_set_graphs(int on) {
    tun_update([&](Tun &t) { t.graphs = on != 0; }, false);
    return 0;
}
int b200msm_set_max_chunk(size_t max_points_per_pass) {
    tun_update([&](Tun &t) { t.max_chunk_override = max_points_per_pass; });
    return 0;
}
int b200msm_set_profiling(int on) {
    tun_update([&](Tun &t) { t.profiling = on != 0; }, false);
    return 0;
}
int b200msm_plan_query(int group, size_t n, int glv_mode, int out[4]) {
    if (group != B200MSM_G1 && group != B200MSM_G2) return fail(B200MSM_EINVAL, "bad group");
    if (!out || n == 0 || glv_mode < -1 || glv_mode > 2) return fail(B200MSM_EINVAL, "bad argument");
    Plan pl;
    auto_plan(n, group == B200MSM_G2, glv_mode, 0, pl, tun_snapshot().batch_affine);   // host arithmetic only: no device needed
    out[0] = pl.c;
    out[1] = pl.nwin;
    out[2] = pl.glv ? (pl.parts == 4 ? 3 : 1) + (pl.split ? 1 : 0) : 0;
    out[3] = (int)std::min<uint64_t>(pl.nb, 0x7fffffff);
    return 0;
}
int b200msm_last_plan(int out[4]) {
    if (!g_eng.inited) return fail(B200MSM_EINVAL, "engine not initialised");
    DeviceCtx *cx = ctx_for_current_device();
    if (!cx) cx = g_eng.ctx[0].get();
    std::lock_guard<std::mutex> lk(cx->mu);
    memcpy(out, cx->last_plan, sizeof cx->last_plan);
    return 0;
}
int b200msm_last_phase_ms(double out[8]) {
    if (!g_eng.inited) return fail(B200MSM_EINVAL, "engine not initialised");
    DeviceCtx *cx = ctx_for_current_device();
    if (!cx) cx = g_eng.ctx[0].get();
    std::lock_guard<std::mutex> lk(cx->mu);
    if (cx->phase_pending) {
        CUDA_TRY(cudaEventSynchronize(cx->ev[6]));
        float ms;
        for (int i = 0; i < 6; i++) {
            CUDA_TRY(cudaEventElapsedTime(&ms, cx->ev[i], cx->ev[i + 1]));
            cx->phase_ms[i] = ms;
        }
        CUDA_TRY(cudaEventElapsedTime(&ms, cx->ev[0], cx->ev[6]));
        cx->phase_ms[6] = ms;
        cx->phase_ms[7] = 1;
        cx->phase_pending = false;
    }
    memcpy(out, cx->phase_ms, sizeof cx->phase_ms);
    return 0;
}

int b200msm_synth_bases_device(int group, uint64_t seed, size_t n, void *d_out, void *stream) {
    if (!d_out && n) return fail(B200MSM_EINVAL, "null pointer");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (group == B200MSM_G1) launch_synth_bases_g1(seed, n, (uint32_t *)d_out, st);
    else if (group == B200MSM_G2) launch_synth_bases_g2(seed, n, (uint32_t *)d_out, st);
    else return fail(B200MSM_EINVAL, "bad group");
    CUDA_TRY(cudaGetLastError());
    return 0;
}
int b200msm_synth_scalars_device(uint64_t seed, size_t n, int montgomery, void *d_out, void *stream) {
    if (!d_out && n) return fail(B200MSM_EINVAL, "null pointer");
    if (n == 0) return 0;
    launch_synth_scalars(seed, n, montgomery, (uint32_t *)d_out, (cudaStream_t)stream);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int b200msm_imad_peak(double out[3]) {
    cudaDeviceProp p;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaGetDeviceProperties(&p, dev));
    const int blocks = p.multiProcessorCount * 4, threads = 256, iters = 2000;
    uint32_t *buf = nullptr;
    CUDA_TRY(cudaMalloc(&buf, (size_t)blocks * threads * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double rate[2] = {0, 0};
    for (int mode = 0; mode < 2; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            launch_imad_peak(mode, blocks, threads, buf, rep ? iters : 10, 0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        double ops = (double)blocks * threads * iters * 64.0 * 8.0 * (mode == 1 ? 2.0 : 1.0);
        rate[mode] = ops / (best * 1e-3);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    CUDA_TRY(cudaGetLastError());
    out[0] = rate[0];
    out[1] = rate[1];
    out[2] = rate[0] / (64.0 * p.multiProcessorCount) / 1e6;  // MHz if the pipe issues 64 IMAD/clk/SM
    return 0;
}

// ---- unit hooks ----
int b200msm_dbg_field_op(int is_fp2, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    if (n == 0) return 0;
    const size_t EB = is_fp2 ? 96 : 48;
    uint32_t *da = nullptr, *db = nullptr, *dout = nullptr;
    CUDA_TRY(cudaMalloc(&da, n * EB));
    CUDA_TRY(cudaMalloc(&db, n * EB));
    CUDA_TRY(cudaMalloc(&dout, n * EB));
    CUDA_TRY(cudaMemcpy(da, a, n * EB, cudaMemcpyHostToDevice));
    if (b) CUDA_TRY(cudaMemcpy(db, b, n * EB, cudaMemcpyHostToDevice));
    if (op == 6) launch_dbg_inv_sg(is_fp2, da, dout, n, 0);
    else launch_dbg_field_op(is_fp2, op, da, db, dout, n);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, dout, n * EB, cudaMemcpyDeviceToHost));
    cudaFree(da);
    cudaFree(db);
    cudaFree(dout);
    return 0;
}
int b200msm_dbg_point_op(int group, int op, const uint64_t *acc, const uint64_t *q, uint64_t *out, size_t n) {
    if (n == 0) return 0;
    const size_t FB = group == B200MSM_G2 ? 96 : 48;
    const size_t QB = (op == 0 ? 2 : 4) * FB;  // ops 3/4: quad-distributed add / dbl
    uint32_t *dacc = nullptr, *dq = nullptr, *dout = nullptr;
    CUDA_TRY(cudaMalloc(&dacc, n * 4 * FB));
    CUDA_TRY(cudaMalloc(&dq, n * QB));
    CUDA_TRY(cudaMalloc(&dout, n * 3 * FB));
    CUDA_TRY(cudaMemcpy(dacc, acc, n * 4 * FB, cudaMemcpyHostToDevice));
    if (q && op != 2 && op != 4) CUDA_TRY(cudaMemcpy(dq, q, n * QB, cudaMemcpyHostToDevice));
    (group == B200MSM_G2 ? launch_dbg_point_op_g2 : launch_dbg_point_op_g1)(op, dacc, dq, dout, n);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, dout, n * 3 * FB, cudaMemcpyDeviceToHost));
    cudaFree(dacc);
    cudaFree(dq);
    cudaFree(dout);
    return 0;
}
int b200msm_dbg_digits(const uint64_t *scalars, size_t n, int montgomery, int c, int32_t *out, int *nwin) {
    if (c < 2 || c > 24 || !out || !nwin) return fail(B200MSM_EINVAL, "bad argument");
    int W = (256 + c - 1) / c;
    *nwin = W;
    if (n == 0) return 0;
    uint32_t *ds = nullptr;
    int *dout = nullptr;
    CUDA_TRY(cudaMalloc(&ds, n * 32));
    CUDA_TRY(cudaMalloc(&dout, n * (size_t)W * 4));
    CUDA_TRY(cudaMemcpy(ds, scalars, n * 32, cudaMemcpyHostToDevice));
    launch_digits_dbg(ds, n, montgomery, c, W, dout, 0);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, dout, n * (size_t)W * 4, cudaMemcpyDeviceToHost));
    cudaFree(ds);
    cudaFree(dout);
    return 0;
}

}  // extern "C"
