// tapir_cuda.cu -- C ABI (include/tapir_cuda.h) over the sm_100a kernels in kernels.cuh.
//
// Host-side responsibilities: validate a CSR batch, build the work lists (time tiles, site
// splits, quad blocks), own device memory / streams, launch, time with CUDA events, copy back.
// There is no CPU fallback anywhere in this file: without a device every call fails.
#include "../../include/tapir_cuda.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <numeric>
#include <string>
#include <vector>
#include <chrono>

#include "kernels.cuh"
#include "mom_epoch.h"

using namespace tapir;

namespace {

thread_local std::string g_last_error;

int fail(tapir_cuda_ctx* ctx, int code, const char* fmt, ...);

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&p, want);
        }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

struct DevView {  // a slice of a DevBuf
    void* p = nullptr;
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

struct HostBuf {  // pinned staging
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes + bytes / 8 + 256);
        if (e == cudaSuccess) cap = bytes + bytes / 8 + 256;
        return e;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

struct Plan {
    int64_t L = 0, T = 0, nnz = 0;
    int32_t K = 0, E = 0;
    uint32_t flags = 0;
    int rh = 4;       // profile kernel variant
    bool mom_form = false;  // profile_mom_kernel<rh> (moment form) instead of profile_kernel<rh>
    int mom_terms = 32;     // moments per 128-point block of the moment form (32 or 16, see choose_moments)
    bool fp32 = false;      // TAPIR_FP32_MODE: profile_f32_kernel (16 moments, class 64 c <= 2.5, 1,024-point tiles)
    int32_t TT = 1024, n_tiles = 1;
    int64_t n_prof_items = 0, n_quad_items = 0, ws_rows = 0, n_parts = 0;
    std::vector<EpochBatch> batches;
    std::vector<MomEpochPlan> mom;   // per epoch (input order): is it served by the moment functional, and how
    int32_t n_mom = 0;
    uint64_t mom_mask = 0;
    size_t small_off[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, small_bytes = 0;  // layout of d_small / h_small
    bool valid = false;
};

}  // namespace

// One in-flight batch (or chunk of a batch): its own streams, device buffers and pinned staging, so that
// the lanes can overlap H2D, kernels and D2H of consecutive chunks.
struct Lane {
    cudaStream_t stream = nullptr;    // copies, profile, finalize
    cudaStream_t stream_q = nullptr;  // epoch-integral kernels, concurrent with the profile kernel
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;  // upload landed / kernels done / download landed
    cudaEvent_t ev_dstart = nullptr;   // trace only: the download stream has reached this lane's copies
    double host_us[4] = {0, 0, 0, 0};  // trace only: host clock when upload / compute / download were enqueued
    cudaGraphExec_t graph_exec = nullptr;  // staged path: the captured launch sequence of the resident batch
    int graph_launches = 0;
    Plan plan;
    // inputs: the rate stream, and ONE buffer holding every work list (a single H2D copy per batch)
    DevBuf d_rates, d_work, d_counter;
    DevView d_prof_items, d_quad_items, d_nsplit, d_ws_base, d_qitem_begin, d_time_idx, d_part_base, d_mom_epochs, d_mom_p;
    // intermediates
    DevBuf d_ws, d_tile_max, d_tile_arg, d_qpart, d_fix, d_mom_int, d_rare_list;
    // results: the profile, and ONE buffer holding every small per-locus vector (a single D2H copy per batch)
    DevBuf d_pi_net, d_small;
    DevView d_pi_times, d_integral, d_abserr, d_max_pi, d_argmax_t, d_argmax_epoch, d_mean_rate, d_n_finite, d_n_multi;
    HostBuf h_stage, h_small;
    // host destination of the small vectors of the batch in flight (scattered after the D2H copy lands)
    tapir_results pending_out = {};
    int64_t pending_row0 = 0;
    bool pending = false;
    DevBuf* all_bufs[12] = {&d_rates, &d_work, &d_counter, &d_ws, &d_tile_max, &d_tile_arg, &d_qpart, &d_fix, &d_pi_net, &d_small,
                            &d_mom_int, &d_rare_list};
};

constexpr int kLanes = 8;  // upper bound; ctx->n_lanes of them rotate (TAPIR_CUDA_LANES)

struct tapir_cuda_ctx {
    int device = 0;
    int sm_count = 148;
    int serial_kernels = 0;  // 1: profile and integral kernels on one stream (per-kernel timing is then exact)
    int use_graphs = 1;      // staged path replays a captured CUDA graph (TAPIR_CUDA_GRAPHS=0 disables)
    int n_lanes = 8;         // lanes the chunks of tapir_cuda_run rotate over
    int ramp = 4;            // tapir_cuda_run opens with chunks of 1/ramp, 2/ramp, ... of the regular size (<= 1: off)
    double* d_mt_pow = nullptr;       // profile_mom_kernel's table of powers (kernels.cuh mt_pow_table), written once
    long long* prof_trace = nullptr;  // TAPIR_PROF_TRACE=<file>: per-warp time stamps of profile_kernel, dumped by tapir_cuda_compute
    size_t prof_trace_words = 0;
    std::string prof_trace_path;
    int prof_mom = 1;        // the tile-size model may pick profile_mom_kernel (slow sites in moment form); TAPIR_PROF_MOM=0: never
    int mom_int = 1;         // epoch integrals of the moment class from the locus' moments (mom_epoch.h); TAPIR_MOM_INT=0: per site
    int last_rh = 0, last_mom_form = 0, last_n_mom = 0;  // what the last tapir_cuda_compute launched (tapir_timing)
    // per-device launch configuration, filled once in tapir_cuda_create (index: log2(RH))
    int prof_per_sm[4] = {0, 0, 0, 0}, mom_per_sm[4] = {0, 0, 0, 0}, mom16_per_sm[4] = {0, 0, 0, 0};
    int force_terms = 0;                       // TAPIR_MOM_TERMS = 16 / 32
    int force_fp32 = 0;                        // TAPIR_FP32_MODE=1 in the environment: every batch of this context
    int f32_per_sm = 0;
    // plans of the moment-functional epochs seen so far (they depend on the epoch and the tile geometry only; a
    // chunked run would otherwise recompute them for every chunk)
    struct MomCacheEntry { double a, b; int64_t n_tiles; int32_t TT, km; double class_x; MomEpochPlan plan; };
    std::vector<MomCacheEntry> mom_cache;
    int prof_ctas_cap = 0, quad_ctas_cap = 0;  // TAPIR_PROF_CTAS_PER_SM / TAPIR_QUAD_CTAS_PER_SM (co-residency experiments)
    int force_rh = 0;                          // TAPIR_CUDA_RH
    int trace = 0;           // TAPIR_CUDA_TRACE=1: tapir_cuda_run prints the timeline of its chunks to stderr
    std::string last_error;
    Lane lanes[kLanes];
    cudaStream_t stream = nullptr;  // == lanes[0].stream; used by the element-wise / measurement entry points
    // every batch copy goes through ONE host-to-device and ONE device-to-host stream: copies of one direction are
    // serial on the link anyway, and a lane stream that alternates directions was seen to queue its next upload
    // behind other lanes' pending downloads (uploads then waited ~200 us for the download engine)
    cudaStream_t stream_h2d = nullptr, stream_d2h = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    DevBuf d_scratch_a, d_scratch_b, d_scratch_c, d_flush, d_counter;
    uint64_t bytes_h2d = 0, bytes_d2h = 0;
    // gather target (see tapir_cuda_gather_*): a buffer on the owner's GPU, local or mapped through CUDA IPC
    DevBuf d_gather;                  // owner only
    void* gather_mapped = nullptr;    // non-owner: base pointer returned by cudaIpcOpenMemHandle
    char* gather_slot_ptr = nullptr;  // where this rank's small block goes
    size_t gather_slot_bytes = 0, gather_total_bytes = 0;
    // layout of the batch being gathered when it is processed in chunks (tapir_cuda_run): section offsets of the
    // FULL batch and the first row of the chunk in flight; 0/unused for the staged single-batch path
    bool gather_chunked = false;
    size_t gather_full_off[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};

namespace {

int fail(tapir_cuda_ctx* ctx, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) ctx->last_error = buf;
    return code;
}

#define CU(ctx, call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess) {                                                                          \
            cudaGetLastError();                                                                            \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? TAPIR_E_NOMEM : TAPIR_E_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e__));                                                          \
        }                                                                                                  \
    } while (0)

double now_us()
{
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int bind(tapir_cuda_ctx* ctx)
{
    if (!ctx) return fail(nullptr, TAPIR_E_INVALID, "null context");
    CU(ctx, cudaSetDevice(ctx->device));
    return TAPIR_OK;
}

// Tile-size model: time of one locus' profile per site (ps), a + b / sites per time tile, for the direct form
// (profile_kernel<RH>: TT = 256 RH points per tile) and the moment form (profile_mom_kernel<RH>, RH >= 4: a quarter of
// the FMAs and cheaper tables, but a moments -> times product per item that its sites share).  `sites` = mean sites
// per locus.  The constants are fitted to the many-tile rows of profiles/r02_tile_sweep.txt (B200, loci of 100 / 500 /
// 2,000 sites, grids of 200 .. 5,000 points; one-tile rows carry the same ~20 % tail of a 2-round launch in every variant).
double tile_cost(int rh, bool mom, int64_t T, double sites)
{
    static const double a_dir[4] = {19.8, 29.4, 52.0, 99.0}, b_dir[4] = {116.0, 630.0, 1010.0, 2160.0};
    static const double a_mom[4] = {0.0, 0.0, 24.0, 44.7}, b_mom[4] = {0.0, 0.0, 4500.0, 10480.0};
    const int64_t TT = 32 * 8 * rh;
    const int64_t nt = (T + TT - 1) / TT;
    const int i = rh == 1 ? 0 : rh == 2 ? 1 : rh == 4 ? 2 : 3;
    const double s = std::max(sites, 1.0);
    return double(nt) * (mom ? a_mom[i] + b_mom[i] / s : a_dir[i] + b_dir[i] / s);
}

void choose_tile(const tapir_cuda_ctx* ctx, const tapir_problem* p, int64_t T, double sites, int& rh_out, bool& mom_out);

template <int RH>
cudaError_t configure_profile(tapir_cuda_ctx* ctx, int idx)
{
    using Cfg = ProfCfg<RH>;
    cudaError_t e = cudaFuncSetAttribute(profile_kernel<RH>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SmemBytes));
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->prof_per_sm[idx], profile_kernel<RH>, kProfWarps * 32, Cfg::SmemBytes);
    if (e == cudaSuccess && ctx->prof_trace)
        e = cudaFuncSetAttribute(profile_kernel<RH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SmemBytes));
    if constexpr (RH >= 4) {
        using Mom = MomCfg<RH>;
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(profile_mom_kernel<RH>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Mom::SmemBytes));
        if (e == cudaSuccess)
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->mom_per_sm[idx], profile_mom_kernel<RH>, kProfWarps * 32, Mom::SmemBytes);
        if (e == cudaSuccess && ctx->prof_trace)
            e = cudaFuncSetAttribute(profile_mom_kernel<RH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Mom::SmemBytes));
        using Mom16 = MomCfg<RH, 16>;
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(profile_mom_kernel<RH, false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Mom16::SmemBytes));
        if (e == cudaSuccess)
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->mom16_per_sm[idx], profile_mom_kernel<RH, false, 16>, kProfWarps * 32, Mom16::SmemBytes);
        if (e == cudaSuccess && ctx->prof_trace)
            e = cudaFuncSetAttribute(profile_mom_kernel<RH, true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Mom16::SmemBytes));
    }
    return e;
}

int rh_index(int rh) { return rh == 1 ? 0 : rh == 2 ? 1 : rh == 4 ? 2 : 3; }

template <int RH>
cudaError_t launch_profile(tapir_cuda_ctx* ctx, Lane& ln, const ProfArgs& pa, int64_t n_items, bool mom, int km)
{
    using Cfg = ProfCfg<RH>;
    const int idx = rh_index(RH);
    int per_sm = std::max(1, mom ? (km == 16 ? ctx->mom16_per_sm[idx] : ctx->mom_per_sm[idx]) : ctx->prof_per_sm[idx]);
    if (ctx->prof_ctas_cap >= 1 && ctx->prof_ctas_cap < per_sm) per_sm = ctx->prof_ctas_cap;  // TAPIR_PROF_CTAS_PER_SM
    const int64_t need = (n_items + kProfWarps - 1) / kProfWarps;
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(need, int64_t(ctx->sm_count) * per_sm));
    if constexpr (RH >= 4) {
        if (mom && km == 16) {
            using Mom16 = MomCfg<RH, 16>;
            if (pa.trace)
                profile_mom_kernel<RH, true, 16><<<unsigned(grid), kProfWarps * 32, Mom16::SmemBytes, ln.stream>>>(pa);
            else
                profile_mom_kernel<RH, false, 16><<<unsigned(grid), kProfWarps * 32, Mom16::SmemBytes, ln.stream>>>(pa);
            return cudaGetLastError();
        }
        if (mom) {
            using Mom = MomCfg<RH>;
            if (pa.trace)
                profile_mom_kernel<RH, true><<<unsigned(grid), kProfWarps * 32, Mom::SmemBytes, ln.stream>>>(pa);
            else
                profile_mom_kernel<RH><<<unsigned(grid), kProfWarps * 32, Mom::SmemBytes, ln.stream>>>(pa);
            return cudaGetLastError();
        }
    }
    if (pa.trace)  // measurement only
        profile_kernel<RH, true><<<unsigned(grid), kProfWarps * 32, Cfg::SmemBytes, ln.stream>>>(pa);
    else
        profile_kernel<RH><<<unsigned(grid), kProfWarps * 32, Cfg::SmemBytes, ln.stream>>>(pa);
    return cudaGetLastError();
}

// Epoch integrals ride along: with the moment form the epochs mom_epoch.h serves cost a streaming scan of the rates
// (~6 ps per site), every other (site, epoch) ~10 ps in quad_kernel.
void choose_tile(const tapir_cuda_ctx* ctx, const tapir_problem* p, int64_t T, double sites, int& rh_out, bool& mom_out)
{
    rh_out = 1;
    mom_out = false;
    double best = 1e300;
    for (int mom = 0; mom <= (ctx->prof_mom ? 1 : 0); ++mom)
        for (int rh : {1, 2, 4, 8}) {
            if (mom && rh < 4) continue;
            const int64_t TT = 256 * rh, span = ((T + TT - 1) / TT) * TT;
            int served = 0;
            if (mom && ctx->mom_int)
                for (int e = 0; e < p->n_epochs; ++e) {
                    const double a = p->epoch_ab[2 * e], b = p->epoch_ab[2 * e + 1], h = 0.5 * (b - a);
                    if (a >= 0.0 && h >= 0.5 && h <= 32.0 && 0.5 * (a + b) < double(span)) ++served;
                }
            const double c = tile_cost(rh, mom != 0, T, sites) + 10.3 * (p->n_epochs - served) + (served ? 6.0 : 0.0);
            if (c < best) { best = c; rh_out = rh; mom_out = mom != 0; }
        }
    if (ctx->force_rh) {  // TAPIR_CUDA_RH: 1, 2, 4 or 8
        rh_out = ctx->force_rh;
        mom_out = ctx->prof_mom && rh_out >= 4;
    }
}

// Moments per 128-point block for the moment form: 32 (class 64 c <= 5) or 16 (class 64 c <= 1.25: half the work per
// slow site and per item, but every site between the two bounds goes through the power-table pass of the fast sites,
// and through quad_kernel for its epoch integrals).  Decided per batch from a strided sample of its rates: expected
// FP64-pipe cycles per (locus, tile) item -- the fast groups weighted 2.5x for their exposed latency -- plus what the
// per-site integrals of the sites outside the class cost in the same unit (~0.17 ns = 135 cycles each through the
// chunked quad_kernel).  On config 5 the model says 21.1 k against 24.9 k cycles per locus for 16 against 32 moments,
// the measurement 14.1 against 16.8 ms per step (profiles/r02_moments_sweep.txt).
int choose_moments(const tapir_cuda_ctx* ctx, const tapir_problem* p, int64_t l0, int64_t l1, int rh, int n_tiles)
{
    if (ctx->force_terms) return ctx->force_terms;
    const int64_t s0 = p->offsets[l0], s1 = p->offsets[l1];
    const int64_t nnz = s1 - s0, L = l1 - l0;
    if (nnz <= 0 || L <= 0) return 32;
    // 128 runs of 64 consecutive rates spread over the batch: 8,192 samples from ~1,000 cache lines of the (pinned) host
    // array -- a strided sample of the same size costs a cache miss per rate, a millisecond on a call that takes one
    const int64_t run_len = 64, n_runs = std::min<int64_t>(128, std::max<int64_t>(1, nnz / run_len));
    int64_t n_sample = 0, n_act = 0, n_ge16 = 0, n_ge32 = 0;
    for (int64_t k = 0; k < n_runs; ++k) {
        const int64_t b0 = s0 + int64_t(double(k) * double(nnz - std::min(nnz, run_len)) / double(std::max<int64_t>(1, n_runs - 1)));
        for (int64_t i = b0; i < std::min(s1, b0 + run_len); ++i) {
            const double r = p->rates[i];
            ++n_sample;
            const double a = 16.0 * (r * r);
            if (!(std::fabs(a) <= DBL_MAX) || a == 0.0) continue;
            ++n_act;
            if (!(r >= 0.0 && r < mom_rate_bound(16))) ++n_ge16;
            if (!(r >= 0.0 && r < mom_rate_bound(32))) ++n_ge32;
        }
    }
    if (n_act == 0) return 32;
    const double sites = double(nnz) / double(L);                              // sites per locus
    const double per_item = sites * double(n_act) / double(n_sample);         // of them with a non-zero finite rate
    const int NT = rh / 4, NB = 8 * NT;
    auto cost = [&](int km, int64_t n_ge) {
        const double n_fast = per_item * double(n_ge) / double(n_act), n_slow = per_item - n_fast;
        const double groups16 = std::max(1.0 - std::exp(-n_fast), (n_fast + std::sqrt(n_fast)) / 16.0);
        const double fast = groups16 * 2.5 * (256.0 * rh + 220.0) * (n_fast > 24.0 ? 1.2 : 1.0);
        const double profile = n_slow * (0.5 * km * NT + (33.0 + km + NB) / 16.0) + 64.0 * km * NT + fast;
        const double quad = p->n_epochs > 0 ? 135.0 * n_fast : 0.0;   // per-site integrals of the sites outside the class
        return double(n_tiles) * profile + quad;
    };
    return cost(16, n_ge16) < cost(32, n_ge32) ? 16 : 32;
}

void fill_batches(Plan& plan, const double* epoch_ab)
{
    plan.batches.clear();
    const int E = plan.E;
    std::vector<int> order(E);
    std::iota(order.begin(), order.end(), 0);
    // group equal half-lengths so the e^{+/- c h x_j} tables are shared
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        return 0.5 * (epoch_ab[2 * x + 1] - epoch_ab[2 * x]) < 0.5 * (epoch_ab[2 * y + 1] - epoch_ab[2 * y]);
    });
    for (int start = 0; start < E; start += kEpochsPerLaunch) {
        EpochBatch eb;
        memset(&eb, 0, sizeof eb);
        eb.n = std::min(kEpochsPerLaunch, E - start);
        eb.first = (start == 0);
        eb.e_total = E;
        eb.nonneg = 1;
        eb.bmax = 0.0;
        for (int k = 0; k < eb.n; ++k) {
            const int e = order[start + k];
            const double a = epoch_ab[2 * e], b = epoch_ab[2 * e + 1];
            eb.col[k] = e;
            eb.a[k] = a;
            eb.b[k] = b;
            // dqk21: centr = 0.5*(a+b), hlgth = 0.5*(b-a), absc = hlgth*xgk[j], nodes centr -/+ absc
            eb.center[k] = 0.5 * (a + b);
            eb.half[k] = 0.5 * (b - a);
            for (int j = 0; j < 10; ++j) {
                const double absc = eb.half[k] * h_xgk[j];
                eb.tm[k][j] = eb.center[k] - absc;
                eb.tp[k][j] = eb.center[k] + absc;
            }
            eb.new_half[k] = (k == 0) || (eb.half[k] != eb.half[k - 1]);
            if (k == 0) {
                eb.mom_all = 1;
                eb.n_win = 0;
                eb.scan_lo = 0u;
                eb.mom_terms = plan.mom_terms;
                eb.class_hi = plan.fp32 ? mom_rate_bound_hi_x(kMomClassF32) : mom_rate_bound_hi(plan.mom_terms);
                eb.scan_hi = eb.class_hi;
            }
            if (!(size_t(e) < plan.mom.size() && plan.mom[e].eligible)) eb.mom_all = 0;
            if (size_t(e) < plan.mom.size() && plan.mom[e].eligible) {
                const MomEpochPlan& mp = plan.mom[e];
                auto bits = [](double v) { unsigned long long u; memcpy(&u, &v, 8); return u; };
                eb.mom_any = 1;
                eb.mom[k] = 1;
                eb.rmin_bits[k] = bits(mp.r_min);
                eb.rmax_bits[k] = bits(mp.r_max);
                eb.wlo_bits[k] = bits(mp.w_lo);
                eb.whi_bits[k] = bits(mp.w_hi);
                eb.tb[k] = mp.tb;
                // high words for the scan: strictly inside the range, windows widened to whole high words
                eb.scan_lo = std::max(eb.scan_lo, uint32_t(eb.rmin_bits[k] >> 32));
                eb.scan_hi = std::min(eb.scan_hi, uint32_t(eb.rmax_bits[k] >> 32));
                if (mp.w_lo <= mp.w_hi) {
                    eb.win_lo[eb.n_win] = uint32_t(eb.wlo_bits[k] >> 32);
                    eb.win_hi[eb.n_win] = uint32_t(eb.whi_bits[k] >> 32);
                    ++eb.n_win;
                }
            }
            if (a < 0.0) eb.nonneg = 0;
            eb.bmax = std::max(eb.bmax, std::max(std::fabs(a), std::fabs(b)));
        }
        if (eb.scan_hi <= eb.scan_lo + 1u) eb.mom_all = 0;   // empty common range: nothing is served for every epoch
        plan.batches.push_back(eb);
    }
    if (E == 0) {  // still need the rate statistics
        EpochBatch eb;
        memset(&eb, 0, sizeof eb);
        eb.first = 1;
        eb.nonneg = 1;
        plan.batches.push_back(eb);
    }
}

int validate(tapir_cuda_ctx* ctx, const tapir_problem* p)
{
    if (!p) return fail(ctx, TAPIR_E_INVALID, "null problem");
    if (p->n_loci < 0 || p->n_grid < 0 || p->n_times < 0 || p->n_epochs < 0)
        return fail(ctx, TAPIR_E_INVALID, "negative size");
    if (p->n_loci > 0 && !p->offsets) return fail(ctx, TAPIR_E_INVALID, "offsets is NULL");
    if (p->n_loci >= (int64_t(1) << 31)) return fail(ctx, TAPIR_E_INVALID, "too many loci for one batch");
    if (p->n_epochs > TAPIR_MAX_EPOCHS) return fail(ctx, TAPIR_E_INVALID, "more than %d epochs", TAPIR_MAX_EPOCHS);
    if (p->n_times > TAPIR_MAX_TIMES) return fail(ctx, TAPIR_E_INVALID, "more than %d times", TAPIR_MAX_TIMES);
    if (p->n_epochs > 0 && !p->epoch_ab) return fail(ctx, TAPIR_E_INVALID, "epoch_ab is NULL");
    if (p->n_times > 0 && !p->time_idx) return fail(ctx, TAPIR_E_INVALID, "time_idx is NULL");
    if (p->n_loci > 0) {
        if (p->offsets[0] != 0) return fail(ctx, TAPIR_E_INVALID, "offsets[0] != 0");
        for (int64_t l = 0; l < p->n_loci; ++l)
            if (p->offsets[l + 1] < p->offsets[l]) return fail(ctx, TAPIR_E_INVALID, "offsets decrease at locus %lld", (long long)l);
        if (p->offsets[p->n_loci] > 0 && !p->rates) return fail(ctx, TAPIR_E_INVALID, "rates is NULL");
    }
    for (int e = 0; e < p->n_epochs; ++e) {
        const double a = p->epoch_ab[2 * e], b = p->epoch_ab[2 * e + 1];
        if (!(a < b) || !std::isfinite(a) || !std::isfinite(b))
            // compute.py:90-91 asserts span[0] < span[1]
            return fail(ctx, TAPIR_E_INVALID, "Start time [%g] is sooner than end time [%g]", a, b);
    }
    for (int k = 0; k < p->n_times; ++k) {
        const int64_t t = p->time_idx[k];
        if (t >= p->n_grid || t < -p->n_grid)  // numpy fancy-index bounds (compute.py:78)
            return fail(ctx, TAPIR_E_INVALID, "index %lld is out of bounds for axis 0 with size %lld", (long long)t,
                        (long long)p->n_grid);
    }
    return TAPIR_OK;
}

}  // namespace

extern "C" {

int tapir_cuda_version(void) { return TAPIR_CUDA_ABI_VERSION; }

int tapir_cuda_device_count(int* count)
{
    if (!count) return fail(nullptr, TAPIR_E_INVALID, "count is NULL");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return fail(nullptr, TAPIR_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return TAPIR_OK;
}

const char* tapir_cuda_last_error(const tapir_cuda_ctx* ctx)
{
    if (ctx && !ctx->last_error.empty()) return ctx->last_error.c_str();
    return g_last_error.c_str();
}

int tapir_cuda_create(int device, tapir_cuda_ctx** out)
{
    if (!out) return fail(nullptr, TAPIR_E_INVALID, "ctx out pointer is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, TAPIR_E_CUDA, "no CUDA device available (%s); libtapir_cuda has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= n) return fail(nullptr, TAPIR_E_INVALID, "device %d out of range [0,%d)", device, n);
    tapir_cuda_ctx* ctx = new (std::nothrow) tapir_cuda_ctx();
    if (!ctx) return fail(nullptr, TAPIR_E_NOMEM, "out of host memory");
    ctx->device = device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        delete ctx;
        cudaGetLastError();
        return fail(nullptr, TAPIR_E_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    }
    if (prop.major < 10) {
        delete ctx;
        return fail(nullptr, TAPIR_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    }
    ctx->sm_count = prop.multiProcessorCount;
    {
        const char* env = getenv("TAPIR_CUDA_SERIAL_KERNELS");
        ctx->serial_kernels = (env && env[0] == '1') ? 1 : 0;
        env = getenv("TAPIR_CUDA_GRAPHS");
        ctx->use_graphs = (env && env[0] == '0') ? 0 : 1;
        if ((env = getenv("TAPIR_CUDA_LANES"))) ctx->n_lanes = std::max(1, std::min(kLanes, atoi(env)));
        if ((env = getenv("TAPIR_CUDA_RAMP"))) ctx->ramp = atoi(env);
        if ((env = getenv("TAPIR_CUDA_TRACE"))) ctx->trace = atoi(env) != 0;
        if ((env = getenv("TAPIR_PROF_MOM"))) ctx->prof_mom = atoi(env) != 0;
        if ((env = getenv("TAPIR_MOM_INT"))) ctx->mom_int = atoi(env) != 0;
        if ((env = getenv("TAPIR_PROF_CTAS_PER_SM"))) ctx->prof_ctas_cap = atoi(env);
        if ((env = getenv("TAPIR_QUAD_CTAS_PER_SM"))) ctx->quad_ctas_cap = atoi(env);
        if ((env = getenv("TAPIR_FP32_MODE"))) ctx->force_fp32 = atoi(env) != 0;
        if ((env = getenv("TAPIR_MOM_TERMS"))) {
            const int v = atoi(env);
            if (v == 16 || v == 32) ctx->force_terms = v;
        }
        if ((env = getenv("TAPIR_CUDA_RH"))) {
            const int v = atoi(env);
            if (v == 1 || v == 2 || v == 4 || v == 8) ctx->force_rh = v;
        }
        if ((env = getenv("TAPIR_PROF_TRACE")) && env[0]) {
            ctx->prof_trace_path = env;
            ctx->prof_trace_words = size_t(prop.multiProcessorCount) * 32 * 32;
            if (cudaMalloc(&ctx->prof_trace, ctx->prof_trace_words * 8) != cudaSuccess) ctx->prof_trace = nullptr;
            else cudaMemset(ctx->prof_trace, 0, ctx->prof_trace_words * 8);
            ctx->use_graphs = 0;
        }
    }
    {
        std::vector<double> tab(kMtPowDoubles);
        mt_pow_table(tab.data());
        e = cudaMalloc(&ctx->d_mt_pow, tab.size() * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpy(ctx->d_mt_pow, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice);
    }
    // launch configuration of the profile kernels on this device (shared-memory opt-in, resident CTAs per SM)
    if (e == cudaSuccess) e = configure_profile<1>(ctx, 0);
    if (e == cudaSuccess) e = configure_profile<2>(ctx, 1);
    if (e == cudaSuccess) e = configure_profile<4>(ctx, 2);
    if (e == cudaSuccess) e = configure_profile<8>(ctx, 3);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->f32_per_sm, profile_f32_kernel, 32, 0);
    for (Lane& ln : ctx->lanes) {
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ln.stream_q, cudaStreamNonBlocking);
        for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&ln.ev[i]);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ln.ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ln.ev_join, cudaEventDisableTiming);
        const unsigned evf = ctx->trace ? cudaEventDefault : cudaEventDisableTiming;
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ln.ev_in, evf);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ln.ev_out, evf);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ln.ev_done, evf);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ln.ev_dstart, evf);
    }
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream_h2d, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream_d2h, cudaStreamNonBlocking);
    ctx->stream = ctx->lanes[0].stream;
    if (e != cudaSuccess) {
        tapir_cuda_destroy(ctx);
        cudaGetLastError();
        return fail(nullptr, TAPIR_E_CUDA, "stream/event creation: %s", cudaGetErrorString(e));
    }
    *out = ctx;
    return TAPIR_OK;
}

int tapir_cuda_destroy(tapir_cuda_ctx* ctx)
{
    if (!ctx) return TAPIR_OK;
    cudaSetDevice(ctx->device);
    for (Lane& ln : ctx->lanes) {
        if (ln.stream) cudaStreamSynchronize(ln.stream);
        if (ln.stream_q) cudaStreamSynchronize(ln.stream_q);
        for (DevBuf* b : ln.all_bufs) b->release();
        ln.h_stage.release();
        ln.h_small.release();
        for (auto& ev : ln.ev)
            if (ev) cudaEventDestroy(ev);
        if (ln.graph_exec) cudaGraphExecDestroy(ln.graph_exec);
        if (ln.ev_fork) cudaEventDestroy(ln.ev_fork);
        if (ln.ev_join) cudaEventDestroy(ln.ev_join);
        if (ln.ev_in) cudaEventDestroy(ln.ev_in);
        if (ln.ev_out) cudaEventDestroy(ln.ev_out);
        if (ln.ev_done) cudaEventDestroy(ln.ev_done);
        if (ln.ev_dstart) cudaEventDestroy(ln.ev_dstart);
        if (ln.stream) cudaStreamDestroy(ln.stream);
        if (ln.stream_q) cudaStreamDestroy(ln.stream_q);
    }
    if (ctx->stream_h2d) { cudaStreamSynchronize(ctx->stream_h2d); cudaStreamDestroy(ctx->stream_h2d); }
    if (ctx->stream_d2h) { cudaStreamSynchronize(ctx->stream_d2h); cudaStreamDestroy(ctx->stream_d2h); }
    if (ctx->gather_mapped) cudaIpcCloseMemHandle(ctx->gather_mapped);
    if (ctx->prof_trace) cudaFree(ctx->prof_trace);
    if (ctx->d_mt_pow) cudaFree(ctx->d_mt_pow);
    DevBuf* bufs[] = {&ctx->d_scratch_a, &ctx->d_scratch_b, &ctx->d_scratch_c, &ctx->d_flush, &ctx->d_counter, &ctx->d_gather};
    for (DevBuf* b : bufs) b->release();
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    cudaGetLastError();
    delete ctx;
    return TAPIR_OK;
}

int tapir_cuda_host_alloc(void** ptr, size_t bytes)
{
    if (!ptr) return fail(nullptr, TAPIR_E_INVALID, "ptr is NULL");
    *ptr = nullptr;
    if (bytes == 0) bytes = 1;
    cudaError_t e = cudaMallocHost(ptr, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, TAPIR_E_NOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
    }
    return TAPIR_OK;
}

int tapir_cuda_host_free(void* ptr)
{
    if (!ptr) return TAPIR_OK;
    cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, TAPIR_E_CUDA, "cudaFreeHost: %s", cudaGetErrorString(e));
    }
    return TAPIR_OK;
}

}  // extern "C"

namespace {

void drop_graph(Lane& ln)
{
    if (ln.graph_exec) cudaGraphExecDestroy(ln.graph_exec);
    ln.graph_exec = nullptr;
    ln.graph_launches = 0;
}

// Build the work lists of loci [l0, l1) of an already validated problem on lane `ln` and start their H2D copies.
// Which profile kernel a call runs: decided ONCE per call from the whole batch (tile size and form from the grid and
// the mean locus, moments per block from a sample of the rates), so that every chunk of a pipelined tapir_cuda_run
// uses the same kernel -- results do not depend on how a batch was cut -- and the sampling is not paid per chunk.
struct Decision {
    int rh = 1;
    bool mom_form = false;
    int mom_terms = 32;
    bool fp32 = false;
};

Decision decide_batch(const tapir_cuda_ctx* ctx, const tapir_problem* p)
{
    Decision d;
    const int64_t L = p->n_loci, nnz = L ? p->offsets[L] : 0, T = p->n_grid;
    choose_tile(ctx, p, std::max<int64_t>(T, 1), L ? double(nnz) / double(L) : 1.0, d.rh, d.mom_form);
    d.fp32 = ((p->flags & TAPIR_FP32_MODE) || ctx->force_fp32) && T > 0 && L > 0 && !ctx->prof_trace;
    if (d.fp32) {   // profile_f32_kernel: 1,024-point tiles, moment form
        d.rh = 4;
        d.mom_form = true;
    }
    const int64_t TT = 256 * d.rh;
    const int n_tiles = int(std::max<int64_t>(1, (T + TT - 1) / TT));
    d.mom_terms = d.mom_form ? choose_moments(ctx, p, 0, L, d.rh, n_tiles) : 32;
    if (d.fp32) d.mom_terms = kF32Terms;
    return d;
}

int lane_upload(tapir_cuda_ctx* ctx, Lane& ln, const tapir_problem* p, int64_t l0, int64_t l1, bool sync, const Decision& dec)
{
    ln.plan.valid = false;
    drop_graph(ln);
    Plan& plan = ln.plan;
    const int64_t* off = &p->offsets[l0];       // off[0..L]; site indices are taken relative to off[0]
    const int64_t base = (l1 > l0) ? off[0] : 0;
    plan.L = l1 - l0;
    plan.T = p->n_grid;
    plan.K = p->n_times;
    plan.E = p->n_epochs;
    plan.flags = p->flags;
    plan.nnz = plan.L ? off[plan.L] - base : 0;
    plan.rh = dec.rh;
    plan.mom_form = dec.mom_form;
    plan.fp32 = dec.fp32 && plan.L > 0;
    plan.TT = 32 * 8 * plan.rh;
    plan.n_tiles = int32_t(std::max<int64_t>(1, (plan.T + plan.TT - 1) / plan.TT));
    // epochs whose sums come out of the profile's moments (mom_epoch.h); the others, and the sites outside an
    // epoch's rate range, are integrated per site by quad_kernel
    plan.mom_terms = dec.mom_terms;
    const double class_x = plan.fp32 ? kMomClassF32 : mom_class_x(plan.mom_terms);
    plan.mom.assign(size_t(plan.E), MomEpochPlan());
    plan.n_mom = 0;
    plan.mom_mask = 0;
    if (ctx->mom_int && plan.mom_form && plan.T > 0 && plan.L > 0)
        for (int e = 0; e < plan.E; ++e) {
            const double ea = p->epoch_ab[2 * e], eb_ = p->epoch_ab[2 * e + 1];
            const tapir_cuda_ctx::MomCacheEntry* hit = nullptr;
            for (const auto& ce : ctx->mom_cache)
                if (ce.a == ea && ce.b == eb_ && ce.n_tiles == plan.n_tiles && ce.TT == plan.TT && ce.km == plan.mom_terms &&
                    ce.class_x == class_x) { hit = &ce; break; }
            if (!hit) {
                if (ctx->mom_cache.size() >= 4096) ctx->mom_cache.clear();
                tapir_cuda_ctx::MomCacheEntry ce{ea, eb_, plan.n_tiles, plan.TT, plan.mom_terms, class_x, MomEpochPlan()};
                mom_epoch_plan(ea, eb_, plan.n_tiles, plan.TT, plan.mom_terms, &ce.plan, class_x);
                ctx->mom_cache.push_back(ce);
                hit = &ctx->mom_cache.back();
            }
            plan.mom[size_t(e)] = hit->plan;
            if (plan.mom[size_t(e)].eligible) {
                ++plan.n_mom;
                plan.mom_mask |= uint64_t(1) << e;
            }
        }
    fill_batches(plan, p->epoch_ab);
    const int64_t L = plan.L;

    // ---- site splits: only when there are too few (locus, tile) pairs to fill the machine ----
    const int64_t warp_slots = int64_t(ctx->sm_count) * 12;
    const int64_t target_items = 8 * warp_slots;
    const int64_t pairs = L * plan.n_tiles;
    int64_t chunk_sites = INT32_MAX;
    if (pairs > 0 && pairs < target_items && plan.nnz > 0) {
        const double total_site_tiles = double(plan.nnz) * plan.n_tiles;
        chunk_sites = int64_t(total_site_tiles / double(target_items)) + 1;
        chunk_sites = std::max<int64_t>(1024, ((chunk_sites + 31) / 32) * 32);
    }
    std::vector<int32_t> nsplit(L);
    std::vector<int64_t> ws_base(L), qitem_begin(L + 1), part_base(L + 1);
    int64_t n_prof = 0, n_quad = 0, ws_rows = 0, n_parts = 0;
    for (int64_t l = 0; l < L; ++l) {
        const int64_t s = off[l + 1] - off[l];
        if (s > INT32_MAX) return fail(ctx, TAPIR_E_INVALID, "locus %lld has too many sites", (long long)l);
        int64_t ns = (s + chunk_sites - 1) / chunk_sites;
        if (ns < 1) ns = 1;
        nsplit[l] = int32_t(ns);
        part_base[l] = n_parts;
        n_parts += ns;
        ws_base[l] = ws_rows;
        if (ns > 1) ws_rows += ns * plan.n_tiles;
        n_prof += ns * plan.n_tiles;
        qitem_begin[l] = n_quad;
        n_quad += (s + kQuadBlock - 1) / kQuadBlock;
    }
    qitem_begin[L] = n_quad;
    part_base[L] = n_parts;
    plan.n_parts = n_parts;
    if (n_prof >= (int64_t(1) << 31) || n_quad >= (int64_t(1) << 31) || n_parts >= (int64_t(1) << 31))
        return fail(ctx, TAPIR_E_INVALID, "batch too large: split it into several calls");
    plan.n_prof_items = n_prof;
    plan.n_quad_items = n_quad;
    plan.ws_rows = ws_rows;

    // ---- stage work lists in pinned memory ----
    const size_t b_prof = size_t(n_prof) * sizeof(ProfItem), b_quad = size_t(n_quad) * sizeof(QuadItem);
    const size_t b_ns = size_t(L) * sizeof(int32_t), b_wb = size_t(L) * sizeof(int64_t);
    const size_t b_qb = size_t(L + 1) * sizeof(int64_t), b_ti = size_t(plan.K) * sizeof(int64_t);
    const size_t b_pb = size_t(L + 1) * sizeof(int64_t), b_me = size_t(plan.n_mom) * sizeof(MomEpochDev);
    const size_t b_mp = size_t(plan.E) * kMomTerms * sizeof(double);
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t o_prof = 0, o_quad = o_prof + al(b_prof), o_ns = o_quad + al(b_quad), o_wb = o_ns + al(b_ns),
                 o_qb = o_wb + al(b_wb), o_ti = o_qb + al(b_qb), o_pb = o_ti + al(b_ti), o_me = o_pb + al(b_pb),
                 o_mp = o_me + al(b_me), total = o_mp + al(b_mp);
    CU(ctx, ln.h_stage.ensure(total));
    char* hs = ln.h_stage.as<char>();
    ProfItem* pi = reinterpret_cast<ProfItem*>(hs + o_prof);
    QuadItem* qi = reinterpret_cast<QuadItem*>(hs + o_quad);
    {
        int64_t k = 0, q = 0;
        bool ragged = false;
        for (int64_t l = 0; l < L; ++l) {
            const int64_t s0 = off[l] - base, s = off[l + 1] - off[l];
            if (l && s != off[1] - off[0]) ragged = true;
            const int64_t ns = nsplit[l];
            // equal site ranges rounded to multiples of 32
            const int64_t per = ns > 1 ? (((s + ns - 1) / ns + 31) / 32) * 32 : s;
            for (int tile = 0; tile < plan.n_tiles; ++tile)
                for (int64_t j = 0; j < ns; ++j) {
                    const int64_t b0 = std::min(s, j * per), b1 = std::min(s, (j + 1) * per);
                    ProfItem it;
                    it.site_begin = s0 + b0;
                    it.n_sites = int32_t(b1 - b0);
                    it.locus = int32_t(l);
                    it.tile = tile;
                    it.slot = ns > 1 ? int32_t(ws_base[l] + int64_t(tile) * ns + j) : -1;
                    it.part = int32_t(part_base[l] + j);
                    it.pad_ = 0;
                    pi[k++] = it;
                }
            for (int64_t b0 = 0; b0 < s; b0 += kQuadBlock) {
                QuadItem it;
                it.site_begin = s0 + b0;
                it.count = int32_t(std::min<int64_t>(kQuadBlock, s - b0));
                it.locus = int32_t(l);
                qi[q++] = it;
            }
        }
        if (ws_rows >= (int64_t(1) << 31)) return fail(ctx, TAPIR_E_INVALID, "split workspace too large");
        // longest-first so the dynamic queue balances ragged loci
        if (ragged) std::stable_sort(pi, pi + n_prof, [](const ProfItem& x, const ProfItem& y) { return x.n_sites > y.n_sites; });
    }
    memcpy(hs + o_ns, nsplit.data(), b_ns);
    memcpy(hs + o_wb, ws_base.data(), b_wb);
    memcpy(hs + o_qb, qitem_begin.data(), b_qb);
    memcpy(hs + o_pb, part_base.data(), b_pb);
    {
        MomEpochDev* me = reinterpret_cast<MomEpochDev*>(hs + o_me);
        double* mp = reinterpret_cast<double*>(hs + o_mp);
        int k = 0;
        for (int e = 0; e < plan.E; ++e) {
            const MomEpochPlan& m = plan.mom[size_t(e)];
            for (int j = 0; j < kMomTerms; ++j) mp[size_t(e) * kMomTerms + j] = m.eligible ? m.p[j] : 0.0;
            if (m.eligible) me[k++] = MomEpochDev{e, m.tile, m.hb, 0};
        }
    }
    {
        int64_t* ti = reinterpret_cast<int64_t*>(hs + o_ti);
        for (int k = 0; k < plan.K; ++k) ti[k] = p->time_idx[k] < 0 ? p->time_idx[k] + plan.T : p->time_idx[k];
    }

    // ---- device buffers ----
    const size_t qrow = size_t(2 * plan.E + 4);
    const size_t Ls = size_t(L), Ks = size_t(plan.K), Es = size_t(plan.E);
    const size_t small_sizes[9] = {Ls * Ks * 8, Ls * Es * 8, Ls * Es * 8, Ls * 8, Ls * 8, Ls * 8, Ls * 8, Ls * 8, Ls * 4};
    plan.small_bytes = 0;
    for (int i = 0; i < 9; ++i) {
        plan.small_off[i] = plan.small_bytes;
        plan.small_bytes += al(small_sizes[i]);
    }
    CU(ctx, ln.d_rates.ensure(size_t(plan.nnz) * 8 + 8));
    CU(ctx, ln.d_work.ensure(total + 8));
    CU(ctx, ln.d_counter.ensure(256));
    CU(ctx, ln.d_ws.ensure(size_t(ws_rows) * plan.TT * 8 + 8));
    CU(ctx, ln.d_tile_max.ensure(size_t(L) * plan.n_tiles * 8 + 8));
    CU(ctx, ln.d_tile_arg.ensure(size_t(L) * plan.n_tiles * 8 + 8));
    CU(ctx, ln.d_qpart.ensure(size_t(n_quad) * qrow * 8 + 8));
    CU(ctx, ln.d_fix.ensure(size_t(n_quad) * kEpochsPerLaunch * sizeof(FixEntry) + 8));
    CU(ctx, ln.d_mom_int.ensure(size_t(n_parts) * size_t(plan.E) * 8 + 8));
    CU(ctx, ln.d_rare_list.ensure(plan.n_mom ? size_t(n_quad) * plan.batches.size() * (kQuadBlock / 32) * 4 + 64 : 64));
    CU(ctx, ln.d_pi_net.ensure(size_t(L) * plan.T * 8 + 16));
    CU(ctx, ln.d_small.ensure(plan.small_bytes + 8));
    CU(ctx, ln.h_small.ensure(plan.small_bytes + 8));
    {
        char* w = ln.d_work.as<char>();
        ln.d_prof_items.p = w + o_prof;
        ln.d_quad_items.p = w + o_quad;
        ln.d_nsplit.p = w + o_ns;
        ln.d_ws_base.p = w + o_wb;
        ln.d_qitem_begin.p = w + o_qb;
        ln.d_time_idx.p = w + o_ti;
        ln.d_part_base.p = w + o_pb;
        ln.d_mom_epochs.p = w + o_me;
        ln.d_mom_p.p = w + o_mp;
        for (EpochBatch& eb : plan.batches) eb.mom_p = ln.d_mom_p.as<double>();
        char* r = ln.d_small.as<char>();
        DevView* views[9] = {&ln.d_pi_times, &ln.d_integral,  &ln.d_abserr,   &ln.d_max_pi,      &ln.d_argmax_t,
                             &ln.d_mean_rate, &ln.d_n_finite, &ln.d_n_multi, &ln.d_argmax_epoch};
        for (int i = 0; i < 9; ++i) views[i]->p = r + plan.small_off[i];
    }

    cudaStream_t st = ctx->stream_h2d;
    ctx->bytes_h2d += size_t(plan.nnz) * 8 + total;
    if (plan.nnz) CU(ctx, cudaMemcpyAsync(ln.d_rates.p, p->rates + base, size_t(plan.nnz) * 8, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(ln.d_work.p, hs, total, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaEventRecord(ln.ev_in, st));
    CU(ctx, cudaStreamWaitEvent(ln.stream, ln.ev_in, 0));  // the lane's kernels start when its inputs have landed
    if (sync) CU(ctx, cudaStreamSynchronize(st));
    plan.valid = true;
    return TAPIR_OK;
}

// Enqueue the kernels of the lane's resident batch on its stream(s).  Unless ctx->serial_kernels is set, the
// epoch-integral kernels go to the lane's second stream, concurrently with the profile kernel, and join before
// finalize.  Also what the CUDA graph of the staged path captures.
int enqueue_kernels(tapir_cuda_ctx* ctx, Lane& ln, bool inner_events, int& launches)
{
    Plan& plan = ln.plan;
    cudaStream_t st = ln.stream;
    const bool fork = !ctx->serial_kernels && plan.L > 0 && plan.n_quad_items > 0 && plan.n_prof_items > 0 && plan.T > 0;
    cudaStream_t sq = fork ? ln.stream_q : st;
    if (plan.L > 0) {
        CU(ctx, cudaMemsetAsync(ln.d_counter.p, 0, 256, st));
        if (fork) {
            CU(ctx, cudaEventRecord(ln.ev_fork, st));
            CU(ctx, cudaStreamWaitEvent(sq, ln.ev_fork, 0));
        }
        if (plan.n_prof_items > 0 && plan.T > 0) {

            ProfArgs pa;
            pa.rates = ln.d_rates.as<double>();
            pa.items = ln.d_prof_items.as<ProfItem>();
            pa.counter = ln.d_counter.as<uint32_t>();
            pa.pi_net = ln.d_pi_net.as<double>();
            pa.ws = ln.d_ws.as<double>();
            pa.tile_max = ln.d_tile_max.as<double>();
            pa.tile_arg = ln.d_tile_arg.as<int64_t>();
            pa.T = plan.T;
            pa.n_items = int32_t(plan.n_prof_items);
            pa.n_tiles = plan.n_tiles;
            pa.trace = ctx->prof_trace;
            pa.mom_epochs = ln.d_mom_epochs.as<MomEpochDev>();
            pa.mom_p = ln.d_mom_p.as<double>();
            pa.mom_int = ln.d_mom_int.as<double>();
            pa.n_mom_epochs = plan.mom_form ? plan.n_mom : 0;
            pa.e_total = plan.E;
            pa.mt_pow = ctx->d_mt_pow;
            cudaError_t e = cudaSuccess;
            if (plan.fp32) {
                const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(plan.n_prof_items, int64_t(ctx->sm_count) * std::max(1, ctx->f32_per_sm)));
                profile_f32_kernel<<<unsigned(grid), 32, 0, st>>>(pa);
                e = cudaGetLastError();
            } else
            switch (plan.rh) {
                case 1: e = launch_profile<1>(ctx, ln, pa, plan.n_prof_items, false, 32); break;
                case 2: e = launch_profile<2>(ctx, ln, pa, plan.n_prof_items, false, 32); break;
                case 4: e = launch_profile<4>(ctx, ln, pa, plan.n_prof_items, plan.mom_form, plan.mom_terms); break;
                default: e = launch_profile<8>(ctx, ln, pa, plan.n_prof_items, plan.mom_form, plan.mom_terms); break;
            }
            CU(ctx, e);
            ++launches;
        }
        if (inner_events) CU(ctx, cudaEventRecord(ln.ev[1], st));
        if (plan.n_quad_items > 0) {
            QuadArgs qa;
            qa.rates = ln.d_rates.as<double>();
            qa.items = ln.d_quad_items.as<QuadItem>();
            qa.qpart = ln.d_qpart.as<double>();
            qa.fix = ln.d_fix.as<FixEntry>();
            qa.n_items = int32_t(plan.n_quad_items);
            int k = 0;
            for (const EpochBatch& eb : plan.batches) {
                // two counters per batch (d_counter[1..]); slot 0 is the profile kernel's work queue head
                qa.fix_count = ln.d_counter.as<uint32_t>() + 1 + 2 * k;
                qa.queue = ln.d_counter.as<uint32_t>() + 2 + 2 * k;
                ++k;
                // two counters per batch in d_counter[1..16], the batch's rare-item count in d_counter[17 + k]
                qa.list = nullptr;
                qa.list_count = nullptr;
                if (eb.mom_any) {
                    qa.list = ln.d_rare_list.as<uint32_t>() + size_t(k - 1) * size_t(plan.n_quad_items) * (kQuadBlock / 32);
                    qa.list_count = ln.d_counter.as<uint32_t>() + 17 + (k - 1);
                    const int64_t warps = kScanThreads / 32;
                    const unsigned sgrid = unsigned(std::max<int64_t>(1, std::min<int64_t>((plan.n_quad_items + warps - 1) / warps, int64_t(ctx->sm_count) * 8)));
                    quad_scan_kernel<<<sgrid, kScanThreads, 0, sq>>>(qa, eb);
                    CU(ctx, cudaGetLastError());
                    ++launches;
                }
                const bool chunked = eb.mom_any != 0;   // candidate sites only, several items per work unit
                int q_per_sm = 12 / kQuadWarps;  // 12 warps per SM (168 registers)
                if (ctx->quad_ctas_cap >= 1 && ctx->quad_ctas_cap <= q_per_sm) q_per_sm = ctx->quad_ctas_cap;
                // unit size: as many items as keep every resident warp supplied with a few units
                qa.chunk = int32_t(std::max<int64_t>(1, std::min<int64_t>(kQuadChunk, plan.n_quad_items / (4 * int64_t(ctx->sm_count) * q_per_sm))));
                const int64_t want = chunked ? (plan.n_quad_items + qa.chunk - 1) / qa.chunk : plan.n_quad_items;
                const unsigned qgrid = unsigned(std::max<int64_t>(1, std::min<int64_t>(want, int64_t(ctx->sm_count) * q_per_sm)));
#define TAPIR_QUAD_LAUNCH(NE)                                                                    \
    if (chunked) quad_kernel<NE, kQuadChunk><<<qgrid, kQuadWarps * 32, 0, sq>>>(qa, eb);          \
    else quad_kernel<NE><<<qgrid, kQuadWarps * 32, 0, sq>>>(qa, eb)
                switch (eb.n) {
                    case 0: quad_kernel<0><<<qgrid, kQuadWarps * 32, 0, sq>>>(qa, eb); break;  // rate statistics only
                    case 1: TAPIR_QUAD_LAUNCH(1); break;
                    case 2: TAPIR_QUAD_LAUNCH(2); break;
                    case 3: TAPIR_QUAD_LAUNCH(3); break;
                    case 4: TAPIR_QUAD_LAUNCH(4); break;
                    case 5: TAPIR_QUAD_LAUNCH(5); break;
                    case 6: TAPIR_QUAD_LAUNCH(6); break;
                    case 7: TAPIR_QUAD_LAUNCH(7); break;
                    default: TAPIR_QUAD_LAUNCH(8); break;
                }
#undef TAPIR_QUAD_LAUNCH
                CU(ctx, cudaGetLastError());
                quad_fix_kernel<<<unsigned(ctx->sm_count * 4), 128, 0, sq>>>(qa, eb);
                CU(ctx, cudaGetLastError());
                launches += 2;
            }
        }
        if (fork) {
            CU(ctx, cudaEventRecord(ln.ev_join, sq));
            CU(ctx, cudaStreamWaitEvent(st, ln.ev_join, 0));
        }
        if (inner_events) CU(ctx, cudaEventRecord(ln.ev[2], st));
        {
            FinalArgs fa;
            fa.n_loci = plan.L;
            fa.T = plan.T;
            fa.TT = plan.TT;
            fa.n_tiles = plan.T > 0 ? plan.n_tiles : 0;
            fa.K = plan.T > 0 ? plan.K : 0;
            fa.E = plan.E;
            fa.nsplit = ln.d_nsplit.as<int32_t>();
            fa.ws_base = ln.d_ws_base.as<int64_t>();
            fa.qitem_begin = ln.d_qitem_begin.as<int64_t>();
            fa.time_idx = ln.d_time_idx.as<int64_t>();
            fa.ws = ln.d_ws.as<double>();
            fa.tile_max = ln.d_tile_max.as<double>();
            fa.tile_arg = ln.d_tile_arg.as<int64_t>();
            fa.qpart = ln.d_qpart.as<double>();
            fa.pi_net = ln.d_pi_net.as<double>();
            fa.pi_times = ln.d_pi_times.as<double>();
            fa.integral = ln.d_integral.as<double>();
            fa.abserr = ln.d_abserr.as<double>();
            fa.max_pi = ln.d_max_pi.as<double>();
            fa.argmax_t = ln.d_argmax_t.as<int64_t>();
            fa.argmax_epoch = ln.d_argmax_epoch.as<int32_t>();
            fa.mean_rate = ln.d_mean_rate.as<double>();
            fa.n_finite = ln.d_n_finite.as<int64_t>();
            fa.n_multi = ln.d_n_multi.as<int64_t>();
            fa.part_base = ln.d_part_base.as<int64_t>();
            fa.mom_int = ln.d_mom_int.as<double>();
            fa.mom_mask = (plan.mom_form && plan.T > 0 && plan.n_prof_items > 0) ? plan.mom_mask : 0;
            const int64_t blocks = (plan.L * 32 + 127) / 128;
            finalize_kernel<<<unsigned(blocks), 128, 0, st>>>(fa);
            CU(ctx, cudaGetLastError());
            ++launches;
        }
    } else {
        if (inner_events) CU(ctx, cudaEventRecord(ln.ev[1], st));
        if (inner_events) CU(ctx, cudaEventRecord(ln.ev[2], st));
    }
    if (ctx->gather_slot_ptr && plan.L > 0) {
        // NVLink push: the batch's small block goes into this rank's slot of the owner's buffer (peer memory),
        // by stores issued from this GPU on the same stream as finalize -- no collective launch
        const size_t K = size_t(plan.T > 0 ? plan.K : 0), E = size_t(plan.E);
        const size_t row_bytes[9] = {K * 8, E * 8, E * 8, 8, 8, 8, 8, 8, 4};
        PushArgs pa;
        pa.src = ln.d_small.as<uint32_t>();
        pa.dst = reinterpret_cast<uint32_t*>(ctx->gather_slot_ptr);
        size_t end = 0;
        for (int i = 0; i < 9; ++i) {
            const size_t dst_off = ctx->gather_chunked ? ctx->gather_full_off[i] + size_t(ln.pending_row0) * row_bytes[i]
                                                       : plan.small_off[i];
            pa.src_off[i] = uint32_t(plan.small_off[i] / 4);
            pa.dst_off[i] = uint32_t(dst_off / 4);
            pa.words[i] = uint32_t(size_t(plan.L) * row_bytes[i] / 4);
            end = std::max(end, dst_off + size_t(plan.L) * row_bytes[i]);
        }
        if (size_t(plan.L) * std::max(K, E) * 2 > size_t(UINT32_MAX))
            return fail(ctx, TAPIR_E_INVALID, "batch too large for the gather push: split it into several calls");
        if (end > ctx->gather_slot_bytes)
            return fail(ctx, TAPIR_E_STATE, "gather slot of %zu bytes is smaller than the batch's %zu", ctx->gather_slot_bytes, end);
        push_small_kernel<<<unsigned(std::min<int64_t>(ctx->sm_count, (plan.L * 16 + 255) / 256 + 1)), 256, 0, st>>>(pa);
        CU(ctx, cudaGetLastError());
        ++launches;
    }
    return TAPIR_OK;
}

// Run the kernels of the lane's resident batch.  The staged path (tapir_cuda_compute) repeats the same launch
// sequence on the same buffers for as long as the batch stays resident, so it is captured once into a CUDA graph
// (memset, fork, profile || integrals + fix-up, join, finalize, gather push) and replayed with a single launch.
int lane_compute(tapir_cuda_ctx* ctx, Lane& ln, tapir_timing* timing, bool sync, bool graph_ok)
{
    Plan& plan = ln.plan;
    if (!plan.valid) return fail(ctx, TAPIR_E_STATE, "tapir_cuda_compute called before a successful tapir_cuda_upload");
    cudaStream_t st = ln.stream;
    const bool use_graph = graph_ok && ctx->use_graphs && !ctx->serial_kernels && plan.L > 0;
    int launches = 0;
    if (timing) CU(ctx, cudaEventRecord(ln.ev[0], st));
    if (use_graph) {
        if (!ln.graph_exec) {
            CU(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
            int n = 0;
            const int rc = enqueue_kernels(ctx, ln, false, n);
            cudaGraph_t graph = nullptr;
            const cudaError_t e = cudaStreamEndCapture(st, &graph);
            if (rc != TAPIR_OK) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                return rc;
            }
            CU(ctx, e);
            const cudaError_t ei = cudaGraphInstantiate(&ln.graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            CU(ctx, ei);
            ln.graph_launches = n;
        }
        CU(ctx, cudaGraphLaunch(ln.graph_exec, st));
        launches = ln.graph_launches;
    } else {
        const int rc = enqueue_kernels(ctx, ln, timing != nullptr, launches);
        if (rc != TAPIR_OK) return rc;
    }
    if (timing) CU(ctx, cudaEventRecord(ln.ev[3], st));
    if (sync || timing) CU(ctx, cudaStreamSynchronize(st));
    if (timing) {
        memset(timing, 0, sizeof *timing);
        CU(ctx, cudaEventElapsedTime(&timing->total_ms, ln.ev[0], ln.ev[3]));
        if (!use_graph) {  // the per-kernel events are not part of the captured graph
            CU(ctx, cudaEventElapsedTime(&timing->profile_ms, ln.ev[0], ln.ev[1]));
            CU(ctx, cudaEventElapsedTime(&timing->quad_ms, ln.ev[1], ln.ev[2]));
            CU(ctx, cudaEventElapsedTime(&timing->finalize_ms, ln.ev[2], ln.ev[3]));
        }
        timing->launches = launches;
        timing->profile_tile_points = plan.TT;
        timing->profile_moment_form = plan.mom_form ? plan.mom_terms : 0;
        timing->moment_epochs = plan.mom_form ? plan.n_mom : 0;
        timing->profile_fp32 = plan.fp32 ? 1 : 0;
    }
    return TAPIR_OK;
}

// After the lane's D2H copy has landed in h_small: scatter the small vectors into the caller's arrays.
int lane_finish(tapir_cuda_ctx* ctx, Lane& ln)
{
    CU(ctx, cudaStreamSynchronize(ln.stream));
    if (!ln.pending) return TAPIR_OK;
    CU(ctx, cudaEventSynchronize(ln.ev_out));  // this lane's downloads (recorded on the download stream)
    ln.pending = false;
    const Plan& plan = ln.plan;
    const tapir_results& out = ln.pending_out;
    const size_t L = size_t(plan.L), r0 = size_t(ln.pending_row0);
    const size_t K = size_t(plan.T > 0 ? plan.K : 0), E = size_t(plan.E);
    const char* h = ln.h_small.as<char>();
    void* dst[9] = {out.pi_times, out.integral, out.abserr, out.max_pi, out.argmax_t, out.mean_rate, out.n_finite,
                    out.n_multi_panel, out.argmax_epoch};
    const size_t row_bytes[9] = {K * 8, E * 8, E * 8, 8, 8, 8, 8, 8, 4};
    for (int i = 0; i < 9; ++i)
        if (dst[i] && L * row_bytes[i]) memcpy(static_cast<char*>(dst[i]) + r0 * row_bytes[i], h + plan.small_off[i], L * row_bytes[i]);
    return TAPIR_OK;
}

// Start copying the lane's results to the host: the profile straight into rows [l0, l0 + L) of the caller's
// array, the small vectors as one block into pinned staging (lane_finish scatters them).
int lane_download(tapir_cuda_ctx* ctx, Lane& ln, const tapir_results* out, int64_t l0, bool sync)
{
    const Plan& plan = ln.plan;
    if (!plan.valid) return fail(ctx, TAPIR_E_STATE, "tapir_cuda_download called before upload/compute");
    if (!out) return fail(ctx, TAPIR_E_INVALID, "null results");
    cudaStream_t st = ctx->stream_d2h;
    CU(ctx, cudaEventRecord(ln.ev_done, ln.stream));  // after the lane's kernels
    CU(ctx, cudaStreamWaitEvent(st, ln.ev_done, 0));
    if (ctx->trace) CU(ctx, cudaEventRecord(ln.ev_dstart, st));
    const size_t L = size_t(plan.L);
    if ((plan.flags & TAPIR_WANT_PROFILE) && out->pi_net && L * size_t(plan.T)) {
        const size_t row = size_t(plan.T) * 8;
        ctx->bytes_d2h += L * row;
        CU(ctx, cudaMemcpyAsync(reinterpret_cast<char*>(out->pi_net) + size_t(l0) * row, ln.d_pi_net.p, L * row,
                                cudaMemcpyDeviceToHost, st));
    }
    if (plan.small_bytes) {
        ctx->bytes_d2h += plan.small_bytes;
        CU(ctx, cudaMemcpyAsync(ln.h_small.p, ln.d_small.p, plan.small_bytes, cudaMemcpyDeviceToHost, st));
    }
    CU(ctx, cudaEventRecord(ln.ev_out, st));
    ln.pending_out = *out;
    ln.pending_row0 = l0;
    ln.pending = true;
    return sync ? lane_finish(ctx, ln) : TAPIR_OK;
}

}  // namespace

extern "C" {

int tapir_cuda_device_results(tapir_cuda_ctx* ctx, tapir_results* d)
{
    if (!ctx || !d) return fail(ctx, TAPIR_E_INVALID, "null argument");
    const Lane& ln = ctx->lanes[0];
    if (!ln.plan.valid) return fail(ctx, TAPIR_E_STATE, "no resident batch");
    d->pi_net = ln.d_pi_net.as<double>();
    d->pi_times = ln.d_pi_times.as<double>();
    d->integral = ln.d_integral.as<double>();
    d->abserr = ln.d_abserr.as<double>();
    d->max_pi = ln.d_max_pi.as<double>();
    d->argmax_t = ln.d_argmax_t.as<int64_t>();
    d->argmax_epoch = ln.d_argmax_epoch.as<int32_t>();
    d->mean_rate = ln.d_mean_rate.as<double>();
    d->n_finite = ln.d_n_finite.as<int64_t>();
    d->n_multi_panel = ln.d_n_multi.as<int64_t>();
    return TAPIR_OK;
}

int tapir_cuda_device_small_block(tapir_cuda_ctx* ctx, void** block, size_t* bytes, size_t section_offset[9])
{
    if (!ctx || !block || !bytes || !section_offset) return fail(ctx, TAPIR_E_INVALID, "null argument");
    const Lane& ln = ctx->lanes[0];
    if (!ln.plan.valid) return fail(ctx, TAPIR_E_STATE, "no resident batch");
    *block = ln.d_small.p;
    *bytes = ln.plan.small_bytes;
    for (int i = 0; i < 9; ++i) section_offset[i] = ln.plan.small_off[i];
    return TAPIR_OK;
}

int tapir_cuda_rank(tapir_cuda_ctx* ctx, int32_t n_top, int64_t* top_locus, double* top_pi, int64_t* epoch_count)
{
    int rc = bind(ctx);
    if (rc) return rc;
    Lane& ln = ctx->lanes[0];
    const Plan& plan = ln.plan;
    if (!plan.valid) return fail(ctx, TAPIR_E_STATE, "tapir_cuda_rank needs a resident, computed batch (upload + compute)");
    if (n_top < 0 || n_top > kRankMaxTop) return fail(ctx, TAPIR_E_INVALID, "n_top must be in [0, %d]", kRankMaxTop);
    const int E = plan.E;
    if (E == 0) return TAPIR_OK;
    cudaStream_t st = ln.stream;
    const size_t b_idx = size_t(E) * n_top * 8, b_cnt = size_t(E) * 8;
    CU(ctx, ctx->d_scratch_a.ensure(b_idx + 8));
    CU(ctx, ctx->d_scratch_b.ensure(b_idx + 8));
    CU(ctx, ctx->d_scratch_c.ensure(b_cnt + 8));
    if (n_top > 0 && (top_locus || top_pi)) {
        rank_kernel<<<unsigned(E), kRankThreads, 0, st>>>(ln.d_integral.as<double>(), plan.L, E, n_top,
                                                          ctx->d_scratch_a.as<int64_t>(), ctx->d_scratch_b.as<double>());
        CU(ctx, cudaGetLastError());
        if (top_locus) CU(ctx, cudaMemcpyAsync(top_locus, ctx->d_scratch_a.p, b_idx, cudaMemcpyDeviceToHost, st));
        if (top_pi) CU(ctx, cudaMemcpyAsync(top_pi, ctx->d_scratch_b.p, b_idx, cudaMemcpyDeviceToHost, st));
    }
    if (epoch_count) {
        CU(ctx, cudaMemsetAsync(ctx->d_scratch_c.p, 0, b_cnt, st));
        if (plan.L > 0) {
            const int64_t blocks = std::min<int64_t>((plan.L + 255) / 256, int64_t(ctx->sm_count) * 8);
            epoch_count_kernel<<<unsigned(blocks), 256, 0, st>>>(ln.d_argmax_epoch.as<int32_t>(), plan.L, E,
                                                                 ctx->d_scratch_c.as<unsigned long long>());
            CU(ctx, cudaGetLastError());
        }
        CU(ctx, cudaMemcpyAsync(epoch_count, ctx->d_scratch_c.p, b_cnt, cudaMemcpyDeviceToHost, st));
    }
    CU(ctx, cudaStreamSynchronize(st));
    return TAPIR_OK;
}

int tapir_cuda_gather_detach(tapir_cuda_ctx* ctx)
{
    int rc = bind(ctx);
    if (rc) return rc;
    for (Lane& ln : ctx->lanes)
        if (ln.stream) cudaStreamSynchronize(ln.stream);
    if (ctx->gather_mapped) cudaIpcCloseMemHandle(ctx->gather_mapped);
    ctx->gather_mapped = nullptr;
    ctx->gather_slot_ptr = nullptr;
    for (Lane& ln : ctx->lanes) drop_graph(ln);
    ctx->gather_slot_bytes = ctx->gather_total_bytes = 0;
    ctx->d_gather.release();
    cudaGetLastError();
    return TAPIR_OK;
}

int tapir_cuda_gather_create(tapir_cuda_ctx* ctx, size_t slot_bytes, int n_slots, void* handle_out)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!handle_out || slot_bytes == 0 || n_slots < 1) return fail(ctx, TAPIR_E_INVALID, "bad gather geometry");
    static_assert(sizeof(cudaIpcMemHandle_t) == TAPIR_IPC_HANDLE_BYTES, "IPC handle size");
    tapir_cuda_gather_detach(ctx);
    ctx->d_gather.release();
    CU(ctx, ctx->d_gather.ensure(slot_bytes * size_t(n_slots)));
    CU(ctx, cudaMemset(ctx->d_gather.p, 0, slot_bytes * size_t(n_slots)));
    cudaIpcMemHandle_t h;
    CU(ctx, cudaIpcGetMemHandle(&h, ctx->d_gather.p));
    memcpy(handle_out, &h, sizeof h);
    ctx->gather_total_bytes = slot_bytes * size_t(n_slots);
    return TAPIR_OK;
}

int tapir_cuda_gather_attach(tapir_cuda_ctx* ctx, const void* handle, size_t slot_bytes, int slot)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (slot < 0 || slot_bytes == 0) return fail(ctx, TAPIR_E_INVALID, "bad gather slot");
    char* base = nullptr;
    if (!handle) {
        if (!ctx->d_gather.p) return fail(ctx, TAPIR_E_STATE, "tapir_cuda_gather_attach(NULL) on a context that owns no gather buffer");
        if (slot_bytes * size_t(slot + 1) > ctx->gather_total_bytes) return fail(ctx, TAPIR_E_INVALID, "slot outside the gather buffer");
        base = ctx->d_gather.as<char>();
    } else {
        if (ctx->gather_mapped) cudaIpcCloseMemHandle(ctx->gather_mapped);
        ctx->gather_mapped = nullptr;
        cudaIpcMemHandle_t h;
        memcpy(&h, handle, sizeof h);
        void* p = nullptr;
        CU(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->gather_mapped = p;
        base = static_cast<char*>(p);
    }
    ctx->gather_slot_ptr = base + slot_bytes * size_t(slot);
    ctx->gather_slot_bytes = slot_bytes;
    for (Lane& ln : ctx->lanes) drop_graph(ln);  // the push kernel joins the captured sequence
    return TAPIR_OK;
}

int tapir_cuda_gather_read(tapir_cuda_ctx* ctx, void* host_dst, size_t bytes)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!ctx->d_gather.p) return fail(ctx, TAPIR_E_STATE, "this context owns no gather buffer");
    if (!host_dst || bytes > ctx->gather_total_bytes) return fail(ctx, TAPIR_E_INVALID, "bad destination");
    ctx->bytes_d2h += bytes;
    CU(ctx, cudaMemcpyAsync(host_dst, ctx->d_gather.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return TAPIR_OK;
}

int tapir_cuda_set_serial_kernels(tapir_cuda_ctx* ctx, int serial)
{
    if (!ctx) return fail(nullptr, TAPIR_E_INVALID, "null context");
    ctx->serial_kernels = serial ? 1 : 0;
    return TAPIR_OK;
}

int tapir_cuda_transfer_bytes(tapir_cuda_ctx* ctx, uint64_t* h2d, uint64_t* d2h)
{
    if (!ctx) return fail(nullptr, TAPIR_E_INVALID, "null context");
    if (h2d) *h2d = ctx->bytes_h2d;
    if (d2h) *d2h = ctx->bytes_d2h;
    return TAPIR_OK;
}

int tapir_cuda_upload(tapir_cuda_ctx* ctx, const tapir_problem* p)
{
    int rc = bind(ctx);
    if (rc) return rc;
    ctx->lanes[0].plan.valid = false;
    if ((rc = validate(ctx, p)) != TAPIR_OK) return rc;
    return lane_upload(ctx, ctx->lanes[0], p, 0, p->n_loci, true, decide_batch(ctx, p));
}

int tapir_cuda_compute(tapir_cuda_ctx* ctx, tapir_timing* timing)
{
    int rc = bind(ctx);
    if (rc) return rc;
    rc = lane_compute(ctx, ctx->lanes[0], timing, true, true);
    if (rc == TAPIR_OK && ctx->prof_trace) {  // measurement only: the stamps of this step's profile_kernel launch
        std::vector<long long> h(ctx->prof_trace_words);
        CU(ctx, cudaMemcpy(h.data(), ctx->prof_trace, h.size() * 8, cudaMemcpyDeviceToHost));
        if (FILE* f = fopen(ctx->prof_trace_path.c_str(), "wb")) {
            fwrite(h.data(), 8, h.size(), f);
            fclose(f);
        }
    }
    return rc;
}

int tapir_cuda_download(tapir_cuda_ctx* ctx, const tapir_results* out)
{
    int rc = bind(ctx);
    if (rc) return rc;
    return lane_download(ctx, ctx->lanes[0], out, 0, true);
}

// Whole path with host buffers.  Large batches are cut into chunks of loci that rotate over the
// lanes, so the H2D copy of chunk k+1 and the D2H copy of chunk k-1 overlap the kernels of chunk k (PCIe is
// full duplex; a step of this path moves ~3x more bytes out than in).
int tapir_cuda_run(tapir_cuda_ctx* ctx, const tapir_problem* p, const tapir_results* host_out)
{
    int rc = bind(ctx);
    if (rc) return rc;
    for (Lane& ln : ctx->lanes) ln.plan.valid = false;
    if ((rc = validate(ctx, p)) != TAPIR_OK) return rc;
    if (!host_out) return fail(ctx, TAPIR_E_INVALID, "null results");
    const int64_t L = p->n_loci;
    const int64_t nnz = L ? p->offsets[L] : 0;
    const double bytes = double(nnz) * 8.0 + ((p->flags & TAPIR_WANT_PROFILE) ? double(L) * double(p->n_grid) * 8.0 : 0.0);
    double chunk_bytes = double(13 << 20);  // measured on config 2 (20 MB in, 41 MB out): 8 MB chunks 1.19 ms, 12-14 MB 1.03-1.05 ms
    if (const char* env = getenv("TAPIR_CUDA_CHUNK_MB")) {  // tuning override
        const double v = atof(env);
        if (v > 0.0) chunk_bytes = v * double(1 << 20);
    }
    int64_t n_chunks = int64_t(bytes / chunk_bytes);
    n_chunks = std::max<int64_t>(1, std::min<int64_t>(n_chunks, 16));
    if (L < 64 * n_chunks) n_chunks = 1;  // few (possibly huge) loci: one pass, split over sites instead
    if (ctx->gather_slot_ptr) {  // where each chunk's rows go inside the full batch's small-block layout
        const size_t Ls = size_t(L), Ks = size_t(p->n_times), Es = size_t(p->n_epochs);  // as lane_upload lays it out
        const size_t sizes[9] = {Ls * Ks * 8, Ls * Es * 8, Ls * Es * 8, Ls * 8, Ls * 8, Ls * 8, Ls * 8, Ls * 8, Ls * 4};
        size_t o = 0;
        for (int i = 0; i < 9; ++i) {
            ctx->gather_full_off[i] = o;
            o += (sizes[i] + 255) & ~size_t(255);
        }
        ctx->gather_chunked = true;
    }
    // Chunk boundaries at cumulative shares of the sites.  Nothing flows back to the host before the first chunk
    // has been uploaded and computed, so the run opens with short chunks (1/ramp, 2/ramp, ... of the regular size).
    std::vector<int64_t> cut;  // site count at which chunk k ends
    {
        const int64_t per = nnz / n_chunks;
        int64_t done = 0;
        if (ctx->ramp > 1 && n_chunks >= 2 && L >= 256 * n_chunks) {
            // ctx->ramp = 4: a quarter- and a half-size chunk; 8: an eighth, a quarter, a half
            for (int64_t piece = per / ctx->ramp; piece < per && piece > 0; piece *= 2) cut.push_back(done += piece);
        }
        while (nnz - done > per + per / 2) cut.push_back(done += per);
        cut.push_back(nnz);
    }
    n_chunks = int64_t(cut.size());
    const int n_lanes = ctx->n_lanes;
    const Decision dec = decide_batch(ctx, p);
    const double t_host0 = now_us();
    if (ctx->trace) CU(ctx, cudaEventRecord(ctx->ev[0], ctx->stream_h2d));
    int64_t l0 = 0;
    for (int64_t k = 0; k < n_chunks && l0 < L; ++k) {
        int64_t l1 = L;
        if (k + 1 < n_chunks) {
            l1 = std::lower_bound(p->offsets + l0, p->offsets + L, cut[k]) - p->offsets;
            l1 = std::max(l1, l0 + 1);
            l1 = std::min(l1, L);
        }
        Lane& ln = ctx->lanes[k % n_lanes];
        if (k >= n_lanes && (rc = lane_finish(ctx, ln)) != TAPIR_OK) break;  // its previous chunk (and its staging) is done
        if (ctx->trace) ln.host_us[0] = now_us() - t_host0;
        if ((rc = lane_upload(ctx, ln, p, l0, l1, false, dec)) != TAPIR_OK) break;
        ln.pending_row0 = l0;
        if (ctx->trace) ln.host_us[1] = now_us() - t_host0;
        if ((rc = lane_compute(ctx, ln, nullptr, false, false)) != TAPIR_OK) break;
        if (ctx->trace) ln.host_us[2] = now_us() - t_host0;
        if ((rc = lane_download(ctx, ln, host_out, l0, false)) != TAPIR_OK) break;
        if (ctx->trace) ln.host_us[3] = now_us() - t_host0;
        l0 = l1;
    }
    for (Lane& ln : ctx->lanes) {
        const int rc2 = lane_finish(ctx, ln);
        if (rc == TAPIR_OK) rc = rc2;
    }
    if (ctx->trace && rc == TAPIR_OK) {
        // chunk k ran on lane k % n_lanes; with at least as many lanes as chunks every lane still holds its chunk's events
        fprintf(stderr, "[tapir trace] %lld chunks on %d lanes, host done at %.0f us\n", (long long)n_chunks, n_lanes, now_us() - t_host0);
        fprintf(stderr, "[tapir trace] chunk  loci   host: upload compute download enq'd |  device: H2D landed, kernels done, D2H start, D2H done (us)\n");
        for (int64_t k = 0; k < std::min<int64_t>(n_chunks, n_lanes); ++k) {
            Lane& ln = ctx->lanes[k];
            float t[4] = {0, 0, 0, 0};
            cudaEventElapsedTime(&t[0], ctx->ev[0], ln.ev_in);
            cudaEventElapsedTime(&t[1], ctx->ev[0], ln.ev_done);
            cudaEventElapsedTime(&t[2], ctx->ev[0], ln.ev_dstart);
            cudaEventElapsedTime(&t[3], ctx->ev[0], ln.ev_out);
            fprintf(stderr, "[tapir trace] %5lld %5lld   %6.0f %6.0f %6.0f %6.0f | %7.0f %7.0f %7.0f %7.0f\n", (long long)k,
                    (long long)ln.plan.L, ln.host_us[0], ln.host_us[1], ln.host_us[2], ln.host_us[3], t[0] * 1e3, t[1] * 1e3,
                    t[2] * 1e3, t[3] * 1e3);
        }
        cudaGetLastError();
    }
    ctx->gather_chunked = false;
    if (n_chunks > 1) ctx->lanes[0].plan.valid = false;  // the resident data is one chunk, not the batch
    return rc;
}

int tapir_cuda_pi_matrix(tapir_cuda_ctx* ctx, const double* time, int64_t n_time, const double* rates, int64_t n_rates,
                         double* out)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n_time < 0 || n_rates < 0) return fail(ctx, TAPIR_E_INVALID, "negative size");
    const int64_t n = n_time * n_rates;
    if (n == 0) return TAPIR_OK;
    if (!time || !rates || !out) return fail(ctx, TAPIR_E_INVALID, "null pointer");
    CU(ctx, ctx->d_scratch_a.ensure(size_t(n_time) * 8));
    CU(ctx, ctx->d_scratch_b.ensure(size_t(n_rates) * 8));
    CU(ctx, ctx->d_scratch_c.ensure(size_t(n) * 8));
    cudaStream_t st = ctx->stream;
    CU(ctx, cudaMemcpyAsync(ctx->d_scratch_a.p, time, size_t(n_time) * 8, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(ctx->d_scratch_b.p, rates, size_t(n_rates) * 8, cudaMemcpyHostToDevice, st));
    const int64_t blocks = std::min<int64_t>((n + 255) / 256, int64_t(ctx->sm_count) * 16);
    pi_matrix_kernel<<<unsigned(blocks), 256, 0, st>>>(ctx->d_scratch_a.as<double>(), n_time, ctx->d_scratch_b.as<double>(),
                                                        n_rates, ctx->d_scratch_c.as<double>());
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(out, ctx->d_scratch_c.p, size_t(n) * 8, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    return TAPIR_OK;
}

int tapir_cuda_nansum_rows(tapir_cuda_ctx* ctx, const double* matrix, int64_t n_rows, int64_t n_cols, double* out)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n_rows < 0 || n_cols < 0) return fail(ctx, TAPIR_E_INVALID, "negative size");
    if (n_rows == 0) return TAPIR_OK;
    if (!out || (n_cols > 0 && !matrix)) return fail(ctx, TAPIR_E_INVALID, "null pointer");
    const size_t n = size_t(n_rows) * size_t(n_cols);
    CU(ctx, ctx->d_scratch_c.ensure(n * 8 + 8));
    CU(ctx, ctx->d_scratch_a.ensure(size_t(n_rows) * 8));
    cudaStream_t st = ctx->stream;
    if (n) CU(ctx, cudaMemcpyAsync(ctx->d_scratch_c.p, matrix, n * 8, cudaMemcpyHostToDevice, st));
    const int64_t blocks = (n_rows * 32 + 255) / 256;
    nansum_rows_kernel<<<unsigned(blocks), 256, 0, st>>>(ctx->d_scratch_c.as<double>(), n_rows, n_cols,
                                                         ctx->d_scratch_a.as<double>());
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(out, ctx->d_scratch_a.p, size_t(n_rows) * 8, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    return TAPIR_OK;
}

int tapir_cuda_site_integrals(tapir_cuda_ctx* ctx, double a, double b, const double* rates, int64_t n, double* integral,
                              double* abserr, int32_t* panels)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (n < 0) return fail(ctx, TAPIR_E_INVALID, "negative size");
    if (n == 0) return TAPIR_OK;
    if (!rates || !integral || !abserr) return fail(ctx, TAPIR_E_INVALID, "null pointer");
    CU(ctx, ctx->d_scratch_a.ensure(size_t(n) * 8));
    CU(ctx, ctx->d_scratch_b.ensure(size_t(n) * 8));
    CU(ctx, ctx->d_scratch_c.ensure(size_t(n) * 8 + size_t(n) * 4));
    cudaStream_t st = ctx->stream;
    double* d_err = ctx->d_scratch_c.as<double>();
    int32_t* d_pan = reinterpret_cast<int32_t*>(d_err + n);
    CU(ctx, cudaMemcpyAsync(ctx->d_scratch_a.p, rates, size_t(n) * 8, cudaMemcpyHostToDevice, st));
    site_integrals_kernel<<<unsigned((n + 127) / 128), 128, 0, st>>>(a, b, ctx->d_scratch_a.as<double>(), n,
                                                                     ctx->d_scratch_b.as<double>(), d_err, d_pan);
    CU(ctx, cudaGetLastError());
    CU(ctx, cudaMemcpyAsync(integral, ctx->d_scratch_b.p, size_t(n) * 8, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaMemcpyAsync(abserr, d_err, size_t(n) * 8, cudaMemcpyDeviceToHost, st));
    if (panels) CU(ctx, cudaMemcpyAsync(panels, d_pan, size_t(n) * 4, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    return TAPIR_OK;
}

int tapir_cuda_fp64_peak(tapir_cuda_ctx* ctx, double* tflops, float* ms_out)
{
    int rc = bind(ctx);
    if (rc) return rc;
    if (!tflops) return fail(ctx, TAPIR_E_INVALID, "null pointer");
    CU(ctx, ctx->d_counter.ensure(256));
    CU(ctx, ctx->d_scratch_a.ensure(32 * 8));
    cudaStream_t st = ctx->stream;
    double seed[32];
    for (int i = 0; i < 32; ++i) seed[i] = 1.0 + 1e-3 * i;
    CU(ctx, cudaMemcpyAsync(ctx->d_scratch_a.p, seed, sizeof seed, cudaMemcpyHostToDevice, st));
    const int iters = 4096, blocks = ctx->sm_count * 4, threads = 128;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CU(ctx, cudaEventRecord(ctx->ev[0], st));
        fp64_peak_kernel<<<blocks, threads, 0, st>>>(ctx->d_counter.as<double>() + 8, iters, ctx->d_scratch_a.as<double>());
        CU(ctx, cudaGetLastError());
        CU(ctx, cudaEventRecord(ctx->ev[1], st));
        CU(ctx, cudaStreamSynchronize(st));
        float ms = 0.f;
        CU(ctx, cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
        if (rep > 0 && ms < best) best = ms;
    }
    const double flops = 2.0 * 64.0 * double(iters) * double(blocks) * double(threads);
    *tflops = flops / (double(best) * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return TAPIR_OK;
}

int tapir_mom_epoch_plan(double a, double b, int64_t n_tiles, int32_t tile_points, int32_t n_moments, double class_x,
                         tapir_mom_epoch* out)
{
    static_assert(sizeof(tapir_mom_epoch) == sizeof(MomEpochPlan), "public and internal plan structs agree");
    if (!out || n_tiles < 1 || tile_points < 1024 || tile_points % 1024 || (n_moments != 16 && n_moments != 32))
        return fail(nullptr, TAPIR_E_INVALID, "bad plan geometry");
    MomEpochPlan mp;
    if (!(class_x == 0.0 || class_x == 1.25 || class_x == 2.5 || class_x == 5.0)) return fail(nullptr, TAPIR_E_INVALID, "class_x is 0, 1.25, 2.5 or 5");
    mom_epoch_plan(a, b, n_tiles, tile_points, n_moments, &mp, class_x);
    memcpy(out, &mp, sizeof mp);
    return TAPIR_OK;
}

int tapir_cuda_flush_l2(tapir_cuda_ctx* ctx)
{
    int rc = bind(ctx);
    if (rc) return rc;
    const size_t bytes = size_t(256) << 20;  // > 126 MB L2
    CU(ctx, ctx->d_flush.ensure(bytes));
    CU(ctx, cudaMemsetAsync(ctx->d_flush.p, 0, bytes, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return TAPIR_OK;
}

}  // extern "C"
