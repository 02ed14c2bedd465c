// tapir_cuda.cu -- C ABI (include/tapir_cuda.h) over the sm_100a kernels in kernels.cuh.
//
// Host-side responsibilities: validate a CSR batch, build the work lists (time tiles, site
// splits, quad blocks), own device memory / streams, launch, time with CUDA events, copy back.
// There is no CPU fallback anywhere in this file: without a device every call fails.
#include "../../include/tapir_cuda.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <numeric>
#include <string>
#include <vector>

#include "kernels.cuh"

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
    int32_t TT = 1024, n_tiles = 1;
    int64_t n_prof_items = 0, n_quad_items = 0, ws_rows = 0;
    std::vector<EpochBatch> batches;
    bool valid = false;
};

}  // namespace

struct tapir_cuda_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string last_error;
    Plan plan;
    // inputs / work lists
    DevBuf d_rates, d_prof_items, d_quad_items, d_nsplit, d_ws_base, d_qitem_begin, d_time_idx, d_counter;
    // intermediates
    DevBuf d_ws, d_tile_max, d_tile_arg, d_qpart, d_fix;
    // results
    DevBuf d_pi_net, d_pi_times, d_integral, d_abserr, d_max_pi, d_argmax_t, d_argmax_epoch, d_mean_rate, d_n_finite,
        d_n_multi;
    DevBuf d_scratch_a, d_scratch_b, d_scratch_c, d_flush;
    HostBuf h_stage;
    uint64_t bytes_h2d = 0, bytes_d2h = 0;
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

int bind(tapir_cuda_ctx* ctx)
{
    if (!ctx) return fail(nullptr, TAPIR_E_INVALID, "null context");
    CU(ctx, cudaSetDevice(ctx->device));
    return TAPIR_OK;
}

// Estimated FP64 instructions per site for one time tile of the RH variant (table build + FMAs).
double tile_cost(int rh, int64_t T)
{
    const int64_t TT = 32 * 8 * rh;
    const int64_t nt = (T + TT - 1) / TT;
    const double gen = 45.0 + 32.0 + 8.0 * rh + (nt > 1 ? 28.0 : 0.0);
    return double(nt) * (double(TT) + gen);
}

int choose_rh(int64_t T)
{
    int best = 1;
    double bc = tile_cost(1, T);
    for (int rh : {2, 4, 8}) {
        const double c = tile_cost(rh, T);
        if (c < bc) { bc = c; best = rh; }
    }
    return best;
}

template <int RH>
cudaError_t launch_profile(tapir_cuda_ctx* ctx, const ProfArgs& pa, int64_t n_items)
{
    using Cfg = ProfCfg<RH>;
    static bool configured[64] = {false};
    if (!configured[ctx->device & 63]) {
        cudaError_t e = cudaFuncSetAttribute(profile_kernel<RH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             int(Cfg::SmemBytes));
        if (e != cudaSuccess) return e;
        configured[ctx->device & 63] = true;
    }
    int per_sm = 0;
    cudaError_t e =
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, profile_kernel<RH>, kProfWarps * 32, Cfg::SmemBytes);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = int64_t(ctx->sm_count) * per_sm;
    const int64_t need = (n_items + kProfWarps - 1) / kProfWarps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    profile_kernel<RH><<<unsigned(grid), kProfWarps * 32, Cfg::SmemBytes, ctx->stream>>>(pa);
    return cudaGetLastError();
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
            if (a < 0.0) eb.nonneg = 0;
            eb.bmax = std::max(eb.bmax, std::max(std::fabs(a), std::fabs(b)));
        }
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
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&ctx->ev[i]);
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
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->d_rates,     &ctx->d_prof_items, &ctx->d_quad_items, &ctx->d_nsplit,    &ctx->d_ws_base,
                      &ctx->d_qitem_begin, &ctx->d_time_idx, &ctx->d_counter,    &ctx->d_ws,        &ctx->d_tile_max,
                      &ctx->d_tile_arg,  &ctx->d_qpart,      &ctx->d_fix,        &ctx->d_pi_net,     &ctx->d_pi_times,  &ctx->d_integral,
                      &ctx->d_abserr,    &ctx->d_max_pi,     &ctx->d_argmax_t,   &ctx->d_argmax_epoch, &ctx->d_mean_rate,
                      &ctx->d_n_finite,  &ctx->d_n_multi,    &ctx->d_scratch_a,  &ctx->d_scratch_b, &ctx->d_scratch_c,
                      &ctx->d_flush};
    for (DevBuf* b : bufs) b->release();
    ctx->h_stage.release();
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
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

int tapir_cuda_upload(tapir_cuda_ctx* ctx, const tapir_problem* p)
{
    int rc = bind(ctx);
    if (rc) return rc;
    ctx->plan.valid = false;
    if ((rc = validate(ctx, p)) != TAPIR_OK) return rc;

    Plan& plan = ctx->plan;
    plan.L = p->n_loci;
    plan.T = p->n_grid;
    plan.K = p->n_times;
    plan.E = p->n_epochs;
    plan.flags = p->flags;
    plan.nnz = plan.L ? p->offsets[plan.L] : 0;
    plan.rh = choose_rh(std::max<int64_t>(plan.T, 1));
    plan.TT = 32 * 8 * plan.rh;
    plan.n_tiles = int32_t(std::max<int64_t>(1, (plan.T + plan.TT - 1) / plan.TT));
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
    std::vector<int64_t> ws_base(L), qitem_begin(L + 1);
    int64_t n_prof = 0, n_quad = 0, ws_rows = 0;
    for (int64_t l = 0; l < L; ++l) {
        const int64_t s = p->offsets[l + 1] - p->offsets[l];
        if (s > INT32_MAX) return fail(ctx, TAPIR_E_INVALID, "locus %lld has too many sites", (long long)l);
        int64_t ns = (s + chunk_sites - 1) / chunk_sites;
        if (ns < 1) ns = 1;
        nsplit[l] = int32_t(ns);
        ws_base[l] = ws_rows;
        if (ns > 1) ws_rows += ns * plan.n_tiles;
        n_prof += ns * plan.n_tiles;
        qitem_begin[l] = n_quad;
        n_quad += (s + kQuadBlock - 1) / kQuadBlock;
    }
    qitem_begin[L] = n_quad;
    if (n_prof >= (int64_t(1) << 31) || n_quad >= (int64_t(1) << 31))
        return fail(ctx, TAPIR_E_INVALID, "batch too large: split it into several calls");
    plan.n_prof_items = n_prof;
    plan.n_quad_items = n_quad;
    plan.ws_rows = ws_rows;

    // ---- stage work lists in pinned memory ----
    const size_t b_prof = size_t(n_prof) * sizeof(ProfItem), b_quad = size_t(n_quad) * sizeof(QuadItem);
    const size_t b_ns = size_t(L) * sizeof(int32_t), b_wb = size_t(L) * sizeof(int64_t);
    const size_t b_qb = size_t(L + 1) * sizeof(int64_t), b_ti = size_t(plan.K) * sizeof(int64_t);
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t o_prof = 0, o_quad = o_prof + al(b_prof), o_ns = o_quad + al(b_quad), o_wb = o_ns + al(b_ns),
                 o_qb = o_wb + al(b_wb), o_ti = o_qb + al(b_qb), total = o_ti + al(b_ti);
    CU(ctx, ctx->h_stage.ensure(total));
    char* hs = ctx->h_stage.as<char>();
    ProfItem* pi = reinterpret_cast<ProfItem*>(hs + o_prof);
    QuadItem* qi = reinterpret_cast<QuadItem*>(hs + o_quad);
    {
        int64_t k = 0, q = 0;
        bool ragged = false;
        for (int64_t l = 0; l < L; ++l) {
            const int64_t s0 = p->offsets[l], s = p->offsets[l + 1] - s0;
            if (l && s != p->offsets[1] - p->offsets[0]) ragged = true;
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
    {
        int64_t* ti = reinterpret_cast<int64_t*>(hs + o_ti);
        for (int k = 0; k < plan.K; ++k) ti[k] = p->time_idx[k] < 0 ? p->time_idx[k] + plan.T : p->time_idx[k];
    }

    // ---- device buffers ----
    const size_t qrow = size_t(2 * plan.E + 4);
    CU(ctx, ctx->d_rates.ensure(size_t(plan.nnz) * 8 + 8));
    CU(ctx, ctx->d_prof_items.ensure(b_prof + 8));
    CU(ctx, ctx->d_quad_items.ensure(b_quad + 8));
    CU(ctx, ctx->d_nsplit.ensure(b_ns + 8));
    CU(ctx, ctx->d_ws_base.ensure(b_wb + 8));
    CU(ctx, ctx->d_qitem_begin.ensure(b_qb));
    CU(ctx, ctx->d_time_idx.ensure(b_ti + 8));
    CU(ctx, ctx->d_counter.ensure(256));
    CU(ctx, ctx->d_ws.ensure(size_t(ws_rows) * plan.TT * 8 + 8));
    CU(ctx, ctx->d_tile_max.ensure(size_t(L) * plan.n_tiles * 8 + 8));
    CU(ctx, ctx->d_tile_arg.ensure(size_t(L) * plan.n_tiles * 8 + 8));
    CU(ctx, ctx->d_qpart.ensure(size_t(n_quad) * qrow * 8 + 8));
    CU(ctx, ctx->d_fix.ensure(size_t(n_quad) * kEpochsPerLaunch * sizeof(FixEntry) + 8));
    CU(ctx, ctx->d_pi_net.ensure(size_t(L) * plan.T * 8 + 16));
    CU(ctx, ctx->d_pi_times.ensure(size_t(L) * plan.K * 8 + 8));
    CU(ctx, ctx->d_integral.ensure(size_t(L) * plan.E * 8 + 8));
    CU(ctx, ctx->d_abserr.ensure(size_t(L) * plan.E * 8 + 8));
    CU(ctx, ctx->d_max_pi.ensure(size_t(L) * 8 + 8));
    CU(ctx, ctx->d_argmax_t.ensure(size_t(L) * 8 + 8));
    CU(ctx, ctx->d_argmax_epoch.ensure(size_t(L) * 4 + 8));
    CU(ctx, ctx->d_mean_rate.ensure(size_t(L) * 8 + 8));
    CU(ctx, ctx->d_n_finite.ensure(size_t(L) * 8 + 8));
    CU(ctx, ctx->d_n_multi.ensure(size_t(L) * 8 + 8));

    cudaStream_t st = ctx->stream;
    ctx->bytes_h2d += size_t(plan.nnz) * 8 + b_prof + b_quad + b_ns + b_wb + b_qb + b_ti;
    if (plan.nnz) CU(ctx, cudaMemcpyAsync(ctx->d_rates.p, p->rates, size_t(plan.nnz) * 8, cudaMemcpyHostToDevice, st));
    if (b_prof) CU(ctx, cudaMemcpyAsync(ctx->d_prof_items.p, hs + o_prof, b_prof, cudaMemcpyHostToDevice, st));
    if (b_quad) CU(ctx, cudaMemcpyAsync(ctx->d_quad_items.p, hs + o_quad, b_quad, cudaMemcpyHostToDevice, st));
    if (b_ns) CU(ctx, cudaMemcpyAsync(ctx->d_nsplit.p, hs + o_ns, b_ns, cudaMemcpyHostToDevice, st));
    if (b_wb) CU(ctx, cudaMemcpyAsync(ctx->d_ws_base.p, hs + o_wb, b_wb, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(ctx->d_qitem_begin.p, hs + o_qb, b_qb, cudaMemcpyHostToDevice, st));
    if (b_ti) CU(ctx, cudaMemcpyAsync(ctx->d_time_idx.p, hs + o_ti, b_ti, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaStreamSynchronize(st));
    plan.valid = true;
    return TAPIR_OK;
}

int tapir_cuda_compute(tapir_cuda_ctx* ctx, tapir_timing* timing)
{
    int rc = bind(ctx);
    if (rc) return rc;
    Plan& plan = ctx->plan;
    if (!plan.valid) return fail(ctx, TAPIR_E_STATE, "tapir_cuda_compute called before a successful tapir_cuda_upload");
    cudaStream_t st = ctx->stream;
    int launches = 0;
    CU(ctx, cudaEventRecord(ctx->ev[0], st));
    if (plan.L > 0) {
        CU(ctx, cudaMemsetAsync(ctx->d_counter.p, 0, 256, st));
        if (plan.n_prof_items > 0 && plan.T > 0) {

            ProfArgs pa;
            pa.rates = ctx->d_rates.as<double>();
            pa.items = ctx->d_prof_items.as<ProfItem>();
            pa.counter = ctx->d_counter.as<uint32_t>();
            pa.pi_net = ctx->d_pi_net.as<double>();
            pa.ws = ctx->d_ws.as<double>();
            pa.tile_max = ctx->d_tile_max.as<double>();
            pa.tile_arg = ctx->d_tile_arg.as<int64_t>();
            pa.T = plan.T;
            pa.n_items = int32_t(plan.n_prof_items);
            pa.n_tiles = plan.n_tiles;
            cudaError_t e = cudaSuccess;
            switch (plan.rh) {
                case 1: e = launch_profile<1>(ctx, pa, plan.n_prof_items); break;
                case 2: e = launch_profile<2>(ctx, pa, plan.n_prof_items); break;
                case 4: e = launch_profile<4>(ctx, pa, plan.n_prof_items); break;
                default: e = launch_profile<8>(ctx, pa, plan.n_prof_items); break;
            }
            CU(ctx, e);
            ++launches;
        }
        CU(ctx, cudaEventRecord(ctx->ev[1], st));
        if (plan.n_quad_items > 0) {
            QuadArgs qa;
            qa.rates = ctx->d_rates.as<double>();
            qa.items = ctx->d_quad_items.as<QuadItem>();
            qa.qpart = ctx->d_qpart.as<double>();
            qa.fix = ctx->d_fix.as<FixEntry>();
            qa.n_items = int32_t(plan.n_quad_items);
            int k = 0;
            for (const EpochBatch& eb : plan.batches) {
                // one counter per batch (d_counter[1..]); slot 0 is the profile kernel's work queue head
                qa.fix_count = ctx->d_counter.as<uint32_t>() + 1 + k++;
                const unsigned qgrid = unsigned(std::min<int64_t>(plan.n_quad_items, int64_t(ctx->sm_count) * 2));
                switch (eb.n) {
                    case 0: quad_kernel<0><<<qgrid, kQuadBlock, 0, st>>>(qa, eb); break;  // rate statistics only
                    case 1: quad_kernel<1><<<qgrid, kQuadBlock, 0, st>>>(qa, eb); break;
                    case 2: quad_kernel<2><<<qgrid, kQuadBlock, 0, st>>>(qa, eb); break;
                    case 3: quad_kernel<3><<<qgrid, kQuadBlock, 0, st>>>(qa, eb); break;
                    default: quad_kernel<4><<<qgrid, kQuadBlock, 0, st>>>(qa, eb); break;
                }
                CU(ctx, cudaGetLastError());
                quad_fix_kernel<<<unsigned(ctx->sm_count * 4), 128, 0, st>>>(qa, eb);
                CU(ctx, cudaGetLastError());
                launches += 2;
            }
        }
        CU(ctx, cudaEventRecord(ctx->ev[2], st));
        {
            FinalArgs fa;
            fa.n_loci = plan.L;
            fa.T = plan.T;
            fa.TT = plan.TT;
            fa.n_tiles = plan.T > 0 ? plan.n_tiles : 0;
            fa.K = plan.T > 0 ? plan.K : 0;
            fa.E = plan.E;
            fa.nsplit = ctx->d_nsplit.as<int32_t>();
            fa.ws_base = ctx->d_ws_base.as<int64_t>();
            fa.qitem_begin = ctx->d_qitem_begin.as<int64_t>();
            fa.time_idx = ctx->d_time_idx.as<int64_t>();
            fa.ws = ctx->d_ws.as<double>();
            fa.tile_max = ctx->d_tile_max.as<double>();
            fa.tile_arg = ctx->d_tile_arg.as<int64_t>();
            fa.qpart = ctx->d_qpart.as<double>();
            fa.pi_net = ctx->d_pi_net.as<double>();
            fa.pi_times = ctx->d_pi_times.as<double>();
            fa.integral = ctx->d_integral.as<double>();
            fa.abserr = ctx->d_abserr.as<double>();
            fa.max_pi = ctx->d_max_pi.as<double>();
            fa.argmax_t = ctx->d_argmax_t.as<int64_t>();
            fa.argmax_epoch = ctx->d_argmax_epoch.as<int32_t>();
            fa.mean_rate = ctx->d_mean_rate.as<double>();
            fa.n_finite = ctx->d_n_finite.as<int64_t>();
            fa.n_multi = ctx->d_n_multi.as<int64_t>();
            const int64_t blocks = (plan.L * 32 + 127) / 128;
            finalize_kernel<<<unsigned(blocks), 128, 0, st>>>(fa);
            CU(ctx, cudaGetLastError());
            ++launches;
        }
    } else {
        CU(ctx, cudaEventRecord(ctx->ev[1], st));
        CU(ctx, cudaEventRecord(ctx->ev[2], st));
    }
    CU(ctx, cudaEventRecord(ctx->ev[3], st));
    CU(ctx, cudaStreamSynchronize(st));
    if (timing) {
        memset(timing, 0, sizeof *timing);
        CU(ctx, cudaEventElapsedTime(&timing->total_ms, ctx->ev[0], ctx->ev[3]));
        CU(ctx, cudaEventElapsedTime(&timing->profile_ms, ctx->ev[0], ctx->ev[1]));
        CU(ctx, cudaEventElapsedTime(&timing->quad_ms, ctx->ev[1], ctx->ev[2]));
        CU(ctx, cudaEventElapsedTime(&timing->finalize_ms, ctx->ev[2], ctx->ev[3]));
        timing->launches = launches;
    }
    return TAPIR_OK;
}

int tapir_cuda_download(tapir_cuda_ctx* ctx, const tapir_results* out)
{
    int rc = bind(ctx);
    if (rc) return rc;
    const Plan& plan = ctx->plan;
    if (!plan.valid) return fail(ctx, TAPIR_E_STATE, "tapir_cuda_download called before upload/compute");
    if (!out) return fail(ctx, TAPIR_E_INVALID, "null results");
    cudaStream_t st = ctx->stream;
    const size_t L = size_t(plan.L);
    auto copy = [&](void* dst, const DevBuf& src, size_t bytes) -> cudaError_t {
        if (!dst || bytes == 0) return cudaSuccess;
        ctx->bytes_d2h += bytes;
        return cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, st);
    };
    if (plan.flags & TAPIR_WANT_PROFILE) CU(ctx, copy(out->pi_net, ctx->d_pi_net, L * size_t(plan.T) * 8));
    CU(ctx, copy(out->pi_times, ctx->d_pi_times, L * size_t(plan.T > 0 ? plan.K : 0) * 8));
    CU(ctx, copy(out->integral, ctx->d_integral, L * size_t(plan.E) * 8));
    CU(ctx, copy(out->abserr, ctx->d_abserr, L * size_t(plan.E) * 8));
    CU(ctx, copy(out->max_pi, ctx->d_max_pi, L * 8));
    CU(ctx, copy(out->argmax_t, ctx->d_argmax_t, L * 8));
    CU(ctx, copy(out->argmax_epoch, ctx->d_argmax_epoch, L * 4));
    CU(ctx, copy(out->mean_rate, ctx->d_mean_rate, L * 8));
    CU(ctx, copy(out->n_finite, ctx->d_n_finite, L * 8));
    CU(ctx, copy(out->n_multi_panel, ctx->d_n_multi, L * 8));
    CU(ctx, cudaStreamSynchronize(st));
    return TAPIR_OK;
}

int tapir_cuda_device_results(tapir_cuda_ctx* ctx, tapir_results* d)
{
    if (!ctx || !d) return fail(ctx, TAPIR_E_INVALID, "null argument");
    if (!ctx->plan.valid) return fail(ctx, TAPIR_E_STATE, "no resident batch");
    d->pi_net = ctx->d_pi_net.as<double>();
    d->pi_times = ctx->d_pi_times.as<double>();
    d->integral = ctx->d_integral.as<double>();
    d->abserr = ctx->d_abserr.as<double>();
    d->max_pi = ctx->d_max_pi.as<double>();
    d->argmax_t = ctx->d_argmax_t.as<int64_t>();
    d->argmax_epoch = ctx->d_argmax_epoch.as<int32_t>();
    d->mean_rate = ctx->d_mean_rate.as<double>();
    d->n_finite = ctx->d_n_finite.as<int64_t>();
    d->n_multi_panel = ctx->d_n_multi.as<int64_t>();
    return TAPIR_OK;
}

int tapir_cuda_transfer_bytes(tapir_cuda_ctx* ctx, uint64_t* h2d, uint64_t* d2h)
{
    if (!ctx) return fail(nullptr, TAPIR_E_INVALID, "null context");
    if (h2d) *h2d = ctx->bytes_h2d;
    if (d2h) *d2h = ctx->bytes_d2h;
    return TAPIR_OK;
}

int tapir_cuda_run(tapir_cuda_ctx* ctx, const tapir_problem* problem, const tapir_results* host_out)
{
    int rc = tapir_cuda_upload(ctx, problem);
    if (rc) return rc;
    if ((rc = tapir_cuda_compute(ctx, nullptr)) != TAPIR_OK) return rc;
    return tapir_cuda_download(ctx, host_out);
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
