// kernels.cuh -- sm_100a kernels of the Townsend-PI hot path.
//
//   profile_kernel<RH>   pi_net[l,t] = t * sum_i a_i q_i^t           (compute.py:46-48 + nansum,
//                        with a_i = 16 r_i^2, q_i = exp(-4 r_i)        bin/tapir_compute.py:119)
//   profile_mom_kernel<RH, KM>  the same profile for tiles of 1024+ grid points (the shipped default there): slow
//                        sites as KM = 16 / 32 Taylor moments per block -- the whole tile, 256 or 128 points by rate
//                        class -- an eighth of profile_kernel's DMMAs or less; also the slow sites' epoch sums
//   profile_f32_kernel   FP32 mode of the same (TAPIR_FP32_MODE)
//   quad_scan_kernel, quad_kernel<NE, CH>   rate statistics; first GK21 panel of every (site, epoch) the moments do
//                        not serve, collapsed (compute.py:50-52, 81-94)
//   quad_fix_kernel      full QAGS for the sites QAGS would bisect   (scipy.integrate.quad = QUADPACK dqagse)
//   finalize_kernel      per-locus reductions, --times picks, maxima (compute.py:76-79, README.rst:178-185)
//   rank_kernel, epoch_count_kernel   top-N loci per interval, loci per max-PI epoch
//                                     (bin/tapir_compare.py:93-101, README.rst:178-185)
//   push_small_kernel    multi-GPU: small result block -> rank 0's buffer over NVLink (peer stores)
//   pi_matrix_kernel, site_integrals_kernel, nansum_rows_kernel   element-wise parity with tapir.compute
//   fp64_peak_kernel     measurement: FP64 FMA peak of the device
//
// The profile is evaluated as a blocked Vandermonde product instead of one exp per
// (site, time): with t = T0 + hi*32 + lo,
//        a q^t = [a e^{-c T0} q^lo] * [(q^32)^hi] = U[lo] * W[hi],
// so pi_net[lo,hi] / t = sum_s U[s][lo] W[s][hi] is a dense FP64 contraction over sites.  One warp
// owns a 32 x H time tile of one locus, builds the two short power tables of 32 sites in shared
// memory (one exp per site plus DMUL chains), and contracts them with the FP64 MMA
// (mma.sync m8n8k4 -> DMMA.8x8x4): one FP64 FMA per (site, time), 256 per instruction.  DMMA and
// DFMA share one pipe on B200 (tools/microbench/fp64_pipes.cu: 37.0 vs 36.6 TFLOP/s, not additive);
// the MMA form is used because it needs 8x fewer issue slots and a third of the shared-memory
// fetches of a register-blocked DFMA outer product.  All arithmetic is IEEE double.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "fastexp.cuh"
#include "gk21.cuh"

namespace tapir {

// ------------------------------------------------------------------------------------------
// work descriptors (built on the host at upload time)
// ------------------------------------------------------------------------------------------
struct ProfItem {      // one (locus, time tile, site range)
    int64_t site_begin;  // index into rates
    int32_t n_sites;
    int32_t locus;
    int32_t tile;        // time tile index: T0 = tile * TT
    int32_t slot;        // >= 0: row of the split workspace to write raw partial sums to; -1: direct
    int32_t part;        // row of mom_int this item's moment-functional epoch sums go to (locus' first row + split index)
    int32_t pad_;
};

// One epoch served by the moment functional (mom_epoch.h): the anchor block of its centre and the output column.
struct MomEpochDev {
    int32_t col, tile, hb, pad_;
};

struct QuadItem {      // up to kQuadBlock consecutive sites of one locus
    int64_t site_begin;
    int32_t count;
    int32_t locus;
};

constexpr int kQuadBlock = 256;  // sites per quad item (one warp takes an item)
constexpr int kEpochsPerLaunch = 8;
constexpr int kProfWarps = 1;  // warps per CTA of the profile kernel (warps are autonomous)
constexpr int kB = 32;         // lo extent of a time tile
constexpr int kUStride = kB + 2;

struct EpochBatch {
    int32_t n;          // epochs in this launch
    int32_t first;      // first batch of the step: also writes the per-item rate statistics
    int32_t e_total;    // total epochs (row length of the partial table is 2*e_total + 4)
    int32_t nonneg;     // all limits >= 0: the integrand is >= 0 on every node
    int32_t col[kEpochsPerLaunch];  // output column of each epoch (epochs are batched sorted by half-length)
    double bmax;        // largest |limit| in the batch
    double a[kEpochsPerLaunch], b[kEpochsPerLaunch];
    double center[kEpochsPerLaunch], half[kEpochsPerLaunch];
    double tm[kEpochsPerLaunch][10], tp[kEpochsPerLaunch][10];  // centre -/+ half*xgk[j] as dqk21 forms them
    int32_t new_half[kEpochsPerLaunch];  // half-length differs from the previous epoch in the batch
    // moment-functional epochs (mom_epoch.h): a site of profile_mom_kernel's moment class whose rate lies in
    // [rmin, rmax] and outside [wlo, whi] is integrated through the locus' moments, not here (rates compared as
    // bit patterns: non-negative doubles order like unsigned integers)
    int32_t mom_any;                     // some epoch of this launch is served by the moments
    int32_t mom[kEpochsPerLaunch];
    unsigned long long rmin_bits[kEpochsPerLaunch], rmax_bits[kEpochsPerLaunch];
    unsigned long long wlo_bits[kEpochsPerLaunch], whi_bits[kEpochsPerLaunch];
    double tb[kEpochsPerLaunch];         // centre of the anchor block (absolute time)
    // quad_scan_kernel's conservative form on the high words of the rates: a site is served for every epoch of the
    // launch when mom_all is set, scan_lo < hi(rate) < scan_hi (the tightest range of the launch) and hi(rate) lies in
    // none of the n_win windows [win_lo, win_hi]
    int32_t mom_all, n_win;
    int32_t mom_terms;                   // moments per block the profile kernel runs with (32 or 16)
    uint32_t class_hi;                   // high word of the moment class bound (rate_is_slow)
    uint32_t scan_lo, scan_hi;
    uint32_t win_lo[kEpochsPerLaunch], win_hi[kEpochsPerLaunch];
    const double* mom_p;                // device [e_total][32]: coefficients of the functional, row = output column
};

struct ProfArgs {
    const double* rates;
    const ProfItem* items;
    uint32_t* counter;   // dynamic work queue head (zeroed before launch)
    double* pi_net;      // [L*T]
    double* ws;          // split workspace [slots][TT]
    double* tile_max;    // [L*n_tiles]
    int64_t* tile_arg;   // [L*n_tiles]
    int64_t T;
    int32_t n_items;
    int32_t n_tiles;
    long long* trace;    // measurement only (TAPIR_PROF_TRACE, profile_kernel<RH, true>): 32 words per warp
    // moment-functional epoch sums (profile_mom_kernel only; n_mom_epochs = 0 otherwise)
    const MomEpochDev* mom_epochs;
    const double* mom_p;  // [e_total][32]
    double* mom_int;      // [parts][e_total]
    int32_t n_mom_epochs;
    int32_t e_total;
    // profile_mom_kernel: seeds of the A fragments of the moments -> times product, [16 row tiles][32 lanes][2]:
    // entry (mt, lane) = {x^fk, x^4}, x = (8 mt + fg - 64)/64 (kMtPowDoubles doubles, filled by mt_pow_table; one
    // 16-byte load per row tile, 512 contiguous bytes per warp)
    const double* mt_pow;
};
constexpr int kMtPowDoubles = 16 * 32 * 2;
// (x is exact; x^2, x^3, x^4 are formed as the kernel formed them in registers before the table existed)
inline void mt_pow_table(double* out)
{
    for (int mt = 0; mt < 16; ++mt)
        for (int lane = 0; lane < 32; ++lane) {
            const int fk = lane & 3, fg = lane >> 2;
            const double x = double(8 * mt + fg - 64) * 0.015625, xx = x * x;
            out[(mt * 32 + lane) * 2] = (fk == 0) ? 1.0 : (fk == 1) ? x : (fk == 2) ? xx : xx * x;
            out[(mt * 32 + lane) * 2 + 1] = xx * xx;
        }
}

template <int RH>
struct ProfCfg {
    static constexpr int H = 8 * RH;            // hi extent = RH n-tiles of the 8x8x4 FP64 MMA
    static constexpr int TT = kB * H;           // times per tile
    static constexpr int WStride = (RH == 1) ? 12 : H + 2;
    static constexpr int WarpDoubles = 32 * (kUStride + WStride) + 64;  // U, W tables + the active-rate queue
    static constexpr size_t SmemBytes = size_t(kProfWarps) * WarpDoubles * sizeof(double);
    // resident 4-warp CTAs per SM the register budget is sized for (shared memory allows 3 / 2 at RH = 4 / 8)
    static constexpr int MinBlocks = (RH >= 8 ? 2 : (RH == 4 ? 3 : 4)) * 4 / kProfWarps;
};

__device__ __forceinline__ void st_shared_v2(double* p, double x, double y)
{
    *reinterpret_cast<double2*>(p) = make_double2(x, y);
}

// D(8x8) += A(8x4, row) * B(4x8, col) in IEEE double on the FP64 MMA path (SASS: DMMA.8x8x4).
// Fragment layout (PTX ISA, mma.m8n8k4 .f64): with g = lane >> 2 and k = lane & 3 a lane holds
// A[g][k], B[k][g] and D[g][2k], D[g][2k+1].
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}

// Row/column g of an MMA tile is mapped to a chunk of consecutive table entries so that one lane
// fetches its fragments for all m- (n-) tiles with 16-byte loads; the chunk order is permuted so that
// the 8 lanes of a quarter warp (k = 0..3, two values of g) touch 8 distinct 16-byte bank groups.
__device__ __forceinline__ int chunk_pair(int g) { return (g & 4) | ((g & 1) << 1) | ((g >> 1) & 1); }  // 0 2 1 3 4 6 5 7
__device__ __forceinline__ int chunk_quad(int g) { return ((g & 1) << 2) | (g >> 1); }                  // 0 4 1 5 2 6 3 7
template <int RH>
__device__ __forceinline__ int w_chunk(int g)
{
    if constexpr (RH == 4) return chunk_pair(g);
    else if constexpr (RH == 2) return chunk_quad(g);
    else return g;
}

// ------------------------------------------------------------------------------------------
// PI profile
// ------------------------------------------------------------------------------------------
// TRACE (measurement only, tools/prof_trace.py): lane 0 of every warp stamps clock64 into args.trace --
// [0] SM id, [1] entry, [2] first item claimed, [3] first contraction, [4..11] end of item k, [12] exit, [13] items done,
// [16 + 4 j + k], j = 0..3: groups 2..5 of the warp's second item, k = 0 group popped, 1 tables built, 2 contracted.
// The production instantiation (TRACE = false) carries none of it: the kernel's schedule is sensitive to every
// extra predicate (a run-time switch for the same stamps cost 8 %).
template <int RH, bool TRACE = false>
__global__ void __launch_bounds__(kProfWarps * 32, ProfCfg<RH>::MinBlocks) profile_kernel(const ProfArgs args)
{
    using Cfg = ProfCfg<RH>;
    constexpr int H = Cfg::H, TT = Cfg::TT, WS = Cfg::WStride, US = kUStride;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int fk = lane & 3, fg = lane >> 2;  // MMA fragment coordinates of this lane
    double* Us = smem + size_t(warp) * Cfg::WarpDoubles;
    double* Ws = Us + 32 * US;
    double* Qs = Ws + 32 * WS;
    double* myU = Us + lane * US;  // generation phase: lane == site
    double* myW = Ws + lane * WS;
    // fragment sources: site row fk (+4 per k-step); lo chunk = 4 entries, hi chunk = RH entries
    const double* fragU = Us + fk * US + 4 * chunk_pair(fg);
    const double* fragW = Ws + fk * WS + RH * w_chunk<RH>(fg);
    long long* tr = nullptr;
    int n_done = 0, n_group = 0;
    if constexpr (TRACE) {
        tr = args.trace + (size_t(blockIdx.x) * kProfWarps + warp) * 32;
        if (lane == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            tr[0] = smid;
            tr[1] = clock64();
        }
    }

    for (;;) {
        uint32_t it = 0;
        if (lane == 0) it = atomicAdd(args.counter, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= uint32_t(args.n_items)) break;
        const ProfItem item = args.items[it];
        const int64_t T0 = int64_t(item.tile) * TT;
        const double t0d = double(T0);
        if constexpr (TRACE) {
            if (lane == 0 && n_done == 0) tr[2] = clock64() + (item.n_sites == -12345);
            n_group = 0;
        }

        // acc[mt][nt][c]: lo = 4*chunk_pair(fg) + mt, hi = RH*w_chunk(2*fk + c) + nt
        double acc[4][RH][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < RH; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        // Sites with a zero rate (invariant sites: HyPhy prints 0.0000) and culled sites (NaN/Inf) add exactly 0
        // to every grid point.  They are dropped right after the load: the active rates of the item are
        // compacted, in order, through a 64-entry per-warp queue and the tables are built 32 active sites at
        // a time, so the contraction only ever sees sites that contribute.
        const double* rp = args.rates + item.site_begin;
        int qh = 0, qn = 0;  // queue head / tail (running counts)
        // rates are fetched two rounds ahead: a round that completes no group of 32 active sites consumes its
        // load immediately, so a distance of one would expose the DRAM latency every other round
        double lam_in = (lane < item.n_sites) ? __ldg(rp + lane) : 0.0;
        double lam_in2 = (32 + lane < item.n_sites) ? __ldg(rp + 32 + lane) : 0.0;
        for (int base = 0;; base += 32) {
            const bool more = base < item.n_sites;
            if (more) {
                const double lam_new = lam_in;
                lam_in = lam_in2;
                {
                    const int nb = base + 64 + lane;
                    lam_in2 = (nb < item.n_sites) ? __ldg(rp + nb) : 0.0;
                }
                const double a_new = 16.0 * (lam_new * lam_new);
                const bool act = (fabs(a_new) <= DBL_MAX) && (a_new != 0.0);  // lanes past the item hold 0
                const unsigned ballot = __ballot_sync(0xffffffffu, act);
                if (act) Qs[(qn + __popc(ballot & ((1u << lane) - 1u))) & 63] = lam_new;
                qn += __popc(ballot);
                __syncwarp();
            }
            while (qn - qh >= 32 || (!more && qn > qh)) {
                const int cnt = min(32, qn - qh);
                const double lam = (lane < cnt) ? Qs[(qh + lane) & 63] : 0.0;
                qh += cnt;
                int tg = -1;  // TRACE: which of the stamped groups this is
                if constexpr (TRACE) {
                    if (lane == 0 && n_done == 0 && n_group == 0) tr[3] = clock64() + (lam == -1.0);
                    tg = (n_done == 1 && n_group >= 2 && n_group < 6) ? n_group - 2 : -1;
                    ++n_group;
                    if (tg >= 0 && lane == 0) tr[16 + 4 * tg] = clock64() + (lam == -1.0);
                }
                // ---- generation: this lane's site -> U[0..31], W[0..H-1] (power tables) ----
                // rows of lanes >= cnt are U = 0, W = 1: they add exactly 0
                {
                    double a = 16.0 * (lam * lam);
                    const bool ok = lane < cnt;  // queue entries are finite and non-zero; idle lanes build a zero row
                    const double c = ok ? 4.0 * lam : 0.0;
                    a = ok ? a : 0.0;
                    const double q = exp_full(-c);
                    const double u0 = (T0 == 0) ? a : a * exp_full(-(c * t0d));
                    const double q2 = q * q, q4 = q2 * q2, q8 = q4 * q4, q16 = q8 * q8;
                    const double qB = q16 * q16;  // q^32
                    double xa = u0, xb = u0 * q16;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        st_shared_v2(myU + 2 * k, xa, xa * q);
                        st_shared_v2(myU + 16 + 2 * k, xb, xb * q);
                        xa *= q2;
                        xb *= q2;
                    }
                    if constexpr (RH >= 2) {
                        const double qB2 = qB * qB;
                        double qBh = qB2;  // -> qB^(H/2)
#pragma unroll
                        for (int k = 2; k < H / 2; k *= 2) qBh *= qBh;
                        double wa = 1.0, wb = qBh;
#pragma unroll
                        for (int k = 0; k < H / 4; ++k) {
                            st_shared_v2(myW + 2 * k, wa, wa * qB);
                            st_shared_v2(myW + H / 2 + 2 * k, wb, wb * qB);
                            wa *= qB2;
                            wb *= qB2;
                        }
                    } else {
                        const double qB2 = qB * qB;
                        double wa = 1.0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            st_shared_v2(myW + 2 * k, wa, wa * qB);
                            wa *= qB2;
                        }
                    }
                }
                __syncwarp();
                if constexpr (TRACE) {
                    if (tg >= 0 && lane == 0) tr[16 + 4 * tg + 1] = clock64();
                }
                // ---- contraction over the chunk's sites: acc(lo,hi) += sum_s U[s][lo] * W[s][hi] ----
                // 4 sites per k-step; fragments of step s+1 are fetched while the MMAs of step s issue
                {
                    auto fetch = [&](double (&a)[4], double (&b)[RH], int step) {
                        const double* up = fragU + step * (4 * US);
                        const double* wp = fragW + step * (4 * WS);
                        const double2 u01 = *reinterpret_cast<const double2*>(up);
                        const double2 u23 = *reinterpret_cast<const double2*>(up + 2);
                        a[0] = u01.x; a[1] = u01.y; a[2] = u23.x; a[3] = u23.y;
                        if constexpr (RH == 1) {
                            b[0] = wp[0];
                        } else {
#pragma unroll
                            for (int r = 0; r < RH / 2; ++r) {
                                const double2 v = *reinterpret_cast<const double2*>(wp + 2 * r);
                                b[2 * r] = v.x; b[2 * r + 1] = v.y;
                            }
                        }
                    };
                    auto mma_step = [&](const double (&a)[4], const double (&b)[RH]) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < RH; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                    };
                    const int nk = (cnt + 3) >> 2;  // 1..8 k-steps; rows past cnt are zero, so odd counts round up
                    double a0[4], b0[RH], a1[4], b1[RH];
                    fetch(a0, b0, 0);
                    for (int s = 0; s < nk; s += 2) {
                        fetch(a1, b1, s + 1);
                        mma_step(a0, b0);
                        fetch(a0, b0, min(s + 2, 7));
                        mma_step(a1, b1);
                    }
                }
                __syncwarp();
                if constexpr (TRACE) {
                    if (tg >= 0 && lane == 0) tr[16 + 4 * tg + 2] = clock64() + (acc[0][0][0] == -1.0);
                }
            }  // groups of 32 active sites
            if (!more) break;
        }

        // ---- epilogue: this lane holds 4 consecutive lo for each of its 2*RH hi ----
        const int lo0 = 4 * chunk_pair(fg);
        if (item.slot >= 0) {
            // split locus: raw partial sums, reduced in a fixed order by finalize_kernel
            double* out = args.ws + size_t(item.slot) * TT;
#pragma unroll
            for (int j = 0; j < RH; ++j)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int hi = RH * w_chunk<RH>(2 * fk + c) + j;
                    double* o = out + hi * 32 + lo0;
                    *reinterpret_cast<double2*>(o) = make_double2(acc[0][j][c], acc[1][j][c]);
                    *reinterpret_cast<double2*>(o + 2) = make_double2(acc[2][j][c], acc[3][j][c]);
                }
        } else {
            const int64_t row = int64_t(item.locus) * args.T;
            double* out = args.pi_net + row;
            const bool vec_ok = ((row & 1) == 0);
            // Tile maximum: the values are >= 0, so their bit patterns order like the values and the comparisons
            // run on the integer pipe instead of the FP64 pipe the other warps' MMAs are using.  A lane visits
            // its offsets in increasing order (c outer, j inner: w_chunk(2 fk) < w_chunk(2 fk + 1)), so a strict
            // comparison keeps the first maximum, as numpy.argmax does.
            long long best_bits = -1;
            int best_i = 0;  // offset within the tile
            const int64_t t_left = args.T - T0;
            const int t_end = int(t_left < int64_t(TT) ? t_left : int64_t(TT));  // valid offsets of this tile
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int j = 0; j < RH; ++j) {
                    const int hi = RH * w_chunk<RH>(2 * fk + c) + j;
                    const int o = hi * 32 + lo0;
                    const double td = t0d + double(o);  // exact: integers far below 2^53
                    double v[4];
                    v[0] = acc[0][j][c] * td;
                    v[1] = acc[1][j][c] * (td + 1.0);
                    v[2] = acc[2][j][c] * (td + 2.0);
                    v[3] = acc[3][j][c] * (td + 3.0);
                    double* dst = out + T0 + o;
                    if (o + 3 < t_end && vec_ok) {
                        *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
                        *reinterpret_cast<double2*>(dst + 2) = make_double2(v[2], v[3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (o + i < t_end) dst[i] = v[i];
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const long long bits = __double_as_longlong(v[i]);
                        if (o + i < t_end && bits > best_bits) {
                            best_bits = bits;
                            best_i = o + i;
                        }
                    }
                }
            int64_t best_t = T0 + best_i;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const long long ob = __shfl_xor_sync(0xffffffffu, best_bits, off);
                const int64_t ot = __shfl_xor_sync(0xffffffffu, best_t, off);
                if (ob > best_bits || (ob == best_bits && ot < best_t)) { best_bits = ob; best_t = ot; }
            }
            const double best = __longlong_as_double(best_bits);
            if (lane == 0) {
                const int64_t k = int64_t(item.locus) * args.n_tiles + item.tile;
                args.tile_max[k] = best;
                args.tile_arg[k] = best_t;
            }
        }
        if constexpr (TRACE) {
            __syncwarp();
            if (lane == 0 && n_done < 8) tr[4 + n_done] = clock64();
            ++n_done;
        }
    }
    if constexpr (TRACE) {
        if (lane == 0) { tr[12] = clock64(); tr[13] = n_done; }
    }
}


// finite, and 16 r^2 neither 0 nor Inf -- decided on the bit pattern: 16*(r*r) == 0 up to |r| = 0x1e56a09e667f3bcc
// (1.57e-162) and is finite up to |r| = 0x5fcfffffffffffff (3.35e153); NaN and Inf compare above both
__device__ __forceinline__ bool rate_is_active(double r)
{
    // on the two 32-bit halves, so that no part of the test is scheduled on the FP64 pipe
    const uint32_t hi = uint32_t(__double2hiint(r)) & 0x7fffffffu, lo = uint32_t(__double2loint(r));
    const bool above = hi > 0x1e56a09eu || (hi == 0x1e56a09eu && lo > 0x667f3bccu);
    return above && hi <= 0x5fcfffffu;
}

// ------------------------------------------------------------------------------------------
// PI profile
// ------------------------------------------------------------------------------------------
// power tables of one group: lane = site; rows of lanes >= cnt are U = 0, W = 1 (they add exactly 0)
template <int RH>
__device__ __forceinline__ void prof_build_tables(double lam, bool ok, int64_t T0, double t0d, double* myU, double* myW)
{
    constexpr int H = 8 * RH;
    double a = 16.0 * (lam * lam);
    const double c = ok ? 4.0 * lam : 0.0;
    a = ok ? a : 0.0;
    const double q = exp_full(-c);
    const double u0 = (T0 == 0) ? a : a * exp_full(-(c * t0d));
    const double q2 = q * q, q4 = q2 * q2, q8 = q4 * q4, q16 = q8 * q8;
    const double qB = q16 * q16;  // q^32
    double xa = u0, xb = u0 * q16;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        st_shared_v2(myU + 2 * k, xa, xa * q);
        st_shared_v2(myU + 16 + 2 * k, xb, xb * q);
        xa *= q2;
        xb *= q2;
    }
    if constexpr (RH >= 2) {
        const double qB2 = qB * qB;
        double qBh = qB2;  // -> qB^(H/2)
#pragma unroll
        for (int k = 2; k < H / 2; k *= 2) qBh *= qBh;
        double wa = 1.0, wb = qBh;
#pragma unroll
        for (int k = 0; k < H / 4; ++k) {
            st_shared_v2(myW + 2 * k, wa, wa * qB);
            st_shared_v2(myW + H / 2 + 2 * k, wb, wb * qB);
            wa *= qB2;
            wb *= qB2;
        }
    } else {
        const double qB2 = qB * qB;
        double wa = 1.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            st_shared_v2(myW + 2 * k, wa, wa * qB);
            wa *= qB2;
        }
    }
}

// acc(lo,hi) += sum_s U[s][lo] * W[s][hi] over nk k-steps of 4 sites; `last` = last k-step the tables hold (the
// look-ahead fetch is clamped to it: the tables may end where the warp's shared memory ends)
template <int RH, bool ROLLED = false>   // ROLLED: the k-step loop is not unrolled (profile_mom_kernel: code size)
__device__ __forceinline__ void prof_contract(double (&acc)[4][RH][2], const double* fragU, const double* fragW, int nk, int last)
{
    constexpr int WS = ProfCfg<RH>::WStride, US = kUStride;
    auto fetch = [&](double (&a)[4], double (&b)[RH], int step) {
        const double* up = fragU + step * (4 * US);
        const double* wp = fragW + step * (4 * WS);
        const double2 u01 = *reinterpret_cast<const double2*>(up);
        const double2 u23 = *reinterpret_cast<const double2*>(up + 2);
        a[0] = u01.x; a[1] = u01.y; a[2] = u23.x; a[3] = u23.y;
        if constexpr (RH == 1) {
            b[0] = wp[0];
        } else {
#pragma unroll
            for (int r = 0; r < RH / 2; ++r) {
                const double2 v = *reinterpret_cast<const double2*>(wp + 2 * r);
                b[2 * r] = v.x; b[2 * r + 1] = v.y;
            }
        }
    };
    auto mma_step = [&](const double (&a)[4], const double (&b)[RH]) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < RH; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    };
    double a0[4], b0[RH], a1[4], b1[RH];
    fetch(a0, b0, 0);
    if constexpr (ROLLED) {
#pragma unroll 1
        for (int s = 0; s < nk; s += 2) {
            fetch(a1, b1, min(s + 1, last));
            mma_step(a0, b0);
            fetch(a0, b0, min(s + 2, last));
            mma_step(a1, b1);
        }
    } else {
        for (int s = 0; s < nk; s += 2) {
            fetch(a1, b1, min(s + 1, last));
            mma_step(a0, b0);
            fetch(a0, b0, min(s + 2, last));
            mma_step(a1, b1);
        }
    }
}

// one item's accumulators -> pi_net (times t) and the tile maximum, or the split workspace
template <int RH>
__device__ __forceinline__ void prof_epilogue(const ProfArgs& args, const double (&acc)[4][RH][2], int locus, int tile, int slot,
                                              int lane)
{
    constexpr int TT = ProfCfg<RH>::TT;
    const int fk = lane & 3, fg = lane >> 2;
    const int64_t T0 = int64_t(tile) * TT;
    const double t0d = double(T0);
    const int lo0 = 4 * chunk_pair(fg);
    if (slot >= 0) {
        double* out = args.ws + size_t(slot) * TT;
#pragma unroll
        for (int j = 0; j < RH; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int hi = RH * w_chunk<RH>(2 * fk + c) + j;
                double* o = out + hi * 32 + lo0;
                *reinterpret_cast<double2*>(o) = make_double2(acc[0][j][c], acc[1][j][c]);
                *reinterpret_cast<double2*>(o + 2) = make_double2(acc[2][j][c], acc[3][j][c]);
            }
        return;
    }
    const int64_t row = int64_t(locus) * args.T;
    double* out = args.pi_net + row;
    const bool vec_ok = ((row & 1) == 0);
    long long best_bits = -1;
    int best_i = 0;
    const int64_t t_left = args.T - T0;
    const int t_end = int(t_left < int64_t(TT) ? t_left : int64_t(TT));
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < RH; ++j) {
            const int hi = RH * w_chunk<RH>(2 * fk + c) + j;
            const int o = hi * 32 + lo0;
            const double td = t0d + double(o);
            double v[4];
            v[0] = acc[0][j][c] * td;
            v[1] = acc[1][j][c] * (td + 1.0);
            v[2] = acc[2][j][c] * (td + 2.0);
            v[3] = acc[3][j][c] * (td + 3.0);
            double* dst = out + T0 + o;
            if (o + 3 < t_end && vec_ok) {
                *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
                *reinterpret_cast<double2*>(dst + 2) = make_double2(v[2], v[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (o + i < t_end) dst[i] = v[i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long bits = __double_as_longlong(v[i]);
                if (o + i < t_end && bits > best_bits) {
                    best_bits = bits;
                    best_i = o + i;
                }
            }
        }
    int64_t best_t = T0 + best_i;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const long long ob = __shfl_xor_sync(0xffffffffu, best_bits, off);
        const int64_t ot = __shfl_xor_sync(0xffffffffu, best_t, off);
        if (ob > best_bits || (ob == best_bits && ot < best_t)) { best_bits = ob; best_t = ot; }
    }
    if (lane == 0) {
        const int64_t k = int64_t(locus) * args.n_tiles + tile;
        args.tile_max[k] = __longlong_as_double(best_bits);
        args.tile_arg[k] = best_t;
    }
}

// The same for a sub-tile of 256 RH points that starts at grid point T0 (profile_mom_kernel's 1,024-point halves of a
// 2,048-point tile): values times t into `row` (the locus' row of pi_net) with the running maximum of the lane kept in
// best_bits / best_t for the caller to close, or -- raw -- the partial sums into `raw` (split workspace, at the sub-tile).
template <int RH>
__device__ __forceinline__ void prof_epilogue_at(const ProfArgs& args, const double (&acc)[4][RH][2], double* row, double* raw,
                                                 int64_t T0, int lane, long long& best_bits, int64_t& best_t)
{
    constexpr int TT = ProfCfg<RH>::TT;
    const int fk = lane & 3, fg = lane >> 2;
    const double t0d = double(T0);
    const int lo0 = 4 * chunk_pair(fg);
    if (raw) {
#pragma unroll
        for (int j = 0; j < RH; ++j)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int hi = RH * w_chunk<RH>(2 * fk + c) + j;
                double* o = raw + hi * 32 + lo0;
                *reinterpret_cast<double2*>(o) = make_double2(acc[0][j][c], acc[1][j][c]);
                *reinterpret_cast<double2*>(o + 2) = make_double2(acc[2][j][c], acc[3][j][c]);
            }
        return;
    }
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
    const int64_t t_left = args.T - T0;
    const int t_end = int(t_left < int64_t(TT) ? (t_left < 0 ? 0 : t_left) : int64_t(TT));
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int j = 0; j < RH; ++j) {
            const int hi = RH * w_chunk<RH>(2 * fk + c) + j;
            const int o = hi * 32 + lo0;
            const double td = t0d + double(o);
            double v[4];
            v[0] = acc[0][j][c] * td;
            v[1] = acc[1][j][c] * (td + 1.0);
            v[2] = acc[2][j][c] * (td + 2.0);
            v[3] = acc[3][j][c] * (td + 3.0);
            double* dst = row + T0 + o;
            if (o + 3 < t_end && vec_ok) {
                *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
                *reinterpret_cast<double2*>(dst + 2) = make_double2(v[2], v[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (o + i < t_end) dst[i] = v[i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long bits = __double_as_longlong(v[i]);
                if (o + i < t_end && bits > best_bits) {   // offsets rise within a lane: the first maximum is kept
                    best_bits = bits;
                    best_t = T0 + o + i;
                }
            }
        }
}

constexpr int kLinBlock = 130;  // doubles per 128-point block of the moments -> times scratch (128 + 2: the four row groups
                                // of a store land 8 banks apart instead of on the same banks)
// The same from shared memory: `lin` holds the sub-tile's 256 RH sums in their natural order, kLinBlock doubles per
// 128-point block.  A rolled loop, one block per round -- the warp 1 KB of the row per store pair -- instead of the
// unrolled walk over the accumulator registers: 110 instructions of code for 860 (the moment kernel's code has to stay
// inside the instruction cache: every straight-line stretch counts).
template <int RH>
__device__ __forceinline__ void prof_epilogue_lin(const ProfArgs& args, const double* lin, double* row, double* raw, int64_t T0,
                                                  int lane, long long& best_bits, int64_t& best_t)
{
    constexpr int TT = ProfCfg<RH>::TT;
    const int64_t t_left = args.T - T0;
    const int t_end = int(t_left < int64_t(TT) ? (t_left < 0 ? 0 : t_left) : int64_t(TT));
    const bool vec_ok = raw || ((reinterpret_cast<uintptr_t>(row + T0) & 15) == 0);
    // per 128-point block a lane takes the times 2 lane, 2 lane + 1 and 64 + 2 lane, 65 + 2 lane: both 16-byte loads of
    // the warp are 512 contiguous bytes of shared memory (four consecutive times per lane would be two-way conflicted),
    // both stores 512 contiguous bytes of the row
    const double* src = lin + 2 * lane;
    double* dst = raw ? raw + 2 * lane : row + T0 + 2 * lane;
    double td = double(T0) + double(2 * lane);
#pragma unroll 1
    for (int b = 0; b < 2 * RH; ++b, src += kLinBlock, dst += 128, td += 128.0) {
        const double2 a01 = *reinterpret_cast<const double2*>(src);
        const double2 a23 = *reinterpret_cast<const double2*>(src + 64);
        const int o = 128 * b + 2 * lane;
        double v[4] = {a01.x, a01.y, a23.x, a23.y};
        if (!raw) {
            v[0] *= td; v[1] *= td + 1.0; v[2] *= td + 64.0; v[3] *= td + 65.0;
        }
        if ((raw || o + 65 < t_end) && vec_ok) {
            *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
            *reinterpret_cast<double2*>(dst + 64) = make_double2(v[2], v[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int oi = o + (i & 1) + 32 * (i & 2);
                if (oi < t_end) dst[(i & 1) + 32 * (i & 2)] = v[i];
            }
        }
        if (!raw) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int oi = o + (i & 1) + 32 * (i & 2);
                const long long bits = __double_as_longlong(v[i]);
                if (oi < t_end && bits > best_bits) {   // offsets rise within a lane: the first maximum is kept
                    best_bits = bits;
                    best_t = T0 + oi;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// PI profile, moment form
// ------------------------------------------------------------------------------------------
// Inside a 128-point block of the time tile, t = T0 + 128 hb + 64 + d, a slowly decaying site is a polynomial in d:
//   a q^t = [a e^{-c T0}] [q^{128 hb + 64}] e^{-c d},      e^{-c d} = sum_k (-64 c)^k / k! (d/64)^k.
// For 64 c <= 5 (rate < 0.0195: 99.8 % of the active sites of UCE-like data) 32 terms leave a remainder below 1.3e-11
// of the term (5^32/32! against e^{-5}; the alternating sum loses < 4 digits to cancellation, ~1e-12), so the site is
// represented by 32 moments  U'[k] = a e^{-c T0} (-64 c)^k  (the 1/k! is applied with the powers of d) and meets only NB = TT/128 columns
//   W'[hb] = q^{128 hb + 64}:      M[k][hb] = sum_s U'[s][k] W'[s][hb]      -- 4 x NT DMMAs per four sites (NT = NB/8),
// a quarter of profile_kernel's 4 x RH (RH = 4 NT), with a W table a quarter as long.  One product per item turns
// moments back into times,  out[128 hb + t] = sum_k ((t - 64)/64)^k M[k][hb]  (128 NT DMMAs, against ~1400 RH/4 for the
// sites of a 350-site locus), written through shared memory into profile_kernel's accumulator layout.  The few faster
// sites of the item are then added by profile_kernel's power-table contraction (second pass over the item's rates,
// skipped when there is none) and the epilogue is profile_kernel's.  All arithmetic is IEEE double; the result differs
// from profile_kernel's by the series remainder, the cancellation noted above and a different association of the
// same sums (observed <= 1e-12 relative; the parity gate is 1e-9).
// (KM = 16: sites with 64 c <= 1.25, rate < 0.0048828125, need only 16 moments for the same remainder -- half the
// DMMAs, multiplies and shared-memory traffic of the moment pass and of the moments -> times step; the host picks KM
// per batch from a sample of the rates, tapir_cuda.cu choose_moments)
// The slower a site, the longer the block on which KM terms suffice: sites below half the class bound are expanded on
// 256-point blocks (2,048-point tiles: one column tile of DMMAs instead of two), sites with (TT/2) c <= 1.25 on the
// whole tile with 16 terms, summed per lane in registers without any table or DMMA.  After the pass the coarser sums
// are shifted onto the 128-point blocks -- an exact change of variable of the truncated polynomials, a few DMMAs per
// item -- so everything downstream (moments -> times, the epoch functional) sees one set of moments.  DESIGN.md 3.1a.
// 1/k!, k = 0..31
__device__ __constant__ double c_inv_fact[32] = {
    1.0, 1.0, 0.5, 0.16666666666666666, 0.041666666666666664, 0.008333333333333333, 0.001388888888888889,
    0.0001984126984126984, 2.48015873015873e-05, 2.7557319223985893e-06, 2.755731922398589e-07,
    2.505210838544172e-08, 2.08767569878681e-09, 1.6059043836821613e-10, 1.1470745597729725e-11,
    7.647163731819816e-13, 4.779477332387385e-14, 2.8114572543455206e-15, 1.5619206968586225e-16,
    8.22063524662433e-18, 4.110317623312165e-19, 1.9572941063391263e-20, 8.896791392450574e-22,
    3.868170170630684e-23, 1.6117375710961184e-24, 6.446950284384474e-26, 2.4795962632247976e-27,
    9.183689863795546e-29, 3.279889237069838e-30, 1.1309962886447716e-31, 3.7699876288159054e-33,
    1.216125041553518e-34};

// the same table in global memory, for lane-varying indices (a constant-bank read with 32 different indices is serialised)
__device__ const double g_inv_fact[32] = {
    1.0, 1.0, 0.5, 0.16666666666666666, 0.041666666666666664, 0.008333333333333333, 0.001388888888888889,
    0.0001984126984126984, 2.48015873015873e-05, 2.7557319223985893e-06, 2.755731922398589e-07,
    2.505210838544172e-08, 2.08767569878681e-09, 1.6059043836821613e-10, 1.1470745597729725e-11,
    7.647163731819816e-13, 4.779477332387385e-14, 2.8114572543455206e-15, 1.5619206968586225e-16,
    8.22063524662433e-18, 4.110317623312165e-19, 1.9572941063391263e-20, 8.896791392450574e-22,
    3.868170170630684e-23, 1.6117375710961184e-24, 6.446950284384474e-26, 2.4795962632247976e-27,
    9.183689863795546e-29, 3.279889237069838e-30, 1.1309962886447716e-31, 3.7699876288159054e-33,
    1.216125041553518e-34};

// Code size of the 2,048-point kernels: they carry three classes of slow sites, and beyond ~2,500 instructions (40 KB)
// the SM's instruction cache misses and the GPC-level cache behind it saturates -- the kernel is then bound by
// instruction fetch (measured: hit rate 87 %, 10.2 ms; 99.7 %, 9.9 ms with less unrolling).  So their k-step loops are
// unrolled less than the compiler would (four-fold: ~250 instructions each); MomCfg::GroupUnroll, RolledContract.
#ifndef TAPIR_G8
#define TAPIR_G8 1
#endif
#ifndef TAPIR_ROLL8
#define TAPIR_ROLL8 1
#endif

// Shared memory of one warp of profile_mom_kernel: the moment tables U' [32][34] + W' [32][WA]; later, over them, the
// scratch of the moments -> times step ([32][NB+4] moments, then [TT] times) and the 16-site power tables of the
// fast-site pass; behind all of them the list of up to 32 fast rates (filled during the moment pass, read after the
// scratch has been used), the 64-entry compaction queue(s) and the slots of the rates in flight.  13.25 KB for RH = 4,
// KM = 32: sixteen warps per SM.
template <int RH, int KM = 32>
struct MomCfg {
    static constexpr int TT = ProfCfg<RH>::TT, NB = TT / 128, NT = NB / 8, WA = (NB == 8) ? 12 : NB + 2;
    static constexpr int MS = KM + 2;  // doubles per moment row (+ 2: rows 16 bytes off a multiple of 128, conflict-free 16-byte fragments)
    static constexpr int Tables = 32 * MS + 32 * WA;
    // moments -> times, the fast sites and the epilogue work on sub-tiles of 256 QR points, so that the accumulators of a
    // 2,048-point tile (128 registers) never exist at once: twelve warps per SM instead of eight
    static constexpr int QR = (RH == 8) ? 4 : RH, NH = RH / QR, QNB = 2 * QR;
    static constexpr int Lin = kLinBlock * QNB;   // moments -> times scratch, one padded row per 128-point block
    static constexpr int Pow = 16 * (kUStride + ProfCfg<QR>::WStride);   // 16-site power tables of the fast-site pass
    static constexpr int FastOff = (Tables > Lin ? Tables : Lin) > Pow ? (Tables > Lin ? Tables : Lin) : Pow;
    static constexpr int QueueOff = FastOff + 32;
    // 2,048-point tiles keep two classes of slow sites (see profile_mom_kernel): a second 64-entry ring for the coarse one
#if defined(TAPIR_NO_COARSE)   // measurement builds only (tools/gpu_variants.sh)
    static constexpr bool Coarse = false;
#else
    static constexpr bool Coarse = (NT == 2);
#endif
#if defined(TAPIR_NO_WHOLE)
    static constexpr bool Whole = false;
#else
    static constexpr bool Whole = true;
#endif
    // the ring of the whole-tile class behind them, then four slots of 32 rates for the rounds in flight from global
    // memory (cp.async, lane-private).  With eight-column W' tables (WA = 12) the slots are the four unused doubles at
    // the end of every W' row instead: those kernels run sixteen warps per SM and have no shared memory to spare.
    // Ring length: 2,048-point tiles run twelve warps per SM and have shared memory to spare -- rings of 256 let an
    // item of up to ~250 sites per class finish its rounds first and then drain class by class, every loop's code
    // entered once per item instead of once per group (the kernel is bound by instruction fetch)
#ifndef TAPIR_RING8
#define TAPIR_RING8 256
#endif
#ifndef TAPIR_WARPS8
#define TAPIR_WARPS8 12
#endif
    static constexpr int Ring = Coarse ? TAPIR_RING8 : 64;
    static constexpr int TileQOff = QueueOff + (Coarse ? 2 : 1) * Ring;
    static constexpr bool RateInPad = (WA == 12);
    static constexpr int RateOff = TileQOff + Ring;
    static constexpr int WarpDoubles = RateOff + (RateInPad ? 0 : 4 * 32);
    static constexpr size_t SmemBytes = size_t(kProfWarps) * WarpDoubles * sizeof(double);
    static constexpr int MinBlocks = (RH == 4 ? 14 : TAPIR_WARPS8) / kProfWarps;  // 144 / 168 registers
    static constexpr int GroupUnroll = (RH == 8) ? TAPIR_G8 : 4;           // pairs of k-steps per loop body of mom_group
    static constexpr bool RolledContract = (RH == 8) && TAPIR_ROLL8;       // fast-site contraction not unrolled
    static_assert(Pow <= FastOff, "16-site power tables fit below the fast-site list");
    static_assert(NT == 1 || NT == 2, "one or two column tiles");
    static_assert(KM * (NB + 4) + (Coarse ? KM * 12 : 0) + 32 <= FastOff, "moment scratch fits");
    static_assert(KM == 16 || KM == 32, "16 or 32 moments per block");
    static_assert((kLinBlock * 8) % 16 == 0, "16-byte reads of the scratch stay aligned");
};

// a slow site: 64 c = 256 r <= 5 (KM = 32: r < 0x3F94000000000000 = 0.01953125) or <= 1.25 (KM = 16: r <
// 0x3F74000000000000 = 0.0048828125), r >= 0 -- on the bit pattern; `bound_hi` is the high word of the bound
__device__ __forceinline__ bool rate_is_slow(double r, uint32_t bound_hi) { return uint32_t(__double2hiint(r)) < bound_hi; }
template <int KM>
__device__ __forceinline__ constexpr uint32_t mom_bound_hi() { return KM == 16 ? 0x3F740000u : 0x3F940000u; }

// Whole-tile class of profile_mom_kernel: V[n][j] = (2 j - 7)^n / n!, the shift of a polynomial on the tile (variable
// y in [-1, 1)) onto its eight equal parts (y = (x + 2 j - 7)/8); correctly rounded
__device__ const double g_tile_shift[16 * 8] = {
    1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0,
    -7.0, -5.0, -3.0, -1.0, 1.0, 3.0, 5.0, 7.0,
    24.5, 12.5, 4.5, 0.5, 0.5, 4.5, 12.5, 24.5,
    -57.166666666666664, -20.833333333333332, -4.5, -0.16666666666666666, 0.16666666666666666, 4.5, 20.833333333333332, 57.166666666666664,
    100.04166666666667, 26.041666666666668, 3.375, 0.041666666666666664, 0.041666666666666664, 3.375, 26.041666666666668, 100.04166666666667,
    -140.05833333333334, -26.041666666666668, -2.025, -0.008333333333333333, 0.008333333333333333, 2.025, 26.041666666666668, 140.05833333333334,
    163.4013888888889, 21.70138888888889, 1.0125, 0.001388888888888889, 0.001388888888888889, 1.0125, 21.70138888888889, 163.4013888888889,
    -163.4013888888889, -15.500992063492063, -0.43392857142857144, -0.0001984126984126984, 0.0001984126984126984, 0.43392857142857144, 15.500992063492063, 163.4013888888889,
    142.97621527777778, 9.68812003968254, 0.1627232142857143, 2.48015873015873e-05, 2.48015873015873e-05, 0.1627232142857143, 9.68812003968254, 142.97621527777778,
    -111.20372299382716, -5.3822889109347445, -0.05424107142857143, -2.7557319223985893e-06, 2.7557319223985893e-06, 0.05424107142857143, 5.3822889109347445, 111.20372299382716,
    77.84260609567902, 2.6911444554673722, 0.016272321428571428, 2.755731922398589e-07, 2.755731922398589e-07, 0.016272321428571428, 2.6911444554673722, 77.84260609567902,
    -49.53620387906846, -1.2232474797578965, -0.004437905844155844, -2.505210838544172e-08, 2.505210838544172e-08, 0.004437905844155844, 1.2232474797578965, 49.53620387906846,
    28.896118929456602, 0.5096864498991235, 0.001109476461038961, 2.08767569878681e-09, 2.08767569878681e-09, 0.001109476461038961, 0.5096864498991235, 28.896118929456602,
    -15.559448654322786, -0.19603324996120136, -0.00025603302947052944, -1.6059043836821613e-10, 1.6059043836821613e-10, 0.00025603302947052944, 0.19603324996120136, 15.559448654322786,
    7.779724327161393, 0.07001187498614334, 5.486422060082774e-05, 1.1470745597729725e-11, 1.1470745597729725e-11, 5.486422060082774e-05, 0.07001187498614334, 7.779724327161393,
    -3.6305380193419836, -0.02333729166204778, -1.0972844120165549e-05, -7.647163731819816e-13, 7.647163731819816e-13, 1.0972844120165549e-05, 0.02333729166204778, 3.6305380193419836,
};

// v[0..15] of every lane -> the 16 sums over the warp; lane L returns the sum of index L >> 1 (halving exchange:
// 16 double shuffles and 16 additions instead of 16 x 5)
__device__ __forceinline__ double transpose_reduce16(const double (&v)[16], int lane)
{
    double a[8], b[4], c[2];
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (h4 ? v[8 + i] : v[i]) + __shfl_xor_sync(0xffffffffu, h4 ? v[i] : v[8 + i], 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = (h3 ? a[4 + i] : a[i]) + __shfl_xor_sync(0xffffffffu, h3 ? a[i] : a[4 + i], 8);
#pragma unroll
    for (int i = 0; i < 2; ++i) c[i] = (h2 ? b[2 + i] : b[i]) + __shfl_xor_sync(0xffffffffu, h2 ? b[i] : b[2 + i], 4);
    const double d = (h1 ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, h1 ? c[0] : c[1], 2);
    return d + __shfl_xor_sync(0xffffffffu, d, 1);
}

// Position of column hb of an eight-column W' table inside rows that are 18 doubles apart (the stride of the
// sixteen-column table of the same kernel): 0 1 8 9 2 3 10 11.  The 8-byte fragment loads of a half warp (four rows fk,
// four columns) then fall on sixteen different banks -- in natural order rows 2 banks apart collide two-way -- and the
// pairs (hb, hb + 1) stay adjacent for the 16-byte stores.
template <bool Swz>
__device__ __forceinline__ constexpr int wcol(int hb)
{
    return Swz ? ((hb & 1) | ((hb & 2) << 2) | ((hb & 4) >> 1)) : hb;
}


// One group of up to 32 slow sites of profile_mom_kernel's moment pass, for a class with NG column tiles: blocks of
// TT / (8 NG) points (half-length HB), the group's tables U' [32][MS] (moments u0 z^k, z = -HB c) and W' [32][WA]
// (q^HB (q^2HB)^hb) built in the warp's shared memory, then  acE/acO[i][j][c] += U'^T W'  on the FP64 MMA path.
template <int KM, int NG, int TT, int WA, int MS, int UNROLL>
__device__ __forceinline__ void mom_group(const double lam, const int cnt, const int lane, const int64_t T0, const double t0d,
                                          double* myM, double* myWA, const double* fragM, const double* WAs,
                                          double (&acE)[KM / 8][NG][2], double (&acO)[KM / 8][NG][2])
{
    constexpr int MI = KM / 8;
    constexpr int GB = 8 * NG;                   // blocks of this class per tile
    constexpr double HB = double(TT / GB / 2);   // 64 (fine) or 128 (coarse)
    constexpr bool Swz = (NG == 1 && WA == 18);  // eight columns in rows made for sixteen: see wcol
    const int fk = lane & 3, fg = lane >> 2;
    {
        // rows of lanes >= cnt: moments 0, W' = 1
        const bool ok = lane < cnt;
        const double c = ok ? 4.0 * lam : 0.0;
        const double a = ok ? 16.0 * (lam * lam) : 0.0;
        const double u0 = (T0 == 0) ? a : a * exp_full(-(c * t0d));
        // q^HB directly (HB c inside the class bound): one bounded exponential instead of exp(-c) and squarings
        const double z = -HB * c;
        const double q1 = exp_bounded_short(z), q2 = q1 * q1, q4 = q2 * q2;
        // moments u0 z^k (the 1/k! lives in the moments -> times step): four chains of stride 4
        const double z2 = z * z, z4 = z2 * z2;
        double m0 = u0, m1 = u0 * z, m2 = u0 * z2, m3 = m1 * z2;
#pragma unroll
        for (int k = 0; k < KM / 4; ++k) {
            st_shared_v2(myM + 4 * k, m0, m1);
            st_shared_v2(myM + 4 * k + 2, m2, m3);
            m0 *= z4; m1 *= z4; m2 *= z4; m3 *= z4;
        }
        // W'[hb] = q^HB (q^(2 HB))^hb, two chains of stride 4; column hb at position wcol<Swz>(hb)
        const double q8 = q4 * q4;
        double w = q1, v = q1 * q4;
#pragma unroll
        for (int k = 0; k < GB / 4; ++k) {
            st_shared_v2(myWA + wcol<Swz>(4 * k), w, w * q2);
            st_shared_v2(myWA + wcol<Swz>(4 * k + 2), v, v * q2);
            w *= q8; v *= q8;
        }
    }
    __syncwarp();
    {
        const double* fragWG = WAs + fk * WA + (Swz ? wcol<true>(fg) : NG * w_chunk<NG>(fg));
        auto fetch = [&](double (&a)[MI], double (&b)[NG], int step) {
            const double* mp = fragM + step * (4 * MS);
            const double2 m01 = *reinterpret_cast<const double2*>(mp);
            a[0] = m01.x; a[1] = m01.y;
            if constexpr (MI == 4) {
                const double2 m23 = *reinterpret_cast<const double2*>(mp + 2);
                a[2] = m23.x; a[3] = m23.y;
            }
            const double* wp = fragWG + step * (4 * WA);
            if constexpr (NG == 1) {
                b[0] = wp[0];
            } else {
#pragma unroll
                for (int r = 0; r < NG / 2; ++r) {
                    const double2 v = *reinterpret_cast<const double2*>(wp + 2 * r);
                    b[2 * r] = v.x; b[2 * r + 1] = v.y;
                }
            }
        };
        auto mma_step = [&](double (&ac)[MI][NG][2], const double (&a)[MI], const double (&b)[NG]) {
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NG; ++j) dmma884(ac[i][j][0], ac[i][j][1], a[i], b[j]);
        };
        const int nk = (cnt + 3) >> 2;  // rows past cnt are zero, so an odd count rounds up
        double a0[MI], b0[NG], a1[MI], b1[NG];
        fetch(a0, b0, 0);
#pragma unroll(UNROLL)
        for (int st = 0; st < nk; st += 2) {
            fetch(a1, b1, st + 1);
            mma_step(acE, a0, b0);
            fetch(a0, b0, min(st + 2, 7));
            mma_step(acO, a1, b1);
        }
    }
    __syncwarp();
}

// TRACE (measurement only): [0] SM id, [1] entry, [12] exit, [13] items done; for the warp's second item
// [16] claimed, [17] moment pass done, [18] moments -> times done, [19] fast-site pass done, [20] epilogue done,
// [22] groups of the moment pass, [24] n_fast
template <int RH, bool TRACE = false, int KM = 32>
__global__ void __launch_bounds__(kProfWarps * 32, MomCfg<RH, KM>::MinBlocks) profile_mom_kernel(const ProfArgs args)
{
    using Cfg = ProfCfg<RH>;
    using Mom = MomCfg<RH, KM>;
    constexpr int TT = Cfg::TT, WS = ProfCfg<Mom::QR>::WStride, US = kUStride, MS = Mom::MS;
    constexpr int MI = KM / 8;   // moments per fragment row (moment = MI * mrow(fg) + i)
    constexpr int KS = KM / 4;   // k-steps of the moments -> times product
    constexpr uint32_t kBound = mom_bound_hi<KM>();
    static_assert(RH >= 4, "moment form: tiles of 1024 points or more");
    constexpr int NB = Mom::NB, NT = Mom::NT, WA = Mom::WA;
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int fk = lane & 3, fg = lane >> 2;
    double* Us = smem + size_t(warp) * Mom::WarpDoubles;
    double* Qs = Us + Mom::QueueOff;
    // fast-site pass: power tables of 16 sites in the sub-tile's geometry, U [16][US] then W [16][WS]
    constexpr int QR = Mom::QR, NH = Mom::NH, QTT = 256 * QR;
    double* Ws = Us + 16 * US;
    double* myU = Us + (lane & 15) * US;
    double* myW = Ws + (lane & 15) * WS;
    const double* fragU = Us + fk * US + 4 * chunk_pair(fg);
    const double* fragW = Ws + fk * WS + QR * w_chunk<QR>(fg);
    // moment pass: U' [32][MS] where profile_kernel keeps U (same stride), W' [32][WA] behind it
    double* myM = Us + lane * MS;
    double* WAs = Us + 32 * MS;
    double* myWA = WAs + lane * WA;
    // 16-byte chunks of a moment row, permuted so that the eight lanes of a quarter warp (four rows fk, two values of fg)
    // read eight different bank groups: rows are 8 + 1 (MI = 2) or 16 + 1 (MI = 4) chunks apart
    const int mrow = (MI == 4) ? chunk_pair(fg) : chunk_quad(fg);
    const double* fragM = Us + fk * MS + MI * mrow;
    const double* fragWA = WAs + fk * WA + NT * w_chunk<NT>(fg);
    double* Fs = Us + Mom::FastOff;  // up to 32 fast rates of the item, behind everything that is written before they are read
    long long* tr = nullptr;
    int n_done = 0, n_groups = 0;
    if constexpr (TRACE) {
        tr = args.trace + (size_t(blockIdx.x) * kProfWarps + warp) * 32;
        if (lane == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            tr[0] = smid;
            tr[1] = clock64();
        }
    }
    auto stamp = [&](int slot) {
        if constexpr (TRACE) {
            __syncwarp();
            if (lane == 0 && n_done == 1) tr[slot] = clock64();
        }
    };

    // TRACE: cycles of the second item spent in [0] the groups of the moment pass, [1] moments -> times, [2] the fast
    // sites, [3] the epilogue (sums over groups / sub-tiles; words 25..28)
    long long t_sum[4] = {0, 0, 0, 0}, t_mark = 0;
    auto tick = [&]() {
        if constexpr (TRACE) {
            __syncwarp();
            t_mark = clock64();
        }
    };
    auto tock = [&](int k) {
        if constexpr (TRACE) {
            __syncwarp();
            if (n_done == 1) t_sum[k] += clock64() - t_mark;
        }
    };

    for (;;) {
        uint32_t it = 0;
        if (lane == 0) it = atomicAdd(args.counter, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= uint32_t(args.n_items)) break;
        const ProfItem item = args.items[it];
        const int64_t T0 = int64_t(item.tile) * TT;
        const double t0d = double(T0);
        const double* rp = args.rates + item.site_begin;
        stamp(16);
        n_groups = 0;

        // ---------------- moment pass: slow sites ----------------
        // accA[i][j][c]: moment = MI mrow + i, hb = NT w_chunk<NT>(2 fk + c) + j
        // (two sets, for the even and the odd k-steps: with NT = 1 a single set is a dependent chain of four DMMAs)
        // 2,048-point tiles (Coarse): a site with 128 c inside the class bound -- half the rate bound -- is expanded on
        // the eight 256-point blocks of the tile instead (accP/accQ, the same KM moments in z = -128 c: one column tile,
        // half the DMMAs and half the W' table); the coarse moments are shifted onto the 128-point blocks after the pass.
        constexpr bool Coarse = Mom::Coarse;
        double accA[MI][NT][2], accB[MI][NT][2];
        double accP[MI][1][2], accQ[MI][1][2];   // unused (and removed) without the coarse class
#pragma unroll
        for (int i = 0; i < MI; ++i) {
#pragma unroll
            for (int j = 0; j < NT; ++j) accA[i][j][0] = accA[i][j][1] = accB[i][j][0] = accB[i][j][1] = 0.0;
            accP[i][0][0] = accP[i][0][1] = accQ[i][0][0] = accQ[i][0][1] = 0.0;
        }
        // Whole-tile class: a site with (TT/2) c <= 1.25 (rate < 0.0048828125 x 128/TT) is one polynomial of 16 terms on
        // the whole tile; its moments u z^k, z = -(TT/2) c, u = a e^{-c (T0 + TT/2)}, are summed per lane in registers --
        // no table, no DMMA -- and shifted onto the tile's eight parts after the pass.
        constexpr uint32_t kBoundTile = 0x3F740000u - (TT == 2048 ? 4u : 3u) * 0x00100000u;
        double tileS[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) tileS[k] = 0.0;
        int n_fast = 0;
        bool any_coarse = false, any_tile = false;
        {
            int qh = 0, qn = 0, ph = 0, pn = 0, ah = 0, an = 0;   // rings: fine class (Qs), coarse class (Ps), whole tile (As)
            constexpr int RM = Mom::Ring - 1;
            double* Ps = Qs + Mom::Ring;
            double* As = Us + Mom::TileQOff;
            // Rates are fetched three rounds ahead, asynchronously into lane-private shared-memory slots (a register
            // queue must be shifted every round, and the shift waits for the load issued one round earlier: rounds
            // without a group in between then expose the whole DRAM latency).  One commit group per round, issued or
            // empty, so that "all but the newest two" always means this round's rates have landed; four slots, so the
            // slot a round refills is never the one it has just read.
            constexpr int kAhead = 3;
            // this lane's four slots: 32 doubles apart, or -- in the W' padding, rows 12 doubles apart -- rotated by
            // lane / 4 so that the sixteen lanes of a half warp touch sixteen different banks
            double* ring = Mom::RateInPad ? myWA + 8 : Us + Mom::RateOff + lane;
            auto slot_of = [&](int r) { return Mom::RateInPad ? ((r + (lane >> 2)) & 3) : 32 * (r & 3); };
            auto fetch_round = [&](int r) {   // rates [32 r, 32 r + 32) of the item -> slot r mod 4; zeros past the item
                const int i = 32 * r + lane;
                const bool in = i < item.n_sites;
                const unsigned dst = unsigned(__cvta_generic_to_shared(ring + slot_of(r)));
                const double* src = in ? rp + i : rp;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(in ? 8 : 0) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
#pragma unroll
            for (int k = 0; k < kAhead; ++k) fetch_round(k);
            for (int base = 0, rnd = 0;; base += 32, ++rnd) {
                const bool more = base < item.n_sites;
                if (more) {
                    asm volatile("cp.async.wait_group %0;" ::"n"(kAhead - 1) : "memory");
                    const double lam_new = ring[slot_of(rnd)];
                    fetch_round(rnd + kAhead);
                    const bool act = rate_is_active(lam_new);  // lanes past the item hold 0
                    const bool slow = act && rate_is_slow(lam_new, kBound);
                    const bool whole = Mom::Whole && act && rate_is_slow(lam_new, kBoundTile);
                    const bool coarse = Coarse && act && !whole && rate_is_slow(lam_new, kBound - 0x00100000u);   // half the bound
                    const unsigned ballot = __ballot_sync(0xffffffffu, slow && !coarse && !whole);
                    {   // the first 32 fast sites are kept aside, so that their pass need not stream the item again
                        const unsigned bf = __ballot_sync(0xffffffffu, act && !slow);
                        const int pos = n_fast + __popc(bf & ((1u << lane) - 1u));
                        if (act && !slow && pos < 32) Fs[pos] = lam_new;
                        n_fast += __popc(bf);
                    }
                    if (slow && !coarse && !whole) Qs[(qn + __popc(ballot & ((1u << lane) - 1u))) & RM] = lam_new;
                    qn += __popc(ballot);
                    if constexpr (Mom::Whole) {
                        const unsigned ba = __ballot_sync(0xffffffffu, whole);
                        if (whole) As[(an + __popc(ba & ((1u << lane) - 1u))) & RM] = lam_new;
                        an += __popc(ba);
                    }
                    if constexpr (Coarse) {
                        const unsigned bc = __ballot_sync(0xffffffffu, coarse);
                        if (coarse) Ps[(pn + __popc(bc & ((1u << lane) - 1u))) & RM] = lam_new;
                        pn += __popc(bc);
                    }
                    __syncwarp();
                }
                if (Mom::Whole && (an - ah > Mom::Ring - 32 || !more))
                while (an - ah >= 32 || (!more && an > ah)) {
                    const int cnt = min(32, an - ah);
                    const bool ok = lane < cnt;
                    const double lam = ok ? As[(ah + lane) & RM] : 0.0;
                    ah += cnt;
                    any_tile = true;
                    tick();
                    const double c = 4.0 * lam;               // lanes >= cnt: lam = 0, a = 0
                    const double a = 16.0 * (lam * lam);
                    const double z = -double(TT / 2) * c;
                    const double m = (T0 == 0) ? a * exp_bounded_short(z) : a * exp_full(-(c * (t0d + double(TT / 2))));
                    // u z^k: four chains of stride 4 (depth 5 instead of 16)
                    const double z2 = z * z, z4 = z2 * z2;
                    double m0 = m, m1 = m * z, m2 = m * z2, m3 = m1 * z2;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        tileS[4 * k] += m0; tileS[4 * k + 1] += m1; tileS[4 * k + 2] += m2; tileS[4 * k + 3] += m3;
                        if (k < 3) { m0 *= z4; m1 *= z4; m2 *= z4; m3 *= z4; }
                    }
                    tock(0);
                }
                if constexpr (Coarse) {
                    if (pn - ph > Mom::Ring - 32 || !more)
                    while (pn - ph >= 32 || (!more && pn > ph)) {
                        const int cnt = min(32, pn - ph);
                        const double lam = (lane < cnt) ? Ps[(ph + lane) & RM] : 0.0;
                        ph += cnt;
                        any_coarse = true;
                        ++n_groups;
                        tick();
                        mom_group<KM, 1, TT, WA, MS, Mom::GroupUnroll>(lam, cnt, lane, T0, t0d, myM, myWA, fragM, WAs, accP, accQ);
                        tock(0);
                    }
                }
                if (qn - qh > Mom::Ring - 32 || !more)
                while (qn - qh >= 32 || (!more && qn > qh)) {
                    const int cnt = min(32, qn - qh);
                    const double lam = (lane < cnt) ? Qs[(qh + lane) & RM] : 0.0;
                    qh += cnt;
                    ++n_groups;
                    tick();
                    mom_group<KM, NT, TT, WA, MS, Mom::GroupUnroll>(lam, cnt, lane, T0, t0d, myM, myWA, fragM, WAs, accA, accB);
                    tock(0);
                }
                if (!more) break;
            }
        }
        stamp(17);

        // ---------------- moments -> B fragments, the slow sites' epoch sums ----------------
        double bf[KS][NT];  // B fragments: M[4 ks + fk][8 nt + fg] / k!
        {
            constexpr int SA = NB + 4;
            double* MA = Us;   // [KM][SA]
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        MA[(MI * mrow + i) * SA + NT * w_chunk<NT>(2 * fk + c) + j] = accA[i][j][c] + accB[i][j][c];
            // Whole-tile sums onto the tile's eight parts (256-point blocks of a 2,048-point tile: the coarse moments;
            // 128-point blocks of a 1,024-point tile: the moments themselves).  With y = (x + o_j)/8, o_j = 2 j - 7,
            //   N[m][j] += sum_n T[m][n] V[n][j],   T[m][n] = 8^-(m+n) S[m+n] (zero from m + n = 16),  V[n][j] = o_j^n / n!
            // -- six DMMAs (the zero blocks of the Hankel matrix T are skipped).
            constexpr int SBc = 12;
            double* MBc = MA + KM * SA;                          // [KM][SBc], coarse moments (Coarse only)
            double* TS = MBc + (Coarse ? KM * SBc : 0);          // [32]: scaled sums, zero-padded
            if (any_tile) {
                const double tot = transpose_reduce16(tileS, lane);
                const int k = lane >> 1;
                if (lane & 1) TS[16 + k] = 0.0;                                        // odd lanes: the padding
                else TS[k] = tot * __hiloint2double((1023 - 3 * k) << 20, 0);         // 8^-k, exact
            }
            if constexpr (Coarse) {
                if (any_coarse || any_tile) {
#pragma unroll
                    for (int i = 0; i < MI; ++i)
#pragma unroll
                        for (int c = 0; c < 2; ++c) MBc[(MI * mrow + i) * SBc + 2 * fk + c] = accP[i][0][c] + accQ[i][0][c];
                }
            }
            __syncwarp();
            if (any_tile) {
                double* dstN = Coarse ? MBc : MA;
                constexpr int SN = Coarse ? SBc : SA;
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    double d[2] = {0.0, 0.0};
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        if (8 * mt + 4 * ks >= 16) continue;   // T[m][n] = 0 from m + n = 16
                        const double tv = TS[8 * mt + fg + 4 * ks + fk];
                        dmma884(d[0], d[1], tv, __ldg(g_tile_shift + 8 * (4 * ks + fk) + fg));
                    }
#pragma unroll
                    for (int c = 0; c < 2; ++c) dstN[(8 * mt + fg) * SN + 2 * fk + c] += d[c];
                }
                __syncwarp();
            }
            if constexpr (Coarse) {
                // Coarse moments onto the fine blocks.  A 256-point block (variable y = (t - centre)/128) holds the
                // polynomial sum_k N[k]/k! y^k; on its two halves y = (x -+ 1)/2, so the 128-point blocks 2 j + s receive
                //   M[m][2 j + s] += sum_{k >= m} S_s[m][k] N[k][j],   S_s[m][k] = 2^-k (-+1)^(k-m) / (k-m)!
                // -- an exact change of variable of the (truncated) polynomial, one small product per half on the
                // moments' own fragment layout: rows m, columns j, the upper-triangular S as the A operand.
                if (any_coarse || any_tile) {
                    double nb[KS];
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) nb[ks] = MBc[(4 * ks + fk) * SBc + fg];
#pragma unroll
                    for (int mt = 0; mt < MI; ++mt) {
                        double dl[2] = {0.0, 0.0}, dr[2] = {0.0, 0.0};
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            if (4 * ks + 3 < 8 * mt) continue;   // S is upper triangular
                            const int k = 4 * ks + fk, n = k - (8 * mt + fg);
                            // 2^-k / n!, zero below the diagonal
                            const double sr = (n < 0) ? 0.0 : g_inv_fact[n] * __hiloint2double((1023 - k) << 20, 0);
                            const double sl = (n & 1) ? -sr : sr;
                            dmma884(dl[0], dl[1], sl, nb[ks]);
                            dmma884(dr[0], dr[1], sr, nb[ks]);
                        }
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            double* dst = MA + (8 * mt + fg) * SA + 2 * (2 * fk + c);
                            dst[0] += dl[c];
                            dst[1] += dr[c];
                        }
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) bf[ks][nt] = MA[(4 * ks + fk) * SA + 8 * nt + fg];
            // Epoch integrals of the item's slow sites: for every epoch whose centre lies in this tile the sum of the
            // sites' first GK21 panels is a linear functional of the moments of the block around the centre
            // (mom_epoch.h),  I[e] = sum_k p[e][k] M[k][hb_e].  Eight epochs at a time as one more product on the B
            // fragments that are in registers anyway: A = p (row fg = epoch, column 4 ks + fk), eight DMMAs per column
            // tile; the lane that holds column hb_e of row e stores it.  The order of the additions is the MMA's, fixed.
            for (int g0 = 0; g0 < args.n_mom_epochs; g0 += 8) {
                const bool valid = g0 + fg < args.n_mom_epochs;
                MomEpochDev ep;
                ep.col = 0; ep.tile = -1; ep.hb = 0; ep.pad_ = 0;
                if (valid) ep = args.mom_epochs[g0 + fg];
                const double* pc = args.mom_p + 32 * ep.col + fk;
                double pa[KS];
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) pa[ks] = valid ? __ldg(pc + 4 * ks) : 0.0;
                double dd[NT][2];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) dd[nt][0] = dd[nt][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) dmma884(dd[nt][0], dd[nt][1], pa[ks], bf[ks][nt]);
                if (valid && ep.tile == item.tile && ((ep.hb & 7) >> 1) == fk) {
                    double v = 0.0;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        if ((ep.hb >> 3) == nt) v = (ep.hb & 1) ? dd[nt][1] : dd[nt][0];
                    args.mom_int[size_t(item.part) * args.e_total + ep.col] = v;
                }
            }
            // the 1/k! of the Taylor terms goes onto the moments (8 NT multiplies per lane), so that the A fragments
            // below are pure powers of x
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const double f = c_inv_fact[4 * ks + fk];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) bf[ks][nt] *= f;
            }
            __syncwarp();
        }
        stamp(18);

        // ---------------- per sub-tile of 256 QR points: moments -> times, the fast sites, the epilogue ----------------
        long long best_bits = -1;
        int64_t best_t = T0;
#pragma unroll 1
        for (int h = 0; h < NH; ++h) {
            const int64_t T0q = T0 + int64_t(h) * QTT;
            const double t0q = double(T0q);
            tick();
            {
                double* lin = Us;  // [QNB][kLinBlock]: natural layout [128 hb + t]
                double bq[KS];     // this sub-tile's column tile of the B fragments
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) bq[ks] = (NT == 1 || h == 0) ? bf[ks][0] : bf[ks][NT - 1];
                // rows t = 8 mt + fg of the block: A fragment of step ks is x^(4 ks + fk), x = (t - 64)/64 (exact), one
                // multiply per step; eight row tiles at a time, so that the DMMAs of a tile are not one dependent chain
                // (x^fk and x^4 come from an 8 KB table that stays in L1, args.mt_pow: forming them in registers took an
                // int -> double conversion, three multiplies and three selects per row tile)
                const double2* xp = reinterpret_cast<const double2*>(args.mt_pow) + lane;
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    double pw[8], x4[8], d[8][2];
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        const double2 v = __ldg(xp + 32 * (8 * half + m));
                        pw[m] = v.x;
                        x4[m] = v.y;
                        d[m][0] = d[m][1] = 0.0;
                    }
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                        for (int m = 0; m < 8; ++m) {
                            dmma884(d[m][0], d[m][1], pw[m], bq[ks]);
                            if (ks + 1 < KS) pw[m] *= x4[m];  // -> x^(k+4), k = 4 ks + fk
                        }
                    }
#pragma unroll
                    for (int m = 0; m < 8; ++m)
#pragma unroll
                        for (int c = 0; c < 2; ++c) lin[kLinBlock * (2 * fk + c) + 8 * (8 * half + m) + fg] = d[m][c];
                }
                __syncwarp();
            }
            tock(1);
            tick();
            // the item's fast sites, power tables as in profile_kernel (sub-tile geometry), sixteen at a time in their
            // order in the item: from the side list, or -- more than 32 of them -- compacted from a second stream of the
            // item's rates.  One build-and-contract site for both.  The sums go through profile_kernel's accumulator
            // layout: read from the scratch (the tables overwrite it), contracted onto, written back.
            if (n_fast > 0) {
                double* lin = Us;
                double acc[4][QR][2];
#pragma unroll
                for (int j = 0; j < QR; ++j)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int hi = QR * w_chunk<QR>(2 * fk + c) + j;
                        const double* src = lin + kLinBlock * (hi >> 2) + 32 * (hi & 3) + 4 * chunk_pair(fg);
                        const double2 v01 = *reinterpret_cast<const double2*>(src);
                        const double2 v23 = *reinterpret_cast<const double2*>(src + 2);
                        acc[0][j][c] = v01.x; acc[1][j][c] = v01.y; acc[2][j][c] = v23.x; acc[3][j][c] = v23.y;
                    }
                __syncwarp();
                const bool listed = n_fast <= 32;
                int qh = 0, qn = 0, base = 0;
                double lam_in = 0.0, lam_in2 = 0.0;
                if (!listed) {
                    lam_in = (lane < item.n_sites) ? __ldg(rp + lane) : 0.0;
                    lam_in2 = (32 + lane < item.n_sites) ? __ldg(rp + 32 + lane) : 0.0;
                }
                for (int g0 = 0; g0 < n_fast; g0 += 16) {
                    const int cnt = min(16, n_fast - g0);
                    if (!listed) {
                        while (qn - qh < cnt) {   // n_fast counted these sites: the stream holds them
                            const double lam_new = lam_in;
                            lam_in = lam_in2;
                            {
                                const int nb = base + 64 + lane;
                                lam_in2 = (nb < item.n_sites) ? __ldg(rp + nb) : 0.0;
                            }
                            base += 32;
                            const bool fast = rate_is_active(lam_new) && !rate_is_slow(lam_new, kBound);
                            const unsigned ballot = __ballot_sync(0xffffffffu, fast);
                            if (fast) Qs[(qn + __popc(ballot & ((1u << lane) - 1u))) & 63] = lam_new;
                            qn += __popc(ballot);
                            __syncwarp();
                        }
                    }
                    const double* from = listed ? Fs + g0 : Qs;
                    const int at = listed ? lane : (qh + lane) & 63;
                    const double lam = (lane < cnt) ? from[at] : 0.0;
                    qh += cnt;
                    __syncwarp();
                    if (lane < 16) prof_build_tables<QR>(lam, lane < cnt, T0q, t0q, myU, myW);
                    __syncwarp();
                    prof_contract<QR, Mom::RolledContract>(acc, fragU, fragW, (cnt + 3) >> 2, 3);
                    __syncwarp();
                }
#pragma unroll
                for (int j = 0; j < QR; ++j)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int hi = QR * w_chunk<QR>(2 * fk + c) + j;
                        double* d = lin + kLinBlock * (hi >> 2) + 32 * (hi & 3) + 4 * chunk_pair(fg);
                        st_shared_v2(d, acc[0][j][c], acc[1][j][c]);
                        st_shared_v2(d + 2, acc[2][j][c], acc[3][j][c]);
                    }
                __syncwarp();
            }
            tock(2);
            tick();
            prof_epilogue_lin<QR>(args, Us, args.pi_net + int64_t(item.locus) * args.T,
                                  item.slot >= 0 ? args.ws + size_t(item.slot) * TT + size_t(h) * QTT : nullptr, T0q, lane,
                                  best_bits, best_t);
            __syncwarp();
            tock(3);
        }
        stamp(19);
        if (item.slot < 0) {   // the tile's maximum over its sub-tiles and lanes (first maximum, as numpy.argmax)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const long long ob = __shfl_xor_sync(0xffffffffu, best_bits, off);
                const int64_t ot = __shfl_xor_sync(0xffffffffu, best_t, off);
                if (ob > best_bits || (ob == best_bits && ot < best_t)) { best_bits = ob; best_t = ot; }
            }
            if (lane == 0) {
                const int64_t k = int64_t(item.locus) * args.n_tiles + item.tile;
                args.tile_max[k] = __longlong_as_double(best_bits);
                args.tile_arg[k] = best_t;
            }
        }
        stamp(20);
        if constexpr (TRACE) {
            if (lane == 0 && n_done == 1) {
                tr[22] = n_groups; tr[24] = n_fast;
                tr[25] = t_sum[0]; tr[26] = t_sum[1]; tr[27] = t_sum[2]; tr[28] = t_sum[3];
            }
            ++n_done;
        }
    }
    if constexpr (TRACE) {
        if (lane == 0) { tr[12] = clock64(); tr[13] = n_done; }
    }
}

// ------------------------------------------------------------------------------------------
// PI profile, FP32 mode (TAPIR_FP32_MODE; BASELINE: "FP32 with compensated summation ... <= 1e-5 in FP32 mode")
// ------------------------------------------------------------------------------------------
// The moment form again, in single precision and without any shared-memory tables: lane = site.  A slow site
// (64 c <= 2.5, rate < 0.009765625: 16 Taylor moments per 128-point block leave a remainder of 1.3e-6, inside this mode's
// 1e-5 gate) forms its 16 moments u0 z^k and its 8 block weights q^(128 hb + 64) in registers and adds their outer
// product to 128 per-lane accumulators -- 128 independent FFMAs per site on the FP32 pipe (128 lanes per SM and clock
// against the FP64 pipe's 64), no tensor-core fragment layout to feed.  Summation: every lane sums its own sites
// (a 32nd of the locus) and a shuffle transpose-reduction tree closes the 128 sums over the lanes, so the error grows
// with log2(32) + S/32 additions, not with S (the "compensated summation" the north star asks of an FP32 mode; measured
// error ~1e-6).  The moments go back to times by a 16-term Horner evaluation per grid point (lane = grid point), faster
// sites are added as they are met by the direct power form a q^(T0 + lane) (q^32)^j on the same 32 per-lane outputs, and
// the epilogue is the FP64 kernels' (values leave as float64, times t, tile maximum).  The epoch sums of the slow
// sites are the same linear functional of the moments (mom_epoch.h with 16 terms and the class bound 2.5), evaluated
// in double on the FP32 moments.  One-warp CTAs, 1,024-point tiles only.
constexpr int kF32Terms = 16;
constexpr uint32_t kF32BoundHi = 0x3F840000u;  // high word of 0.009765625 = 2.5 / 256

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// sum over the 32 lanes of v[lane'] for every lane' at once: lane L returns the total of element L (31 shuffles)
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane)
{
#pragma unroll
    for (int n = 16; n >= 1; n >>= 1) {
        const bool up = (lane & n) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float send = up ? v[i] : v[i + n];
            const float keep = up ? v[i + n] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, n);
        }
    }
    return v[0];
}

__global__ void __launch_bounds__(32, 8) profile_f32_kernel(const ProfArgs args)
{
    constexpr int TT = 1024, NB = 8, KM = kF32Terms;
    __shared__ __align__(16) double s_q[64];     // compaction queue of the slow rates
    __shared__ __align__(16) float s_m[NB * KM]; // moments [hb][k]
    const int lane = threadIdx.x & 31;
    constexpr float kLog2e = 1.4426950408889634f;

    for (;;) {
        uint32_t it = 0;
        if (lane == 0) it = atomicAdd(args.counter, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= uint32_t(args.n_items)) break;
        const ProfItem item = args.items[it];
        const int64_t T0 = int64_t(item.tile) * TT;
        const double t0d = double(T0);
        const double* rp = args.rates + item.site_begin;

        float acc[KM][NB];   // per-lane sums of m_k w_hb over this lane's slow sites
        float o[32];         // outputs of this lane: grid point T0 + 32 j + lane, value / t
#pragma unroll
        for (int k = 0; k < KM; ++k)
#pragma unroll
            for (int hb = 0; hb < NB; ++hb) acc[k][hb] = 0.0f;
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = 0.0f;

        int qh = 0, qn = 0;
        constexpr int kAhead = 4;
        double lam_q[kAhead];
#pragma unroll
        for (int k = 0; k < kAhead; ++k) lam_q[k] = (32 * k + lane < item.n_sites) ? __ldg(rp + 32 * k + lane) : 0.0;
        for (int base = 0;; base += 32) {
            const bool more = base < item.n_sites;
            if (more) {
                const double lam_new = lam_q[0];
#pragma unroll
                for (int k = 0; k + 1 < kAhead; ++k) lam_q[k] = lam_q[k + 1];
                {
                    const int nb = base + 32 * kAhead + lane;
                    lam_q[kAhead - 1] = (nb < item.n_sites) ? __ldg(rp + nb) : 0.0;
                }
                const bool act = rate_is_active(lam_new);  // lanes past the item hold 0
                const bool slow = act && rate_is_slow(lam_new, kF32BoundHi);
                const unsigned ballot = __ballot_sync(0xffffffffu, slow);
                unsigned bfast = __ballot_sync(0xffffffffu, act && !slow);
                if (slow) s_q[(qn + __popc(ballot & ((1u << lane) - 1u))) & 63] = lam_new;
                qn += __popc(ballot);
                // faster sites, one at a time for the whole warp: lane = offset inside a 32-point run,
                // a q^t = [a q^(T0 + lane)] (q^32)^j
                while (bfast) {
                    const int src = __ffs(bfast) - 1;
                    bfast &= bfast - 1;
                    // (clamped: 16 r^2 and r t stay finite in single precision; such a site is 0 from t = 1 on either way)
                    const float lamf = fminf(fmaxf(float(__shfl_sync(0xffffffffu, lam_new, src)), -1.0e18f), 1.0e18f);
                    const float l2q = -4.0f * lamf * kLog2e;              // log2 q (negative rates: growth, as the reference)
                    const float a = fminf(16.0f * lamf * lamf, 3.0e38f);  // (a huge rate: q^t is 0 from t = 1 on, a 0 = 0)
                    const float ulo = a * ex2_approx(l2q * float(T0 + lane));
                    const float qB = ex2_approx(32.0f * l2q);
                    float w = 1.0f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        o[j] = fmaf(ulo, w, o[j]);
                        w *= qB;
                    }
                }
                __syncwarp();
            }
            while (qn - qh >= 32 || (!more && qn > qh)) {
                const int cnt = min(32, qn - qh);
                const double lam = (lane < cnt) ? s_q[(qh + lane) & 63] : 0.0;
                qh += cnt;
                __syncwarp();
                const float lamf = float(lam);                     // 0 for the lanes past cnt: every moment is 0
                const float c = 4.0f * lamf;
                float u0 = 16.0f * lamf * lamf;
                if (T0 != 0) {
                    // e^{-c T0}: the exponent is formed and reduced in double (c T0 reaches hundreds), 2^frac in single
                    const double arg = -(4.0 * lam) * t0d * 1.4426950408889634;
                    const double kd = rint(arg);
                    u0 = ldexpf(u0 * ex2_approx(float(arg - kd)), int(fmax(kd, -300.0)));
                }
                const float z = -64.0f * c;
                const float q64 = ex2_approx(z * kLog2e), q128 = q64 * q64;
                float m[KM];
                m[0] = u0;
#pragma unroll
                for (int k = 1; k < KM; ++k) m[k] = m[k - 1] * z;
                float w = q64;
#pragma unroll
                for (int hb = 0; hb < NB; ++hb) {
#pragma unroll
                    for (int k = 0; k < KM; ++k) acc[k][hb] = fmaf(m[k], w, acc[k][hb]);
                    w *= q128;
                }
            }
            if (!more) break;
        }

        // ---- close the 128 sums over the lanes: lane L of block b ends with M[4 b + (L >> 3)][L & 7] ----
#pragma unroll
        for (int b = 0; b < KM * NB / 32; ++b) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = acc[4 * b + (i >> 3)][i & 7];
            const float total = transpose_reduce32(v, lane);
            s_m[(lane & 7) * KM + 4 * b + (lane >> 3)] = total;
        }
        __syncwarp();
        // ---- epoch integrals of the slow sites from the moments (mom_epoch.h), in double ----
        for (int me = 0; me < args.n_mom_epochs; ++me) {
            const MomEpochDev ep = args.mom_epochs[me];
            if (ep.tile != item.tile) continue;
            double v = (lane < KM) ? __ldg(args.mom_p + 32 * ep.col + lane) * double(s_m[ep.hb * KM + lane]) : 0.0;
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
            if (lane == 0) args.mom_int[size_t(item.part) * args.e_total + ep.col] = v;
        }
        __syncwarp();
        // 1/k! onto the moments
#pragma unroll
        for (int j = 0; j < KM * NB / 32; ++j) {
            const int i = lane + 32 * j;
            s_m[i] *= float(c_inv_fact[i & (KM - 1)]);
        }
        __syncwarp();
        // ---- moments -> times: lane = grid point, 16-term Horner per point ----
#pragma unroll
        for (int hb = 0; hb < NB; ++hb) {
            float cf[KM];
#pragma unroll
            for (int q = 0; q < KM / 4; ++q) {
                const float4 t4 = *reinterpret_cast<const float4*>(s_m + hb * KM + 4 * q);
                cf[4 * q] = t4.x; cf[4 * q + 1] = t4.y; cf[4 * q + 2] = t4.z; cf[4 * q + 3] = t4.w;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float x = float(32 * r + lane - 64) * 0.015625f;
                float pv = cf[KM - 1];
#pragma unroll
                for (int k = KM - 2; k >= 0; --k) pv = fmaf(pv, x, cf[k]);
                o[4 * hb + r] += pv;
            }
        }
        __syncwarp();

        // ---- epilogue: float64 values times t, tile maximum (first maximum, as numpy.argmax), or the split workspace ----
        if (item.slot >= 0) {
            double* out = args.ws + size_t(item.slot) * TT;
#pragma unroll
            for (int j = 0; j < 32; ++j) out[32 * j + lane] = double(o[j]);
        } else {
            double* out = args.pi_net + int64_t(item.locus) * args.T;
            long long best_bits = -1;
            int64_t best_t = T0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int64_t t = T0 + 32 * j + lane;
                if (t < args.T) {
                    const double v = double(o[j]) * double(t);
                    out[t] = v;
                    const long long bits = __double_as_longlong(v);
                    if (bits > best_bits) { best_bits = bits; best_t = t; }
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const long long ob = __shfl_xor_sync(0xffffffffu, best_bits, off);
                const int64_t ot = __shfl_xor_sync(0xffffffffu, best_t, off);
                if (ob > best_bits || (ob == best_bits && ot < best_t)) { best_bits = ob; best_t = ot; }
            }
            if (lane == 0) {
                const int64_t k = int64_t(item.locus) * args.n_tiles + item.tile;
                args.tile_max[k] = __longlong_as_double(best_bits);
                args.tile_arg[k] = best_t;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// epoch integrals
// ------------------------------------------------------------------------------------------
// One (quad item, epoch) pair with at least one site whose first GK21 panel QAGS does not accept
// (or whose arguments fall outside the factored fast path): the sites are marked in 8 ballots.
struct FixEntry {
    int32_t item;
    int32_t slot;          // epoch slot within the batch
    uint32_t mask[kQuadBlock / 32];
};

struct QuadArgs {
    const double* rates;
    const QuadItem* items;
    double* qpart;       // [n_items][2*e_total + 4]: integral[e], abserr[e] ..., rate_sum, n_notnan, n_finite, n_multi
    FixEntry* fix;       // [n_items * kEpochsPerLaunch] worst case
    uint32_t* fix_count; // entries appended by this launch (zeroed before it)
    uint32_t* queue;     // work-queue head of this launch (zeroed before it)
    int32_t n_items;
    // moment-functional batches: quad_scan_kernel marks the sites quad_kernel may have to integrate, one ballot per
    // 32 sites (list[8 item + round]); the chunked quad_kernel<NE, kQuadChunk> fetches only those (list == nullptr:
    // quad_kernel<NE> takes every item, rate statistics included)
    uint32_t* list;
    uint32_t* list_count;  // unused
    int32_t chunk;         // items per work unit of the chunked form (1 .. kQuadChunk)
};

// GK21 on one epoch with the exponentials factored as e^{-c(centre -/+ h x_j)} = E0 * e^{+/- c h x_j}:
// D[j] = e^{-c h x_j} and R[j] = e^{+c h x_j} are shared by all epochs of equal half-length, and the
// common factor A0 = a E0 is applied once to the sums (every dqk21 quantity is homogeneous of degree one
// in the node values, and A0 >= 1e-270 on this path, so nothing under- or overflows differently).
// Node values are >= 0 here, so dqk21's resabs sum equals resk's.
// `scale` = a e^{-c centre} h (>= 0).
// Returns true when the whole warp took the short way out: for every live lane the rounding floor
// 50 eps resabs is PROVED to be dqk21's error estimate (and dqagse accepts the panel) from a lower bound of resasc
// -- two of its 21 non-negative terms -- so the other 19 are never formed; resasc is then not set.  The proof is
// the existing floor test, which is monotone in resasc, plus resasc > floor (dqagse compares abserr with resasc);
// whenever it holds the full computation yields the same abserr bit for bit.  `live`: the lane's result is used.
__device__ __forceinline__ bool gk21_shared_exp(const EpochBatch& eb, int e, double scale, bool live, const double (&D)[10],
                                                const double (&R)[10], double& result, double& abserr, double& resabs,
                                                double& resasc)
{
    const double fc = eb.center[e];
    double fv1[10], fv2[10];
    // Kronrod and Gauss sums over the node pairs.  dqk21 adds the 21 terms into one accumulator; here each sum
    // is split over independent partial accumulators (Gauss nodes / Kronrod-only nodes, two halves each) that are
    // combined at the end: same terms, a different association of the additions (differences of a few 1e-16
    // relative, below the noise of resk - resg that the reference itself carries), and a dependent chain of
    // 3 FMAs instead of 21.
    double rk0 = c_wgk[10] * fc, rk1 = 0.0, rk2 = 0.0, rk3 = 0.0;
    double rg0 = 0.0, rg1 = 0.0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int jt = 2 * j + 1;
        const double f1 = eb.tm[e][jt] * R[jt];
        const double f2 = eb.tp[e][jt] * D[jt];
        fv1[jt] = f1; fv2[jt] = f2;
        const double fsum = f1 + f2;
        if (j & 1) { rg1 += c_wg[j] * fsum; rk1 += c_wgk[jt] * fsum; }
        else       { rg0 += c_wg[j] * fsum; rk0 += c_wgk[jt] * fsum; }
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int jt = 2 * j;
        const double f1 = eb.tm[e][jt] * R[jt];
        const double f2 = eb.tp[e][jt] * D[jt];
        fv1[jt] = f1; fv2[jt] = f2;
        if (j & 1) rk3 += c_wgk[jt] * (f1 + f2);
        else       rk2 += c_wgk[jt] * (f1 + f2);
    }
    const double resk = (rk0 + rk1) + (rk2 + rk3);
    const double resg = rg0 + rg1;
    const double reskh = resk * 0.5;
    result = resk * scale;
    resabs = result;
    // dqk21's error heuristic: err = resasc * min(1, (200 e0 / resasc)^1.5), then raised to the rounding
    // floor 50 eps resabs.  Whether the floor wins is decided without the divide and the square root:
    // resasc (n/resasc)^1.5 <= floor  <=>  n^3 <= floor^2 resasc, compared in units of resk (scaled by a
    // power of two taken from resk's exponent so that neither side can overflow or underflow to a tie).
    const double e0 = fabs(resk - resg);
    const double unit = __hiloint2double((2046 - ((__double2hiint(resk) >> 20) & 0x7ff)) << 20, 0);  // ~ 1 / resk
    const double n = (200.0 * e0) * unit, fl = ((kEpmach * 50.0) * resk) * unit;
    const bool above_uflow = resabs > kUflow / (50.0 * kEpmach);
    {
        const double lb = (c_wgk[4] * (fabs(fv1[4] - reskh) + fabs(fv2[4] - reskh)) +
                           c_wgk[5] * (fabs(fv1[5] - reskh) + fabs(fv2[5] - reskh))) * unit;  // <= resasc * unit
        const bool proved = above_uflow && n < lb && (n * n) * n <= (fl * fl) * lb && lb > fl;
        if (__all_sync(0xffffffffu, proved || !live)) {
            abserr = (kEpmach * 50.0) * resabs;
            return true;
        }
    }
    double ra0 = c_wgk[10] * fabs(fc - reskh), ra1 = 0.0, ra2 = 0.0, ra3 = 0.0;
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        const double t = c_wgk[j] * (fabs(fv1[j] - reskh) + fabs(fv2[j] - reskh));
        if ((j & 3) == 0) ra0 += t;
        else if ((j & 3) == 1) ra1 += t;
        else if ((j & 3) == 2) ra2 += t;
        else ra3 += t;
    }
    const double rasc = (ra0 + ra1) + (ra2 + ra3);
    resasc = rasc * scale;
    const double ra = rasc * unit;
    if (above_uflow && (ra == 0.0 || e0 == 0.0)) {
        abserr = fmax((kEpmach * 50.0) * resabs, e0 * scale);
    } else if (above_uflow && n < ra && (n * n) * n <= (fl * fl) * ra) {
        abserr = (kEpmach * 50.0) * resabs;
    } else {
        abserr = gk_error(resk, resg, scale, resabs, resasc);
    }
    return false;
}

// The same panel without forming the node values.  With the common factor taken out, the pair sum of node j is
//   f1_j + f2_j = (centre - h x_j) R_j + (centre + h x_j) D_j = centre (R_j + D_j) - h x_j (R_j - D_j),
// linear in the epoch's centre, so for all epochs of one half-length
//   resk = centre Ak - Bk,   resk - resg = centre dA - dB,
//   Ak = wgk_10 + sum wgk_j (R_j + D_j),  Bk = h sum wgk_j x_j (R_j - D_j),  dA, dB: the same with the weights
//   c_dl (Kronrod minus Gauss), summed directly so that the cancellation happens once, in the site's tables.
// An epoch then costs two FMAs plus the proof that the rounding floor is dqk21's error estimate (lower bound of
// resasc from node pairs 4 and 5, which are formed as before).  Returns true when the proof holds for every live
// lane of the warp; otherwise nothing is written and the caller runs gk21_shared_exp.  The integral differs from
// the node-by-node sum by rounding only (a different association, ~1e-16 relative).
struct PanelSums {
    double Ak, Bk, dA, dB;
    double e0rel;           // series branch: bound of |resk - resg| / resk that stands in for dA, dB (else 0)
    double D4, R4, D5, R5;  // the two node pairs the proof looks at
};

// Slow sites (y x_0 log2(e) < 1/2 for the whole warp, the bulk of UCE-like rates): the pair sums are entire
// functions of y with known series (c_ak, c_bk: the rule's moments), seven terms each, so the ten exponential pairs
// are not formed at all -- only the two pairs the floor proof looks at.  resk - resg is then not computed: in exact
// arithmetic it is below 1e-27 resk here (Kronrod and Gauss both integrate the polynomial part exactly), and what
// dqk21 sees is the rounding noise of its two sums, at most ~25 eps resk; the proof runs with the bound 32 eps resk
// (it then needs resasc > 2.3e-8 resk, which only an integrand flat to eight digits across the epoch misses --
// the warp then takes the node-by-node path like any other unproved case).
__device__ __forceinline__ void panel_sums_series(double h, double y, PanelSums& ps)
{
    const double s = y * y;
    double pa = c_ak[6], pb = c_bk[6];
#pragma unroll
    for (int n = 5; n >= 0; --n) {
        pa = fma(pa, s, c_ak[n]);
        pb = fma(pb, s, c_bk[n]);
    }
    ps.Ak = pa;
    ps.Bk = (pb * y) * h;
    ps.dA = 0.0;
    ps.dB = 0.0;
    ps.e0rel = 32.0 * kEpmach;
    exp_pm_small(y * c_xgk[4], ps.D4, ps.R4);
    exp_pm_small(y * c_xgk[5], ps.D5, ps.R5);
}

__device__ __forceinline__ void panel_sums(double h, const double (&D)[10], const double (&R)[10], PanelSums& ps)
{
    double ak0 = c_wgk[10], ak1 = 0.0, bk0 = 0.0, bk1 = 0.0, da0 = c_wgk[10], da1 = 0.0, db0 = 0.0, db1 = 0.0;
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        const double sj = R[j] + D[j], dj = R[j] - D[j];
        if (j & 1) {
            ak1 = fma(c_wgk[j], sj, ak1); bk1 = fma(c_wx[j], dj, bk1);
            da1 = fma(c_dl[j], sj, da1);  db1 = fma(c_dlx[j], dj, db1);
        } else {
            ak0 = fma(c_wgk[j], sj, ak0); bk0 = fma(c_wx[j], dj, bk0);
            da0 = fma(c_dl[j], sj, da0);  db0 = fma(c_dlx[j], dj, db0);
        }
    }
    ps.Ak = ak0 + ak1;
    ps.Bk = (bk0 + bk1) * h;
    ps.dA = da0 + da1;
    ps.dB = (db0 + db1) * h;
    ps.e0rel = 0.0;
    ps.D4 = D[4]; ps.R4 = R[4]; ps.D5 = D[5]; ps.R5 = R[5];
}

__device__ __forceinline__ bool gk21_collapsed(const EpochBatch& eb, int e, double scale, bool live, const PanelSums& ps,
                                               double& result, double& abserr)
{
    const double fc = eb.center[e];
    const double resk = fma(fc, ps.Ak, -ps.Bk);
    const double e0 = fma(ps.e0rel, resk, fabs(fma(fc, ps.dA, -ps.dB)));
    const double reskh = resk * 0.5;
    const double unit = __hiloint2double((2046 - ((__double2hiint(resk) >> 20) & 0x7ff)) << 20, 0);  // ~ 1 / resk
    const double n = (200.0 * e0) * unit, fl = ((kEpmach * 50.0) * resk) * unit;
    const double f14 = eb.tm[e][4] * ps.R4, f24 = eb.tp[e][4] * ps.D4, f15 = eb.tm[e][5] * ps.R5, f25 = eb.tp[e][5] * ps.D5;
    const double lb = (c_wgk[4] * (fabs(f14 - reskh) + fabs(f24 - reskh)) + c_wgk[5] * (fabs(f15 - reskh) + fabs(f25 - reskh))) * unit;
    const double res = resk * scale;
    const bool proved = res > kUflow / (50.0 * kEpmach) && n < lb && (n * n) * n <= (fl * fl) * lb && lb > fl;
    if (!__all_sync(0xffffffffu, proved || !live)) return false;
    result = res;
    abserr = (kEpmach * 50.0) * res;
    return true;
}

// First GK21 panel of every (site, epoch) of the batch.  A site whose panel is what QAGS returns (the
// overwhelmingly common case) is summed here; the others contribute 0 and are handed to
// quad_fix_kernel through the FixEntry list.
//
// Warps are autonomous (no block barrier): a warp takes one item (<= 256 consecutive sites of one locus) from
// a work queue, streams its rates 32 at a time, and packs the sites that need arithmetic -- finite, 16 r^2 != 0;
// a zero-rate site integrates to exactly (0, 0) and QAGS accepts that as is -- through a 64-entry queue in
// shared memory, so every pass over the GK21 code runs with 32 useful lanes.  Each lane adds its sites' values
// into its own shared-memory accumulators (site order), and one shuffle tree per column closes the item: the
// summation order depends on the item's content only, so results are bit-reproducible.
//
// NE = eb.n (0..kEpochsPerLaunch) is a template parameter so that the per-site computation is one straight-line
// block the scheduler can interleave (the NE exponentials of the common factors in particular).
constexpr int kQuadWarps = 1;  // warps per CTA
constexpr int kQuadChunk = 32; // items per work unit of the chunked form (moment-functional batches)

// CH = 1: a work unit is one item; every site of the item is integrated here (and the item's rate statistics are
// formed).  CH = kQuadChunk (moment-functional batches): a work unit is CH consecutive items of which only those that
// quad_scan_kernel flagged are streamed at all, and of those only the sites the moments do not serve; the few such
// sites of the whole unit share the 32-lane passes (a pass per item would run with one or two live lanes), a
// segmented scan over the lanes hands every item its own sums.  Either way the order of the additions depends on the
// unit's content only.
template <int NE, int CH = 1>
__global__ void __launch_bounds__(kQuadWarps * 32, 12 / kQuadWarps) quad_kernel(const QuadArgs args, const EpochBatch eb)
{
    static_assert(kQuadWarps == 1 && CH >= 1 && CH <= 32, "one-warp CTAs; a unit's items map to the lanes");
    constexpr int NA = NE > 0 ? 2 * NE : 1;
    constexpr int QN = (CH == 1) ? 64 : 512;                     // queue entries: a pass' worth left over + one round / one item
    __shared__ double s_q[QN];                                   // queue of active rates
    __shared__ uint16_t s_src[QN];                               // position inside the item (low byte), item of the unit (high byte)
    __shared__ double s_acc[CH == 1 ? NA : 1][32];               // CH = 1: per-lane sums, integral / abserr per epoch
    __shared__ double s_tot[CH == 1 ? 1 : CH][NA];               // CH > 1: per-item sums
    __shared__ uint32_t s_flag[CH][kEpochsPerLaunch][kQuadBlock / 32];
    const int lane = threadIdx.x & 31;
    double* q = s_q;
    uint16_t* qsrc = s_src;
    const int row = 2 * eb.e_total + 4;
    // bit e: the site takes the per-site path for epoch e (all epochs unless the moment functional serves it)
    auto need_mask = [&](double lam) -> uint32_t {
        const uint32_t all = (NE > 0) ? ((1u << NE) - 1u) : 1u;
        if (!eb.mom_any) return all;
        if (!(rate_is_active(lam) && rate_is_slow(lam, eb.class_hi))) return all;
        const unsigned long long rb = (unsigned long long)__double_as_longlong(lam);
        uint32_t need = 0;
#pragma unroll 1
        for (int e = 0; e < eb.n; ++e) {
            const bool served = eb.mom[e] && rb >= eb.rmin_bits[e] && rb <= eb.rmax_bits[e] &&
                                !(rb >= eb.wlo_bits[e] && rb <= eb.whi_bits[e]);
            need |= served ? 0u : (1u << e);
        }
        return need;
    };
    const int ch = (CH == 1) ? 1 : args.chunk;   // items per work unit (<= CH)
    const int n_units = (args.n_items + ch - 1) / ch;

    for (;;) {
        int unit = 0;
        if (lane == 0) unit = int(atomicAdd(args.queue, 1u));
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= n_units) break;
        const int it0 = unit * ch, n_local = min(ch, args.n_items - it0);
        // which items of the unit hold a site for this kernel (CH = 1: the item itself, always): lane = item, its
        // eight ballots of candidate sites stay in the lane's registers
        uint32_t todo = 1u;
        uint32_t cand[kQuadBlock / 32];
        long long my_begin = 0;   // CH > 1: first site of this lane's item
#pragma unroll
        for (int r = 0; r < kQuadBlock / 32; ++r) cand[r] = 0u;
        if constexpr (CH > 1) {
            if (lane < n_local) {
                my_begin = args.items[it0 + lane].site_begin;
                const uint4* lp = reinterpret_cast<const uint4*>(args.list + size_t(it0 + lane) * (kQuadBlock / 32));
                const uint4 w0 = __ldg(lp), w1 = __ldg(lp + 1);
                cand[0] = w0.x; cand[1] = w0.y; cand[2] = w0.z; cand[3] = w0.w;
                cand[4] = w1.x; cand[5] = w1.y; cand[6] = w1.z; cand[7] = w1.w;
            }
            uint32_t any = 0u;
#pragma unroll
            for (int r = 0; r < kQuadBlock / 32; ++r) any |= cand[r];
            todo = __ballot_sync(0xffffffffu, any != 0u);
        }
        if (todo == 0u) continue;
        const uint32_t todo_all = todo;
        if constexpr (CH == 1) {
#pragma unroll
            for (int k = 0; k < NA; ++k) s_acc[k][lane] = 0.0;
        } else {
            for (int i = lane; i < CH * NA; i += 32) (&s_tot[0][0])[i] = 0.0;
        }
        for (int i = lane; i < CH * kEpochsPerLaunch * (kQuadBlock / 32); i += 32) (&s_flag[0][0][0])[i] = 0u;
        double rs = 0.0, nn = 0.0, nf = 0.0;  // rate sum, #not-NaN, #finite of this lane's sites (CH = 1)
        int qh = 0, qn = 0;
        __syncwarp();

        // one pass of the GK21 code over `cnt` queued sites
        auto run_pass = [&](int cnt) {
            const bool on = lane < cnt;
            const double lam2 = on ? q[(qh + lane) & (QN - 1)] : 0.0;
            const uint32_t src = on ? qsrc[(qh + lane) & (QN - 1)] : 0xff00u;
            qh += cnt;
            const uint32_t pos = src & 0xffu, li = src >> 8;
            const double a = 16.0 * (lam2 * lam2);
            const double c = 4.0 * lam2;
            // fast path: all exponent arguments stay well inside the normal range, every node value >= 0
            const bool fast = on && eb.nonneg && (c >= 0.0) && (c * eb.bmax <= 600.0) && (a <= DBL_MAX);
            const double cs = fast ? c : 0.0, as = fast ? a : 0.0;  // keeps idle lanes' arithmetic finite
            const uint32_t need = on ? need_mask(lam2) : 0u;
            const bool in_moments = on && eb.mom_any && rate_is_active(lam2) && rate_is_slow(lam2, eb.class_hi);
            // CH > 1: the lanes of one item are consecutive; the last lane of a run adds the run's total to the item
            uint32_t li_next = 0xffffu;
            if constexpr (CH > 1) li_next = __shfl_down_sync(0xffffffffu, li, 1);
            const bool seg_last = on && (lane == cnt - 1 || li_next != li);
            PanelSums ps;
            // one copy of the GK21 code, looped over the epochs: the fully unrolled form (4 x ~350
            // instructions) thrashed the instruction cache once the warps ran out of step (18 % of the
            // stall samples were instruction fetches)
            // (for NE <= 2 the trip count is read from the batch, eb.n == NE: with a compile-time count of 1 or 2
            // the compiler flattens the loop and the longer live ranges spill 140-190 bytes -- E = 2: 0.105 ->
            // 0.087 ms on config 2)
            const int ne = NE <= 2 ? eb.n : NE;
#pragma unroll 1
            for (int e = 0; e < ne; ++e) {
                if (eb.new_half[e]) {
                    double D[10], R[10];
                    const double y = cs * eb.half[e];
                    // slow sites (the bulk of UCE-like data), every y x_j below ln(2)/2 (x_0 log2 e = 1.4364...):
                    // the sums from their power series, no node tables
                    if (__all_sync(0xffffffffu, y * 1.4365 < 0.4999)) {
                        panel_sums_series(eb.half[e], y, ps);
                    } else {
#pragma unroll
                        for (int j = 0; j < 10; ++j) exp_pm_bounded(y * c_xgk[j], D[j], R[j]);  // 0 <= y x_j <= 600
                        panel_sums(eb.half[e], D, R, ps);
                    }
                }
                // common factor a e^{-c centre} h; |c * centre| <= 600 on this path
                const double scale = (as * exp_bounded(-(cs * eb.center[e]))) * eb.half[e];
                double result, abserr, resabs, resasc = 0.0;
                const bool wanted = (need >> e) & 1u;   // this (site, epoch) is not served by the moments
                const bool live = fast && wanted;
                bool by_floor = gk21_collapsed(eb, e, scale, live, ps, result, abserr);
                if (!by_floor) {
                    // rare: some site's estimate is not the floor.  The node tables are rebuilt here (same
                    // function, same argument, same values) instead of being kept alive across the epochs.
                    double D[10], R[10];
                    const double y = cs * eb.half[e];
#pragma unroll
                    for (int j = 0; j < 10; ++j) exp_pm_bounded(y * c_xgk[j], D[j], R[j]);
                    by_floor = gk21_shared_exp(eb, e, scale, live, D, R, result, abserr, resabs, resasc);
                }
                // by_floor: abserr = 50 eps resabs <= 100 eps resabs and abserr != resasc, so dqagse's test
                // (round-off exit above errbnd, plain acceptance below) passes either way
                const bool keep = live && (by_floor || qags_first_panel_final(result, abserr, resabs, resasc));
                double vi = keep ? result : 0.0, ve = keep ? abserr : 0.0;
                if (eb.mom_any && eb.mom[e]) {
                    // a site of the moment class that is integrated here (rate inside the epoch's window or
                    // outside its range) is also part of the locus' moments: take out what the functional
                    // adds for it -- a e^{-c tb} P(z), z = -64 c, the same polynomial on the same site
                    const bool sub = in_moments && wanted;
                    if (__any_sync(0xffffffffu, sub)) {
                        const double* pc = eb.mom_p + 32 * eb.col[e];
                        const double z = -64.0 * c;
                        double hz = __ldg(pc + eb.mom_terms - 1);
#pragma unroll 1
                        for (int k = eb.mom_terms - 2; k >= 0; --k) hz = fma(hz, z, __ldg(pc + k));
                        const double pv = (a * exp(-(c * eb.tb[e]))) * hz;
                        if (sub) { vi -= pv; ve -= (kEpmach * 50.0) * pv; }
                    }
                }
                if constexpr (CH == 1) {
                    s_acc[2 * e][lane] += vi;
                    s_acc[2 * e + 1][lane] += ve;
                } else {
                    // segmented inclusive scan over the lanes (runs of equal li), then one add per run and column
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const double ui = __shfl_up_sync(0xffffffffu, vi, d), ue = __shfl_up_sync(0xffffffffu, ve, d);
                        const uint32_t lu = __shfl_up_sync(0xffffffffu, li, d);
                        if (lane >= d && lu == li) { vi += ui; ve += ue; }
                    }
                    if (seg_last) {
                        s_tot[li][2 * e] += vi;
                        s_tot[li][2 * e + 1] += ve;
                    }
                }
                if (on && wanted && !keep)  // rare: hand the site (by its position in the item) to quad_fix_kernel
                    atomicOr(&s_flag[CH == 1 ? 0 : li][e][pos >> 5], 1u << (pos & 31));
            }
            __syncwarp();
        };

        // Fill the queue until a full pass is there (or nothing is left), run ONE pass, repeat: a single call site of
        // the GK21 code per instantiation (inlined at every place it used to be called from, it no longer fitted the
        // instruction cache: 60 % of the stall samples were instruction fetches).
        QuadItem item;
        item.site_begin = 0; item.count = 0; item.locus = 0;
        const double* rp = args.rates;
        int base = 0;
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        double lam_in = qnan, lam_in2 = qnan;
        if constexpr (CH == 1) {
            item = args.items[it0];
            rp = args.rates + item.site_begin;
            lam_in = (lane < item.count) ? __ldg(rp + lane) : qnan;             // fetched two rounds ahead, as in
            lam_in2 = (32 + lane < item.count) ? __ldg(rp + 32 + lane) : qnan;  // profile_kernel
        }
        for (;;) {
            if constexpr (CH == 1) {
                while (qn - qh < 32 && base < item.count) {
                    const double lam = lam_in;
                    lam_in = lam_in2;
                    {
                        const int nb = base + 64 + lane;
                        lam_in2 = (nb < item.count) ? __ldg(rp + nb) : qnan;
                    }
                    const bool notnan = lam == lam;  // lanes past the item hold NaN
                    const bool fin = fabs(lam) <= DBL_MAX;
                    rs += notnan ? lam : 0.0;
                    nn += notnan ? 1.0 : 0.0;
                    nf += fin ? 1.0 : 0.0;
                    const bool act = fin && (16.0 * (lam * lam) != 0.0) && need_mask(lam) != 0u;
                    const unsigned ballot = __ballot_sync(0xffffffffu, act);
                    if (act) {
                        const int pos = (qn + __popc(ballot & ((1u << lane) - 1u))) & (QN - 1);
                        q[pos] = lam;
                        qsrc[pos] = uint16_t(base + lane);
                    }
                    qn += __popc(ballot);
                    base += 32;
                    __syncwarp();
                }
            } else {
                while (qn - qh < 32 && todo) {
                    // the next flagged item: only its candidate sites are fetched -- all of them at once, so that one
                    // memory latency is exposed per item; the exact test is run_pass's (a site the moments serve
                    // after all adds nothing there)
                    const int li_now = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const double* ip = args.rates + __shfl_sync(0xffffffffu, my_begin, li_now);
                    uint32_t w[kQuadBlock / 32];
                    double val[kQuadBlock / 32];
#pragma unroll
                    for (int r = 0; r < kQuadBlock / 32; ++r) {
                        w[r] = __shfl_sync(0xffffffffu, cand[r], li_now);
                        val[r] = ((w[r] >> lane) & 1u) ? __ldg(ip + 32 * r + lane) : 0.0;
                    }
#pragma unroll
                    for (int r = 0; r < kQuadBlock / 32; ++r) {
                        if ((w[r] >> lane) & 1u) {
                            const int pos = (qn + __popc(w[r] & ((1u << lane) - 1u))) & (QN - 1);
                            q[pos] = val[r];
                            qsrc[pos] = uint16_t((32 * r + lane) | (li_now << 8));
                        }
                        qn += __popc(w[r]);
                    }
                    __syncwarp();
                }
            }
            if (qn == qh) break;
            if constexpr (NE > 0) run_pass(min(32, qn - qh));
            else qh = qn;
        }

        // ---- close the unit ----
        if constexpr (CH == 1) {
            // one shuffle tree per column, fixed order
            const int it = it0;
            double* out = args.qpart + size_t(it) * row;
#pragma unroll
            for (int k = 0; k < 2 * NE; ++k) {
                double v = s_acc[k][lane];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
                if (lane == 0) out[2 * eb.col[k >> 1] + (k & 1)] = v;
            }
            if (eb.first) {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    rs += __shfl_down_sync(0xffffffffu, rs, off);
                    nn += __shfl_down_sync(0xffffffffu, nn, off);
                    nf += __shfl_down_sync(0xffffffffu, nf, off);
                }
                if (lane == 0) {
                    out[2 * eb.e_total + 0] = rs;
                    out[2 * eb.e_total + 1] = nn;
                    out[2 * eb.e_total + 2] = nf;
                    out[2 * eb.e_total + 3] = 0.0;  // multi-panel count, accumulated by quad_fix_kernel
                }
            }
        } else {
            // the flagged items' cells (quad_scan_kernel zeroed every item's, and wrote the statistics)
            __syncwarp();
            for (int i = lane; i < n_local * 2 * NE; i += 32) {
                const int li = i / (2 * NE), k = i % (2 * NE);
                if ((todo_all >> li) & 1u)
                    args.qpart[size_t(it0 + li) * row + 2 * eb.col[k >> 1] + (k & 1)] = s_tot[li][k];
            }
        }
        __syncwarp();
        for (int i = lane; i < n_local * NE; i += 32) {
            const int li = i / max(NE, 1), e = i % max(NE, 1);
            uint32_t any = 0;
#pragma unroll
            for (int w = 0; w < kQuadBlock / 32; ++w) any |= s_flag[li][e][w];
            if (any) {
                FixEntry* fe = args.fix + atomicAdd(args.fix_count, 1u);
                fe->item = it0 + li;
                fe->slot = e;
#pragma unroll
                for (int w = 0; w < kQuadBlock / 32; ++w) fe->mask[w] = s_flag[li][e][w];
            }
        }
        __syncwarp();
    }
}

// Moment-functional batches: almost every site is served by the locus' moments (mom_epoch.h), so what is left of
// the per-site path is a streaming pass over the rates that (i) writes the rate statistics of every item (the same
// per-lane sums and shuffle tree as quad_kernel, bit for bit), (ii) zeroes the item's epoch cells and (iii) flags the
// items holding at least one site quad_kernel must integrate.  The test is the conservative form of quad_kernel's
// need_mask on the high words of the rates (a listed item is then filtered exactly by quad_kernel itself), integer
// pipe only; the eight rounds of an item are loaded up front so that a warp keeps 2 KB in flight.
constexpr int kScanThreads = 256;

__global__ void __launch_bounds__(kScanThreads) quad_scan_kernel(const QuadArgs args, const EpochBatch eb)
{
    const int lane = threadIdx.x & 31;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int row = 2 * eb.e_total + 4;
    for (int it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < args.n_items; it += n_warps) {
        const QuadItem item = args.items[it];
        const double* rp = args.rates + item.site_begin;
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);
        double v[kQuadBlock / 32];
#pragma unroll
        for (int k = 0; k < kQuadBlock / 32; ++k) v[k] = (32 * k + lane < item.count) ? __ldg(rp + 32 * k + lane) : qnan;
        double rs = 0.0;
        int nn = 0, nf = 0;
        uint32_t mine = 0u;   // lane k keeps the ballot of round k
#pragma unroll
        for (int k = 0; k < kQuadBlock / 32; ++k) {
            const uint32_t hs = uint32_t(__double2hiint(v[k])), hi = hs & 0x7fffffffu, lo = uint32_t(__double2loint(v[k]));
            const bool fin = hi < 0x7ff00000u;
            const bool notnan = fin || (hi == 0x7ff00000u && lo == 0u);
            rs += notnan ? v[k] : 0.0;
            nn += notnan ? 1 : 0;
            nf += fin ? 1 : 0;
            // quad_kernel's `act`: finite and 16 r^2 != 0 (|r| above 0x1e56a09e667f3bcc, see rate_is_active)
            const bool act = fin && (hi > 0x1e56a09eu || (hi == 0x1e56a09eu && lo > 0x667f3bccu));
            // inside the launch's common range (positive, in the moment class): one unsigned comparison
            bool served = eb.mom_all && (hs - eb.scan_lo - 1u) < (eb.scan_hi - eb.scan_lo - 1u);
#pragma unroll 1
            for (int w = 0; w < eb.n_win; ++w) served = served && (hs - eb.win_lo[w]) > (eb.win_hi[w] - eb.win_lo[w]);
            const uint32_t b = __ballot_sync(0xffffffffu, act && !served);
            if (lane == k) mine = b;
        }
        nn = __reduce_add_sync(0xffffffffu, nn);
        nf = __reduce_add_sync(0xffffffffu, nf);
        double* out = args.qpart + size_t(it) * row;
        if (lane < 2 * eb.n) out[2 * eb.col[lane >> 1] + (lane & 1)] = 0.0;
        if (eb.first) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) rs += __shfl_down_sync(0xffffffffu, rs, off);
            if (lane == 0) {
                out[2 * eb.e_total + 0] = rs;
                out[2 * eb.e_total + 1] = double(nn);
                out[2 * eb.e_total + 2] = double(nf);
                out[2 * eb.e_total + 3] = 0.0;
            }
        }
        if (lane < kQuadBlock / 32) args.list[size_t(it) * (kQuadBlock / 32) + lane] = mine;   // 32 bytes per item
    }
}

// Full QAGS (bisection + epsilon extrapolation, gk21.cuh) for the flagged sites.  One warp per
// FixEntry, so every (item, epoch) cell of qpart has a single writer and a fixed summation order.
__global__ void __launch_bounds__(128) quad_fix_kernel(const QuadArgs args, const EpochBatch eb)
{
    const int lane = threadIdx.x & 31;
    const uint32_t n = *args.fix_count;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const int row = 2 * eb.e_total + 4;
    for (uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += warps) {
        const FixEntry fe = args.fix[k];
        const QuadItem item = args.items[fe.item];
        double si = 0.0, se = 0.0, nm = 0.0;
        for (int w = 0; w < kQuadBlock / 32; ++w) {
            const uint32_t m = fe.mask[w];
            if (m == 0) continue;
            double result = 0.0, abserr = 0.0, multi = 0.0;
            if ((m >> lane) & 1u) {
                const double lam = __ldg(args.rates + item.site_begin + w * 32 + lane);
                int last;
                qags_device(eb.a[fe.slot], eb.b[fe.slot], 16.0 * (lam * lam), 4.0 * lam, &result, &abserr, &last);
                multi = last > 1 ? 1.0 : 0.0;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                result += __shfl_down_sync(0xffffffffu, result, off);
                abserr += __shfl_down_sync(0xffffffffu, abserr, off);
                multi += __shfl_down_sync(0xffffffffu, multi, off);
            }
            si += result;
            se += abserr;
            nm += multi;
        }
        if (lane == 0) {
            double* out = args.qpart + size_t(fe.item) * row;
            out[2 * eb.col[fe.slot]] += si;
            out[2 * eb.col[fe.slot] + 1] += se;
            if (nm != 0.0) atomicAdd(out + 2 * eb.e_total + 3, nm);  // small integers: exact in any order
        }
    }
}

// ------------------------------------------------------------------------------------------
// per-locus finalize: one warp per locus
// ------------------------------------------------------------------------------------------
struct FinalArgs {
    int64_t n_loci;
    int64_t T;
    int32_t TT;        // times per tile
    int32_t n_tiles;
    int32_t K;         // --times
    int32_t E;         // epochs
    const int32_t* nsplit;       // [L]
    const int64_t* ws_base;      // [L] first workspace row of a split locus
    const int64_t* qitem_begin;  // [L+1] quad item range per locus
    const int64_t* time_idx;     // [K] normalised to 0..T-1
    const double* ws;
    const double* tile_max;
    const int64_t* tile_arg;
    const double* qpart;
    double* pi_net;
    double* pi_times;
    double* integral;
    double* abserr;
    double* max_pi;
    int64_t* argmax_t;
    int32_t* argmax_epoch;
    double* mean_rate;
    int64_t* n_finite;
    int64_t* n_multi;
    // moment-functional epoch sums written by profile_mom_kernel: rows part_base[l] .. part_base[l+1]-1 of mom_int
    const int64_t* part_base;    // [L+1]
    const double* mom_int;       // [parts][E]
    unsigned long long mom_mask; // bit e: epoch e is served by the moments
};

__global__ void __launch_bounds__(128) finalize_kernel(const FinalArgs fa)
{
    const int lane = threadIdx.x & 31;
    const int64_t l = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (l >= fa.n_loci) return;
    // One warp per locus and a handful of dependent loads: the kernel is a chain of DRAM round trips, so everything that
    // does not depend on another load is fetched first (the locus' descriptors), then everything that depends on those
    // only (tile maxima, the --times picks, the epoch cells), and the arithmetic comes last.
    const int ns = __ldg(fa.nsplit + l);
    const int64_t q0 = __ldg(fa.qitem_begin + l), q1 = __ldg(fa.qitem_begin + l + 1);
    const bool mom = fa.mom_mask != 0ull;
    const int64_t p0 = mom ? __ldg(fa.part_base + l) : 0, p1 = mom ? __ldg(fa.part_base + l + 1) : 0;
    const int64_t tix = (lane < fa.K) ? __ldg(fa.time_idx + lane) : 0;
    const int row = 2 * fa.E + 4;

    double best = -1.0;
    int64_t best_t = 0;
    double pick = 0.0;
    if (ns > 1) {
        double* out = fa.pi_net + l * fa.T;
        for (int tile = 0; tile < fa.n_tiles; ++tile) {
            const double* w0 = fa.ws + size_t(fa.ws_base[l] + int64_t(tile) * ns) * fa.TT;
            for (int i = lane; i < fa.TT; i += 32) {
                const int64_t t = int64_t(tile) * fa.TT + i;
                if (t >= fa.T) break;
                double s = 0.0;
                for (int j = 0; j < ns; ++j) s += w0[size_t(j) * fa.TT + i];
                const double v = s * double(t);
                out[t] = v;
                if (v > best) { best = v; best_t = t; }
            }
        }
        __syncwarp();
        if (lane < fa.K) pick = fa.pi_net[l * fa.T + tix];   // after the row has been written
    } else {
        if (lane < fa.K) pick = fa.pi_net[l * fa.T + tix];
        for (int tile = lane; tile < fa.n_tiles; tile += 32) {
            const double v = fa.tile_max[l * fa.n_tiles + tile];
            const int64_t t = fa.tile_arg[l * fa.n_tiles + tile];
            if (v > best || (v == best && t < best_t)) { best = v; best_t = t; }
        }
    }
    // epoch sums over this locus' quad items, in item (= site) order; lanes 0..3 of the last round also take the four
    // statistics columns (loads issued before the reductions below need anything)
    double si0 = 0.0, se0 = 0.0, sm0 = 0.0, st = 0.0;
    const bool first_e = lane < fa.E;
    if (first_e) {
#pragma unroll 4
        for (int64_t q = q0; q < q1; ++q) {
            si0 += fa.qpart[q * row + 2 * lane];
            se0 += fa.qpart[q * row + 2 * lane + 1];
        }
        if ((fa.mom_mask >> lane) & 1ull)
            for (int64_t p = p0; p < p1; ++p) sm0 += fa.mom_int[p * fa.E + lane];
    }
    if (lane < 4) {
#pragma unroll 4
        for (int64_t q = q0; q < q1; ++q) st += fa.qpart[q * row + 2 * fa.E + lane];
    }

#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int64_t ot = __shfl_xor_sync(0xffffffffu, best_t, off);
        if (ob > best || (ob == best && ot < best_t)) { best = ob; best_t = ot; }
    }
    if (lane == 0) {
        fa.max_pi[l] = best;
        fa.argmax_t[l] = best_t;
    }
    if (lane < fa.K) fa.pi_times[l * fa.K + lane] = pick;
    for (int k = lane + 32; k < fa.K; k += 32) fa.pi_times[l * fa.K + k] = fa.pi_net[l * fa.T + fa.time_idx[k]];

    double ebest = 0.0;
    int ebest_idx = 0x7fffffff;
    bool have = false;
    for (int e = lane; e < fa.E; e += 32) {
        double si = si0, se = se0, sm = sm0;
        if (e != lane) {   // more than 32 epochs: the later rounds as before
            si = se = sm = 0.0;
            for (int64_t q = q0; q < q1; ++q) {
                si += fa.qpart[q * row + 2 * e];
                se += fa.qpart[q * row + 2 * e + 1];
            }
            if ((fa.mom_mask >> e) & 1ull)
                for (int64_t p = p0; p < p1; ++p) sm += fa.mom_int[p * fa.E + e];
        }
        if ((fa.mom_mask >> e) & 1ull) {
            // the locus' slow sites: sum of their first GK21 panels from the moments; every one of them has the
            // rounding floor 50 eps resabs (= 50 eps result, the integrand is >= 0) as QUADPACK's error estimate
            si += sm;
            se += (kEpmach * 50.0) * sm;
        }
        fa.integral[l * fa.E + e] = si;
        fa.abserr[l * fa.E + e] = se;
        if (!have || si > ebest) { ebest = si; ebest_idx = e; have = true; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, ebest, off);
        const int oi = __shfl_xor_sync(0xffffffffu, ebest_idx, off);
        if (oi != 0x7fffffff && (ebest_idx == 0x7fffffff || ob > ebest || (ob == ebest && oi < ebest_idx))) {
            ebest = ob;
            ebest_idx = oi;
        }
    }
    if (lane < 4) {
        const double rs = __shfl_sync(0x0000000fu, st, 0);
        if (lane == 1) fa.mean_rate[l] = rs / st;  // 0/0 -> NaN for a locus without any non-NaN rate
        if (lane == 2) fa.n_finite[l] = int64_t(st);
        if (lane == 3) fa.n_multi[l] = int64_t(st);
        if (lane == 0) fa.argmax_epoch[l] = (fa.E > 0) ? ebest_idx : -1;
    }
}

// ------------------------------------------------------------------------------------------
// ranking queries on the resident results (tapir_compare.py:93-101 "ORDER BY pi DESC limit N" per interval;
// README.rst:178-185 count of loci per max-PI epoch).  One CTA per epoch: 8-bit radix select of the N-th
// largest integral, index-ordered compaction of the winners, bitonic sort of the <= 1024 candidates.
// Integer atomics only, so the result is deterministic; ties go to the lower locus index.
// ------------------------------------------------------------------------------------------
constexpr int kRankThreads = 1024;
constexpr int kRankMaxTop = 1024;

__device__ __forceinline__ unsigned long long rank_key(double v)
{
    // order-preserving map of a double to an unsigned integer (larger value -> larger key)
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(kRankThreads) rank_kernel(const double* __restrict__ integral, int64_t L, int32_t E,
                                                            int32_t n_top, int64_t* __restrict__ top_locus,
                                                            double* __restrict__ top_pi)
{
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_need, s_count, s_warp[kRankThreads / 32];
    __shared__ unsigned long long ckey[kRankMaxTop];
    __shared__ int64_t cidx[kRankMaxTop];
    const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t want = L < n_top ? L : n_top;
    // ---- radix select: key of the want-th largest element ----
    if (tid == 0) { s_prefix = 0; s_need = (unsigned int)want; }
    __syncthreads();
    for (int shift = 56; shift >= 0 && want > 0; shift -= 8) {
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const unsigned long long mask = shift == 56 ? 0ull : (~0ull << (shift + 8));
        for (int64_t l = tid; l < L; l += kRankThreads) {
            const unsigned long long k = rank_key(integral[l * E + e]);
            if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 0xff], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int need = s_need;
            int b = 255;
            for (; b > 0; --b) {
                if (hist[b] >= need) break;
                need -= hist[b];
            }
            s_prefix = prefix | ((unsigned long long)b << shift);
            s_need = need;  // how many elements equal to the final key are still wanted
        }
        __syncthreads();
    }
    const unsigned long long tau = s_prefix;
    const unsigned int need_equal = s_need;
    // ---- index-ordered compaction: everything above tau, then the first need_equal elements equal to tau ----
    if (tid == 0) s_count = 0;
    __syncthreads();
    for (int pass = 0; pass < 2 && want > 0; ++pass) {
        const unsigned int limit = pass == 0 ? (unsigned int)want : (unsigned int)want;
        unsigned int taken_equal = 0;
        for (int64_t base = 0; base < L; base += kRankThreads) {
            const int64_t l = base + tid;
            unsigned long long k = 0;
            bool hit = false;
            if (l < L) {
                k = rank_key(integral[l * E + e]);
                hit = pass == 0 ? (k > tau) : (k == tau);
            }
            const unsigned int ballot = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_warp[warp] = __popc(ballot);
            __syncthreads();
            unsigned int before = 0, total = 0;
            for (int w = 0; w < kRankThreads / 32; ++w) {
                if (w < warp) before += s_warp[w];
                total += s_warp[w];
            }
            const unsigned int start = s_count;
            const unsigned int pos = start + before + __popc(ballot & ((1u << lane) - 1u));
            const unsigned int cap = pass == 0 ? limit : start + (need_equal - taken_equal);
            if (hit && pos < cap && pos < (unsigned int)kRankMaxTop) { ckey[pos] = k; cidx[pos] = l; }
            __syncthreads();
            if (tid == 0) s_count = min(start + total, cap);
            if (pass == 1) taken_equal = min(taken_equal + total, need_equal);
            __syncthreads();
            if (pass == 1 && taken_equal >= need_equal) break;
        }
    }
    const unsigned int n = s_count;
    // ---- bitonic sort of the candidates: key descending, index ascending ----
    for (int i = tid; i < kRankMaxTop; i += kRankThreads)
        if (i >= (int)n) { ckey[i] = 0; cidx[i] = 0x7fffffffffffffffLL; }
    __syncthreads();
    for (int size = 2; size <= kRankMaxTop; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const int i = tid, j = i ^ stride;
            if (j > i) {
                const bool desc = (i & size) == 0;
                const bool i_first = ckey[i] > ckey[j] || (ckey[i] == ckey[j] && cidx[i] < cidx[j]);
                if (i_first != desc) {
                    const unsigned long long tk = ckey[i]; ckey[i] = ckey[j]; ckey[j] = tk;
                    const int64_t ti = cidx[i]; cidx[i] = cidx[j]; cidx[j] = ti;
                }
            }
            __syncthreads();
        }
    for (int i = tid; i < n_top; i += kRankThreads) {
        const bool ok = i < (int)n;
        top_locus[int64_t(e) * n_top + i] = ok ? cidx[i] : -1;
        top_pi[int64_t(e) * n_top + i] = ok ? integral[cidx[i] * E + e] : __longlong_as_double(0x7ff8000000000000LL);
    }
}

__global__ void epoch_count_kernel(const int32_t* __restrict__ argmax_epoch, int64_t L, int32_t E,
                                   unsigned long long* __restrict__ count)
{
    for (int64_t l = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; l < L; l += int64_t(gridDim.x) * blockDim.x) {
        const int32_t e = argmax_epoch[l];
        if (e >= 0 && e < E) atomicAdd(count + e, 1ull);
    }
}

// ------------------------------------------------------------------------------------------
// multi-GPU gather: copy the 9 sections of a batch's small-result block into this rank's slot of the owner's
// buffer.  `dst` is peer memory (CUDA IPC mapping of a buffer on another GPU): the stores travel over NVLink.
// A chunk of a larger batch lands at its row offset inside the full batch's layout.
// ------------------------------------------------------------------------------------------
struct PushArgs {
    const uint32_t* src;
    uint32_t* dst;
    uint32_t src_off[9], dst_off[9], words[9];  // all in 4-byte words
};

__global__ void __launch_bounds__(256) push_small_kernel(const PushArgs pa)
{
    const uint32_t stride = gridDim.x * blockDim.x;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const uint32_t* s = pa.src + pa.src_off[i];
        uint32_t* d = pa.dst + pa.dst_off[i];
        for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < pa.words[i]; w += stride) d[w] = s[w];
    }
}

// ------------------------------------------------------------------------------------------
// API-parity element-wise kernels (tapir.compute.get_townsend_pi / get_integral_over_times)
// ------------------------------------------------------------------------------------------
// out[t*S + s] = ((16 r^2) t) exp(-((4 r) t)) in the reference's evaluation order (compute.py:48).
__global__ void pi_matrix_kernel(const double* __restrict__ time, int64_t n_time, const double* __restrict__ rates,
                                 int64_t n_rates, double* __restrict__ out)
{
    const int64_t n = n_time * n_rates;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const double t = time[i / n_rates];
        const double r = rates[i % n_rates];
        const double a = __dmul_rn(__dmul_rn(16.0, __dmul_rn(r, r)), t);
        const double x = -__dmul_rn(__dmul_rn(4.0, r), t);
        out[i] = __dmul_rn(a, exp(x));
    }
}

// numpy.nansum(pi, axis=1) of a row-major (n_rows, n_cols) matrix: one warp per row, NaN terms skipped
// (compute.py:78, bin/tapir_compute.py:119).
__global__ void nansum_rows_kernel(const double* __restrict__ m, int64_t n_rows, int64_t n_cols, double* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const double* p = m + row * n_cols;
    double s = 0.0;
    for (int64_t j = lane; j < n_cols; j += 32) {
        const double v = p[j];
        s += (v == v) ? v : 0.0;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if (lane == 0) out[row] = s;
}

__global__ void site_integrals_kernel(double a, double b, const double* __restrict__ rates, int64_t n,
                                      double* __restrict__ integral, double* __restrict__ abserr,
                                      int32_t* __restrict__ panels)
{
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double lam = rates[i];
    double res, err;
    int last;
    qags_device(a, b, 16.0 * (lam * lam), 4.0 * lam, &res, &err, &last);
    integral[i] = res;
    abserr[i] = err;
    if (panels) panels[i] = last;
}

// ------------------------------------------------------------------------------------------
// measurement helper: FP64 pipe peak.  A register-resident 8x4 outer-product of independent DFMAs
// (64 per iteration, no memory traffic); tools/microbench/fp64_pipes.cu shows this form reaches the
// pipe's 64 FMA/clk/SM while a plain x = fma(x, m, b) chain stops ~13 % short of it.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) fp64_peak_kernel(double* sink, int iters, const double* seed)
{
    double acc[8][4], u0[8], w0[4], u1[8], w1[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { u0[i] = seed[i] + threadIdx.x; u1[i] = seed[8 + i] * 0.5; }
#pragma unroll
    for (int j = 0; j < 4; ++j) { w0[j] = seed[16 + j]; w1[j] = seed[20 + j] + threadIdx.x; }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(u0[i], w0[j], acc[i][j]);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(u1[i], w1[j], acc[i][j]);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += acc[i][j];
    if (s == 123.456) sink[0] = s;  // never true; keeps the accumulators alive
}

}  // namespace tapir
