/*
 * tapir_cuda.h -- C ABI of libtapir_cuda.so, the B200 (sm_100a) implementation of tapir's
 * Townsend phylogenetic-informativeness hot path.
 *
 * The reference (faircloth-lab/tapir) is pure Python and has no FFI; the boundary this
 * library replaces is the per-locus arithmetic that `worker()` performs through
 * `tapir.compute` (citations relative to the reference checkout):
 *
 *   tapir/compute.py:46-48   get_townsend_pi            PI(t,i) = 16 r_i^2 t exp(-4 r_i t)
 *   bin/tapir_compute.py:119 numpy.nansum(pi, axis=1)    -> pi_net[t]        (`net` table)
 *   tapir/compute.py:76-79   get_net_pi_for_periods      -> pi_net[times]    (`discrete` table)
 *   tapir/compute.py:50-52   get_integral_over_times     scipy quad == QUADPACK QAGS / GK21
 *   tapir/compute.py:81-94   get_net_integral_for_epochs -> sum(integral), sum(error) (`interval`)
 *   bin/tapir_compute.py:110 masked mean of the rates    -> mean_rate
 *   README.rst:178-185       per-locus max(pi) over intervals -> argmax_epoch
 *
 * One call processes MANY loci packed as CSR (offsets, rates); the reference maps worker()
 * over loci one at a time (bin/tapir_compute.py:159-164).
 *
 * Conventions: every entry point returns 0 on success or a negative TAPIR_E_* code and never
 * throws or exits; tapir_cuda_last_error() gives the message.  The caller owns all host
 * buffers; the library owns device memory and streams inside a context.  A context is bound to
 * one device and is not thread-safe; distinct contexts are independent.  There is NO CPU
 * fallback: without a CUDA device every compute entry point fails with TAPIR_E_CUDA.
 */
#ifndef TAPIR_CUDA_H
#define TAPIR_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAPIR_CUDA_ABI_VERSION 2

/* error codes */
#define TAPIR_OK 0
#define TAPIR_E_INVALID (-1) /* bad argument (NULL, negative size, a >= b epoch, time index out of range) */
#define TAPIR_E_CUDA (-2)    /* CUDA runtime / driver error, or no device */
#define TAPIR_E_NOMEM (-3)   /* host or device allocation failed */
#define TAPIR_E_STATE (-4)   /* call order error (compute before upload, ...) */

/* tapir_problem.flags */
#define TAPIR_WANT_PROFILE 1u /* copy pi_net[L*T] back to the host (it is always computed) */
/* FP32 mode (BASELINE: "FP64, or FP32 with compensated summation ... <= 1e-5 in FP32 mode"): the PI profile is summed in
 * single precision by profile_f32_kernel (per-lane partial sums closed by a reduction tree) and the epoch sums of the
 * slow sites are taken from its FP32 moments; outputs stay float64 arrays, accurate to ~1e-6 relative (gate 1e-5).
 * Per-site integrals (fast sites, rate statistics) and everything else stay FP64.  Also switched on for every call of a
 * context created with TAPIR_FP32_MODE=1 in the environment. */
#define TAPIR_FP32_MODE 2u

#define TAPIR_MAX_EPOCHS 64
#define TAPIR_MAX_TIMES 4096

typedef struct tapir_cuda_ctx tapir_cuda_ctx;

/* One batch of loci.  All pointers are HOST pointers for tapir_cuda_run / tapir_cuda_upload. */
typedef struct tapir_problem {
    int64_t n_loci;          /* L */
    const int64_t *offsets;  /* [L+1], offsets[0] == 0, non-decreasing: CSR row pointers into rates */
    const double *rates;     /* [offsets[L]] corrected site rates; NaN/Inf = culled site (skipped) */
    int64_t n_grid;          /* T: the grid is t = 0..T-1, step 1 (compute.py:54-57, get_time(0,int(depth))) */
    int32_t n_times;         /* K */
    const int64_t *time_idx; /* [K] grid indices as numpy fancy indexing takes them: -T <= idx < T */
    int32_t n_epochs;        /* E <= TAPIR_MAX_EPOCHS */
    const double *epoch_ab;  /* [2E] a0,b0,a1,b1,... with a < b (compute.py:90) */
    uint32_t flags;
} tapir_problem;

/* Result arrays (host pointers for run/download; device pointers from tapir_cuda_device_results).
 * Any pointer may be NULL to skip that output. */
typedef struct tapir_results {
    double *pi_net;         /* [L*T]  net PI profile (only filled when TAPIR_WANT_PROFILE) */
    double *pi_times;       /* [L*K]  pi_net at time_idx */
    double *integral;       /* [L*E]  sum over finite sites of the QAGS integral */
    double *abserr;         /* [L*E]  sum over finite sites of the QAGS error estimate */
    double *max_pi;         /* [L]    max_t pi_net */
    int64_t *argmax_t;      /* [L]    first t attaining it (numpy.argmax) */
    int32_t *argmax_epoch;  /* [L]    first epoch (in the given order) with the largest integral */
    double *mean_rate;      /* [L]    mean over non-NaN rates (NaN when there are none) */
    int64_t *n_finite;      /* [L]    number of finite-rate sites (the sites that are integrated) */
    int64_t *n_multi_panel; /* [L]    (site,epoch) pairs for which QAGS went beyond one GK21 panel */
} tapir_results;

/* Device-side timing of the last tapir_cuda_compute (CUDA events on the library's stream). */
typedef struct tapir_timing {
    float total_ms;    /* first launch -> last launch finished */
    float profile_ms;  /* PI profile kernel */
    float quad_ms;     /* GK21/QAGS epoch-integral kernel(s) */
    float finalize_ms; /* per-locus reduction / picks / maxima kernel */
    int32_t launches;  /* kernel launches issued */
    /* what was launched for the profile: points per time tile (256 .. 2048), the moment form's moments per block
     * (0 = profile_kernel, direct form), and how many of the batch's epochs took their sums from the moments */
    int32_t profile_tile_points;
    int32_t profile_moment_form; /* 0, or the number of moments per 128-point block (32 or 16) */
    int32_t moment_epochs;
    int32_t profile_fp32;        /* 1: profile_f32_kernel ran (TAPIR_FP32_MODE) */
    int32_t reserved;
} tapir_timing;

int tapir_cuda_version(void);
int tapir_cuda_device_count(int *count);
/* message of the last failure on this thread (ctx may be NULL) */
const char *tapir_cuda_last_error(const tapir_cuda_ctx *ctx);

int tapir_cuda_create(int device, tapir_cuda_ctx **ctx);
int tapir_cuda_destroy(tapir_cuda_ctx *ctx);

/* pinned host memory for the CSR pack and the result arrays */
int tapir_cuda_host_alloc(void **ptr, size_t bytes);
int tapir_cuda_host_free(void *ptr);

/* --- whole path with HOST buffers: H2D, kernels, D2H (chunked and overlapped) --- */
int tapir_cuda_run(tapir_cuda_ctx *ctx, const tapir_problem *problem, const tapir_results *host_out);

/* --- staged path: inputs resident in HBM, kernels timed on their own --- */
int tapir_cuda_upload(tapir_cuda_ctx *ctx, const tapir_problem *problem);
int tapir_cuda_compute(tapir_cuda_ctx *ctx, tapir_timing *timing /* may be NULL */);
int tapir_cuda_download(tapir_cuda_ctx *ctx, const tapir_results *host_out);
/* device pointers of the resident result arrays (valid until the next upload/destroy) */
int tapir_cuda_device_results(tapir_cuda_ctx *ctx, tapir_results *device_ptrs);

/* The small per-locus result vectors of the resident batch live in ONE device block (this is what a multi-GPU
 * driver gathers): *block / *bytes describe it, section_offset[i] is the byte offset of, in order,
 * pi_times, integral, abserr, max_pi, argmax_t, mean_rate, n_finite, n_multi_panel, argmax_epoch. */
int tapir_cuda_device_small_block(tapir_cuda_ctx *ctx, void **block, size_t *bytes, size_t section_offset[9]);

/* --- ranking queries on the resident, computed batch (results stay on the device until asked for) ---
 * top_locus / top_pi [E][n_top]: per epoch the n_top (<= 1024) loci with the largest summed integral, descending,
 * ties to the lower locus index -- the "ORDER BY pi DESC limit N" of bin/tapir_compare.py:93-101; rows are padded
 * with locus -1 / NaN when the batch has fewer loci.  epoch_count[E]: number of loci whose max-PI epoch is e
 * (README.rst:178-185; a tie counts for the first epoch).  Any output pointer may be NULL. */
int tapir_cuda_rank(tapir_cuda_ctx *ctx, int32_t n_top, int64_t *top_locus, double *top_pi, int64_t *epoch_count);

/* --- multi-GPU gather without a collective launch: every rank pushes its small block over NVLink into one
 * buffer that lives on the owner's GPU (CUDA IPC peer memory), on the same stream right after finalize ---
 * owner (rank 0): tapir_cuda_gather_create allocates n_slots * slot_bytes and returns a 64-byte IPC handle to
 * pass to the other processes; every rank (the owner with handle == NULL) then calls tapir_cuda_gather_attach
 * with its slot; from then on each tapir_cuda_compute ends with the push.  tapir_cuda_gather_read copies the
 * whole buffer to the owner's host after the ranks have synchronised (e.g. a barrier). */
#define TAPIR_IPC_HANDLE_BYTES 64
int tapir_cuda_gather_create(tapir_cuda_ctx *ctx, size_t slot_bytes, int n_slots, void *handle_out);
int tapir_cuda_gather_attach(tapir_cuda_ctx *ctx, const void *handle, size_t slot_bytes, int slot);
int tapir_cuda_gather_read(tapir_cuda_ctx *ctx, void *host_dst, size_t bytes);
int tapir_cuda_gather_detach(tapir_cuda_ctx *ctx);

/* serial != 0: run the profile and the epoch-integral kernels back to back on one stream instead of
 * concurrently on two (per-kernel tapir_timing fields are only exact in this mode; also set by
 * TAPIR_CUDA_SERIAL_KERNELS=1 in the environment at context creation) */
int tapir_cuda_set_serial_kernels(tapir_cuda_ctx *ctx, int serial);

/* bytes moved host->device / device->host by this context since creation (for bench.py's e2e accounting) */
int tapir_cuda_transfer_bytes(tapir_cuda_ctx *ctx, uint64_t *h2d, uint64_t *d2h);

/* --- element-wise API parity with tapir.compute --- */
/* get_townsend_pi(time, rates) -> out[T*S] row-major (T,S); NaN propagate as numpy does */
int tapir_cuda_pi_matrix(tapir_cuda_ctx *ctx, const double *time, int64_t n_time, const double *rates,
                         int64_t n_rates, double *out);
/* numpy.nansum(pi, axis=1) of a row-major (n_rows, n_cols) matrix (compute.py:78, bin/tapir_compute.py:119) */
int tapir_cuda_nansum_rows(tapir_cuda_ctx *ctx, const double *matrix, int64_t n_rows, int64_t n_cols, double *out);
/* vectorize(get_integral_over_times)(a, b, rates) -> integral[n], abserr[n], panels[n] (QAGS `last`) */
int tapir_cuda_site_integrals(tapir_cuda_ctx *ctx, double a, double b, const double *rates, int64_t n,
                              double *integral, double *abserr, int32_t *panels);

/* --- plan of the moment-functional epoch integrals (host only, no device needed; what tapir_cuda_upload decides
 * per epoch when the profile runs in moment form, exported so that tests can pin it against QUADPACK) ---
 * For the epoch [a, b] on a grid of n_tiles tiles of tile_points points: eligible != 0 when the sum over a locus'
 * slowly decaying sites of scipy.integrate.quad's value (tapir/compute.py:50-52, 81-94) is taken as
 * sum_j p[j] M[j][hb] from the n_moments (32 or 16) moments of block hb of tile `tile` (block centre tb); only sites
 * with r_min <= rate <= r_max outside [w_lo, w_hi] (and below the moment class bound, 0.01953125 for 32 moments,
 * 0.0048828125 for 16) are served that way.  class_x = 0 selects that FP64 class; class_x = 2.5 with 16 moments is the
 * FP32 mode's (bound 0.009765625). */
typedef struct tapir_mom_epoch {
    int32_t eligible;
    int32_t tile, hb;
    double tb;
    double p[32];
    double r_min, r_max;
    double w_lo, w_hi;
} tapir_mom_epoch;
int tapir_mom_epoch_plan(double a, double b, int64_t n_tiles, int32_t tile_points, int32_t n_moments, double class_x,
                         tapir_mom_epoch *out);

/* --- batched rate ingest (host only, no device needed): parse_site_rates' read half (compute.py:30-39) for
 * n_files HyPhy site-rate JSON files on n_threads threads into one CSR pack.  offsets[n_files + 1] is always
 * filled and *needed = total number of sites; rates (each divided by `correction`) is filled only when
 * capacity >= *needed, so a caller may first call with capacity 0 to size a (pinned) buffer.  Files are not
 * modified.  Returns TAPIR_E_INVALID on an unreadable or malformed file (tapir_ingest_last_error()). */
int tapir_ingest_rates(const char *const *paths, int64_t n_files, int n_threads, double correction, int64_t *offsets,
                       double *rates, int64_t capacity, int64_t *needed);
/* The same in three steps, so that every file is read and parsed once: open parses all files and keeps the rates
 * (offsets[n_files + 1] and *needed are filled), copy writes rates / correction into a buffer of >= *needed doubles
 * (e.g. pinned memory sized from *needed), close releases the handle.  A file json.load would reject (truncated,
 * trailing data) is an error. */
typedef struct tapir_ingest tapir_ingest;
int tapir_ingest_open(const char *const *paths, int64_t n_files, int n_threads, tapir_ingest **handle, int64_t *offsets,
                      int64_t *needed);
int tapir_ingest_copy(const tapir_ingest *handle, double correction, double *rates, int64_t capacity);
int tapir_ingest_close(tapir_ingest *handle);
const char *tapir_ingest_last_error(void);

/* --- measurement helpers used by bench.py --- */
/* sustained FP64 FMA rate of this device, TFLOP/s (2 flop per DFMA), dependent-chain-free register kernel */
int tapir_cuda_fp64_peak(tapir_cuda_ctx *ctx, double *tflops, float *ms);
/* overwrite a buffer larger than L2 so the next timed step starts cold */
int tapir_cuda_flush_l2(tapir_cuda_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* TAPIR_CUDA_H */
