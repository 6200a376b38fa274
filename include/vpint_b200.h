/*
 * vpint_b200.h -- C ABI of the B200-native VPint value-propagation path.
 *
 * The reference (LaurensArp/VPint 0.2.4) is pure Python + NumPy and has no
 * plugin / FFI layer; its boundary for this path is the Python method surface
 *   SD_SMRP.run   VPint/SD_MRP.py:51      WP_SMRP.run   VPint/WP_MRP.py:82-91
 *   SD_STMRP.run  VPint/SD_MRP.py:286     WP_STMRP.run  VPint/WP_MRP.py:1289
 *   VPint2_interpolator.run / run_serial  VPint/VPint2.py:58,109
 * This header is the C boundary a maintainer would bind (ctypes stub in
 * INTEGRATION.md) to run those methods on a B200.  Each entry point names the
 * reference code it replaces.
 *
 * Conventions
 *   - plain C: opaque handle, pointers and sizes, int status (0 = ok).  No
 *     exceptions, no torch / Python types.  vpint_last_error() gives the text.
 *   - every "device pointer" is caller-owned CUDA memory on the current device;
 *     the library allocates nothing on the device except through the workspace
 *     the caller binds (vpint_solver_workspace_bytes / vpint_solver_bind).
 *   - all work is enqueued on the caller's stream (cudaStream_t passed as
 *     void*); only vpint_solver_run blocks the host (it polls one pinned flag
 *     per chunk of sweeps, never per sweep).
 *   - arithmetic is IEEE fp64 in both storage modes; VPINT_F32 only changes the
 *     element type of the device-resident state and weights.
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point fails with VPINT_ERR_CUDA.
 */
#ifndef VPINT_B200_H
#define VPINT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPINT_ABI_VERSION 2

/* status codes */
#define VPINT_OK               0
#define VPINT_ERR_INVALID      1   /* bad argument / shape / state            */
#define VPINT_ERR_CUDA         2   /* CUDA runtime error (text in last_error)  */
#define VPINT_ERR_UNSUPPORTED  3   /* valid request this build cannot serve    */
#define VPINT_ERR_WORKSPACE    4   /* workspace missing or too small           */

/* element types */
#define VPINT_F64 0
#define VPINT_F32 1
#define VPINT_U16 2   /* raw rasters of the fused VPint2 load only (Sentinel-2 bands, VPint/utils/EO_utils.py:66) */
#define VPINT_U8  3   /* cloud masks only */

/* weight modes */
#define VPINT_W_SCALAR 0   /* SD_*: per-problem gamma (and tau in 3-D)          */
#define VPINT_W_PLANES 1   /* WP_*: 4 static planes (2-D) / 4 spatial + 1 temporal (3-D) */

/* pairwise weight methods, WP_SMRP.predict_weight (VPint/WP_MRP.py:396-427) */
#define VPINT_METHOD_EXACT          0
#define VPINT_METHOD_COSINE         1
#define VPINT_METHOD_EXACT_INVERSE  2

/* loop drivers for vpint_solver_run */
#define VPINT_LOOP_CHUNKED 0   /* host enqueues chunks of launches, polls a pinned flag per chunk */
#define VPINT_LOOP_GRAPH   1   /* CUDA graph with a device-side WHILE node (conditional graph node): one graph
                                  launch and one host sync per run, no per-sweep host work at all */

typedef struct vpint_solver vpint_solver;

/* Shape of one batched solve.  Every problem in the batch has the same shape;
 * problems are independent (own mask, weights, stopping sweep), which is how
 * VPint2 bands (VPint/VPint2.py:71-83), patches and hyper-parameter trials map
 * onto one launch. */
typedef struct vpint_desc {
    int32_t ndim;         /* 2: (H, W) grids.  3: (H, W, T) cubes (any strides; device state is time-major) */
    int32_t nprob;        /* problems in the batch                                   */
    int32_t nprob_inner;  /* problem p = outer * nprob_inner + inner for addressing caller arrays
                             (patch, band); 0 or 1 = flat                            */
    int32_t H, W, T;      /* T must be 1 when ndim == 2                              */
    int32_t storage;      /* VPINT_F64 or VPINT_F32 device state                     */
    int32_t weight_mode;  /* VPINT_W_SCALAR or VPINT_W_PLANES                        */
    int32_t k_block;      /* sweeps fused per launch (temporal blocking); 0 = default */
    /* Row-band sharding of ONE global grid across ranks.  For an
     * unsharded solve leave all four at 0 (then global_H = H, everything owned). */
    int32_t global_H;     /* rows of the whole grid                                  */
    int32_t row_offset;   /* global row index of local row 0 (may be negative: never)*/
    int32_t own_row0;     /* first local row this rank owns                          */
    int32_t own_rows;     /* number of owned local rows (the rest are halo rows)     */
    int32_t features;     /* VPINT_FEAT_* bits: optional workspace the solve will need */
} vpint_desc;

/* vpint_desc.features */
#define VPINT_FEAT_RESISTANCE 1   /* keeps a copy of the known values: run() may use `resistance` */
#define VPINT_FEAT_BLOCKING_WAIT 2 /* host waits of load() and of the cluster-resident run() sleep instead of
                                      spinning: for callers that drive several solvers from several host threads
                                      (the pipelined VPint2 path), where spinning threads would fight for cores */

/* Arguments of one run() call.  Defaults of the reference in brackets. */
typedef struct vpint_run_params {
    int32_t max_iter;        /* sweep cap: `iterations` if >-1, else 5000 (WP 2-D,
                                VPint/WP_MRP.py:112-115) or 10000 (SD_MRP.py:63-66,296-299) */
    int32_t auto_terminate;  /* stop when delta <= threshold [True unless iterations>-1] */
    double  threshold;       /* auto_terminate_threshold [1e-4]                         */
    double  clip_val;        /* new[new > clip_val] = clip_val [inf], WP_MRP.py:322     */
    int32_t loop_mode;       /* VPINT_LOOP_*                                            */
    int32_t chunk;           /* launches per host poll in chunked mode; 0 = default     */
    double *delta_out;       /* device, [nprob][max_iter] per-sweep delta (track_delta,
                                SD_MRP.py:140-141) or NULL                              */
    int32_t resistance;      /* WP_SMRP 'elastic band' damping (WP_MRP.py:327-348): after the restore every
                                cell -- known ones too -- with new > mu becomes old + (new-old) - epsilon*old
                                [False].  Needs VPINT_FEAT_RESISTANCE and k_block 1.           */
    double  epsilon, mu;     /* [0.01, 1.0]                                                     */
} vpint_run_params;

/* ---- library ---------------------------------------------------------- */
int         vpint_abi_version(void);
const char *vpint_last_error(void);              /* thread-local, never NULL */
int64_t     vpint_row_pitch(int64_t row_elems);  /* padded device row length (elements) */

/* ---- solver life cycle ------------------------------------------------ */
int    vpint_solver_create(const vpint_desc *desc, vpint_solver **out);
void   vpint_solver_destroy(vpint_solver *s);
size_t vpint_solver_workspace_bytes(const vpint_solver *s);
/* workspace: device memory, 256-byte aligned, >= workspace_bytes.  It must stay
 * valid until destroy / the next bind. */
int    vpint_solver_bind(vpint_solver *s, void *workspace, size_t bytes);

/* ---- A1 + known-cell bookkeeping --------------------------------------
 * Replaces the per-run setup around MRP.original_grid / pred_grid
 * (VPint/MRP.py:47-52, restore at SD_MRP.py:133-134, WP_MRP.py:321,323).
 *   original : device, dtype `dtype`, NaN marks an unknown cell.
 *   pred0    : device, same layout, the state run() starts from
 *              (self.pred_grid); may be NULL, then unknown cells start at
 *              fill[p] (the init_strategy='mean'/'zero' value computed by the
 *              caller exactly as MRP.init_pred_grid does, MRP.py:64-82).
 *   fill     : HOST array [nprob]; only read when pred0 == NULL.  With pred0 == NULL and
 *              fill == NULL the init_strategy='mean' value is computed on the device:
 *              np.nanmean of each problem's grid in ITS dtype, bit-identical to NumPy's
 *              pairwise summation (grids up to 131072 cells in one kernel; for larger grids call
 *              vpint_solver_fill_leaves + vpint_solver_fill_combine on the band-interleaved raster
 *              first -- the next load with pred0 == NULL and fill == NULL uses those values -- or
 *              pass fill).
 *   stride   : 5 element strides {outer problem, inner problem, row, column, t}
 *              of original/pred0 (t ignored for ndim == 2).  This is how the
 *              band-interleaved (N, H, W, B) VPint2 arrays are read in place:
 *              outer = patch, inner = band.
 * Builds the bit mask of unknown cells, both ping-pong state buffers, the
 * active-tile list and the sums over known cells the stopping rule needs. */
int vpint_solver_load(vpint_solver *s, const void *original, const void *pred0,
                      const double *fill, int dtype, const int64_t *stride, void *stream);

/* ---- A2 weights -------------------------------------------------------
 * VPINT_W_PLANES only.  feat: device (..., F) features with 5 element strides
 * {outer problem, inner problem, row, column, feature}; zeros are read as 0.01
 * (WP_MRP.py:143-145).
 * 2-D exact:  w_dir = mean_k f[i,j,k] / f[nb_dir,k], missing neighbour -> 0
 *             (WP_MRP.py:149-170).  method selects predict_weight's variants
 *             (WP_MRP.py:396-427) with the [min_gamma, max_gamma] clamp, which
 *             the vectorised exact branch does not apply (pass clamp = 0).
 * 3-D      :  features are (H, W, F), static in time (WP_MRP.py:1319-1355). */
int vpint_solver_weights(vpint_solver *s, const void *feat, int dtype, const int64_t *stride,
                         int F, int method, int clamp, double min_gamma, double max_gamma,
                         void *stream);
/* prioritise_identity (WP_MRP.py:243-276, 309-315), 2-D VPINT_W_PLANES only, after vpint_solver_weights.
 * Builds the static priority planes p from the weights (beta == 0: ones with zero borders; otherwise the
 * reference's two sequential rewrites p>1 -> (1/p)*((1/p)/((1/p)*beta)), then p<1 -> p*(p/((p+0.001)*beta))),
 * optionally multiplies them by `bias` (device double [nprob][H][W], the known_value_bias sub-solve of
 * WP_MRP.py:263-276; zeros of the product become 0.01) and folds them into the weights:
 * w_dir <- w_dir * p_dir / sum(p), with the division by the neighbour count switched off -- the
 * update new = sum(w*v*p) / sum(p) of WP_MRP.py:311-315.  p_out (device double [nprob][H][W][4], may be
 * NULL) receives the planes (self.priority_grid). */
int vpint_solver_priority(vpint_solver *s, double beta, const double *bias, double *p_out, void *stream);

/* Batched trials of a parameter search (WP_SMRP.auto_adapt, VPint/WP_MRP.py:429-759: the same grid solved with many
 * candidate (beta, epsilon, mu)): HOST arrays [nprob], NULL = not per problem.  beta replaces the scalar of the next
 * vpint_solver_priority call, epsilon / mu (together) the scalars of vpint_run_params in resistance runs.  Stays in
 * force until called again (all NULL switches it off). */
int vpint_solver_set_trial_params(vpint_solver *s, const double *beta, const double *epsilon, const double *mu, void *stream);

/* VPINT_W_SCALAR: HOST arrays [nprob]; tau is ignored for ndim == 2
 * (SD_MRP.py:87, 315-321).  The copy is ordered on `stream` like every other call (the host arrays may be
 * reused as soon as the call returns). */
int vpint_solver_set_discounts(vpint_solver *s, const double *gamma, const double *tau, void *stream);

/* ---- A4 + A5 the loop --------------------------------------------------
 * Jacobi sweeps with the reference's stopping rule
 *   delta = nanmean(|new-old|) / nanmean(old) <= threshold   (SD_MRP.py:138-147,
 *   WP_MRP.py:351-362), commit-then-break, exact stopping sweep (temporal-block
 * overshoot is rolled back and replayed).  Blocks the host until every problem
 * stopped or hit max_iter.  Can be called again to continue from the current
 * state (run(50); run(50) == run(100)). */
int vpint_solver_run(vpint_solver *s, const vpint_run_params *prm, void *stream);

/* Split form of ONE launch for callers that put a collective between the
 * local sums and the stop decision (row-sharded multi-GPU):
 *   step_begin : k_block sweeps on the local rows + local per-sweep sums
 *   (caller all-reduces vpint_solver_sums_ptr(): [nprob][k_block][4] doubles,
 *    and exchanges halo rows of vpint_solver_state_ptr())
 *   step_end   : stop decision; *active_out (HOST, may be NULL) is only valid
 *                after the stream is synchronised. */
int     vpint_solver_begin_run(vpint_solver *s, const vpint_run_params *prm, void *stream);
int     vpint_solver_step_begin(vpint_solver *s, void *stream);
int     vpint_solver_step_end(vpint_solver *s, void *stream);
/* step_begin with CUDA events around the stencil launch only (on `stream`); synchronises and
 * returns that kernel's duration in milliseconds.  For roofline measurements. */
int     vpint_solver_step_profile(vpint_solver *s, void *stream, float *stencil_ms);
double *vpint_solver_sums_ptr(vpint_solver *s);
int     vpint_solver_sums_count(const vpint_solver *s);
void   *vpint_solver_state_ptr(vpint_solver *s, int which);   /* ping-pong buffer 0/1, [nprob][H][pitch] (3-D: [nprob][T][H][pitch]) */
int32_t*vpint_solver_active_ptr(vpint_solver *s);             /* device int32: problems still running   */
int32_t*vpint_solver_cur_ptr(vpint_solver *s);                /* device int32[nprob]: buffer holding the current state */

/* Row-sharded scenes, used between step_end and the next step_begin.
 *   known_ptr : [2][nprob][4] doubles, the known-cell terms of the stopping rule summed over the OWNED
 *               rows at load; all-reduce (SUM) once after vpint_solver_load so that every rank divides
 *               by the statistics of the whole grid (VPint/SD_MRP.py:139).
 *   halo_pack : copies the `rows` first / last owned rows of every problem's current state into the
 *               contiguous messages to_up / to_down ([nprob][rows][pitch], storage type; NULL = skip).
 *   halo_unpack : writes messages received from the upper / lower neighbour into the halo rows just
 *               above / below the owned band.
 *   any_dirty : nonzero when the start state differed from `original` on a known cell (such a start
 *               state must be avoided in sharded runs: ranks would disagree on the first launch). */
double *vpint_solver_known_ptr(vpint_solver *s);
int     vpint_solver_known_count(const vpint_solver *s);
int     vpint_solver_any_dirty(const vpint_solver *s);
int     vpint_solver_halo_pack(vpint_solver *s, int rows, void *to_up, void *to_down, void *stream);
int     vpint_solver_halo_unpack(vpint_solver *s, int rows, const void *from_up, const void *from_down, void *stream);

/* Non-blocking form of vpint_solver_run for solves the cluster-resident kernel serves (small 2-D grids, one
 * feature layer or scalar gamma): run_async enqueues the whole run() -- one kernel launch -- and returns; the host
 * may enqueue vpint_solver_result* behind it and go on to other solvers / streams.  run_finish waits for it and, in
 * the rare case the kernel declined (non-finite data), completes the run on the tiled path -- after which results
 * read earlier must be read again; it returns 1 in *reran (may be NULL) then.  For solves the resident kernel
 * does not serve run_async only records the parameters and run_finish does all the work (== vpint_solver_run). */
int vpint_solver_run_async(vpint_solver *s, const vpint_run_params *prm, void *stream);
int vpint_solver_run_finish(vpint_solver *s, void *stream, int *reran);

/* ---- fused VPint2 load: constructor + A1 + A2 from the raw rasters ---------------------------------------
 * Replaces, for a band-interleaved stack (n_img, H, W, B) -- nprob = n_img * B, nprob_inner = B --
 *   VPint2_interpolator.__init__ (VPint/VPint2.py:13-55): astype(DTYPE), clip_target / clip_features, apply_mask
 *   (mask > threshold -> NaN on every band, :122-132), buffer_clouds (:135-150, incl. the column bound of :146),
 *   MRP.init_pred_grid 'mean' (VPint/MRP.py:73-77: np.nanmean in DTYPE, NumPy's pairwise summation bit for bit, ANY
 *   grid size) and the zero -> 0.01 feature rewrite of WP_SMRP.run (VPint/WP_MRP.py:143-145)
 * with device kernels that read the rasters once.  All pointers are device memory unless stated otherwise. */

/* cloud[i] = mask[i] > threshold (mask dtype VPINT_F64 / F32 / U16 / U8; n = n_img * H * W). */
int vpint_cloud_mask(const void *mask, int dtype, double threshold, int64_t n, uint8_t *cloud_out, void *stream);

/* buffer_clouds on the all-band mask.  cloud (in/out, [n_img][H][W]) holds the thresholded mask when have_cloud,
 * else it is output only.  target (dtype F32 / F64; U16 has no NaN) contributes pixels with a NaN in any band to
 * the dilated footprint.  scratch: 3 * n_img * H * W bytes. */
int vpint_buffer_clouds(const void *target, int dtype, uint8_t *cloud, int have_cloud, int n_img, int H, int W, int B,
                        int buffer_size, int passes, uint8_t *scratch, void *stream);

typedef struct vpint_ingest_params {
    const void *target, *features;  /* (n_img, H, W, B) contiguous, band fastest; H = the solver's (local) rows   */
    int32_t raster_dtype;           /* VPINT_U16 / VPINT_F32 / VPINT_F64, both rasters                             */
    int32_t value_dtype;            /* the interpolator's DTYPE (VPINT_F32 = VPint2's default, or VPINT_F64)       */
    const uint8_t *cloud;           /* [n_img][H][W], nonzero = unknown on every band; may be NULL                 */
    int32_t clip_target, clip_features;             /* 0 / 1                                                      */
    double  target_min, target_max, features_min, features_max;
    int32_t fill_mode;              /* VPINT_FILL_DEVICE, VPINT_FILL_READY or VPINT_FILL_HOST                     */
    const double *fill_host;        /* HOST [nprob], VPINT_FILL_HOST only                                          */
    void   *scratch;                /* VPINT_FILL_DEVICE: vpint_fill_scratch_bytes() bytes                         */
    void   *out;                    /* known cells of the result are written here (full output); may be NULL      */
    int32_t out_dtype;              /* VPINT_F32 / VPINT_F64 / VPINT_U16 (write-back of a product: clip to [0, 65535],
                                       truncate, NaN -> 0)                                                        */
    int64_t out_stride[4];          /* element strides {image, band, row, column} of out                          */
} vpint_ingest_params;

#define VPINT_FILL_DEVICE 0   /* np.nanmean per band on the device, inside vpint_solver_ingest_bands (unsharded)     */
#define VPINT_FILL_READY  1   /* the solver already holds the fill values (vpint_solver_fill_combine)                */
#define VPINT_FILL_HOST   2   /* fill_host[nprob]                                                                    */

size_t vpint_fill_scratch_bytes(int64_t cells_global, int nprob, int value_dtype);
/* Row shards (and anyone who wants the mean alone): fill_leaves sums the leaves of NumPy's summation tree that START in
 * this solver's owned rows into scratch (zeros elsewhere; a leaf runs up to 127 cells past the owned rows, which the
 * halo rows must cover), the caller all-reduces (SUM) vpint_fill_leaf_count() values of value_dtype at scratch and
 * nprob uint64 counts at scratch + vpint_fill_count_offset(), and fill_combine adds the leaf sums up in NumPy's
 * order -- every rank gets the bit-identical np.nanmean of the WHOLE band. */
int     vpint_solver_fill_leaves(vpint_solver *s, const vpint_ingest_params *p, void *stream);
int64_t vpint_fill_leaf_count(int64_t cells_global, int nprob);
size_t  vpint_fill_count_offset(int64_t cells_global, int nprob, int value_dtype);
int     vpint_solver_fill_combine(vpint_solver *s, int value_dtype, void *scratch, void *stream);
double *vpint_solver_fill_ptr(vpint_solver *s);                /* device double[nprob] */

/* The load itself: state, unknown-cell mask, feature plane, known-cell sums (what vpint_solver_load +
 * vpint_solver_weights(exact, F = 1) build), without a host synchronisation. */
int vpint_solver_ingest_bands(vpint_solver *s, const vpint_ingest_params *p, void *stream);

/* The unknown cells of the result only (the known ones are the caller's own input, or were written by
 * vpint_solver_ingest_bands through vpint_ingest_params.out).  out may be device memory or pinned host memory
 * mapped into the device (cudaHostAlloc / cudaHostRegister): the stores leave as contiguous runs, so only the
 * cloud cells cross the host link.  dtype VPINT_F64 / VPINT_F32 / VPINT_U16.  stride: {outer problem, inner problem,
 * row, column}. */
int vpint_solver_result_unknown(vpint_solver *s, void *out, int dtype, const int64_t *stride, int32_t *niter_out,
                                void *stream);

/* Low-latency form of the row-sharded loop (k_block 1): the stop decision lags one sweep, so the all-reduce of sweep
 * t's sums travels while sweep t + 1 computes, and halo rows go straight into the neighbour's memory.
 *   step_begin                 : sweep t + local sums (as above)
 *   (caller copies vpint_solver_sums_ptr() aside and all-reduces the copy on a side stream)
 *   step_commit_lagged(prev)   : prev = the reduced sums of sweep t - 1 (device, [nprob][4]; NULL in the first launch of
 *                                a run): problems whose delta(t - 1) <= threshold stop WITHOUT committing sweep t (commit-
 *                                then-break with the exact stopping sweep, VPint/SD_MRP.py:143-147); the others commit it
 *   halo_push / halo_pull      : the `rows` boundary rows of the current state are stored into the neighbours' receive
 *                                slots ([nprob][rows][pitch] doubles in THEIR memory, mapped into this process: CUDA IPC /
 *                                symmetric memory over NVLink) and the sequence number `seq` is published in their flag;
 *                                pull waits for flag >= seq and copies the received rows into the halo rows.  Receive
 *                                slots are double buffered by the caller (a sender is at most one exchange ahead). */
int vpint_solver_step_commit_lagged(vpint_solver *s, const double *sums_prev, void *stream);
int vpint_solver_halo_push(vpint_solver *s, int rows, void *peer_up, int32_t *flag_up, void *peer_down, int32_t *flag_down,
                           int seq, void *stream);
int vpint_solver_halo_pull(vpint_solver *s, int rows, const void *from_up, const int32_t *flag_up, const void *from_down,
                           const int32_t *flag_down, int seq, void *stream);

/* The same loop entirely inside the library, without a collective library in it: every rank owns a PEER BLOCK of
 * vpint_solver_peer_bytes() bytes (zero-filled symmetric / IPC memory that every rank has mapped; peer_base[r] = rank
 * r's block as seen from this process, ranks in row order).  Per launch: sweep + local sums, ONE kernel that waits for
 * the previous sweep's sums of all ranks, adds them up in rank order (every rank the same order -> the same decision),
 * decides, commits and publishes this sweep's sums in every rank's block, ONE kernel that pushes / pulls the halo rows.
 * The host stays a few launches ahead and reads "problems still running" from mapped pinned memory; all ranks stop
 * after the same launch.  Call after vpint_solver_load / ingest on every rank, with the known-cell sums
 * (vpint_solver_known_ptr) already all-reduced.  launches_out (may be NULL): launch groups issued. */
size_t vpint_solver_peer_bytes(const vpint_solver *s, int world);
int    vpint_solver_peer_bind(vpint_solver *s, int world, int rank, void *const *peer_base);
int    vpint_solver_run_sharded(vpint_solver *s, const vpint_run_params *prm, void *stream, int64_t *launches_out);

/* ---- results -----------------------------------------------------------
 * out: device, dtype `dtype`, element strides as in load.  This is
 * self.pred_grid after run().  niter_out: device int32[nprob] sweeps executed
 * per problem (== len(delta_vec) under track_delta); may be NULL. */
int vpint_solver_result(vpint_solver *s, void *out, int dtype, const int64_t *stride,
                        int32_t *niter_out, void *stream);

/* Introspection used by benchmarks and tests (host values, valid after load). */
int64_t vpint_solver_active_tiles(const vpint_solver *s);
int64_t vpint_solver_total_tiles(const vpint_solver *s);
int64_t vpint_solver_launches(const vpint_solver *s);   /* kernels launched so far */
int     vpint_solver_k_block(const vpint_solver *s);

#ifdef __cplusplus
}
#endif
#endif /* VPINT_B200_H */
