// Internal declarations shared by the vpint_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/vpint_b200.h"

namespace vpint {

constexpr int kMaskWordBits = 32;
constexpr int kPitchAlign = 64;     // row pitch in elements: 64-bit mask words stay aligned, rows 512 B (f64) aligned
constexpr int kMaxKBlock = 16;      // per-launch fused sweeps (register-resident per-sweep sums)

// Sum slots of one sweep: {sum |new-old|, sum old, #NaN in |new-old|, #NaN in old}.
// Counts are kept as doubles so one buffer can be all-reduced as fp64.
constexpr int kSumSlots = 4;

// Geometry of a batched solve.  Device state is a stack of (H, pitch) SLICES: one per problem in 2-D, T per
// problem in 3-D (time-major: slice s = p * T + t), so that every slice is an image whose unknown cells are
// spatially coherent -- whatever the caller's memory order (the reference's cubes are (H, W, T) with T fastest).
struct Geom {
    int ndim, P, H, W, T;
    int Pin;          // inner problem count: p = outer * Pin + inner when addressing caller arrays
    int S;            // slices per problem: 1 (2-D) or T (3-D)
    int L;            // row length in elements = W
    int pitch;        // padded row length (multiple of kPitchAlign)
    int mpitch;       // 32-bit mask words per row = pitch / 32
    int wpitch;       // == pitch (kept for the (H, W) weight planes of the 3-D plane mode)
    int TH, TW;       // owned tile: rows x row-elements
    int tiles_y, tiles_x, tiles_per_prob;   // per SLICE
    int KB;           // sweeps per launch
    int global_H;     // rows of the whole grid (== H when unsharded)
    int row_offset;   // global row of local row 0
    int own_row0;     // first owned local row
    int own_rows;     // owned rows
    double n_cells;   // cells of the WHOLE grid (global_H * W * T): nanmean denominators
};

// Per-problem loop state (device).
struct ProbState {
    int iters_done;   // committed sweeps
    int k_run;        // sweeps the next launch executes for this problem (0 = idle)
    int cur;          // ping-pong buffer holding the current state
    int status;       // 0 running, 1 finished
    int replay;       // 1: next launch replays k_run sweeps after an overshoot, then finishes
    int dirty;        // 1: known cells of the start state differ from original (first launch runs 1 sweep)
    int fresh;        // 1: no sweep committed since load (first sweep uses the known0 sums)
    int n_unknown;    // unknown cells of the problem (set by load; orders the resident kernel's queue)
};

struct TileRef {
    int p, y0, x0, pad;      // p = SLICE index (problem * S + t)
};

// Everything a stencil launch needs.
struct StepArgs {
    Geom g;
    const TileRef *tiles;
    ProbState *state;
    void *buf[2];
    const uint32_t *mask;
    const void *w[5];        // planes: up, right, down, left (, temporal for 3-D)
    const float *f32plane;   // ratio form: the feature plane as float32 when every feature is exactly one (else null)
    const double *gamma;     // [P] scalar mode
    const double *tau;       // [P] scalar mode, 3-D
    double *partial;         // [tile][KB][kSumSlots]
    double *partial8;        // [tile][8 warps][kSumSlots]: per-WARP partials of the ratio kernel when the caller reduces in
                             // a separate launch (no block barrier in the stencil kernel); null = block partials
    double clip;
    int no_nc;               // 1: weights already carry the priority normalisation, no division by the neighbour count
    int resist;              // 1: resistance sweep over ALL cells (vpint_step_k1.cu, jacobi2d_resist_kernel)
    double eps, mu;
    const double *eps_arr, *mu_arr;   // [P] per-problem epsilon / mu (batched trials of a parameter search) or null
    const void *orig;        // [P][H][pitch] known values (resistance)
    // fused finish (vpint_solver_run on an unsharded solver): the last tile block of a problem reduces the
    // per-tile partials and takes the stop decision, so a launch group is ONE kernel
    int fused;
    int *done;               // [P] tickets of the blocks of this launch
    const int *tile_start;   // [P * S + 1]
    double *sums;            // [P][KB][kSumSlots]
    const double *known0, *known;
    double *delta_out;
    int *n_active;
    int max_iter, auto_terminate;
    double threshold;
};

struct DecideArgs {
    Geom g;
    ProbState *state;
    const double *sums;      // [P][KB][kSumSlots] (already combined over ranks when sharded)
    const double *known0;    // [P][kSumSlots] known-cell terms of the first sweep after load
    const double *known;     // [P][kSumSlots] known-cell terms of every later sweep
    double *delta_out;       // [P][max_iter] or null
    int *n_active;
    int max_iter;
    int auto_terminate;
    double threshold;
    int all_cells;           // 1: the sums already cover every cell (resistance), no known-cell constants
};

// ---- cluster-resident solve (vpint_resident.cu) -----------------------------
constexpr int kResidentThreads = 512;   // 15 worker warps + the sync warp
constexpr int kResidentCtasPerSm = 1;  // measured: 2 x 256 threads (clusters of 8) and 1 x 1024 threads are both slower than 1 x 512
constexpr int kResidentMaxSeg = 9;    // 4-cell segments per worker thread (new values kept in registers); 480 x 9 >= 64 rows x 64 segments
constexpr int kResidentMaxRows = 512; // grid rows of a problem the resident kernel takes (row map in static shared memory)
constexpr int kResidentMaxC = 8;      // portable cluster size

struct ResidentPlan {
    int C, R, SP, fcap;     // cluster size, rows per CTA, shared-memory row pitch, on-chip feature capacity
    int BS, LB;             // rows per round-robin block, blocks per CTA (R = BS * LB)
    size_t smem_bytes;      // dynamic shared memory per CTA
};

struct ResidentArgs {
    Geom g;
    ProbState *state;
    void *buf[2];
    const uint32_t *mask;
    const double *fplane;    // RATIO form: [P][H][pitch] features with zeros already replaced
    const double *gamma;     // SCALAR form: [P]
    const double *known0, *known;
    double *delta_out;       // [P][max_iter] or null
    int *n_active;
    int *queue;              // zeroed before the launch
    const int *order;        // queue position -> problem, largest clouds first; null = index order
    const int *count;        // problems in the queue (device; the launch of one class); null = all g.P problems
    const int *bad_flag;     // nonzero: decline (non-finite features); null = none
    const int *f32_flag;     // zero: every feature is exactly a float32 value; null = unknown (keep fp64 on chip)
    int max_iter, auto_terminate;
    double threshold, clip;
    int C, R, SP, fcap, BS, LB;
};

bool resident_plan(const Geom &g, bool ratio, int smem_limit, int C, ResidentPlan *out);
cudaError_t launch_resident2d(const ResidentArgs &a, bool ratio, size_t smem_bytes, int max_clusters, cudaStream_t st,
                              int64_t *launches);
constexpr int kResidentOrderMax = 32768;   // problems per solver up to which the queue is ordered (rank by counting)
cudaError_t launch_resident_order(const Geom &g, const ProbState *state, const double *rowpart, const int *caps, int c_max,
                                  int *cls_of, int *order, int *cls_count, cudaStream_t st, int64_t *launches);
cudaError_t launch_fplane(const Geom &g, const void *feat, int dtype, const int64_t *stride, double *fplane, int *bad_flag,
                          double *extreme, cudaStream_t st, int64_t *launches);
cudaError_t launch_planes_from_fplane(const Geom &g, const double *fplane, int storage, void *const *planes,
                                      cudaStream_t st, int64_t *launches);

// ---- host-side launchers (one per .cu) ------------------------------------
cudaError_t launch_prepare(const Geom &g, const void *original, const void *pred0, const double *fill_dev,
                           int dtype, const int64_t *stride, int storage, void *bufA, void *bufB, void *orig_copy,
                           uint32_t *mask, double *rowpart, double *known0, double *known, double *extreme,
                           ProbState *state, int *counters, cudaStream_t st, int64_t *launches);
bool fill_mean_supported(const Geom &g);
cudaError_t launch_fill_mean(const Geom &g, const void *original, int dtype, const int64_t *stride, double *fill_dev,
                             cudaStream_t st, int64_t *launches);
cudaError_t launch_tiles(const Geom &g, const uint32_t *mask, int *tile_flag, TileRef *tiles, int *tile_start,
                         int *counters, cudaStream_t st, int64_t *launches);
cudaError_t launch_recompact_tiles(const Geom &g, const int *tile_flag, TileRef *tiles, int *tile_start, int *counters,
                                   const ProbState *state, cudaStream_t st, int64_t *launches);
cudaError_t launch_weights2d(const Geom &g, const void *feat, int dtype, const int64_t *stride, int F, int method,
                             int clamp, double min_gamma, double max_gamma, int storage, void *const *planes,
                             cudaStream_t st, int64_t *launches);
cudaError_t launch_weights3d(const Geom &g, const void *feat, int dtype, const int64_t *stride, int F, int method,
                             int storage, void *const *planes, cudaStream_t st, int64_t *launches);
cudaError_t launch_step2d(const StepArgs &a, int storage, int weight_mode, int64_t n_tiles, cudaStream_t st,
                          int64_t *launches);
cudaError_t launch_step3d(const StepArgs &a, int storage, int weight_mode, int64_t n_tiles, cudaStream_t st,
                          int64_t *launches);
size_t reduce_stage_bytes(const Geom &g);
cudaError_t launch_reduce_partials(const Geom &g, const double *partial, const double *partial8, const int *tile_start,
                                   const ProbState *state, double *sums, double *stage, bool all_tiles,
                                   cudaStream_t st, int64_t *launches);
cudaError_t launch_step2d_resist(const StepArgs &a, int storage, int weight_mode, cudaStream_t st, int64_t *launches);
cudaError_t launch_priority(const Geom &g, int storage, void *const *planes, double beta, const double *beta_arr,
                            const double *bias, double *p_out, cudaStream_t st, int64_t *launches);
cudaError_t launch_init_state(const Geom &g, ProbState *state, int *n_active, int max_iter, cudaStream_t st,
                              int64_t *launches);
cudaError_t launch_decide(const DecideArgs &a, cudaStream_t st, int64_t *launches);
cudaError_t launch_step2d_ratio(const StepArgs &a, int64_t n_tiles, cudaStream_t st, int64_t *launches);
bool step2d_ratio_bulk_supported(const StepArgs &a);                      // vpint_step_bulk.cu
cudaError_t launch_step2d_ratio_bulk(const StepArgs &a, int64_t n_tiles, cudaStream_t st, int64_t *launches);
cudaError_t launch_to_ratio(const Geom &g, void *const *buf, const double *f, double *save, cudaStream_t st,
                            int64_t *launches);
cudaError_t launch_from_ratio(const Geom &g, const ProbState *state, void *const *buf, const double *f,
                              const double *save, const uint32_t *mask, cudaStream_t st, int64_t *launches);
cudaError_t launch_loop_condition(cudaGraphConditionalHandle handle, const int *n_active, int *iter_left,
                                  cudaStream_t st, int64_t *launches);
cudaError_t launch_restore_known(const Geom &g, int storage, ProbState *state, void *const *buf,
                                 const uint32_t *mask, cudaStream_t st, int64_t *launches);
cudaError_t launch_halo_rows(const Geom &g, int storage, const ProbState *state, void *const *buf, void *msg,
                             int first_row, int rows, bool pack, cudaStream_t st, int64_t *launches);
cudaError_t launch_result_ratio(const Geom &g, const ProbState *state, void *const *buf, const double *f, const double *save,
                                const uint32_t *mask, void *out, int dtype, const int64_t *stride, int32_t *niter_out,
                                cudaStream_t st, int64_t *launches);
cudaError_t launch_lagged_commit(const DecideArgs &a, const double *sums_prev, cudaStream_t st, int64_t *launches);
cudaError_t launch_halo_push(const Geom &g, const ProbState *state, void *const *buf, void *dst_up, void *dst_down,
                             int *flag_up, int *flag_down, int rows, int seq, int *tickets, cudaStream_t st, int64_t *launches);
cudaError_t launch_halo_pull(const Geom &g, const ProbState *state, void *const *buf, const void *src_up, const void *src_down,
                             const int *flag_up, const int *flag_down, int rows, int seq, cudaStream_t st, int64_t *launches);
cudaError_t launch_result(const Geom &g, int storage, const ProbState *state, void *const *buf, void *out,
                          int dtype, const int64_t *stride, int32_t *niter_out, cudaStream_t st,
                          int64_t *launches);

// ---- fused VPint2 load (vpint_ingest.cu) -----------------------------------
cudaError_t launch_cloud_mask(const void *mask, int dtype, double threshold, int64_t n, uint8_t *cloud, cudaStream_t st);
cudaError_t launch_buffer_clouds(const void *target, int dtype, uint8_t *cloud, bool have_cloud, int n_img, int H, int W, int B,
                                 int buffer_size, int passes, uint8_t *scratch, cudaStream_t st);
size_t fill_general_scratch_bytes(int64_t n, int n_prob, int acc_dtype);
cudaError_t launch_fill_leaves(const void *target, int t_dtype, int acc_dtype, const uint8_t *cloud, int n_img, int B, int64_t n,
                               int64_t cell_offset, int64_t cells_local, int64_t own_cell0, int64_t own_cell1, int clip,
                               double lo, double hi, void *scratch, cudaStream_t st, int64_t *launches);
cudaError_t launch_fill_combine(int acc_dtype, int n_prob, int64_t n, void *scratch, double *fill_dev, cudaStream_t st,
                                int64_t *launches);
bool ingest_bands_supported(const Geom &g, int t_dtype);
cudaError_t launch_ingest_bands(const Geom &g, const void *target, const void *features, int t_dtype, int acc_dtype,
                                const uint8_t *cloud, int clip_t, double t_lo, double t_hi, int clip_f, double f_lo,
                                double f_hi, const double *fill, double *bufA, double *bufB, uint32_t *mask, double *fplane,
                                double *rowpart, double *extreme, int *flags, void *out, int out_dtype,
                                const int64_t *out_stride, double *rowchunk, double *save, float *f32plane, cudaStream_t st,
                                int64_t *launches);
int ingest_row_chunks(const Geom &g);
bool ingest_direct_supported(const Geom &g, int t_dtype);
cudaError_t launch_result_unknown(const Geom &g, const ProbState *state, void *const *buf, const uint32_t *mask, void *out,
                                  int dtype, const int64_t *stride, int32_t *niter_out, int max_blocks, cudaStream_t st,
                                  int64_t *launches);
cudaError_t launch_prepare_reduce(const Geom &g, const double *rowpart, double *known0, double *known, ProbState *state,
                                  int *counters, cudaStream_t st, int64_t *launches);

// ---- row shards over peer-mapped memory (vpint_setup.cu) -----------------------
// Layout of every rank's peer block (symmetric memory; vpint_solver_peer_bytes):
//   [0, 4 msg)                       halo receive slots: [2 exchanges by parity][from above, from below][P][rows][pitch] fp64
//   [4 msg, +256)                    halo flags: int at +0 (from above), int at +128 (from below)
//   then [2 slots by parity][world][P * 4] doubles   per-sweep sums of every rank ("gather")
//   then [world] ints, 128 bytes apart               sequence number of the last sums rank r pushed
constexpr int kPeerMaxWorld = 16;
struct PeerCtx {
    int world, rank;
    char *base[kPeerMaxWorld];        // every rank's peer block, mapped into this process
    size_t off_gather, off_sflag, msg_bytes;
    int up, down;                     // row neighbours (-1 = none)
};
cudaError_t launch_peer_commit(const DecideArgs &a, const double *my_sums, const PeerCtx &pc, int *ctr,
                               unsigned long long *host_ring, cudaStream_t st, int64_t *launches);
cudaError_t launch_peer_halo(const Geom &g, const ProbState *state, void *const *buf, const PeerCtx &pc, int rows, const int *ctr,
                             int *tickets, int mode, cudaStream_t st, int64_t *launches);

// ---- device helpers -------------------------------------------------------
#if defined(__CUDACC__)

// Coalesced copy of `count` elements from global to (this warp's) shared memory: 16-byte vectors, several
// independent loads in flight per lane, when source and length allow; element-wise otherwise.
template <typename T>
__device__ __forceinline__ void warp_stage(T *dst, const T *__restrict__ src, int count, int lane) {
    constexpr int kPer = 16 / (int)sizeof(T);
    if ((count % kPer) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const int4 *s4 = reinterpret_cast<const int4 *>(src);
        int4 *d4 = reinterpret_cast<int4 *>(dst);
        const int n4 = count / kPer;
        int k = lane;
        for (; k + 7 * 32 < n4; k += 8 * 32) {
            int4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = s4[k + u * 32];
#pragma unroll
            for (int u = 0; u < 8; ++u) d4[k + u * 32] = v[u];
        }
        for (; k < n4; k += 32) d4[k] = s4[k];
    } else {
        for (int k = lane; k < count; k += 32) dst[k] = src[k];
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block reduction of NV doubles per thread; result valid in thread 0.
// `scratch` needs NV * (blockDim.x / 32) doubles.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) scratch[warp * NV + i] = v[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double acc = scratch[i];
            for (int w = 1; w < nwarp; ++w) acc += scratch[w * NV + i];
            v[i] = acc;
        }
    }
    __syncthreads();
}

// 2 consecutive elements, 128-bit for double / 64-bit for float.
__device__ __forceinline__ void load2(const double *p, double &a, double &b) {
    const double2 v = *reinterpret_cast<const double2 *>(p);
    a = v.x;
    b = v.y;
}
__device__ __forceinline__ void load2(const float *p, double &a, double &b) {
    const float2 v = *reinterpret_cast<const float2 *>(p);
    a = (double)v.x;
    b = (double)v.y;
}

// The stopping rule and the loop bookkeeping of ONE problem from its per-sweep sums S[KB][kSumSlots]:
//   delta = nanmean(|new - old|) / nanmean(old) over the whole grid (VPint/SD_MRP.py:139,
//   VPint/WP_MRP.py:352); known cells contribute |orig - orig| (0, or NaN for +-inf) and orig.
//   delta <= threshold commits the sweep and stops (SD_MRP.py:142-147); an overshoot inside a
//   temporally blocked launch schedules a replay of exactly `stop` sweeps from the untouched
//   launch-entry buffer.
__device__ __forceinline__ void decide_problem(const Geom &g, ProbState *state, int p, const double *sums_p,
                                               const double *known0, const double *known, double *delta_out,
                                               int *n_active, int max_iter, int auto_terminate, double threshold,
                                               int all_cells) {
    ProbState st = state[p];
    if (st.status != 0 || st.k_run == 0) return;
    const int kr = st.k_run;
    if (st.replay) {
        st.iters_done += kr;
        st.cur ^= 1;
        st.status = 1; st.k_run = 0; st.replay = 0; st.fresh = 0;
        state[p] = st;
        atomicSub(n_active, 1);
        return;
    }
    int stop = -1;
    for (int j = 0; j < kr; ++j) {
        const double *S = sums_p + (size_t)j * kSumSlots;
        const double *K = ((st.fresh && j == 0) ? known0 : known) + (size_t)p * kSumSlots;
        const double kz = all_cells ? 0.0 : 1.0;      // resistance: the sums already cover known cells
        const double s_abs = all_cells ? S[0] : S[0] + K[0], s_old = all_cells ? S[1] : S[1] + K[1];
        const double n_abs = g.n_cells - (S[2] + kz * K[2]), n_old = g.n_cells - (S[3] + kz * K[3]);
        const double delta = (s_abs / n_abs) / (s_old / n_old);
        if (delta_out) delta_out[(size_t)p * max_iter + st.iters_done + j] = delta;
        if (auto_terminate && delta <= threshold) {
            stop = j + 1;
            break;
        }
    }
    if (stop < 0 || stop == kr) {
        st.iters_done += kr;
        st.cur ^= 1;
        st.fresh = 0;
        if (stop == kr || st.iters_done >= max_iter) {
            st.status = 1;
            st.k_run = 0;
            atomicSub(n_active, 1);
        } else {
            st.k_run = min(g.KB, max_iter - st.iters_done);
        }
    } else {
        st.replay = 1;
        st.k_run = stop;
    }
    state[p] = st;
}

// Fused finish, called by every thread of a tile block after the block has written its partial sums: the
// block that takes the last ticket of its problem adds the partials of all the problem's tiles in a fixed
// order (threads stride over the tiles, then the deterministic block sum) and decides.  kr = sweeps of this
// launch.  `scratch` needs kSumSlots * (blockDim.x / 32) doubles.
__device__ __forceinline__ void finish_tile(const StepArgs &a, int p, int kr, double *scratch) {
    if (!a.fused) return;
    __shared__ int s_last;
    const Geom &g = a.g;
    __threadfence();
    __syncthreads();
    int t0, t1;
    if (a.resist) { t0 = p * g.S * g.tiles_per_prob; t1 = t0 + g.S * g.tiles_per_prob; }
    else { t0 = a.tile_start[p * g.S]; t1 = a.tile_start[(p + 1) * g.S]; }
    if (threadIdx.x == 0) {
        const int ticket = atomicAdd(&a.done[p], 1);
        s_last = (ticket == t1 - t0 - 1) ? 1 : 0;
        if (s_last) a.done[p] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double *out = a.sums + (size_t)p * g.KB * kSumSlots;
    for (int j = 0; j < g.KB; ++j) {
        double acc[kSumSlots] = {0, 0, 0, 0};
        if (j < kr) {
            for (int t = t0 + (int)threadIdx.x; t < t1; t += blockDim.x) {
                const double *q = a.partial + ((size_t)t * g.KB + j) * kSumSlots;
#pragma unroll
                for (int i = 0; i < kSumSlots; ++i) acc[i] += __ldcg(q + i);
            }
        }
        block_sum<kSumSlots>(acc, scratch);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < kSumSlots; ++i) out[j * kSumSlots + i] = acc[i];
        }
    }
    if (threadIdx.x == 0)
        decide_problem(g, a.state, p, out, a.known0, a.known, a.delta_out, a.n_active, a.max_iter, a.auto_terminate,
                       a.threshold, a.resist);
}

template <typename T>
__device__ __forceinline__ double ld_as_double(const void *base, int64_t idx) {
    return (double)reinterpret_cast<const T *>(base)[idx];
}

#endif  // __CUDACC__

}  // namespace vpint
