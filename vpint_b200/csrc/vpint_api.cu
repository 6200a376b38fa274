// C ABI of the vpint_b200 library (declared in include/vpint_b200.h): solver object, workspace carving,
// and the loop drivers around the stencil / reduction / decision kernels.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "vpint_internal.cuh"

namespace vpint {
void k1_tile_shape(int ndim, int *th, int *tw);
void tb_tile_shape(int k_block, int *th, int *tw);
bool tb_supported(const Geom &g, int storage);
cudaError_t tb_make_maps(const Geom &g, int storage, void *buf0, void *buf1, void *maps_out);
size_t tb_maps_bytes();
cudaError_t launch_step2d_tb(const StepArgs &a, const void *maps, int storage, int weight_mode, int64_t n_tiles,
                             cudaStream_t st, int64_t *launches);
}  // namespace vpint

using namespace vpint;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

int fail_cuda(cudaError_t e, const char *where) {
    g_last_error = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return VPINT_ERR_CUDA;
}

#define VP_CUDA(call, where)                                   \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return fail_cuda(e__, where);  \
    } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Small pinned host blocks (the per-solver counters the loop drivers poll).  cudaHostAlloc / cudaFreeHost cost
// milliseconds and the free synchronises the whole device, so blocks are recycled through a process-wide
// free list and only returned to the driver at process exit.
constexpr size_t kPinnedBlock = 256;
std::mutex g_pinned_mutex;
std::vector<void *> g_pinned_free;

cudaError_t pinned_acquire(void **out) {
    {
        std::lock_guard<std::mutex> lock(g_pinned_mutex);
        if (!g_pinned_free.empty()) {
            *out = g_pinned_free.back();
            g_pinned_free.pop_back();
            return cudaSuccess;
        }
    }
    return cudaHostAlloc(out, kPinnedBlock, cudaHostAllocDefault);
}

void pinned_release(void *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_pinned_mutex);
    g_pinned_free.push_back(p);
}

enum Counter { kCntDirty = 0, kCntActiveTiles = 1, kCntTileless = 2, kCntActiveProblems = 8, kCntTotal = 16 };

}  // namespace

struct vpint_solver {
    vpint_desc desc;
    Geom g;
    int storage, weight_mode, n_planes;
    size_t elem;
    int64_t tiles_total;
    // workspace offsets
    size_t off_buf[2], off_mask, off_w[5], off_gamma, off_tau, off_fill, off_rowpart, off_known0, off_known,
        off_state, off_counters, off_tileflag, off_tiles, off_tilestart, off_partial, off_sums, ws_bytes;
    char *ws = nullptr;
    bool loaded = false, weights_ready = false, run_begun = false, first_step_pending = false;
    bool tiles_pending = false;       // fused load: tile list / counters not built yet (only the tiled kernels need them)
    bool bufB_pending = false;        // fused load: second ping-pong buffer not written yet (ditto)
    bool async_pending = false;       // vpint_solver_run_async launched the resident kernel; run_finish has not checked it
    bool use_tb = false;
    int halo_up = 0, halo_down = 0;   // halo rows above / below the owned band (row-sharded scenes)
    // cluster-resident solve (vpint_resident.cu): chosen when k_block == 0 and one problem fits on chip
    bool res_ok = false;              // plans valid for this solver's weight form
    ResidentPlan rplans[kResidentMaxC + 1];   // per cluster size (problems run on the smallest cluster that holds the rows
    int res_caps[kResidentMaxC + 1] = {0};    //   their sweeps touch); caps = rows a cluster of C CTAs holds, 0 = no plan
    int res_cmax = 0;                 // smallest cluster that holds ALL rows of a problem
    // row shards over peer-mapped memory (vpint_solver_peer_bind / vpint_solver_run_sharded)
    bool peer_bound = false;
    PeerCtx peer;
    size_t off_peerctr = 0;           // int[2]: lifetime launch counter of the sharded loop, first-launch-of-a-run flag
    unsigned long long *h_ring = nullptr, *d_ring = nullptr;   // mapped pinned ring: (launch << 32) | problems still running
    int peer_t = 0;                   // host mirror of the launch counter
    size_t off_partial8 = 0;          // per-warp partials of the ratio stencil (step API / sharded runs: no block barrier)
    size_t off_rowchunk = 0;          // fused load of long rows: per-chunk partials of the known-cell row sums
    size_t off_trial = 0;             // double[3][P]: per-problem beta, epsilon, mu of batched search trials
    bool trial_beta = false, trial_res = false;
    size_t off_cls = 0;               // int[P] class of every problem, int[32]: [0..8] problems per class, [16..24] class queues
    cudaStream_t cls_stream[kResidentMaxC + 1] = {nullptr};   // the launches of the classes run side by side: forked from the
    cudaEvent_t cls_fork = nullptr, cls_join[kResidentMaxC + 1] = {nullptr};   //   caller's stream and joined back into it
    bool planes_valid = false;        // the 4 (5) weight planes hold current weights
    bool fplane_valid = false;        // w[4] holds the single-layer feature plane (RATIO form)
    size_t off_flags = 0;             // int[16]: [0] non-finite feature seen, [1] problem queue, [2] sweeps left (graph loop), [3] a feature that is not a float32 value
    size_t off_order = 0;             // int[P]: queue order of the resident kernel
    size_t off_orig = 0;              // copy of the known values (VPINT_FEAT_RESISTANCE)
    size_t off_extreme = 0;           // double[2]: cells with non-finite / huge values, ... features
    size_t off_stage = 0;             // chunk sums of the two-stage per-problem reduction
    size_t off_done = 0;              // int[P] tickets of the fused finish
    bool can_fuse = false;
    int tileless = 0;                 // problems without any unknown cell (decided by the separate kernel)
    int64_t tiles_full = 0;           // active tiles of the full list built at load
    bool tiles_trimmed = false;       // the list currently covers only the problems that were still running
    bool has_orig = false;
    bool blocking_wait = false;       // VPINT_FEAT_BLOCKING_WAIT: s->ev sleeps, load() waits on it
    bool no_nc = false;               // weights carry the priority normalisation (vpint_solver_priority)
    // ratio form of the tiled path (jacobi2d_k1_ratio_kernel): state buffers hold P / f while a run is in flight
    bool state_ratio = false;         // buffers are in ratio space right now
    bool f32plane_valid = false;      // plane slot 1 holds the feature plane as float32 (ratio-direct load; slots 1-3 are free in ratio space)
    int f32_exact = -1;               // -1 unknown, 1 every feature is a float32 value (flags[3] == 0 at the range check), 0 not
    std::vector<cudaEvent_t> prof_ev;  // VPINT_SHARD_PROFILE=1: events around the kernels of every sharded sweep (diagnostics)
    bool prof_on = false;
    bool early_push = false;          // in-library sharded loop: halo rows leave right after the stencil, before the sums are reduced
    bool fill_ready = false;          // vpint_solver_fill_combine left the start values at off_fill; the next load may use them
    bool ratio_fresh = false;         // ... written so by the fused load, no sweep yet: plane slot 0 holds the exact start state
    int ratio_verdict = -1;           // -1 not evaluated since load/weights, 0 declined (extreme data), 1 allowed
    double *h_extreme = nullptr;      // pinned copy of the two range-check counters
    int64_t tiles_active = 0;
    int any_dirty = 0;
    int64_t launches = 0;
    int *h_counters = nullptr;   // pinned
    cudaEvent_t ev = nullptr;
    vpint_run_params prm;
    alignas(64) unsigned char tb_maps[512];   // two CUtensorMap (ping-pong state buffers)

    template <typename T>
    T *at(size_t off) const { return reinterpret_cast<T *>(ws + off); }
};

extern "C" {

int vpint_abi_version(void) { return VPINT_ABI_VERSION; }

const char *vpint_last_error(void) { return g_last_error.c_str(); }

int64_t vpint_row_pitch(int64_t row_elems) {
    if (row_elems < 1) row_elems = 1;
    return (row_elems + kPitchAlign - 1) / kPitchAlign * kPitchAlign;
}

int vpint_solver_create(const vpint_desc *d, vpint_solver **out) {
    if (!d || !out) return fail(VPINT_ERR_INVALID, "vpint_solver_create: null argument");
    *out = nullptr;
    if (d->ndim != 2 && d->ndim != 3) return fail(VPINT_ERR_INVALID, "ndim must be 2 or 3");
    if (d->nprob < 1 || d->H < 1 || d->W < 1 || d->T < 1) return fail(VPINT_ERR_INVALID, "nprob, H, W, T must be >= 1");
    if (d->ndim == 2 && d->T != 1) return fail(VPINT_ERR_INVALID, "T must be 1 for ndim == 2");
    if (d->storage != VPINT_F64 && d->storage != VPINT_F32) return fail(VPINT_ERR_INVALID, "bad storage type");
    if (d->weight_mode != VPINT_W_SCALAR && d->weight_mode != VPINT_W_PLANES)
        return fail(VPINT_ERR_INVALID, "bad weight_mode");
    if (d->k_block < 0 || d->k_block > kMaxKBlock) return fail(VPINT_ERR_INVALID, "k_block out of range");
    if (d->nprob_inner < 0 || (d->nprob_inner > 1 && d->nprob % d->nprob_inner != 0))
        return fail(VPINT_ERR_INVALID, "nprob_inner must divide nprob");
    if ((int64_t)d->W > (int64_t)1 << 30 || (int64_t)d->nprob * d->T > (int64_t)1 << 30)
        return fail(VPINT_ERR_INVALID, "row too long or too many slices");

    vpint_solver *s = new (std::nothrow) vpint_solver();
    if (!s) return fail(VPINT_ERR_INVALID, "out of host memory");
    s->desc = *d;
    Geom &g = s->g;
    g.ndim = d->ndim; g.P = d->nprob; g.H = d->H; g.W = d->W; g.T = d->T;
    g.Pin = d->nprob_inner > 1 ? d->nprob_inner : 1;
    g.S = (d->ndim == 3) ? d->T : 1;           // time-major slice stack (vpint_internal.cuh, Geom)
    g.L = d->W;
    g.pitch = (int)vpint_row_pitch(g.L);
    g.mpitch = g.pitch / kMaskWordBits;
    g.wpitch = g.pitch;
    if (d->global_H == 0 && d->own_rows == 0 && d->row_offset == 0 && d->own_row0 == 0) {
        g.global_H = g.H; g.row_offset = 0; g.own_row0 = 0; g.own_rows = g.H;
    } else {
        g.global_H = d->global_H; g.row_offset = d->row_offset; g.own_row0 = d->own_row0; g.own_rows = d->own_rows;
        if (g.own_row0 < 0 || g.own_rows < 1 || g.own_row0 + g.own_rows > g.H || g.row_offset < 0 ||
            g.row_offset + g.H > g.global_H) {
            delete s;
            return fail(VPINT_ERR_INVALID, "inconsistent row sharding (global_H, row_offset, own_row0, own_rows)");
        }
    }
    g.n_cells = (double)g.global_H * (double)g.W * (double)g.T;
    s->halo_up = g.own_row0;
    s->halo_down = g.H - g.own_row0 - g.own_rows;
    s->storage = d->storage;
    s->weight_mode = d->weight_mode;
    s->elem = (d->storage == VPINT_F64) ? 8 : 4;
    s->n_planes = (d->weight_mode == VPINT_W_PLANES) ? (d->ndim == 2 ? 4 : 5) : 0;
    const bool unsharded = (g.global_H == g.H && g.own_row0 == 0 && g.own_rows == g.H);
    if (d->k_block == 0 && d->ndim == 2 && d->storage == VPINT_F64 && unsharded && !getenv("VPINT_NO_RESIDENT")) {
        const int smem_optin = (233472 - kResidentCtasPerSm * 1024) / kResidentCtasPerSm;   // sm_100a: 228 KB per SM, 1 KB reserved per CTA
        for (int c = 1; c <= kResidentMaxC && s->res_cmax == 0; ++c) {
            if (!resident_plan(g, d->weight_mode == VPINT_W_PLANES, smem_optin, c, &s->rplans[c])) continue;
            s->res_caps[c] = c * s->rplans[c].R;
            if (s->res_caps[c] >= g.H) s->res_cmax = c;
        }
        s->res_ok = s->res_cmax > 0;
        if (!s->res_ok)
            for (int c = 0; c <= kResidentMaxC; ++c) s->res_caps[c] = 0;
    }
    if (d->ndim == 2 && d->weight_mode == VPINT_W_PLANES && d->storage == VPINT_F64) s->n_planes = 5;   // + the feature plane

    int kb = d->k_block;
    if (d->ndim == 3) {
        if (kb > 1) { delete s; return fail(VPINT_ERR_UNSUPPORTED, "temporal blocking (k_block > 1) is 2-D only"); }
        kb = 1;
    }
    if (kb == 0) kb = 1;
    g.KB = kb;
    s->use_tb = (kb > 1);
    if (s->use_tb) tb_tile_shape(kb, &g.TH, &g.TW);
    else k1_tile_shape(g.ndim, &g.TH, &g.TW);
    if (s->use_tb && !tb_supported(g, s->storage)) {
        delete s;
        return fail(VPINT_ERR_UNSUPPORTED, "temporally blocked kernel does not support this configuration");
    }
    {
        // a launch reads KB rows beyond the owned band unless that side is the physical border of the grid
        const bool top_border = (g.row_offset + g.own_row0 == 0);
        const bool bottom_border = (g.row_offset + g.own_row0 + g.own_rows == g.global_H);
        if ((!top_border && s->halo_up < g.KB) || (!bottom_border && s->halo_down < g.KB)) {
            delete s;
            return fail(VPINT_ERR_INVALID, "row sharding needs at least k_block halo rows on every interior side");
        }
    }
    g.tiles_y = (g.own_rows + g.TH - 1) / g.TH;
    g.tiles_x = (g.L + g.TW - 1) / g.TW;
    g.tiles_per_prob = g.tiles_y * g.tiles_x;
    s->tiles_total = (int64_t)g.P * g.S * g.tiles_per_prob;
    if (s->tiles_total >= ((int64_t)1 << 31) - 1024 || (int64_t)g.P * g.S * g.H >= ((int64_t)1 << 31) - 1) {
        delete s;
        return fail(VPINT_ERR_INVALID, "batch too large for one solver (tile or row count exceeds 2^31)");
    }

    // workspace layout
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t plane = (size_t)g.P * g.S * g.H * g.pitch * s->elem;
    s->off_buf[0] = take(plane);
    s->off_buf[1] = take(plane);
    s->off_mask = take((size_t)g.P * g.S * g.H * g.mpitch * sizeof(uint32_t));
    const size_t wplane = (size_t)g.P * g.H * g.pitch * s->elem;      // (H, W) per problem, also in 3-D
    for (int i = 0; i < 5; ++i) s->off_w[i] = (i < s->n_planes) ? take(wplane) : 0;
    s->off_gamma = take((size_t)g.P * sizeof(double));
    s->off_tau = take((size_t)g.P * sizeof(double));
    s->off_fill = take((size_t)g.P * sizeof(double));
    s->off_rowpart = take((size_t)g.P * g.S * g.H * 8 * sizeof(double));
    // known0, known, then two range-check counters (values, features): one contiguous block, all-reduced
    // as a whole by row-sharded runs
    s->off_known0 = take((2 * (size_t)g.P * kSumSlots + 2) * sizeof(double));
    s->off_known = s->off_known0 + (size_t)g.P * kSumSlots * sizeof(double);
    s->off_extreme = s->off_known0 + 2 * (size_t)g.P * kSumSlots * sizeof(double);
    s->off_state = take((size_t)g.P * sizeof(ProbState));
    s->off_counters = take(kCntTotal * sizeof(int));
    s->off_tileflag = take((size_t)s->tiles_total * sizeof(int));
    s->off_tiles = take((size_t)s->tiles_total * sizeof(TileRef));
    s->off_tilestart = take((size_t)(g.P * g.S + 1) * sizeof(int));
    s->off_partial = take((size_t)s->tiles_total * g.KB * kSumSlots * sizeof(double));
    if (g.ndim == 2 && g.KB == 1 && d->storage == VPINT_F64 && d->weight_mode == VPINT_W_PLANES && !getenv("VPINT_NO_WARP_PARTIALS"))
        s->off_partial8 = take((size_t)s->tiles_total * 8 * kSumSlots * sizeof(double));
    s->off_sums = take((size_t)g.P * g.KB * kSumSlots * sizeof(double));
    s->off_flags = take(16 * sizeof(int));
    s->off_stage = take(reduce_stage_bytes(g));
    s->off_done = take((size_t)g.P * sizeof(int));
    s->off_order = take((size_t)(kResidentMaxC + 1) * g.P * sizeof(int));
    s->off_cls = take(((size_t)g.P + 32) * sizeof(int));
    s->off_trial = take(3 * (size_t)g.P * sizeof(double));
    s->off_peerctr = take(4 * sizeof(int));
    {
        const int xc = (g.ndim == 2) ? ingest_row_chunks(g) : 1;
        s->off_rowchunk = (xc > 1) ? take((size_t)g.P * g.H * xc * 3 * sizeof(double)) : 0;
    }
    // one kernel per launch group unless a problem has so many tiles that its reduction wants many blocks
    // fused finish (the last tile block of a problem reduces and decides): saves two small launches per sweep, but the
    // copy-engine-fed ratio stencil (vpint_step_bulk.cu) only runs with a separate reduction and is ~7 ns per tile faster,
    // which outweighs the launches from ~2000 tiles on
    s->can_fuse = ((int64_t)g.S * g.tiles_per_prob < 8192) && !(s->off_partial8 && s->tiles_total >= 2048) &&
                  !getenv("VPINT_NO_FUSED_FINISH");
    s->has_orig = (d->features & VPINT_FEAT_RESISTANCE) != 0;
    s->blocking_wait = (d->features & VPINT_FEAT_BLOCKING_WAIT) != 0;
    if (s->has_orig) s->off_orig = take(plane);
    s->ws_bytes = off;
    *out = s;
    return VPINT_OK;
}

void vpint_solver_destroy(vpint_solver *s) {
    if (!s) return;
    pinned_release(s->h_counters);
    if (s->ev) cudaEventDestroy(s->ev);
    if (s->h_ring) cudaFreeHost(s->h_ring);
    if (s->cls_fork) cudaEventDestroy(s->cls_fork);
    for (int c = 0; c <= kResidentMaxC; ++c) {
        if (s->cls_join[c]) cudaEventDestroy(s->cls_join[c]);
        if (s->cls_stream[c]) cudaStreamDestroy(s->cls_stream[c]);
    }
    delete s;
}

size_t vpint_solver_workspace_bytes(const vpint_solver *s) { return s ? s->ws_bytes : 0; }

int vpint_solver_bind(vpint_solver *s, void *workspace, size_t bytes) {
    if (!s) return fail(VPINT_ERR_INVALID, "null solver");
    if (!workspace || bytes < s->ws_bytes) return fail(VPINT_ERR_WORKSPACE, "workspace missing or too small");
    if ((uintptr_t)workspace % 256 != 0) return fail(VPINT_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    if (!s->h_counters) {
        static_assert(kCntTotal * sizeof(int) <= 128 && 128 + 32 + sizeof(int) <= kPinnedBlock, "pinned block too small");
        void *blk = nullptr;
        VP_CUDA(pinned_acquire(&blk), "cudaHostAlloc");
        s->h_counters = static_cast<int *>(blk);
        s->h_extreme = reinterpret_cast<double *>(static_cast<char *>(blk) + 128);
        VP_CUDA(cudaEventCreateWithFlags(&s->ev, cudaEventDisableTiming | (s->blocking_wait ? cudaEventBlockingSync : 0)),
                "cudaEventCreate");
    }
    s->ws = static_cast<char *>(workspace);
    s->loaded = s->weights_ready = s->run_begun = false;
    s->planes_valid = s->fplane_valid = false;
    s->f32plane_valid = false;
    // bind has no stream argument: the memset runs on the legacy default stream and is waited for, so that it is
    // ordered before whatever the caller enqueues next on any (also non-blocking) stream
    VP_CUDA(cudaMemsetAsync(s->at<int>(s->off_flags), 0, 16 * sizeof(int), cudaStreamLegacy), "memset flags");
    VP_CUDA(cudaStreamSynchronize(cudaStreamLegacy), "bind sync");
    s->tiles_pending = s->bufB_pending = s->async_pending = false;
    if (s->use_tb) {
        if (tb_maps_bytes() > sizeof(s->tb_maps)) return fail(VPINT_ERR_INVALID, "internal: tensor map storage too small");
        VP_CUDA(tb_make_maps(s->g, s->storage, s->at<void>(s->off_buf[0]), s->at<void>(s->off_buf[1]), s->tb_maps),
                "cuTensorMapEncodeTiled");
    }
    return VPINT_OK;
}

// Active-tile list + the host copies of the load counters.  vpint_solver_load does this at once (it needs any_dirty);
// the fused load leaves it to the first tiled launch, because the cluster-resident kernel needs none of it.
static int resolve_tiles(vpint_solver *s, cudaStream_t st) {
    if (!s->tiles_pending) return VPINT_OK;
    const Geom &g = s->g;
    if (s->bufB_pending) {
        const size_t plane = (size_t)g.P * g.S * g.H * g.pitch * s->elem;
        VP_CUDA(cudaMemcpyAsync(s->at<void>(s->off_buf[1]), s->at<void>(s->off_buf[0]), plane, cudaMemcpyDeviceToDevice, st),
                "second state buffer");
        s->bufB_pending = false;
    }
    VP_CUDA(launch_tiles(g, s->at<uint32_t>(s->off_mask), s->at<int>(s->off_tileflag), s->at<TileRef>(s->off_tiles),
                         s->at<int>(s->off_tilestart), s->at<int>(s->off_counters), st, &s->launches),
            "tile list");
    VP_CUDA(cudaMemcpyAsync(s->h_counters, s->at<int>(s->off_counters), kCntTotal * sizeof(int), cudaMemcpyDeviceToHost, st),
            "read counters");
    if (s->blocking_wait) {
        VP_CUDA(cudaEventRecord(s->ev, st), "event record");
        VP_CUDA(cudaEventSynchronize(s->ev), "load sync");
    } else {
        VP_CUDA(cudaStreamSynchronize(st), "load sync");
    }
    s->any_dirty = s->h_counters[kCntDirty];
    s->tiles_active = s->h_counters[kCntActiveTiles];
    s->tiles_full = s->tiles_active;
    s->tiles_trimmed = false;
    s->tileless = s->h_counters[kCntTileless];
    s->tiles_pending = false;
    return VPINT_OK;
}

int vpint_solver_load(vpint_solver *s, const void *original, const void *pred0, const double *fill, int dtype,
                      const int64_t *stride, void *stream) {
    if (!s || !s->ws) return fail(VPINT_ERR_WORKSPACE, "solver has no workspace bound");
    if (!original || !stride) return fail(VPINT_ERR_INVALID, "vpint_solver_load: null argument");
    const bool fill_kept = (!pred0 && !fill && s->fill_ready);      // fill_leaves + fill_combine ran before this load
    s->fill_ready = false;
    const bool device_fill = (!pred0 && !fill && !fill_kept);
    if (device_fill && !fill_mean_supported(s->g))
        return fail(VPINT_ERR_UNSUPPORTED, "vpint_solver_load: grid too large for the single-kernel mean init; run vpint_solver_fill_leaves + vpint_solver_fill_combine first, or pass fill");
    if (device_fill && (s->g.global_H != s->g.H))
        return fail(VPINT_ERR_INVALID, "vpint_solver_load: a row shard cannot take the mean of the whole grid; pass fill");
    if (dtype != VPINT_F64 && dtype != VPINT_F32) return fail(VPINT_ERR_INVALID, "bad dtype");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Geom &g = s->g;
    if (device_fill)
        VP_CUDA(launch_fill_mean(g, original, dtype, stride, s->at<double>(s->off_fill), st, &s->launches), "fill mean");
    else if (!pred0 && !fill_kept)
        VP_CUDA(cudaMemcpyAsync(s->at<double>(s->off_fill), fill, (size_t)g.P * sizeof(double), cudaMemcpyHostToDevice, st),
                "copy fill");
    VP_CUDA(cudaMemsetAsync(s->at<int>(s->off_counters), 0, kCntTotal * sizeof(int), st), "memset counters");
    VP_CUDA(cudaMemsetAsync(s->at<double>(s->off_extreme), 0, sizeof(double), st), "memset range check");
    s->state_ratio = false;
    s->ratio_fresh = false;
    s->ratio_verdict = -1;
    s->f32plane_valid = false;
    VP_CUDA(launch_prepare(g, original, pred0, s->at<double>(s->off_fill), dtype, stride, s->storage,
                           s->at<void>(s->off_buf[0]), s->at<void>(s->off_buf[1]),
                           s->has_orig ? s->at<void>(s->off_orig) : nullptr, s->at<uint32_t>(s->off_mask),
                           s->at<double>(s->off_rowpart), s->at<double>(s->off_known0), s->at<double>(s->off_known),
                           s->at<double>(s->off_extreme), s->at<ProbState>(s->off_state), s->at<int>(s->off_counters), st, &s->launches),
            "prepare");
    s->tiles_pending = true;
    s->bufB_pending = false;
    s->async_pending = false;
    s->loaded = true;
    s->run_begun = false;
    return resolve_tiles(s, st);
}

static int leave_ratio(vpint_solver *s, cudaStream_t st);

int vpint_solver_weights(vpint_solver *s, const void *feat, int dtype, const int64_t *stride, int F, int method,
                         int clamp, double min_gamma, double max_gamma, void *stream) {
    if (!s || !s->ws) return fail(VPINT_ERR_WORKSPACE, "solver has no workspace bound");
    if (s->weight_mode != VPINT_W_PLANES) return fail(VPINT_ERR_INVALID, "solver was created with scalar weights");
    if (!feat || !stride || F < 1) return fail(VPINT_ERR_INVALID, "vpint_solver_weights: bad argument");
    if (dtype != VPINT_F64 && dtype != VPINT_F32) return fail(VPINT_ERR_INVALID, "bad dtype");
    if (method != VPINT_METHOD_EXACT && method != VPINT_METHOD_COSINE && method != VPINT_METHOD_EXACT_INVERSE)
        return fail(VPINT_ERR_INVALID, "bad weight method");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    void *planes[5];
    for (int i = 0; i < 5; ++i) planes[i] = (i < s->n_planes) ? s->at<void>(s->off_w[i]) : nullptr;
    {
        // the state may still be in ratio space (a tiled run of a single-layer solve): bring it back with the OLD
        // feature plane before the planes are rewritten
        const int rc = leave_ratio(s, st);
        if (rc != VPINT_OK) return rc;
    }
    s->planes_valid = s->fplane_valid = false;
    s->f32plane_valid = false;
    s->no_nc = false;
    s->ratio_verdict = -1;
    VP_CUDA(cudaMemsetAsync(s->at<double>(s->off_extreme) + 1, 0, sizeof(double), st), "memset range check");
    if (s->g.ndim == 2 && s->n_planes == 5 && F == 1 && method == VPINT_METHOD_EXACT && !clamp) {
        // one feature layer: keep the feature plane only; the ratio planes are derived on demand
        VP_CUDA(cudaMemsetAsync(s->at<int>(s->off_flags), 0, sizeof(int), st), "memset flag");
        VP_CUDA(cudaMemsetAsync(s->at<int>(s->off_flags) + 3, 0, sizeof(int), st), "memset flag");
        VP_CUDA(launch_fplane(s->g, feat, dtype, stride, s->at<double>(s->off_w[4]), s->at<int>(s->off_flags),
                              s->at<double>(s->off_extreme) + 1, st, &s->launches),
                "fplane");
        s->fplane_valid = true;
    } else if (s->g.ndim == 2) {
        VP_CUDA(launch_weights2d(s->g, feat, dtype, stride, F, method, clamp, min_gamma, max_gamma, s->storage, planes,
                                 st, &s->launches),
                "weights2d");
        s->planes_valid = true;
    } else {
        if (method == VPINT_METHOD_EXACT_INVERSE)
            return fail(VPINT_ERR_UNSUPPORTED, "exact_inverse is not a WP_STMRP method (VPint/WP_MRP.py:1443-1468)");
        VP_CUDA(launch_weights3d(s->g, feat, dtype, stride, F, method, s->storage, planes, st, &s->launches), "weights3d");
        s->planes_valid = true;
    }
    s->weights_ready = true;
    return VPINT_OK;
}

int vpint_solver_set_discounts(vpint_solver *s, const double *gamma, const double *tau, void *stream) {
    if (!s || !s->ws) return fail(VPINT_ERR_WORKSPACE, "solver has no workspace bound");
    if (s->weight_mode != VPINT_W_SCALAR) return fail(VPINT_ERR_INVALID, "solver was created with weight planes");
    if (!gamma || (s->g.ndim == 3 && !tau)) return fail(VPINT_ERR_INVALID, "vpint_solver_set_discounts: null argument");
    const size_t bytes = (size_t)s->g.P * sizeof(double);
    // pageable sources are staged before cudaMemcpyAsync returns; the copy itself is ordered on the caller's stream
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    VP_CUDA(cudaMemcpyAsync(s->at<double>(s->off_gamma), gamma, bytes, cudaMemcpyHostToDevice, st), "copy gamma");
    if (tau) VP_CUDA(cudaMemcpyAsync(s->at<double>(s->off_tau), tau, bytes, cudaMemcpyHostToDevice, st), "copy tau");
    s->weights_ready = true;
    return VPINT_OK;
}

int vpint_solver_set_trial_params(vpint_solver *s, const double *beta, const double *epsilon, const double *mu, void *stream) {
    if (!s || !s->ws) return fail(VPINT_ERR_WORKSPACE, "solver has no workspace bound");
    if ((epsilon == nullptr) != (mu == nullptr)) return fail(VPINT_ERR_INVALID, "epsilon and mu come together");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t bytes = (size_t)s->g.P * sizeof(double);
    double *base = s->at<double>(s->off_trial);
    if (beta) VP_CUDA(cudaMemcpyAsync(base, beta, bytes, cudaMemcpyHostToDevice, st), "copy beta");
    if (epsilon) {
        VP_CUDA(cudaMemcpyAsync(base + s->g.P, epsilon, bytes, cudaMemcpyHostToDevice, st), "copy epsilon");
        VP_CUDA(cudaMemcpyAsync(base + 2 * (size_t)s->g.P, mu, bytes, cudaMemcpyHostToDevice, st), "copy mu");
    }
    s->trial_beta = beta != nullptr;
    s->trial_res = epsilon != nullptr;
    return VPINT_OK;
}

static int recompact_tiles(vpint_solver *s, cudaStream_t st, bool trim);

// Start of a run(): parameter checks + per-problem loop state.  Nothing here needs the tile list.
static int begin_run_light(vpint_solver *s, const vpint_run_params *prm, cudaStream_t st) {
    if (!s || !s->ws) return fail(VPINT_ERR_WORKSPACE, "solver has no workspace bound");
    if (!prm) return fail(VPINT_ERR_INVALID, "null run params");
    if (!s->loaded) return fail(VPINT_ERR_INVALID, "vpint_solver_load has not been called");
    if (!s->weights_ready) return fail(VPINT_ERR_INVALID, "weights / discounts have not been set");
    if (prm->max_iter < 0) return fail(VPINT_ERR_INVALID, "max_iter must be >= 0");
    if (prm->resistance) {
        if (!s->has_orig) return fail(VPINT_ERR_INVALID, "resistance needs a solver created with VPINT_FEAT_RESISTANCE");
        if (s->g.ndim != 2 || s->use_tb) return fail(VPINT_ERR_UNSUPPORTED, "resistance runs on the 2-D one-sweep-per-launch path (k_block 0 or 1)");
    }
    s->prm = *prm;
    VP_CUDA(launch_init_state(s->g, s->at<ProbState>(s->off_state), s->at<int>(s->off_counters) + kCntActiveProblems,
                              prm->max_iter, st, &s->launches),
            "init state");
    s->run_begun = false;
    s->async_pending = false;
    return VPINT_OK;
}

// What the tiled kernels need on top of begin_run_light.
static int begin_run_tiled(vpint_solver *s, cudaStream_t st) {
    int rc = resolve_tiles(s, st);
    if (rc != VPINT_OK) return rc;
    if (s->tiles_trimmed) {
        rc = recompact_tiles(s, st, false);
        if (rc != VPINT_OK) return rc;
    }
    VP_CUDA(cudaMemsetAsync(s->at<int>(s->off_done), 0, (size_t)s->g.P * sizeof(int), st), "memset tickets");
    s->run_begun = true;
    s->first_step_pending = true;
    return VPINT_OK;
}

int vpint_solver_begin_run(vpint_solver *s, const vpint_run_params *prm, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rc = begin_run_light(s, prm, st);
    if (rc != VPINT_OK) return rc;
    return begin_run_tiled(s, st);
}

static int step_begin_impl(vpint_solver *s, void *stream, float *stencil_ms, bool fused = false);
static int step_end_impl(vpint_solver *s, void *stream, bool fused);

// The four ratio planes are built lazily from the single-layer feature plane (vpint_solver_weights keeps only
// that plane when the cluster-resident kernel can serve the run).
// Rebuild the active-tile list: for the running problems only (trim) or for every problem again (before a
// new run() on a solver whose list was trimmed).  Synchronises to learn the new launch size.
static int recompact_tiles(vpint_solver *s, cudaStream_t st, bool trim) {
    VP_CUDA(launch_recompact_tiles(s->g, s->at<int>(s->off_tileflag), s->at<TileRef>(s->off_tiles),
                                   s->at<int>(s->off_tilestart), s->at<int>(s->off_counters),
                                   trim ? s->at<ProbState>(s->off_state) : nullptr, st, &s->launches),
            "recompact tiles");
    VP_CUDA(cudaMemcpyAsync(s->h_counters + kCntActiveTiles, s->at<int>(s->off_counters) + kCntActiveTiles, sizeof(int),
                            cudaMemcpyDeviceToHost, st),
            "read tile count");
    VP_CUDA(cudaStreamSynchronize(st), "recompact sync");
    s->tiles_active = s->h_counters[kCntActiveTiles];
    s->tiles_trimmed = trim;
    return VPINT_OK;
}

// ---- ratio form of the tiled path -----------------------------------------------------------------------
static int leave_ratio(vpint_solver *s, cudaStream_t st) {
    if (!s->state_ratio) return VPINT_OK;
    if (s->ratio_fresh) {
        // nothing has run since the fused load wrote the ratio form: the saved plane IS the start state, bit for bit
        const size_t plane = (size_t)s->g.P * s->g.S * s->g.H * s->g.pitch * s->elem;
        VP_CUDA(cudaMemcpyAsync(s->at<void>(s->off_buf[0]), s->at<void>(s->off_w[0]), plane, cudaMemcpyDeviceToDevice, st), "restore state");
        VP_CUDA(cudaMemcpyAsync(s->at<void>(s->off_buf[1]), s->at<void>(s->off_w[0]), plane, cudaMemcpyDeviceToDevice, st), "restore state");
        s->state_ratio = false;
        s->ratio_fresh = false;
        return VPINT_OK;
    }
    void *bufs[2] = {s->at<void>(s->off_buf[0]), s->at<void>(s->off_buf[1])};
    VP_CUDA(launch_from_ratio(s->g, s->at<ProbState>(s->off_state), bufs, s->at<double>(s->off_w[4]),
                              s->at<double>(s->off_w[0]), s->at<uint32_t>(s->off_mask), st, &s->launches),
            "from ratio");
    s->state_ratio = false;
    return VPINT_OK;
}

// May this run use the ratio form?  One feature layer with exact weights, fp64, one sweep per launch, and
// every value / feature of moderate size (counters filled by load and weights; in a row-sharded run the
// caller has all-reduced them together with the known-cell sums, so every rank answers alike).
static int ratio_allowed(vpint_solver *s, cudaStream_t st, bool *out) {
    *out = false;
    if (s->weight_mode != VPINT_W_PLANES || s->g.ndim != 2 || s->storage != VPINT_F64 || s->use_tb || !s->fplane_valid ||
        s->no_nc || s->prm.resistance || getenv("VPINT_NO_RATIO_TILES"))
        return VPINT_OK;
    if (s->ratio_verdict < 0) {
        VP_CUDA(cudaMemcpyAsync(s->h_extreme, s->at<double>(s->off_extreme), 2 * sizeof(double), cudaMemcpyDeviceToHost, st),
                "read range check");
        int *h_wide = reinterpret_cast<int *>(reinterpret_cast<char *>(s->h_extreme) + 32);     // same pinned block
        VP_CUDA(cudaMemcpyAsync(h_wide, s->at<int>(s->off_flags) + 3, sizeof(int), cudaMemcpyDeviceToHost, st), "read float32 flag");
        VP_CUDA(cudaStreamSynchronize(st), "range check sync");
        s->ratio_verdict = (s->h_extreme[0] == 0.0 && s->h_extreme[1] == 0.0) ? 1 : 0;
        s->f32_exact = (*h_wide == 0) ? 1 : 0;
    }
    *out = (s->ratio_verdict == 1);
    return VPINT_OK;
}

static int ensure_planes(vpint_solver *s, cudaStream_t st) {
    {
        const int rc = leave_ratio(s, st);      // the plane slots double as the ratio form's save area
        if (rc != VPINT_OK) return rc;
    }
    if (s->weight_mode != VPINT_W_PLANES || s->planes_valid) return VPINT_OK;
    if (!s->fplane_valid) return fail(VPINT_ERR_INVALID, "weights have not been set");
    void *planes[5];
    for (int i = 0; i < 5; ++i) planes[i] = (i < s->n_planes) ? s->at<void>(s->off_w[i]) : nullptr;
    VP_CUDA(launch_planes_from_fplane(s->g, s->at<double>(s->off_w[4]), s->storage, planes, st, &s->launches),
            "planes from feature plane");
    s->planes_valid = true;
    s->f32plane_valid = false;        // slot 1 is a weight plane again
    return VPINT_OK;
}

int vpint_solver_priority(vpint_solver *s, double beta, const double *bias, double *p_out, void *stream) {
    if (!s || !s->ws) return fail(VPINT_ERR_WORKSPACE, "solver has no workspace bound");
    if (s->weight_mode != VPINT_W_PLANES || s->g.ndim != 2)
        return fail(VPINT_ERR_INVALID, "prioritise_identity applies to 2-D solvers with weight planes");
    if (!s->weights_ready) return fail(VPINT_ERR_INVALID, "vpint_solver_weights has not been called");
    if (s->no_nc) return fail(VPINT_ERR_INVALID, "priority planes are already folded into the weights");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rc = ensure_planes(s, st);
    if (rc != VPINT_OK) return rc;
    void *planes[5];
    for (int i = 0; i < 5; ++i) planes[i] = (i < s->n_planes) ? s->at<void>(s->off_w[i]) : nullptr;
    VP_CUDA(launch_priority(s->g, s->storage, planes, beta, s->trial_beta ? s->at<double>(s->off_trial) : nullptr, bias, p_out,
                            st, &s->launches),
            "priority");
    s->fplane_valid = false;          // the ratio form no longer describes these weights
    s->no_nc = true;
    return VPINT_OK;
}

int vpint_solver_step_begin(vpint_solver *s, void *stream) { return step_begin_impl(s, stream, nullptr, false); }

int vpint_solver_step_profile(vpint_solver *s, void *stream, float *stencil_ms) {
    if (!stencil_ms) return fail(VPINT_ERR_INVALID, "null stencil_ms");
    return step_begin_impl(s, stream, stencil_ms, false);
}

static int step_begin_impl(vpint_solver *s, void *stream, float *stencil_ms, bool fused) {
    if (!s || !s->run_begun) return fail(VPINT_ERR_INVALID, "vpint_solver_begin_run has not been called");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (stencil_ms) {
        VP_CUDA(cudaEventCreate(&e0), "event create");
        VP_CUDA(cudaEventCreate(&e1), "event create");
        VP_CUDA(cudaEventRecord(e0, st), "event record");
    }
    bool use_ratio = false;
    {
        int rc = ratio_allowed(s, st, &use_ratio);
        if (rc != VPINT_OK) return rc;
        if (use_ratio) {
            if (!s->state_ratio) {
                void *bufs[2] = {s->at<void>(s->off_buf[0]), s->at<void>(s->off_buf[1])};
                VP_CUDA(launch_to_ratio(s->g, bufs, s->at<double>(s->off_w[4]), s->at<double>(s->off_w[0]), st, &s->launches),
                        "to ratio");
                s->state_ratio = true;
                s->planes_valid = false;         // plane slot 0 now holds the saved known values
                s->f32plane_valid = false;       // (only the ratio-direct load writes the float32 copy)
            }
        } else {
            rc = ensure_planes(s, st);
            if (rc != VPINT_OK) return rc;
        }
    }
    StepArgs a;
    a.g = s->g;
    a.tiles = s->at<TileRef>(s->off_tiles);
    a.state = s->at<ProbState>(s->off_state);
    a.buf[0] = s->at<void>(s->off_buf[0]);
    a.buf[1] = s->at<void>(s->off_buf[1]);
    a.mask = s->at<uint32_t>(s->off_mask);
    for (int i = 0; i < 5; ++i) a.w[i] = (i < s->n_planes) ? s->at<void>(s->off_w[i]) : nullptr;
    a.f32plane = (use_ratio && s->f32plane_valid && s->f32_exact == 1 && !getenv("VPINT_NO_F32_PLANE")) ? s->at<float>(s->off_w[1]) : nullptr;
    a.gamma = s->at<double>(s->off_gamma);
    a.tau = s->at<double>(s->off_tau);
    a.partial = s->at<double>(s->off_partial);
    // the ratio stencil of a launch whose sums are reduced by a separate kernel writes per-warp partials (no barrier)
    static const bool ratio_v4 = !getenv("VPINT_RATIO_KERNEL") || atoi(getenv("VPINT_RATIO_KERNEL")) != 1;   // triage switch
    a.partial8 = (use_ratio && !fused && s->off_partial8 && ratio_v4) ? s->at<double>(s->off_partial8) : nullptr;
    a.clip = s->prm.clip_val;
    a.no_nc = s->no_nc ? 1 : 0;
    a.resist = s->prm.resistance ? 1 : 0;
    a.eps = s->prm.epsilon;
    a.mu = s->prm.mu;
    a.eps_arr = s->trial_res ? s->at<double>(s->off_trial) + s->g.P : nullptr;
    a.mu_arr = s->trial_res ? s->at<double>(s->off_trial) + 2 * (size_t)s->g.P : nullptr;
    a.orig = s->has_orig ? s->at<void>(s->off_orig) : nullptr;
    a.fused = fused ? 1 : 0;
    a.done = s->at<int>(s->off_done);
    a.tile_start = s->at<int>(s->off_tilestart);
    a.sums = s->at<double>(s->off_sums);
    a.known0 = s->at<double>(s->off_known0);
    a.known = s->at<double>(s->off_known);
    a.delta_out = s->prm.delta_out;
    a.n_active = s->at<int>(s->off_counters) + kCntActiveProblems;
    a.max_iter = s->prm.max_iter;
    a.auto_terminate = s->prm.auto_terminate;
    a.threshold = s->prm.threshold;
    s->ratio_fresh = false;              // (leave_ratio above has already used it if the ratio form is not taken)
    if (a.resist) {
        VP_CUDA(launch_step2d_resist(a, s->storage, s->weight_mode, st, &s->launches), "step2d_resist");
    } else if (use_ratio) {
        VP_CUDA(launch_step2d_ratio(a, s->tiles_active, st, &s->launches), "step2d_ratio");
    } else if (s->g.ndim == 2) {
        if (s->use_tb)
            VP_CUDA(launch_step2d_tb(a, s->tb_maps, s->storage, s->weight_mode, s->tiles_active, st, &s->launches), "step2d_tb");
        else
            VP_CUDA(launch_step2d(a, s->storage, s->weight_mode, s->tiles_active, st, &s->launches), "step2d");
    } else {
        VP_CUDA(launch_step3d(a, s->storage, s->weight_mode, s->tiles_active, st, &s->launches), "step3d");
    }
    if (s->prof_on) {
        cudaEvent_t ev;
        VP_CUDA(cudaEventCreate(&ev), "event create");
        VP_CUDA(cudaEventRecord(ev, st), "event record");
        s->prof_ev.push_back(ev);
    }
    if (s->early_push) {
        void *hb[2] = {s->at<void>(s->off_buf[0]), s->at<void>(s->off_buf[1])};
        VP_CUDA(launch_peer_halo(s->g, s->at<ProbState>(s->off_state), hb, s->peer, s->g.KB, s->at<int>(s->off_peerctr),
                                 s->at<int>(s->off_flags) + 4, 1, st, &s->launches),
                "peer halo push");
    }
    if (stencil_ms) {
        VP_CUDA(cudaEventRecord(e1, st), "event record");
        VP_CUDA(cudaEventSynchronize(e1), "event sync");
        VP_CUDA(cudaEventElapsedTime(stencil_ms, e0, e1), "event elapsed");
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    if (fused) return VPINT_OK;          // the last tile block of every problem has reduced and decided
    VP_CUDA(launch_reduce_partials(s->g, s->at<double>(s->off_partial), a.partial8, s->at<int>(s->off_tilestart),
                                   s->at<ProbState>(s->off_state), s->at<double>(s->off_sums), s->at<double>(s->off_stage),
                                   s->prm.resistance != 0, st, &s->launches),
            "reduce partials");
    return VPINT_OK;
}

int vpint_solver_step_end(vpint_solver *s, void *stream) { return step_end_impl(s, stream, false); }

static int step_end_impl(vpint_solver *s, void *stream, bool fused) {
    if (!s || !s->run_begun) return fail(VPINT_ERR_INVALID, "vpint_solver_begin_run has not been called");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!fused) {
    DecideArgs d;
    d.g = s->g;
    d.state = s->at<ProbState>(s->off_state);
    d.sums = s->at<double>(s->off_sums);
    d.known0 = s->at<double>(s->off_known0);
    d.known = s->at<double>(s->off_known);
    d.delta_out = s->prm.delta_out;
    d.n_active = s->at<int>(s->off_counters) + kCntActiveProblems;
    d.max_iter = s->prm.max_iter;
    d.auto_terminate = s->prm.auto_terminate;
    d.threshold = s->prm.threshold;
    d.all_cells = s->prm.resistance ? 1 : 0;
    VP_CUDA(launch_decide(d, st, &s->launches), "decide");
    }
    if (s->first_step_pending) {
        s->first_step_pending = false;
        if (s->any_dirty) {
            void *bufs[2] = {s->at<void>(s->off_buf[0]), s->at<void>(s->off_buf[1])};
            VP_CUDA(launch_restore_known(s->g, s->storage, s->at<ProbState>(s->off_state), bufs,
                                         s->at<uint32_t>(s->off_mask), st, &s->launches),
                    "restore known");
            s->any_dirty = 0;
        }
    }
    return VPINT_OK;
}

// Device-side loop: ONE graph launch runs sweeps until every problem has stopped.  The graph holds a
// conditional WHILE node whose body is one launch group (stencil, per-problem reduction, stop decision) and
// a one-thread kernel that re-arms the condition from the "problems still running" counter, so the host
// neither enqueues per sweep nor polls.
static int run_graph_loop(vpint_solver *s, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // first launch group outside the graph: it may be followed by the one-off known-cell repair
    const bool fused = s->can_fuse && s->tileless == 0;
    int rc = step_begin_impl(s, stream, nullptr, fused);
    if (rc != VPINT_OK) return rc;
    rc = step_end_impl(s, stream, fused);
    if (rc != VPINT_OK) return rc;
    int *iter_left = s->at<int>(s->off_flags) + 2;
    const int budget = 2 * s->prm.max_iter + 2;
    VP_CUDA(cudaMemcpyAsync(iter_left, &budget, sizeof(int), cudaMemcpyHostToDevice, st), "graph budget");

    cudaGraph_t graph = nullptr, body = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaStream_t cap = nullptr;
    cudaGraphConditionalHandle handle;
    int status = VPINT_OK;
    auto cleanup = [&]() {
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
        if (cap) cudaStreamDestroy(cap);
    };
#define VP_G(call, where)                                                   \
    do {                                                                    \
        cudaError_t e__ = (call);                                           \
        if (e__ != cudaSuccess) { cleanup(); return fail_cuda(e__, where); } \
    } while (0)
    VP_G(cudaGraphCreate(&graph, 0), "cudaGraphCreate");
    VP_G(cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault), "cudaGraphConditionalHandleCreate");
    cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = handle;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t node;
    VP_G(cudaGraphAddNode(&node, graph, nullptr, 0, &np), "cudaGraphAddNode(conditional)");
    body = np.conditional.phGraph_out[0];
    VP_G(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking), "cudaStreamCreate");
    VP_G(cudaStreamBeginCaptureToGraph(cap, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed), "begin capture");
    status = step_begin_impl(s, cap, nullptr, fused);
    if (status == VPINT_OK) status = step_end_impl(s, cap, fused);
    if (status == VPINT_OK) {
        cudaError_t e = launch_loop_condition(handle, s->at<int>(s->off_counters) + kCntActiveProblems, iter_left, cap,
                                              &s->launches);
        if (e != cudaSuccess) status = fail_cuda(e, "loop condition");
    }
    {
        cudaGraph_t captured = nullptr;
        cudaError_t e = cudaStreamEndCapture(cap, &captured);
        if (status == VPINT_OK && e != cudaSuccess) status = fail_cuda(e, "end capture");
    }
    if (status != VPINT_OK) { cleanup(); return status; }
    VP_G(cudaGraphInstantiate(&exec, graph, 0), "cudaGraphInstantiate");
    VP_G(cudaGraphLaunch(exec, st), "cudaGraphLaunch");
    VP_G(cudaMemcpyAsync(s->h_counters + kCntActiveProblems, s->at<int>(s->off_counters) + kCntActiveProblems,
                         sizeof(int), cudaMemcpyDeviceToHost, st),
         "read active count");
    VP_G(cudaEventRecord(s->ev, st), "event record");
    VP_G(cudaEventSynchronize(s->ev), "event sync");
#undef VP_G
    cleanup();
    if (s->h_counters[kCntActiveProblems] > 0)
        return fail(VPINT_ERR_INVALID, "internal error: graph loop did not terminate within its launch budget");
    return VPINT_OK;
}

int vpint_solver_run_async(vpint_solver *s, const vpint_run_params *prm, void *stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = begin_run_light(s, prm, st);
    if (rc != VPINT_OK) return rc;
    if (prm->loop_mode != VPINT_LOOP_CHUNKED && prm->loop_mode != VPINT_LOOP_GRAPH)
        return fail(VPINT_ERR_INVALID, "bad loop_mode");
    if (prm->max_iter == 0) return VPINT_OK;
    const bool ratio = (s->weight_mode == VPINT_W_PLANES);
    // any_dirty is known without a sync: the fused load has no separate start state, vpint_solver_load has read it
    if (s->res_ok && !s->any_dirty && !prm->resistance && !s->no_nc && (!ratio || s->fplane_valid)) {
        // whole run() in one launch: a cluster per problem, state resident in shared memory
        rc = leave_ratio(s, st);                   // a previous tiled run may have left the state in ratio space
        if (rc != VPINT_OK) return rc;
        int *flags = s->at<int>(s->off_flags);
        int *cls_of = s->at<int>(s->off_cls), *cls_count = cls_of + s->g.P, *cls_queue = cls_count + 16;
        VP_CUDA(cudaMemsetAsync(cls_count, 0, 32 * sizeof(int), st), "memset queues");
        ResidentArgs ra;
        ra.g = s->g;
        ra.state = s->at<ProbState>(s->off_state);
        ra.buf[0] = s->at<void>(s->off_buf[0]);
        ra.buf[1] = s->at<void>(s->off_buf[1]);
        ra.mask = s->at<uint32_t>(s->off_mask);
        ra.fplane = ratio ? s->at<double>(s->off_w[4]) : nullptr;
        ra.gamma = s->at<double>(s->off_gamma);
        ra.known0 = s->at<double>(s->off_known0);
        ra.known = s->at<double>(s->off_known);
        ra.delta_out = prm->delta_out;
        ra.n_active = s->at<int>(s->off_counters) + kCntActiveProblems;
        ra.bad_flag = ratio ? flags : nullptr;
        ra.f32_flag = ratio ? flags + 3 : nullptr;
        ra.max_iter = prm->max_iter;
        ra.auto_terminate = prm->auto_terminate;
        ra.threshold = prm->threshold;
        ra.clip = prm->clip_val;
        auto launch_class = [&](int c, const int *order, const int *count, cudaStream_t on) -> cudaError_t {
            const ResidentPlan &pl = s->rplans[c];
            ra.queue = cls_queue + c;
            ra.order = order;
            ra.count = count;
            ra.C = pl.C; ra.R = pl.R; ra.SP = pl.SP; ra.fcap = pl.fcap; ra.BS = pl.BS; ra.LB = pl.LB;
            return launch_resident2d(ra, ratio, pl.smem_bytes, 0, on, &s->launches);
        };
        if (s->g.P > 2 * 33 && s->g.P <= kResidentOrderMax && !getenv("VPINT_NO_QUEUE_ORDER")) {
            // A batch: every problem runs on the smallest cluster that holds the rows its sweeps touch (a class per
            // cluster size, one launch each, largest first), and inside a class the longest problems go first -- big
            // clouds need the most sweeps, so they must not start last.
            int *order = s->at<int>(s->off_order);
            const int c_top = getenv("VPINT_ONE_CLASS") ? 0 : s->res_cmax;
            int caps[kResidentMaxC + 1];
            for (int c = 0; c <= kResidentMaxC; ++c) caps[c] = (c_top && c <= c_top) ? s->res_caps[c] : 0;
            if (!c_top) caps[s->res_cmax] = s->res_caps[s->res_cmax];
            if (const char *m = getenv("VPINT_CLASS_MASK")) {        // triage: bit c set = cluster size c may be used
                const int mask = atoi(m);
                for (int c = 1; c < s->res_cmax; ++c)
                    if (!((mask >> c) & 1)) caps[c] = 0;
            }
            VP_CUDA(launch_resident_order(s->g, ra.state, s->at<double>(s->off_rowpart), caps, s->res_cmax, cls_of, order,
                                          cls_count, st, &s->launches),
                    "queue order");
            // Every class is launched on a stream of its own, forked from the caller's stream and joined back into it:
            // the classes run side by side (a cluster of 4 cannot use the last 2 SMs of a GPC, a cluster of 2 or 3 can;
            // the long problems of one class drain while another fills the machine).  The class streams have the LOWEST
            // priority, so that whatever else the caller has in flight on higher-priority streams -- the load / result
            // kernels of other chunks of a pipeline -- gets the next SM a persistent cluster leaves.
            const bool fork = !getenv("VPINT_CLASS_SERIAL");
            if (fork && !s->cls_fork) {
                int lo = 0, hi = 0;
                VP_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi), "priority range");
                VP_CUDA(cudaEventCreateWithFlags(&s->cls_fork, cudaEventDisableTiming), "cudaEventCreate");
                for (int c = 1; c <= s->res_cmax; ++c) {
                    if (caps[c] <= 0) continue;
                    // among the classes the larger clusters go first (they hold the largest clouds = the longest runs, and
                    // a cluster of 4 needs 4 free SMs of one GPC at once), but all stay below the default priority
                    int pr = lo - (c - 1);
                    if (pr < hi + 1) pr = hi + 1;
                    if (pr > lo) pr = lo;
                    VP_CUDA(cudaStreamCreateWithPriority(&s->cls_stream[c], cudaStreamNonBlocking, pr), "cudaStreamCreate");
                    VP_CUDA(cudaEventCreateWithFlags(&s->cls_join[c], cudaEventDisableTiming), "cudaEventCreate");
                }
            }
            if (fork) VP_CUDA(cudaEventRecord(s->cls_fork, st), "event record");
            for (int c = s->res_cmax; c >= 1; --c) {
                if (caps[c] <= 0) continue;
                cudaStream_t on = (fork && s->cls_stream[c]) ? s->cls_stream[c] : st;
                if (on != st) VP_CUDA(cudaStreamWaitEvent(on, s->cls_fork, 0), "stream wait");
                VP_CUDA(launch_class(c, order + (size_t)c * s->g.P, cls_count + c, on), "resident2d");
                if (on != st) VP_CUDA(cudaEventRecord(s->cls_join[c], on), "event record");
            }
            for (int c = 1; c <= s->res_cmax; ++c)
                if (fork && s->cls_stream[c] && caps[c] > 0) VP_CUDA(cudaStreamWaitEvent(st, s->cls_join[c], 0), "stream wait");
        } else {
            VP_CUDA(launch_class(s->res_cmax, nullptr, nullptr, st), "resident2d");
        }
        VP_CUDA(cudaMemcpyAsync(s->h_counters + kCntActiveProblems, s->at<int>(s->off_counters) + kCntActiveProblems,
                                sizeof(int), cudaMemcpyDeviceToHost, st),
                "read active count");
        VP_CUDA(cudaEventRecord(s->ev, st), "event record");
        s->async_pending = true;
    }
    return VPINT_OK;
}

int vpint_solver_run_finish(vpint_solver *s, void *stream, int *reran) {
    if (reran) *reran = 0;
    if (!s || !s->ws) return fail(VPINT_ERR_WORKSPACE, "solver has no workspace bound");
    if (!s->loaded) return fail(VPINT_ERR_INVALID, "vpint_solver_load has not been called");
    if (s->prm.max_iter == 0) return VPINT_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (s->async_pending) {
        s->async_pending = false;
        VP_CUDA(cudaEventSynchronize(s->ev), "event sync");
        if (s->h_counters[kCntActiveProblems] <= 0) return VPINT_OK;
        // the kernel declined (non-finite features): continue on the tiled path from the untouched state
    }
    if (reran) *reran = 1;
    int rc = begin_run_tiled(s, st);
    if (rc != VPINT_OK) return rc;
    const vpint_run_params *prm = &s->prm;
    if (prm->loop_mode == VPINT_LOOP_GRAPH) return run_graph_loop(s, stream);
    int chunk = prm->chunk > 0 ? prm->chunk : 4;
    const int chunk_cap = prm->chunk > 0 ? prm->chunk : 64;
    // Every launch either commits >= 1 sweep of every running problem or is its single replay, so
    // 2 * max_iter + 2 launches always suffice; the bound only guards against a logic error.
    int64_t budget = 2 * (int64_t)prm->max_iter + 2;
    int covered = s->g.P;                 // problems the current tile list covers
    while (budget > 0) {
        for (int i = 0; i < chunk; ++i) {
            rc = step_begin_impl(s, stream, nullptr, s->can_fuse && s->tileless == 0);
            if (rc != VPINT_OK) return rc;
            rc = step_end_impl(s, stream, s->can_fuse && s->tileless == 0);
            if (rc != VPINT_OK) return rc;
        }
        budget -= chunk;
        VP_CUDA(cudaMemcpyAsync(s->h_counters + kCntActiveProblems, s->at<int>(s->off_counters) + kCntActiveProblems,
                                sizeof(int), cudaMemcpyDeviceToHost, st),
                "read active count");
        VP_CUDA(cudaEventRecord(s->ev, st), "event record");
        VP_CUDA(cudaEventSynchronize(s->ev), "event sync");
        if (s->h_counters[kCntActiveProblems] <= 0) return VPINT_OK;
        if (s->g.P > 1 && !s->prm.resistance && s->h_counters[kCntActiveProblems] * 2 <= covered) {
            // many problems of the batch have stopped: later launches should not schedule their tiles
            rc = recompact_tiles(s, st, true);
            if (rc != VPINT_OK) return rc;
            covered = s->h_counters[kCntActiveProblems];
        }
        if (chunk < chunk_cap) chunk = (chunk * 2 < chunk_cap) ? chunk * 2 : chunk_cap;
    }
    return fail(VPINT_ERR_INVALID, "internal error: loop did not terminate within its launch budget");
}

int vpint_solver_run(vpint_solver *s, const vpint_run_params *prm, void *stream) {
    const int rc = vpint_solver_run_async(s, prm, stream);
    if (rc != VPINT_OK) return rc;
    return vpint_solver_run_finish(s, stream, nullptr);
}

// ---- fused VPint2 load (vpint_ingest.cu) ---------------------------------------------------------------------
int vpint_cloud_mask(const void *mask, int dtype, double threshold, int64_t n, uint8_t *cloud_out, void *stream) {
    if (!mask || !cloud_out || n < 0) return fail(VPINT_ERR_INVALID, "vpint_cloud_mask: bad argument");
    if (dtype != VPINT_F64 && dtype != VPINT_F32 && dtype != VPINT_U16 && dtype != VPINT_U8)
        return fail(VPINT_ERR_INVALID, "vpint_cloud_mask: bad dtype");
    VP_CUDA(launch_cloud_mask(mask, dtype, threshold, n, cloud_out, static_cast<cudaStream_t>(stream)), "cloud mask");
    return VPINT_OK;
}

int vpint_buffer_clouds(const void *target, int dtype, uint8_t *cloud, int have_cloud, int n_img, int H, int W, int B,
                        int buffer_size, int passes, uint8_t *scratch, void *stream) {
    if (!cloud || !scratch || n_img < 1 || H < 1 || W < 1 || B < 1) return fail(VPINT_ERR_INVALID, "vpint_buffer_clouds: bad argument");
    if (dtype != VPINT_F64 && dtype != VPINT_F32 && dtype != VPINT_U16) return fail(VPINT_ERR_INVALID, "vpint_buffer_clouds: bad dtype");
    if (dtype != VPINT_U16 && !target) return fail(VPINT_ERR_INVALID, "vpint_buffer_clouds: null target");
    VP_CUDA(launch_buffer_clouds(target, dtype, cloud, have_cloud != 0, n_img, H, W, B, buffer_size, passes, scratch,
                                 static_cast<cudaStream_t>(stream)),
            "buffer clouds");
    return VPINT_OK;
}

size_t vpint_fill_scratch_bytes(int64_t cells_global, int nprob, int value_dtype) {
    return fill_general_scratch_bytes(cells_global, nprob, value_dtype);
}

int64_t vpint_fill_leaf_count(int64_t cells_global, int nprob) { return (int64_t)nprob * ((cells_global + 63) / 64); }

size_t vpint_fill_count_offset(int64_t cells_global, int nprob, int value_dtype) {
    const size_t esz = (value_dtype == VPINT_F64) ? 8 : 4;
    return ((size_t)vpint_fill_leaf_count(cells_global, nprob) * esz + 255) / 256 * 256;
}

static int check_ingest(const vpint_solver *s, const vpint_ingest_params *p, const char *who, bool mean_only = false) {
    if (!s || !s->ws) return fail(VPINT_ERR_WORKSPACE, "solver has no workspace bound");
    if (!p || !p->target) return fail(VPINT_ERR_INVALID, std::string(who) + ": null argument");
    if (p->raster_dtype != VPINT_U16 && p->raster_dtype != VPINT_F32 && p->raster_dtype != VPINT_F64)
        return fail(VPINT_ERR_INVALID, std::string(who) + ": raster dtype must be VPINT_U16, VPINT_F32 or VPINT_F64");
    if (p->value_dtype != VPINT_F32 && p->value_dtype != VPINT_F64)
        return fail(VPINT_ERR_INVALID, std::string(who) + ": value dtype must be VPINT_F32 or VPINT_F64");
    if (mean_only) {                  // the mean init reads the raster only: any 2-D solver
        if (s->g.ndim != 2) return fail(VPINT_ERR_UNSUPPORTED, std::string(who) + ": needs a 2-D solver");
        return VPINT_OK;
    }
    if (s->g.ndim != 2 || s->storage != VPINT_F64 || s->weight_mode != VPINT_W_PLANES)
        return fail(VPINT_ERR_UNSUPPORTED, std::string(who) + ": needs a 2-D fp64 solver with weight planes");
    return VPINT_OK;
}

int vpint_solver_fill_leaves(vpint_solver *s, const vpint_ingest_params *p, void *stream) {
    int rc = check_ingest(s, p, "vpint_solver_fill_leaves", true);
    if (rc != VPINT_OK) return rc;
    if (!p->scratch) return fail(VPINT_ERR_INVALID, "vpint_solver_fill_leaves: null scratch");
    const Geom &g = s->g;
    const int64_t n = (int64_t)g.global_H * g.W;
    if (n < 128) return fail(VPINT_ERR_UNSUPPORTED, "vpint_solver_fill_leaves: grids below 128 cells take the mean through vpint_solver_load");
    const int n_img = g.P / g.Pin;
    const bool sharded = (g.global_H != g.H || g.own_row0 != 0 || g.own_rows != g.H);
    if (sharded && n_img != 1) return fail(VPINT_ERR_UNSUPPORTED, "vpint_solver_fill_leaves: a row shard holds one image");
    const int64_t cell_offset = (int64_t)g.row_offset * g.W, cells_local = (int64_t)g.H * g.W;
    const int64_t own0 = (int64_t)(g.row_offset + g.own_row0) * g.W, own1 = own0 + (int64_t)g.own_rows * g.W;
    if (own1 < n && own1 + 127 > cell_offset + cells_local)
        return fail(VPINT_ERR_INVALID, "vpint_solver_fill_leaves: the halo rows below the owned band must cover 127 cells");
    VP_CUDA(launch_fill_leaves(p->target, p->raster_dtype, p->value_dtype, p->cloud, n_img, g.Pin, n, cell_offset, cells_local,
                               own0, own1, p->clip_target, p->target_min, p->target_max, p->scratch,
                               static_cast<cudaStream_t>(stream), &s->launches),
            "fill leaves");
    return VPINT_OK;
}

int vpint_solver_fill_combine(vpint_solver *s, int value_dtype, void *scratch, void *stream) {
    if (!s || !s->ws) return fail(VPINT_ERR_WORKSPACE, "solver has no workspace bound");
    if (!scratch) return fail(VPINT_ERR_INVALID, "vpint_solver_fill_combine: null scratch");
    if (value_dtype != VPINT_F32 && value_dtype != VPINT_F64) return fail(VPINT_ERR_INVALID, "bad value dtype");
    const Geom &g = s->g;
    VP_CUDA(launch_fill_combine(value_dtype, g.P, (int64_t)g.global_H * g.W, scratch, s->at<double>(s->off_fill),
                                static_cast<cudaStream_t>(stream), &s->launches),
            "fill combine");
    s->fill_ready = true;
    return VPINT_OK;
}

double *vpint_solver_fill_ptr(vpint_solver *s) { return (s && s->ws) ? s->at<double>(s->off_fill) : nullptr; }

int vpint_solver_ingest_bands(vpint_solver *s, const vpint_ingest_params *p, void *stream) {
    int rc = check_ingest(s, p, "vpint_solver_ingest_bands");
    if (rc != VPINT_OK) return rc;
    if (!p->features) return fail(VPINT_ERR_INVALID, "vpint_solver_ingest_bands: null features");
    if (s->n_planes != 5) return fail(VPINT_ERR_UNSUPPORTED, "vpint_solver_ingest_bands: solver has no feature plane");
    if (!ingest_bands_supported(s->g, p->raster_dtype))
        return fail(VPINT_ERR_UNSUPPORTED, "vpint_solver_ingest_bands: rows too long for the staged load");
    if (p->out && p->out_dtype != VPINT_F32 && p->out_dtype != VPINT_F64 && p->out_dtype != VPINT_U16)
        return fail(VPINT_ERR_INVALID, "bad out dtype");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Geom &g = s->g;
    if (p->fill_mode == VPINT_FILL_DEVICE) {
        const bool sharded = (g.global_H != g.H || g.own_row0 != 0 || g.own_rows != g.H);
        if (sharded) return fail(VPINT_ERR_INVALID, "vpint_solver_ingest_bands: a row shard takes the mean through fill_leaves / all-reduce / fill_combine");
        rc = vpint_solver_fill_leaves(s, p, stream);
        if (rc != VPINT_OK) return rc;
        rc = vpint_solver_fill_combine(s, p->value_dtype, p->scratch, stream);
        if (rc != VPINT_OK) return rc;
    } else if (p->fill_mode == VPINT_FILL_HOST) {
        if (!p->fill_host) return fail(VPINT_ERR_INVALID, "vpint_solver_ingest_bands: null fill_host");
        VP_CUDA(cudaMemcpyAsync(s->at<double>(s->off_fill), p->fill_host, (size_t)g.P * sizeof(double), cudaMemcpyHostToDevice, st),
                "copy fill");
    } else if (p->fill_mode != VPINT_FILL_READY) {
        return fail(VPINT_ERR_INVALID, "vpint_solver_ingest_bands: bad fill_mode");
    }
    s->fill_ready = false;            // consumed here: a later vpint_solver_load takes its own mean
    VP_CUDA(cudaMemsetAsync(s->at<int>(s->off_counters), 0, kCntTotal * sizeof(int), st), "memset counters");
    VP_CUDA(cudaMemsetAsync(s->at<double>(s->off_extreme), 0, 2 * sizeof(double), st), "memset range check");
    VP_CUDA(cudaMemsetAsync(s->at<int>(s->off_flags), 0, sizeof(int), st), "memset flag");
    VP_CUDA(cudaMemsetAsync(s->at<int>(s->off_flags) + 3, 0, sizeof(int), st), "memset flag");
    s->state_ratio = false;
    s->ratio_fresh = false;
    s->ratio_verdict = -1;
    s->planes_valid = false;
    s->no_nc = false;
    // the second ping-pong buffer is only read by the tiled kernels: written lazily when the resident kernel will run
    const bool skip_b = s->res_ok && !s->has_orig;
    // a solve the tiled ratio stencil will run (one sweep per launch, no resident plan): state written in ratio space now
    const bool direct = !s->res_ok && !s->has_orig && !s->use_tb && s->g.KB == 1 && ingest_direct_supported(g, p->raster_dtype) &&
                        !getenv("VPINT_NO_RATIO_DIRECT") && !getenv("VPINT_NO_RATIO_TILES");
    VP_CUDA(launch_ingest_bands(g, p->target, p->features, p->raster_dtype, p->value_dtype, p->cloud, p->clip_target,
                                p->target_min, p->target_max, p->clip_features, p->features_min, p->features_max,
                                s->at<double>(s->off_fill), s->at<double>(s->off_buf[0]),
                                skip_b ? nullptr : s->at<double>(s->off_buf[1]), s->at<uint32_t>(s->off_mask),
                                s->at<double>(s->off_w[4]), s->at<double>(s->off_rowpart), s->at<double>(s->off_extreme),
                                s->at<int>(s->off_flags), p->out, p->out_dtype, p->out_stride,
                                s->off_rowchunk ? s->at<double>(s->off_rowchunk) : nullptr,
                                direct ? s->at<double>(s->off_w[0]) : nullptr, direct ? s->at<float>(s->off_w[1]) : nullptr, st,
                                &s->launches),
            "ingest bands");
    s->f32plane_valid = direct;
    s->f32_exact = -1;
    if (direct) {
        s->state_ratio = true;
        s->ratio_fresh = true;
    }
    if (s->has_orig)
        VP_CUDA(cudaMemcpyAsync(s->at<void>(s->off_orig), s->at<void>(s->off_buf[0]),
                                (size_t)g.P * g.S * g.H * g.pitch * s->elem, cudaMemcpyDeviceToDevice, st),
                "copy known values");
    VP_CUDA(launch_prepare_reduce(g, s->at<double>(s->off_rowpart), s->at<double>(s->off_known0), s->at<double>(s->off_known),
                                  s->at<ProbState>(s->off_state), s->at<int>(s->off_counters), st, &s->launches),
            "prepare reduce");
    s->fplane_valid = true;
    s->weights_ready = true;
    s->loaded = true;
    s->run_begun = false;
    s->any_dirty = 0;                 // no separate start state: it equals the original on every known cell
    s->tiles_pending = true;
    s->bufB_pending = skip_b;
    s->async_pending = false;
    return VPINT_OK;
}

int vpint_solver_result_unknown(vpint_solver *s, void *out, int dtype, const int64_t *stride, int32_t *niter_out,
                                void *stream) {
    if (!s || !s->ws) return fail(VPINT_ERR_WORKSPACE, "solver has no workspace bound");
    if (!s->loaded) return fail(VPINT_ERR_INVALID, "vpint_solver_load has not been called");
    if (!out || !stride) return fail(VPINT_ERR_INVALID, "vpint_solver_result_unknown: null argument");
    if (dtype != VPINT_F64 && dtype != VPINT_F32 && dtype != VPINT_U16) return fail(VPINT_ERR_INVALID, "bad dtype");
    if (s->g.ndim != 2 || s->storage != VPINT_F64) return fail(VPINT_ERR_UNSUPPORTED, "vpint_solver_result_unknown: 2-D fp64 solvers");
    {
        const int rc = leave_ratio(s, static_cast<cudaStream_t>(stream));
        if (rc != VPINT_OK) return rc;
    }
    int max_blocks = 0;
    {
        // pinned host memory is written through its device mapping (same address under unified addressing, but ask)
        cudaPointerAttributes attr;
        VP_CUDA(cudaPointerGetAttributes(&attr, out), "cudaPointerGetAttributes");
        if (attr.type == cudaMemoryTypeHost) {
            if (!attr.devicePointer) return fail(VPINT_ERR_INVALID, "vpint_solver_result_unknown: host memory is not mapped into the device");
            out = attr.devicePointer;
            // bound by the host link, not by SMs: a few blocks keep enough stores in flight and leave the SMs to the
            // solve kernels of the other streams
            max_blocks = getenv("VPINT_RESULT_BLOCKS") ? atoi(getenv("VPINT_RESULT_BLOCKS")) : 64;
        } else if (attr.type == cudaMemoryTypeUnregistered) {
            return fail(VPINT_ERR_INVALID, "vpint_solver_result_unknown: out must be device memory or pinned (mapped) host memory");
        }
    }
    void *bufs[2] = {s->at<void>(s->off_buf[0]), s->at<void>(s->off_buf[1])};
    VP_CUDA(launch_result_unknown(s->g, s->at<ProbState>(s->off_state), bufs, s->at<uint32_t>(s->off_mask), out, dtype,
                                  stride, niter_out, max_blocks, static_cast<cudaStream_t>(stream), &s->launches),
            "result unknown");
    return VPINT_OK;
}

double *vpint_solver_sums_ptr(vpint_solver *s) { return (s && s->ws) ? s->at<double>(s->off_sums) : nullptr; }

int vpint_solver_sums_count(const vpint_solver *s) { return s ? s->g.P * s->g.KB * kSumSlots : 0; }

void *vpint_solver_state_ptr(vpint_solver *s, int which) {
    return (s && s->ws && (which == 0 || which == 1)) ? s->at<void>(s->off_buf[which]) : nullptr;
}

int32_t *vpint_solver_active_ptr(vpint_solver *s) {
    return (s && s->ws) ? s->at<int32_t>(s->off_counters) + kCntActiveProblems : nullptr;
}

int32_t *vpint_solver_cur_ptr(vpint_solver *s) {
    // ProbState is 8 int32; `cur` is field 2.  Exposed as a strided view: element p at [p * 8].
    return (s && s->ws) ? reinterpret_cast<int32_t *>(s->at<ProbState>(s->off_state)) + 2 : nullptr;
}

double *vpint_solver_known_ptr(vpint_solver *s) { return (s && s->ws) ? s->at<double>(s->off_known0) : nullptr; }

int vpint_solver_known_count(const vpint_solver *s) { return s ? 2 * s->g.P * kSumSlots + 2 : 0; }

int vpint_solver_any_dirty(const vpint_solver *s) { return s ? s->any_dirty : 0; }

static int halo_impl(vpint_solver *s, int rows, void *up, void *down, bool pack, void *stream) {
    if (!s || !s->ws) return fail(VPINT_ERR_WORKSPACE, "solver has no workspace bound");
    if (!s->loaded) return fail(VPINT_ERR_INVALID, "vpint_solver_load has not been called");
    const Geom &g = s->g;
    if (rows < 1 || rows > g.own_rows) return fail(VPINT_ERR_INVALID, "halo rows out of range");
    if ((up && rows > s->halo_up) || (down && rows > s->halo_down))
        return fail(VPINT_ERR_INVALID, "halo message deeper than the halo rows of this shard");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    void *bufs[2] = {s->at<void>(s->off_buf[0]), s->at<void>(s->off_buf[1])};
    const ProbState *state = s->at<ProbState>(s->off_state);
    // pack: the owned rows a neighbour needs; unpack: the halo rows just outside the owned band
    const int up_row = pack ? g.own_row0 : g.own_row0 - rows;
    const int down_row = pack ? g.own_row0 + g.own_rows - rows : g.own_row0 + g.own_rows;
    if (up) VP_CUDA(launch_halo_rows(g, s->storage, state, bufs, up, up_row, rows, pack, st, &s->launches), "halo rows");
    if (down) VP_CUDA(launch_halo_rows(g, s->storage, state, bufs, down, down_row, rows, pack, st, &s->launches), "halo rows");
    return VPINT_OK;
}

int vpint_solver_halo_pack(vpint_solver *s, int rows, void *to_up, void *to_down, void *stream) {
    return halo_impl(s, rows, to_up, to_down, true, stream);
}

int vpint_solver_halo_unpack(vpint_solver *s, int rows, const void *from_up, const void *from_down, void *stream) {
    return halo_impl(s, rows, const_cast<void *>(from_up), const_cast<void *>(from_down), false, stream);
}

int vpint_solver_step_commit_lagged(vpint_solver *s, const double *sums_prev, void *stream) {
    if (!s || !s->run_begun) return fail(VPINT_ERR_INVALID, "vpint_solver_begin_run has not been called");
    if (s->g.KB != 1) return fail(VPINT_ERR_UNSUPPORTED, "the lagged stop decision needs one sweep per launch (k_block 1)");
    if (s->prm.delta_out || s->prm.resistance) return fail(VPINT_ERR_UNSUPPORTED, "the lagged stop decision does not record deltas / run resistance sweeps");
    if (s->any_dirty) return fail(VPINT_ERR_INVALID, "the lagged stop decision needs a start state that equals the original on known cells");
    DecideArgs d;
    d.g = s->g;
    d.state = s->at<ProbState>(s->off_state);
    d.sums = nullptr;
    d.known0 = s->at<double>(s->off_known0);
    d.known = s->at<double>(s->off_known);
    d.delta_out = nullptr;
    d.n_active = s->at<int>(s->off_counters) + kCntActiveProblems;
    d.max_iter = s->prm.max_iter;
    d.auto_terminate = s->prm.auto_terminate;
    d.threshold = s->prm.threshold;
    d.all_cells = 0;
    VP_CUDA(launch_lagged_commit(d, sums_prev, static_cast<cudaStream_t>(stream), &s->launches), "lagged commit");
    s->first_step_pending = false;
    return VPINT_OK;
}

int vpint_solver_halo_push(vpint_solver *s, int rows, void *peer_up, int32_t *flag_up, void *peer_down, int32_t *flag_down,
                           int seq, void *stream) {
    if (!s || !s->ws) return fail(VPINT_ERR_WORKSPACE, "solver has no workspace bound");
    if (!s->loaded) return fail(VPINT_ERR_INVALID, "vpint_solver_load has not been called");
    if (s->storage != VPINT_F64) return fail(VPINT_ERR_UNSUPPORTED, "peer halo exchange: fp64 state only");
    if (rows < 1 || rows > s->g.own_rows || (peer_up && !flag_up) || (peer_down && !flag_down))
        return fail(VPINT_ERR_INVALID, "vpint_solver_halo_push: bad argument");
    void *bufs[2] = {s->at<void>(s->off_buf[0]), s->at<void>(s->off_buf[1])};
    VP_CUDA(launch_halo_push(s->g, s->at<ProbState>(s->off_state), bufs, peer_up, peer_down, flag_up, flag_down, rows, seq,
                             s->at<int>(s->off_flags) + 4, static_cast<cudaStream_t>(stream), &s->launches),
            "halo push");
    return VPINT_OK;
}

int vpint_solver_halo_pull(vpint_solver *s, int rows, const void *from_up, const int32_t *flag_up, const void *from_down,
                           const int32_t *flag_down, int seq, void *stream) {
    if (!s || !s->ws) return fail(VPINT_ERR_WORKSPACE, "solver has no workspace bound");
    if (!s->loaded) return fail(VPINT_ERR_INVALID, "vpint_solver_load has not been called");
    if (s->storage != VPINT_F64) return fail(VPINT_ERR_UNSUPPORTED, "peer halo exchange: fp64 state only");
    if (rows < 1 || (from_up && (rows > s->halo_up || !flag_up)) || (from_down && (rows > s->halo_down || !flag_down)))
        return fail(VPINT_ERR_INVALID, "vpint_solver_halo_pull: bad argument");
    void *bufs[2] = {s->at<void>(s->off_buf[0]), s->at<void>(s->off_buf[1])};
    VP_CUDA(launch_halo_pull(s->g, s->at<ProbState>(s->off_state), bufs, from_up, from_down, flag_up, flag_down, rows, seq,
                             static_cast<cudaStream_t>(stream), &s->launches),
            "halo pull");
    return VPINT_OK;
}

// ---- row shards over peer-mapped memory: the whole loop in the library --------------------------------------------
static size_t peer_layout(const vpint_solver *s, int world, int rows, size_t *msg, size_t *off_gather, size_t *off_sflag) {
    const Geom &g = s->g;
    const size_t m = (size_t)g.P * rows * g.pitch * sizeof(double);
    const size_t og = align_up(4 * m + 256, 256);
    const size_t os = align_up(og + 2 * (size_t)world * g.P * kSumSlots * sizeof(double), 256);
    if (msg) *msg = m;
    if (off_gather) *off_gather = og;
    if (off_sflag) *off_sflag = os;
    return os + (size_t)world * 128;
}

size_t vpint_solver_peer_bytes(const vpint_solver *s, int world) {
    if (!s || world < 1 || world > kPeerMaxWorld) return 0;
    return peer_layout(s, world, s->g.KB, nullptr, nullptr, nullptr);
}

int vpint_solver_peer_bind(vpint_solver *s, int world, int rank, void *const *peer_base) {
    if (!s || !s->ws) return fail(VPINT_ERR_WORKSPACE, "solver has no workspace bound");
    if (world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world || !peer_base)
        return fail(VPINT_ERR_INVALID, "vpint_solver_peer_bind: bad argument (up to 16 ranks)");
    if (s->g.KB != 1 || s->storage != VPINT_F64 || s->g.ndim != 2)
        return fail(VPINT_ERR_UNSUPPORTED, "the peer-memory loop runs 2-D fp64 solvers with one sweep per launch (k_block 1)");
    PeerCtx &pc = s->peer;
    pc.world = world; pc.rank = rank;
    for (int r = 0; r < kPeerMaxWorld; ++r) pc.base[r] = (r < world) ? static_cast<char *>(peer_base[r]) : nullptr;
    for (int r = 0; r < world; ++r)
        if (!pc.base[r]) return fail(VPINT_ERR_INVALID, "vpint_solver_peer_bind: null peer block");
    peer_layout(s, world, s->g.KB, &pc.msg_bytes, &pc.off_gather, &pc.off_sflag);
    const Geom &g = s->g;
    const bool top = (g.row_offset + g.own_row0 == 0), bottom = (g.row_offset + g.own_row0 + g.own_rows == g.global_H);
    pc.up = top ? -1 : rank - 1;
    pc.down = bottom ? -1 : rank + 1;
    if (pc.up < -1 || pc.down >= world) return fail(VPINT_ERR_INVALID, "vpint_solver_peer_bind: rank order must follow the row order");
    if (!s->h_ring) {
        VP_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&s->h_ring), 64 * sizeof(unsigned long long), cudaHostAllocMapped),
                "cudaHostAlloc");
        memset(s->h_ring, 0, 64 * sizeof(unsigned long long));
        VP_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&s->d_ring), s->h_ring, 0), "cudaHostGetDevicePointer");
    }
    VP_CUDA(cudaMemsetAsync(s->at<int>(s->off_peerctr), 0, 4 * sizeof(int), cudaStreamLegacy), "memset launch counter");
    VP_CUDA(cudaStreamSynchronize(cudaStreamLegacy), "peer bind sync");
    s->peer_t = 0;
    s->peer_bound = true;
    return VPINT_OK;
}

int vpint_solver_run_sharded(vpint_solver *s, const vpint_run_params *prm, void *stream, int64_t *launches_out) {
    if (launches_out) *launches_out = 0;
    if (!s || !s->peer_bound) return fail(VPINT_ERR_INVALID, "vpint_solver_peer_bind has not been called");
    if (!prm) return fail(VPINT_ERR_INVALID, "null run params");
    if (prm->delta_out || prm->resistance) return fail(VPINT_ERR_UNSUPPORTED, "the peer-memory loop does not record deltas / run resistance sweeps");
    int rc = vpint_solver_begin_run(s, prm, stream);
    if (rc != VPINT_OK) return rc;
    if (s->any_dirty) return fail(VPINT_ERR_INVALID, "a row-sharded run needs a start state that equals the original on known cells");
    if (prm->max_iter == 0) return VPINT_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int *ctr = s->at<int>(s->off_peerctr);
    const int one = 1;
    VP_CUDA(cudaMemcpyAsync(ctr + 1, &one, sizeof(int), cudaMemcpyHostToDevice, st), "first-launch flag");
    DecideArgs d;
    d.g = s->g;
    d.state = s->at<ProbState>(s->off_state);
    d.sums = nullptr;
    d.known0 = s->at<double>(s->off_known0);
    d.known = s->at<double>(s->off_known);
    d.delta_out = nullptr;
    d.n_active = s->at<int>(s->off_counters) + kCntActiveProblems;
    d.max_iter = prm->max_iter;
    d.auto_terminate = prm->auto_terminate;
    d.threshold = prm->threshold;
    d.all_cells = 0;
    void *bufs[2] = {s->at<void>(s->off_buf[0]), s->at<void>(s->off_buf[1])};
    const int kDepth = 4;                          // launches the host stays ahead of the device
    const int64_t budget = 2 * (int64_t)prm->max_iter + 8;
    s->prof_on = getenv("VPINT_SHARD_PROFILE") != nullptr;
    s->prof_ev.clear();
    struct EarlyPush {                             // cleared on every way out of the loop
        vpint_solver *s;
        ~EarlyPush() { s->early_push = false; }
    } early_guard{s};
    s->early_push = !getenv("VPINT_HALO_LATE");
    int64_t n = 0;
    bool done = false;
    while (n < budget && !done) {
        auto stamp = [&]() {                                    // diagnostics only (VPINT_SHARD_PROFILE=1)
            if (!s->prof_on) return;
            cudaEvent_t ev;
            if (cudaEventCreate(&ev) != cudaSuccess) return;
            cudaEventRecord(ev, st);
            s->prof_ev.push_back(ev);
        };
        stamp();
        rc = step_begin_impl(s, stream, nullptr, false);          // sweep + local per-sweep sums (stamps after the stencil)
        if (rc != VPINT_OK) return rc;
        stamp();
        VP_CUDA(launch_peer_commit(d, s->at<double>(s->off_sums), s->peer, ctr, s->d_ring, st, &s->launches), "peer commit");
        stamp();
        VP_CUDA(launch_peer_halo(s->g, d.state, bufs, s->peer, s->g.KB, ctr, s->at<int>(s->off_flags) + 4, s->early_push ? 2 : 0, st,
                                 &s->launches),
                "peer halo");
        stamp();
        s->first_step_pending = false;
        ++n;
        ++s->peer_t;
        if (n > kDepth) {
            // the count the launch peer_t - kDepth left behind (every rank reads the same value: same decisions)
            const unsigned want = (unsigned)(s->peer_t - kDepth);
            volatile unsigned long long *slot = s->h_ring + ((want - 1u) & 63u);
            unsigned long long v;
            int spins = 0;
            while ((unsigned)((v = *slot) >> 32) != want) {
                if (++spins > 64) {
                    spins = 0;
                    if (cudaStreamQuery(st) == cudaSuccess && (unsigned)((v = *slot) >> 32) != want)
                        return fail(VPINT_ERR_INVALID, "internal error: sharded loop lost its launch record");
                }
            }
            if ((unsigned)(v & 0xffffffffu) == 0u) done = true;
        }
    }
    VP_CUDA(cudaStreamSynchronize(st), "sharded run sync");
    if (launches_out) *launches_out = n;
    if (s->prof_on) {
        // per sweep: [start, after stencil, after reduce, after commit, after halo]
        double acc[4] = {0, 0, 0, 0};
        const size_t per = 5, sweeps = s->prof_ev.size() / per;
        for (size_t i = 0; i < sweeps; ++i)
            for (int k = 0; k < 4; ++k) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, s->prof_ev[i * per + k], s->prof_ev[i * per + k + 1]);
                acc[k] += ms;
            }
        float total = 0.f;
        if (sweeps) cudaEventElapsedTime(&total, s->prof_ev[0], s->prof_ev[sweeps * per - 1]);
        fprintf(stderr, "[vpint shard profile] rows %d tiles %lld sweeps %zu: stencil %.3f reduce %.3f commit %.3f halo %.3f total %.3f ms\n",
                s->g.own_rows, (long long)s->tiles_active, sweeps, acc[0], acc[1], acc[2], acc[3], total);
        for (cudaEvent_t ev : s->prof_ev) cudaEventDestroy(ev);
        s->prof_ev.clear();
        s->prof_on = false;
    }
    {
        int timed_out = 0;
        VP_CUDA(cudaMemcpy(&timed_out, ctr + 2, sizeof(int), cudaMemcpyDeviceToHost), "read time-out flag");
        if (timed_out) {
            VP_CUDA(cudaMemset(ctr + 2, 0, sizeof(int)), "clear time-out flag");
            return fail(VPINT_ERR_CUDA, "vpint_solver_run_sharded: a rank did not publish its sums / halo rows within 10 s (peer died?)");
        }
    }
    if (!done) {
        int left = 0;
        VP_CUDA(cudaMemcpy(&left, d.n_active, sizeof(int), cudaMemcpyDeviceToHost), "read active count");
        if (left > 0) return fail(VPINT_ERR_INVALID, "internal error: sharded loop did not terminate within its launch budget");
    }
    return VPINT_OK;
}

int vpint_solver_result(vpint_solver *s, void *out, int dtype, const int64_t *stride, int32_t *niter_out, void *stream) {
    if (!s || !s->ws) return fail(VPINT_ERR_WORKSPACE, "solver has no workspace bound");
    if (!s->loaded) return fail(VPINT_ERR_INVALID, "vpint_solver_load has not been called");
    if (!out || !stride) return fail(VPINT_ERR_INVALID, "vpint_solver_result: null argument");
    if (dtype != VPINT_F64 && dtype != VPINT_F32) return fail(VPINT_ERR_INVALID, "bad dtype");
    void *bufs[2] = {s->at<void>(s->off_buf[0]), s->at<void>(s->off_buf[1])};
    if (s->state_ratio && !s->ratio_fresh && s->g.ndim == 2) {
        // gather straight out of ratio space (unknown cells f * r, known cells their saved value); the state stays there
        VP_CUDA(launch_result_ratio(s->g, s->at<ProbState>(s->off_state), bufs, s->at<double>(s->off_w[4]),
                                    s->at<double>(s->off_w[0]), s->at<uint32_t>(s->off_mask), out, dtype, stride, niter_out,
                                    static_cast<cudaStream_t>(stream), &s->launches),
                "result (ratio)");
        return VPINT_OK;
    }
    {
        const int rc = leave_ratio(s, static_cast<cudaStream_t>(stream));
        if (rc != VPINT_OK) return rc;
    }
    VP_CUDA(launch_result(s->g, s->storage, s->at<ProbState>(s->off_state), bufs, out, dtype, stride, niter_out,
                          static_cast<cudaStream_t>(stream), &s->launches),
            "result");
    return VPINT_OK;
}

int64_t vpint_solver_active_tiles(const vpint_solver *s) { return s ? s->tiles_active : 0; }
int64_t vpint_solver_total_tiles(const vpint_solver *s) { return s ? s->tiles_total : 0; }
int64_t vpint_solver_launches(const vpint_solver *s) { return s ? s->launches : 0; }
int vpint_solver_k_block(const vpint_solver *s) { return s ? s->g.KB : 0; }

}  // extern "C"
