// Setup and bookkeeping kernels: state staging (A1), unknown-cell bit mask, active-tile list,
// per-problem loop state, result gather.  Reference semantics cited per kernel.
#include "vpint_internal.cuh"

namespace vpint {

namespace {

struct Strides {
    int64_t sO, sI, sY, sX, sT;
};

// Element offset of cell (row y, column x) of slice `sl` in the caller's array.
__device__ __forceinline__ int64_t cell_offset(const Geom &g, const Strides &s, int sl, int y, int x) {
    const int p = sl / g.S, t = sl - p * g.S;
    const int po = p / g.Pin, pi = p - po * g.Pin;
    return (int64_t)po * s.sO + (int64_t)pi * s.sI + (int64_t)y * s.sY + (int64_t)x * s.sX + (int64_t)t * s.sT;
}

// A1 on the device: init_strategy='mean' fills unknown cells with np.nanmean(grid) taken in the grid's own
// dtype (VPint/MRP.py:73-77).  NumPy evaluates it as  dtype( float64(pairwise_sum(grid with NaN -> 0)) /
// count )  where pairwise_sum is its blocked pairwise summation of the row-major copy (blocks of <= 128
// elements with 8 running accumulators, halves split at multiples of 8).  This kernel reproduces that
// summation tree bit for bit (checked against NumPy in tests/test_gpu_parity.py), one block per problem:
// thread 0 walks the recursion and lists the leaf blocks, 8 lanes per leaf own the 8 accumulators, thread 0
// combines the leaf sums in recursion order.
constexpr int kFillMaxLeaves = 2048;        // leaves hold 64..128 elements: grids up to 131072 cells

// Walks the row-major cell order of one 2-D problem without a division per element.
template <typename TIn>
struct FillCursor {
    const TIn *base;      // problem origin in the caller's array
    int64_t sY, sX;
    int L, y, x;
    __device__ __forceinline__ void seek(int64_t idx) { y = (int)(idx / L); x = (int)(idx - (int64_t)y * L); }
    __device__ __forceinline__ void advance(int n) {
        x += n;
        while (x >= L) { x -= L; ++y; }
    }
    __device__ __forceinline__ TIn get() const {
        const TIn v = base[(int64_t)y * sY + (int64_t)x * sX];
        return (v != v) ? (TIn)0 : v;
    }
};

template <typename TIn>
__global__ void __launch_bounds__(256) fill_mean_kernel(Geom g, const TIn *__restrict__ original, Strides sd,
                                                        double *__restrict__ fill) {
    __shared__ int leaf_lo[kFillMaxLeaves];
    __shared__ unsigned char leaf_sz[kFillMaxLeaves];
    __shared__ TIn leaf_sum[kFillMaxLeaves];
    __shared__ int n_leaves;
    __shared__ int cnt_part[8];
    const int p = blockIdx.x;
    const int64_t n = (int64_t)g.H * g.L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int po = p / g.Pin, pi = p - po * g.Pin;
    const TIn *base = original + (int64_t)po * sd.sO + (int64_t)pi * sd.sI;
    if (tid == 0) {                               // DFS over the recursion, left first
        int st_lo[40], st_n[40];
        int sp = 0, nl = 0;
        st_lo[0] = 0; st_n[0] = (int)n;
        while (sp >= 0) {
            const int lo = st_lo[sp], m = st_n[sp];
            --sp;
            if (m <= 128) {
                leaf_lo[nl] = lo; leaf_sz[nl] = (unsigned char)m; ++nl;
            } else {
                int h = m / 2;
                h -= h % 8;
                ++sp; st_lo[sp] = lo + h; st_n[sp] = m - h;     // right is popped after left
                ++sp; st_lo[sp] = lo; st_n[sp] = h;
            }
        }
        n_leaves = nl;
    }
    // count of non-NaN cells (order does not matter): rows x strided columns, no division per cell
    int cnt = 0;
    for (int y = 0; y < g.H; ++y) {
        const TIn *row = base + (int64_t)y * sd.sY;
        for (int x = tid; x < g.L; x += blockDim.x) {
            const TIn v = row[(int64_t)x * sd.sX];
            cnt += (v == v) ? 1 : 0;
        }
    }
    cnt = warp_sum(cnt);
    if (lane == 0) cnt_part[warp] = cnt;
    __syncthreads();
    // leaf sums: 8 lanes per leaf = the 8 accumulators of NumPy's unrolled loop
    const int sub = lane >> 3, j = lane & 7;
    for (int l0 = warp * 4; l0 < n_leaves; l0 += 8 * 4) {
        const int l = l0 + sub;
        TIn res = (TIn)0;
        if (l < n_leaves) {
            const int lo = leaf_lo[l], m = leaf_sz[l];
            FillCursor<TIn> cur{base, sd.sY, sd.sX, g.L, 0, 0};
            if (m < 8) {
                if (j == 0) {
                    cur.seek(lo);
                    for (int i = 0; i < m; ++i) { res += cur.get(); cur.advance(1); }
                }
            } else {
                cur.seek((int64_t)lo + j);
                TIn r = cur.get();
                int i = 8;
                for (; i < m - (m % 8); i += 8) { cur.advance(8); r += cur.get(); }
                // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) within each group of 8 lanes
                const unsigned grp = 0xffu << (sub * 8);
                TIn a = r + __shfl_down_sync(grp, r, 1, 8);     // valid in j = 0, 2, 4, 6
                TIn b = a + __shfl_down_sync(grp, a, 2, 8);     // valid in j = 0, 4
                TIn c = b + __shfl_down_sync(grp, b, 4, 8);     // valid in j = 0
                if (j == 0) {
                    res = c;
                    cur.seek((int64_t)lo + i);
                    for (; i < m; ++i) { res += cur.get(); cur.advance(1); }
                }
            }
            if (j == 0) leaf_sum[l] = res;
        }
    }
    __syncthreads();
    if (tid == 0) {
        // combine in recursion order: res(node) = res(left) + res(right)
        int st_n[40], st_state[40];
        TIn st_acc[40];
        int sp = 0, next_leaf = 0;
        st_n[0] = (int)n; st_state[0] = 0;
        TIn ret = (TIn)0;
        while (sp >= 0) {
            const int m = st_n[sp];
            if (m <= 128) { ret = leaf_sum[next_leaf++]; --sp; continue; }
            int h = m / 2;
            h -= h % 8;
            if (st_state[sp] == 0) { st_state[sp] = 1; ++sp; st_n[sp] = h; st_state[sp] = 0; }
            else if (st_state[sp] == 1) { st_acc[sp] = ret; st_state[sp] = 2; ++sp; st_n[sp] = m - h; st_state[sp] = 0; }
            else { ret = st_acc[sp] + ret; --sp; }
        }
        int total = 0;
        for (int w = 0; w < 8; ++w) total += cnt_part[w];
        const TIn total_sum = (TIn)0 + ret;                      // add.reduce starts from the identity
        fill[p] = (double)(TIn)((double)total_sum / (double)total);
    }
}

// The same mean for a band-interleaved stack (N, H, W, B) with B fastest -- VPint2's layout -- and H * W a power of
// two: one block per patch takes all B bands at once.  A leaf of the summation tree is 128 consecutive cells of a
// band, i.e. a contiguous run of 128 * B values; a warp stages that run with coalesced loads and then sums it per
// band the way NumPy does (8 lanes = the 8 accumulators).  With n = 2^k the tree above the leaves is balanced, so
// the leaf sums are combined level by level in parallel.  (The general kernel reads every band with a stride of B
// elements and walks the tree on one thread.)
constexpr int kFillBandsMax = 16;

template <typename TIn>
__global__ void __launch_bounds__(256) fill_mean_bands_kernel(Geom g, const TIn *__restrict__ original, int64_t sO,
                                                              double *__restrict__ fill) {
    extern __shared__ __align__(16) unsigned char fill_raw[];
    __shared__ int cnt_s[kFillBandsMax];
    const int B = g.Pin;
    const int n = g.H * g.L, NL = n / 128, run = 128 * B;
    TIn *stage = reinterpret_cast<TIn *>(fill_raw);            // [8 warps][128 * B]
    TIn *leaf = stage + 8 * run;                               // [B][NL]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sub = lane >> 3, j = lane & 7;
    const TIn *base = original + (int64_t)blockIdx.x * sO;
    if (tid < kFillBandsMax) cnt_s[tid] = 0;
    __syncthreads();
    int cnt[kFillBandsMax / 4] = {0, 0, 0, 0};
    TIn *my = stage + warp * run;
    for (int l = warp; l < NL; l += 8) {
        const TIn *src = base + (int64_t)l * run;
        warp_stage(my, src, run, lane);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < kFillBandsMax / 4; ++q) {
            if (q * 4 < B) {                                   // warp-uniform
                const int b = q * 4 + sub;
                const bool live = b < B;
                TIn r = (TIn)0;
                if (live) {
                    const TIn *col = my + j * B + b;
                    TIn v = col[0];
                    cnt[q] += (v == v) ? 1 : 0;
                    r = (v != v) ? (TIn)0 : v;
                    for (int i = 8; i < 128; i += 8) {
                        v = col[i * B];
                        cnt[q] += (v == v) ? 1 : 0;
                        r += (v != v) ? (TIn)0 : v;
                    }
                }
                // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) within each group of 8 lanes
                const TIn a = r + __shfl_down_sync(0xffffffffu, r, 1, 8);
                const TIn c = a + __shfl_down_sync(0xffffffffu, a, 2, 8);
                const TIn d = c + __shfl_down_sync(0xffffffffu, c, 4, 8);
                if (live && j == 0) leaf[b * NL + l] = d;
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int q = 0; q < kFillBandsMax / 4; ++q)
        if (q * 4 + sub < B && cnt[q]) atomicAdd(&cnt_s[q * 4 + sub], cnt[q]);
    __syncthreads();
    for (int st = 1; st < NL; st <<= 1) {                      // node = left + right, level by level
        const int pairs = NL / (2 * st);
        for (int idx = tid; idx < B * pairs; idx += blockDim.x) {
            const int b = idx / pairs, i = (idx - b * pairs) * 2 * st;
            leaf[b * NL + i] = leaf[b * NL + i] + leaf[b * NL + i + st];
        }
        __syncthreads();
    }
    if (tid < B) {
        const TIn total_sum = (TIn)0 + leaf[tid * NL];          // add.reduce starts from the identity
        fill[(size_t)blockIdx.x * B + tid] = (double)(TIn)((double)total_sum / (double)cnt_s[tid]);
    }
}

// One block per (problem, row).  Stages the start state of run():
//   bufA = pred0 (self.pred_grid, VPint/MRP.py:64-91 or whatever the caller left there),
//   bufB = original at known cells (the restore of VPint/SD_MRP.py:133-134 / WP_MRP.py:321,323 is
//          materialised once: later sweeps only write unknown cells),
//   mask = 1 bit per cell, set where original is NaN,
// and the known-cell terms of the stopping statistic (VPint/SD_MRP.py:139):
//   rowpart[8] = {sum|orig-pred0|, sum pred0, #NaN of the former, #NaN of the latter,
//                 sum orig, #inf orig, dirty (pred0 != orig on a known cell), #unknown}.
template <typename TIn, typename TSt>
__global__ void __launch_bounds__(256) prepare_kernel(Geom g, const TIn *__restrict__ original,
                                                      const TIn *__restrict__ pred0,
                                                      const double *__restrict__ fill, Strides sd,
                                                      TSt *__restrict__ bufA,
                                                      TSt *__restrict__ bufB, TSt *__restrict__ orig_copy,
                                                      uint32_t *__restrict__ mask,
                                                      double *__restrict__ rowpart, double *__restrict__ extreme) {
    __shared__ double scratch[8 * 8];
    bool wild = false;        // a non-finite or huge value: the ratio form of the tiled path must not be used
    const int p = blockIdx.x / g.H, y = blockIdx.x - p * g.H;      // p: slice
    const size_t row = ((size_t)p * g.H + y) * g.pitch;
    const size_t mrow = ((size_t)p * g.H + y) * g.mpitch;
    const bool owned = (y >= g.own_row0) && (y < g.own_row0 + g.own_rows);
    const double fillv = pred0 ? 0.0 : fill[p / g.S];
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int e = threadIdx.x; e < g.pitch; e += blockDim.x) {
        bool unknown = false;
        TSt va = (TSt)0, vb = (TSt)0;
        if (e < g.L) {
            const int64_t off = cell_offset(g, sd, p, y, e);
            const double o = (double)original[off];
            unknown = (o != o);
            const double v0 = pred0 ? (double)pred0[off] : (unknown ? fillv : o);
            va = (TSt)v0;
            vb = unknown ? va : (TSt)o;
            if (!(fabs(v0) < 1e150) || (!unknown && !(fabs(o) < 1e150))) wild = true;
            if (owned) {
                if (unknown) {
                    acc[7] += 1.0;
                } else {
                    const double os = (double)vb, vs = (double)va;
                    const double a0 = fabs(os - vs);
                    if (a0 != a0) acc[2] += 1.0; else acc[0] += a0;
                    if (vs != vs) acc[3] += 1.0; else acc[1] += vs;
                    acc[4] += os;
                    if (isinf(os)) acc[5] += 1.0;
                    if (!(vs == os)) acc[6] += 1.0;
                }
            }
        }
        bufA[row + e] = va;
        bufB[row + e] = vb;
        if (orig_copy) orig_copy[row + e] = vb;
        const unsigned word = __ballot_sync(0xffffffffu, unknown);   // pitch % 32 == 0: whole warps iterate together
        if ((threadIdx.x & 31) == 0) mask[mrow + (e >> 5)] = word;
    }
    if (__syncthreads_or(wild) && threadIdx.x == 0) atomicAdd(extreme, 1.0);
    block_sum<8>(acc, scratch);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) rowpart[(size_t)blockIdx.x * 8 + i] = acc[i];
    }
}

// prepare_kernel for a band-interleaved stack (B fastest, no pred0): one warp per (patch, row) stages the row of all
// B bands with coalesced loads and then writes every band's plane row.  The general kernel reads each band with a
// stride of B elements and pays a block-wide reduction of 8 values per 256 cells; here the known-cell terms shrink
// to {sum orig, #inf, #unknown} because pred0 equals orig on every known cell.
constexpr int kPrepWarps = 4;

template <typename TIn, typename TSt>
__global__ void __launch_bounds__(kPrepWarps * 32) prepare_bands_kernel(Geom g, const TIn *__restrict__ original,
                                                                        const double *__restrict__ fill, Strides sd,
                                                                        TSt *__restrict__ bufA, TSt *__restrict__ bufB,
                                                                        TSt *__restrict__ orig_copy, uint32_t *__restrict__ mask,
                                                                        double *__restrict__ rowpart, double *__restrict__ extreme) {
    extern __shared__ __align__(16) unsigned char prep_raw[];
    const int B = g.Pin, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * kPrepWarps + warp;          // (patch, row)
    const int64_t n_rows = (int64_t)(g.P / B) * g.H;
    if (r >= n_rows) return;
    const int po = (int)(r / g.H), y = (int)(r - (int64_t)po * g.H);
    TIn *my = reinterpret_cast<TIn *>(prep_raw) + (size_t)warp * g.L * B;
    const TIn *src = original + (int64_t)po * sd.sO + (int64_t)y * sd.sY;
    warp_stage(my, src, g.L * B, lane);
    __syncwarp();
    const bool owned = (y >= g.own_row0) && (y < g.own_row0 + g.own_rows);
    bool wild = false;
    for (int b = 0; b < B; ++b) {
        const int p = po * B + b;
        const size_t row = ((size_t)p * g.H + y) * g.pitch;
        const size_t mrow = ((size_t)p * g.H + y) * g.mpitch;
        const double fillv = fill[p];
        double sum_o = 0.0;
        int counts = 0;                                                 // #inf << 16 | #unknown
        for (int e = lane; e < g.pitch; e += 32) {
            bool unknown = false;
            TSt va = (TSt)0, vb = (TSt)0;
            if (e < g.L) {
                const double o = (double)my[e * B + b];
                unknown = (o != o);
                const double v0 = unknown ? fillv : o;
                va = (TSt)v0;
                vb = va;
                if (!(fabs(v0) < 1e150)) wild = true;
                if (owned) {
                    if (unknown) counts += 1;
                    else {
                        const double os = (double)vb;
                        sum_o += os;
                        if (isinf(os)) counts += 1 << 16;
                    }
                }
            }
            bufA[row + e] = va;
            bufB[row + e] = vb;
            if (orig_copy) orig_copy[row + e] = vb;
            const unsigned word = __ballot_sync(0xffffffffu, unknown);
            if (lane == 0) mask[mrow + (e >> 5)] = word;
        }
        sum_o = warp_sum(sum_o);
        counts = warp_sum(counts);
        if (lane == 0) {
            double *rp = rowpart + ((size_t)p * g.H + y) * 8;
            rp[0] = 0.0; rp[1] = sum_o; rp[2] = 0.0; rp[3] = 0.0;
            rp[4] = sum_o; rp[5] = (double)(counts >> 16); rp[6] = 0.0; rp[7] = (double)(counts & 0xffff);
        }
    }
    if (__any_sync(0xffffffffu, wild) && lane == 0) atomicAdd(extreme, 1.0);
}

// One block per problem: fixed-order reduction of the row partials.
__global__ void __launch_bounds__(256) prepare_reduce_kernel(Geom g, const double *__restrict__ rowpart,
                                                             double *__restrict__ known0,
                                                             double *__restrict__ known, ProbState *state,
                                                             int *counters) {
    __shared__ double scratch[8 * 8];
    const int p = blockIdx.x;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int y = threadIdx.x; y < g.S * g.H; y += blockDim.x) {      // the rows of every slice of the problem
        const double *r = rowpart + ((size_t)p * g.S * g.H + y) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += r[i];
    }
    block_sum<8>(acc, scratch);
    if (threadIdx.x == 0) {
        double *k0 = known0 + (size_t)p * kSumSlots, *k = known + (size_t)p * kSumSlots;
        k0[0] = acc[0]; k0[1] = acc[1]; k0[2] = acc[2]; k0[3] = acc[3];
        k[0] = 0.0; k[1] = acc[4]; k[2] = acc[5]; k[3] = 0.0;
        ProbState st;
        st.iters_done = 0; st.k_run = 0; st.cur = 0; st.status = 1; st.replay = 0;
        st.dirty = acc[6] > 0.0 ? 1 : 0; st.fresh = 1; st.n_unknown = (int)acc[7];
        state[p] = st;
        if (st.dirty) atomicAdd(&counters[0], 1);
        if (acc[7] == 0.0) atomicAdd(&counters[2], 1);     // a problem without unknown cells has no tile to finish it
    }
}

// One warp per tile: does the owned tile contain an unknown cell?
__global__ void __launch_bounds__(256) tile_flag_kernel(Geom g, const uint32_t *__restrict__ mask,
                                                        int *__restrict__ tile_flag, int64_t n_tiles) {
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= n_tiles) return;
    const int lane = threadIdx.x & 31;
    const int p = (int)(t / g.tiles_per_prob);
    const int r = (int)(t - (int64_t)p * g.tiles_per_prob);
    const int ty = r / g.tiles_x, tx = r - ty * g.tiles_x;
    const int y0 = g.own_row0 + ty * g.TH, y1 = min(y0 + g.TH, g.own_row0 + g.own_rows);
    const int x0 = tx * g.TW, x1 = min(x0 + g.TW, g.L);
    unsigned any = 0;
    if (x1 > x0) {
        const int w0 = x0 >> 5, w1 = (x1 - 1) >> 5;
        for (int y = y0 + lane; y < y1; y += 32) {
            const uint32_t *mr = mask + ((size_t)p * g.H + y) * g.mpitch;
            for (int w = w0; w <= w1; ++w) {
                unsigned bits = mr[w];
                if (w == w0) bits &= 0xffffffffu << (x0 & 31);
                if (w == w1 && ((x1 & 31) != 0)) bits &= 0xffffffffu >> (32 - (x1 & 31));
                any |= bits;
            }
        }
    }
    const unsigned vote = __ballot_sync(0xffffffffu, any != 0);
    if (lane == 0) tile_flag[t] = vote ? 1 : 0;
}

// Single block: ordered compaction of the flagged tiles (problem-major, then row, then column),
// tile_start[p] = index of problem p's first active tile, counters[1] = number of active tiles.
// With `state` the list is rebuilt for the problems that are still running only (the loop driver does that
// when many problems of a batch have stopped, so that later launches do not schedule their tiles).
__global__ void __launch_bounds__(1024) tile_compact_kernel(Geom g, const int *__restrict__ tile_flag,
                                                            TileRef *__restrict__ tiles,
                                                            int *__restrict__ tile_start, int *counters,
                                                            int64_t n_tiles, const ProbState *__restrict__ state) {
    __shared__ int counts[1024];
    const int tid = threadIdx.x;
    const int64_t chunk = (n_tiles + 1023) / 1024;
    const int64_t a = min((int64_t)tid * chunk, n_tiles), b = min(a + chunk, n_tiles);
    auto live = [&](int64_t i) {
        if (!tile_flag[i]) return 0;
        if (!state) return 1;
        return state[(int)(i / g.tiles_per_prob) / g.S].status == 0 ? 1 : 0;
    };
    int c = 0;
    for (int64_t i = a; i < b; ++i) c += live(i);
    counts[tid] = c;
    __syncthreads();
    // inclusive Hillis-Steele scan over 1024 counts
    for (int off = 1; off < 1024; off <<= 1) {
        const int v = (tid >= off) ? counts[tid - off] : 0;
        __syncthreads();
        counts[tid] += v;
        __syncthreads();
    }
    int pos = counts[tid] - c;
    for (int64_t i = a; i < b; ++i) {
        const int p = (int)(i / g.tiles_per_prob);
        const int r = (int)(i - (int64_t)p * g.tiles_per_prob);
        if (r == 0) tile_start[p] = pos;
        if (live(i)) {
            const int ty = r / g.tiles_x, tx = r - ty * g.tiles_x;
            TileRef tr;
            tr.p = p; tr.y0 = g.own_row0 + ty * g.TH; tr.x0 = tx * g.TW; tr.pad = 0;
            tiles[pos++] = tr;
        }
    }
    if (tid == 1023) {
        tile_start[g.P * g.S] = counts[1023];
        counters[1] = counts[1023];
    }
}

// Start of a run(): `iterations` bookkeeping of VPint/SD_MRP.py:63-66 / WP_MRP.py:112-115 is done by the
// caller (max_iter, auto_terminate); iterations == 0 returns the start state untouched.
__global__ void init_state_kernel(Geom g, ProbState *state, int *n_active, int max_iter) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.P) return;
    ProbState st = state[p];
    st.iters_done = 0;
    st.replay = 0;
    if (max_iter <= 0) {
        st.status = 1;
        st.k_run = 0;
    } else {
        st.status = 0;
        st.k_run = st.dirty ? 1 : min(g.KB, max_iter);
        atomicAdd(n_active, 1);
    }
    state[p] = st;
}

// Sum slot i of tile t (sweep j).  With per-warp partials the tile's value is formed first, the 8 warps added in order
// -- exactly what the block reduction of a stencil kernel does -- so both forms give the same bits.
__device__ __forceinline__ double tile_partial(const double *__restrict__ partial, const double *__restrict__ partial8, int KB,
                                               int t, int j, int i) {
    if (!partial8) return partial[((size_t)t * KB + j) * kSumSlots + i];
    const double *q = partial8 + (size_t)t * 8 * kSumSlots + i;
    double acc = q[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) acc += q[w * kSumSlots];
    return acc;
}

// Block per problem: fixed-order sum of the per-tile partials of the sweeps this launch executed.
__global__ void __launch_bounds__(128) reduce_partials_kernel(Geom g, const double *__restrict__ partial,
                                                              const double *__restrict__ partial8,
                                                              const int *__restrict__ tile_start,
                                                              const ProbState *__restrict__ state,
                                                              double *__restrict__ sums, int all_tiles) {
    __shared__ double scratch[kSumSlots * 4];
    const int p = blockIdx.x;
    const int kr = state[p].k_run;
    double *out = sums + (size_t)p * g.KB * kSumSlots;
    if (kr == 0) {
        // keep the buffer finite so an external all-reduce never mixes stale values in
        for (int i = threadIdx.x; i < g.KB * kSumSlots; i += blockDim.x) out[i] = 0.0;
        return;
    }
    const int t0 = all_tiles ? p * g.S * g.tiles_per_prob : tile_start[p * g.S];
    const int t1 = all_tiles ? t0 + g.S * g.tiles_per_prob : tile_start[(p + 1) * g.S];
    for (int j = 0; j < g.KB; ++j) {
        double acc[kSumSlots] = {0, 0, 0, 0};
        if (j < kr) {
            for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
#pragma unroll
                for (int i = 0; i < kSumSlots; ++i) acc[i] += tile_partial(partial, partial8, g.KB, t, j, i);
            }
        }
        block_sum<kSumSlots>(acc, scratch);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < kSumSlots; ++i) out[j * kSumSlots + i] = acc[i];
        }
    }
}

// Two-stage form for few problems with very many tiles (one large grid / cube): stage 1 lets kReduceChunks
// blocks per problem sum contiguous chunks of the tile partials, stage 2 adds the chunk sums in order.
constexpr int kReduceChunks = 64;

__global__ void __launch_bounds__(128) reduce_partials_stage1_kernel(Geom g, const double *__restrict__ partial,
                                                                     const double *__restrict__ partial8,
                                                                     const int *__restrict__ tile_start,
                                                                     const ProbState *__restrict__ state,
                                                                     double *__restrict__ stage, int all_tiles) {
    __shared__ double scratch[kSumSlots * 4];
    const int p = blockIdx.x / kReduceChunks, c = blockIdx.x - p * kReduceChunks;
    const int kr = state[p].k_run;
    double *out = stage + ((size_t)p * kReduceChunks + c) * g.KB * kSumSlots;
    const int t0 = all_tiles ? p * g.S * g.tiles_per_prob : tile_start[p * g.S];
    const int t1 = all_tiles ? t0 + g.S * g.tiles_per_prob : tile_start[(p + 1) * g.S];
    const int per = (t1 - t0 + kReduceChunks - 1) / kReduceChunks;
    const int a0 = min(t0 + c * per, t1), a1 = min(a0 + per, t1);
    for (int j = 0; j < g.KB; ++j) {
        double acc[kSumSlots] = {0, 0, 0, 0};
        if (j < kr) {
            for (int t = a0 + threadIdx.x; t < a1; t += blockDim.x) {
#pragma unroll
                for (int i = 0; i < kSumSlots; ++i) acc[i] += tile_partial(partial, partial8, g.KB, t, j, i);
            }
        }
        block_sum<kSumSlots>(acc, scratch);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < kSumSlots; ++i) out[j * kSumSlots + i] = acc[i];
        }
    }
}

__global__ void reduce_partials_stage2_kernel(Geom g, const double *__restrict__ stage, double *__restrict__ sums) {
    const int p = blockIdx.x;
    const int i = threadIdx.x;                      // (j, slot)
    if (i >= g.KB * kSumSlots) return;
    double acc = 0.0;
    for (int c = 0; c < kReduceChunks; ++c) acc += stage[((size_t)p * kReduceChunks + c) * g.KB * kSumSlots + i];
    sums[(size_t)p * g.KB * kSumSlots + i] = acc;
}

// Thread per problem: the stopping rule and the loop bookkeeping.
//   delta = nanmean(|new - old|) / nanmean(old) over the whole grid (VPint/SD_MRP.py:139,
//   VPint/WP_MRP.py:352); known cells contribute |orig - orig| (0, or NaN for +-inf) and orig.
//   delta <= threshold commits the sweep and stops (SD_MRP.py:142-147); an overshoot inside a
//   temporally blocked launch schedules a replay of exactly `stop` sweeps from the untouched
//   launch-entry buffer.
__global__ void decide_kernel(DecideArgs a) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.g.P) return;
    decide_problem(a.g, a.state, p, a.sums + (size_t)p * a.g.KB * kSumSlots, a.known0, a.known, a.delta_out, a.n_active,
                   a.max_iter, a.auto_terminate, a.threshold, a.all_cells);
}

// Ratio form of the tiled path (vpint_step_k1.cu, jacobi2d_k1_ratio_kernel): move both state buffers into
// r = P / f.  `save` keeps buffer 1, which holds the original value of every known cell, for the way back.
// The form is only used when load / weights found every value and feature finite and of moderate size
// (|.| < 1e150, |f| > 1e-150), so that f * r cannot overflow where the reference's w * P would not.
__global__ void __launch_bounds__(256) to_ratio_kernel(Geom g, double *__restrict__ buf0, double *__restrict__ buf1,
                                                       const double *__restrict__ f, double *__restrict__ save) {
    const size_t row = (size_t)blockIdx.x * g.pitch;
    for (int e = threadIdx.x; e < g.L; e += blockDim.x) {
        const double a = buf0[row + e], b = buf1[row + e], ff = f[row + e];
        save[row + e] = b;
        buf0[row + e] = a / ff;
        buf1[row + e] = b / ff;
    }
}

// Back to value space: unknown cells P = f * r of the current buffer, known cells their original value,
// written to BOTH buffers (the state both hold after any ordinary launch).
__global__ void __launch_bounds__(256) from_ratio_kernel(Geom g, const ProbState *__restrict__ state, double *buf0,
                                                         double *buf1, const double *__restrict__ f,
                                                         const double *__restrict__ save,
                                                         const uint32_t *__restrict__ mask) {
    const int p = blockIdx.x / g.H;
    const size_t row = (size_t)blockIdx.x * g.pitch;
    const uint32_t *mr = mask + (size_t)blockIdx.x * g.mpitch;
    const double *cur = state[p].cur ? buf1 : buf0;
    for (int e = threadIdx.x; e < g.L; e += blockDim.x) {
        const bool unk = (mr[e >> 5] >> (e & 31)) & 1u;
        const double v = unk ? f[row + e] * cur[row + e] : save[row + e];
        buf0[row + e] = v;
        buf1[row + e] = v;
    }
}

// Tail of the WHILE body of the graph loop (vpint_api.cu run_graph_loop): keep looping while a problem is
// still running and the launch budget lasts.
__global__ void loop_condition_kernel(cudaGraphConditionalHandle handle, const int *n_active, int *iter_left) {
    const int left = *iter_left - 1;
    *iter_left = left;
    cudaGraphSetConditional(handle, (*n_active > 0 && left > 0) ? 1u : 0u);
}

// Rare path: the start state disagreed with `original` on known cells (user-edited pred_grid, or a
// previous resistance run).  The first launch ran one sweep from it; the other ping-pong buffer still
// holds those stale known cells, so copy the restored ones over.
template <typename TSt>
__global__ void __launch_bounds__(256) restore_known_kernel(Geom g, const ProbState *__restrict__ state,
                                                            TSt *buf0, TSt *buf1,
                                                            const uint32_t *__restrict__ mask) {
    const int p = blockIdx.x / g.H, y = blockIdx.x - p * g.H;      // p: slice
    const ProbState st = state[p / g.S];
    if (!st.dirty || st.fresh) return;
    const TSt *src = st.cur ? buf1 : buf0;
    TSt *dst = st.cur ? buf0 : buf1;
    const size_t row = ((size_t)p * g.H + y) * g.pitch;
    const uint32_t *mr = mask + ((size_t)p * g.H + y) * g.mpitch;
    for (int e = threadIdx.x; e < g.pitch; e += blockDim.x) {
        if (!((mr[e >> 5] >> (e & 31)) & 1u)) dst[row + e] = src[row + e];
    }
}

__global__ void clear_dirty_kernel(Geom g, ProbState *state) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < g.P && !state[p].fresh) state[p].dirty = 0;
}

// self.pred_grid after run(): gather the current ping-pong buffer of every problem into the caller's
// layout (VPint/SD_MRP.py:149-156; VPint2 reassembles (H, W, B) at VPint/VPint2.py:96-99,112-116).
// One warp per row; rows that are contiguous in the caller's array move as 16-byte vectors.
constexpr int kResultWarps = 8;

template <typename TSt, typename TOut>
__global__ void __launch_bounds__(kResultWarps * 32) result_kernel(Geom g, const ProbState *__restrict__ state,
                                                                   const TSt *__restrict__ buf0, const TSt *__restrict__ buf1,
                                                                   TOut *__restrict__ out, Strides sd, int32_t *niter_out,
                                                                   int vec_ok) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kResultWarps + (threadIdx.x >> 5);
    if (r >= (int64_t)g.P * g.S * g.H) return;
    const int p = (int)(r / g.H), y = (int)(r - (int64_t)p * g.H);      // p: slice
    const ProbState st = state[p / g.S];
    const TSt *src = (st.cur ? buf1 : buf0) + ((size_t)p * g.H + y) * g.pitch;
    TOut *dst = out + cell_offset(g, sd, p, y, 0);
    if (vec_ok && sizeof(TSt) == 8 && sizeof(TOut) == 8) {              // unit stride, even length, 16-byte aligned rows
        const double2 *s2 = reinterpret_cast<const double2 *>(src);
        double2 *d2 = reinterpret_cast<double2 *>(dst);
        for (int e = lane; e < g.L / 2; e += 32) d2[e] = s2[e];
    } else {
        for (int e = lane; e < g.L; e += 32) dst[(int64_t)e * sd.sX] = (TOut)src[e];
    }
    if (niter_out && y == 0 && lane == 0 && p % g.S == 0) niter_out[p / g.S] = st.iters_done;
}

// The same gather while the state is in ratio space (tiled ratio path): unknown cells P = f * r of the current buffer,
// known cells their saved original value -- what from_ratio_kernel would write, without the extra pass over both state
// buffers; the solver stays in ratio space.
template <typename TOut>
__global__ void __launch_bounds__(kResultWarps * 32) result_ratio_kernel(Geom g, const ProbState *__restrict__ state,
                                                                         const double *__restrict__ buf0,
                                                                         const double *__restrict__ buf1,
                                                                         const double *__restrict__ f,
                                                                         const double *__restrict__ save,
                                                                         const uint32_t *__restrict__ mask, TOut *__restrict__ out,
                                                                         Strides sd, int32_t *niter_out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kResultWarps + (threadIdx.x >> 5);
    if (r >= (int64_t)g.P * g.H) return;
    const int p = (int)(r / g.H), y = (int)(r - (int64_t)p * g.H);
    const ProbState st = state[p];
    const size_t row = ((size_t)p * g.H + y) * g.pitch;
    const double *cur = (st.cur ? buf1 : buf0) + row;
    const uint32_t *mr = mask + ((size_t)p * g.H + y) * g.mpitch;
    TOut *dst = out + cell_offset(g, sd, p, y, 0);
    for (int e0 = 0; e0 < g.L; e0 += 32) {
        const int e = e0 + lane;
        const unsigned word = mr[e0 >> 5];
        if (e < g.L) {
            const double v = ((word >> lane) & 1u) ? f[row + e] * cur[e] : save[row + e];
            dst[(int64_t)e * sd.sX] = (TOut)v;
        }
    }
    if (niter_out && y == 0 && lane == 0) niter_out[p] = st.iters_done;
}

// Row-sharded scenes: copy `rows` owned boundary rows of every problem's CURRENT buffer into a
// contiguous [P][rows][pitch] message (pack), or a received message into the halo rows next to the owned
// band (unpack).  `first_row` is the local row the block of rows starts at.
template <typename TSt, bool PACK>
__global__ void __launch_bounds__(256) halo_rows_kernel(Geom g, const ProbState *__restrict__ state, TSt *buf0,
                                                        TSt *buf1, TSt *msg, int first_row, int rows) {
    const int p = blockIdx.x / rows, r = blockIdx.x - p * rows;
    TSt *grid_row = (state[p].cur ? buf1 : buf0) + ((size_t)p * g.H + first_row + r) * g.pitch;
    TSt *msg_row = msg + ((size_t)p * rows + r) * g.pitch;
    for (int e = threadIdx.x; e < g.pitch; e += blockDim.x) {
        if (PACK) msg_row[e] = grid_row[e];
        else grid_row[e] = msg_row[e];
    }
}

// Wait until *flag >= want (system-scope acquire), for at most ~10 s: a rank that died must not hang its neighbours'
// kernels for ever.  Returns false on a time-out (the caller records it; the results of the run are void then).
__device__ __forceinline__ bool wait_flag(const int *flag, int want) {
    unsigned long long t0 = 0;
    for (unsigned spins = 0;; ++spins) {
        int v;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v >= want) return true;
        if ((spins & 1023u) == 1023u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 10000000000ull) return false;
        }
    }
}

// ---- row-sharded scenes, low-latency protocol ----------------------------------------------------------------
// Stop decision with a lag of one sweep (the trick of the cluster-resident kernel): launch t computes sweep t and its
// local sums; while their all-reduce travels, launch t + 1 already computes sweep t + 1.  This kernel runs after the
// stencil of launch t + 1: it first takes the decision on sweep t from the reduced sums (stop = sweep t + 1 is simply
// not committed -- commit-then-break, VPint/SD_MRP.py:143-147, with the exact stopping sweep and no rollback), and
// otherwise commits sweep t + 1 (flips the ping-pong buffer).  One sweep per launch only.
__global__ void lagged_commit_kernel(Geom g, ProbState *state, const double *__restrict__ sums_prev,
                                     const double *__restrict__ known0, const double *__restrict__ known, int *n_active,
                                     int max_iter, int auto_terminate, double threshold) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.P) return;
    ProbState st = state[p];
    if (st.status != 0 || st.k_run == 0) return;
    if (sums_prev) {
        const double *S = sums_prev + (size_t)p * kSumSlots;
        const double *K = (st.fresh ? known0 : known) + (size_t)p * kSumSlots;
        const double s_abs = S[0] + K[0], s_old = S[1] + K[1];
        const double n_abs = g.n_cells - (S[2] + K[2]), n_old = g.n_cells - (S[3] + K[3]);
        const double delta = (s_abs / n_abs) / (s_old / n_old);
        st.fresh = 0;
        if (auto_terminate && delta <= threshold) {
            st.status = 1; st.k_run = 0;
            state[p] = st;
            atomicSub(n_active, 1);
            return;
        }
    }
    st.iters_done += 1;
    st.cur ^= 1;
    if (st.iters_done >= max_iter) {
        st.status = 1; st.k_run = 0; st.fresh = 0;
        atomicSub(n_active, 1);
    }
    state[p] = st;
}

// Halo rows straight into the neighbour's memory (peer-mapped over NVLink) instead of pack -> NCCL send/recv -> unpack:
// every block copies one boundary row of one problem's CURRENT state into the neighbour's receive slot; the block that
// finishes last publishes the sequence number of this exchange in the neighbour's flag (system-scope release).
// blockIdx.y: 0 = rows for the upper neighbour (my first owned rows), 1 = rows for the lower neighbour.
__global__ void __launch_bounds__(256) halo_push_kernel(Geom g, const ProbState *__restrict__ state, const double *buf0,
                                                        const double *buf1, double *dst_up, double *dst_down, int *flag_up,
                                                        int *flag_down, int rows, int seq, int *tickets) {
    const int side = blockIdx.y;
    double *dst = side ? dst_down : dst_up;
    int *flag = side ? flag_down : flag_up;
    if (!dst) return;
    const int p = blockIdx.x / rows, r = blockIdx.x - p * rows;
    const int first = side ? g.own_row0 + g.own_rows - rows : g.own_row0;
    const double *src = (state[p].cur ? buf1 : buf0) + ((size_t)p * g.H + first + r) * g.pitch;
    double *out = dst + ((size_t)p * rows + r) * g.pitch;
    for (int e = threadIdx.x; e < g.pitch; e += blockDim.x) out[e] = src[e];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(&tickets[side], 1);
        if (done == (int)gridDim.x - 1) {
            tickets[side] = 0;
            __threadfence_system();
            asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
        }
    }
}

// Receiver: wait until the neighbour has published exchange `seq`, then copy its rows into the halo rows next to the
// owned band.  blockIdx.y: 0 = from the upper neighbour (rows above the band), 1 = from the lower neighbour.
__global__ void __launch_bounds__(256) halo_pull_kernel(Geom g, const ProbState *__restrict__ state, double *buf0, double *buf1,
                                                        const double *src_up, const double *src_down, const int *flag_up,
                                                        const int *flag_down, int rows, int seq) {
    const int side = blockIdx.y;
    const double *src = side ? src_down : src_up;
    const int *flag = side ? flag_down : flag_up;
    if (!src) return;
    if (threadIdx.x == 0) wait_flag(flag, seq);      // bounded: a dead neighbour must not hang this GPU
    __syncthreads();
    const int p = blockIdx.x / rows, r = blockIdx.x - p * rows;
    const int first = side ? g.own_row0 + g.own_rows : g.own_row0 - rows;
    double *dstrow = (state[p].cur ? buf1 : buf0) + ((size_t)p * g.H + first + r) * g.pitch;
    const double *in = src + ((size_t)p * rows + r) * g.pitch;
    for (int e = threadIdx.x; e < g.pitch; e += blockDim.x) dstrow[e] = __ldcv(in + e);
}

// The whole reduction + stop decision + commit of one launch of a row-sharded run, without a collective library: every
// rank stores its per-sweep sums into every rank's peer block (plain stores over NVLink + one flag per source rank), and
// adds the `world` contributions up in rank order -- the same order on every rank, so all ranks take the same decision.
// Launch t (lifetime counter ctr[0]):  wait for the sums of launch t - 1 from every rank; decide on sweep t - 1 (a stop
// = sweep t is not committed, see lagged_commit_kernel); commit sweep t; publish the sums of sweep t; count the launch;
// report "problems still running" to the host through mapped pinned memory (tagged with the launch number, so that the
// host polls without events or copies).  One block.
__global__ void __launch_bounds__(256) peer_commit_kernel(Geom g, ProbState *state, const double *__restrict__ my_sums,
                                                          const double *__restrict__ known0, const double *__restrict__ known,
                                                          int *n_active, int max_iter, int auto_terminate, double threshold,
                                                          PeerCtx pc, int *ctr, unsigned long long *host_ring) {
    const int t = ctr[0], first = ctr[1];
    const int tid = threadIdx.x;
    const size_t slot_elems = (size_t)pc.world * g.P * kSumSlots;
    char *mine = pc.base[pc.rank];
    if (!first && tid < pc.world) {
        const int *flag = reinterpret_cast<const int *>(mine + pc.off_sflag + (size_t)tid * 128);
        if (!wait_flag(flag, t)) ctr[2] = 1;          // a rank never published its sums: recorded, reported by the host
    }
    __syncthreads();
    const double *prev = reinterpret_cast<const double *>(mine + pc.off_gather) + (size_t)((t - 1) & 1) * slot_elems;
    for (int p = tid; p < g.P; p += blockDim.x) {
        ProbState st = state[p];
        if (st.status != 0 || st.k_run == 0) continue;
        bool stop = false;
        if (!first) {
            double S[kSumSlots] = {0.0, 0.0, 0.0, 0.0};
            for (int r = 0; r < pc.world; ++r) {
                const double *q = prev + ((size_t)r * g.P + p) * kSumSlots;
#pragma unroll
                for (int i = 0; i < kSumSlots; ++i) S[i] += __ldcv(q + i);
            }
            const double *K = (st.fresh ? known0 : known) + (size_t)p * kSumSlots;
            const double s_abs = S[0] + K[0], s_old = S[1] + K[1];
            const double n_abs = g.n_cells - (S[2] + K[2]), n_old = g.n_cells - (S[3] + K[3]);
            const double delta = (s_abs / n_abs) / (s_old / n_old);
            st.fresh = 0;
            stop = auto_terminate && delta <= threshold;
        }
        if (stop) {
            st.status = 1; st.k_run = 0;
            atomicSub(n_active, 1);
        } else {
            st.iters_done += 1;
            st.cur ^= 1;
            if (st.iters_done >= max_iter) {
                st.status = 1; st.k_run = 0; st.fresh = 0;
                atomicSub(n_active, 1);
            }
        }
        state[p] = st;
    }
    // publish the sums of sweep t
    const int n_mine = g.P * kSumSlots;
    for (int idx = tid; idx < pc.world * n_mine; idx += blockDim.x) {
        const int r = idx / n_mine, e = idx - r * n_mine;
        double *dst = reinterpret_cast<double *>(pc.base[r] + pc.off_gather) + (size_t)(t & 1) * slot_elems + (size_t)pc.rank * n_mine;
        dst[e] = my_sums[e];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < pc.world) {
        int *flag = reinterpret_cast<int *>(pc.base[tid] + pc.off_sflag + (size_t)pc.rank * 128);
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(t + 1) : "memory");
    }
    if (tid == 0) {
        ctr[0] = t + 1;
        ctr[1] = 0;
        const int left = *reinterpret_cast<volatile int *>(n_active);
        *reinterpret_cast<volatile unsigned long long *>(host_ring + ((unsigned)t & 63u)) =
            ((unsigned long long)(unsigned)(t + 1) << 32) | (unsigned)(left < 0 ? 0 : left);
    }
}

// Halo rows of the new state to both row neighbours (push blocks) and the neighbours' rows into the halo rows (pull blocks).
// A row is moved by kHaloChunks blocks as 16-byte vectors.
//   mode 1 (push, launched right AFTER THE STENCIL, before the sums are reduced and committed): the exchange number is the
//          one peer_commit_kernel is about to publish (ctr[0] + 1) and the source is the buffer the stencil has just
//          written (cur ^ 1 of a running problem) -- the neighbour has the rows while this rank still reduces its sums;
//   mode 2 (pull, launched after peer_commit_kernel): waits for the neighbours' flags of exchange ctr[0] and copies their
//          rows into the halo rows of the (now committed) current buffer;
//   mode 0: both in one launch after the commit (push blocks [0, 2m), pull blocks [2m, 4m): every push block is
//          dispatched before any spinning pull block).
constexpr int kHaloChunks = 4;

__global__ void __launch_bounds__(256) peer_halo_kernel(Geom g, const ProbState *__restrict__ state, double *buf0, double *buf1,
                                                        PeerCtx pc, int rows, const int *__restrict__ ctr, int *tickets, int mode) {
    const int seq = ctr[0] + (mode == 1 ? 1 : 0);
    const int slot = seq & 1;
    const int n = g.P * rows, m = n * kHaloChunks;
    const int bid = (int)blockIdx.x;
    const bool pull = (mode == 2) || (mode == 0 && bid >= 2 * m);
    const int local = (mode == 0 && pull) ? bid - 2 * m : bid;
    const int side = (local / m) & 1;                 // 0: the upper neighbour, 1: the lower neighbour
    const int b = (local % m) / kHaloChunks, chunk = local % kHaloChunks;
    const int nb = side ? pc.down : pc.up;
    if (nb < 0) return;
    const int p = b / rows, r = b - p * rows;
    const size_t msg = pc.msg_bytes, row_off = ((size_t)p * rows + r) * g.pitch;
    const ProbState ps = state[p];
    // before the commit a running problem's new state is in the other buffer; a finished one keeps its buffer
    const int which = (mode == 1 && ps.k_run != 0) ? (ps.cur ^ 1) : ps.cur;
    double *cur = which ? buf1 : buf0;
    const int v_all = g.pitch >> 1;                    // double2 elements of a row (pitch is a multiple of 64)
    const int v_per = (v_all + kHaloChunks - 1) / kHaloChunks;
    const int v0 = chunk * v_per, v1 = min(v_all, v0 + v_per);
    if (!pull) {
        // my first owned rows -> the upper neighbour's "from below" slot; my last owned rows -> the lower one's "from above"
        double2 *dst = reinterpret_cast<double2 *>(reinterpret_cast<double *>(pc.base[nb] + (size_t)(2 * slot + (side ? 0 : 1)) * msg) + row_off);
        int *flag = reinterpret_cast<int *>(pc.base[nb] + 4 * msg + (side ? 0 : 128));
        const int first = side ? g.own_row0 + g.own_rows - rows : g.own_row0;
        const double2 *src = reinterpret_cast<const double2 *>(cur + ((size_t)p * g.H + first + r) * g.pitch);
        for (int e = v0 + threadIdx.x; e < v1; e += blockDim.x) dst[e] = src[e];
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const int done = atomicAdd(&tickets[side], 1);
            if (done == m - 1) {
                tickets[side] = 0;
                __threadfence_system();
                asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
            }
        }
    } else {
        char *mine = pc.base[pc.rank];
        const double2 *in = reinterpret_cast<const double2 *>(reinterpret_cast<const double *>(mine + (size_t)(2 * slot + side) * msg) + row_off);
        const int *flag = reinterpret_cast<const int *>(mine + 4 * msg + (side ? 128 : 0));
        if (threadIdx.x == 0 && !wait_flag(flag, seq)) const_cast<int *>(ctr)[2] = 1;
        __syncthreads();
        const int first = side ? g.own_row0 + g.own_rows : g.own_row0 - rows;
        double2 *dstrow = reinterpret_cast<double2 *>(cur + ((size_t)p * g.H + first + r) * g.pitch);
        for (int e = v0 + threadIdx.x; e < v1; e += blockDim.x) dstrow[e] = __ldcv(in + e);
    }
}

}  // namespace

#define VP_LAUNCH_CHECK()                         \
    do {                                          \
        cudaError_t e__ = cudaGetLastError();     \
        if (e__ != cudaSuccess) return e__;       \
        if (launches) ++*launches;                \
    } while (0)

bool fill_mean_supported(const Geom &g) { return g.ndim == 2 && (int64_t)g.H * g.L <= (int64_t)kFillMaxLeaves * 64; }

cudaError_t launch_fill_mean(const Geom &g, const void *original, int dtype, const int64_t *stride, double *fill_dev,
                             cudaStream_t st, int64_t *launches) {
    const Strides sd{stride[0], stride[1], stride[2], stride[3], (g.ndim == 3) ? stride[4] : 0};
    const int64_t n = (int64_t)g.H * g.L;
    const size_t esz = (dtype == VPINT_F64) ? 8 : 4;
    const size_t bands_smem = ((size_t)8 * 128 * g.Pin + (size_t)g.Pin * (n / 128)) * esz;
    if (g.ndim == 2 && g.Pin >= 2 && g.Pin <= kFillBandsMax && sd.sI == 1 && sd.sX == g.Pin && sd.sY == (int64_t)g.L * g.Pin &&
        n >= 128 && (n & (n - 1)) == 0 && bands_smem <= 200 * 1024) {
        cudaError_t e;
        if (dtype == VPINT_F64) {
            e = cudaFuncSetAttribute(fill_mean_bands_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bands_smem);
            if (e != cudaSuccess) return e;
            fill_mean_bands_kernel<double><<<g.P / g.Pin, 256, bands_smem, st>>>(g, (const double *)original, sd.sO, fill_dev);
        } else {
            e = cudaFuncSetAttribute(fill_mean_bands_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bands_smem);
            if (e != cudaSuccess) return e;
            fill_mean_bands_kernel<float><<<g.P / g.Pin, 256, bands_smem, st>>>(g, (const float *)original, sd.sO, fill_dev);
        }
        VP_LAUNCH_CHECK();
        return cudaSuccess;
    }
    if (dtype == VPINT_F64) fill_mean_kernel<double><<<g.P, 256, 0, st>>>(g, (const double *)original, sd, fill_dev);
    else fill_mean_kernel<float><<<g.P, 256, 0, st>>>(g, (const float *)original, sd, fill_dev);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_prepare(const Geom &g, const void *original, const void *pred0, const double *fill_dev,
                           int dtype, const int64_t *stride, int storage, void *bufA, void *bufB, void *orig_copy,
                           uint32_t *mask, double *rowpart, double *known0, double *known, double *extreme,
                           ProbState *state, int *counters, cudaStream_t st, int64_t *launches) {
    const Strides sd{stride[0], stride[1], stride[2], stride[3], (g.ndim == 3) ? stride[4] : 0};
    const unsigned grid = (unsigned)((int64_t)g.P * g.S * g.H);
    const size_t esz = (dtype == VPINT_F64) ? 8 : 4;
    const size_t bands_smem = (size_t)kPrepWarps * g.L * g.Pin * esz;
    if (g.ndim == 2 && !pred0 && g.Pin >= 2 && sd.sI == 1 && sd.sX == g.Pin && g.L <= 32768 && bands_smem <= 100 * 1024) {
        const int64_t n_rows = (int64_t)(g.P / g.Pin) * g.H;
        const unsigned bgrid = (unsigned)((n_rows + kPrepWarps - 1) / kPrepWarps);
#define VP_PREPB(TIN, TST)                                                                                          \
    do {                                                                                                            \
        cudaError_t e__ = cudaFuncSetAttribute(prepare_bands_kernel<TIN, TST>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               (int)bands_smem);                                                    \
        if (e__ != cudaSuccess) return e__;                                                                         \
        prepare_bands_kernel<TIN, TST><<<bgrid, kPrepWarps * 32, bands_smem, st>>>(                                 \
            g, (const TIN *)original, fill_dev, sd, (TST *)bufA, (TST *)bufB, (TST *)orig_copy, mask, rowpart, extreme); \
    } while (0)
        if (dtype == VPINT_F64 && storage == VPINT_F64) VP_PREPB(double, double);
        else if (dtype == VPINT_F32 && storage == VPINT_F64) VP_PREPB(float, double);
        else if (dtype == VPINT_F64 && storage == VPINT_F32) VP_PREPB(double, float);
        else VP_PREPB(float, float);
#undef VP_PREPB
        VP_LAUNCH_CHECK();
        prepare_reduce_kernel<<<g.P, 256, 0, st>>>(g, rowpart, known0, known, state, counters);
        VP_LAUNCH_CHECK();
        return cudaSuccess;
    }
#define VP_PREP(TIN, TST)                                                                                  \
    prepare_kernel<TIN, TST><<<grid, 256, 0, st>>>(g, (const TIN *)original, (const TIN *)pred0, fill_dev, \
                                                   sd, (TST *)bufA, (TST *)bufB, (TST *)orig_copy, mask, rowpart, extreme)
    if (dtype == VPINT_F64 && storage == VPINT_F64) VP_PREP(double, double);
    else if (dtype == VPINT_F32 && storage == VPINT_F64) VP_PREP(float, double);
    else if (dtype == VPINT_F64 && storage == VPINT_F32) VP_PREP(double, float);
    else VP_PREP(float, float);
#undef VP_PREP
    VP_LAUNCH_CHECK();
    prepare_reduce_kernel<<<g.P, 256, 0, st>>>(g, rowpart, known0, known, state, counters);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_prepare_reduce(const Geom &g, const double *rowpart, double *known0, double *known, ProbState *state,
                                  int *counters, cudaStream_t st, int64_t *launches) {
    prepare_reduce_kernel<<<g.P, 256, 0, st>>>(g, rowpart, known0, known, state, counters);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_tiles(const Geom &g, const uint32_t *mask, int *tile_flag, TileRef *tiles, int *tile_start,
                         int *counters, cudaStream_t st, int64_t *launches) {
    const int64_t n_tiles = (int64_t)g.P * g.S * g.tiles_per_prob;
    const unsigned grid = (unsigned)((n_tiles + 7) / 8);
    tile_flag_kernel<<<grid, 256, 0, st>>>(g, mask, tile_flag, n_tiles);
    VP_LAUNCH_CHECK();
    tile_compact_kernel<<<1, 1024, 0, st>>>(g, tile_flag, tiles, tile_start, counters, n_tiles, nullptr);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_recompact_tiles(const Geom &g, const int *tile_flag, TileRef *tiles, int *tile_start, int *counters,
                                   const ProbState *state, cudaStream_t st, int64_t *launches) {
    const int64_t n_tiles = (int64_t)g.P * g.S * g.tiles_per_prob;
    tile_compact_kernel<<<1, 1024, 0, st>>>(g, tile_flag, tiles, tile_start, counters, n_tiles, state);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_init_state(const Geom &g, ProbState *state, int *n_active, int max_iter, cudaStream_t st,
                              int64_t *launches) {
    cudaError_t e = cudaMemsetAsync(n_active, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    init_state_kernel<<<(g.P + 127) / 128, 128, 0, st>>>(g, state, n_active, max_iter);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

size_t reduce_stage_bytes(const Geom &g) { return (size_t)g.P * kReduceChunks * g.KB * kSumSlots * sizeof(double); }

cudaError_t launch_reduce_partials(const Geom &g, const double *partial, const double *partial8, const int *tile_start,
                                   const ProbState *state, double *sums, double *stage, bool all_tiles,
                                   cudaStream_t st, int64_t *launches) {
    // one block per problem is enough unless a problem has tens of thousands of tiles
    if (stage && (int64_t)g.S * g.tiles_per_prob >= 8192) {
        reduce_partials_stage1_kernel<<<g.P * kReduceChunks, 128, 0, st>>>(g, partial, partial8, tile_start, state, stage,
                                                                           all_tiles ? 1 : 0);
        VP_LAUNCH_CHECK();
        reduce_partials_stage2_kernel<<<g.P, 64, 0, st>>>(g, stage, sums);
        VP_LAUNCH_CHECK();
        return cudaSuccess;
    }
    reduce_partials_kernel<<<g.P, 128, 0, st>>>(g, partial, partial8, tile_start, state, sums, all_tiles ? 1 : 0);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_decide(const DecideArgs &a, cudaStream_t st, int64_t *launches) {
    decide_kernel<<<(a.g.P + 127) / 128, 128, 0, st>>>(a);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_to_ratio(const Geom &g, void *const *buf, const double *f, double *save, cudaStream_t st,
                            int64_t *launches) {
    to_ratio_kernel<<<(unsigned)((int64_t)g.P * g.H), 256, 0, st>>>(g, (double *)buf[0], (double *)buf[1], f, save);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_from_ratio(const Geom &g, const ProbState *state, void *const *buf, const double *f,
                              const double *save, const uint32_t *mask, cudaStream_t st, int64_t *launches) {
    from_ratio_kernel<<<(unsigned)((int64_t)g.P * g.H), 256, 0, st>>>(g, state, (double *)buf[0], (double *)buf[1], f, save,
                                                                       mask);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_loop_condition(cudaGraphConditionalHandle handle, const int *n_active, int *iter_left,
                                  cudaStream_t st, int64_t *launches) {
    loop_condition_kernel<<<1, 1, 0, st>>>(handle, n_active, iter_left);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_restore_known(const Geom &g, int storage, ProbState *state, void *const *buf,
                                 const uint32_t *mask, cudaStream_t st, int64_t *launches) {
    const unsigned grid = (unsigned)((int64_t)g.P * g.S * g.H);
    if (storage == VPINT_F64)
        restore_known_kernel<double><<<grid, 256, 0, st>>>(g, state, (double *)buf[0], (double *)buf[1], mask);
    else
        restore_known_kernel<float><<<grid, 256, 0, st>>>(g, state, (float *)buf[0], (float *)buf[1], mask);
    VP_LAUNCH_CHECK();
    clear_dirty_kernel<<<(g.P + 127) / 128, 128, 0, st>>>(g, state);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_halo_rows(const Geom &g, int storage, const ProbState *state, void *const *buf, void *msg,
                             int first_row, int rows, bool pack, cudaStream_t st, int64_t *launches) {
    const unsigned grid = (unsigned)((int64_t)g.P * rows);
    if (storage == VPINT_F64) {
        if (pack) halo_rows_kernel<double, true><<<grid, 256, 0, st>>>(g, state, (double *)buf[0], (double *)buf[1], (double *)msg, first_row, rows);
        else halo_rows_kernel<double, false><<<grid, 256, 0, st>>>(g, state, (double *)buf[0], (double *)buf[1], (double *)msg, first_row, rows);
    } else {
        if (pack) halo_rows_kernel<float, true><<<grid, 256, 0, st>>>(g, state, (float *)buf[0], (float *)buf[1], (float *)msg, first_row, rows);
        else halo_rows_kernel<float, false><<<grid, 256, 0, st>>>(g, state, (float *)buf[0], (float *)buf[1], (float *)msg, first_row, rows);
    }
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_result_ratio(const Geom &g, const ProbState *state, void *const *buf, const double *f, const double *save,
                                const uint32_t *mask, void *out, int dtype, const int64_t *stride, int32_t *niter_out,
                                cudaStream_t st, int64_t *launches) {
    const Strides sd{stride[0], stride[1], stride[2], stride[3], 0};
    const unsigned grid = (unsigned)(((int64_t)g.P * g.H + kResultWarps - 1) / kResultWarps);
    if (dtype == VPINT_F64)
        result_ratio_kernel<double><<<grid, kResultWarps * 32, 0, st>>>(g, state, (const double *)buf[0], (const double *)buf[1], f,
                                                                       save, mask, (double *)out, sd, niter_out);
    else
        result_ratio_kernel<float><<<grid, kResultWarps * 32, 0, st>>>(g, state, (const double *)buf[0], (const double *)buf[1], f,
                                                                      save, mask, (float *)out, sd, niter_out);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_lagged_commit(const DecideArgs &a, const double *sums_prev, cudaStream_t st, int64_t *launches) {
    lagged_commit_kernel<<<(a.g.P + 127) / 128, 128, 0, st>>>(a.g, a.state, sums_prev, a.known0, a.known, a.n_active,
                                                             a.max_iter, a.auto_terminate, a.threshold);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_halo_push(const Geom &g, const ProbState *state, void *const *buf, void *dst_up, void *dst_down,
                             int *flag_up, int *flag_down, int rows, int seq, int *tickets, cudaStream_t st, int64_t *launches) {
    if (!dst_up && !dst_down) return cudaSuccess;
    halo_push_kernel<<<dim3((unsigned)((int64_t)g.P * rows), 2), 256, 0, st>>>(
        g, state, (const double *)buf[0], (const double *)buf[1], (double *)dst_up, (double *)dst_down, flag_up, flag_down, rows,
        seq, tickets);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_halo_pull(const Geom &g, const ProbState *state, void *const *buf, const void *src_up, const void *src_down,
                             const int *flag_up, const int *flag_down, int rows, int seq, cudaStream_t st, int64_t *launches) {
    if (!src_up && !src_down) return cudaSuccess;
    halo_pull_kernel<<<dim3((unsigned)((int64_t)g.P * rows), 2), 256, 0, st>>>(
        g, state, (double *)buf[0], (double *)buf[1], (const double *)src_up, (const double *)src_down, flag_up, flag_down, rows,
        seq);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_peer_commit(const DecideArgs &a, const double *my_sums, const PeerCtx &pc, int *ctr,
                               unsigned long long *host_ring, cudaStream_t st, int64_t *launches) {
    peer_commit_kernel<<<1, 256, 0, st>>>(a.g, a.state, my_sums, a.known0, a.known, a.n_active, a.max_iter, a.auto_terminate,
                                          a.threshold, pc, ctr, host_ring);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_peer_halo(const Geom &g, const ProbState *state, void *const *buf, const PeerCtx &pc, int rows, const int *ctr,
                             int *tickets, int mode, cudaStream_t st, int64_t *launches) {
    if (pc.up < 0 && pc.down < 0) return cudaSuccess;
    const int64_t m = (int64_t)g.P * rows * kHaloChunks;
    peer_halo_kernel<<<(unsigned)((mode == 0 ? 4 : 2) * m), 256, 0, st>>>(g, state, (double *)buf[0], (double *)buf[1], pc, rows, ctr,
                                                                         tickets, mode);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_result(const Geom &g, int storage, const ProbState *state, void *const *buf, void *out,
                          int dtype, const int64_t *stride, int32_t *niter_out, cudaStream_t st,
                          int64_t *launches) {
    const Strides sd{stride[0], stride[1], stride[2], stride[3], (g.ndim == 3) ? stride[4] : 0};
    const unsigned grid = (unsigned)(((int64_t)g.P * g.S * g.H + kResultWarps - 1) / kResultWarps);
    // 16-byte vectors need unit stride along the row, an even row length and every row start on a 16-byte boundary
    const int vec_ok = (sd.sX == 1 && (g.L & 1) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (sd.sO & 1) == 0 &&
                        (sd.sI & 1) == 0 && (sd.sY & 1) == 0 && (sd.sT & 1) == 0 && (g.pitch & 1) == 0) ? 1 : 0;
#define VP_RES(TST, TOUT)                                                                                              \
    result_kernel<TST, TOUT><<<grid, kResultWarps * 32, 0, st>>>(g, state, (const TST *)buf[0], (const TST *)buf[1],   \
                                                                 (TOUT *)out, sd, niter_out, vec_ok)
    if (storage == VPINT_F64 && dtype == VPINT_F64) VP_RES(double, double);
    else if (storage == VPINT_F64 && dtype == VPINT_F32) VP_RES(double, float);
    else if (storage == VPINT_F32 && dtype == VPINT_F64) VP_RES(float, double);
    else VP_RES(float, float);
#undef VP_RES
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

}  // namespace vpint
