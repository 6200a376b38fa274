// N3 + N4 (device side) fused with A1 + A2 for VPint2's band-interleaved stacks.
//
// The reference builds a VPint2_interpolator on the host (VPint/VPint2.py:13-55: dtype cast, optional clip, cloud
// mask -> NaN on every band, optional cloud buffering, all in per-pixel Python loops) and every band's WP_SMRP then
// initialises its unknown cells with np.nanmean (VPint/MRP.py:64-91) and replaces zero features (VPint/WP_MRP.py:
// 143-145).  Here the RAW rasters -- uint16 as Sentinel-2 products are stored (VPint/utils/EO_utils.py:66), or
// float32 / float64 -- go to the device as they are and one pass over them produces everything run() needs:
//
//   cloud_mask_kernel      mask > threshold                                   (VPint2.py:122-132)
//   footprint / dilate_*   cloud buffering as a separable dilation            (VPint2.py:135-150, :146 quirk kept)
//   fill_leaf_kernel, fill_subtree_kernel, fill_top_kernel
//                          np.nanmean of every band in the interpolator's dtype, NumPy's pairwise summation
//                          tree reproduced bit for bit for ANY grid size (MRP.py:73-77); row shards compute the
//                          leaves they own and all-reduce the leaf sums
//   ingest_bands_kernel    state plane, unknown-cell bit mask, feature plane, known-cell sums of the stopping rule
//                          and -- for a full result -- the known cells of the output, from ONE read of the rasters
//   result_unknown_kernel  the unknown cells of the result only, written as contiguous runs (the caller's array may
//                          be pinned host memory mapped into the device: only the cloud cells cross the link)
#include <cstdlib>

#include "vpint_internal.cuh"

namespace vpint {

namespace {

struct Strides {
    int64_t sO, sI, sY, sX;
};

// dtype cast + optional clip of the constructor (VPint2.py:23-41): np.clip keeps NaN and works in the array's dtype.
template <typename TAcc, typename TIn>
__device__ __forceinline__ TAcc conv_in(TIn raw, int clip, double lo, double hi) {
    TAcc v = (TAcc)raw;
    if (clip) {
        const TAcc l = (TAcc)lo, h = (TAcc)hi;
        v = (v < l) ? l : v;
        v = (v > h) ? h : v;
    }
    return v;
}

// Result element in the caller's dtype.  uint16 is the write-back of a Sentinel-2 product (N4): what
// np.clip(pred, 0, 65535).astype(np.uint16) gives (truncation toward zero, NaN -> 0).
template <typename TOut>
__device__ __forceinline__ TOut to_out(double v) { return (TOut)v; }
template <>
__device__ __forceinline__ uint16_t to_out<uint16_t>(double v) {
    if (!(v > 0.0)) return 0;
    return (uint16_t)(v >= 65535.0 ? 65535.0 : v);
}

__device__ __forceinline__ void store_out(void *out, int code, int64_t idx, double v) {
    if (code == VPINT_F32) reinterpret_cast<float *>(out)[idx] = (float)v;
    else if (code == VPINT_U16) reinterpret_cast<uint16_t *>(out)[idx] = to_out<uint16_t>(v);
    else reinterpret_cast<double *>(out)[idx] = v;
}

template <typename TM>
__global__ void __launch_bounds__(256) cloud_mask_kernel(const TM *__restrict__ mask, double threshold, int64_t n,
                                                         uint8_t *__restrict__ cloud) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cloud[i] = ((double)mask[i] > threshold) ? 1 : 0;
}

// Pixels with a NaN in any band (np.isnan(img[i, j, :]).any(), VPint2.py:142), OR the all-band cloud mask.
template <typename TIn>
__global__ void __launch_bounds__(256) footprint_kernel(const TIn *__restrict__ target, const uint8_t *__restrict__ cloud,
                                                        int64_t n_px, int B, uint8_t *__restrict__ foot) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    bool any = cloud ? (cloud[i] != 0) : false;
    const TIn *px = target + i * B;
    for (int b = 0; b < B && !any; ++b) {
        const TIn v = px[b];
        any = (v != v);
    }
    foot[i] = any ? 1 : 0;
}

// A source pixel (i, j) blanks rows [i-b, i+b) and columns [j-b, min(H, j+b)) (VPint2.py:143-147), so an output pixel
// (y, x) is covered iff a source lies in rows (y-b, y+b] and columns (x-b, x+b] -- and x < H, the reference's
// column bound.  Rows first, then columns.
__global__ void __launch_bounds__(256) dilate_rows_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int H,
                                                          int W, int b, int64_t n_px) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    const int64_t img = i / ((int64_t)H * W), r = i - img * H * W;
    const int y = (int)(r / W), x = (int)(r - (int64_t)y * W);
    const uint8_t *base = in + img * H * W;
    bool any = false;
    for (int yy = max(0, y - b + 1); yy <= min(H - 1, y + b) && !any; ++yy) any = base[(int64_t)yy * W + x] != 0;
    out[i] = any ? 1 : 0;
}

// grown = dilate_cols(tall) & (x < col_limit); next footprint = foot0 | grown; all-band mask = cloud0 | grown.
__global__ void __launch_bounds__(256) dilate_cols_kernel(const uint8_t *__restrict__ tall, const uint8_t *__restrict__ foot0,
                                                          const uint8_t *__restrict__ cloud0, uint8_t *__restrict__ foot_next,
                                                          uint8_t *__restrict__ cloud_out, int H, int W, int b, int col_limit,
                                                          int64_t n_px) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    const int64_t row0 = i - (i % W);
    const int x = (int)(i - row0);
    bool any = false;
    if (x < col_limit)
        for (int xx = max(0, x - b + 1); xx <= min(W - 1, x + b) && !any; ++xx) any = tall[row0 + xx] != 0;
    foot_next[i] = (any || foot0[i]) ? 1 : 0;
    if (cloud_out) cloud_out[i] = (any || (cloud0 && cloud0[i])) ? 1 : 0;     // in place on cloud0 is fine: own cell only
}

// ---- np.nanmean of every band, any grid size ---------------------------------------------------------
// NumPy sums the row-major copy of a band with its blocked pairwise scheme: a range of n > 128 elements splits at
// n2 = n/2 - (n/2) % 8, a range of <= 128 elements (a LEAF) is summed with 8 running accumulators.  Every range
// above the leaves has >= 129 elements, so both halves have >= 64: a leaf starts at a multiple of 8, holds 64..128
// cells, and at most one leaf starts inside any aligned run of 64 cells (a CHUNK).  Chunks therefore index the
// leaves without a table: the warp of chunk c walks down the tree to the leaf that holds cell 64c + 63 and owns it if
// that leaf starts inside the chunk.
__device__ __forceinline__ void leaf_of(int64_t n, int64_t idx, int64_t &lo, int &m) {
    int64_t a = 0, len = n;
    while (len > 128) {
        int64_t h = len / 2;
        h -= h % 8;
        if (idx < a + h) len = h;
        else { a += h; len -= h; }
    }
    lo = a;
    m = (int)len;
}

constexpr int kLeafWarps = 8;
constexpr int kLeafPerWarp = 8;      // consecutive chunks per warp (amortises the per-band count atomics)
constexpr int kLeafMaxBands = 32;

// One warp per kLeafPerWarp chunks; all B bands of a leaf at once (a leaf of m cells is a contiguous run of m * B
// values in a band-interleaved raster).  leaf_sum[(img * B + b) * n_chunks + c]; count[img * B + b] += known cells.
// Row shards: cell index = global row-major index; the shard holds cells [cell_offset, cell_offset + cells_local)
// and owns the leaves that START in [own_cell0, own_cell1).
template <typename TIn, typename TAcc>
__global__ void __launch_bounds__(kLeafWarps * 32) fill_leaf_kernel(const TIn *__restrict__ target,
                                                                    const uint8_t *__restrict__ cloud, int n_img, int B,
                                                                    int64_t n, int64_t n_chunks, int64_t groups_per_img,
                                                                    int64_t cell_offset, int64_t cells_local,
                                                                    int64_t own_cell0, int64_t own_cell1, int clip,
                                                                    double lo_v, double hi_v, TAcc *__restrict__ leaf_sum,
                                                                    unsigned long long *__restrict__ count) {
    extern __shared__ __align__(16) unsigned char leaf_raw[];
    __shared__ unsigned cnt_s[kLeafMaxBands];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < kLeafMaxBands) cnt_s[threadIdx.x] = 0;
    __syncthreads();
    // a block stays inside one image: blockIdx.x = img * blocks_per_img + k
    const int64_t blocks_per_img = (groups_per_img + kLeafWarps - 1) / kLeafWarps;
    const int img = (int)(blockIdx.x / blocks_per_img);
    const int64_t grp = (int64_t)(blockIdx.x - (int64_t)img * blocks_per_img) * kLeafWarps + warp;
    TIn *my = reinterpret_cast<TIn *>(leaf_raw) + (size_t)warp * 128 * B;
    uint8_t *myc = leaf_raw + (size_t)kLeafWarps * 128 * B * sizeof(TIn) + (size_t)warp * 128;
    const int sub = lane >> 3, j = lane & 7;
    int cnt[kLeafMaxBands / 4];
#pragma unroll
    for (int q = 0; q < kLeafMaxBands / 4; ++q) cnt[q] = 0;
    for (int k = 0; k < kLeafPerWarp; ++k) {
        const int64_t c = grp * kLeafPerWarp + k;
        if (grp >= groups_per_img || c >= n_chunks) break;                // warp-uniform
        int64_t lo;
        int m;
        leaf_of(n, min(64 * c + 63, n - 1), lo, m);
        if (lo < 64 * c || lo < own_cell0 || lo >= own_cell1) continue;   // no leaf starts here / not ours
        const int64_t l0 = lo - cell_offset;                              // local cell index
        if (l0 < 0 || l0 + m > cells_local) continue;                     // the host checked that the halo covers it
        const TIn *src = target + ((int64_t)img * cells_local + l0) * B;
        __syncwarp();
        warp_stage(my, src, m * B, lane);
        if (cloud)
            for (int i = lane; i < m; i += 32) myc[i] = cloud[(int64_t)img * cells_local + l0 + i];
        __syncwarp();
        const int m8 = m - (m % 8);
#pragma unroll
        for (int q = 0; q < kLeafMaxBands / 4; ++q) {
            if (q * 4 >= B) break;                                        // warp-uniform
            const int b = q * 4 + sub;
            const bool live = b < B;
            TAcc r = (TAcc)0;
            int cn = 0;
            if (live) {
                for (int i = j; i < m8; i += 8) {
                    TAcc v = conv_in<TAcc>(my[i * B + b], clip, lo_v, hi_v);
                    const bool unk = (v != v) || (cloud && myc[i]);
                    cn += unk ? 0 : 1;
                    v = unk ? (TAcc)0 : v;
                    r = (i == j) ? v : r + v;
                }
            }
            // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) within each group of 8 lanes
            const TAcc a1 = r + __shfl_down_sync(0xffffffffu, r, 1, 8);
            const TAcc a2 = a1 + __shfl_down_sync(0xffffffffu, a1, 2, 8);
            TAcc res = a2 + __shfl_down_sync(0xffffffffu, a2, 4, 8);
            if (live && j == 0) {
                for (int i = m8; i < m; ++i) {                            // the tail of the grid's last leaf
                    TAcc v = conv_in<TAcc>(my[i * B + b], clip, lo_v, hi_v);
                    const bool unk = (v != v) || (cloud && myc[i]);
                    cn += unk ? 0 : 1;
                    res += unk ? (TAcc)0 : v;
                }
                leaf_sum[((size_t)img * B + b) * n_chunks + c] = res;
            }
            cnt[q] += cn;
        }
    }
#pragma unroll
    for (int q = 0; q < kLeafMaxBands / 4; ++q) {
        if (q * 4 >= B) break;
        if (q * 4 + sub < B && cnt[q]) atomicAdd(&cnt_s[q * 4 + sub], (unsigned)cnt[q]);
    }
    __syncthreads();
    if (threadIdx.x < B && cnt_s[threadIdx.x]) atomicAdd(&count[(size_t)img * B + threadIdx.x], (unsigned long long)cnt_s[threadIdx.x]);
}

// Sum of the subtree that covers cells [a, a + len) from the leaf sums, in NumPy's recursion order (explicit stack).
template <typename TAcc>
__device__ TAcc subtree_sum(const TAcc *__restrict__ leaf_sum, int64_t a0, int64_t len0) {
    int64_t st_a[48], st_len[48];
    TAcc st_acc[48];
    int st_state[48];
    int sp = 0;
    st_a[0] = a0; st_len[0] = len0; st_state[0] = 0;
    TAcc ret = (TAcc)0;
    while (sp >= 0) {
        const int64_t a = st_a[sp], len = st_len[sp];
        if (len <= 128) { ret = leaf_sum[a >> 6]; --sp; continue; }
        int64_t h = len / 2;
        h -= h % 8;
        if (st_state[sp] == 0) { st_state[sp] = 1; ++sp; st_a[sp] = a; st_len[sp] = h; st_state[sp] = 0; }
        else if (st_state[sp] == 1) { st_acc[sp] = ret; st_state[sp] = 2; ++sp; st_a[sp] = a + h; st_len[sp] = len - h; st_state[sp] = 0; }
        else { ret = st_acc[sp] + ret; --sp; }
    }
    return ret;
}

// The tree is cut at depth D (2^D subtrees, each far above leaf size).  fill_subtree_kernel: one thread per subtree
// sums its leaves sequentially in recursion order.  fill_top_kernel: one block per band problem adds the 2^D subtree
// sums level by level -- above depth D every range still splits, so that part of the tree is a complete binary tree
// and node i of a level is node 2i + node 2i+1 of the level below -- and forms
// fill = dtype(double(sum) / count), as np.nanmean does for a float32 grid (float64 division, result cast back).
constexpr int kFillMaxDepth = 14;

template <typename TAcc>
__global__ void __launch_bounds__(128) fill_subtree_kernel(const TAcc *__restrict__ leaf_sum, int64_t n, int64_t n_chunks, int D,
                                                           TAcc *__restrict__ part) {
    const int p = blockIdx.x;
    const int t = blockIdx.y * blockDim.x + threadIdx.x;
    if (t >= (1 << D)) return;
    int64_t a = 0, len = n;
    for (int d = D - 1; d >= 0; --d) {
        int64_t h = len / 2;
        h -= h % 8;
        if ((t >> d) & 1) { a += h; len -= h; }
        else len = h;
    }
    part[((size_t)p << D) + t] = subtree_sum<TAcc>(leaf_sum + (size_t)p * n_chunks, a, len);
}

template <typename TAcc>
__global__ void __launch_bounds__(1024) fill_top_kernel(TAcc *__restrict__ part, const unsigned long long *__restrict__ count,
                                                        int D, double *__restrict__ fill) {
    const int p = blockIdx.x;
    TAcc *v = part + ((size_t)p << D);
    for (int lvl = D - 1; lvl >= 0; --lvl) {
        const int nodes = 1 << lvl;
        // node i <- node 2i + node 2i+1: read both children before anyone overwrites them
        for (int i0 = 0; i0 < nodes; i0 += blockDim.x) {
            const int i = i0 + threadIdx.x;
            TAcc sum = (TAcc)0;
            if (i < nodes) sum = v[2 * i] + v[2 * i + 1];
            __syncthreads();
            if (i < nodes) v[i] = sum;
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        const TAcc total = (TAcc)0 + v[0];                               // add.reduce starts from the identity
        fill[p] = (double)(TAcc)((double)total / (double)count[p]);
    }
}

// ---- the fused load ------------------------------------------------------------------------------------
// One warp per (image, row): the row of all bands is staged once with coalesced loads (first the target, then the
// features reuse the same shared memory) and every band's plane row is written from there.
//   bufA   [P][H][pitch] fp64 start state: known cells, unknown cells = fill[p]        (MRP.py:64-91)
//   bufB   the same (second ping-pong buffer of the tiled kernels); skipped when null
//   mask   1 bit per cell, set = unknown (cloud on every band, or NaN in this band)
//   fplane features with zeros -> 0.01 (WP_MRP.py:143-145) + the flags fplane_kernel keeps
//   rowpart known-cell terms of the stopping rule, as prepare_bands_kernel
//   out    known cells of the result in the caller's layout / dtype (null: none)
constexpr int kIngestWarps = 4;
constexpr int kIngestMaxBands = 32;
constexpr int kIngestTile = 256;     // pixels per staged tile (a multiple of 32: mask words)

struct IngestArgs {
    Geom g;
    const void *target, *features;
    const uint8_t *cloud;
    int clip_t, clip_f;
    double t_lo, t_hi, f_lo, f_hi;
    const double *fill;
    double *bufA, *bufB, *fplane, *rowpart, *extreme;
    uint32_t *mask;
    int *flags;
    void *out;
    int out_code;            // VPINT_F64 / VPINT_F32 / VPINT_U16
    Strides osd;
    int TW;
    int XC;                  // column chunks per row (one warp each)
    double *rowchunk;        // [P][H][XC][3] chunk partials when XC > 1
    double *save;            // ratio-direct load: [P][H][pitch] the start state in value space (bufA / bufB then hold value / f)
    float *f32plane;         // ratio-direct load: the feature plane once more as float32 (exact unless flags[3] gets set)
};

// rowpart of a row = its chunk partials added in chunk order (fixed order: the known-cell sums are reproducible).
__global__ void __launch_bounds__(256) ingest_rowsum_kernel(int64_t n_rows_all, int XC, const double *__restrict__ rowchunk,
                                                            double *__restrict__ rowpart) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows_all) return;
    double s = 0.0, ninf = 0.0, nunk = 0.0;
    for (int c = 0; c < XC; ++c) {
        const double *rc = rowchunk + ((size_t)i * XC + c) * 3;
        s += rc[0]; ninf += rc[1]; nunk += rc[2];
    }
    double *rp = rowpart + (size_t)i * 8;
    rp[0] = 0.0; rp[1] = s; rp[2] = 0.0; rp[3] = 0.0; rp[4] = s; rp[5] = ninf; rp[6] = 0.0; rp[7] = nunk;
}

// DIRECT: the tiled ratio stencil will run this solve (a scene, one feature layer), so the state is written in ratio space
// at once -- bufA = bufB = value / feature, `save` = value -- instead of a separate pass over both buffers before the first
// sweep (to_ratio_kernel).  Target and feature tiles are then staged side by side.
template <typename TIn, typename TAcc, bool DIRECT>
__global__ void __launch_bounds__(kIngestWarps * 32) ingest_bands_kernel(const IngestArgs a) {
    extern __shared__ __align__(16) unsigned char ing_raw[];
    __shared__ double racc[kIngestWarps][kIngestMaxBands][2];           // per band: sum of known values, packed counts
    const Geom &g = a.g;
    const int B = g.Pin, W = g.L, TW = a.TW, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // a warp takes one of XC column chunks of one row (long rows of a scene are split so that there are enough warps)
    const int64_t wid = (int64_t)blockIdx.x * kIngestWarps + warp;
    const int64_t n_rows = (int64_t)(g.P / B) * g.H;
    const int XC = a.XC;
    if (wid >= n_rows * XC) return;
    const int64_t r = wid / XC;
    const int xc = (int)(wid - r * XC);
    const int po = (int)(r / g.H), y = (int)(r - (int64_t)po * g.H);
    const int tiles = (g.pitch + TW - 1) / TW, per = (tiles + XC - 1) / XC;
    const int x_begin = xc * per * TW, x_end = min(g.pitch, (xc + 1) * per * TW);
    TIn *my = reinterpret_cast<TIn *>(ing_raw) + (size_t)warp * TW * B;
    TIn *myf = DIRECT ? reinterpret_cast<TIn *>(ing_raw) + (size_t)(kIngestWarps + warp) * TW * B : my;   // features beside the target
    uint8_t *myc = ing_raw + (size_t)(DIRECT ? 2 : 1) * kIngestWarps * TW * B * sizeof(TIn) + (size_t)warp * TW;
    const int64_t row_elems = (int64_t)W * B;
    const TIn *trow = reinterpret_cast<const TIn *>(a.target) + ((int64_t)po * g.H + y) * row_elems;
    const TIn *frow = reinterpret_cast<const TIn *>(a.features) + ((int64_t)po * g.H + y) * row_elems;
    const bool owned = (y >= g.own_row0) && (y < g.own_row0 + g.own_rows);
    const bool out_interleaved = a.out && a.osd.sI == 1 && a.osd.sX == B;
    bool wild = false, bad = false, fwild = false, wide = false;
    for (int b = lane; b < B; b += 32) { racc[warp][b][0] = 0.0; racc[warp][b][1] = 0.0; }
    // the row goes through shared memory in tiles of TW pixels (one tile for the 256-pixel patches)
    for (int x0 = x_begin; x0 < x_end; x0 += TW) {
        const int tw = max(0, min(TW, W - x0));                          // pixels of this tile inside the grid
        const int te = min(TW, g.pitch - x0);                            // plane elements of this tile (incl. padding)
        __syncwarp();
        if (tw > 0) warp_stage(my, trow + (int64_t)x0 * B, tw * B, lane);
        if (DIRECT && tw > 0) warp_stage(myf, frow + (int64_t)x0 * B, tw * B, lane);
        for (int x = lane; x < tw; x += 32) myc[x] = a.cloud ? a.cloud[((int64_t)po * g.H + y) * W + x0 + x] : 0;
        __syncwarp();
        // known cells of the result, band-interleaved output: the staged tile has the output's own order
        if (out_interleaved && owned) {
            const int64_t obase = (int64_t)po * a.osd.sO + (int64_t)y * a.osd.sY + (int64_t)x0 * B;
            for (int i = lane; i < tw * B; i += 32) {
                const int x = i / B;
                const TAcc v = conv_in<TAcc>(my[i], a.clip_t, a.t_lo, a.t_hi);
                if (!((v != v) || myc[x])) store_out(a.out, a.out_code, obase + i, (double)v);
            }
        }
        for (int b = 0; b < B; ++b) {
            const int p = po * B + b;
            const size_t row = ((size_t)p * g.H + y) * g.pitch + x0;
            const size_t mrow = ((size_t)p * g.H + y) * g.mpitch + (x0 >> 5);
            const double fillv = a.fill[p];
            const int64_t obase = (int64_t)po * a.osd.sO + (int64_t)b * a.osd.sI + (int64_t)y * a.osd.sY;
            double sum_o = 0.0;
            int counts = 0;                                             // #inf << 16 | #unknown
            for (int e = lane; e < te; e += 32) {
                bool unknown = false;
                double va = 0.0, fe = 1.0;
                if (e < tw) {
                    const double o = (double)conv_in<TAcc>(my[e * B + b], a.clip_t, a.t_lo, a.t_hi);
                    unknown = (o != o) || myc[e];
                    va = unknown ? fillv : o;
                    if (!(fabs(va) < 1e150)) wild = true;
                    if (DIRECT) {
                        fe = (double)conv_in<TAcc>(myf[e * B + b], a.clip_f, a.f_lo, a.f_hi);
                        if (fe == 0.0) fe = 0.01;                 // VPint/WP_MRP.py:143-145
                        if (!(fabs(fe) <= 1.7976931348623157e308)) bad = true;
                        if (!(fabs(fe) < 1e150) || !(fabs(fe) > 1e-150)) fwild = true;
                        if ((double)(float)fe != fe) wide = true;
                    }
                    if (owned) {
                        if (unknown) counts += 1;
                        else {
                            sum_o += o;
                            if (isinf(o)) counts += 1 << 16;
                            if (a.out && !out_interleaved) store_out(a.out, a.out_code, obase + (int64_t)(x0 + e) * a.osd.sX, o);
                        }
                    }
                }
                if (DIRECT) {
                    const double ra = va / fe;                  // to_ratio_kernel's quotient, formed here
                    a.bufA[row + e] = ra;
                    a.bufB[row + e] = ra;
                    a.save[row + e] = va;
                    a.fplane[row + e] = fe;
                    a.f32plane[row + e] = (float)fe;
                } else {
                    a.bufA[row + e] = va;
                    if (a.bufB) a.bufB[row + e] = va;
                }
                const unsigned word = __ballot_sync(0xffffffffu, unknown);
                if (lane == 0) a.mask[mrow + (e >> 5)] = word;
            }
            sum_o = warp_sum(sum_o);
            counts = warp_sum(counts);
            if (lane == 0) {
                racc[warp][b][0] += sum_o;
                racc[warp][b][1] += (double)counts;                     // exact: both fields stay far below 2^53
            }
        }
        __syncwarp();
        if (DIRECT) continue;                                   // the feature plane was written with the state
        // features of the same tile reuse the staging area
        if (tw > 0) warp_stage(my, frow + (int64_t)x0 * B, tw * B, lane);
        __syncwarp();
        for (int b = 0; b < B; ++b) {
            double *out = a.fplane + ((size_t)(po * B + b) * g.H + y) * g.pitch + x0;
            for (int x = lane; x < te; x += 32) {
                double f = 1.0;
                if (x < tw) {
                    f = (double)conv_in<TAcc>(my[x * B + b], a.clip_f, a.f_lo, a.f_hi);
                    if (f == 0.0) f = 0.01;                 // VPint/WP_MRP.py:143-145
                    if (!(fabs(f) <= 1.7976931348623157e308)) bad = true;
                    if (!(fabs(f) < 1e150) || !(fabs(f) > 1e-150)) fwild = true;
                    if ((double)(float)f != f) wide = true;  // not a float32 value (0.01 is not): fp64 on chip
                }
                out[x] = f;
            }
        }
    }
    __syncwarp();
    for (int b = lane; b < B; b += 32) {
        const double sum_o = racc[warp][b][0];
        const long long counts = (long long)racc[warp][b][1];
        if (XC == 1) {
            double *rp = a.rowpart + ((size_t)(po * B + b) * g.H + y) * 8;
            rp[0] = 0.0; rp[1] = sum_o; rp[2] = 0.0; rp[3] = 0.0;
            rp[4] = sum_o; rp[5] = (double)(counts >> 16); rp[6] = 0.0; rp[7] = (double)(counts & 0xffff);
        } else {                                                        // chunk partials, added up in order by ingest_rowsum_kernel
            double *rc = a.rowchunk + (((size_t)(po * B + b) * g.H + y) * XC + xc) * 3;
            rc[0] = sum_o; rc[1] = (double)(counts >> 16); rc[2] = (double)(counts & 0xffff);
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicExch(a.flags, 1);
    if (__any_sync(0xffffffffu, wide) && lane == 0) atomicExch(a.flags + 3, 1);
    if (__any_sync(0xffffffffu, wild) && lane == 0) atomicAdd(a.extreme, 1.0);
    if (__any_sync(0xffffffffu, fwild) && lane == 0) atomicAdd(a.extreme + 1, 1.0);
}

// ---- unknown cells of the result -------------------------------------------------------------------------
// One warp per (image, row) of a band-interleaved stack (or per (problem, row) otherwise).  Rows without an unknown
// cell are skipped after reading their mask words; the others write every unknown cell of every band, lanes running
// over the output row in ITS memory order, so that a run of cloudy pixels leaves as one contiguous run of stores.
constexpr int kUnkWarps = 8;

template <typename TOut>
__global__ void __launch_bounds__(kUnkWarps * 32) result_unknown_kernel(Geom g, const ProbState *__restrict__ state,
                                                                        const double *__restrict__ buf0,
                                                                        const double *__restrict__ buf1,
                                                                        const uint32_t *__restrict__ mask, TOut *__restrict__ out,
                                                                        Strides sd, int interleaved, int32_t *niter_out) {
    const int lane = threadIdx.x & 31;
    const int B = interleaved ? g.Pin : 1;
    const int64_t n_rows = (int64_t)(g.P / B) * g.H;
    // grid-stride over the rows: the launcher caps the grid when `out` is host memory (the kernel is then bound by the
    // host link, and a small grid keeps it from taking SMs the solve kernels of other streams could use)
    for (int64_t r = (int64_t)blockIdx.x * kUnkWarps + (threadIdx.x >> 5); r < n_rows; r += (int64_t)gridDim.x * kUnkWarps) {
    const int po = (int)(r / g.H), y = (int)(r - (int64_t)po * g.H);
    if (niter_out && y == 0)
        for (int b = lane; b < B; b += 32) niter_out[po * B + b] = state[po * B + b].iters_done;
    if (y < g.own_row0 || y >= g.own_row0 + g.own_rows) continue;
    const size_t plane = (size_t)g.H * g.pitch;
    if (!interleaved) {
        const int p = po;
        const uint32_t *mr = mask + ((size_t)p * g.H + y) * g.mpitch;
        const double *src = (state[p].cur ? buf1 : buf0) + (size_t)p * plane + (size_t)y * g.pitch;
        const int pi = p % g.Pin, pq = p / g.Pin;
        TOut *dst = out + (int64_t)pq * sd.sO + (int64_t)pi * sd.sI + (int64_t)y * sd.sY;
        for (int x0 = 0; x0 < g.L; x0 += 32) {
            const unsigned word = mr[x0 >> 5];
            if (word == 0) continue;
            const int x = x0 + lane;
            if ((word >> lane) & 1u) dst[(int64_t)x * sd.sX] = to_out<TOut>(src[x]);
        }
        continue;
    }
    // any unknown cell in this row, in any band?
    unsigned any = 0;
    const int words = (g.L + 31) >> 5;
    for (int k = lane; k < B * words; k += 32) {
        const int b = k / words, w = k - b * words;
        any |= mask[((size_t)(po * B + b) * g.H + y) * g.mpitch + w];
    }
    if (!__any_sync(0xffffffffu, any != 0)) continue;
    TOut *dst = out + (int64_t)po * sd.sO + (int64_t)y * sd.sY;
    const int total = g.L * B;
    for (int i0 = 0; i0 < total; i0 += 32) {
        const int i = i0 + lane;
        if (i < total) {
            const int x = i / B, b = i - x * B;
            const int p = po * B + b;
            const unsigned word = mask[((size_t)p * g.H + y) * g.mpitch + (x >> 5)];
            if ((word >> (x & 31)) & 1u) {
                const double *src = (state[p].cur ? buf1 : buf0) + (size_t)p * plane + (size_t)y * g.pitch;
                dst[i] = to_out<TOut>(src[x]);
            }
        }
    }
    }
}

}  // namespace

#define VP_LAUNCH_CHECK()                         \
    do {                                          \
        cudaError_t e__ = cudaGetLastError();     \
        if (e__ != cudaSuccess) return e__;       \
        if (launches) ++*launches;                \
    } while (0)

cudaError_t launch_cloud_mask(const void *mask, int dtype, double threshold, int64_t n, uint8_t *cloud, cudaStream_t st) {
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (n <= 0) return cudaSuccess;
    switch (dtype) {
        case VPINT_F64: cloud_mask_kernel<double><<<grid, 256, 0, st>>>((const double *)mask, threshold, n, cloud); break;
        case VPINT_F32: cloud_mask_kernel<float><<<grid, 256, 0, st>>>((const float *)mask, threshold, n, cloud); break;
        case VPINT_U16: cloud_mask_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t *)mask, threshold, n, cloud); break;
        case VPINT_U8: cloud_mask_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)mask, threshold, n, cloud); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// cloud (in/out): all-band mask [n_img][H][W]; scratch: 3 * n_img * H * W bytes.
cudaError_t launch_buffer_clouds(const void *target, int dtype, uint8_t *cloud, bool have_cloud, int n_img, int H, int W, int B,
                                 int buffer_size, int passes, uint8_t *scratch, cudaStream_t st) {
    const int64_t n_px = (int64_t)n_img * H * W;
    if (n_px <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n_px + 255) / 256);
    uint8_t *foot0 = scratch, *cur = scratch + n_px, *tall = scratch + 2 * n_px;
    const uint8_t *cloud0 = have_cloud ? cloud : nullptr;
    switch (dtype) {
        case VPINT_F64: footprint_kernel<double><<<grid, 256, 0, st>>>((const double *)target, cloud0, n_px, B, foot0); break;
        case VPINT_F32: footprint_kernel<float><<<grid, 256, 0, st>>>((const float *)target, cloud0, n_px, B, foot0); break;
        case VPINT_U16:
            if (have_cloud) {
                cudaError_t e = cudaMemcpyAsync(foot0, cloud, (size_t)n_px, cudaMemcpyDeviceToDevice, st);
                if (e != cudaSuccess) return e;
            } else {
                cudaError_t e = cudaMemsetAsync(foot0, 0, (size_t)n_px, st);
                if (e != cudaSuccess) return e;
            }
            break;
        default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (!have_cloud) {
        e = cudaMemsetAsync(cloud, 0, (size_t)n_px, st);
        if (e != cudaSuccess) return e;
    }
    if (buffer_size <= 0 || passes <= 0) return cudaSuccess;
    // pass k: grown_k = dilate(footprint_{k-1}); footprint_k = foot0 | grown_k (every pass starts again from the
    // original image, VPint2.py:138); the all-band mask after the last pass is cloud0 | grown_last
    const int col_limit = (H < W) ? H : W;
    const uint8_t *src = foot0;
    for (int k = 0; k < passes; ++k) {
        dilate_rows_kernel<<<grid, 256, 0, st>>>(src, tall, H, W, buffer_size, n_px);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        const bool last = (k == passes - 1);
        dilate_cols_kernel<<<grid, 256, 0, st>>>(tall, foot0, cloud0, cur, last ? cloud : nullptr, H, W, buffer_size, col_limit,
                                                 n_px);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        src = cur;
    }
    return cudaSuccess;
}

// scratch: [leaf sums: n_prob x n_chunks][known-cell counts: n_prob x uint64][subtree sums: n_prob x 2^kFillMaxDepth]
size_t fill_general_scratch_bytes(int64_t n, int n_prob, int acc_dtype) {
    const int64_t n_chunks = (n + 63) / 64;
    const size_t esz = (acc_dtype == VPINT_F64) ? 8 : 4;
    const size_t leaf_bytes = ((size_t)n_prob * n_chunks * esz + 255) / 256 * 256;
    const size_t count_bytes = ((size_t)n_prob * sizeof(unsigned long long) + 255) / 256 * 256;
    return leaf_bytes + count_bytes + ((size_t)n_prob << kFillMaxDepth) * esz + 256;
}

// Leaf sums of the leaves this shard owns (zeros elsewhere) + known-cell counts.  leaf_sum / count live in `scratch`.
cudaError_t launch_fill_leaves(const void *target, int t_dtype, int acc_dtype, const uint8_t *cloud, int n_img, int B, int64_t n,
                               int64_t cell_offset, int64_t cells_local, int64_t own_cell0, int64_t own_cell1, int clip,
                               double lo, double hi, void *scratch, cudaStream_t st, int64_t *launches) {
    if (B < 1 || B > kLeafMaxBands || n < 128) return cudaErrorInvalidValue;
    const int64_t n_chunks = (n + 63) / 64;
    const int n_prob = n_img * B;
    const size_t esz = (acc_dtype == VPINT_F64) ? 8 : 4;
    const size_t leaf_bytes = ((size_t)n_prob * n_chunks * esz + 255) / 256 * 256;
    cudaError_t e = cudaMemsetAsync(scratch, 0, leaf_bytes + (size_t)n_prob * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    unsigned long long *count = reinterpret_cast<unsigned long long *>(static_cast<char *>(scratch) + leaf_bytes);
    const size_t tsz = (t_dtype == VPINT_F64) ? 8 : (t_dtype == VPINT_F32 ? 4 : 2);
    const size_t smem = (size_t)kLeafWarps * 128 * B * tsz + (size_t)kLeafWarps * 128;
    const int64_t groups_per_img = (n_chunks + kLeafPerWarp - 1) / kLeafPerWarp;
    const int64_t blocks_per_img = (groups_per_img + kLeafWarps - 1) / kLeafWarps;
    if (blocks_per_img * n_img >= ((int64_t)1 << 31)) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)(blocks_per_img * n_img);
#define VP_LEAF(TIN, TACC)                                                                                              \
    do {                                                                                                                \
        e = cudaFuncSetAttribute(fill_leaf_kernel<TIN, TACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
        if (e != cudaSuccess) return e;                                                                                 \
        fill_leaf_kernel<TIN, TACC><<<grid, kLeafWarps * 32, smem, st>>>(                                               \
            (const TIN *)target, cloud, n_img, B, n, n_chunks, groups_per_img, cell_offset, cells_local, own_cell0,     \
            own_cell1, clip, lo, hi,                                                                                    \
            (TACC *)scratch, count);                                                                                    \
    } while (0)
    if (t_dtype == VPINT_U16 && acc_dtype == VPINT_F32) VP_LEAF(uint16_t, float);
    else if (t_dtype == VPINT_U16 && acc_dtype == VPINT_F64) VP_LEAF(uint16_t, double);
    else if (t_dtype == VPINT_F32 && acc_dtype == VPINT_F32) VP_LEAF(float, float);
    else if (t_dtype == VPINT_F32 && acc_dtype == VPINT_F64) VP_LEAF(float, double);
    else if (t_dtype == VPINT_F64 && acc_dtype == VPINT_F64) VP_LEAF(double, double);
    else if (t_dtype == VPINT_F64 && acc_dtype == VPINT_F32) VP_LEAF(double, float);
    else return cudaErrorInvalidValue;
#undef VP_LEAF
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_fill_combine(int acc_dtype, int n_prob, int64_t n, void *scratch, double *fill_dev, cudaStream_t st,
                                int64_t *launches) {
    const int64_t n_chunks = (n + 63) / 64;
    const size_t esz = (acc_dtype == VPINT_F64) ? 8 : 4;
    const size_t leaf_bytes = ((size_t)n_prob * n_chunks * esz + 255) / 256 * 256;
    const unsigned long long *count = reinterpret_cast<const unsigned long long *>(static_cast<char *>(scratch) + leaf_bytes);
    const size_t count_bytes = ((size_t)n_prob * sizeof(unsigned long long) + 255) / 256 * 256;
    void *part = static_cast<char *>(scratch) + leaf_bytes + count_bytes;
    int D = 0;
    while (D < kFillMaxDepth && (n >> (D + 1)) >= 2048) ++D;        // subtrees of >= 2048 cells: far above leaf size
    const dim3 grid((unsigned)n_prob, (unsigned)(((1 << D) + 127) / 128));
    if (acc_dtype == VPINT_F64) {
        fill_subtree_kernel<double><<<grid, 128, 0, st>>>((const double *)scratch, n, n_chunks, D, (double *)part);
        VP_LAUNCH_CHECK();
        fill_top_kernel<double><<<n_prob, 1024, 0, st>>>((double *)part, count, D, fill_dev);
    } else {
        fill_subtree_kernel<float><<<grid, 128, 0, st>>>((const float *)scratch, n, n_chunks, D, (float *)part);
        VP_LAUNCH_CHECK();
        fill_top_kernel<float><<<n_prob, 1024, 0, st>>>((float *)part, count, D, fill_dev);
    }
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

// Column chunks per row: enough warps to fill the machine (a scene shard has ~10^3 long rows), one chunk for patches.
int ingest_row_chunks(const Geom &g) {
    const int TW = g.pitch < kIngestTile ? g.pitch : kIngestTile;
    const int tiles = (g.pitch + TW - 1) / TW;
    const int64_t n_rows = (int64_t)(g.P / g.Pin) * g.H;
    int64_t want = (16384 + n_rows - 1) / n_rows;
    if (want > tiles) want = tiles;
    if (want > 16) want = 16;
    return want < 1 ? 1 : (int)want;
}

bool ingest_direct_supported(const Geom &g, int t_dtype) {
    const size_t tsz = (t_dtype == VPINT_F64) ? 8 : (t_dtype == VPINT_F32 ? 4 : 2);
    const int TW = g.pitch < kIngestTile ? g.pitch : kIngestTile;
    return 2 * (size_t)kIngestWarps * TW * g.Pin * tsz + (size_t)kIngestWarps * TW <= 200 * 1024;
}

bool ingest_bands_supported(const Geom &g, int t_dtype) {
    const size_t tsz = (t_dtype == VPINT_F64) ? 8 : (t_dtype == VPINT_F32 ? 4 : 2);
    const int TW = g.pitch < kIngestTile ? g.pitch : kIngestTile;
    const size_t smem = (size_t)kIngestWarps * TW * g.Pin * tsz + (size_t)kIngestWarps * TW;
    return g.ndim == 2 && g.Pin >= 1 && g.Pin <= kIngestMaxBands && g.L < 65536 && smem <= 200 * 1024;
}

cudaError_t launch_ingest_bands(const Geom &g, const void *target, const void *features, int t_dtype, int acc_dtype,
                                const uint8_t *cloud, int clip_t, double t_lo, double t_hi, int clip_f, double f_lo,
                                double f_hi, const double *fill, double *bufA, double *bufB, uint32_t *mask, double *fplane,
                                double *rowpart, double *extreme, int *flags, void *out, int out_dtype,
                                const int64_t *out_stride, double *rowchunk, double *save, float *f32plane, cudaStream_t st,
                                int64_t *launches) {
    IngestArgs a;
    a.save = save;
    a.f32plane = f32plane;
    a.g = g;
    a.target = target; a.features = features; a.cloud = cloud;
    a.clip_t = clip_t; a.clip_f = clip_f; a.t_lo = t_lo; a.t_hi = t_hi; a.f_lo = f_lo; a.f_hi = f_hi;
    a.fill = fill; a.bufA = bufA; a.bufB = bufB; a.fplane = fplane; a.rowpart = rowpart; a.extreme = extreme;
    a.mask = mask; a.flags = flags;
    a.out = out; a.out_code = out_dtype;
    a.osd = out ? Strides{out_stride[0], out_stride[1], out_stride[2], out_stride[3]} : Strides{0, 0, 0, 0};
    const size_t tsz = (t_dtype == VPINT_F64) ? 8 : (t_dtype == VPINT_F32 ? 4 : 2);
    const bool direct = save != nullptr;
    // the ratio-direct load stages target AND features: half the tile keeps twice the blocks per SM resident
    const int tile = (direct && !getenv("VPINT_INGEST_TILE256")) ? kIngestTile / 2 : kIngestTile;
    a.TW = g.pitch < tile ? g.pitch : tile;
    const size_t smem = (size_t)(direct ? 2 : 1) * kIngestWarps * a.TW * g.Pin * tsz + (size_t)kIngestWarps * a.TW;
    const int64_t n_rows = (int64_t)(g.P / g.Pin) * g.H;
    a.XC = ingest_row_chunks(g);
    a.rowchunk = rowchunk;
    if (a.XC > 1 && !rowchunk) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)((n_rows * a.XC + kIngestWarps - 1) / kIngestWarps);
    cudaError_t e;
#define VP_ING1(TIN, TACC, DIR)                                                                                               \
    do {                                                                                                                      \
        e = cudaFuncSetAttribute(ingest_bands_kernel<TIN, TACC, DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                                                       \
        ingest_bands_kernel<TIN, TACC, DIR><<<grid, kIngestWarps * 32, smem, st>>>(a);                                        \
    } while (0)
#define VP_ING(TIN, TACC)                    \
    do {                                     \
        if (direct) VP_ING1(TIN, TACC, true); \
        else VP_ING1(TIN, TACC, false);      \
    } while (0)
    if (t_dtype == VPINT_U16 && acc_dtype == VPINT_F32) VP_ING(uint16_t, float);
    else if (t_dtype == VPINT_U16 && acc_dtype == VPINT_F64) VP_ING(uint16_t, double);
    else if (t_dtype == VPINT_F32 && acc_dtype == VPINT_F32) VP_ING(float, float);
    else if (t_dtype == VPINT_F32 && acc_dtype == VPINT_F64) VP_ING(float, double);
    else if (t_dtype == VPINT_F64 && acc_dtype == VPINT_F64) VP_ING(double, double);
    else if (t_dtype == VPINT_F64 && acc_dtype == VPINT_F32) VP_ING(double, float);
    else return cudaErrorInvalidValue;
#undef VP_ING
#undef VP_ING1
    VP_LAUNCH_CHECK();
    if (a.XC > 1) {
        const int64_t n_all = (int64_t)g.P * g.H;
        ingest_rowsum_kernel<<<(unsigned)((n_all + 255) / 256), 256, 0, st>>>(n_all, a.XC, rowchunk, rowpart);
        VP_LAUNCH_CHECK();
    }
    return cudaSuccess;
}

cudaError_t launch_result_unknown(const Geom &g, const ProbState *state, void *const *buf, const uint32_t *mask, void *out,
                                  int dtype, const int64_t *stride, int32_t *niter_out, int max_blocks, cudaStream_t st,
                                  int64_t *launches) {
    const Strides sd{stride[0], stride[1], stride[2], stride[3]};
    const int interleaved = (g.Pin >= 2 && sd.sI == 1 && sd.sX == g.Pin) ? 1 : 0;
    const int64_t n_rows = (int64_t)(interleaved ? g.P / g.Pin : g.P) * g.H;
    unsigned grid = (unsigned)((n_rows + kUnkWarps - 1) / kUnkWarps);
    if (max_blocks > 0 && grid > (unsigned)max_blocks) grid = (unsigned)max_blocks;
    if (dtype == VPINT_F64)
        result_unknown_kernel<double><<<grid, kUnkWarps * 32, 0, st>>>(g, state, (const double *)buf[0], (const double *)buf[1],
                                                                       mask, (double *)out, sd, interleaved, niter_out);
    else if (dtype == VPINT_U16)
        result_unknown_kernel<uint16_t><<<grid, kUnkWarps * 32, 0, st>>>(g, state, (const double *)buf[0], (const double *)buf[1],
                                                                         mask, (uint16_t *)out, sd, interleaved, niter_out);
    else
        result_unknown_kernel<float><<<grid, kUnkWarps * 32, 0, st>>>(g, state, (const double *)buf[0], (const double *)buf[1],
                                                                      mask, (float *)out, sd, interleaved, niter_out);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

}  // namespace vpint
