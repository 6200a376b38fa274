// A4 + A5 (local part): one Jacobi sweep per launch straight from global memory ("k1" kernels).
//   new = (sum_dir w_dir * P[nb_dir]) / neighbour_count, cells outside the grid read as 0
//   (VPint/WP_MRP.py:291-319, VPint/SD_MRP.py:112-132, 357-385); clip before the restore
//   (WP_MRP.py:322); only unknown cells are written, known cells were materialised once at load.
//   Per tile: {sum |new-old|, sum old, NaN counts} over unknown owned cells (VPint/SD_MRP.py:139).
// These kernels serve k_block == 1 and the 3-D path; the temporally blocked 2-D kernel lives in
// vpint_step_tb.cu.
#include <cstdlib>

#include "vpint_internal.cuh"

namespace vpint {

namespace {

constexpr int kTileW2 = 128, kTileH2 = 32;   // 2-D k1 tile: 8 warps as 2 (columns) x 4 (rows), 2 cells x 8 rows per thread
constexpr int kTileW3 = 256, kTileH3 = 8;    // 3-D k1 tile: 256 row-elements x 8 rows, 1 element x 8 rows per thread

template <typename T>
struct Acc {
    double s_abs = 0.0, s_old = 0.0;
    int c_abs = 0, c_old = 0;
    __device__ __forceinline__ T commit(double nv, double old) {
        const T stored = (T)nv;
        const double d = fabs((double)stored - old);
        if (d != d) ++c_abs; else s_abs += d;
        if (old != old) ++c_old; else s_old += old;
        return stored;
    }
};

// Block sum of the tile's statistics -> its partial; then (fused mode) the problem's finish.
template <typename T>
__device__ __forceinline__ void write_partial(Acc<T> &a, const StepArgs &args, int problem) {
    __shared__ double scratch[kSumSlots * 8];
    double v[kSumSlots] = {a.s_abs, a.s_old, (double)a.c_abs, (double)a.c_old};
    block_sum<kSumSlots>(v, scratch);
    if (threadIdx.x == 0) {
        double *partial_tile = args.partial + (size_t)blockIdx.x * args.g.KB * kSumSlots;
#pragma unroll
        for (int i = 0; i < kSumSlots; ++i) partial_tile[i] = v[i];
    }
    finish_tile(args, problem, 1, scratch);
}

template <typename T, bool PLANES>
__global__ void __launch_bounds__(256) jacobi2d_k1_kernel(const StepArgs a) {
    const TileRef tile = a.tiles[blockIdx.x];
    const int p = tile.p;
    const ProbState st = a.state[p];
    if (st.k_run == 0) return;
    const Geom &g = a.g;
    const size_t plane = (size_t)g.H * g.pitch;
    const T *__restrict__ in = reinterpret_cast<const T *>(a.buf[st.cur]) + (size_t)p * plane;
    T *__restrict__ out = reinterpret_cast<T *>(a.buf[st.cur ^ 1]) + (size_t)p * plane;
    const T *__restrict__ w_up = PLANES ? reinterpret_cast<const T *>(a.w[0]) + (size_t)p * plane : nullptr;
    const T *__restrict__ w_rt = PLANES ? reinterpret_cast<const T *>(a.w[1]) + (size_t)p * plane : nullptr;
    const T *__restrict__ w_dn = PLANES ? reinterpret_cast<const T *>(a.w[2]) + (size_t)p * plane : nullptr;
    const T *__restrict__ w_lf = PLANES ? reinterpret_cast<const T *>(a.w[3]) + (size_t)p * plane : nullptr;
    const double gamma = PLANES ? 0.0 : a.gamma[p];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int xw = tile.x0 + (warp & 1) * 64;
    const int x = xw + 2 * lane;
    const int yb = tile.y0 + (warp >> 1) * 8;
    const int y_lim = min(tile.y0 + kTileH2, g.own_row0 + g.own_rows);
    Acc<T> acc;
    if (xw < g.L) {
        const uint64_t *m64 = reinterpret_cast<const uint64_t *>(a.mask + (size_t)p * g.H * g.mpitch) + (xw >> 6);
        const int mp64 = g.mpitch >> 1;
        for (int r = 0; r < 8; ++r) {
            const int y = yb + r;
            if (y >= y_lim) break;
            const uint64_t bits = m64[(size_t)y * mp64];
            const unsigned mb = (unsigned)(bits >> (2 * lane)) & 3u;
            if (mb == 0) continue;
            const size_t row = (size_t)y * g.pitch;
            const int yg = y + g.row_offset;
            const bool has_up = yg > 0, has_dn = yg < g.global_H - 1;
            double c0, c1, u0 = 0.0, u1 = 0.0, d0 = 0.0, d1 = 0.0, lf = 0.0, rt = 0.0;
            load2(in + row + x, c0, c1);                       // padding columns hold 0
            if (has_up) load2(in + row - g.pitch + x, u0, u1);
            if (has_dn) load2(in + row + g.pitch + x, d0, d1);
            if (x > 0) lf = (double)in[row + x - 1];
            if (x + 2 < g.L) rt = (double)in[row + x + 2];
            double n0, n1;
            if (PLANES) {
                double wu0, wu1, wr0, wr1, wd0, wd1, wl0, wl1;
                load2(w_up + row + x, wu0, wu1);
                load2(w_rt + row + x, wr0, wr1);
                load2(w_dn + row + x, wd0, wd1);
                load2(w_lf + row + x, wl0, wl1);
                n0 = fma(lf, wl0, fma(d0, wd0, fma(c1, wr0, u0 * wu0)));
                n1 = fma(c0, wl1, fma(d1, wd1, fma(rt, wr1, u1 * wu1)));
            } else {
                n0 = fma(lf, gamma, fma(d0, gamma, fma(c1, gamma, u0 * gamma)));
                n1 = fma(c0, gamma, fma(d1, gamma, fma(rt, gamma, u1 * gamma)));
            }
            const int ncy = 4 - (yg == 0 ? 1 : 0) - (yg == g.global_H - 1 ? 1 : 0);
            const int nc0 = ncy - (x == 0 ? 1 : 0) - (x == g.L - 1 ? 1 : 0);
            const int nc1 = ncy - (x + 1 == g.L - 1 ? 1 : 0);
            if (!a.no_nc) {
                n0 = (nc0 == 4) ? n0 * 0.25 : n0 / (double)nc0;
                n1 = (nc1 == 4) ? n1 * 0.25 : n1 / (double)nc1;
            }
            if (n0 > a.clip) n0 = a.clip;
            if (n1 > a.clip) n1 = a.clip;
            if (mb & 1u) out[row + x] = acc.commit(n0, c0);
            if (mb & 2u) out[row + x + 1] = acc.commit(n1, c1);
        }
    }
    write_partial(acc, a, p);
}

// Ratio form of the 2-D sweep for ONE feature layer with exact weights (the VPint2 case): the state buffers
// hold r = P / f, the update is r_new = (sum of in-grid neighbour r) / nc (see vpint_resident.cu for the
// algebra), and P = f * r is formed only for the stopping statistic and the clip.  Per unknown cell and
// sweep the kernel moves 8 B state in + 8 B feature + 8 B state out instead of 48 B with weight planes.
template <int DUMMY>
__global__ void __launch_bounds__(256) jacobi2d_k1_ratio_kernel(const StepArgs a) {
    const TileRef tile = a.tiles[blockIdx.x];
    const int p = tile.p;
    const ProbState st = a.state[p];
    if (st.k_run == 0) return;
    const Geom &g = a.g;
    const size_t plane = (size_t)g.H * g.pitch;
    const double *__restrict__ in = reinterpret_cast<const double *>(a.buf[st.cur]) + (size_t)p * plane;
    double *__restrict__ out = reinterpret_cast<double *>(a.buf[st.cur ^ 1]) + (size_t)p * plane;
    const double *__restrict__ fpl = reinterpret_cast<const double *>(a.w[4]) + (size_t)p * plane;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int xw = tile.x0 + (warp & 1) * 64;
    const int x = xw + 2 * lane;
    const int yb = tile.y0 + (warp >> 1) * 8;
    const int y_lim = min(tile.y0 + kTileH2, g.own_row0 + g.own_rows);
    double s_abs = 0.0, s_old = 0.0;
    int c_abs = 0, c_old = 0;
    if (xw < g.L) {
        const uint64_t *m64 = reinterpret_cast<const uint64_t *>(a.mask + (size_t)p * g.H * g.mpitch) + (xw >> 6);
        const int mp64 = g.mpitch >> 1;
        for (int r = 0; r < 8; ++r) {
            const int y = yb + r;
            if (y >= y_lim) break;
            const uint64_t bits = m64[(size_t)y * mp64];
            const unsigned mb = (unsigned)(bits >> (2 * lane)) & 3u;
            if (mb == 0) continue;
            const size_t row = (size_t)y * g.pitch;
            const int yg = y + g.row_offset;
            const bool has_up = yg > 0, has_dn = yg < g.global_H - 1;
            double c0, c1, u0 = 0.0, u1 = 0.0, d0 = 0.0, d1 = 0.0, lf = 0.0, rt = 0.0, f0, f1;
            load2(in + row + x, c0, c1);
            load2(fpl + row + x, f0, f1);
            if (has_up) load2(in + row - g.pitch + x, u0, u1);
            if (has_dn) load2(in + row + g.pitch + x, d0, d1);
            if (x > 0) lf = in[row + x - 1];
            if (x + 2 < g.L) rt = in[row + x + 2];
            double n0 = ((u0 + c1) + d0) + lf;
            double n1 = ((u1 + rt) + d1) + c0;
            const int ncy = 4 - (yg == 0 ? 1 : 0) - (yg == g.global_H - 1 ? 1 : 0);
            const int nc0 = ncy - (x == 0 ? 1 : 0) - (x == g.L - 1 ? 1 : 0);
            const int nc1 = ncy - (x + 1 == g.L - 1 ? 1 : 0);
            n0 = (nc0 == 4) ? n0 * 0.25 : n0 / (double)nc0;
            n1 = (nc1 == 4) ? n1 * 0.25 : n1 / (double)nc1;
            double v0 = f0 * n0, v1 = f1 * n1;
            if (v0 > a.clip) { v0 = a.clip; n0 = a.clip / f0; }
            if (v1 > a.clip) { v1 = a.clip; n1 = a.clip / f1; }
            if (mb & 1u) {
                const double o = f0 * c0, d = fabs(v0 - o);
                if (d != d) ++c_abs; else s_abs += d;
                if (o != o) ++c_old; else s_old += o;
                out[row + x] = n0;
            }
            if (mb & 2u) {
                const double o = f1 * c1, d = fabs(v1 - o);
                if (d != d) ++c_abs; else s_abs += d;
                if (o != o) ++c_old; else s_old += o;
                out[row + x + 1] = n1;
            }
        }
    }
    Acc<double> acc;
    acc.s_abs = s_abs; acc.s_old = s_old; acc.c_abs = c_abs; acc.c_old = c_old;
    write_partial(acc, a, p);
}

// The default form of the same sweep (VPINT_RATIO_KERNEL=1 selects the kernel above).  A warp takes its 8 rows as two
// half-strips of 4: the 6 state rows and 4 feature rows of a half-strip are requested before any arithmetic (warp-uniform
// tests of the 64-bit mask words decide which: no divergence, ~10 loads in flight per warp), the left / right
// neighbours come from the adjacent lanes by shuffle (two edge loads per row instead of 64), tiles that do not touch
// the grid border skip the neighbour counts, and -- when the caller reduces the sums in a separate launch (step API,
// row-sharded runs) -- there is no block barrier at all: every warp writes its own partial sums and the reduction
// kernel adds the 8 warps of a tile in order.  Thread <-> cell mapping, arithmetic and summation order are those of the
// kernel above: results are bit-identical; half the warp instructions (ncu: 273 M -> 140 M per launch).
// Measured and dropped: a plain sliding window over the 8 rows (fewest loads, but too few in flight: 0.51 against
// 0.42 ms per sweep) and all 10 + 8 rows prefetched at once (128 registers, 2 blocks per SM, barrier stalls: 0.51).
template <bool INTERIOR>
__device__ __forceinline__ void ratio_half_strip(const StepArgs &a, const Geom &g, const double *__restrict__ in,
                                                 double *__restrict__ out, const double *__restrict__ fpl, int xw, int lane,
                                                 int y0, unsigned any, unsigned mine, double &s_abs, double &s_old, int &c_abs,
                                                 int &c_old) {
    // any / mine: bits of the 4 rows y0 .. y0+3 (row r: bit r of any, bits 2r, 2r+1 of mine)
    const int x = xw + 2 * lane;
    const unsigned need = (any << 2) | (any << 1) | any;      // state row i = grid row y0 - 1 + i
    double v0[6], v1[6], f0[4], f1[4], edge[4];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int y = y0 - 1 + i, yg = y + g.row_offset;
        v0[i] = 0.0; v1[i] = 0.0;
        if (((need >> i) & 1u) && yg >= 0 && yg < g.global_H && y >= 0 && y < g.H) load2(in + (size_t)y * g.pitch + x, v0[i], v1[i]);
    }
    const bool has_left = xw > 0, has_right = xw + 64 < g.L;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        f0[r] = 1.0; f1[r] = 1.0; edge[r] = 0.0;
        if ((any >> r) & 1u) {
            const size_t row = (size_t)(y0 + r) * g.pitch;
            load2(fpl + row + x, f0[r], f1[r]);
            if (lane == 0 && has_left) edge[r] = in[row + x - 1];
            if (lane == 31 && has_right) edge[r] = in[row + x + 2];
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (!((any >> r) & 1u)) continue;                       // warp-uniform
        const int y = y0 + r;
        const size_t row = (size_t)y * g.pitch;
        const unsigned mb = (mine >> (2 * r)) & 3u;
        const double u0 = v0[r], u1 = v1[r], c0 = v0[r + 1], c1 = v1[r + 1], d0 = v0[r + 2], d1 = v1[r + 2];
        double lf = __shfl_up_sync(0xffffffffu, c1, 1), rt = __shfl_down_sync(0xffffffffu, c0, 1);
        if (lane == 0) lf = edge[r];
        if (lane == 31) rt = edge[r];
        double n0 = ((u0 + c1) + d0) + lf;
        double n1 = ((u1 + rt) + d1) + c0;
        if (INTERIOR) {
            n0 *= 0.25; n1 *= 0.25;
        } else {
            const int yg = y + g.row_offset;
            const int ncy = 4 - (yg == 0 ? 1 : 0) - (yg == g.global_H - 1 ? 1 : 0);
            const int nc0 = ncy - (x == 0 ? 1 : 0) - (x == g.L - 1 ? 1 : 0);
            const int nc1 = ncy - (x + 1 == g.L - 1 ? 1 : 0);
            n0 = (nc0 == 4) ? n0 * 0.25 : n0 / (double)nc0;
            n1 = (nc1 == 4) ? n1 * 0.25 : n1 / (double)nc1;
        }
        double w0 = f0[r] * n0, w1 = f1[r] * n1;
        if (w0 > a.clip) { w0 = a.clip; n0 = a.clip / f0[r]; }
        if (w1 > a.clip) { w1 = a.clip; n1 = a.clip / f1[r]; }
        if (mb & 1u) {
            const double o = f0[r] * c0, d = fabs(w0 - o);
            if (d != d) ++c_abs; else s_abs += d;
            if (o != o) ++c_old; else s_old += o;
            out[row + x] = n0;
        }
        if (mb & 2u) {
            const double o = f1[r] * c1, d = fabs(w1 - o);
            if (d != d) ++c_abs; else s_abs += d;
            if (o != o) ++c_old; else s_old += o;
            out[row + x + 1] = n1;
        }
    }
}

__global__ void __launch_bounds__(256, 3) jacobi2d_k1_ratio4_kernel(const StepArgs a) {
    const TileRef tile = a.tiles[blockIdx.x];
    const int p = tile.p;
    const ProbState st = a.state[p];
    if (st.k_run == 0) return;
    const Geom &g = a.g;
    const size_t plane = (size_t)g.H * g.pitch;
    const double *__restrict__ in = reinterpret_cast<const double *>(a.buf[st.cur]) + (size_t)p * plane;
    double *__restrict__ out = reinterpret_cast<double *>(a.buf[st.cur ^ 1]) + (size_t)p * plane;
    const double *__restrict__ fpl = reinterpret_cast<const double *>(a.w[4]) + (size_t)p * plane;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int xw = tile.x0 + (warp & 1) * 64;
    const int yb = tile.y0 + (warp >> 1) * 8;
    const int y_lim = min(tile.y0 + kTileH2, g.own_row0 + g.own_rows);
    double s_abs = 0.0, s_old = 0.0;
    int c_abs = 0, c_old = 0;
    if (xw < g.L && yb < y_lim) {
        const uint64_t *m64 = reinterpret_cast<const uint64_t *>(a.mask + (size_t)p * g.H * g.mpitch) + (xw >> 6);
        const int mp64 = g.mpitch >> 1;
        uint64_t mword = 0;
        if (lane < 8 && yb + lane < y_lim) mword = m64[(size_t)(yb + lane) * mp64];
        unsigned any = 0, mine = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint64_t w = __shfl_sync(0xffffffffu, mword, r);
            any |= (w != 0ull) ? (1u << r) : 0u;
            mine |= ((unsigned)(w >> (2 * lane)) & 3u) << (2 * r);
        }
        // no cell of the tile on the border of the GLOBAL grid: every neighbour count is 4
        const bool interior = (tile.y0 + g.row_offset > 0) && (tile.y0 + kTileH2 + g.row_offset < g.global_H) && (tile.x0 > 0) &&
                              (tile.x0 + kTileW2 < g.L);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const unsigned any_h = (any >> (4 * h)) & 15u, mine_h = (mine >> (8 * h)) & 255u;
            if (any_h == 0) continue;
            if (interior) ratio_half_strip<true>(a, g, in, out, fpl, xw, lane, yb + 4 * h, any_h, mine_h, s_abs, s_old, c_abs, c_old);
            else ratio_half_strip<false>(a, g, in, out, fpl, xw, lane, yb + 4 * h, any_h, mine_h, s_abs, s_old, c_abs, c_old);
        }
    }
    if (a.partial8) {
        // per-warp partial sums, no block barrier: the reduction kernel adds the 8 warps of a tile in order
        const double q0 = warp_sum(s_abs), q1 = warp_sum(s_old);
        const int q2 = warp_sum(c_abs), q3 = warp_sum(c_old);
        if (lane == 0) {
            double *dst = a.partial8 + ((size_t)blockIdx.x * 8 + warp) * kSumSlots;
            dst[0] = q0; dst[1] = q1; dst[2] = (double)q2; dst[3] = (double)q3;
        }
        return;
    }
    Acc<double> acc;
    acc.s_abs = s_abs; acc.s_old = s_old; acc.c_abs = c_abs; acc.c_old = c_old;
    write_partial(acc, a, p);
}

// Resistance sweep (VPint/WP_MRP.py:327-348): the damping  new > mu -> old + (new - old) - epsilon * old  is
// applied after the restore to EVERY cell, so known cells move too and nothing can be skipped: one block
// per tile of the full grid (no active-tile list), every cell written, sums over all cells.
template <typename T, bool PLANES>
__global__ void __launch_bounds__(256) jacobi2d_resist_kernel(const StepArgs a) {
    const Geom &g = a.g;
    const int p = blockIdx.x / g.tiles_per_prob;
    const int tr = blockIdx.x - p * g.tiles_per_prob;
    const int ty = tr / g.tiles_x, tx = tr - ty * g.tiles_x;
    const ProbState st = a.state[p];
    Acc<T> acc;
    if (st.k_run != 0) {
        const size_t plane = (size_t)g.H * g.pitch;
        const T *__restrict__ in = reinterpret_cast<const T *>(a.buf[st.cur]) + (size_t)p * plane;
        T *__restrict__ out = reinterpret_cast<T *>(a.buf[st.cur ^ 1]) + (size_t)p * plane;
        const T *__restrict__ orig = reinterpret_cast<const T *>(a.orig) + (size_t)p * plane;
        const uint32_t *mrow = a.mask + (size_t)p * g.H * g.mpitch;
        const double gamma = PLANES ? 0.0 : a.gamma[p];
        const int x = tx * kTileW2 + (threadIdx.x & (kTileW2 - 1));
        const int y_lim = min(g.own_row0 + (ty + 1) * kTileH2, g.own_row0 + g.own_rows);
        if (x < g.L) {
            for (int y = g.own_row0 + ty * kTileH2 + (threadIdx.x >> 7); y < y_lim; y += 2) {
                const size_t row = (size_t)y * g.pitch;
                const int yg = y + g.row_offset;
                const double ov = (double)in[row + x];
                double nv;
                if ((mrow[(size_t)y * g.mpitch + (x >> 5)] >> (x & 31)) & 1u) {
                    const double up = (yg > 0) ? (double)in[row - g.pitch + x] : 0.0;
                    const double dn = (yg < g.global_H - 1) ? (double)in[row + g.pitch + x] : 0.0;
                    const double lf = (x > 0) ? (double)in[row + x - 1] : 0.0;
                    const double rt = (x < g.L - 1) ? (double)in[row + x + 1] : 0.0;
                    if (PLANES) {
                        const double wu = (double)reinterpret_cast<const T *>(a.w[0])[(size_t)p * plane + row + x];
                        const double wr = (double)reinterpret_cast<const T *>(a.w[1])[(size_t)p * plane + row + x];
                        const double wd = (double)reinterpret_cast<const T *>(a.w[2])[(size_t)p * plane + row + x];
                        const double wl = (double)reinterpret_cast<const T *>(a.w[3])[(size_t)p * plane + row + x];
                        nv = fma(lf, wl, fma(dn, wd, fma(rt, wr, up * wu)));
                    } else {
                        nv = fma(lf, gamma, fma(dn, gamma, fma(rt, gamma, up * gamma)));
                    }
                    if (!a.no_nc) {
                        const int nc = 4 - (yg == 0 ? 1 : 0) - (yg == g.global_H - 1 ? 1 : 0) - (x == 0 ? 1 : 0) - (x == g.L - 1 ? 1 : 0);
                        nv = (nc == 4) ? nv * 0.25 : nv / (double)nc;
                    }
                    if (nv > a.clip) nv = a.clip;
                } else {
                    nv = (double)orig[row + x];                // restore (WP_MRP.py:321,323)
                }
                const double dy = nv - ov;
                double t = nv;
                const double mu = a.mu_arr ? a.mu_arr[p] : a.mu, eps = a.eps_arr ? a.eps_arr[p] : a.eps;   // per problem: batched trials
                if (nv <= mu) t = ov + dy;                     // WP_MRP.py:340-344
                if (t > mu) t = (ov + dy) - eps * ov;          // WP_MRP.py:346-348
                out[row + x] = acc.commit(t, ov);
            }
        }
    }
    write_partial(acc, a, p);
}

// x / n for a small positive integer n, correctly rounded like the IEEE division the reference performs:
// q = x * fl(1/n), one fma residual correction (Markstein; checked against x / n on 1e8 random operands);
// a handful of instructions instead of the ~30 of a general fp64 division.  Non-finite or extreme x and
// n outside 1..6 take the real division.
__device__ __forceinline__ double div_small(double x, int n) {
    const double rcp = (n == 6) ? (1.0 / 6.0) : (n == 5) ? 0.2 : (n == 4) ? 0.25 : (n == 3) ? (1.0 / 3.0) : (n == 2) ? 0.5 : 1.0;
    const double ax = fabs(x);
    if (n >= 1 && n <= 6 && ax < 1e300 && (ax > 1e-290 || x == 0.0)) {
        const double d = (double)n;
        const double q = x * rcp;
        const double r = fma(-d, q, x);
        return fma(r, rcp, q);
    }
    return x / (double)n;
}

// 3-D sweep on the time-major slice stack: slice s = p * T + t is an (H, pitch) image, so the kernel is the
// 2-D one (same tiles, two cells per thread, 128-bit loads) plus the two temporal neighbours one slice away.
// Value slots 4 / 5 are the t+1 / t-1 neighbours (VPint/SD_MRP.py:372-378).  Plane mode keeps the
// reference's pairing: the temporal weight multiplies the t+1 value only where t > 0 and the t-1 value only
// where t < T-1 (VPint/WP_MRP.py:1344-1353 vs 1406-1412).  Weight planes are (H, W) per problem: features are
// static in time.
template <typename T, bool PLANES>
__global__ void __launch_bounds__(256, 5) jacobi3d_k1_kernel(const StepArgs a) {
    const TileRef tile = a.tiles[blockIdx.x];
    const Geom &g = a.g;
    const int sl = tile.p;
    const int p = sl / g.S, t = sl - p * g.S;
    const ProbState st = a.state[p];
    if (st.k_run == 0) return;
    const size_t plane = (size_t)g.H * g.pitch;
    const T *__restrict__ in = reinterpret_cast<const T *>(a.buf[st.cur]) + (size_t)sl * plane;
    T *__restrict__ out = reinterpret_cast<T *>(a.buf[st.cur ^ 1]) + (size_t)sl * plane;
    const bool has_nx = t < g.S - 1, has_pv = t > 0;
    const T *__restrict__ in_nx = in + plane, *__restrict__ in_pv = in - plane;      // only dereferenced when they exist
    const size_t wbase = (size_t)p * plane;
    const double gamma = PLANES ? 0.0 : a.gamma[p];
    const double tau = PLANES ? 0.0 : a.tau[p];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int xw = tile.x0 + (warp & 1) * 64;
    const int x = xw + 2 * lane;
    const int yb = tile.y0 + (warp >> 1) * 8;
    const int y_lim = min(tile.y0 + kTileH2, g.own_row0 + g.own_rows);
    const int nct = 6 - (has_nx ? 0 : 1) - (has_pv ? 0 : 1);
    Acc<T> acc;
    if (xw < g.L) {
        const uint64_t *m64 = reinterpret_cast<const uint64_t *>(a.mask + (size_t)sl * g.H * g.mpitch) + (xw >> 6);
        const int mp64 = g.mpitch >> 1;
        for (int r = 0; r < 8; ++r) {
            const int y = yb + r;
            if (y >= y_lim) break;
            const uint64_t bits = m64[(size_t)y * mp64];
            const unsigned mb = (unsigned)(bits >> (2 * lane)) & 3u;
            if (mb == 0) continue;
            const size_t row = (size_t)y * g.pitch;
            const int yg = y + g.row_offset;
            double c0, c1, u0 = 0.0, u1 = 0.0, d0 = 0.0, d1 = 0.0, lf = 0.0, rt = 0.0, n0 = 0.0, n1 = 0.0, q0 = 0.0, q1 = 0.0;
            load2(in + row + x, c0, c1);                       // padding columns hold 0
            if (yg > 0) load2(in + row - g.pitch + x, u0, u1);
            if (yg < g.global_H - 1) load2(in + row + g.pitch + x, d0, d1);
            if (x > 0) lf = (double)in[row + x - 1];
            if (x + 2 < g.L) rt = (double)in[row + x + 2];
            if (has_nx) load2(in_nx + row + x, n0, n1);
            if (has_pv) load2(in_pv + row + x, q0, q1);
            double v0, v1;
            if (PLANES) {
                double wu0, wu1, wr0, wr1, wd0, wd1, wl0, wl1, wt0, wt1;
                load2(reinterpret_cast<const T *>(a.w[0]) + wbase + row + x, wu0, wu1);
                load2(reinterpret_cast<const T *>(a.w[1]) + wbase + row + x, wr0, wr1);
                load2(reinterpret_cast<const T *>(a.w[2]) + wbase + row + x, wd0, wd1);
                load2(reinterpret_cast<const T *>(a.w[3]) + wbase + row + x, wl0, wl1);
                load2(reinterpret_cast<const T *>(a.w[4]) + wbase + row + x, wt0, wt1);
                const double w40 = has_pv ? wt0 : 0.0, w41 = has_pv ? wt1 : 0.0;   // slot 4: set where t > 0, applied to t+1
                const double w50 = has_nx ? wt0 : 0.0, w51 = has_nx ? wt1 : 0.0;   // slot 5: set where t < T-1, applied to t-1
                v0 = fma(q0, w50, fma(n0, w40, fma(lf, wl0, fma(d0, wd0, fma(c1, wr0, u0 * wu0)))));
                v1 = fma(q1, w51, fma(n1, w41, fma(c0, wl1, fma(d1, wd1, fma(rt, wr1, u1 * wu1)))));
            } else {
                v0 = fma(q0, tau, fma(n0, tau, fma(lf, gamma, fma(d0, gamma, fma(c1, gamma, u0 * gamma)))));
                v1 = fma(q1, tau, fma(n1, tau, fma(c0, gamma, fma(d1, gamma, fma(rt, gamma, u1 * gamma)))));
            }
            const int ncy = nct - (yg == 0 ? 1 : 0) - (yg == g.global_H - 1 ? 1 : 0);
            const int nc0 = ncy - (x == 0 ? 1 : 0) - (x == g.L - 1 ? 1 : 0);
            const int nc1 = ncy - (x + 1 == g.L - 1 ? 1 : 0);
            v0 = div_small(v0, nc0);
            v1 = div_small(v1, nc1);
            if (mb & 1u) out[row + x] = acc.commit(v0, c0);
            if (mb & 2u) out[row + x + 1] = acc.commit(v1, c1);
        }
    }
    write_partial(acc, a, p);
}

}  // namespace

void k1_tile_shape(int ndim, int *th, int *tw) {
    (void)ndim;                 // 3-D slices use the 2-D tile
    *th = kTileH2;
    *tw = kTileW2;
}

#define VP_LAUNCH_CHECK()                         \
    do {                                          \
        cudaError_t e__ = cudaGetLastError();     \
        if (e__ != cudaSuccess) return e__;       \
        if (launches) ++*launches;                \
    } while (0)

cudaError_t launch_step2d(const StepArgs &a, int storage, int weight_mode, int64_t n_tiles, cudaStream_t st,
                          int64_t *launches) {
    if (n_tiles <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)n_tiles;
    const bool planes = (weight_mode == VPINT_W_PLANES);
    if (storage == VPINT_F64) {
        if (planes) jacobi2d_k1_kernel<double, true><<<grid, 256, 0, st>>>(a);
        else jacobi2d_k1_kernel<double, false><<<grid, 256, 0, st>>>(a);
    } else {
        if (planes) jacobi2d_k1_kernel<float, true><<<grid, 256, 0, st>>>(a);
        else jacobi2d_k1_kernel<float, false><<<grid, 256, 0, st>>>(a);
    }
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_step2d_ratio(const StepArgs &a, int64_t n_tiles, cudaStream_t st, int64_t *launches) {
    if (n_tiles <= 0) return cudaSuccess;
    // 5 (default): the copy-engine-fed kernel of vpint_step_bulk.cu wherever the sums are reduced by a separate launch;
    // 4: the register-fed kernel below everywhere; 1: round 1's kernel
    const char *ver_env = getenv("VPINT_RATIO_KERNEL");           // read per launch: a triage switch, also toggled by the tests
    const int ver = ver_env ? atoi(ver_env) : 5;
    if (ver >= 5 && step2d_ratio_bulk_supported(a)) return launch_step2d_ratio_bulk(a, n_tiles, st, launches);
    if (ver == 1) jacobi2d_k1_ratio_kernel<0><<<(unsigned)n_tiles, 256, 0, st>>>(a);
    else jacobi2d_k1_ratio4_kernel<<<(unsigned)n_tiles, 256, 0, st>>>(a);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_step2d_resist(const StepArgs &a, int storage, int weight_mode, cudaStream_t st, int64_t *launches) {
    const unsigned grid = (unsigned)((int64_t)a.g.P * a.g.tiles_per_prob);
    const bool planes = (weight_mode == VPINT_W_PLANES);
    if (storage == VPINT_F64) {
        if (planes) jacobi2d_resist_kernel<double, true><<<grid, 256, 0, st>>>(a);
        else jacobi2d_resist_kernel<double, false><<<grid, 256, 0, st>>>(a);
    } else {
        if (planes) jacobi2d_resist_kernel<float, true><<<grid, 256, 0, st>>>(a);
        else jacobi2d_resist_kernel<float, false><<<grid, 256, 0, st>>>(a);
    }
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_step3d(const StepArgs &a, int storage, int weight_mode, int64_t n_tiles, cudaStream_t st,
                          int64_t *launches) {
    if (n_tiles <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)n_tiles;
    const bool planes = (weight_mode == VPINT_W_PLANES);
    if (storage == VPINT_F64) {
        if (planes) jacobi3d_k1_kernel<double, true><<<grid, 256, 0, st>>>(a);
        else jacobi3d_k1_kernel<double, false><<<grid, 256, 0, st>>>(a);
    } else {
        if (planes) jacobi3d_k1_kernel<float, true><<<grid, 256, 0, st>>>(a);
        else jacobi3d_k1_kernel<float, false><<<grid, 256, 0, st>>>(a);
    }
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

}  // namespace vpint
