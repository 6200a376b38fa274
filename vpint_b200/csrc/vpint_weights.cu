// A2: feature bands -> per-cell direction weight planes.
//   2-D 'exact' (vectorised branch)        VPint/WP_MRP.py:138-170
//   pairwise methods (predict_weight)      VPint/WP_MRP.py:173-195,396-427
//   3-D (features static in time)          VPint/WP_MRP.py:1319-1355,1443-1468
// Planes are stored planar ([P][H][pitch] per direction) in the solver's storage type.
#include "vpint_internal.cuh"

namespace vpint {

namespace {

// NumPy's summation order for a contiguous reduction (np.mean / np.sum over the feature axis):
// < 8 terms sequential from 0.0; <= 128 terms eight running accumulators combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail; larger inputs split in halves
// (multiples of 8).  Verified bit-for-bit against NumPy 2.3 for F in 1..300.
template <typename Fn>
__device__ double np_sum_block(Fn term, int off, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += term(off + i);
        return r;
    }
    double r0 = term(off + 0), r1 = term(off + 1), r2 = term(off + 2), r3 = term(off + 3);
    double r4 = term(off + 4), r5 = term(off + 5), r6 = term(off + 6), r7 = term(off + 7);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
        r0 += term(off + i + 0); r1 += term(off + i + 1); r2 += term(off + i + 2); r3 += term(off + i + 3);
        r4 += term(off + i + 4); r5 += term(off + i + 5); r6 += term(off + i + 6); r7 += term(off + i + 7);
    }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res += term(off + i);
    return res;
}

template <typename Fn>
__device__ double np_sum(Fn term, int off, int n) {
    if (n <= 128) return np_sum_block(term, off, n);
    // iterative form of the recursive halving: explicit stack of (off, n) ranges, left first
    int s_off[24], s_n[24];
    double s_acc[24];
    int s_state[24];
    int sp = 0;
    s_off[0] = off; s_n[0] = n; s_state[0] = 0; s_acc[0] = 0.0;
    double ret = 0.0;
    while (sp >= 0) {
        const int o = s_off[sp], m = s_n[sp];
        if (m <= 128) {
            ret = np_sum_block(term, o, m);
            --sp;
            continue;
        }
        int h = m / 2;
        h -= h % 8;
        if (s_state[sp] == 0) {            // descend left
            s_state[sp] = 1;
            ++sp;
            s_off[sp] = o; s_n[sp] = h; s_state[sp] = 0;
        } else if (s_state[sp] == 1) {     // left done -> descend right
            s_acc[sp] = ret;
            s_state[sp] = 2;
            ++sp;
            s_off[sp] = o + h; s_n[sp] = m - h; s_state[sp] = 0;
        } else {                           // both done
            ret = s_acc[sp] + ret;
            --sp;
        }
    }
    return ret;
}

template <typename TIn>
struct FeatRef {
    const TIn *base;   // features of one cell
    int64_t sF;
    __device__ double operator()(int k) const { return (double)base[(int64_t)k * sF]; }
};

// One weight between a neighbour cell and the own cell.
//   method EXACT, replace_own = 1 : vectorised branch, zeros of BOTH cells read as 0.01 (WP_MRP.py:143-145)
//   method EXACT, replace_own = 0 : predict_weight, only the neighbour's zeros replaced (WP_MRP.py:414-417)
template <typename TIn>
__device__ double pair_weight(const FeatRef<TIn> nb, const FeatRef<TIn> own, int F, int method, int replace_own,
                              int clamp, double lo, double hi) {
    double g;
    if (method == VPINT_METHOD_EXACT) {
        auto term = [&](int k) {
            double o = own(k), d = nb(k);
            if (replace_own && o == 0.0) o = 0.01;
            if (d == 0.0) d = 0.01;
            return o / d;
        };
        g = np_sum(term, 0, F) / (double)F;
    } else if (method == VPINT_METHOD_EXACT_INVERSE) {
        auto term = [&](int k) {
            double o = own(k);
            if (o == 0.0) o = 0.01;
            return nb(k) / o;
        };
        g = np_sum(term, 0, F) / (double)F;
    } else {   // cosine_similarity: dot(f1, f2) / max(sum(f1) * sum(f2), 0.01)
        double dot = 0.0;
        for (int k = 0; k < F; ++k) dot = fma(nb(k), own(k), dot);
        const double s1 = np_sum([&](int k) { return nb(k); }, 0, F);
        const double s2 = np_sum([&](int k) { return own(k); }, 0, F);
        const double prod = s1 * s2;
        const double den = (0.01 > prod) ? 0.01 : prod;   // Python max(prod, 0.01): prod unless 0.01 > prod
        g = dot / den;
    }
    if (clamp) {   // max(min_gamma, min(g, max_gamma)) with Python's comparison order (NaN -> min_gamma)
        g = (hi < g) ? hi : g;
        g = (g > lo) ? g : lo;
    }
    return g;
}

// Block per (problem, row); threads stride over the padded row.  n_planes = 4 (2-D) or 5 (3-D: plus the
// temporal self-weight).  Rows/columns without the neighbour get weight 0 (WP_MRP.py:162-168, 176-193).
template <typename TIn, typename TSt>
__global__ void __launch_bounds__(256) weights_kernel(int Pin, int H, int W, int wpitch, int row_offset, int global_H,
                                                      const TIn *__restrict__ feat, int64_t sO, int64_t sI, int64_t sY,
                                                      int64_t sX, int64_t sF, int F, int method, int replace_own,
                                                      int clamp, double lo, double hi, int n_planes,
                                                      TSt *__restrict__ w_up, TSt *__restrict__ w_right,
                                                      TSt *__restrict__ w_down, TSt *__restrict__ w_left,
                                                      TSt *__restrict__ w_time) {
    const int p = blockIdx.x / H, y = blockIdx.x - p * H;
    const size_t row = ((size_t)p * H + y) * wpitch;
    const int yg = y + row_offset;
    const int po = p / Pin, pi = p - po * Pin;
    const TIn *frow = feat + (int64_t)po * sO + (int64_t)pi * sI + (int64_t)y * sY;
    for (int x = threadIdx.x; x < wpitch; x += blockDim.x) {
        double up = 0.0, right = 0.0, down = 0.0, left = 0.0, tw = 0.0;
        if (x < W) {
            const FeatRef<TIn> own{frow + (int64_t)x * sX, sF};
            if (yg > 0 && y > 0)
                up = pair_weight(FeatRef<TIn>{own.base - sY, sF}, own, F, method, replace_own, clamp, lo, hi);
            if (x < W - 1)
                right = pair_weight(FeatRef<TIn>{own.base + sX, sF}, own, F, method, replace_own, clamp, lo, hi);
            if (yg < global_H - 1 && y < H - 1)
                down = pair_weight(FeatRef<TIn>{own.base + sY, sF}, own, F, method, replace_own, clamp, lo, hi);
            if (x > 0)
                left = pair_weight(FeatRef<TIn>{own.base - sX, sF}, own, F, method, replace_own, clamp, lo, hi);
            if (n_planes == 5) tw = pair_weight(own, own, F, method, replace_own, clamp, lo, hi);
        }
        w_up[row + x] = (TSt)up;
        w_right[row + x] = (TSt)right;
        w_down[row + x] = (TSt)down;
        w_left[row + x] = (TSt)left;
        if (n_planes == 5) w_time[row + x] = (TSt)tw;
    }
}

// Fast path for one planar fp64 feature band per problem (F = 1, unit column stride, 16-byte aligned
// rows): two cells per thread, every global access a 128-bit transaction.
__global__ void __launch_bounds__(256) weights_exact_f1_kernel(int P, int H, int W, int pitch, int row_offset,
                                                               int global_H, const double *__restrict__ feat,
                                                               int64_t sP, int64_t sY,
                                                               double *__restrict__ w_up,
                                                               double *__restrict__ w_right,
                                                               double *__restrict__ w_down,
                                                               double *__restrict__ w_left) {
    const int p = blockIdx.x / H, y = blockIdx.x - p * H;
    const size_t row = ((size_t)p * H + y) * pitch;
    const int yg = y + row_offset;
    const bool has_up = (yg > 0 && y > 0), has_down = (yg < global_H - 1 && y < H - 1);
    const double *frow = feat + (int64_t)p * sP + (int64_t)y * sY;
    auto fix = [](double v) { return v == 0.0 ? 0.01 : v; };
    for (int x = 2 * threadIdx.x; x < pitch; x += 2 * blockDim.x) {
        double2 up = {0, 0}, right = {0, 0}, down = {0, 0}, left = {0, 0};
        if (x < W) {
            const bool two = (x + 1 < W);
            double c0, c1 = 1.0;
            if (two) { const double2 c = *reinterpret_cast<const double2 *>(frow + x); c0 = fix(c.x); c1 = fix(c.y); }
            else c0 = fix(frow[x]);
            if (has_up) {
                if (two) { const double2 u = *reinterpret_cast<const double2 *>(frow - sY + x); up.x = c0 / fix(u.x); up.y = c1 / fix(u.y); }
                else up.x = c0 / fix(frow[x - sY]);
            }
            if (has_down) {
                if (two) { const double2 d = *reinterpret_cast<const double2 *>(frow + sY + x); down.x = c0 / fix(d.x); down.y = c1 / fix(d.y); }
                else down.x = c0 / fix(frow[x + sY]);
            }
            if (x > 0) left.x = c0 / fix(frow[x - 1]);
            if (two) {
                left.y = c1 / c0;
                right.x = c0 / c1;
                if (x + 2 < W) right.y = c1 / fix(frow[x + 2]);
            }
        }
        *reinterpret_cast<double2 *>(w_up + row + x) = up;
        *reinterpret_cast<double2 *>(w_right + row + x) = right;
        *reinterpret_cast<double2 *>(w_down + row + x) = down;
        *reinterpret_cast<double2 *>(w_left + row + x) = left;
    }
}

// One feature layer, 'exact' weights: keep the (zero -> 0.01) feature itself as a planar fp64 plane.  The
// cluster-resident kernel iterates P / f with it (vpint_resident.cu); the four ratio planes are only built
// from it if a tiled launch is needed (planes_from_fplane_kernel).  bad_flag is raised for a non-finite
// feature, which sends the solve down the plane path where NaN / inf propagate exactly as in the reference.
template <typename TIn>
__global__ void __launch_bounds__(256) fplane_kernel(int Pin, int H, int W, int pitch, const TIn *__restrict__ feat,
                                                     int64_t sO, int64_t sI, int64_t sY, int64_t sX,
                                                     double *__restrict__ fplane, int *bad_flag,
                                                     double *__restrict__ extreme) {
    const int p = blockIdx.x / H, y = blockIdx.x - p * H;
    const int po = p / Pin, pi = p - po * Pin;
    const TIn *frow = feat + (int64_t)po * sO + (int64_t)pi * sI + (int64_t)y * sY;
    double *out = fplane + ((size_t)p * H + y) * pitch;
    bool bad = false, wild = false, wide = false;
    for (int x = threadIdx.x; x < pitch; x += blockDim.x) {
        double f = 1.0;
        if (x < W) {
            f = (double)frow[(int64_t)x * sX];
            if (f == 0.0) f = 0.01;                     // VPint/WP_MRP.py:143-145
            if (!(fabs(f) <= 1.7976931348623157e308)) bad = true;
            if (!(fabs(f) < 1e150) || !(fabs(f) > 1e-150)) wild = true;
            if ((double)(float)f != f) wide = true;      // not a float32 value (0.01 is not): fp64 on chip
        }
        out[x] = f;
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicExch(bad_flag, 1);
    if (__syncthreads_or(wide) && threadIdx.x == 0) atomicExch(bad_flag + 3, 1);
    if (__syncthreads_or(wild) && threadIdx.x == 0) atomicAdd(extreme, 1.0);
}

// fplane_kernel for a band-interleaved stack (band fastest): one warp per (patch, row) stages the row of all bands
// with coalesced loads and writes every band's plane row from shared memory.
constexpr int kFplaneWarps = 4;

template <typename TIn>
__global__ void __launch_bounds__(kFplaneWarps * 32) fplane_bands_kernel(int n_patches, int Pin, int H, int W, int pitch,
                                                                         const TIn *__restrict__ feat, int64_t sO, int64_t sY,
                                                                         double *__restrict__ fplane, int *bad_flag,
                                                                         double *__restrict__ extreme) {
    extern __shared__ __align__(16) unsigned char fp_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * kFplaneWarps + warp;
    if (r >= (int64_t)n_patches * H) return;
    const int po = (int)(r / H), y = (int)(r - (int64_t)po * H);
    TIn *my = reinterpret_cast<TIn *>(fp_raw) + (size_t)warp * W * Pin;
    const TIn *src = feat + (int64_t)po * sO + (int64_t)y * sY;
    warp_stage(my, src, W * Pin, lane);
    __syncwarp();
    bool bad = false, wild = false, wide = false;
    for (int b = 0; b < Pin; ++b) {
        double *out = fplane + ((size_t)(po * Pin + b) * H + y) * pitch;
        for (int x = lane; x < pitch; x += 32) {
            double f = 1.0;
            if (x < W) {
                f = (double)my[x * Pin + b];
                if (f == 0.0) f = 0.01;                 // VPint/WP_MRP.py:143-145
                if (!(fabs(f) <= 1.7976931348623157e308)) bad = true;
                if (!(fabs(f) < 1e150) || !(fabs(f) > 1e-150)) wild = true;
                if ((double)(float)f != f) wide = true;  // not a float32 value (0.01 is not): fp64 on chip
            }
            out[x] = f;
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicExch(bad_flag, 1);
    if (__any_sync(0xffffffffu, wide) && lane == 0) atomicExch(bad_flag + 3, 1);
    if (__any_sync(0xffffffffu, wild) && lane == 0) atomicAdd(extreme, 1.0);
}

// w_dir = f[i,j] / f[neighbour], missing neighbour -> 0 (VPint/WP_MRP.py:149-170 with one feature layer).
template <typename TSt>
__global__ void __launch_bounds__(256) planes_from_fplane_kernel(int H, int W, int pitch, int row_offset, int global_H,
                                                                 const double *__restrict__ fplane,
                                                                 TSt *__restrict__ w_up, TSt *__restrict__ w_right,
                                                                 TSt *__restrict__ w_down, TSt *__restrict__ w_left) {
    const int p = blockIdx.x / H, y = blockIdx.x - p * H;
    const size_t row = ((size_t)p * H + y) * pitch;
    const int yg = y + row_offset;
    const bool has_up = (yg > 0 && y > 0), has_down = (yg < global_H - 1 && y < H - 1);
    for (int x = threadIdx.x; x < pitch; x += blockDim.x) {
        double up = 0.0, right = 0.0, down = 0.0, left = 0.0;
        if (x < W) {
            const double c = fplane[row + x];
            if (has_up) up = c / fplane[row - pitch + x];
            if (x < W - 1) right = c / fplane[row + x + 1];
            if (has_down) down = c / fplane[row + pitch + x];
            if (x > 0) left = c / fplane[row + x - 1];
        }
        w_up[row + x] = (TSt)up;
        w_right[row + x] = (TSt)right;
        w_down[row + x] = (TSt)down;
        w_left[row + x] = (TSt)left;
    }
}

// prioritise_identity (VPint/WP_MRP.py:243-276): static priority planes from the weights, folded back into
// the weights so that the sweep  new = sum(w * v * p) / sum(p)  (WP_MRP.py:311-315) runs through the ordinary
// stencil with the neighbour-count division switched off.
template <typename TSt>
__global__ void __launch_bounds__(256) priority_kernel(int H, int W, int pitch, int row_offset, int global_H,
                                                       TSt *__restrict__ w0, TSt *__restrict__ w1,
                                                       TSt *__restrict__ w2, TSt *__restrict__ w3, double beta_all,
                                                       const double *__restrict__ beta_arr,
                                                       const double *__restrict__ bias, double *__restrict__ p_out) {
    const int p = blockIdx.x / H, y = blockIdx.x - p * H;
    const double beta = beta_arr ? beta_arr[p] : beta_all;      // per problem: batched trials of a parameter search
    const size_t row = ((size_t)p * H + y) * pitch;
    const int yg = y + row_offset;
    TSt *wp[4] = {w0, w1, w2, w3};
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        double w[4], pr[4];
#pragma unroll
        for (int d = 0; d < 4; ++d) w[d] = (double)wp[d][row + x];
        if (beta == 0.0) {                          // ones, zero where the neighbour is missing (:247-255)
            pr[0] = (yg > 0) ? 1.0 : 0.0;
            pr[1] = (x < W - 1) ? 1.0 : 0.0;
            pr[2] = (yg < global_H - 1) ? 1.0 : 0.0;
            pr[3] = (x > 0) ? 1.0 : 0.0;
        } else {
#pragma unroll
            for (int d = 0; d < 4; ++d) {           // two SEQUENTIAL rewrites (:257-259)
                double q = w[d];
                if (q > 1.0) { const double a = 1.0 / q; q = a * (a / (a * beta)); }
                if (q < 1.0) q = q * (q / ((q + 0.001) * beta));
                pr[d] = q;
            }
        }
        if (p_out) {                                // self.priority_grid is stored before the bias (:260)
#pragma unroll
            for (int d = 0; d < 4; ++d) p_out[(((size_t)p * H + y) * W + x) * 4 + d] = pr[d];
        }
        if (bias) {                                 // known_value_bias (:263-276)
            const double b = bias[((size_t)p * H + y) * W + x];
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                pr[d] = pr[d] * b;
                if (pr[d] == 0.0) pr[d] = 0.01;
            }
        }
        const double psum = (((0.0 + pr[0]) + pr[1]) + pr[2]) + pr[3];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            wp[d][row + x] = (TSt)((w[d] * pr[d]) / psum);
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

cudaError_t launch_weights2d(const Geom &g, const void *feat, int dtype, const int64_t *stride, int F, int method,
                             int clamp, double min_gamma, double max_gamma, int storage, void *const *planes,
                             cudaStream_t st, int64_t *launches) {
    const unsigned grid = (unsigned)((int64_t)g.P * g.H);
    const int64_t sO = stride[0], sI = stride[1], sY = stride[2], sX = stride[3], sF = stride[4];
    const int64_t sP = sO;
    const int replace_own = (method == VPINT_METHOD_EXACT && !clamp) ? 1 : 0;
    const bool fast = (dtype == VPINT_F64 && storage == VPINT_F64 && F == 1 && sX == 1 && method == VPINT_METHOD_EXACT &&
                       !clamp && g.Pin == 1 && (sY % 2 == 0) && (sP % 2 == 0) && ((uintptr_t)feat % 16 == 0));
    if (fast) {
        weights_exact_f1_kernel<<<grid, 256, 0, st>>>(g.P, g.H, g.W, g.pitch, g.row_offset, g.global_H,
                                                      (const double *)feat, sP, sY, (double *)planes[0],
                                                      (double *)planes[1], (double *)planes[2],
                                                      (double *)planes[3]);
        VP_LAUNCH_CHECK();
        return cudaSuccess;
    }
#define VP_W(TIN, TST)                                                                                         \
    weights_kernel<TIN, TST><<<grid, 256, 0, st>>>(g.Pin, g.H, g.W, g.pitch, g.row_offset, g.global_H,         \
                                                   (const TIN *)feat, sO, sI, sY, sX, sF, F, method, replace_own, \
                                                   clamp, min_gamma, max_gamma, 4, (TST *)planes[0],           \
                                                   (TST *)planes[1], (TST *)planes[2], (TST *)planes[3],       \
                                                   (TST *)nullptr)
    if (dtype == VPINT_F64 && storage == VPINT_F64) VP_W(double, double);
    else if (dtype == VPINT_F32 && storage == VPINT_F64) VP_W(float, double);
    else if (dtype == VPINT_F64 && storage == VPINT_F32) VP_W(double, float);
    else VP_W(float, float);
#undef VP_W
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_fplane(const Geom &g, const void *feat, int dtype, const int64_t *stride, double *fplane, int *bad_flag,
                          double *extreme, cudaStream_t st, int64_t *launches) {
    const unsigned grid = (unsigned)((int64_t)g.P * g.H);
    const size_t bands_smem = (size_t)kFplaneWarps * g.W * g.Pin * ((dtype == VPINT_F64) ? 8 : 4);
    if (g.Pin >= 2 && stride[1] == 1 && stride[3] == g.Pin && bands_smem <= 100 * 1024) {
        const int n_patches = g.P / g.Pin;
        const unsigned bgrid = (unsigned)(((int64_t)n_patches * g.H + kFplaneWarps - 1) / kFplaneWarps);
        cudaError_t e;
        if (dtype == VPINT_F64) {
            e = cudaFuncSetAttribute(fplane_bands_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bands_smem);
            if (e != cudaSuccess) return e;
            fplane_bands_kernel<double><<<bgrid, kFplaneWarps * 32, bands_smem, st>>>(
                n_patches, g.Pin, g.H, g.W, g.pitch, (const double *)feat, stride[0], stride[2], fplane, bad_flag, extreme);
        } else {
            e = cudaFuncSetAttribute(fplane_bands_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bands_smem);
            if (e != cudaSuccess) return e;
            fplane_bands_kernel<float><<<bgrid, kFplaneWarps * 32, bands_smem, st>>>(
                n_patches, g.Pin, g.H, g.W, g.pitch, (const float *)feat, stride[0], stride[2], fplane, bad_flag, extreme);
        }
        VP_LAUNCH_CHECK();
        return cudaSuccess;
    }
    if (dtype == VPINT_F64)
        fplane_kernel<double><<<grid, 256, 0, st>>>(g.Pin, g.H, g.W, g.pitch, (const double *)feat, stride[0], stride[1],
                                                    stride[2], stride[3], fplane, bad_flag, extreme);
    else
        fplane_kernel<float><<<grid, 256, 0, st>>>(g.Pin, g.H, g.W, g.pitch, (const float *)feat, stride[0], stride[1],
                                                   stride[2], stride[3], fplane, bad_flag, extreme);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_planes_from_fplane(const Geom &g, const double *fplane, int storage, void *const *planes,
                                      cudaStream_t st, int64_t *launches) {
    const unsigned grid = (unsigned)((int64_t)g.P * g.H);
    if (storage == VPINT_F64)
        planes_from_fplane_kernel<double><<<grid, 256, 0, st>>>(g.H, g.W, g.pitch, g.row_offset, g.global_H, fplane,
                                                                (double *)planes[0], (double *)planes[1],
                                                                (double *)planes[2], (double *)planes[3]);
    else
        planes_from_fplane_kernel<float><<<grid, 256, 0, st>>>(g.H, g.W, g.pitch, g.row_offset, g.global_H, fplane,
                                                               (float *)planes[0], (float *)planes[1],
                                                               (float *)planes[2], (float *)planes[3]);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_priority(const Geom &g, int storage, void *const *planes, double beta, const double *beta_arr,
                            const double *bias, double *p_out, cudaStream_t st, int64_t *launches) {
    const unsigned grid = (unsigned)((int64_t)g.P * g.H);
    if (storage == VPINT_F64)
        priority_kernel<double><<<grid, 256, 0, st>>>(g.H, g.W, g.pitch, g.row_offset, g.global_H, (double *)planes[0],
                                                      (double *)planes[1], (double *)planes[2], (double *)planes[3],
                                                      beta, beta_arr, bias, p_out);
    else
        priority_kernel<float><<<grid, 256, 0, st>>>(g.H, g.W, g.pitch, g.row_offset, g.global_H, (float *)planes[0],
                                                     (float *)planes[1], (float *)planes[2], (float *)planes[3], beta,
                                                     beta_arr, bias, p_out);
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

cudaError_t launch_weights3d(const Geom &g, const void *feat, int dtype, const int64_t *stride, int F, int method,
                             int storage, void *const *planes, cudaStream_t st, int64_t *launches) {
    const unsigned grid = (unsigned)((int64_t)g.P * g.H);
    const int64_t sO = stride[0], sI = stride[1], sY = stride[2], sX = stride[3], sF = stride[4];
#define VP_W(TIN, TST)                                                                                         \
    weights_kernel<TIN, TST><<<grid, 256, 0, st>>>(g.Pin, g.H, g.W, g.wpitch, g.row_offset, g.global_H,        \
                                                   (const TIN *)feat, sO, sI, sY, sX, sF, F, method, 0, 0, 0.0, \
                                                   0.0, 5, (TST *)planes[0], (TST *)planes[1],                 \
                                                   (TST *)planes[2], (TST *)planes[3], (TST *)planes[4])
    if (dtype == VPINT_F64 && storage == VPINT_F64) VP_W(double, double);
    else if (dtype == VPINT_F32 && storage == VPINT_F64) VP_W(float, double);
    else if (dtype == VPINT_F64 && storage == VPINT_F32) VP_W(double, float);
    else VP_W(float, float);
#undef VP_W
    VP_LAUNCH_CHECK();
    return cudaSuccess;
}

}  // namespace vpint
