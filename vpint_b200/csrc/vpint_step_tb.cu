// A4 + A5 (local part), temporally blocked: KB Jacobi sweeps per launch on a 64x64 region.
//
// One CTA (512 threads) owns a (64-2*KB)^2 tile and stages the tile plus a KB-deep halo of the state
// into shared memory with ONE TMA load (cp.async.bulk.tensor, 3-D map {row element, row, problem},
// out-of-grid cells arrive as 0 = the reference's zero border, VPint/WP_MRP.py:291-304).  Each thread
// keeps a 1x8 column strip: its 8 current values and -- for unknown cells only -- the 4 direction
// weights stay in registers for all KB sweeps, so the 32 B/cell of weights are read from HBM once per
// KB sweeps instead of once per sweep.  Neighbour values move through a ping-pong pair of shared
// tiles (known cells are constants and live in both), one __syncthreads per sweep.  Halo cells are
// recomputed redundantly; after sweep j only cells >= j deep are valid, owned cells are KB deep.
// Per sweep the CTA records {sum|new-old|, sum old, NaN counts} over owned unknown cells
// (VPint/SD_MRP.py:139); the launch writes only owned unknown cells back.  A problem whose stop
// falls inside the block is replayed by the host loop with k_run < KB from the untouched input.
#include <cuda.h>

#include "vpint_internal.cuh"

namespace vpint {

namespace {

constexpr int kR = 64;            // region edge (cells)
constexpr int kThreads = 512;     // 16 warps: 2 across x 8 down, each thread 1 column x 8 rows
constexpr int kRows = 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "VP_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra VP_DONE_%=;\n"
        "bra VP_WAIT_%=;\n"
        "VP_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

template <typename T, bool PLANES, int KB>
__global__ void __launch_bounds__(kThreads, 1)
jacobi2d_tb_kernel(const StepArgs a, const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *sb0 = reinterpret_cast<T *>(smem_raw);
    T *sb1 = sb0 + kR * kR;
    __shared__ uint64_t bar;
    __shared__ double red[(kThreads / 32) * KB * kSumSlots];

    const TileRef tile = a.tiles[blockIdx.x];
    const int p = tile.p;
    const ProbState st = a.state[p];
    const int kr = st.k_run;
    if (kr == 0) return;
    const Geom &g = a.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ry = tile.y0 - KB, rx = tile.x0 - KB;      // region origin (may be negative)

    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, (uint32_t)(kR * kR * sizeof(T)));
        tma_load_3d(sb0, st.cur ? &map1 : &map0, rx, ry, p, &bar);
    }

    // ---- per-thread static data: mask bits, ownership, weights of unknown cells -----------------
    const int col = (warp & 1) * 32 + lane;
    const int r0 = (warp >> 1) * kRows;
    const int gx = rx + col;
    const int gy0 = ry + r0;
    const bool col_in = (gx >= 0) && (gx < g.L);
    const int own_y1 = min(tile.y0 + g.TH, g.own_row0 + g.own_rows);
    const bool col_own = (gx >= tile.x0) && (gx < min(tile.x0 + g.TW, g.L));
    const size_t plane = (size_t)g.H * g.pitch;
    const uint32_t *mrow = a.mask + (size_t)p * g.H * g.mpitch;

    unsigned mbits = 0, obits = 0;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int gy = gy0 + r;
        if (col_in && gy >= 0 && gy < g.H) {
            const unsigned bit = (mrow[(size_t)gy * g.mpitch + (gx >> 5)] >> (gx & 31)) & 1u;
            mbits |= bit << r;
            if (bit && col_own && gy >= tile.y0 && gy < own_y1) obits |= 1u << r;
        }
    }
    double w_up[kRows], w_rt[kRows], w_dn[kRows], w_lf[kRows];
    double gamma = 0.0;
    if (PLANES) {
        const T *pu = reinterpret_cast<const T *>(a.w[0]) + (size_t)p * plane;
        const T *pr = reinterpret_cast<const T *>(a.w[1]) + (size_t)p * plane;
        const T *pd = reinterpret_cast<const T *>(a.w[2]) + (size_t)p * plane;
        const T *pl = reinterpret_cast<const T *>(a.w[3]) + (size_t)p * plane;
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            w_up[r] = w_rt[r] = w_dn[r] = w_lf[r] = 0.0;
            if ((mbits >> r) & 1u) {
                const size_t off = (size_t)(gy0 + r) * g.pitch + gx;
                w_up[r] = (double)__ldg(pu + off);
                w_rt[r] = (double)__ldg(pr + off);
                w_dn[r] = (double)__ldg(pd + off);
                w_lf[r] = (double)__ldg(pl + off);
            }
        }
    } else {
        gamma = a.gamma[p];
    }
    const int ncx = 4 - (gx == 0 ? 1 : 0) - (gx == g.L - 1 ? 1 : 0);
    const int gyg0 = gy0 + g.row_offset;
    const double clip = a.clip;

    // ---- wait for the tile, seed registers and the second shared tile ---------------------------
    mbar_wait(&bar, 0);
    double v[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const T raw = sb0[(r0 + r) * kR + col];
        v[r] = (double)raw;
        if (!((mbits >> r) & 1u)) sb1[(r0 + r) * kR + col] = raw;   // constants live in both tiles
    }
    __syncthreads();

    double s_abs[KB], s_old[KB];
    unsigned c_abs = 0, c_old = 0;          // 4-bit NaN counters per sweep (<= 8 cells per thread)
#pragma unroll
    for (int j = 0; j < KB; ++j) s_abs[j] = s_old[j] = 0.0;

    T *cur = sb0, *nxt = sb1;
#pragma unroll
    for (int j = 0; j < KB; ++j) {
        if (j < kr) {
            if (mbits) {
                double up = (r0 > 0) ? (double)cur[(r0 - 1) * kR + col] : 0.0;
                const double dn_ext = (r0 + kRows < kR) ? (double)cur[(r0 + kRows) * kR + col] : 0.0;
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    const double c = v[r];
                    const double dn = (r < kRows - 1) ? v[r + 1] : dn_ext;
                    if ((mbits >> r) & 1u) {
                        const T *rowp = cur + (r0 + r) * kR;
                        const double lf = (col > 0) ? (double)rowp[col - 1] : 0.0;
                        const double rt = (col < kR - 1) ? (double)rowp[col + 1] : 0.0;
                        double nv;
                        if (PLANES) nv = fma(lf, w_lf[r], fma(dn, w_dn[r], fma(rt, w_rt[r], up * w_up[r])));
                        else nv = fma(lf, gamma, fma(dn, gamma, fma(rt, gamma, up * gamma)));
                        const int gyg = gyg0 + r;
                        const int nc = ncx - (gyg == 0 ? 1 : 0) - (gyg == g.global_H - 1 ? 1 : 0);
                        if (!a.no_nc) nv = (nc == 4) ? nv * 0.25 : nv / (double)nc;
                        if (nv > clip) nv = clip;
                        const T stored = (T)nv;
                        nv = (double)stored;
                        nxt[(r0 + r) * kR + col] = stored;
                        if ((obits >> r) & 1u) {
                            const double d = fabs(nv - c);
                            if (d != d) c_abs += 1u << (4 * j); else s_abs[j] += d;
                            if (c != c) c_old += 1u << (4 * j); else s_old[j] += c;
                        }
                        v[r] = nv;
                    }
                    up = c;
                }
            }
            __syncthreads();
            T *t = cur; cur = nxt; nxt = t;
        }
    }

    // ---- write owned unknown cells back -----------------------------------------------------------
    if (obits) {
        T *out = reinterpret_cast<T *>(a.buf[st.cur ^ 1]) + (size_t)p * plane;
#pragma unroll
        for (int r = 0; r < kRows; ++r)
            if ((obits >> r) & 1u) out[(size_t)(gy0 + r) * g.pitch + gx] = (T)v[r];
    }

    // ---- per-sweep sums: fixed-order block reduction ---------------------------------------------
#pragma unroll
    for (int j = 0; j < KB; ++j) {
        double q0 = warp_sum(s_abs[j]), q1 = warp_sum(s_old[j]);
        int n0 = warp_sum((int)((c_abs >> (4 * j)) & 15u)), n1 = warp_sum((int)((c_old >> (4 * j)) & 15u));
        if (lane == 0) {
            double *dst = red + (warp * KB + j) * kSumSlots;
            dst[0] = q0; dst[1] = q1; dst[2] = (double)n0; dst[3] = (double)n1;
        }
    }
    __syncthreads();
    if (tid < KB * kSumSlots) {
        double acc = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) acc += red[w * KB * kSumSlots + tid];
        a.partial[(size_t)blockIdx.x * g.KB * kSumSlots + tid] = acc;
    }
    __syncthreads();
    finish_tile(a, p, kr, red);           // red: (kThreads / 32) * KB * kSumSlots doubles >= kSumSlots * 16
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

template <typename T, bool PLANES>
cudaError_t launch_kb(const StepArgs &a, const CUtensorMap *maps, int64_t n_tiles, cudaStream_t st) {
    const size_t smem = 2 * kR * kR * sizeof(T);
#define VP_TB(KBV)                                                                                              \
    {                                                                                                           \
        auto kern = jacobi2d_tb_kernel<T, PLANES, KBV>;                                                         \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        if (e != cudaSuccess) return e;                                                                         \
        kern<<<(unsigned)n_tiles, kThreads, smem, st>>>(a, maps[0], maps[1]);                                   \
    }
    switch (a.g.KB) {
        case 2: VP_TB(2) break;
        case 4: VP_TB(4) break;
        case 6: VP_TB(6) break;
        case 8: VP_TB(8) break;
        default: return cudaErrorInvalidValue;
    }
#undef VP_TB
    return cudaGetLastError();
}

}  // namespace

void tb_tile_shape(int k_block, int *th, int *tw) {
    *th = kR - 2 * k_block;
    *tw = kR - 2 * k_block;
}

bool tb_supported(const Geom &g, int) { return g.ndim == 2 && (g.KB == 2 || g.KB == 4 || g.KB == 6 || g.KB == 8); }

// Two 3-D tensor maps (one per ping-pong buffer): dims {L, H, P}, box {64, 64, 1}, zero OOB fill.
cudaError_t tb_make_maps(const Geom &g, int storage, void *buf0, void *buf1, void *maps_out) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return cudaErrorNotSupported;
    CUtensorMap *maps = static_cast<CUtensorMap *>(maps_out);
    const size_t elem = (storage == VPINT_F64) ? 8 : 4;
    const CUtensorMapDataType dt = (storage == VPINT_F64) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const cuuint64_t dims[3] = {(cuuint64_t)g.L, (cuuint64_t)g.H, (cuuint64_t)g.P};
    const cuuint64_t strides[2] = {(cuuint64_t)g.pitch * elem, (cuuint64_t)g.H * g.pitch * elem};
    const cuuint32_t box[3] = {kR, kR, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    void *bufs[2] = {buf0, buf1};
    for (int i = 0; i < 2; ++i) {
        CUresult r = enc(&maps[i], dt, 3, bufs[i], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
    return cudaSuccess;
}

size_t tb_maps_bytes() { return 2 * sizeof(CUtensorMap); }

cudaError_t launch_step2d_tb(const StepArgs &a, const void *maps, int storage, int weight_mode, int64_t n_tiles,
                             cudaStream_t st, int64_t *launches) {
    if (n_tiles <= 0) return cudaSuccess;
    const CUtensorMap *m = static_cast<const CUtensorMap *>(maps);
    const bool planes = (weight_mode == VPINT_W_PLANES);
    cudaError_t e;
    if (storage == VPINT_F64) e = planes ? launch_kb<double, true>(a, m, n_tiles, st) : launch_kb<double, false>(a, m, n_tiles, st);
    else e = planes ? launch_kb<float, true>(a, m, n_tiles, st) : launch_kb<float, false>(a, m, n_tiles, st);
    if (e != cudaSuccess) return e;
    if (launches) ++*launches;
    return cudaSuccess;
}

}  // namespace vpint
