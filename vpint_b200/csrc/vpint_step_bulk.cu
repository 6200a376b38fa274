// A4 + A5 (local part) for large single-feature scenes: the ratio sweep of vpint_step_k1.cu
// (jacobi2d_k1_ratio4_kernel -- same cells per lane, same arithmetic, same per-warp partial sums, hence the same
// bits) fed through shared memory by the copy engine instead of by the warps' own loads.
//   r_new = mean of the neighbours' r  (P = f * r, VPint/WP_MRP.py:291-319 with exact single-feature weights);
//   sums over unknown owned cells of |f r_new - f r_old| and f r_old (VPint/SD_MRP.py:139).
//
// One persistent CTA per SM walks the active-tile list (tile i -> CTA i mod grid).  A producer warp per ring slot reads
// its tile's mask words, works out which state / feature rows the tile can touch, and issues one cp.async.bulk per
// row into the slot's tile buffer (mbarrier complete_tx); eight consumer warps take the tiles in order,
// compute from shared memory, store the new values of the unknown cells to global memory and hand the buffer
// back (second mbarrier).  Two tiles (~130 KB) are in flight per SM while a third is being computed, which is what
// the HBM latency needs; the register-fed kernel keeps at most ~45 KB per SM in flight at 24 warps.
#include <cmath>
#include <cstdlib>

#include "vpint_internal.cuh"

namespace vpint {

namespace {

constexpr int kTW = 128, kTH = 32;          // tile of the 2-D k1 kernels (vpint_step_k1.cu)
// Ring slots: 3 with the fp64 feature plane (3 x 69 KB), 4 with its float32 copy (4 x 53 KB; StepArgs::f32plane, written by
// the ratio-direct load when every feature is exactly a float32 value -- uint16 rasters always are).
template <typename FT> struct Slots { static constexpr int n = sizeof(FT) == 4 ? 4 : 3; };
constexpr int kSW = kTW + 4;                // state row in shared memory: columns x0 - 2 .. x0 + 129 (16-byte aligned start)
constexpr int kSRows = kTH + 2;             // grid rows y0 - 1 .. y0 + 32
constexpr int kConsumerWarps = 8;
// one producer warp per ring slot: the latency chains tile -> state -> mask -> copies overlap
template <typename FT> struct Threads { static constexpr int n = (kConsumerWarps + Slots<FT>::n) * 32; };

struct __align__(16) TileMeta {
    int active, p, y0, x0;
    int cur, pad0, pad1, pad2;
};

template <typename FT>
struct __align__(128) Stage {
    double st[kSRows][kSW];                 // 35 904 B
    FT f[kTH][kTW];                         // 32 768 B (fp64) / 16 384 B (float32)
    unsigned long long mw[kTH][2];          // mask words of the tile's rows (0 below the owned band)
    TileMeta meta;
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bar_arrive_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra D_%=;\n"
        "bra W_%=;\n"
        "D_%=:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: sleep, do not spin
}
// global -> shared copy of `bytes` (multiple of 16, both ends 16-byte aligned) by the copy engine
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void lds2(const double *p, double &a, double &b) {
    const double2 v = *reinterpret_cast<const double2 *>(p);
    a = v.x; b = v.y;
}
__device__ __forceinline__ void lds2(const float *p, double &a, double &b) {
    const float2 v = *reinterpret_cast<const float2 *>(p);
    a = (double)v.x; b = (double)v.y;
}

// ---- producer warps -----------------------------------------------------------------------------------------
// Producer `slot` owns ring slot `slot`: the tiles it = slot, slot + kStages, ... of this CTA.  The tile reference, the
// problem state and the mask words are fetched into registers BEFORE the wait for the slot, so their latency runs
// while the consumers still work on the slot's previous tile.
template <typename FT>
__device__ __forceinline__ void produce(const StepArgs &a, Stage<FT> &sg, uint64_t *full, uint64_t *empty, int n_tiles, int slot, int lane) {
    constexpr int kStages = Slots<FT>::n;
    const Geom &g = a.g;
    const size_t plane = (size_t)g.H * g.pitch;
    const int mp64 = g.mpitch >> 1;
    unsigned use = 0;
    for (int64_t t = (int64_t)blockIdx.x + (int64_t)slot * gridDim.x; t < n_tiles; t += (int64_t)kStages * gridDim.x, ++use) {
        const TileRef tile = a.tiles[t];
        const int p = tile.p;
        const ProbState ps = a.state[p];
        const bool live = ps.k_run != 0;
        const int y_lim = min(tile.y0 + kTH, g.own_row0 + g.own_rows);
        unsigned long long w0 = 0ull, w1 = 0ull;
        if (live) {
            const unsigned long long *m64 =
                reinterpret_cast<const unsigned long long *>(a.mask + (size_t)p * g.H * g.mpitch) + (tile.x0 >> 6);
            const int ym = tile.y0 + lane;
            if (ym < y_lim) {
                w0 = m64[(size_t)ym * mp64];
                if (tile.x0 + 64 < g.L) w1 = m64[(size_t)ym * mp64 + 1];
            }
        }
        const unsigned rowany = __ballot_sync(0xffffffffu, (w0 | w1) != 0ull);
        if (use > 0) bar_wait(empty, (use - 1) & 1u);              // the consumers are done with the tile that was here
        if (!live) {                                                // problem already stopped: nothing to stage
            if (lane == 0) {
                sg.meta.active = 0;
                bar_arrive(full);
            }
            continue;
        }
        sg.mw[lane][0] = w0;
        sg.mw[lane][1] = w1;
        if (lane == 0) {
            sg.meta.active = 1; sg.meta.p = p; sg.meta.y0 = tile.y0; sg.meta.x0 = tile.x0; sg.meta.cur = ps.cur;
        }
        const unsigned long long any64 = rowany;
        const unsigned long long need = any64 | (any64 << 1) | (any64 << 2);          // state row i = grid row y0 - 1 + i
        const int xs = max(tile.x0 - 2, 0), xe = min(tile.x0 + kTW + 2, g.pitch);
        const unsigned sbytes = (unsigned)(xe - xs) * 8u;
        const unsigned fbytes = (unsigned)(min(tile.x0 + kTW, g.pitch) - tile.x0) * (unsigned)sizeof(FT);
        const double *in = reinterpret_cast<const double *>(a.buf[ps.cur]) + (size_t)p * plane;
        const FT *fpl = (sizeof(FT) == 4 ? reinterpret_cast<const FT *>(a.f32plane) : reinterpret_cast<const FT *>(a.w[4])) + (size_t)p * plane;
        // rows of this lane: state rows lane and lane + 32 (two lanes only), feature row lane
        bool s_on[2];
        unsigned mine = 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int i = lane + 32 * k;
            const int y = tile.y0 - 1 + i, yg = y + g.row_offset;
            s_on[k] = i < kSRows && ((need >> i) & 1ull) && y >= 0 && y < g.H && yg >= 0 && yg < g.global_H;
            mine += s_on[k] ? sbytes : 0u;
        }
        const bool f_on = (rowany >> lane) & 1u;
        mine += f_on ? fbytes : 0u;
        unsigned total = mine;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        __syncwarp();                                               // mask words and meta written before the arrive releases them
        if (lane == 0) bar_arrive_tx(full, total);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (s_on[k]) {
                const int i = lane + 32 * k;
                bulk_g2s(&sg.st[i][xs - (tile.x0 - 2)], in + (size_t)(tile.y0 - 1 + i) * g.pitch + xs, sbytes, full);
            }
        }
        if (f_on) bulk_g2s(&sg.f[lane][0], fpl + (size_t)(tile.y0 + lane) * g.pitch + tile.x0, fbytes, full);
    }
}

// ---- consumer warps -----------------------------------------------------------------------------------------
// Update of the two cells of one lane in one row, shared by both paths below: the arithmetic of ratio_half_strip
// (vpint_step_k1.cu) term for term.
template <bool CLIP>
__device__ __forceinline__ void commit_pair(double clip, double n0, double n1, double c0, double c1, double f0, double f1, unsigned mb,
                                            double *__restrict__ o_row, double &s_abs, double &s_old, int &c_abs, int &c_old) {
    double w0 = __dmul_rn(f0, n0), w1 = __dmul_rn(f1, n1);       // products rounded as the reference rounds them (no fma)
    if (CLIP) {
        if (w0 > clip) { w0 = clip; n0 = clip / f0; }
        if (w1 > clip) { w1 = clip; n1 = clip / f1; }
    }
    if (mb & 1u) {
        const double o = __dmul_rn(f0, c0), d = fabs(w0 - o);
        if (d != d) ++c_abs; else s_abs += d;
        if (o != o) ++c_old; else s_old += o;
        o_row[0] = n0;
    }
    if (mb & 2u) {
        const double o = __dmul_rn(f1, c1), d = fabs(w1 - o);
        if (d != d) ++c_abs; else s_abs += d;
        if (o != o) ++c_old; else s_old += o;
        o_row[1] = n1;
    }
}

// The same without the per-cell NaN tests: a NaN term turns the running sum into NaN, which the caller sees once per
// tile and answers by redoing the tile with commit_pair (same stores, exact nanmean bookkeeping).  Without a NaN term
// both forms add the same numbers in the same order.
template <bool CLIP>
__device__ __forceinline__ void commit_pair_fast(double clip, double n0, double n1, double c0, double c1, double f0, double f1,
                                                 unsigned mb, double *__restrict__ o_row, double &s_abs, double &s_old) {
    double w0 = __dmul_rn(f0, n0), w1 = __dmul_rn(f1, n1);
    if (CLIP) {
        if (w0 > clip) { w0 = clip; n0 = clip / f0; }
        if (w1 > clip) { w1 = clip; n1 = clip / f1; }
    }
    if (mb & 1u) {
        const double o = __dmul_rn(f0, c0);
        s_abs += fabs(w0 - o);
        s_old += o;
        o_row[0] = n0;
    }
    if (mb & 2u) {
        const double o = __dmul_rn(f1, c1);
        s_abs += fabs(w1 - o);
        s_old += o;
        o_row[1] = n1;
    }
}

// Four rows of one warp (64 columns, two per lane) of a tile with no cell on the border of the global grid: every
// neighbour exists and counts, every row the producer skipped belongs to cells that are neither stored nor summed,
// so the loads need no predicates.  sp = &st[R][cx] (state row above the first of the four), fp = &f[R][column].
template <bool CLIP, typename FT>
__device__ __forceinline__ void half_strip_interior(double clip, const double *__restrict__ sp, const FT *__restrict__ fp,
                                                    double *__restrict__ o_base, int pitch, unsigned any, unsigned mine, double &s_abs,
                                                    double &s_old) {
    double v0[6], v1[6], lf[4], rt[4], f0[4], f1[4];
#pragma unroll
    for (int i = 0; i < 6; ++i) lds2(sp + i * kSW, v0[i], v1[i]);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        lds2(fp + r * kTW, f0[r], f1[r]);
        lf[r] = sp[(r + 1) * kSW - 1];
        rt[r] = sp[(r + 1) * kSW + 2];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (!((any >> r) & 1u)) continue;                       // warp-uniform
        const double n0 = (((v0[r] + v1[r + 1]) + v0[r + 2]) + lf[r]) * 0.25;
        const double n1 = (((v1[r] + rt[r]) + v1[r + 2]) + v0[r + 1]) * 0.25;
        commit_pair_fast<CLIP>(clip, n0, n1, v0[r + 1], v1[r + 1], f0[r], f1[r], (mine >> (2 * r)) & 3u, o_base + (size_t)r * pitch,
                               s_abs, s_old);
    }
}

// The same four rows of a tile that touches the border of the global grid (or of the row shard's buffer): rows and
// columns outside read as 0 and the neighbour count drops, as in ratio_half_strip<false>.
template <bool CLIP, typename FT>
__device__ __forceinline__ void half_strip_border(const StepArgs &a, const Geom &g, const Stage<FT> &sg, double *__restrict__ out, int x0,
                                                  int y0, int xo, int lane, int R, unsigned any, unsigned mine, double &s_abs,
                                                  double &s_old, int &c_abs, int &c_old) {
    const int x = x0 + xo + 2 * lane;
    const int cx = xo + 2 * lane + 2;                          // column of x in a shared-memory state row
    const unsigned need = (any << 2) | (any << 1) | any;
    double v0[6], v1[6], lf[4], rt[4], f0[4], f1[4];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int y = y0 + R - 1 + i, yg = y + g.row_offset;
        v0[i] = 0.0; v1[i] = 0.0;
        if (((need >> i) & 1u) && yg >= 0 && yg < g.global_H && y >= 0 && y < g.H) lds2(&sg.st[R + i][cx], v0[i], v1[i]);
    }
    const bool has_left = x > 0, has_right = x + 2 < g.L;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        f0[r] = 1.0; f1[r] = 1.0; lf[r] = 0.0; rt[r] = 0.0;
        if ((any >> r) & 1u) {
            lds2(&sg.f[R + r][xo + 2 * lane], f0[r], f1[r]);
            if (has_left) lf[r] = sg.st[R + r + 1][cx - 1];
            if (has_right) rt[r] = sg.st[R + r + 1][cx + 2];
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (!((any >> r) & 1u)) continue;                       // warp-uniform
        const int y = y0 + R + r;
        double n0 = ((v0[r] + v1[r + 1]) + v0[r + 2]) + lf[r];
        double n1 = ((v1[r] + rt[r]) + v1[r + 2]) + v0[r + 1];
        const int yg = y + g.row_offset;
        const int ncy = 4 - (yg == 0 ? 1 : 0) - (yg == g.global_H - 1 ? 1 : 0);
        const int nc0 = ncy - (x == 0 ? 1 : 0) - (x == g.L - 1 ? 1 : 0);
        const int nc1 = ncy - (x + 1 == g.L - 1 ? 1 : 0);
        n0 = (nc0 == 4) ? n0 * 0.25 : n0 / (double)nc0;
        n1 = (nc1 == 4) ? n1 * 0.25 : n1 / (double)nc1;
        commit_pair<CLIP>(a.clip, n0, n1, v0[r + 1], v1[r + 1], f0[r], f1[r], (mine >> (2 * r)) & 3u, out + (size_t)y * g.pitch + x,
                          s_abs, s_old, c_abs, c_old);
    }
}

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int wsum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool CLIP, typename FT>
__device__ __forceinline__ void consume(const StepArgs &a, Stage<FT> *stages, uint64_t *full, uint64_t *empty, int n_tiles, int warp,
                                        int lane) {
    const Geom &g = a.g;
    const size_t plane = (size_t)g.H * g.pitch;
    constexpr int kStages = Slots<FT>::n;
    const int xo = (warp & 1) * 64, yo = (warp >> 1) * 8;
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int s = it % kStages;
        const unsigned use = (unsigned)(it / kStages);
        const Stage<FT> &sg = stages[s];
        bar_wait(&full[s], use & 1u);
        const TileMeta m = sg.meta;
        if (!m.active) {
            if (lane == 0) bar_arrive(&empty[s]);
            continue;
        }
        double *__restrict__ out = reinterpret_cast<double *>(a.buf[m.cur ^ 1]) + (size_t)m.p * plane;
        const int y_lim = min(m.y0 + kTH, g.own_row0 + g.own_rows);
        double s_abs = 0.0, s_old = 0.0;
        int c_abs = 0, c_old = 0;
        if (m.x0 + xo < g.L && m.y0 + yo < y_lim) {
            unsigned any = 0, mine = 0;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const unsigned long long w = sg.mw[yo + r][warp & 1];
                any |= (w != 0ull) ? (1u << r) : 0u;
                mine |= ((unsigned)(w >> (2 * lane)) & 3u) << (2 * r);
            }
            const bool interior = (m.y0 + g.row_offset > 0) && (m.y0 + kTH + g.row_offset < g.global_H) && (m.x0 > 0) &&
                                  (m.x0 + kTW < g.L);
            const int cx = xo + 2 * lane + 2;
            const double clip = a.clip;
            bool careful = !interior;
            if (interior) {
                // straight code for both halves: every offset is a compile-time constant from these bases
                const double *sp = &sg.st[yo][cx];
                const FT *fp = &sg.f[yo][xo + 2 * lane];
                double *ob = out + (size_t)(m.y0 + yo) * g.pitch + (m.x0 + xo + 2 * lane);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const unsigned any_h = (any >> (4 * h)) & 15u, mine_h = (mine >> (8 * h)) & 255u;
                    if (any_h == 0) continue;                       // warp-uniform
                    half_strip_interior<CLIP, FT>(clip, sp + 4 * h * kSW, fp + 4 * h * kTW, ob + (size_t)(4 * h) * g.pitch, g.pitch, any_h, mine_h,
                                              s_abs, s_old);
                }
                // a NaN term (non-finite data) poisons the sums: redo the tile with the exact nanmean bookkeeping
                careful = __any_sync(0xffffffffu, (s_abs != s_abs) || (s_old != s_old));
                if (careful) { s_abs = 0.0; s_old = 0.0; }
            }
            if (careful) {
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    const unsigned any_h = (any >> (4 * h)) & 15u, mine_h = (mine >> (8 * h)) & 255u;
                    if (any_h == 0) continue;
                    half_strip_border<CLIP, FT>(a, g, sg, out, m.x0, m.y0, xo, lane, yo + 4 * h, any_h, mine_h, s_abs, s_old, c_abs, c_old);
                }
            }
        }
        __syncwarp();                                               // every lane has read its part of the buffer
        if (lane == 0) bar_arrive(&empty[s]);
        const double q0 = wsum(s_abs), q1 = wsum(s_old);
        int q2 = 0, q3 = 0;
        if (__any_sync(0xffffffffu, (c_abs | c_old) != 0)) { q2 = wsum(c_abs); q3 = wsum(c_old); }
        if (lane == 0) {
            double *dst = a.partial8 + ((size_t)t * 8 + warp) * kSumSlots;
            dst[0] = q0; dst[1] = q1; dst[2] = (double)q2; dst[3] = (double)q3;
        }
    }
}

template <bool CLIP, typename FT>
__global__ void __launch_bounds__(Threads<FT>::n, 1) jacobi2d_ratio_bulk_kernel(const StepArgs a, int n_tiles) {
    constexpr int kStages = Slots<FT>::n;
    extern __shared__ __align__(128) unsigned char bulk_raw[];
    __shared__ __align__(8) uint64_t full[kStages], empty[kStages];
    Stage<FT> *stages = reinterpret_cast<Stage<FT> *>(bulk_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            bar_init(&full[s], 1);
            bar_init(&empty[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp >= kConsumerWarps) {
        const int slot = warp - kConsumerWarps;
        produce<FT>(a, stages[slot], &full[slot], &empty[slot], n_tiles, slot, lane);
    } else {
        consume<CLIP, FT>(a, stages, full, empty, n_tiles, warp, lane);
    }
}

template <bool CLIP, typename FT>
cudaError_t launch_bulk(const StepArgs &a, int n_tiles, int sms, bool first, cudaStream_t st) {
    const size_t smem = sizeof(Stage<FT>) * Slots<FT>::n;
    if (first) {
        const cudaError_t e = cudaFuncSetAttribute(jacobi2d_ratio_bulk_kernel<CLIP, FT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
    jacobi2d_ratio_bulk_kernel<CLIP, FT><<<grid, Threads<FT>::n, smem, st>>>(a, n_tiles);
    return cudaGetLastError();
}

}  // namespace

bool step2d_ratio_bulk_supported(const StepArgs &a) {
    return a.partial8 != nullptr && a.g.TW == kTW && a.g.TH == kTH && a.g.ndim == 2;
}

cudaError_t launch_step2d_ratio_bulk(const StepArgs &a, int64_t n_tiles, cudaStream_t st, int64_t *launches) {
    if (n_tiles <= 0) return cudaSuccess;
    static int sms_of[64] = {0};                       // per device: the shared-memory attribute belongs to the device's context
    static unsigned char attr_done[64] = {0};          // bit per kernel variant
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (sms_of[dev] == 0) {
        int n = 0;
        e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        sms_of[dev] = n > 0 ? n : 148;
    }
    const bool clip = a.clip < INFINITY, f32 = a.f32plane != nullptr;       // clip_val = inf is the default
    const int variant = (clip ? 1 : 0) | (f32 ? 2 : 0);
    const bool first = !((attr_done[dev] >> variant) & 1);
    const int nt = (int)n_tiles, sms = sms_of[dev];
    if (f32) e = clip ? launch_bulk<true, float>(a, nt, sms, first, st) : launch_bulk<false, float>(a, nt, sms, first, st);
    else e = clip ? launch_bulk<true, double>(a, nt, sms, first, st) : launch_bulk<false, double>(a, nt, sms, first, st);
    if (e != cudaSuccess) return e;
    attr_done[dev] |= (unsigned char)(1u << variant);
    if (launches) ++*launches;
    return cudaSuccess;
}

}  // namespace vpint
