// A4 + A5 with the whole problem resident on chip: one thread-block CLUSTER per problem.
//
// Small grids (a 256x256 VPint2 band, a 100x100 SD_SMRP grid) fit in the shared memory of a few SMs.
// A cluster of C CTAs splits the rows of one problem, keeps its state in shared memory for ALL sweeps
// of run() and exchanges boundary rows through distributed shared memory, so a problem costs one read
// and one write of its unknown cells in HBM no matter how many sweeps it needs, and problems never wait
// for one another (each cluster pulls the next problem from a queue when its own has stopped).
//
// Two weight forms are supported:
//   SCALAR  SD_SMRP: every weight is gamma (VPint/SD_MRP.py:87).
//   RATIO   WP_SMRP 'exact' weights with ONE feature layer (the VPint2 case, VPint/VPint2.py:105):
//           w_dir = f[i,j] / f[nb] (VPint/WP_MRP.py:161-167), so
//               new[i,j] = (sum_nb w_nb * P[nb]) / nc = f[i,j] * (sum_nb P[nb] / f[nb]) / nc .
//           The kernel iterates r = P / f:  r_new = (sum of in-grid neighbour r) / nc, and forms
//           P = f * r only where the reference's stopping rule and clip need it.  Same iteration in
//           exact arithmetic; in fp64 it differs from the reference by a few ulp per sweep.
// Per sweep every CTA does: compute new values of its unknown cells from shared memory (registers hold
// them), block barrier, store them (own rows + the neighbour CTAs' halo rows through DSMEM) and push its
// partial {sum|new-old|, sum old, NaN counts} to every CTA of the cluster, ONE cluster barrier, then
// every CTA adds the C partials in rank order and takes the same stop decision
// (delta = nanmean|new-old| / nanmean(old) <= threshold, commit-then-break: VPint/SD_MRP.py:138-147,
// VPint/WP_MRP.py:351-362).  Halo rows and partials are double buffered by sweep parity, which is what
// makes one cluster barrier per sweep sufficient.
#include <cooperative_groups.h>

#include "vpint_internal.cuh"

namespace cg = cooperative_groups;

namespace vpint {

namespace {

constexpr int kThreads = kResidentThreads;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxPer = kResidentMaxPer;   // unknown cells per thread, held in registers across the barrier
constexpr int kMaxC = 8;                   // portable cluster size

struct SmemLayout {
    double *plane;       // [R][SP] owned rows, columns 0 and SP-1 are the zero border
    double *halo_up;     // [2][SP]  row above the owned band, by sweep parity
    double *halo_dn;     // [2][SP]  row below
    double *fcomp;       // [fcap]   feature of the first fcap listed cells (RATIO)
    unsigned short *list;  // [n]    shared-memory offset (ly * SP + x + 1) of every unknown owned cell, row-major
};

__device__ __forceinline__ SmemLayout carve(unsigned char *raw, int R, int SP, int fcap) {
    SmemLayout s;
    s.plane = reinterpret_cast<double *>(raw);
    s.halo_up = s.plane + (size_t)R * SP;
    s.halo_dn = s.halo_up + 2 * SP;
    s.fcomp = s.halo_dn + 2 * SP;
    s.list = reinterpret_cast<unsigned short *>(s.fcomp + fcap);
    return s;
}

template <bool RATIO>
__global__ void __launch_bounds__(kThreads, 1) resident2d_kernel(const ResidentArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red[kWarps][kSumSlots];
    __shared__ double cta_part[2][kMaxC][kSumSlots];
    __shared__ int wcount[kWarps];
    __shared__ int next_problem;

    if (a.bad_flag && *a.bad_flag) return;             // grid-uniform: leave everything to the tiled path
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int C = a.C, R = a.R, SP = a.SP, fcap = a.fcap;
    const Geom &g = a.g;
    const int W = g.L, H = g.H;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const SmemLayout sm = carve(smem_raw, R, SP, fcap);
    int *next_remote = cluster.map_shared_rank(&next_problem, 0);
    double *up_halo_remote = (crank > 0) ? cluster.map_shared_rank(sm.halo_dn, crank - 1) : nullptr;     // my first row -> upper CTA's lower halo
    double *dn_halo_remote = (crank < C - 1) ? cluster.map_shared_rank(sm.halo_up, crank + 1) : nullptr; // my last row -> lower CTA's upper halo

    const int gy0 = crank * R;                         // first owned grid row
    const int n_rows = max(0, min(R, H - gy0));
    const size_t plane_elems = (size_t)H * g.pitch;
    const int nwords = (W + 31) >> 5;
    const double n_cells = g.n_cells;

    for (;;) {
        // ---- next problem from the queue (cluster-uniform) ------------------------------------
        if (crank == 0 && tid == 0) next_problem = atomicAdd(a.queue, 1);
        cluster.sync();
        const int p = *next_remote;
        cluster.sync();                                // everyone has read before rank 0 overwrites or exits
        if (p >= g.P) break;
        const ProbState st = a.state[p];
        if (st.status != 0) continue;                  // nothing to run (cluster-uniform)

        double *grid = reinterpret_cast<double *>(a.buf[st.cur]) + (size_t)p * plane_elems;
        const double *fp = RATIO ? a.fplane + (size_t)p * plane_elems : nullptr;
        const uint32_t *mrow = a.mask + (size_t)p * H * g.mpitch;
        const double gamma = RATIO ? 1.0 : a.gamma[p];

        // ---- stage the owned rows and both halo rows (r = P / f in RATIO form) ----------------
        for (int row = warp; row < n_rows + 2; row += kWarps) {
            const int gy = gy0 - 1 + row;              // row 0 = upper halo, n_rows + 1 = lower halo
            double *dst = (row == 0) ? sm.halo_up : (row == n_rows + 1 ? sm.halo_dn : sm.plane + (size_t)(row - 1) * SP);
            const bool in_grid = (gy >= 0) && (gy < H);
            for (int lx = lane; lx < SP; lx += 32) {
                double v = 0.0;
                const int x = lx - 1;
                if (in_grid && x >= 0 && x < W) {
                    v = grid[(size_t)gy * g.pitch + x];
                    if (RATIO) v = v / fp[(size_t)gy * g.pitch + x];
                }
                dst[lx] = v;
                if (row == 0 || row == n_rows + 1) dst[SP + lx] = v;   // second parity copy of a halo row
            }
        }

        // ---- compact row-major list of the unknown owned cells --------------------------------
        const int rows_per_warp = (n_rows + kWarps - 1) / kWarps;
        const int wr0 = min(warp * rows_per_warp, n_rows), wr1 = min(wr0 + rows_per_warp, n_rows);
        int cnt = 0;
        for (int ly = wr0; ly < wr1; ++ly)
            for (int w = lane; w < nwords; w += 32) cnt += __popc(mrow[(size_t)(gy0 + ly) * g.mpitch + w]);
        cnt = warp_sum(cnt);
        if (lane == 0) wcount[warp] = cnt;
        __syncthreads();
        int base = 0, n = 0;
        for (int w = 0; w < kWarps; ++w) {
            if (w < warp) base += wcount[w];
            n += wcount[w];
        }
        for (int ly = wr0; ly < wr1; ++ly) {
            for (int w0 = 0; w0 < nwords; w0 += 32) {
                const int w = w0 + lane;
                unsigned word = (w < nwords) ? mrow[(size_t)(gy0 + ly) * g.mpitch + w] : 0u;
                const int pc = __popc(word);
                int incl = pc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                int pos = base + incl - pc;
                base += __shfl_sync(0xffffffffu, incl, 31);
                while (word) {
                    const int b = __ffs(word) - 1;
                    word &= word - 1;
                    const int x = (w << 5) + b;
                    sm.list[pos] = (unsigned short)(ly * SP + x + 1);
                    if (RATIO && pos < fcap) sm.fcomp[pos] = fp[(size_t)(gy0 + ly) * g.pitch + x];
                    ++pos;
                }
            }
        }
        __syncthreads();

        // ---- per-thread static facts about its cells (entry i = tid + k * kThreads) -----------
        const int K = (n > tid) ? (n - tid + kThreads - 1) / kThreads : 0;
        unsigned top_bits = 0, bot_bits = 0, edge_bits = 0;
        for (int k = 0; k < K; ++k) {
            const int off = sm.list[tid + k * kThreads];
            const int ly = off / SP, x = off - ly * SP - 1, gy = gy0 + ly;
            if (ly == 0) top_bits |= 1u << k;
            if (ly == n_rows - 1) bot_bits |= 1u << k;
            if (gy == 0 || gy == H - 1 || x == 0 || x == W - 1) edge_bits |= 1u << k;
        }
        const int bot_base = (n_rows - 1) * SP;
        const double *K0 = a.known0 + (size_t)p * kSumSlots, *K1 = a.known + (size_t)p * kSumSlots;
        const double k0[kSumSlots] = {K0[0], K0[1], K0[2], K0[3]};
        const double k1[kSumSlots] = {K1[0], K1[1], K1[2], K1[3]};
        cluster.sync();                                // lists and planes of every CTA are in place

        // ---- sweeps -----------------------------------------------------------------------------
        int it = 0;
        int par = 0;
        bool bailed = false;
        bool fresh = (st.fresh != 0);
        for (; it < a.max_iter;) {
            const double *hu = sm.halo_up + par * SP, *hd = sm.halo_dn + par * SP;
            double nr[kMaxPer];
            double s_abs = 0.0, s_old = 0.0;
            int c_abs = 0, c_old = 0;
#pragma unroll
            for (int k = 0; k < kMaxPer; ++k) {
                if (k < K) {
                    const int i = tid + k * kThreads;
                    const int off = sm.list[i];
                    const double c = sm.plane[off];
                    const double lf = sm.plane[off - 1], rt = sm.plane[off + 1];
                    const double up = ((top_bits >> k) & 1u) ? hu[off] : sm.plane[off - SP];
                    const double dn = ((bot_bits >> k) & 1u) ? hd[off - bot_base] : sm.plane[off + SP];
                    double s;
                    if (RATIO) s = ((up + rt) + dn) + lf;
                    else s = fma(lf, gamma, fma(dn, gamma, fma(rt, gamma, up * gamma)));
                    double rn;
                    if ((edge_bits >> k) & 1u) {
                        const int ly = off / SP, x = off - ly * SP - 1, gy = gy0 + ly;
                        const int nc = 4 - (gy == 0 ? 1 : 0) - (gy == H - 1 ? 1 : 0) - (x == 0 ? 1 : 0) - (x == W - 1 ? 1 : 0);
                        rn = s / (double)nc;
                    } else {
                        rn = s * 0.25;
                    }
                    double vn = rn, vo = c;
                    if (RATIO) {
                        const double f = (i < fcap) ? sm.fcomp[i]
                                                    : fp[(size_t)(gy0 + off / SP) * g.pitch + (off - (off / SP) * SP - 1)];
                        vn = f * rn;
                        vo = f * c;
                        if (vn > a.clip) { vn = a.clip; rn = a.clip / f; }
                    } else if (vn > a.clip) {
                        vn = a.clip; rn = a.clip;
                    }
                    const double d = fabs(vn - vo);
                    if (d != d) ++c_abs; else s_abs += d;
                    if (vo != vo) ++c_old; else s_old += vo;
                    nr[k] = rn;
                }
            }
            {
                const double q0 = warp_sum(s_abs), q1 = warp_sum(s_old);
                const int n0 = warp_sum(c_abs), n1 = warp_sum(c_old);
                if (lane == 0) { red[warp][0] = q0; red[warp][1] = q1; red[warp][2] = (double)n0; red[warp][3] = (double)n1; }
            }
            __syncthreads();                           // every read of the old state is done
            const int npar = par ^ 1;
#pragma unroll
            for (int k = 0; k < kMaxPer; ++k) {
                if (k < K) {
                    const int off = sm.list[tid + k * kThreads];
                    sm.plane[off] = nr[k];
                    if (((top_bits >> k) & 1u) && up_halo_remote) up_halo_remote[npar * SP + off] = nr[k];
                    if (((bot_bits >> k) & 1u) && dn_halo_remote) dn_halo_remote[npar * SP + off - bot_base] = nr[k];
                }
            }
            if (warp == 0) {
                double part = 0.0;
                if (lane < kSumSlots) {
                    for (int w = 0; w < kWarps; ++w) part += red[w][lane];
                }
                // lane (q * 4 + slot) stores slot of this CTA's partial into CTA q
                const int q = lane >> 2, slot = lane & 3;
                const double val = __shfl_sync(0xffffffffu, part, slot);
                if (q < C) cluster.map_shared_rank(&cta_part[par][crank][slot], q)[0] = val;
            }
            cluster.sync();                            // new state, halos and partials visible cluster-wide
            double S[kSumSlots] = {0.0, 0.0, 0.0, 0.0};
            for (int q = 0; q < C; ++q) {
#pragma unroll
                for (int i = 0; i < kSumSlots; ++i) S[i] += cta_part[par][q][i];
            }
            // Anything non-finite (NaN start state, overflow of f * r where the reference's w * P would
            // already have overflowed differently, inf inputs): leave this problem, untouched, to the tiled
            // plane path, which reproduces the reference's NaN / inf propagation exactly.
            if (!(S[2] == 0.0 && S[3] == 0.0 && fabs(S[0]) <= 1.7976931348623157e308 && fabs(S[1]) <= 1.7976931348623157e308)) {
                bailed = true;
                break;
            }
            const double *KK = fresh ? k0 : k1;
            const double sa = S[0] + KK[0], so = S[1] + KK[1];
            const double na = n_cells - (S[2] + KK[2]), no = n_cells - (S[3] + KK[3]);
            const double delta = (sa / na) / (so / no);
            if (a.delta_out && crank == 0 && tid == 0) a.delta_out[(size_t)p * a.max_iter + it] = delta;
            fresh = false;
            par = npar;
            ++it;
            if (a.auto_terminate && delta <= a.threshold) break;
        }

        // ---- write the unknown cells back (P = f * r), close the problem ------------------------
        for (int i = tid; i < n && !bailed; i += kThreads) {
            const int off = sm.list[i];
            const int ly = off / SP, x = off - ly * SP - 1;
            const size_t go = (size_t)(gy0 + ly) * g.pitch + x;
            double v = sm.plane[off];
            if (RATIO) v = ((i < fcap) ? sm.fcomp[i] : fp[go]) * v;
            grid[go] = v;
        }
        if (crank == 0 && tid == 0 && !bailed) {
            ProbState ns = st;
            ns.iters_done = st.iters_done + it;
            ns.status = 1; ns.k_run = 0; ns.replay = 0; ns.fresh = 0;
            a.state[p] = ns;
            atomicSub(a.n_active, 1);
        }
        __syncthreads();                               // list / plane reuse by the next problem
    }
}

}  // namespace

// Smallest portable cluster that holds one problem on chip; returns false if none does.
bool resident_plan(const Geom &g, bool ratio, int smem_limit, ResidentPlan *out) {
    if (g.ndim != 2 || g.H < 1 || g.L < 1) return false;
    const int W = g.L, SP = W + 2;
    const int static_bytes = 2048;   // red, cta_part, wcount, next_problem (+ slack)
    for (int C = 1; C <= kMaxC; C *= 2) {
        const int R = (g.H + C - 1) / C;
        if ((g.H + R - 1) / R != C) continue;                       // every CTA must own at least one row
        if ((int64_t)R * W > (int64_t)kThreads * kMaxPer) continue; // register budget for the new values
        if ((int64_t)R * SP > 65535) continue;                      // 16-bit shared-memory offsets
        const int64_t fixed = ((int64_t)R * SP + 4 * SP) * 8 + (((int64_t)R * W * 2 + 15) & ~15LL);
        const int64_t room = (int64_t)smem_limit - static_bytes - fixed;
        if (room < 0) continue;
        int fcap = 0;
        if (ratio) {
            fcap = (int)((room / 8 < (int64_t)R * W) ? room / 8 : (int64_t)R * W);
            if (fcap * 4 < R * W && C < kMaxC) continue;             // want >= 25 % of the cells' features on chip
        }
        out->C = C; out->R = R; out->SP = SP; out->fcap = fcap;
        out->smem_bytes = (size_t)(fixed + (int64_t)fcap * 8);
        return true;
    }
    return false;
}

cudaError_t launch_resident2d(const ResidentArgs &a, bool ratio, size_t smem_bytes, int max_clusters, cudaStream_t st,
                              int64_t *launches) {
    auto kern = ratio ? resident2d_kernel<true> : resident2d_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = a.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int clusters = max_clusters;
    if (clusters <= 0) {
        cfg.gridDim = dim3(a.C, 1, 1);
        e = cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg);
        if (e != cudaSuccess) return e;
        if (clusters < 1) return cudaErrorLaunchOutOfResources;
    }
    if (clusters > a.g.P) clusters = a.g.P;
    cfg.gridDim = dim3((unsigned)(clusters * a.C), 1, 1);
    e = cudaLaunchKernelEx(&cfg, kern, a);
    if (e != cudaSuccess) return e;
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace vpint
