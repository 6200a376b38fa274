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
//
// Rows are dealt to the CTAs of a cluster in BLOCKS of kBlockRows rows, round robin (block b belongs to CTA
// b mod C), not as one contiguous band per CTA: a cloud then spreads over all CTAs of the cluster and the
// busiest CTA -- which sets the pace of every sweep -- carries ~1.2x the average number of unknown cells
// instead of ~2x.  The price is a halo row above and below every block; the neighbours of a CTA are always
// CTA c-1 and c+1 (mod C).
//
// Work unit = a SEGMENT of 4 horizontally adjacent cells (x = 4s .. 4s+3) with at least one unknown cell;
// the CTA keeps a compact row-major list of its segments.  15 WORKER warps own the segments; per sweep they
// wait for the state of the previous sweep (an mbarrier that counts their own arrivals and the bytes of the
// neighbours' boundary rows), compute the new values of their unknown cells into registers (own row / row above
// / row below as 128-bit shared-memory vectors), meet at a block barrier, store the values (own rows, and the
// block-boundary rows into the neighbour CTAs' halo rows with st.async through DSMEM) and arrive.  The 16th
// warp, the SYNC WARP, owns no segments: it adds up the warp partials {sum|new-old|, sum old} of a sweep,
// pushes them to every CTA of the cluster (st.async + mbarrier as well) and takes the stop decision of sweep
// t-1 while the workers compute sweep t (delta = nanmean|new-old| / nanmean(old) <= threshold,
// commit-then-break: VPint/SD_MRP.py:138-147, VPint/WP_MRP.py:351-362) -- if it stops, sweep t is not stored.
// There is no cluster barrier inside the sweep loop; halo rows are double buffered by sweep parity, partials
// triple buffered.
//
// State rows live in shared memory in a split layout (see SmemLayout) that keeps the 128-bit accesses of a
// warp free of bank conflicts; the features of the listed segments live there too, as float pairs when every
// feature is a float32 value.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>

#include "vpint_internal.cuh"

namespace cg = cooperative_groups;

namespace vpint {

namespace {

constexpr int kLaunchThreads = kResidentThreads;
constexpr int kAllWarps = kLaunchThreads / 32;
constexpr int kThreads = kLaunchThreads - 32;   // worker threads: they own the segments; the last warp only does the
constexpr int kWarps = kThreads / 32;           // cross-CTA bookkeeping of a sweep (the "sync warp")
static_assert(kWarps <= 16, "the partial-sum tree below covers 16 warps");
constexpr int kMaxSeg = kResidentMaxSeg;   // segments per thread (4 new values each, in registers)
constexpr int kMaxC = kResidentMaxC;       // portable cluster size

constexpr int kBlockRows = 16;             // rows per round-robin block (clusters of more than one CTA)
constexpr int kMaxBlocks = 8;              // blocks per CTA

// list entry: [15:0] plane offset of the segment's first cell, [19:16] unknown-cell nibble,
// [20] first row of its block (row above is a halo row), [21] last row of its block, [22] touches the grid
// border (nc < 4 somewhere), [23] first row AND a block above exists (row is sent up), [24] last row AND a
// block below exists (row is sent down), [27:25] local block index
// [28] first cell is grid column 0, [29] last cell is grid column W-1, [30] first or last grid row (the three
// border bits drive the branch-free neighbour counts of the fast path)
constexpr unsigned kSegTop = 1u << 20, kSegBot = 1u << 21, kSegEdge = 1u << 22, kSegSendUp = 1u << 23,
                   kSegSendDn = 1u << 24, kSegLeft = 1u << 28, kSegRight = 1u << 29, kSegRowEdge = 1u << 30;

// Cluster barrier with release / acquire at cluster scope.
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// fp64 division that the compiler cannot speculate out of a rarely taken branch (an inlined division is
// ~30 instructions, ~100 with an inf operand such as the default clip_val).
__device__ __forceinline__ double div_here(double a, double b) {
    double r;
    asm volatile("div.rn.f64 %0, %1, %2;" : "=d"(r) : "d"(a), "d"(b));
    return r;
}

// ---- mbarrier / DSMEM primitives (PTX) --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_rank(uint32_t addr, int rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}

// 16-byte store into another CTA's shared memory that signals that CTA's mbarrier with the bytes written:
// data and "it has arrived" travel together, no fence and no cluster barrier needed.
__device__ __forceinline__ void st_async2(uint32_t remote_addr, double a, double b, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];"
                 ::"r"(remote_addr), "d"(a), "d"(b), "r"(remote_bar) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "VPR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra VPR_DONE_%=;\n"
        "bra VPR_WAIT_%=;\n"
        "VPR_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: sleep, do not spin
}

#ifdef VPINT_RES_TRACE
// Debug build only: clock64() stamps of the stages of sweeps 10..17 for thread 0 and the first thread of the sync warp of the
// CTAs of cluster 0 (scripts/res_trace_probe.py).
__device__ long long g_trace[8][2][8][10];
#define VPR_TR(k)                                                                                     \
    do {                                                                                              \
        if (blockIdx.x < 8 && t >= 10 && t < 18 && (c.tid == 0 || c.tid == kThreads))                 \
            g_trace[blockIdx.x][c.tid ? 1 : 0][t - 10][k] = clock64();                                \
    } while (0)
#else
#define VPR_TR(k) do { } while (0)
#endif

// Split-row layout of a state row in shared memory (SP = 4 * nseg + 4 doubles, HALF = SP / 2):
//   [0, HALF)    pair s = cells (4s, 4s+1) at 2s, s = 0 .. nseg-1, then one zero pair
//   [HALF, SP)   one zero pair, then pair s = cells (4s+2, 4s+3) at HALF + 2 + 2s
// A warp whose lanes hold consecutive segments reads each half with 128-bit loads at a 16-byte lane stride --
// no bank conflicts, where the plain row-major layout (32-byte stride) costs two wavefronts per quarter warp.
// With off = row * SP + 2s:  cells 0,1 at off; cells 2,3 at off + HALF + 2; the left neighbour (cell 4s-1) at
// off + HALF + 1; the right neighbour (cell 4s+4) at off + 2; the zero pairs are the grid's left / right border.
struct SmemLayout {
    double *plane;       // [R][SP] owned rows in the split-row layout described below
    double *halo_up;     // [2][LB][SP] row above every owned block, by sweep parity
    double *halo_dn;     // [2][LB][SP] row below
    double *fseg;        // [2][fcap][2] features of the first fcap listed segments: cells 0,1 then cells 2,3 (RATIO);
                         // as float pairs (twice the capacity in the same bytes) when every feature is a float32 value
    unsigned *list;      // [n_seg] segment entries, row-major
};

__device__ __forceinline__ SmemLayout carve(unsigned char *raw, int R, int SP, int LB, int fcap) {
    SmemLayout s;
    s.plane = reinterpret_cast<double *>(raw);
    s.halo_up = s.plane + (size_t)R * SP;
    s.halo_dn = s.halo_up + 2 * LB * SP;
    s.fseg = s.halo_dn + 2 * LB * SP;
    s.list = reinterpret_cast<unsigned *>(s.fseg + (size_t)fcap * 4);
    return s;
}

struct SweepCtx {
    double *plane;
    const double *halo_up, *halo_dn;
    const double *fseg;
    const unsigned *list;
    const double *fp;                 // global feature plane of this problem (segments beyond fcap)
    // neighbour CTAs (shared::cluster addresses): their halo row buffers [2][SP] and their halo barriers [2]
    uint32_t up_halo, dn_halo, up_bar[2], dn_bar[2];
    bool has_up, has_dn;
    int SP, half, fcap, pitch, tid, n, H, W;   // half = SP / 2 (split-row layout)
    int C, crank, BS, LB;             // cluster size / rank, rows per block, blocks per CTA
    int shift_up, shift_dn;           // local block index of block b-1 / b+1 in the neighbour CTA, relative to mine
    int part_lb, part_rows;           // local index / row count of this CTA's partial block (H % BS rows), or -1
    const double *zero4;              // four zeros in shared memory (features of segments that must not count)
    const unsigned short *rowmap;     // compact row -> grid row (only the rows a sweep can touch are on chip)
    double gamma, clip;
    // local row of the last row of block lb
    __device__ __forceinline__ int bot_row(int lb) const { return lb * BS + ((lb == part_lb) ? part_rows : BS) - 1; }
    // grid row of local row ly: blocks of COMPACT rows are dealt round robin
    __device__ __forceinline__ int grid_row(int ly) const {
        const int lb = ly / BS;
        return (int)rowmap[(lb * C + crank) * BS + (ly - lb * BS)];
    }
};

// Per-CTA synchronisation state (static shared memory).
struct SyncShared {
    double part[3][kMaxC][2];          // partial {sum|new-old|, sum old} of every CTA of the cluster (16-byte entries)
    double zero4[4];                   // features of segments that must not count (16-byte aligned)
    double red[2][kWarps][2];          // warp partials by sweep parity (read by the sync warp while the workers go on)
    uint64_t bar_h[2];                 // per sweep parity: this CTA's own stores (one arrival per warp) + neighbours' halo bytes
    uint64_t bar_s[3];                 // per sweep mod 3: the C partial sums (C * 16 bytes)
    int verdict[2];                    // stop decision of the previous sweep by sweep parity, written by the sync warp
    int halo_exp[2];                   // bytes per sweep from the upper / lower neighbour
    int edge_cnt[2];                   // segments on the first / last owned row
};

__device__ __forceinline__ void ld2(const double *p, double &a, double &b) {
    const double2 v = *reinterpret_cast<const double2 *>(p);
    a = v.x; b = v.y;
}

// Stop decision of one sweep from the C partials (the sync warp of every CTA computes the same thing):
// delta = nanmean|new-old| / nanmean(old) with the known-cell constants added (VPint/SD_MRP.py:139).
// Returns 0 continue, 1 stop, 2 non-finite data (leave the problem to the tiled path).
__device__ __forceinline__ int decide_sweep(SyncShared &sh, int slot, unsigned ph_s, int C, const double *KK, double n_cells,
                                            const ResidentArgs &a, int p, int sweep, bool writer) {
    mbar_wait(&sh.bar_s[slot], (ph_s >> slot) & 1u);
    double S0 = 0.0, S1 = 0.0;
    for (int q = 0; q < C; ++q) { S0 += sh.part[slot][q][0]; S1 += sh.part[slot][q][1]; }
    if (!(fabs(S0) <= 1.7976931348623157e308 && fabs(S1) <= 1.7976931348623157e308)) return 2;
    const double sa = S0 + KK[0], so = S1 + KK[1];
    const double na = n_cells - KK[2], no = n_cells - KK[3];
    // delta = (sa / na) / (so / no) costs three dependent fp64 divisions on the critical path of every sweep.
    // (sa * no) against threshold * (so * na) gives the same answer unless delta is within 1e-12 (relative) of
    // the threshold -- the quotient's own rounding error is < 1e-15 -- and only then is it formed exactly.
    const double x = sa * no, y = so * na, ty = a.threshold * y;
    const bool sure = (na > 0.0) && (no > 0.0) && (y > 0.0) && (ty <= 1.7976931348623157e308) &&
                      (fabs(x - ty) > 1e-12 * ty);
    const bool want_delta = (a.delta_out != nullptr);
    if (__all_sync(0xffffffffu, sure && !want_delta)) return (a.auto_terminate && x <= ty) ? 1 : 0;
    const double delta = div_here(div_here(sa, na), div_here(so, no));
    if (want_delta && writer) a.delta_out[(size_t)p * a.max_iter + sweep] = delta;
    return (a.auto_terminate && delta <= a.threshold) ? 1 : 0;
}

// All sweeps of one problem for a CTA whose threads own at most KT segments each (entries tid + k * kThreads).
//
// Sweep t:  (1) wait until the state of sweep t-1 is complete in this CTA: own stores + the neighbours' halo
//               rows (mbarrier bar_h[t & 1]);
//           (2) compute the new values of the thread's segments into registers -- straight-line code, every
//               k is executed (entries past the end re-read segment n-1 and are discarded) so the loads and
//               fp64 chains of the KT segments overlap -- and the warp partial sums;
//           (4) meanwhile the SYNC WARP (warp kWarps, owns no segments) forms the stop decision of sweep t-1,
//               whose partial sums have been travelling since: if it stops, sweep t is simply not stored
//               (commit-then-break without a rollback);
//           (3) block barrier: all reads of the old state are done, the verdict is known;
//           (5) the sync warp adds up the warp partials of sweep t and pushes the CTA's partial to every CTA
//               (st.async -> bar_s[t % 3]) while the workers go on with (6);
//           (6) store the new values: own rows with plain stores, first / last owned row also into the
//               neighbours' halo row of the other parity with st.async (-> their bar_h), then arrive on
//               the own bar_h of the other parity.
// A NaN or inf anywhere poisons the partial sums, which is how data for the tiled path is detected.
// Block barrier shared by the two roles below (the worker warps and the sync warp reach it from different code).
__device__ __forceinline__ void role_barrier() {
    asm volatile("bar.sync 1, %0;" ::"n"(kLaunchThreads) : "memory");
}

// Worker warps: steps (1), (2), (3), (6) of every sweep.
// Features of listed segment i (or four zeros): fp64 pairs, or float pairs when F32.
template <bool F32>
__device__ __forceinline__ void load_fseg(const double *fseg, const double *zero4, int fcap, int i, bool live, double (&f)[4]) {
    if (F32) {
        const float2 *q = reinterpret_cast<const float2 *>(fseg), *z = reinterpret_cast<const float2 *>(zero4);
        const float2 a = *(live ? q + i : z), b = *(live ? q + fcap + i : z);
        f[0] = (double)a.x; f[1] = (double)a.y; f[2] = (double)b.x; f[3] = (double)b.y;
    } else {
        const double *fa = live ? fseg + 2 * i : zero4;
        const double *fb = live ? fseg + 2 * (fcap + i) : zero4 + 2;
        ld2(fa, f[0], f[1]); ld2(fb, f[2], f[3]);
    }
}

template <bool RATIO, int KT, bool FAST, bool F32>
__device__ __forceinline__ void worker_sweeps(const SweepCtx &c, SyncShared &sh, int T, unsigned &ph_h, int &it_out,
                                              bool &bailed_out) {
    const int lane = c.tid & 31, warp = c.tid >> 5;
    const int last = c.n - 1;
    const unsigned halo_bytes = (unsigned)((c.has_up ? sh.halo_exp[0] : 0) + (c.has_dn ? sh.halo_exp[1] : 0));
    int t = 0, verdict = 0;
    for (; t < T; ++t) {
        const int par = t & 1, npar = par ^ 1;
        VPR_TR(0);
        if (t > 0) {
            mbar_wait(&sh.bar_h[par], (ph_h >> par) & 1u);
            ph_h ^= 1u << par;
        }
        VPR_TR(1);
        double nr[KT > 0 ? KT : 1][4];
        double s_abs = 0.0, s_old = 0.0;
        const double *halo_up = c.halo_up + par * c.LB * c.SP, *halo_dn = c.halo_dn + par * c.LB * c.SP;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            const int i0 = c.tid + k * kThreads;
            const bool live = i0 <= last;
            const int i = live ? i0 : last;
            const unsigned e = c.list[i];
            const int off = (int)(e & 0xffffu);
            const unsigned m = live ? ((e >> 16) & 15u) : 0u;
            const double *row = c.plane + off;
            const int lb = (int)((e >> 25) & 7u);
            const double *upp = (e & kSegTop) ? halo_up + (lb * c.SP + off - lb * c.BS * c.SP) : row - c.SP;
            const double *dnp = (e & kSegBot) ? halo_dn + (lb * c.SP + off - c.bot_row(lb) * c.SP) : row + c.SP;
            double c0, c1, c2, c3, u0, u1, u2, u3, d0, d1, d2, d3;
            const int hb = c.half + 2;             // second half of a segment
            ld2(row, c0, c1); ld2(row + hb, c2, c3);
            ld2(upp, u0, u1); ld2(upp + hb, u2, u3);
            ld2(dnp, d0, d1); ld2(dnp + hb, d2, d3);
            const double lf = row[hb - 1], rt = row[2];
            double r[4];
            if (RATIO) {
                r[0] = ((u0 + c1) + d0) + lf;
                r[1] = ((u1 + c2) + d1) + c0;
                r[2] = ((u2 + c3) + d2) + c1;
                r[3] = ((u3 + rt) + d3) + c2;
            } else {
                const double g = c.gamma;
                r[0] = fma(lf, g, fma(d0, g, fma(c1, g, u0 * g)));
                r[1] = fma(c0, g, fma(d1, g, fma(c2, g, u1 * g)));
                r[2] = fma(c1, g, fma(d2, g, fma(c3, g, u2 * g)));
                r[3] = fma(c2, g, fma(d3, g, fma(rt, g, u3 * g)));
            }
            const double cc[4] = {c0, c1, c2, c3};
            if (FAST) {
                // One basic block for all KT segments (no votes, no branches), so that the loads and the fp64
                // chains of different segments overlap.  Neighbour count n in {4, 3, 2} from the border bits;
                // s / n as q = s * fl(1/n) plus one fma residual step, which is the correctly rounded quotient
                // (Markstein; the identity behind div_small() of the tiled kernels; exact no-op for n = 4).
                const bool rowe = (e & kSegRowEdge) != 0u, lft = (e & kSegLeft) != 0u, rgt = (e & kSegRight) != 0u;
                const double third = 1.0 / 3.0;
                const double dm = rowe ? 3.0 : 4.0, im = rowe ? third : 0.25;
                const double dl = lft ? dm - 1.0 : dm, il = lft ? (rowe ? 0.5 : third) : im;
                const double dr = rgt ? dm - 1.0 : dm, ir = rgt ? (rowe ? 0.5 : third) : im;
                const double q0 = r[0] * il, q1 = r[1] * im, q2 = r[2] * im, q3 = r[3] * ir;
                r[0] = fma(fma(-dl, q0, r[0]), il, q0);
                r[1] = fma(fma(-dm, q1, r[1]), im, q1);
                r[2] = fma(fma(-dm, q2, r[2]), im, q2);
                r[3] = fma(fma(-dr, q3, r[3]), ir, q3);
                if (RATIO) {
                    // fseg holds f for the unknown cells of a segment and 0 for the known ones
                    double f[4];
                    load_fseg<F32>(c.fseg, c.zero4, c.fcap, i, live, f);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        s_abs = fma(f[j], fabs(r[j] - cc[j]), s_abs);   // |f r_new - f r_old|
                        s_old = fma(f[j], cc[j], s_old);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool unk = (m >> j) & 1u;
                        s_abs += unk ? fabs(r[j] - cc[j]) : 0.0;
                        s_old += unk ? cc[j] : 0.0;
                    }
                }
            } else {
                const bool edge = (e & kSegEdge) != 0;
                if (__any_sync(0xffffffffu, edge)) {
                    if (edge) {                    // grid border: per-cell neighbour count, true division
                        const int ly = off / c.SP, x0 = 2 * (off - ly * c.SP), gy = c.grid_row(ly);
                        const int ncy = 4 - (gy == 0 ? 1 : 0) - (gy == c.H - 1 ? 1 : 0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int x = x0 + j;
                            const int nc = ncy - (x == 0 ? 1 : 0) - (x == c.W - 1 ? 1 : 0);
                            r[j] = (nc == 4) ? r[j] * 0.25 : div_here(r[j], (double)nc);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) r[j] *= 0.25;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) r[j] *= 0.25;
                }
                if (RATIO) {
                    double f[4];
                    load_fseg<F32>(c.fseg, c.zero4, c.fcap, min(i, c.fcap - 1), true, f);
                    if (i >= c.fcap) {
                        const int ly = off / c.SP;
                        const double *fg = c.fp + (size_t)c.grid_row(ly) * c.pitch + 2 * (off - ly * c.SP);
                        ld2(fg, f[0], f[1]); ld2(fg + 2, f[2], f[3]);
                    }
                    bool clipped = false;
#pragma unroll
                    for (int j = 0; j < 4; ++j) clipped |= ((m >> j) & 1u) && (f[j] * r[j] > c.clip);
                    if (__any_sync(0xffffffffu, clipped)) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (((m >> j) & 1u) && f[j] * r[j] > c.clip) r[j] = div_here(c.clip, f[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double fj = ((m >> j) & 1u) ? f[j] : 0.0;
                        s_abs = fma(fj, fabs(r[j] - cc[j]), s_abs);
                        s_old = fma(fj, cc[j], s_old);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool unk = (m >> j) & 1u;
                        if (r[j] > c.clip) r[j] = c.clip;
                        s_abs += unk ? fabs(r[j] - cc[j]) : 0.0;
                        s_old += unk ? cc[j] : 0.0;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) nr[k][j] = ((m >> j) & 1u) ? r[j] : cc[j];
        }
        VPR_TR(2);
        {
            const double q0 = warp_sum(s_abs), q1 = warp_sum(s_old);
            if (lane == 0) { sh.red[par][warp][0] = q0; sh.red[par][warp][1] = q1; }
        }
        VPR_TR(3);
        role_barrier();                            // (3) every read of the old state is done; red[] / verdict complete
        if (t > 0) {
            verdict = sh.verdict[par];
            if (verdict != 0) break;
        }
        VPR_TR(5);
        const uint32_t npo = (uint32_t)(npar * c.LB * c.SP);   // (6)
        // the rows the neighbour CTAs wait for go out first, the local stores follow
        unsigned ek[KT > 0 ? KT : 1];
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            const int i = c.tid + k * kThreads;
            ek[k] = (i <= last) ? c.list[i] : 0u;
        }
        const int hb = c.half + 2;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            const unsigned e = ek[k];
            if (e & (kSegSendUp | kSegSendDn)) {
                const int off = (int)(e & 0xffffu);
                const int lb = (int)((e >> 25) & 7u);
                if (e & kSegSendUp) {              // first row of my block = lower halo of the block above (CTA c-1)
                    const uint32_t dst = c.up_halo + (npo + (uint32_t)((lb + c.shift_up) * c.SP + off - lb * c.BS * c.SP)) * 8u;
                    st_async2(dst, nr[k][0], nr[k][1], c.up_bar[npar]);
                    st_async2(dst + (uint32_t)hb * 8u, nr[k][2], nr[k][3], c.up_bar[npar]);
                }
                if (e & kSegSendDn) {              // last row of my block = upper halo of the block below (CTA c+1)
                    const uint32_t dst = c.dn_halo + (npo + (uint32_t)((lb + c.shift_dn) * c.SP + off - c.bot_row(lb) * c.SP)) * 8u;
                    st_async2(dst, nr[k][0], nr[k][1], c.dn_bar[npar]);
                    st_async2(dst + (uint32_t)hb * 8u, nr[k][2], nr[k][3], c.dn_bar[npar]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            const int i = c.tid + k * kThreads;
            if (i <= last) {
                const int off = (int)(ek[k] & 0xffffu);
                *reinterpret_cast<double2 *>(c.plane + off) = make_double2(nr[k][0], nr[k][1]);
                *reinterpret_cast<double2 *>(c.plane + off + hb) = make_double2(nr[k][2], nr[k][3]);
            }
        }
        VPR_TR(7);
        __syncwarp();                              // the lanes' stores are ordered before lane 0's (release) arrival
        if (c.tid == 0) mbar_arrive_tx(&sh.bar_h[npar], halo_bytes);
        else if (lane == 0) mbar_arrive(&sh.bar_h[npar]);
        VPR_TR(8);
    }
    if (verdict == 0) {                            // ran to max_iter: close the last sweep
        const int par = T & 1;
        mbar_wait(&sh.bar_h[par], (ph_h >> par) & 1u);
        ph_h ^= 1u << par;
        role_barrier();
        verdict = (sh.verdict[par] == 2) ? 2 : 0;  // stopping at the very last sweep changes nothing
    }
    it_out = t;
    bailed_out = (verdict == 2);
}

// Sync warp: steps (4), (3), (5) of every sweep.
__device__ __forceinline__ void sync_sweeps(int tid, SyncShared &sh, const ResidentArgs &a, int p, int C, int crank, bool fresh,
                                            const double *k0, const double *k1, double n_cells, unsigned &ph_s, int &it_out,
                                            bool &bailed_out) {
    const int lane = tid & 31;
    const int T = a.max_iter;
    const uint32_t my_part = smem_u32(&sh.part[0][0][0]);
    const uint32_t my_bar_s = smem_u32(&sh.bar_s[0]);
    const bool writer = (crank == 0 && lane == 0);
    struct { int tid; } c = {tid};                 // for VPR_TR
    (void)c;
    int t = 0, verdict = 0;
    for (; t < T; ++t) {
        const int par = t & 1, slot = t % 3;
        VPR_TR(0);
        if (t > 0) {                               // (4) runs beside the workers' (2)
            const int v = decide_sweep(sh, (t - 1) % 3, ph_s, C, fresh ? k0 : k1, n_cells, a, p, t - 1, writer);
            if (lane == 0) sh.verdict[par] = v;
        }
        VPR_TR(4);
        role_barrier();
        if (t > 0) {
            ph_s ^= 1u << ((t - 1) % 3);
            fresh = false;
            verdict = sh.verdict[par];
            if (verdict != 0) break;
        }
        VPR_TR(5);
        // (5) fixed-order sum of the warp partials, one 16-byte push per CTA
        double v0 = (lane < kWarps) ? sh.red[par][lane][0] : 0.0, v1 = (lane < kWarps) ? sh.red[par][lane][1] : 0.0;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            v0 += __shfl_down_sync(0xffffffffu, v0, o);
            v1 += __shfl_down_sync(0xffffffffu, v1, o);
        }
        v0 = __shfl_sync(0xffffffffu, v0, 0);
        v1 = __shfl_sync(0xffffffffu, v1, 0);
        if (lane < C)
            st_async2(map_rank(my_part + (uint32_t)(((slot * kMaxC) + crank) * 16), lane), v0, v1,
                      map_rank(my_bar_s + (uint32_t)(slot * 8), lane));
        if (lane == 0) mbar_arrive_tx(&sh.bar_s[slot], (unsigned)(C * 16));
        VPR_TR(6);
    }
    if (verdict == 0) {                            // ran to max_iter: close the last sweep
        const int par = T & 1;
        const int v = decide_sweep(sh, (T - 1) % 3, ph_s, C, fresh ? k0 : k1, n_cells, a, p, T - 1, writer);
        if (lane == 0) sh.verdict[par] = v;
        role_barrier();
        ph_s ^= 1u << ((T - 1) % 3);
        verdict = (sh.verdict[par] == 2) ? 2 : 0;
    }
    it_out = t;
    bailed_out = (verdict == 2);
}

template <bool RATIO>
__global__ void __launch_bounds__(kLaunchThreads, kResidentCtasPerSm) resident2d_kernel(const ResidentArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(16) SyncShared sh;
    __shared__ int wcount[kAllWarps];
    __shared__ int next_problem;
    __shared__ unsigned short rowmap[kResidentMaxRows];    // compact row -> grid row of the current problem
    __shared__ unsigned char rowflag[kResidentMaxRows + 2];

    if (a.bad_flag && *a.bad_flag) return;             // grid-uniform: leave everything to the tiled path
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int C = a.C, R = a.R, SP = a.SP;
    // every feature of the batch is a float32 value (VPint2's inputs are): keep them on chip as floats, twice as many
    const bool f32 = RATIO && a.f32_flag && (*a.f32_flag == 0);
    const int fcap = f32 ? 2 * a.fcap : a.fcap;
    const Geom &g = a.g;
    const int W = g.L, H = g.H;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int BS = a.BS, LB = a.LB;
    const SmemLayout sm = carve(smem_raw, R, SP, LB, a.fcap);
    int *next_remote = cluster.map_shared_rank(&next_problem, 0);
    if (tid < 4) sh.zero4[tid] = 0.0;
    if (tid == 0) {
        mbar_init(&sh.bar_h[0], kWarps); mbar_init(&sh.bar_h[1], kWarps);   // one arrival per warp
        mbar_init(&sh.bar_s[0], 1); mbar_init(&sh.bar_s[1], 1); mbar_init(&sh.bar_s[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned ph_h = 0, ph_s = 0;                       // phase parities of bar_h[2] / bar_s[3], carried across problems

    const size_t plane_elems = (size_t)H * g.pitch;
    const int nwords = (W + 31) >> 5;
    const double n_cells = g.n_cells;

    for (;;) {
        // ---- next problem from the queue (cluster-uniform) ------------------------------------
        if (crank == 0 && tid == 0) next_problem = atomicAdd(a.queue, 1);
        cluster_barrier();
        const int qi = *next_remote;
        cluster_barrier();                             // everyone has read before rank 0 overwrites or exits
        if (qi >= (a.count ? *a.count : g.P)) break;
        const int p = a.order ? a.order[qi] : qi;
        const ProbState st = a.state[p];
        if (st.status != 0) continue;                  // nothing to run (cluster-uniform)

        double *grid = reinterpret_cast<double *>(a.buf[st.cur]) + (size_t)p * plane_elems;
        const double *fp = RATIO ? a.fplane + (size_t)p * plane_elems : nullptr;
        const uint32_t *mrow = a.mask + (size_t)p * H * g.mpitch;
        const double gamma = RATIO ? 1.0 : a.gamma[p];

        // ---- the rows a sweep can touch: rows with an unknown cell and their two neighbours -----------------
        // Only those go on chip, in grid order (the COMPACT rows).  A row with an unknown cell always has both its
        // neighbours next to it in the compact order, so "row above / below" stays an index +-1; rows at the seam of
        // two groups hold no unknown cell and are only ever read.  Hc compact rows are dealt to the CTAs in blocks
        // of BS rows, round robin: local block lb is compact block lb * C + crank.
        for (int y = tid; y < H + 2; y += kLaunchThreads) {
            bool any = false;
            if (y >= 1 && y <= H) {
                const uint32_t *mr = mrow + (size_t)(y - 1) * g.mpitch;
                for (int w = 0; w < nwords && !any; ++w) any = mr[w] != 0u;
            }
            rowflag[y] = any ? 1 : 0;                    // rowflag[y + 1]: grid row y holds an unknown cell
        }
        __syncthreads();
        int Hc = 0;
        {
            // ordered compaction of the needed rows: warps take 32 rows each, in turn
            int base_c = 0;
            for (int y0 = 0; y0 < H; y0 += kLaunchThreads) {
                const int y = y0 + tid;
                const bool need = (y < H) && (rowflag[y] | rowflag[y + 1] | rowflag[y + 2]);
                const unsigned vote = __ballot_sync(0xffffffffu, need);
                if (lane == 0) wcount[warp] = __popc(vote);
                __syncthreads();
                int before = base_c, total = 0;
                for (int w = 0; w < kAllWarps; ++w) {
                    if (w < warp) before += wcount[w];
                    total += wcount[w];
                }
                if (need) rowmap[before + __popc(vote & ((1u << lane) - 1u))] = (unsigned short)y;
                base_c += total;
                __syncthreads();
            }
            Hc = base_c;
        }
        const int NB = (Hc + BS - 1) / BS;
        int n_rows = 0;
        for (int lb = 0; lb < LB; ++lb) {
            const int gb = lb * C + crank;
            if (gb < NB) n_rows += min(BS, Hc - gb * BS);
        }
        auto grid_row = [&](int ly) { const int lb = ly / BS; return (int)rowmap[(lb * C + crank) * BS + (ly - lb * BS)]; };

        // ---- stage the owned rows and both halo rows (r = P / f in RATIO form) ----------------
        // The mask words of the owned rows are parked in the (not yet used) feature area, so that the two passes of
        // the list build below read shared memory instead of paying a global-memory round trip per row.
        unsigned *mask_s = ((size_t)n_rows * nwords * sizeof(unsigned) <= (size_t)a.fcap * 32)
                               ? reinterpret_cast<unsigned *>(sm.fseg) : nullptr;
        for (int row = warp; row < n_rows + 2 * LB; row += kAllWarps) {
            // rows 0 .. n_rows-1: owned rows; then for every block its upper and its lower halo row
            int gy;
            double *dst;
            const bool is_halo = row >= n_rows;
            if (!is_halo) {
                gy = grid_row(row);
                dst = sm.plane + (size_t)row * SP;
            } else {
                const int h = row - n_rows, lb = h >> 1, gb = lb * C + crank;
                const int rows_in = (gb < NB) ? min(BS, Hc - gb * BS) : 0;
                const int cy = (gb < NB) ? ((h & 1) ? gb * BS + rows_in : gb * BS - 1) : -1;     // compact row
                gy = (cy >= 0 && cy < Hc) ? (int)rowmap[cy] : -1;
                dst = ((h & 1) ? sm.halo_dn : sm.halo_up) + (size_t)lb * SP;
            }
            const bool in_grid = (gy >= 0) && (gy < H);
            if (mask_s && !is_halo)
                for (int w = lane; w < nwords; w += 32) mask_s[row * nwords + w] = mrow[(size_t)gy * g.mpitch + w];
            const int half = SP / 2, ncell = 2 * half - 4;          // 4 * segments per row
            if (lane < 2) {                                         // the zero pairs (grid border left / right)
                dst[ncell / 2 + lane] = 0.0; dst[half + lane] = 0.0;
                if (is_halo) { dst[LB * SP + ncell / 2 + lane] = 0.0; dst[LB * SP + half + lane] = 0.0; }
            }
            // coalesced reads, 8 independent loads (and divisions) in flight per lane
            constexpr int kU = 8;
            const double *grow = grid + (size_t)max(gy, 0) * g.pitch;
            const double *frow = RATIO ? fp + (size_t)max(gy, 0) * g.pitch : nullptr;
            for (int x0 = 0; x0 < ncell; x0 += 32 * kU) {
                double v[kU], f[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int x = x0 + u * 32 + lane;
                    const bool ok = in_grid && x < W;
                    v[u] = ok ? grow[x] : 0.0;
                    f[u] = (RATIO && ok) ? frow[x] : 1.0;
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int x = x0 + u * 32 + lane;
                    if (RATIO) v[u] = v[u] / f[u];
                    if (x < ncell) {
                        const int sg = x >> 2, j = x & 3;
                        const int slot = (j < 2) ? 2 * sg + j : half + 2 * sg + j;
                        dst[slot] = v[u];
                        if (is_halo) dst[LB * SP + slot] = v[u];    // second parity copy of a halo row
                    }
                }
            }
        }

        __syncthreads();                               // mask_s complete (rows were staged by other warps)

        // ---- compact row-major list of the segments that hold an unknown cell ------------------
        // A mask word covers 8 segments (nibbles); warps take contiguous row chunks, so the list order
        // (and with it every partial sum) is fixed.
        const int rows_per_warp = (n_rows + kAllWarps - 1) / kAllWarps;
        const int wr0 = min(warp * rows_per_warp, n_rows), wr1 = min(wr0 + rows_per_warp, n_rows);
        auto nibbles = [](unsigned w) {                // bit 4q set <=> nibble q of w is nonzero
            unsigned t = w | (w >> 1);
            t |= t >> 2;
            return t & 0x11111111u;
        };
        int cnt = 0;
        for (int ly = wr0; ly < wr1; ++ly)
            for (int w = lane; w < nwords; w += 32)
                cnt += __popc(nibbles(mask_s ? mask_s[ly * nwords + w] : mrow[(size_t)grid_row(ly) * g.mpitch + w]));
        cnt = warp_sum(cnt);
        if (lane == 0) wcount[warp] = cnt;
        __syncthreads();
        int base = 0, n = 0;
        for (int w = 0; w < kAllWarps; ++w) {
            if (w < warp) base += wcount[w];
            n += wcount[w];
        }
        for (int ly = wr0; ly < wr1; ++ly) {
            const int gy = grid_row(ly);
            const int lb = ly / BS, gb = lb * C + crank;
            const bool top = (ly == lb * BS), bot = (ly == lb * BS + min(BS, Hc - gb * BS) - 1);
            const unsigned row_flags = (top ? kSegTop : 0u) | (bot ? kSegBot : 0u) | ((top && gb > 0) ? kSegSendUp : 0u) |
                                       ((bot && gb < NB - 1) ? kSegSendDn : 0u) | ((unsigned)lb << 25);
            const bool row_edge = (gy == 0) || (gy == H - 1);
            for (int w0 = 0; w0 < nwords; w0 += 32) {
                const int w = w0 + lane;
                const unsigned word = (w < nwords) ? (mask_s ? mask_s[ly * nwords + w] : mrow[(size_t)gy * g.mpitch + w]) : 0u;
                unsigned nz = nibbles(word);
                const int pc = __popc(nz);
                int incl = pc;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                int pos = base + incl - pc;
                base += __shfl_sync(0xffffffffu, incl, 31);
                while (nz) {
                    const int b = __ffs(nz) - 1;       // bit 4q
                    nz &= nz - 1;
                    const int x0 = (w << 5) + b;       // first cell of the segment
                    const unsigned nib = (word >> b) & 15u;
                    const bool left = (x0 == 0), right = (x0 + 3 >= W - 1);
                    const bool edge = row_edge || left || right;
                    sm.list[pos++] = (unsigned)(ly * SP + (x0 >> 1)) | (nib << 16) | row_flags | (edge ? kSegEdge : 0u) |
                                     (left ? kSegLeft : 0u) | (right ? kSegRight : 0u) | (row_edge ? kSegRowEdge : 0u);
                }
            }
        }
        __syncthreads();
        if (RATIO) {                                   // features of the listed segments (pad cells read 1.0)
            for (int i = tid; i < min(n, fcap); i += kLaunchThreads) {
                const unsigned e = sm.list[i];
                const int off = (int)(e & 0xffffu);
                const int ly = off / SP, x0 = 2 * (off - ly * SP);
                const double *fg = fp + (size_t)grid_row(ly) * g.pitch + x0;
                double f0, f1, f2, f3;
                ld2(fg, f0, f1); ld2(fg + 2, f2, f3);
                // 0 for the known cells: they do not count
                f0 = (e & (1u << 16)) ? f0 : 0.0; f1 = (e & (1u << 17)) ? f1 : 0.0;
                f2 = (e & (1u << 18)) ? f2 : 0.0; f3 = (e & (1u << 19)) ? f3 : 0.0;
                if (f32) {
                    float2 *q = reinterpret_cast<float2 *>(sm.fseg);
                    q[i] = make_float2((float)f0, (float)f1);
                    q[fcap + i] = make_float2((float)f2, (float)f3);
                } else {
                    *reinterpret_cast<double2 *>(sm.fseg + 2 * i) = make_double2(f0, f1);
                    *reinterpret_cast<double2 *>(sm.fseg + 2 * (fcap + i)) = make_double2(f2, f3);
                }
            }
        }
        // Segments per thread of THIS warp (entries tid + k * kThreads): the warps that hold no entry of the last
        // round run a shorter variant, so a list slightly longer than a multiple of kThreads costs a few warps one
        // more segment instead of all of them.  Every variant meets the same barriers.
        const int w_first = warp * 32;                     // lowest entry index of the warp's first round
        const int Kmax = (n > w_first) ? (n - 1 - w_first) / kThreads + 1 : 0;   // warp-uniform
        int kt = kMaxSeg;                                  // smallest unrolled variant >= Kmax
        {
            const int kts[8] = {0, 1, 2, 3, 4, 5, 6, 8};
#pragma unroll
            for (int q = 7; q >= 0; --q) if (Kmax <= kts[q]) kt = kts[q];
        }
        // segments on the first / last owned row: what the neighbours must expect from this CTA per sweep
        if (tid < 2) sh.edge_cnt[tid] = 0;
        __syncthreads();
        {
            int n_top = 0, n_bot = 0;
            for (int i = tid; i < n; i += kLaunchThreads) {
                const unsigned e = sm.list[i];
                n_top += (e & kSegSendUp) ? 1 : 0;
                n_bot += (e & kSegSendDn) ? 1 : 0;
            }
            n_top = warp_sum(n_top);
            n_bot = warp_sum(n_bot);
            if (lane == 0) { atomicAdd(&sh.edge_cnt[0], n_top); atomicAdd(&sh.edge_cnt[1], n_bot); }
        }
        __syncthreads();
        if (tid == 0) {
            if (C > 1) {
                *cluster.map_shared_rank(&sh.halo_exp[1], (crank + C - 1) % C) = 32 * sh.edge_cnt[0];   // into its lower halos
                *cluster.map_shared_rank(&sh.halo_exp[0], (crank + 1) % C) = 32 * sh.edge_cnt[1];       // into its upper halos
            }
        }
        SweepCtx ctx;
        ctx.plane = sm.plane; ctx.halo_up = sm.halo_up; ctx.halo_dn = sm.halo_dn; ctx.fseg = sm.fseg; ctx.list = sm.list;
        ctx.fp = fp;
        ctx.has_up = ctx.has_dn = (C > 1);
        const int cu = (crank + C - 1) % C, cd = (crank + 1) % C;
        ctx.up_halo = ctx.has_up ? map_rank(smem_u32(sm.halo_dn), cu) : 0u;
        ctx.dn_halo = ctx.has_dn ? map_rank(smem_u32(sm.halo_up), cd) : 0u;
        for (int q = 0; q < 2; ++q) {
            ctx.up_bar[q] = ctx.has_up ? map_rank(smem_u32(&sh.bar_h[q]), cu) : 0u;
            ctx.dn_bar[q] = ctx.has_dn ? map_rank(smem_u32(&sh.bar_h[q]), cd) : 0u;
        }
        ctx.C = C; ctx.crank = crank; ctx.BS = BS; ctx.LB = LB;
        ctx.shift_up = (crank == 0) ? -1 : 0;       // block b-1 lives in CTA c-1 under the same local index, or one less after the wrap
        ctx.shift_dn = (crank == C - 1) ? 1 : 0;
        ctx.part_lb = -1; ctx.part_rows = BS;      // the partial block, if any, is the grid's last one
        if ((Hc % BS) != 0 && ((NB - 1) % C) == crank) { ctx.part_lb = (NB - 1) / C; ctx.part_rows = Hc % BS; }
        ctx.zero4 = sh.zero4;
        ctx.rowmap = rowmap;
        ctx.SP = SP; ctx.half = SP / 2; ctx.fcap = fcap; ctx.pitch = g.pitch; ctx.tid = tid;
        ctx.n = n; ctx.H = H; ctx.W = W;
        ctx.gamma = gamma; ctx.clip = a.clip;
        const double *K0 = a.known0 + (size_t)p * kSumSlots, *K1 = a.known + (size_t)p * kSumSlots;
        const double k0[kSumSlots] = {K0[0], K0[1], K0[2], K0[3]};
        const double k1[kSumSlots] = {K1[0], K1[1], K1[2], K1[3]};
        cluster_barrier();                             // lists, planes and expected byte counts of every CTA are in place

        // ---- sweeps -----------------------------------------------------------------------------
        int it = 0;
        bool bailed = false;
        const bool fresh = (st.fresh != 0);
        // The fast sweep needs: no clip_val, every feature of the listed segments on chip, the grid's last column in
        // the last cell of a segment, and at most one missing neighbour per direction pair (H >= 3, W >= 4).
        // Each CTA decides for itself -- the protocol between the CTAs does not depend on it.
        const bool fast = !(a.clip <= 1.7976931348623157e308) && (!RATIO || n <= fcap) && (W & 3) == 0 && H >= 3;
#define VPR_RUN(KT_)                                                                                   \
    do {                                                                                               \
        if (RATIO && f32) {                                                                            \
            if (fast) worker_sweeps<RATIO, KT_, true, RATIO>(ctx, sh, a.max_iter, ph_h, it, bailed);   \
            else worker_sweeps<RATIO, KT_, false, RATIO>(ctx, sh, a.max_iter, ph_h, it, bailed);       \
        } else {                                                                                       \
            if (fast) worker_sweeps<RATIO, KT_, true, false>(ctx, sh, a.max_iter, ph_h, it, bailed);   \
            else worker_sweeps<RATIO, KT_, false, false>(ctx, sh, a.max_iter, ph_h, it, bailed);       \
        }                                                                                              \
    } while (0)
        if (warp == kWarps) {
            sync_sweeps(tid, sh, a, p, C, crank, fresh, k0, k1, n_cells, ph_s, it, bailed);
        } else {
#ifdef VPR_ONLY_KT                                  // build-time triage: one variant only, for per-variant ptxas -v / SASS
            VPR_RUN(VPR_ONLY_KT);
#else
            switch (kt) {
                case 0: VPR_RUN(0); break;
                case 1: VPR_RUN(1); break;
                case 2: VPR_RUN(2); break;
                case 3: VPR_RUN(3); break;
                case 4: VPR_RUN(4); break;
                case 5: VPR_RUN(5); break;
                case 6: VPR_RUN(6); break;
                case 8: VPR_RUN(8); break;
                default: VPR_RUN(kMaxSeg); break;
            }
#endif
        }
#undef VPR_RUN
        __syncthreads();

        // ---- write the unknown cells back (P = f * r), close the problem ------------------------
        for (int i = tid; i < n && !bailed; i += kLaunchThreads) {
            const unsigned e = sm.list[i];
            const int off = (int)(e & 0xffffu);
            const int ly = off / SP, x0 = 2 * (off - ly * SP);
            const size_t go = (size_t)grid_row(ly) * g.pitch + x0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if ((e >> (16 + j)) & 1u) {
                    double v = sm.plane[off + ((j < 2) ? j : SP / 2 + j)];
                    if (RATIO) {
                        const int fi = 2 * (((j < 2) ? 0 : fcap) + i) + (j & 1);
                        const double fj = (i >= fcap) ? fp[go + j]
                                                      : (f32 ? (double)reinterpret_cast<const float *>(sm.fseg)[fi] : sm.fseg[fi]);
                        v = fj * v;
                    }
                    grid[go + j] = v;
                }
            }
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

// order[c * P + rank] = p for the problems of class c, ranked by descending unknown-cell count (ties by index): every
// thread counts the problems of its class that come before its own.  P <= kResidentOrderMax keeps the quadratic scan
// in the microseconds.  cls_of == null: one class (0).
__global__ void __launch_bounds__(256) resident_order_kernel(int P, const ProbState *__restrict__ state,
                                                             const int *__restrict__ cls_of, int *__restrict__ order,
                                                             int *__restrict__ cls_count) {
    __shared__ int keys[1024];
    __shared__ int kcls[1024];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int mine = (p < P) ? state[p].n_unknown : 0;
    const int mycls = (p < P && cls_of) ? cls_of[p] : 0;
    int rank = 0;
    for (int q0 = 0; q0 < P; q0 += 1024) {
        __syncthreads();
        for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
            keys[i] = (q0 + i < P) ? state[q0 + i].n_unknown : -1;
            kcls[i] = (q0 + i < P && cls_of) ? cls_of[q0 + i] : 0;
        }
        __syncthreads();
        const int m = min(1024, P - q0);
        for (int i = 0; i < m; ++i) {
            const int k = keys[i];
            rank += (kcls[i] == mycls && (k > mine || (k == mine && q0 + i < p))) ? 1 : 0;
        }
    }
    if (p < P) {
        order[(size_t)mycls * P + rank] = p;
        if (cls_count) atomicAdd(&cls_count[mycls], 1);
    }
}

// Class of every problem = the smallest cluster whose CTAs hold its COMPACT rows (rows with an unknown cell and their
// neighbours; the kernel derives the same set from the bit mask).  rowpart[.][7] = unknown cells of a row (load).
struct ResidentCaps {
    int cap[kResidentMaxC + 1];        // compact rows a cluster of C CTAs holds; 0 = no plan
};

__global__ void __launch_bounds__(256) resident_class_kernel(Geom g, const double *__restrict__ rowpart, ResidentCaps caps,
                                                             int c_max, int *__restrict__ cls_of) {
    __shared__ unsigned char flag[kResidentMaxRows + 2];
    __shared__ int part[8];
    const int p = blockIdx.x, H = g.H;
    for (int y = threadIdx.x; y < H + 2; y += blockDim.x)
        flag[y] = (y >= 1 && y <= H && rowpart[((size_t)p * H + (y - 1)) * 8 + 7] > 0.0) ? 1 : 0;
    __syncthreads();
    int cnt = 0;
    for (int y = threadIdx.x; y < H; y += blockDim.x) cnt += (flag[y] | flag[y + 1] | flag[y + 2]) ? 1 : 0;
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int hc = 0;
        for (int w = 0; w < 8; ++w) hc += part[w];
        int c = c_max;
        for (int q = c_max; q >= 1; --q)
            if (caps.cap[q] >= hc) c = q;
        cls_of[p] = c;
    }
}

cudaError_t launch_resident_order(const Geom &g, const ProbState *state, const double *rowpart, const int *caps, int c_max,
                                  int *cls_of, int *order, int *cls_count, cudaStream_t st, int64_t *launches) {
    ResidentCaps rc;
    for (int c = 0; c <= kResidentMaxC; ++c) rc.cap[c] = caps[c];
    resident_class_kernel<<<g.P, 256, 0, st>>>(g, rowpart, rc, c_max, cls_of);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (launches) ++*launches;
    resident_order_kernel<<<(g.P + 255) / 256, 256, 0, st>>>(g.P, state, cls_of, order, cls_count);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (launches) ++*launches;
    return cudaSuccess;
}

// Plan for a cluster of C CTAs: as many (compact) rows per CTA as shared memory, the 16-bit plane offsets and the
// register budget allow.  C == 1 holds its rows as one block (no halo exchange), larger clusters deal blocks of
// kBlockRows rows round robin.  Returns false if not even one block fits.
bool resident_plan(const Geom &g, bool ratio, int smem_limit, int C, ResidentPlan *out) {
    if (g.ndim != 2 || g.H < 1 || g.L < 1 || (g.H < 2 && g.L < 2)) return false;   // 1x1: neighbour count 0
    if (g.H > kResidentMaxRows || C < 1 || C > kMaxC) return false;
    const int W = g.L;
    const int segs_per_row = (W + 3) / 4;
    const int SP = 4 * segs_per_row + 4;        // split-row layout: two halves of (segs_per_row + 1) pairs each
    if (g.pitch < 4 * segs_per_row) return false;
    const int static_bytes = 3584;              // SyncShared, wcount, next_problem, rowmap, rowflag (+ slack)
    auto fits = [&](int BS, int LB, int *fcap_out, size_t *bytes_out) {
        const int R = LB * BS;
        if ((int64_t)R * segs_per_row > (int64_t)kThreads * kMaxSeg) return false;  // register budget for the new values
        if ((int64_t)R * SP > 65535) return false;                                  // 16-bit shared-memory offsets
        const int64_t fixed = ((int64_t)R * SP + 4 * (int64_t)LB * SP) * 8 + (int64_t)R * segs_per_row * 4;
        const int64_t room = (int64_t)smem_limit - static_bytes - fixed;
        if (room < 0) return false;
        int fcap = 0;
        if (ratio) {
            const int64_t all = (int64_t)R * segs_per_row;
            fcap = (int)((room / 32 < all) ? room / 32 : all);
            if (fcap < 1) return false;
            if (fcap * 4 < all) return false;                                       // want >= 25 % of the features on chip
        }
        *fcap_out = fcap;
        *bytes_out = (size_t)(fixed + (int64_t)fcap * 32);
        return true;
    };
    int BS = 0, LB = 0, fcap = 0;
    size_t bytes = 0;
    if (C == 1) {
        for (int R = g.H; R >= 1; --R)
            if (fits(R, 1, &fcap, &bytes)) { BS = R; LB = 1; break; }
    } else {
        for (int lb = kMaxBlocks; lb >= 1; --lb)
            if (fits(kBlockRows, lb, &fcap, &bytes)) { BS = kBlockRows; LB = lb; break; }
        // no more blocks than the whole grid needs
        const int need = ((g.H + kBlockRows - 1) / kBlockRows + C - 1) / C;
        if (LB > need && need >= 1 && fits(kBlockRows, need, &fcap, &bytes)) LB = need;
    }
    if (LB < 1) return false;
    out->C = C; out->R = LB * BS; out->SP = SP; out->fcap = fcap; out->BS = BS; out->LB = LB;
    out->smem_bytes = bytes;
    return true;
}

cudaError_t launch_resident2d(const ResidentArgs &a_in, bool ratio, size_t smem_bytes, int max_clusters, cudaStream_t st,
                              int64_t *launches) {
    const ResidentArgs &a = a_in;
    auto kern = ratio ? resident2d_kernel<true> : resident2d_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kLaunchThreads, 1, 1);
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
    if (getenv("VPINT_DEBUG"))
        fprintf(stderr, "[vpint] resident2d: C=%d R=%d (blocks of %d rows x %d) SP=%d fcap=%d smem=%zu max_clusters=%d P=%d\n",
                a.C, a.R, a.BS, a.LB, a.SP, a.fcap, smem_bytes, clusters, a.g.P);
    if (clusters > a.g.P) clusters = a.g.P;
    cfg.gridDim = dim3((unsigned)(clusters * a.C), 1, 1);
    e = cudaLaunchKernelEx(&cfg, kern, a);
    if (e != cudaSuccess) return e;
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace vpint

#ifdef VPINT_RES_TRACE
extern "C" int vpint_debug_trace(long long *host_out, int n_bytes) {
    return (int)cudaMemcpyFromSymbol(host_out, vpint::g_trace, (size_t)n_bytes);
}
#endif
