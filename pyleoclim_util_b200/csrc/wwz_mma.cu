// Shared-time-axis transform as a contraction on the FP64 tensor pipe (sm_100a, DMMA.8x8x4).
//
// All surrogates of one target share its time axis (core/surrogateseries.py:170-174), the weights, Neff and
// the basis sums of a cell do not depend on y (utils/wavelet.py:988-1019) and a0, a1, a2 are linear in y
// (:1021-1023).  The S-series batch is therefore
//
//     [ w ; w*sin(omega t) ; w*cos(omega t) ] (3*cells x points)   times   Y (points x series)
//
// with the left operand generated on the fly: per (cell, point) one weight (exp2 on the FP64 pipe), evaluated
// exactly once and used for every series.
//
// Work decomposition.  A warp owns 8 cells (8 consecutive frequency rows at one tau) and NS series.  Per step of
// 4 points, lane (r = lane/4, p = lane%4) evaluates the weight of (cell r, point p) -- that IS the A fragment
// of mma.sync.m8n8k4.f64 (row = lane/4, k = lane%4), no shuffles -- and issues 3*NS/8 DMMAs against the B
// fragments Y[p][8t + lane/4] read from the staged series tile.  W, Q, Ws, Wc are accumulated per lane and
// reduced over the quad at the end.  18 FP64 instructions for the A side, 3*NS/8 DMMAs (16 pipe cycles each); an
// FP64 FMA-pipe instruction issued between DMMAs costs ~4.5 pipe cycles (tools/dmma_mix_micro.cu), which caps the
// tensor pipe at ~82% with the weights evaluated in the loop -- hence the stored-weights mode below.
//
// A CTA = 8 compute warps = 32 frequency rows x 2 taus, + one TMA producer warp.  The tile is compact in tau and
// frequency, so its Gaussian support is one index range of the sorted time axis: the producer finds it (warp-
// cooperative search) and only the chunks that intersect it are staged; points below 2^-100 of the largest weight
// of every cell of the tile are skipped exactly as in wwz_rec.cu, and inside the tile's window every warp skips the
// chunks outside its own (warp_window).  Per chunk of P points: one bulk copy for (t, .), one per frequency row of
// {sin, cos} records, one per point for the [P][NS] series tile (rows skewed by 32 B: conflict-free B fragments),
// one for the stored A fragments; 4-stage ring, full/empty mbarriers.
//
// Stored weights (PRE).  The weights do not depend on the series block: wwz_tile_weights_kernel evaluates the A
// fragments of every tile once, in fragment order, with the y-independent sums of every cell, and the contraction
// kernel streams them through the ring -- 2 DMULs per 24 DMMAs remain, the tensor pipe runs at 90%.
//
// Grid: x = cell tiles (low frequencies, i.e. the long windows, first), y = series blocks: CTAs that run
// together share the series block (streamed from HBM once) and differ in the basis rows (L2 resident).
#include <math_constants.h>

#include <algorithm>
#include <cstring>
#include <cstdlib>

#include "wwz_device.cuh"
#include "wwz_internal.h"

namespace wwz {

constexpr int kMmaStages = 4;
// 4 x 2: 32 frequency rows x 2 taus.  Warp w = wf + 4*wt sits on SMSP wf, so the four row groups of one tau cover the
// four SMSPs: a tau whose window is disjoint from its neighbour's (high frequencies) still keeps the whole SM busy
// (2 x 4 left two SMSPs idle there: config 3 122.6 -> 120.2 ms), and nt = 50 needs no padding.  The price is a coarser
// unit for frequency-sharded runs (distributed.balanced_freq_groups deals 32-row groups); the default multi-GPU
// significance test shards surrogates and always sees the whole grid.
constexpr int kMmaWF = 4;                       // warps along the frequency axis
constexpr int kMmaWT = 2;                       // warps along the tau axis
constexpr int kMmaRows = 8 * kMmaWF;            // frequency rows of a CTA tile
constexpr int kMmaWarps = kMmaWF * kMmaWT;      // compute warps
constexpr int kMmaThreads = (kMmaWarps + 1) * 32;
static_assert(kMmaWarps == 8 && kMmaRows <= 32, "tile geometry");

struct MmaArgs {
    const double2 *ty;      // [n_pad] (t, .)
    const double2 *basis;   // [nf][n_pad] records {s, c}
    const double *Yt;       // [S_pad/NS][n_pad][NS]
    const double *ts;
    const double *tau;
    const double *kappa;
    const double *dcut;
    const double *taugap;   // distance from each tau to the nearest sample
    const int *unsorted;
    const double2 *rot;     // [ncell] (sin, cos)(omega*tau)
    double *neffs;          // [ncell]
    double *wq;             // [2][ncell] W, Q
    double *amp, *phase, *a0, *a1, *a2;   // [ncell][S_pad]
    // stored-weights mode (PRE): A fragments written once by wwz_tile_weights_kernel,
    // wfrag[tile][chunk][warp][step][lane], and the y-independent sums of every cell
    double *wfrag;
    double *cellsums;       // [4][ncell] W, Q, Ws, Wc
    long long n, n_pad, ncell, S_pad;
    long long wf_tile_stride;   // doubles per tile in wfrag = chunks of the whole axis * P * 64
    int nt, nf, P, tiles_t;
    int row_stride2, stage_stride2, y_off2, w_off2;   // double2 units
    double thr;
    AmpScatterDev sc;       // destination of the amplitudes (cells_per == 0: the local buffer `amp`)
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Point window of a cell tile, by one warp: chunks [win[0], win[0] + win[1]) hold every point whose weight reaches
// the cut-off for some cell of the tile ([min tau - dcut, max tau + dcut] with the widest dcut of its rows).
__device__ __forceinline__ void tile_window(const MmaArgs &a, int k_base, int rows, int j_base, int lane, int *win)
{
    double dc = -1.0, tlo = CUDART_INF, thi = -CUDART_INF, gp = 0.0;
    if (lane < rows) dc = a.dcut[k_base + lane];            // -1: NaN row (no constraint of its own)
    if (lane < kMmaWT && j_base + lane < a.nt) {
        tlo = thi = a.tau[j_base + lane];
        gp = a.taugap[j_base + lane];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        gp = fmax(gp, __shfl_xor_sync(0xffffffffu, gp, o));
        dc = fmax(dc, __shfl_xor_sync(0xffffffffu, dc, o));
        tlo = fmin(tlo, __shfl_xor_sync(0xffffffffu, tlo, o));
        thi = fmax(thi, __shfl_xor_sync(0xffffffffu, thi, o));
    }
    long long i0 = 0, i1 = a.n;
    if (dc < 0.0) {
        i1 = 0;                                             // every row is NaN: nothing to add, Neff = 0/0
    } else if (!*a.unsorted && dc < CUDART_INF) {
        const double de = sqrt(fma(gp, gp, dc * dc));       // cut relative to the cells' own largest weight
        i0 = warp_bound<false>(a.ts, 0, a.n, tlo - de, lane);
        i1 = warp_bound<true>(a.ts, i0, a.n, thi + de, lane);
    }
    if (lane == 0) {
        const int c_lo = (int)(i0 / a.P);
        win[0] = c_lo;
        win[1] = i1 > i0 ? (int)((i1 + a.P - 1) / a.P) - c_lo : 0;
    }
}

// The chunks of the tile's window that matter to ONE warp (8 frequency rows at one tau): its own relative window,
// [tau - de, tau + de] with de from the widest dcut of its rows and the gap of its tau.  A warp skips the arithmetic
// of the staged chunks outside it (high-frequency tiles: the 2-tau, 32-row window of the CTA is several times wider
// than what one warp needs).  Returns chunk indices [lo, hi) on the whole axis; all lanes get the same values.
__device__ __forceinline__ void warp_window(const MmaArgs &a, int kc, int jc, bool row_ok, int lane, int &lo, int &hi)
{
    double dc = row_ok ? a.dcut[kc] : -1.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dc = fmax(dc, __shfl_xor_sync(0xffffffffu, dc, o));
    lo = 0;
    hi = (int)((a.n + a.P - 1) / a.P);
    if (dc < 0.0) {
        hi = 0;                                   // NaN rows only
    } else if (!*a.unsorted && dc < CUDART_INF) {
        const double tj = a.tau[jc], gp = a.taugap[jc];
        const double de = sqrt(fma(gp, gp, dc * dc));
        const long long i0 = warp_bound<false>(a.ts, 0, a.n, tj - de, lane);
        const long long i1 = warp_bound<true>(a.ts, i0, a.n, tj + de, lane);
        lo = (int)(i0 / a.P);
        hi = i1 > i0 ? (int)((i1 + a.P - 1) / a.P) : lo;
    }
}

// Stored-weights mode, pass 1: the A fragments of every tile, once for all series blocks.  Same tile geometry,
// window and lane mapping as the contraction kernel (8 warps, no staging: the operands are read straight from
// global memory, L2 resident), plus the y-independent sums W, Q, Ws, Wc, Neff of every cell.
__global__ void __launch_bounds__(kMmaWarps * 32) wwz_tile_weights_kernel(const MmaArgs a)
{
    __shared__ double tbl[16];
    __shared__ int win[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kt = (int)blockIdx.x / a.tiles_t, jt = (int)blockIdx.x - kt * a.tiles_t;
    const int k_base = kt * kMmaRows, j_base = jt * kMmaWT;
    const int rows = min(kMmaRows, a.nf - k_base);
    const int P = a.P;
    load_exp2_table(tbl, tid);
    if (warp == 0) tile_window(a, k_base, rows, j_base, lane, win);
    __syncthreads();
    const int c_lo = win[0], nchunks = win[1];
    const int wf = warp % kMmaWF, wt = warp / kMmaWF;
    const int r = lane >> 2, p = lane & 3;
    const int k = k_base + wf * 8 + r, j = j_base + wt;
    const bool ok = k < a.nf && j < a.nt;
    const int kc = min(k, a.nf - 1), jc = min(j, a.nt - 1);
    const double kap = a.kappa[kc];
    const double tj = a.tau[jc];
    const double2 *const brow = a.basis + (size_t)kc * (size_t)a.n_pad;
    double *const wout = a.wfrag + (size_t)blockIdx.x * (size_t)a.wf_tile_stride + (size_t)warp * (size_t)(P / 4) * 32 + lane;
    double W = 0.0, Q = 0.0, Ws = 0.0, Wc = 0.0;
    int cw_lo, cw_hi;
    warp_window(a, kc, jc, k < a.nf, lane, cw_lo, cw_hi);
    for (int i = 0; i < nchunks; ++i) {
        if (c_lo + i < cw_lo || c_lo + i >= cw_hi) continue;     // outside this warp's own window: never read either
        const long long base = (long long)(c_lo + i) * P;
        double *const wo = wout + (size_t)(c_lo + i) * (size_t)P * 64;
#pragma unroll 4
        for (int q = p; q < P; q += 4) {
            const long long gi = base + q;
            const double2 sc = brow[gi];
            double w = gauss_weight_estrin(kap * (a.ty[gi].x - tj), tbl);
            if (gi >= a.n) w = 0.0;
            W += w;
            Q = fma(w, w, Q);
            Ws = fma(w, sc.x, Ws);
            Wc = fma(w, sc.y, Wc);
            wo[(q >> 2) * 32] = w;
        }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
        W += __shfl_xor_sync(0xffffffffu, W, o);
        Q += __shfl_xor_sync(0xffffffffu, Q, o);
        Ws += __shfl_xor_sync(0xffffffffu, Ws, o);
        Wc += __shfl_xor_sync(0xffffffffu, Wc, o);
    }
    if (ok && p == 0) {
        const size_t cell = (size_t)j * (size_t)a.nf + (size_t)k;
        const size_t nc = (size_t)a.ncell;
        a.cellsums[cell] = W;
        a.cellsums[nc + cell] = Q;
        a.cellsums[2 * nc + cell] = Ws;
        a.cellsums[3 * nc + cell] = Wc;
        a.neffs[cell] = W * W / Q;
        a.wq[cell] = W;
        a.wq[nc + cell] = Q;
    }
}

// PRE: the A fragments come from wfrag (stored-weights mode) instead of being evaluated in the loop.
template <int NS, bool PRE>
__global__ void __launch_bounds__(kMmaThreads, 1) wwz_shared_mma_kernel(const MmaArgs a)
{
    constexpr int NT = NS / 8;   // n-tiles of 8 series
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *const stage_base = reinterpret_cast<double2 *>(smem_raw);
    unsigned char *const tail = smem_raw + (size_t)kMmaStages * (size_t)a.stage_stride2 * sizeof(double2);
    uint64_t *const full = reinterpret_cast<uint64_t *>(tail);
    uint64_t *const empty = full + kMmaStages;
    double *const tbl = reinterpret_cast<double *>(empty + kMmaStages);   // [16] 2^(j/16)
    int *const win = reinterpret_cast<int *>(tbl + 16);                   // first chunk, number of chunks

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kt = (int)blockIdx.x / a.tiles_t, jt = (int)blockIdx.x - kt * a.tiles_t;
    const int k_base = kt * kMmaRows, j_base = jt * kMmaWT;
    const int rows = min(kMmaRows, a.nf - k_base);
    const int P = a.P;
    const long long sb = blockIdx.y;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kMmaStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kMmaWarps);
        }
        mbar_fence_init();
    }
    load_exp2_table(tbl, tid);
    if (warp == kMmaWarps) tile_window(a, k_base, rows, j_base, lane, win);
    __syncthreads();
    const int c_lo = win[0], nchunks = win[1];

    if (warp == kMmaWarps) {
        const uint32_t bytes = (uint32_t)P * 16u + (uint32_t)rows * (uint32_t)P * 16u + (uint32_t)P * NS * 8u +
                               (PRE ? (uint32_t)P * 512u : 0u);
        int s = 0, use = 0;
        for (int i = 0; i < nchunks; ++i) {
            if (use > 0) mbar_wait(&empty[s], (uint32_t)(use - 1) & 1u);
            double2 *const st = stage_base + (size_t)s * a.stage_stride2;
            const size_t start = (size_t)(c_lo + i) * P;
            if (lane == 0) mbar_arrive_expect_tx(&full[s], bytes);
            __syncwarp();
            if (lane < rows) {
                const size_t row = (size_t)(k_base + lane);
                tma_load_1d(st + P + (size_t)lane * a.row_stride2, a.basis + row * (size_t)a.n_pad + start,
                            (uint32_t)P * 16u, &full[s]);
            }
            if (lane == 0) {
                tma_load_1d(st, a.ty + start, (uint32_t)P * 16u, &full[s]);
            } else if (PRE && lane == 1) {
                tma_load_1d(st + a.w_off2, a.wfrag + (size_t)blockIdx.x * (size_t)a.wf_tile_stride + start * 64,
                            (uint32_t)P * 512u, &full[s]);
            }
            // series tile: one copy per point, rows skewed by 32 B in shared memory (conflict-free B fragments)
            const double *const ysrc = a.Yt + ((size_t)sb * a.n_pad + start) * NS;
            double *const ydst = reinterpret_cast<double *>(st + a.y_off2);
            for (int q = lane; q < P; q += 32)
                tma_load_1d(ydst + (size_t)q * (NS + 4), ysrc + (size_t)q * NS, (uint32_t)NS * 8u, &full[s]);
            if (++s == kMmaStages) { s = 0; ++use; }
        }
        return;
    }

    // ---- compute warps
    const int wf = warp % kMmaWF, wt = warp / kMmaWF;
    const int r = lane >> 2, p = lane & 3;
    const int k = k_base + wf * 8 + r, j = j_base + wt;
    const bool ok = k < a.nf && j < a.nt;
    const int kc = min(k, a.nf - 1), jc = min(j, a.nt - 1);
    const double kap = a.kappa[kc];
    const double tj = a.tau[jc];
    const int row_off = P + (kc - k_base) * a.row_stride2;

    double W = 0.0, Q = 0.0, Ws = 0.0, Wc = 0.0;
    double Cy[NT][2], Cs[NT][2], Cc[NT][2];
#pragma unroll
    for (int m = 0; m < NT; ++m) Cy[m][0] = Cy[m][1] = Cs[m][0] = Cs[m][1] = Cc[m][0] = Cc[m][1] = 0.0;

    int cw_lo, cw_hi;
    warp_window(a, kc, jc, k < a.nf, lane, cw_lo, cw_hi);
    {
        int s = 0, use = 0;
        for (int i = 0; i < nchunks; ++i) {
            mbar_wait(&full[s], (uint32_t)use & 1u);
            if (c_lo + i < cw_lo || c_lo + i >= cw_hi) {       // staged for the tile, outside this warp's own window
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                if (++s == kMmaStages) { s = 0; ++use; }
                continue;
            }
            const double2 *const st = stage_base + (size_t)s * a.stage_stride2;
            const double2 *const row = st + row_off;
            const double *const yv = reinterpret_cast<const double *>(st + a.y_off2) + r;
            const long long left = a.n - (long long)(c_lo + i) * P;     // points of this chunk that exist
            const double *const wst = reinterpret_cast<const double *>(st + a.w_off2) + warp * (P / 4) * 32 + lane;
#pragma unroll 2
            for (int q = p; q < P; q += 4) {
                const double2 sc = row[q];
                double w;
                if (PRE) {
                    w = wst[(q >> 2) * 32];
                } else {
                    w = gauss_weight_estrin(kap * (st[q].x - tj), tbl);
                    if (q >= left) w = 0.0;
                }
                const double ws = w * sc.x, wc = w * sc.y;
                if (!PRE) {
                    W += w;
                    Q = fma(w, w, Q);
                    Ws += ws;
                    Wc += wc;
                }
                const double *const yq = yv + q * (NS + 4);
#pragma unroll
                for (int m = 0; m < NT; ++m) {
                    const double b = yq[8 * m];
                    dmma884(Cy[m][0], Cy[m][1], w, b);
                    dmma884(Cs[m][0], Cs[m][1], ws, b);
                    dmma884(Cc[m][0], Cc[m][1], wc, b);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == kMmaStages) { s = 0; ++use; }
        }
    }

    if (!PRE) {
        // quad reduction of the y-independent sums: every lane of a quad then holds the totals of its cell
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            W += __shfl_xor_sync(0xffffffffu, W, o);
            Q += __shfl_xor_sync(0xffffffffu, Q, o);
            Ws += __shfl_xor_sync(0xffffffffu, Ws, o);
            Wc += __shfl_xor_sync(0xffffffffu, Wc, o);
        }
    }
    if (!ok) return;
    const size_t cell = (size_t)j * (size_t)a.nf + (size_t)k;   // output order: tau-major rows
    if (PRE) {
        const size_t nc = (size_t)a.ncell;
        W = a.cellsums[cell];
        Q = a.cellsums[nc + cell];
        Ws = a.cellsums[2 * nc + cell];
        Wc = a.cellsums[3 * nc + cell];
    }
    const double neff = W * W / Q;
    if (!PRE && sb == 0 && p == 0) {
        a.neffs[cell] = neff;
        a.wq[cell] = W;
        a.wq[(size_t)a.ncell + cell] = Q;
    }
    const bool valid = neff > a.thr;
    const double S1 = Ws / W, C1 = Wc / W;
    const double2 rt = a.rot[cell];
    // C fragment: row = lane/4 (the cell), columns 2*(lane%4) + {0, 1} of n-tile m
    const size_t o = cell * (size_t)a.S_pad + (size_t)sb * NS + (size_t)(2 * p);
    // the amplitudes may go to the owner of the cell's block over NVLink (peer memory), see AmpScatter
    double *const arow = amp_row(a.sc, a.amp, cell, a.S_pad) + (size_t)sb * NS + (size_t)(2 * p);
#pragma unroll
    for (int m = 0; m < NT; ++m) {
        double amp2[2], ph2[2], c02[2], c12[2], c22[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double c0 = CUDART_NAN, c1 = CUDART_NAN, c2 = CUDART_NAN;
            if (valid) {
                const double Yb = Cy[m][e] / W;
                const double cyc = Cc[m][e] / W - Yb * C1;
                const double cys = Cs[m][e] / W - Yb * S1;
                c0 = Yb;
                c1 = 2.0 * (rt.y * cyc + rt.x * cys);
                c2 = 2.0 * (rt.y * cys - rt.x * cyc);
            }
            amp2[e] = sqrt(c1 * c1 + c2 * c2);
            ph2[e] = a.phase ? atan2(c2, c1) : 0.0;
            c02[e] = c0;
            c12[e] = c1;
            c22[e] = c2;
        }
        const size_t oo = o + (size_t)(8 * m);
        *reinterpret_cast<double2 *>(arow + (size_t)(8 * m)) = make_double2(amp2[0], amp2[1]);
        if (a.phase) *reinterpret_cast<double2 *>(a.phase + oo) = make_double2(ph2[0], ph2[1]);
        if (a.a0) *reinterpret_cast<double2 *>(a.a0 + oo) = make_double2(c02[0], c02[1]);
        if (a.a1) *reinterpret_cast<double2 *>(a.a1 + oo) = make_double2(c12[0], c12[1]);
        if (a.a2) *reinterpret_cast<double2 *>(a.a2 + oo) = make_double2(c22[0], c22[1]);
    }
}

int shared_series_block(int64_t S)
{
    const char *e = getenv("WWZB200_SERIES_BLOCK");
    if (e && *e) return atoi(e) == 16 ? 16 : 64;
    return S > 32 ? 64 : 16;
}

size_t stored_weights_budget()
{
    const char *e = getenv("WWZB200_WEIGHTS_BYTES");
    if (e && *e) return (size_t)atof(e);
    return (size_t)8 << 30;
}

SharedPlan plan_shared(int64_t n, int64_t nt, int64_t nf, int64_t S)
{
    SharedPlan p{};
    p.series_block = shared_series_block(S);
    int64_t P = 32;
    {
        const char *e = getenv("WWZB200_SHARED_POINTS");
        if (e && *e) P = std::max(8, std::min(128, atoi(e))) & ~3;
    }
    if (P > n) P = ((n + 7) / 8) * 8;
    p.points = (int)P;
    p.n_pad = ((n + P - 1) / P) * P;
    p.row_stride = (P + 4) * 16;                 // {s,c} rows: 64 B skew, conflict-free LDS.128 over (row, point) quads
    int64_t y_off = P * 16 + (int64_t)kMmaRows * p.row_stride;
    y_off = (y_off + 127) & ~(int64_t)127;
    p.y_off = y_off;
    int64_t w_off = y_off + P * (p.series_block + 4) * 8;   // series rows skewed by 32 B
    w_off = (w_off + 127) & ~(int64_t)127;
    p.w_off = w_off;
    // stored-weights mode: worth it once several series blocks reuse the weights, if the fragments fit the budget
    const int64_t tiles = ((nf + kMmaRows - 1) / kMmaRows) * ((nt + kMmaWT - 1) / kMmaWT);
    p.weights_count = (size_t)tiles * (size_t)p.n_pad * 64;
    const int64_t blocks = (S + p.series_block - 1) / p.series_block;
    p.stored_weights = blocks >= 4 && p.weights_count * 8 <= stored_weights_budget();
    {
        const char *e = getenv("WWZB200_STORED_WEIGHTS");
        if (e && *e) p.stored_weights = atoi(e) != 0 && p.weights_count * 8 <= stored_weights_budget();
    }
    p.stage_bytes = (int)(p.stored_weights ? w_off + P * 512 : w_off);
    p.smem_bytes = kMmaStages * p.stage_bytes + 2 * kMmaStages * 8 + 16 * 8 + 16;
    return p;
}

template <int NS, bool PRE>
static cudaError_t launch_mma(const MmaArgs &a, dim3 grid, int smem, cudaStream_t st)
{
    static SmemOptIn done;
    cudaError_t e = ensure_dynamic_smem(wwz_shared_mma_kernel<NS, PRE>, smem, done);
    if (e != cudaSuccess) return e;
    wwz_shared_mma_kernel<NS, PRE><<<grid, kMmaThreads, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_shared(const SharedPlan &plan, const double2 *d_ty, const double2 *d_basis, const double *d_Yt,
                          const double *d_ts, int64_t n, int64_t S_pad, const double *d_tau, int64_t nt,
                          const double *d_kappa, const double *d_dcut, const double *d_taugap, const int *d_unsorted,
                          int64_t nf, const double2 *d_rot, double thr, double *d_wfrag, double *d_cellsums, double *d_neffs,
                          double *d_wq, double *d_amp, double *d_phase, double *d_a0, double *d_a1, double *d_a2,
                          cudaStream_t st, int phase, const AmpScatter *scatter)
{
    const int64_t ncell = nt * nf;
    if (ncell <= 0 || S_pad <= 0) return cudaSuccess;
    MmaArgs a{};
    a.ty = d_ty;
    a.basis = d_basis;
    a.Yt = d_Yt;
    a.ts = d_ts;
    a.tau = d_tau;
    a.kappa = d_kappa;
    a.dcut = d_dcut;
    a.taugap = d_taugap;
    a.unsorted = d_unsorted;
    a.rot = d_rot;
    a.neffs = d_neffs;
    a.wq = d_wq;
    a.amp = d_amp;
    a.phase = d_phase;
    a.a0 = d_a0;
    a.a1 = d_a1;
    a.a2 = d_a2;
    a.wfrag = d_wfrag;
    a.cellsums = d_cellsums;
    a.n = n;
    a.n_pad = plan.n_pad;
    a.ncell = ncell;
    a.S_pad = S_pad;
    a.wf_tile_stride = plan.n_pad * 64;
    a.nt = (int)nt;
    a.nf = (int)nf;
    a.P = plan.points;
    a.tiles_t = (int)((nt + kMmaWT - 1) / kMmaWT);
    a.row_stride2 = (int)(plan.row_stride / 16);
    a.stage_stride2 = plan.stage_bytes / 16;
    a.y_off2 = (int)(plan.y_off / 16);
    a.w_off2 = (int)(plan.w_off / 16);
    a.thr = thr;
    static_assert(sizeof(AmpScatter) == sizeof(AmpScatterDev), "host and device views of the scatter map");
    if (scatter) memcpy(&a.sc, scatter, sizeof a.sc);
    const int NS = plan.series_block;
    const int64_t tiles_k = (nf + kMmaRows - 1) / kMmaRows;
    const dim3 grid((unsigned)(tiles_k * a.tiles_t), (unsigned)(S_pad / NS));
    cudaError_t e;
    if (plan.stored_weights) {
        if (!d_wfrag || !d_cellsums) return cudaErrorInvalidValue;
        if (phase & 1) {     // the y-independent pass: weights in fragment order, W, Q, Ws, Wc, Neff of every cell
            wwz_tile_weights_kernel<<<grid.x, kMmaWarps * 32, 0, st>>>(a);
            bump_launch_counter();
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
        }
        if (!(phase & 2)) return cudaSuccess;
        e = NS == 64 ? launch_mma<64, true>(a, grid, plan.smem_bytes, st) : launch_mma<16, true>(a, grid, plan.smem_bytes, st);
    } else {
        if (!(phase & 2)) return cudaSuccess;
        e = NS == 64 ? launch_mma<64, false>(a, grid, plan.smem_bytes, st) : launch_mma<16, false>(a, grid, plan.smem_bytes, st);
    }
    bump_launch_counter();
    return e;
}

}  // namespace wwz
