// tau-recurrence cell kernel with warp-autonomous tiles (sm_100a).
//
// Same arithmetic as wwz_cells_kernel<8, true> (wwz_kernels.cu: one exp2 pair per 8 tau-adjacent cells of
// a lane for equally spaced tau, utils/wavelet.py:988-1023 reduced to seven weighted sums), different
// work decomposition:
//
//  * the unit of work is a WARP TILE of 32 lane slots = TG tau-groups x TR frequency rows (TG*TR = 32,
//    a tau-group = 8 adjacent cells of one row), i.e. a tile compact in tau AND in frequency, so that the
//    support of its Gaussian weights is a short index range of the sorted time axis;
//  * every warp finds the range of its own tile (warp-cooperative 32-ary search on ts) and walks only
//    that: points whose weight is below 2^-100 of the largest weight of every cell of the tile are never touched;
//  * every warp stages its own points: a private 3-stage shared-memory ring filled by 1-D TMA bulk copies
//    (cp.async.bulk + mbarrier complete_tx) that the warp issues itself -- no producer warp, no CTA-wide
//    barrier in the loop, 12 compute warps per SM (3 CTAs x 4 warps at 168 registers) instead of 8.
//
// Skipped weights.  The window of a tile reaches sqrt(dcut^2 + gap^2) beyond its taus, gap = distance from tau to
// the nearest sample (tau_gap_kernel): a skipped weight is below 2^-100 of the largest weight of the cell, hence of
// W = sum(w) -- 2^-47 ulp -- wherever tau lies, also inside a hiatus (the window widens by itself there).
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>

#include "wwz_device.cuh"
#include "wwz_internal.h"

namespace wwz {

struct RecArgs {
    const double2 *ty;
    const double2 *basis;
    const double *ts;
    const double *tau;
    const double *kappa;
    const double *dcut;
    const double *taugap;
    const int *unsorted;
    const unsigned char *rec_row;
    const double *grid;
    double *sums;
    unsigned long long *walked;   // profiling only (may be null): (lane, point) pairs actually evaluated, summed over warps
    long long n, n_pad, ncell_total, seg_len, ntiles;
    int nt, G, nf_batch, k0, nf_total, P, tg_log2, tiles_g;
    int row_stride2;     // basis row stride inside a stage, double2 units
    int stage_stride2;   // stage stride, double2 units
    int warp_stride;     // bytes of shared memory per warp (ring + barriers)
};

constexpr int kRecWarps = 4;     // warps per CTA; every warp is an independent worker

__global__ void __launch_bounds__(kRecWarps * 32, 3) wwz_rec_kernel(const RecArgs a)
{
    constexpr int M = kRecCells;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double tbl[16];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    load_exp2_table(tbl, tid);
    __syncthreads();                     // the only CTA-wide barrier: warps are independent from here on

    const long long tile = (long long)blockIdx.x * kRecWarps + warp;
    if (tile >= a.ntiles) return;
    const int kt = (int)(tile / a.tiles_g), gt = (int)(tile - (long long)kt * a.tiles_g);
    const int TG = 1 << a.tg_log2, TR = 32 >> a.tg_log2;
    const int g = gt * TG + (lane & (TG - 1));
    const int r_loc = lane >> a.tg_log2;
    const int kb = kt * TR + r_loc;                       // row inside the frequency batch
    const int rows = min(TR, a.nf_batch - kt * TR);       // rows of this tile that exist
    bool ok = g < a.G && kb < a.nf_batch;
    if (ok) ok = a.rec_row[a.k0 + kb] != 0;
    if (!__any_sync(0xffffffffu, ok)) return;             // every row of the tile belongs to the direct kernel

    const int P = a.P;
    const long long p0 = (long long)blockIdx.y * a.seg_len;
    const long long p1 = min(a.n, p0 + a.seg_len);

    // ---- lane constants and the tile's point window
    double kap = 0.0, tau0 = 0.0;
    double wlo = CUDART_INF, whi = -CUDART_INF;
    if (ok) {
        kap = a.kappa[a.k0 + kb];
        const double dc = a.dcut[a.k0 + kb];
        tau0 = a.tau[g * M];
        double gmax = 0.0;                                   // widest sample gap seen by the lane's cells
#pragma unroll
        for (int r = 0; r < M; ++r) gmax = fmax(gmax, a.taugap[min(g * M + r, a.nt - 1)]);
        const double de = sqrt(fma(gmax, gmax, dc * dc));   // cut relative to the cells' own largest weight
        wlo = tau0 - de;
        whi = a.tau[min(g * M + M - 1, a.nt - 1)] + de;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wlo = fmin(wlo, __shfl_xor_sync(0xffffffffu, wlo, o));
        whi = fmax(whi, __shfl_xor_sync(0xffffffffu, whi, o));
    }
    long long i0 = p0, i1 = p1;
    if (!*a.unsorted) {
        i0 = warp_bound<false>(a.ts, p0, p1, wlo, lane);
        i1 = warp_bound<true>(a.ts, i0, p1, whi, lane);
    }
    const int c_lo = (int)((i0 - p0) / P);
    const int nchunks = (i1 > i0) ? (int)((i1 - p0 + P - 1) / P) - c_lo : 0;

    // ---- this warp's ring
    unsigned char *const wbase = smem_raw + (size_t)warp * a.warp_stride;
    double2 *const stage_base = reinterpret_cast<double2 *>(wbase);
    uint64_t *const full = reinterpret_cast<uint64_t *>(wbase + (size_t)kStages * a.stage_stride2 * sizeof(double2));
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncwarp();
    const uint32_t bytes = (uint32_t)P * 16u + (uint32_t)rows * (uint32_t)P * 32u;
    auto issue = [&](int c, int s) {
        // lanes 0..rows each start one bulk copy of chunk c into stage s; the stage was last read before
        // the __syncwarp that precedes every call
        double2 *const st = stage_base + (size_t)s * a.stage_stride2;
        const size_t start = (size_t)p0 + (size_t)c * P;
        if (lane == 0) mbar_arrive_expect_tx(&full[s], bytes);
        __syncwarp();
        if (lane == 0) {
            tma_load_1d(st, a.ty + start, (uint32_t)P * 16u, &full[s]);
        } else if (lane <= rows) {
            const size_t row = (size_t)(kt * TR + lane - 1);
            tma_load_1d(st + P + (size_t)(lane - 1) * a.row_stride2, a.basis + (row * (size_t)a.n_pad + start) * 2,
                        (uint32_t)P * 32u, &full[s]);
        }
    };
#pragma unroll
    for (int s = 0; s < kStages; ++s)
        if (s < nchunks) issue(c_lo + s, s);

    double W[M], Q[M], Ws[M], Wc[M], Wy[M], Wys[M], Wyc[M];
#pragma unroll
    for (int r = 0; r < M; ++r) W[r] = Q[r] = Ws[r] = Wc[r] = Wy[r] = Wys[r] = Wyc[r] = 0.0;

    const double dtau = a.grid[0];
    const double step16 = 2.0 * kap * dtau;
    const int row_off = P + r_loc * a.row_stride2;
    int s = 0;
    uint32_t phase = 0;
    for (int i = 0; i < nchunks; ++i) {
        mbar_wait(&full[s], phase);
        const double2 *const st = stage_base + (size_t)s * a.stage_stride2;
        const double2 *const row = st + row_off;
        const int cnt = (int)min((long long)P, p1 - p0 - (long long)(c_lo + i) * P);
#pragma unroll 2
        for (int p = 0; p < cnt; ++p) {
            const double2 t_y = st[p];
            const double2 sc = row[2 * p];
            const double2 ysc = row[2 * p + 1];
            double w0, rho;
            gauss_weight_ratio(kap * (t_y.x - tau0), step16, tbl, w0, rho);
            double wv[M];
            static_assert(M == 8, "the power tree below is written for 8 cells per lane");
            const double rho2 = rho * rho, rho4 = rho2 * rho2;
            wv[0] = w0;
            wv[1] = w0 * rho;
            wv[2] = w0 * rho2;
            wv[3] = wv[1] * rho2;
            wv[4] = w0 * rho4;
            wv[5] = wv[1] * rho4;
            wv[6] = wv[2] * rho4;
            wv[7] = wv[3] * rho4;
#pragma unroll
            for (int r = 0; r < M; ++r) {
                const double w = wv[r];
                W[r] += w;
                Q[r] = fma(w, w, Q[r]);
                Ws[r] = fma(w, sc.x, Ws[r]);
                Wc[r] = fma(w, sc.y, Wc[r]);
                Wy[r] = fma(w, t_y.y, Wy[r]);
                Wys[r] = fma(w, ysc.x, Wys[r]);
                Wyc[r] = fma(w, ysc.y, Wyc[r]);
            }
        }
        __syncwarp();
        if (i + kStages < nchunks) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(c_lo + i + kStages, s);
        }
        if (++s == kStages) { s = 0; phase ^= 1u; }
    }

    if (a.walked && lane == 0 && nchunks > 0) {
        const long long first = (long long)c_lo * P, last = min(p1 - p0, (long long)(c_lo + nchunks) * P);
        atomicAdd(a.walked, (unsigned long long)(last - first));
    }
    if (ok) {
        const long long kcol = (long long)a.k0 + kb;
        const double D = kap * dtau;
        const size_t nc = (size_t)a.ncell_total;
#pragma unroll
        for (int r = 0; r < M; ++r) {
            const int j = g * M + r;
            if (j < a.nt) {
                const size_t cell = (size_t)blockIdx.y * (size_t)kNumSums * nc + (size_t)j * (size_t)a.nf_total + (size_t)kcol;
                // undo the per-cell scale h_r = 2^(-(r D)^2/16) that was factored out of the weights
                const double h = exp2(-(double)(r * r) * D * D * 0.0625);
                a.sums[0 * nc + cell] = W[r] * h;
                a.sums[1 * nc + cell] = Q[r] * h * h;
                a.sums[2 * nc + cell] = Ws[r] * h;
                a.sums[3 * nc + cell] = Wc[r] * h;
                a.sums[4 * nc + cell] = Wy[r] * h;
                a.sums[5 * nc + cell] = Wys[r] * h;
                a.sums[6 * nc + cell] = Wyc[r] * h;
            }
        }
    }
}

RecPlan plan_rec(int64_t n, int64_t nt, int64_t nf, int64_t seg_len)
{
    RecPlan p{};
    p.groups = (int)((nt + kRecCells - 1) / kRecCells);
    int tg_log2 = 3;                                    // 8 tau-groups x 4 rows per warp tile
    {
        const char *e = getenv("WWZB200_REC_TG_LOG2");
        if (e && *e) tg_log2 = std::max(0, std::min(5, atoi(e)));
    }
    while (tg_log2 > 0 && (1 << (tg_log2 - 1)) >= p.groups) --tg_log2;   // no wider than the tau axis
    p.tg_log2 = tg_log2;
    const int TR = 32 >> tg_log2;
    // ring budget per warp ~14 KB (12 warps per SM -> ~170 KB of the 227 KB)
    int64_t P = (14336 / kStages) / (16 + 32 * TR);
    P &= ~(int64_t)3;
    const int64_t span = seg_len > 0 ? std::min(seg_len, n) : n;
    if (P > span) P = ((span + 3) / 4) * 4;
    if (P < 4) P = 4;
    p.points = (int)P;
    p.n_pad = ((n + P + 15) / 16) * 16;
    p.row_stride = (P + 1) * 32;
    p.stage_bytes = (int)(P * 16 + (int64_t)TR * p.row_stride);
    p.warp_bytes = ((kStages * p.stage_bytes + kStages * 8 + 127) / 128) * 128;
    p.smem_bytes = kRecWarps * p.warp_bytes;
    return p;
}

cudaError_t launch_rec(const RecPlan &plan, const double2 *d_ty, const double2 *d_basis, const double *d_ts, int64_t n,
                       const double *d_tau, int64_t nt, const double *d_kappa, const double *d_dcut,
                       const double *d_taugap, const int *d_unsorted, int64_t k0, int64_t nf_batch, int64_t nf_total, int segs,
                       int64_t seg_len, const unsigned char *d_rec_row, const double *d_grid, double *d_sums,
                       unsigned long long *d_walked, cudaStream_t st)
{
    if (nf_batch <= 0 || nt <= 0) return cudaSuccess;
    RecArgs a{};
    a.ty = d_ty;
    a.basis = d_basis;
    a.ts = d_ts;
    a.tau = d_tau;
    a.kappa = d_kappa;
    a.dcut = d_dcut;
    a.taugap = d_taugap;
    a.unsorted = d_unsorted;
    a.rec_row = d_rec_row;
    a.grid = d_grid;
    a.sums = d_sums;
    a.walked = d_walked;
    a.n = n;
    a.n_pad = plan.n_pad;
    a.ncell_total = nt * nf_total;
    a.seg_len = seg_len;
    a.nt = (int)nt;
    a.G = plan.groups;
    a.nf_batch = (int)nf_batch;
    a.k0 = (int)k0;
    a.nf_total = (int)nf_total;
    a.P = plan.points;
    a.tg_log2 = plan.tg_log2;
    const int TG = 1 << plan.tg_log2, TR = 32 >> plan.tg_log2;
    a.tiles_g = (plan.groups + TG - 1) / TG;
    const int64_t tiles_k = (nf_batch + TR - 1) / TR;
    a.ntiles = (int64_t)a.tiles_g * tiles_k;
    a.row_stride2 = (int)(plan.row_stride / 16);
    a.stage_stride2 = plan.stage_bytes / 16;
    a.warp_stride = plan.warp_bytes;
    static SmemOptIn done;
    cudaError_t e = ensure_dynamic_smem(wwz_rec_kernel, plan.smem_bytes, done);
    if (e != cudaSuccess) return e;
    const int64_t ctas = (a.ntiles + kRecWarps - 1) / kRecWarps;
    wwz_rec_kernel<<<dim3((unsigned)ctas, (unsigned)segs), kRecWarps * 32, plan.smem_bytes, st>>>(a);
    bump_launch_counter();
    return cudaGetLastError();
}

}  // namespace wwz
