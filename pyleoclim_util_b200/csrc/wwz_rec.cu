// Warp-tile cell kernels (sm_100a): the tau-recurrence flavour (default on equally spaced tau) and a direct
// flavour with tau-compact tiles for small tau grids.
//
// Same seven weighted sums as wwz_cells_kernel (wwz_kernels.cu; utils/wavelet.py:988-1023 reduced to
// W, Q, Ws, Wc, Wy, Wys, Wyc), different work decomposition:
//
//  * the unit of work is a WARP TILE of 32 lane slots = TG tau-groups x TR frequency rows (TG*TR = 32; a
//    tau-group = M adjacent cells of one row), i.e. a tile compact in tau AND in frequency, so that the
//    support of its Gaussian weights is a short index range of the sorted time axis;
//  * every warp finds the range of its own tile (warp-cooperative 32-ary search on ts) and walks only
//    that: points whose weight is below 2^-100 of the largest weight of every cell of the tile are never touched;
//  * every warp stages its own points: a private shared-memory ring (three stages; two longer ones with the growth
//    table of rho, see the REC loop) filled by 1-D TMA bulk copies
//    (cp.async.bulk + mbarrier complete_tx) that the warp issues itself -- no producer warp, no CTA-wide
//    barrier in the loop;
//  * the grid is PERSISTENT: one launch = (SMs x resident CTAs) CTAs whose warps pull (tile, point segment) items
//    from a device-side queue, longest windows first (low frequencies first), so there is no partial last wave;
//    the last warp to leave resets the queue for the next launch on the stream.
//
// REC = true (M = 8): one exp2 pair per 8 tau-adjacent cells of a lane for equally spaced tau, then w_0 rho^m through
// a depth-3 product tree; two points per iteration so that four exp2 chains interleave.  (Measured on B200 and not
// kept: hand-pipelining the next point's exp2 pair under the current point's accumulations, and a two-phase loop with
// a per-lane weight buffer in shared memory -- at 168 registers ptxas serialises the former, and both lose more to
// same-bank operand fetches of the accumulating DFMAs than they gain in latency; a register-only microbenchmark of the
// accumulation pattern, tools/rec_micro.cu, runs at 94% of the pipe.)
// REC = false (M = 2): one exp2 per (cell, point), software-pipelined over the points, any tau axis; used
// for the rows the recurrence cannot take when the tau grid is small (the reference default, ntau = 50:
// core/series.py:3545-3551), where a tile of the CTA-wide kernel would span the whole tau axis and its point window
// the whole series.
//
// Skipped weights.  The window of a tile reaches sqrt(dcut^2 + gap^2) beyond its taus, gap = distance from tau to
// the nearest sample (setup_kernel): a skipped weight is below 2^-100 of the largest weight of the cell, hence of
// W = sum(w) -- 2^-47 ulp -- wherever tau lies, also inside a hiatus (the window widens by itself there).
#include <math_constants.h>

#include <algorithm>
#include <cstdio>
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
    const double *estep;          // REC, optional: [member][row tile][n_pad][TR] growth of rho from a point to the next
    int e_off2;                   // offset of the [P][TR] growth block inside a stage, double2 units
    int stages;                   // stages of a warp's ring in use (<= kStages)
    unsigned int *queue;          // [0] next work item, [1] warps that have left (self-resetting)
    unsigned long long *walked;   // profiling only (may be null): (lane, point) pairs actually evaluated, summed over warps
    long long n, n_pad, ncell_total, seg_len, ntiles, nitems;
    long long m_ts, m_tau;        // members of a batch: element strides of ts and tau (scratch is [member][...])
    int members;
    int nt, G, nf_batch, k0, nf_total, P, tg_log2, tiles_g, segs;
    int row_stride2;     // basis row stride inside a stage, double2 units (odd multiple of 16 B: conflict-free rows)
    int stage_stride2;   // stage stride, double2 units
    int warp_stride;     // bytes of shared memory per warp (ring + barriers)
    // ---- fused epilogue (FUSE = true): the warp that finishes the last point segment of a tile adds the slabs in
    // order and writes the tile's finished cells -- no separate epilogue launch, and finished tau blocks can leave for
    // the host while later tiles are still being computed
    const double *omega;         // [member][nf_total] (m_omega = element stride between members)
    long long m_omega, row_stride;
    double thr;
    double *o_neffs, *o_a0, *o_a1, *o_a2, *o_wwa, *o_phase;   // [member][nt][row_stride]; all but o_neffs may be null
    unsigned int *tile_cnt;      // [member][tile] point segments finished (left zeroed by the launch)
    unsigned int *fix_count, *fix_list;   // deep-underflow cells for faithful_cells_kernel (as finalize_kernel)
    unsigned int fix_cap;
    unsigned int *unfused;       // number of tiles skipped here (their rows belong to the direct flavour: finalize_kernel)
    unsigned int *blk_cnt;       // null = no progress reports; [tau block] finished work items (left zeroed)
    volatile unsigned int *blk_flag;   // page-locked host memory, mapped: [tau block] := blk_epoch when the block is final
    unsigned int blk_epoch;
    int blk_gt;                  // tiles along tau per block
#ifdef WWZ_TRACE
    unsigned long long *trace;   // developer build only (-DWWZ_TRACE): [item][8] timeline of the work items, tools/trace_items.py
#endif
};

// Growth of rho = 2^(kappa4^2 dtau (t - tau_0)/8) from point i to point i+1: E = 2^(kappa4^2 dtau (t_{i+1} - t_i)/8).  It depends
// on (row, point) only, so the recurrence flavour advances rho with one multiplication per point instead of a second
// exp2 chain per (lane, point).  Laid out [row tile][point][TR]: a chunk of a tile is one contiguous bulk copy.
__global__ void __launch_bounds__(256) rho_growth_kernel(const RecArgs a, double *__restrict__ estep)
{
    const int TR = 32 >> a.tg_log2;
    const long long per_mem = (long long)((a.nf_batch + TR - 1) / TR) * a.n_pad * TR;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= per_mem * a.members) return;
    const int mem = (int)(e / per_mem);
    const long long q = e - (long long)mem * per_mem;
    const int r = (int)(q % TR);
    const long long i = (q / TR) % a.n_pad;
    const int kb = (int)(q / ((long long)TR * a.n_pad)) * TR + r;
    double v = 1.0;
    if (kb < a.nf_batch && i + 1 < a.n) {
        const double *const ts = a.ts + (long long)mem * a.m_ts;
        const double kap = a.kappa[(size_t)mem * a.nf_total + a.k0 + kb];
        const double step16 = 2.0 * kap * a.grid[2 * mem];
        v = exp2(kap * (ts[i + 1] - ts[i]) * step16 * 0.0625);
    }
    estep[e] = v;
}

// NREG = register budget per thread (__maxnreg__: ptxas interleaves the weight evaluation of the next point with the
// accumulations of the current one only when it has the registers for both).  Every warp of a CTA is an independent
// worker; the number of warps per CTA is a launch parameter.
template <int M, bool REC, int NREG, bool FUSE>
__global__ void __maxnreg__(NREG) wwz_tile_kernel(const RecArgs a)
{
    static_assert(!FUSE || REC, "the fused epilogue exists for the recurrence flavour");
    const int kRecWarps = blockDim.x >> 5;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double tbl[16];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    load_exp2_table(tbl, tid);
    __syncthreads();                     // the only CTA-wide barrier: warps are independent from here on

    // ---- this warp's ring
    const int P = a.P;
    unsigned char *const wbase = smem_raw + (size_t)warp * a.warp_stride;
    double2 *const stage_base = reinterpret_cast<double2 *>(wbase);
    uint64_t *const full = reinterpret_cast<uint64_t *>(wbase + (size_t)a.stages * a.stage_stride2 * sizeof(double2));
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncwarp();
    int s = 0;               // next stage to consume
    uint32_t ph = 0;         // parity of every stage's barrier (bit s)

    const int TG = 1 << a.tg_log2, TR = 32 >> a.tg_log2;
    const int r_loc = lane >> a.tg_log2;
    unsigned long long walked = 0;

    for (;;) {
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(&a.queue[0], 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if ((long long)item >= a.nitems) break;
#ifdef WWZ_TRACE
        unsigned long long tr0 = 0, tr1 = 0, tr_first = 0, tr_loop = 0;
        long long clk_pts = 0, clk_mark = 0;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tr0));
#endif
        // items are numbered tile-major (low frequencies = long windows first), then member, point segment fastest;
        // with progress reports (FUSE, one member): tau block by tau block, inside a block again lowest rows first, so
        // that whole tau blocks of the output are final early and every block still starts with its longest windows
        long long tile;
        int mem, seg, kt, gt;
        if (FUSE && a.blk_gt > 0) {
            const unsigned int tiles_k = (unsigned)(a.ntiles / a.tiles_g);
            const unsigned int full = tiles_k * (unsigned)a.blk_gt * (unsigned)a.segs;   // items of a whole block
            const unsigned int b = item / full;
            const unsigned int rem = item - b * full;
            const unsigned int ngt = (unsigned)min(a.blk_gt, a.tiles_g - (int)b * a.blk_gt);
            kt = (int)(rem / (ngt * (unsigned)a.segs));
            const unsigned int r2 = rem - (unsigned)kt * ngt * (unsigned)a.segs;
            gt = (int)b * a.blk_gt + (int)(r2 / (unsigned)a.segs);
            seg = (int)(r2 % (unsigned)a.segs);
            mem = 0;
            tile = (long long)kt * a.tiles_g + gt;
        } else {
            const unsigned int per_tile = (unsigned)a.segs * (unsigned)a.members;
            tile = item / per_tile;
            const unsigned int rem = item - (unsigned)tile * per_tile;
            mem = (int)(rem / (unsigned)a.segs);
            seg = (int)(rem - (unsigned)mem * (unsigned)a.segs);
            kt = (int)(tile / a.tiles_g);
            gt = (int)(tile - (long long)kt * a.tiles_g);
        }
        // progress of the tile's tau block (FUSE): one count per work item once nothing of it is left to write
        auto report = [&]() {
            if constexpr (FUSE) {
                if (a.blk_gt > 0 && lane == 0) {
                    const int tiles_k = (int)(a.ntiles / a.tiles_g);
                    const int b = gt / a.blk_gt;
                    const unsigned int target =
                        (unsigned)(min(a.blk_gt, a.tiles_g - b * a.blk_gt) * tiles_k) * (unsigned)a.segs;
                    __threadfence();
                    if (atomicAdd(&a.blk_cnt[b], 1u) == target - 1) {
                        a.blk_cnt[b] = 0;
                        __threadfence_system();
                        a.blk_flag[b] = a.blk_epoch;
                    }
                }
            }
        };
        // this member's inputs and scratch
        const double2 *const m_ty = a.ty + (size_t)mem * (size_t)a.n_pad;
        const double2 *const m_basis = a.basis + (size_t)mem * (size_t)a.nf_batch * (size_t)a.n_pad * 2;
        const double *const m_tsp = a.ts + (long long)mem * a.m_ts;
        const double *const m_taup = a.tau + (long long)mem * a.m_tau;
        const double *const m_kappa = a.kappa + (size_t)mem * a.nf_total;
        const double *const m_dcut = a.dcut + (size_t)mem * a.nf_total;
        const double *const m_gap = a.taugap + (size_t)mem * a.nt;
        const unsigned char *const m_rec = a.rec_row ? a.rec_row + (size_t)mem * a.nf_total : nullptr;
        double *const m_sums = a.sums + (size_t)mem * (size_t)a.segs * kNumSums * (size_t)a.ncell_total;
        const double dtau = REC ? a.grid[2 * mem] : 0.0;
        const bool sorted = !a.unsorted[mem * kFlagInts];
        const int g = gt * TG + (lane & (TG - 1));
        const int kb = kt * TR + r_loc;                       // row inside the frequency batch
        const int rows = min(TR, a.nf_batch - kt * TR);       // rows of this tile that exist
        bool ok = g < a.G && kb < a.nf_batch;
        if (ok && m_rec) ok = (m_rec[a.k0 + kb] != 0) == REC;
        if (!__any_sync(0xffffffffu, ok)) {                   // every row of the tile belongs to the other flavour
            if constexpr (FUSE) {
                if (seg == 0 && lane == 0) atomicAdd(a.unfused, 1u);
                report();
            }
            continue;
        }

        const long long p0 = (long long)seg * a.seg_len;
        const long long p1 = min(a.n, p0 + a.seg_len);

        // ---- lane constants and the tile's point window
        double kap = 0.0;
        double tau_r[M];
#pragma unroll
        for (int r = 0; r < M; ++r) tau_r[r] = 0.0;
        double wlo = CUDART_INF, whi = -CUDART_INF;
        if (ok) {
            kap = m_kappa[a.k0 + kb];
            const double dc = m_dcut[a.k0 + kb];
            double gmax = 0.0;                                   // widest sample gap seen by the lane's cells
            double tmin = CUDART_INF, tmax = -CUDART_INF;        // (the tau axis may come in any order)
#pragma unroll
            for (int r = 0; r < M; ++r) {
                const int j = min(g * M + r, a.nt - 1);
                tau_r[r] = m_taup[j];
                gmax = fmax(gmax, m_gap[j]);
                tmin = fmin(tmin, tau_r[r]);
                tmax = fmax(tmax, tau_r[r]);
            }
            if (dc >= 0.0) {
                const double de = sqrt(fma(gmax, gmax, dc * dc));   // cut relative to the cells' own largest weight
                wlo = tmin - de;
                whi = tmax + de;
            } else {                                             // NaN omega: no window constraint, NaN flows through
                wlo = -CUDART_INF;
                whi = CUDART_INF;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            wlo = fmin(wlo, __shfl_xor_sync(0xffffffffu, wlo, o));
            whi = fmax(whi, __shfl_xor_sync(0xffffffffu, whi, o));
        }
        long long i0 = p0, i1 = p1;
        if (sorted) {
            i0 = warp_bound<false>(m_tsp, p0, p1, wlo, lane);
            i1 = warp_bound<true>(m_tsp, i0, p1, whi, lane);
        }
        const int c_lo = (int)((i0 - p0) / P);
        const int nchunks = (i1 > i0) ? (int)((i1 - p0 + P - 1) / P) - c_lo : 0;

        const bool use_e = REC && a.estep != nullptr;
        const uint32_t e_bytes = use_e ? (uint32_t)P * (uint32_t)TR * 8u : 0u;
        const uint32_t bytes = (uint32_t)P * 16u + (uint32_t)rows * (uint32_t)P * 32u + e_bytes;
        auto issue = [&](int c, int sg) {
            // lane 0 starts the bulk copy of the (t, y) pairs, lanes 0..rows-1 one basis row each, of chunk c into
            // stage sg; the stage was last read before the __syncwarp that precedes every call
            double2 *const st = stage_base + (size_t)sg * a.stage_stride2;
            const size_t start = (size_t)p0 + (size_t)c * P;
            if (lane == 0) {
                mbar_arrive_expect_tx(&full[sg], bytes);
                tma_load_1d(st, m_ty + start, (uint32_t)P * 16u, &full[sg]);
            }
            __syncwarp();
            if (lane < rows) {
                const size_t row = (size_t)(kt * TR + lane);
                tma_load_1d(st + P + (size_t)lane * a.row_stride2, m_basis + (row * (size_t)a.n_pad + start) * 2,
                            (uint32_t)P * 32u, &full[sg]);
            } else if (use_e && lane == rows) {
                const double *const m_e = a.estep + (size_t)mem * (size_t)((a.nf_batch + TR - 1) / TR) * (size_t)a.n_pad * TR;
                tma_load_1d(st + a.e_off2, m_e + ((size_t)kt * (size_t)a.n_pad + start) * TR, e_bytes, &full[sg]);
            }
        };
        {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // (the ring was last read by the previous tile)
            int sg = s;
#pragma unroll
            for (int u = 0; u < kStages; ++u) {
                if (u < a.stages && u < nchunks) issue(c_lo + u, sg);
                if (++sg == a.stages) sg = 0;
            }
        }

#ifdef WWZ_TRACE
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tr1));
#endif
        double W[M], Q[M], Ws[M], Wc[M], Wy[M], Wys[M], Wyc[M];
#pragma unroll
        for (int r = 0; r < M; ++r) W[r] = Q[r] = Ws[r] = Wc[r] = Wy[r] = Wys[r] = Wyc[r] = 0.0;

        const double step16 = 2.0 * kap * dtau;
        const int row_off = P + r_loc * a.row_stride2;
        for (int i = 0; i < nchunks; ++i) {
            mbar_wait(&full[s], (ph >> s) & 1u);
            ph ^= 1u << s;
#ifdef WWZ_TRACE
            if (i == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tr_first));
            clk_mark = clock64();
#endif
            const double2 *const st = stage_base + (size_t)s * a.stage_stride2;
            const double2 *const row = st + row_off;
            const int cnt = (int)min((long long)P, p1 - p0 - (long long)(c_lo + i) * P);
            // software pipeline: the weights of point p+1 are evaluated under the accumulations of point p
            double2 t_y = st[0];
            double2 sc = row[0], ysc = row[1];
            if constexpr (REC) {
                static_assert(M == 8, "the power tree below is written for 8 cells per lane");
                // rho advanced from point to point by the tabulated growth E (one multiplication instead of a second
                // exp2 chain), from an exact value at the first point of the chunk -- as long as no lane's rho leaves
                // the range of a double inside the chunk (the exponent a D/8 at both ends of the chunk below 900 in
                // magnitude; t is sorted, so it is monotone in between); otherwise, and on unsorted axes, the exact pair
                bool fast = false;
                if (use_e && sorted) {
                    const double b0 = kap * (st[0].x - tau_r[0]) * step16, b1 = kap * (st[cnt - 1].x - tau_r[0]) * step16;
                    fast = __all_sync(0xffffffffu, fabs(b0) < 14400.0 && fabs(b1) < 14400.0);
                }
                if (fast) {
                    const double *const erow = reinterpret_cast<const double *>(st + a.e_off2) + r_loc;
                    double rho_c = exp2_sixteenth(kap * (st[0].x - tau_r[0]) * step16, tbl);
#pragma unroll 2
                    for (int p = 0; p < cnt; ++p) {
                        const double2 ty = st[p];
                        const double2 sc_ = row[2 * p];
                        const double2 ysc_ = row[2 * p + 1];
                        const double grow = erow[p * TR];
                        bool tiny;
                        const double w0 = gauss_weight_flag(kap * (ty.x - tau_r[0]), tbl, tiny);
                        const double rho = tiny ? 0.0 : rho_c;
                        rho_c *= grow;
                        double wv[M];
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
                            Ws[r] = fma(w, sc_.x, Ws[r]);
                            Wc[r] = fma(w, sc_.y, Wc[r]);
                            Wy[r] = fma(w, ty.y, Wy[r]);
                            Wys[r] = fma(w, ysc_.x, Wys[r]);
                            Wyc[r] = fma(w, ysc_.y, Wyc[r]);
                        }
                    }
                } else {
                    // two points per iteration: their exp2 pairs interleave (four independent chains); w_0 rho^m
                    // through a depth-3 product tree
#pragma unroll 2
                    for (int p = 0; p < cnt; ++p) {
                        const double2 ty = st[p];
                        const double2 sc_ = row[2 * p];
                        const double2 ysc_ = row[2 * p + 1];
                        double w0, rho;
                        gauss_weight_ratio(kap * (ty.x - tau_r[0]), step16, tbl, w0, rho);
                        double wv[M];
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
                            Ws[r] = fma(w, sc_.x, Ws[r]);
                            Wc[r] = fma(w, sc_.y, Wc[r]);
                            Wy[r] = fma(w, ty.y, Wy[r]);
                            Wys[r] = fma(w, ysc_.x, Wys[r]);
                            Wyc[r] = fma(w, ysc_.y, Wyc[r]);
                        }
                    }
                }
            } else {
                double wv[M];
#pragma unroll
                for (int r = 0; r < M; ++r) wv[r] = gauss_weight(kap * (t_y.x - tau_r[r]), tbl);
#pragma unroll 1
                for (int p = 0; p < cnt; ++p) {
                    const int pn = min(p + 1, cnt - 1);
                    const double2 t_yn = st[pn];
                    const double2 scn = row[2 * pn];
                    const double2 yscn = row[2 * pn + 1];
                    double wn[M];
#pragma unroll
                    for (int r = 0; r < M; ++r) wn[r] = gauss_weight(kap * (t_yn.x - tau_r[r]), tbl);
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
                        wv[r] = wn[r];
                    }
                    t_y = t_yn;
                    sc = scn;
                    ysc = yscn;
                }
            }
#ifdef WWZ_TRACE
            clk_pts += clock64() - clk_mark;
#endif
            __syncwarp();
            if (i + a.stages < nchunks) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(c_lo + i + a.stages, s);
            }
            if (++s == a.stages) s = 0;
        }

#ifdef WWZ_TRACE
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tr_loop));
#endif
        if (nchunks > 0) {
            const long long first = (long long)c_lo * P, last = min(p1 - p0, (long long)(c_lo + nchunks) * P);
            walked += (unsigned long long)(last - first);
        }
        if (ok) {
            const long long kcol = (long long)a.k0 + kb;
            const double D = kap * dtau;
            const size_t nc = (size_t)a.ncell_total;
#pragma unroll
            for (int r = 0; r < M; ++r) {
                const int j = g * M + r;
                if (j < a.nt) {
                    const size_t cell = (size_t)seg * (size_t)kNumSums * nc + (size_t)j * (size_t)a.nf_total + (size_t)kcol;
                    // REC: undo the per-cell scale h_r = 2^(-(r D)^2/16) that was factored out of the weights
                    const double h = REC ? exp2(-(double)(r * r) * D * D * 0.0625) : 1.0;
                    m_sums[0 * nc + cell] = W[r] * h;
                    m_sums[1 * nc + cell] = Q[r] * h * h;
                    m_sums[2 * nc + cell] = Ws[r] * h;
                    m_sums[3 * nc + cell] = Wc[r] * h;
                    m_sums[4 * nc + cell] = Wy[r] * h;
                    m_sums[5 * nc + cell] = Wys[r] * h;
                    m_sums[6 * nc + cell] = Wyc[r] * h;
                }
            }
        }
        if constexpr (FUSE) {
            // the last segment of the tile to finish adds the slabs in order (the same order as finalize_kernel) and
            // writes the tile's cells; rows of the tile that belong to the direct flavour (!ok) are left to finalize_kernel
            bool last = true;
            if (a.segs > 1) {
                __threadfence();
                __syncwarp();
                unsigned int done = 0;
                unsigned int *const cnt = a.tile_cnt + (size_t)mem * (size_t)a.ntiles + (size_t)tile;
                if (lane == 0) {
                    done = atomicAdd(cnt, 1u);
                    if (done == (unsigned)a.segs - 1) *cnt = 0;
                }
                done = __shfl_sync(0xffffffffu, done, 0);
                last = done == (unsigned)a.segs - 1;
                if (last) __threadfence();
            }
            if (last && ok) {
                const long long kcol = (long long)a.k0 + kb;
                const size_t nc = (size_t)a.ncell_total;
                const double om = a.omega[(long long)mem * a.m_omega + kcol];
                const long long oo = (long long)mem * a.nt * a.row_stride;
#pragma unroll 1
                for (int r = 0; r < M; ++r) {
                    const int j = g * M + r;
                    if (j >= a.nt) break;
                    const size_t idx = (size_t)j * (size_t)a.nf_total + (size_t)kcol;
                    double acc[kNumSums];
#pragma unroll
                    for (int x = 0; x < kNumSums; ++x) acc[x] = __ldcg(&m_sums[(size_t)x * nc + idx]);
                    for (int sg = 1; sg < a.segs; ++sg) {
                        const double *slab = m_sums + (size_t)sg * kNumSums * nc;
#pragma unroll
                        for (int x = 0; x < kNumSums; ++x) acc[x] += __ldcg(&slab[(size_t)x * nc + idx]);
                    }
                    if (a.fix_list && (acc[0] < 0x1p-200 || acc[1] < 0x1p-400)) {
                        const unsigned int slot = atomicAdd(a.fix_count, 1u);
                        if (slot < a.fix_cap) a.fix_list[slot] = (unsigned int)((size_t)mem * nc + idx);
                    }
                    double neff, c0, c1, c2;
                    cell_from_sums(acc, om, m_taup[j], a.thr, neff, c0, c1, c2);
                    store_cell(oo + (long long)j * a.row_stride + kcol, neff, c0, c1, c2, a.o_neffs, a.o_a0, a.o_a1, a.o_a2,
                               a.o_wwa, a.o_phase);
                }
            }
            report();
        }
#ifdef WWZ_TRACE
        if (a.trace && lane == 0) {
            unsigned long long tr2;
            unsigned int smid;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tr2));
            asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
            const long long pts = nchunks > 0 ? min(p1 - p0, (long long)(c_lo + nchunks) * P) - (long long)c_lo * P : 0;
            unsigned long long *const tr = a.trace + 8 * (size_t)item;
            tr[0] = tr0;                               // item taken from the queue
            tr[1] = tr1;                               // window found, first copies issued
            tr[2] = tr2;                               // partial sums (and, fused, cells) written
            tr[3] = (unsigned long long)pts;           // points walked
            tr[4] = (unsigned long long)clk_pts;       // SM cycles inside the point loops
            tr[5] = smid;
            tr[6] = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
            tr[7] = ((tr_first ? tr_first - tr1 : 0) << 32) | (tr2 - tr_loop);   // wait for the first chunk | epilogue (ns)
        }
#endif
    }

    if (lane == 0) {
        if (a.walked && walked) atomicAdd(a.walked, walked);
        // the last warp of the launch to leave re-arms the queue for the next launch on this stream
        const unsigned int left = atomicAdd(&a.queue[1], 1u);
        if (left == gridDim.x * kRecWarps - 1) {
            a.queue[0] = 0;
            a.queue[1] = 0;
            __threadfence();
        }
    }
}

// Tile shape: 8 tau-groups x 4 rows when the tau axis is long; on small tau grids (the reference default is 50)
// tau-compact tiles -- one or two tau-groups x 32 / 16 rows -- so that the point windows prune.
RecPlan plan_rec(int64_t n, int64_t nt, int64_t nf, int64_t seg_len, int cells_per_lane, bool rho_growth)
{
    RecPlan p{};
    const int M = cells_per_lane;
    p.cells_per_lane = M;
    p.groups = (int)((nt + M - 1) / M);
    int tg_log2 = 3;                                    // 8 tau-groups x 4 rows per warp tile
    if (M == kRecCells) {
        if (p.groups < 8) tg_log2 = 2;                  // nt <= 56: 4 tau-groups x 8 rows (32 rows: chunks of 4 points, slower)
        const char *e = getenv("WWZB200_REC_TG_LOG2");
        if (e && *e) tg_log2 = std::max(0, std::min(5, atoi(e)));
    } else {
        tg_log2 = 2;                                    // direct flavour: 4 tau-groups (8 taus) x 8 rows
        const char *e = getenv("WWZB200_TILE_TG_LOG2");
        if (e && *e) tg_log2 = std::max(0, std::min(5, atoi(e)));
    }
    while (tg_log2 > 0 && (1 << (tg_log2 - 1)) >= p.groups) --tg_log2;   // no wider than the tau axis
    p.tg_log2 = tg_log2;
    const int TR = 32 >> tg_log2;
    // ring budget per warp: ~14 KB at 12 warps per SM (recurrence flavour, 3 CTAs), ~12 KB at 16 (direct, 4 CTAs)
    const int64_t budget = M == kRecCells ? 14336 : 12288;
    // Recurrence flavour with the tabulated growth of rho (single transforms on long tau axes: long windows): room for
    // the [P][TR] growth block in every stage, and two stages of longer chunks -- a chunk takes ~10 us of arithmetic
    // against ~1 us for its bulk copies, and every chunk costs ~1000 cycles of hand-over.  Member batches and small
    // grids (windows of two or three chunks: all of them in flight at once) keep three stages and no table.
    p.growth = (M == kRecCells && rho_growth) ? 1 : 0;
    {
        const char *e = getenv("WWZB200_NO_RHO_GROWTH");
        if (e && *e && atoi(e)) p.growth = 0;
    }
    const int64_t e_per_pt = p.growth ? 8 * TR : 0;
    p.stages = p.growth ? 2 : kStages;
    int64_t P = (budget / p.stages - 16 * TR) / (16 + 32 * TR + e_per_pt);
    P &= ~(int64_t)1;
    const int64_t span = seg_len > 0 ? std::min(seg_len, n) : n;
    if (P > span) P = ((span + 1) / 2) * 2;
    if (P < 2) P = 2;
    p.points = (int)P;
    p.n_pad = ((n + P + 15) / 16) * 16;
    p.row_stride = P * 32 + 16;                         // an odd number of 16 B units: rows never share a bank group
    p.stage_bytes = (int)(P * 16 + (int64_t)TR * p.row_stride + P * e_per_pt);
    p.warp_bytes = ((p.stages * p.stage_bytes + kStages * 8 + 127) / 128) * 128;   // ring, then the barriers
    // CTA shape: 4 independent warps; 168 registers -> 3 CTAs = 12 warps per SM for the recurrence flavour (3 warps per
    // SMSP is all the register file holds at 168; budgets between 168 and 255 all give 2), 128 -> 16 for the direct one
    p.warps = 4;
    p.nreg = M == kRecCells ? 168 : 128;
    p.smem_bytes = p.warps * p.warp_bytes;
    return p;
}

template <int M, bool REC, int NREG, bool FUSE>
static cudaError_t launch_tiles_m(const RecPlan &plan, RecArgs &a, int sms, cudaStream_t st)
{
    const int WARPS = plan.warps;
    static SmemOptIn done;
    cudaError_t e = ensure_dynamic_smem(wwz_tile_kernel<M, REC, NREG, FUSE>, plan.smem_bytes, done);
    if (e != cudaSuccess) return e;
    // resident CTAs per SM for this shared-memory size (cached per device: the query is a host-side computation)
    static int cached_smem[kMaxDevices], cached_ctas[kMaxDevices], cached_warps[kMaxDevices];
    int dev = 0;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    const bool tracked = dev >= 0 && dev < kMaxDevices;
    if (tracked && cached_ctas[dev] > 0 && cached_smem[dev] == plan.smem_bytes && cached_warps[dev] == WARPS) {
        per_sm = cached_ctas[dev];
    } else {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wwz_tile_kernel<M, REC, NREG, FUSE>, WARPS * 32,
                                                          (size_t)plan.smem_bytes);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        if (tracked) {
            cached_smem[dev] = plan.smem_bytes;
            cached_warps[dev] = WARPS;
            cached_ctas[dev] = per_sm;
        }
    }
    const int64_t want = (a.nitems + WARPS - 1) / WARPS;
    const int64_t ctas = std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sms * per_sm));
    wwz_tile_kernel<M, REC, NREG, FUSE><<<(unsigned)ctas, WARPS * 32, plan.smem_bytes, st>>>(a);
    bump_launch_counter();
    return cudaGetLastError();
}

int64_t rec_tile_count(const RecPlan &plan, int64_t nf_batch, int64_t members)
{
    const int TG = 1 << plan.tg_log2, TR = 32 >> plan.tg_log2;
    return (int64_t)((plan.groups + TG - 1) / TG) * ((nf_batch + TR - 1) / TR) * members;
}
int64_t rec_estep_count(const RecPlan &plan, int64_t nf_batch, int64_t members)
{
    const int TR = 32 >> plan.tg_log2;
    return ((nf_batch + TR - 1) / TR) * plan.n_pad * TR * members;
}
int rec_tau_tiles(const RecPlan &plan) { return (plan.groups + (1 << plan.tg_log2) - 1) >> plan.tg_log2; }
int rec_tile_taus(const RecPlan &plan) { return plan.cells_per_lane << plan.tg_log2; }

cudaError_t launch_rec(const RecPlan &plan, const double2 *d_ty, const double2 *d_basis, const double *d_ts, int64_t n,
                       const double *d_tau, int64_t nt, const double *d_kappa, const double *d_dcut,
                       const double *d_taugap, const int *d_unsorted, int64_t k0, int64_t nf_batch, int64_t nf_total, int segs,
                       int64_t seg_len, const unsigned char *d_rec_row, const double *d_grid, double *d_sums,
                       unsigned int *d_queue, unsigned long long *d_walked, int sms, cudaStream_t st, int64_t members,
                       int64_t m_ts, int64_t m_tau, const RecFuse *fuse, double *d_estep)
{
    if (nf_batch <= 0 || nt <= 0 || members <= 0) return cudaSuccess;
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
    a.queue = d_queue;
    a.walked = d_walked;
    a.n = n;
    a.n_pad = plan.n_pad;
    a.ncell_total = nt * nf_total;
    a.seg_len = seg_len;
    a.segs = segs;
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
    a.members = (int)members;
    a.m_ts = m_ts;
    a.m_tau = m_tau;
    a.nitems = a.ntiles * segs * members;
    if (a.nitems >= ((int64_t)1 << 32)) return cudaErrorInvalidValue;
    a.row_stride2 = (int)(plan.row_stride / 16);
    a.stage_stride2 = plan.stage_bytes / 16;
    a.stages = plan.stages;
    a.e_off2 = (int)((plan.points * 16 + (int64_t)TR * plan.row_stride) / 16);
    a.estep = nullptr;
    if (plan.cells_per_lane == kRecCells && plan.growth && d_estep) {
        const int64_t cnt = ((nf_batch + TR - 1) / TR) * plan.n_pad * TR * members;
        rho_growth_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(a, d_estep);
        bump_launch_counter();
        cudaError_t ee = cudaGetLastError();
        if (ee != cudaSuccess) return ee;
        a.estep = d_estep;
    }
    a.warp_stride = plan.warp_bytes;
#ifdef WWZ_TRACE
    {
        // developer build: the previous traced launch is appended to $WWZB200_TRACE (binary) before the next one starts
        static unsigned long long *buf = nullptr;
        static int64_t pending = 0;
        const char *path = getenv("WWZB200_TRACE");
        a.trace = nullptr;
        if (path && *path) {
            if (!buf) cudaMallocManaged(&buf, (size_t)64 << 20);
            if (pending) {
                cudaDeviceSynchronize();
                FILE *f = fopen(path, "ab");
                if (f) {
                    fwrite(&pending, 8, 1, f);
                    fwrite(buf, 64, (size_t)pending, f);
                    fclose(f);
                }
                pending = 0;
            }
            if (plan.cells_per_lane == kRecCells && a.nitems * 64 <= ((int64_t)64 << 20)) {
                a.trace = buf;
                pending = a.nitems;
            }
        }
    }
#endif
    if (plan.cells_per_lane == kRecCells) {
        if (fuse) {
            a.omega = fuse->omega;
            a.m_omega = fuse->m_omega;
            a.row_stride = fuse->row_stride;
            a.thr = fuse->thr;
            a.o_neffs = fuse->neffs;
            a.o_a0 = fuse->a0;
            a.o_a1 = fuse->a1;
            a.o_a2 = fuse->a2;
            a.o_wwa = fuse->wwa;
            a.o_phase = fuse->phase;
            a.tile_cnt = fuse->tile_cnt;
            a.fix_count = fuse->fix_count;
            a.fix_list = fuse->fix_list;
            a.fix_cap = fuse->fix_cap;
            a.unfused = fuse->unfused;
            a.blk_cnt = fuse->blk_cnt;
            a.blk_flag = fuse->blk_flag;
            a.blk_epoch = fuse->blk_epoch;
            a.blk_gt = fuse->blk_gt;      // 0: no progress reports, frequency-major tile order
            return launch_tiles_m<kRecCells, true, 168, true>(plan, a, sms, st);
        }
        return launch_tiles_m<kRecCells, true, 168, false>(plan, a, sms, st);
    }
    return launch_tiles_m<kTileDirectCells, false, 128, false>(plan, a, sms, st);
}

}  // namespace wwz
