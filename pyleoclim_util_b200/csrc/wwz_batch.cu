// Batched kernels: AR(1) surrogate generation, per-series standardisation and re-tiling, the
// shared-time-axis transform (one basis evaluation contracted against S series), per-cell quantiles.
//
// Reference semantics
//   ar1_model / ar1_sim (uneven branch)      pyleoclim/utils/tsmodel.py:59-69, 159-165
//   SurrogateSeries.from_series              pyleoclim/core/surrogateseries.py:137-139, 170-174
//   standardize                              pyleoclim/utils/tsutils.py:1134-1154
//   MultipleScalogram.quantiles              pyleoclim/core/scalograms.py:635-641  (scipy mquantiles)
//   MultiplePSD.quantiles                    pyleoclim/core/psds.py:913
// The weights, Neff and the basis sums of a cell do not depend on y (utils/wavelet.py:988-1019);
// only a0, a1, a2 do, and linearly (:1021-1023).  All surrogates of one target share its time axis,
// so the S-surrogate significance test is one weight evaluation per (cell, point) contracted
// against the S series.
#include <math_constants.h>

#include <algorithm>
#include <cmath>

#include "wwz_device.cuh"
#include "wwz_internal.h"

namespace wwz {

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al., SC'11) + Box-Muller.
// Stream: counter = (pair index lo, hi, series lo, hi), key = (seed lo, hi); the two normals of a
// pair serve steps 2p and 2p+1 of the recurrence.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double u01(uint32_t a, uint32_t b)
{
    // 53 random bits -> (0, 1): ((a>>5)*2^26 + (b>>6) + 0.5) * 2^-53
    return (((double)(a >> 5)) * 67108864.0 + (double)(b >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void normal_pair(uint64_t pair, uint64_t series, uint64_t seed, double &z0, double &z1)
{
    uint32_t r[4];
    philox4x32_10((uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)series, (uint32_t)(series >> 32), (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    const double u1 = u01(r[0], r[1]), u2 = u01(r[2], r[3]);
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z0 = rad * c;
    z1 = rad * s;
}

// rho_i = exp(-(t_i - t_{i-1})/tau), q_i = sqrt(1 - rho_i^2)*sigma   (utils/tsmodel.py:64-66)
__global__ void ar1_coeff_kernel(const double *__restrict__ ts, int64_t n, double tau, double sigma,
                                 double *__restrict__ rho, double *__restrict__ q)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) {
        rho[0] = 0.0;
        q[0] = 0.0;
        return;
    }
    const double scaled_dt = (ts[i] - ts[i - 1]) / tau;
    const double r = exp(-scaled_dt);
    rho[i] = r;
    q[i] = sqrt(1.0 - r * r) * sigma;
}

// One warp = 32 surrogates; the recurrence runs 32 steps at a time and the 32x32 tile is written
// through shared memory so that each surrogate row receives 256 contiguous bytes.
// y[0] = 0 (:61), y[i] = y[i-1]*rho_i + q_i*z_i (:67); stored value is y + mu
// (core/surrogateseries.py:139).
__global__ void __launch_bounds__(128)
    ar1_generate_kernel(const double *__restrict__ rho, const double *__restrict__ q, int64_t n, int64_t S,
                        double mu, uint64_t seed, double *__restrict__ Y)
{
    __shared__ double tile[4][32][33];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t s0 = ((int64_t)blockIdx.x * 4 + wib) * 32;
    if (s0 >= S) return;
    const int64_t s = s0 + lane;
    double y = 0.0, zodd = 0.0;
    for (int64_t i0 = 0; i0 < n; i0 += 32) {
        const int cnt = (int)min((int64_t)32, n - i0);
        for (int j = 0; j < cnt; ++j) {
            const int64_t i = i0 + j;
            if (i > 0) {
                double z;
                if ((i & 1) == 0 || i == 1) {
                    double z0, z1;
                    normal_pair((uint64_t)(i >> 1), (uint64_t)s, seed, z0, z1);
                    zodd = z1;
                    z = (i & 1) ? z1 : z0;
                } else {
                    z = zodd;
                }
                y = y * rho[i] + q[i] * z;
            }
            tile[wib][lane][j] = y + mu;
        }
        __syncwarp();
        for (int r = 0; r < 32; ++r) {
            const int64_t sr = s0 + r;
            if (sr < S && lane < cnt) Y[sr * n + i0 + lane] = tile[wib][r][lane];
        }
        __syncwarp();
    }
}

cudaError_t launch_ar1(const double *d_ts, int64_t n, double tau, double sigma, double mu, int64_t S, uint64_t seed,
                       double *d_rho, double *d_q, double *d_Y, cudaStream_t st)
{
    if (n <= 0 || S <= 0) return cudaSuccess;
    ar1_coeff_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_ts, n, tau, sigma, d_rho, d_q);
    ar1_generate_kernel<<<(unsigned)((S + 127) / 128), 128, 0, st>>>(d_rho, d_q, n, S, mu, seed, d_Y);
    bump_launch_counter(2);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Per-series mean / std (ddof = 0), one CTA per series (utils/tsutils.py:1134-1135), and the
// standardised, re-tiled copy  Yt[S/NS][n_pad][NS]  that the shared-axis kernel stages with one
// bulk copy per chunk.  A (nearly) constant series (|std| < 1e-3) is left unscaled (:1140-1146).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double *sh)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(256)
    row_stats_kernel(const double *__restrict__ Y, int64_t n, int standardize, double *__restrict__ mu_out,
                     double *__restrict__ sig_out)
{
    __shared__ double sh[8];
    const int64_t s = blockIdx.x;
    const double *row = Y + s * n;
    double mu = 0.0, sig = 1.0;
    if (standardize) {
        double a = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) a += row[i];
        const double mean = block_sum(a, sh) / (double)n;
        double v = 0.0;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
            const double d = row[i] - mean;
            v += d * d;
        }
        const double sd = sqrt(block_sum(v, sh) / (double)n);
        if (fabs(sd) < 1e-3) {
            mu = 0.0;
            sig = 1.0;
        } else {
            mu = mean;
            sig = sd;
        }
    }
    if (threadIdx.x == 0) {
        mu_out[s] = mu;
        sig_out[s] = sig;
    }
}

template <int NS>
__global__ void retile_kernel(const double *__restrict__ Y, int64_t n, int64_t n_pad, int64_t S,
                              const double *__restrict__ mu, const double *__restrict__ sig, double *__restrict__ Yt)
{
    // thread -> (block b, point i, lane j): Yt[(b*n_pad + i)*NS + j] = (Y[(b*NS+j)*n + i] - mu)/sig
    __shared__ double tile[NS][33];
    const int64_t b = blockIdx.y;
    const int64_t i0 = (int64_t)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // blockDim = 32 x NS
    const int64_t s = b * NS + ty;
    const int64_t i = i0 + tx;
    double v = 0.0;
    if (s < S && i < n) v = (Y[s * n + i] - mu[s]) / sig[s];
    tile[ty][tx] = v;
    __syncthreads();
    // write: NS contiguous doubles per point
    for (int idx = threadIdx.x; idx < 32 * NS; idx += blockDim.x) {
        const int p = idx / NS, j = idx - p * NS;
        const int64_t ip = i0 + p;
        if (ip < n_pad) Yt[(b * n_pad + ip) * NS + j] = tile[j][p];
    }
}

cudaError_t launch_standardize_retile(const double *d_Y, int64_t n, int64_t n_pad, int64_t S, bool standardize,
                                      double *d_mu, double *d_sig, double *d_Yt, cudaStream_t st)
{
    if (S <= 0 || n <= 0) return cudaSuccess;
    row_stats_kernel<<<(unsigned)S, 256, 0, st>>>(d_Y, n, standardize ? 1 : 0, d_mu, d_sig);
    const int64_t blocks = (S + kSurrBlock - 1) / kSurrBlock;
    dim3 grid((unsigned)((n_pad + 31) / 32), (unsigned)blocks);
    retile_kernel<kSurrBlock><<<grid, 32 * kSurrBlock, 0, st>>>(d_Y, n, n_pad, S, d_mu, d_sig, d_Yt);
    bump_launch_counter(2);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Shared-time-axis transform.  One lane = one cell (as in wwz_cells_kernel with M = 1); a CTA owns
// 256 consecutive cells and a block of NS = 16 series.  Per (cell, point) the weight is evaluated
// once (12 FP64 instructions + 6 for W, Q, Ws, Wc and the products w*s, w*c) and contracted
// against the 16 series values of that point (3 DFMA each, broadcast LDS from the staged
// Yt tile): 66 FP64 instructions per (cell, point, 16 series) = 4.1 per (cell, point, series)
// against the 19 of a full transform per series.
// Staging: TMA producer warp as in wwz_cells_kernel; per chunk one bulk copy of (t,y) pairs, one per
// frequency row of {s, c} records, one for the [P][NS] tile of series values.
// Epilogue in registers: Neff / mask once per cell, then per series the coefficients through the
// pre-computed rotation (sin, cos)(omega*tau); outputs are written cell-major [ncell][S] (each lane
// stores 128 contiguous bytes), the layout the quantile kernel consumes.
// ------------------------------------------------------------------------------------------------
struct SharedArgs {
    const double2 *ty;
    const double2 *basis;   // [nf][n_pad] records {s, c}
    const double *Yt;       // [S/NS][n_pad][NS]
    const double *tau;
    const double *kappa;
    const double2 *rot;     // [ncell] (sin, cos)(omega*tau)
    double *neffs;          // [ncell] (written by surrogate block 0)
    double *wq;             // [2][ncell] W, Q (written by surrogate block 0; deep-underflow detection)
    double *amp, *phase, *a0, *a1, *a2;   // [ncell][S_pad], any may be null except amp
    long long n, n_pad, ncell, S_pad;
    int nt, nf, P, nchunks;
    int row_stride2, stage_stride2, y_off2;   // double2 units
    double thr;
};

template <int NS>
__global__ void __launch_bounds__(kCellThreads, 1) wwz_shared_kernel(const SharedArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *const stage_base = reinterpret_cast<double2 *>(smem_raw);
    unsigned char *const tail = smem_raw + (size_t)kStages * (size_t)a.stage_stride2 * sizeof(double2);
    uint64_t *const full = reinterpret_cast<uint64_t *>(tail);
    uint64_t *const empty = full + kStages;
    double *const tbl = reinterpret_cast<double *>(empty + kStages);   // [16] 2^(j/16)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool consumer = warp < kConsumerWarps;
    const long long q0 = (long long)blockIdx.x * kConsumerThreads;
    const long long qlast = min(q0 + (long long)kConsumerThreads - 1, a.ncell - 1);
    // cells are numbered frequency-major here: q = k*nt + j
    const int k_first = (int)(q0 / a.nt), k_last = (int)(qlast / a.nt);
    const int Ft = k_last - k_first + 1;
    const int P = a.P;
    const long long sb = blockIdx.y;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        mbar_fence_init();
    }
    load_exp2_table(tbl, tid);
    __syncthreads();

    if (!consumer) {
        const uint32_t bytes = (uint32_t)P * 16u + (uint32_t)Ft * (uint32_t)P * 16u + (uint32_t)P * NS * 8u;
        int s = 0, use = 0;
        for (int c = 0; c < a.nchunks; ++c) {
            if (use > 0) mbar_wait(&empty[s], (uint32_t)(use - 1) & 1u);
            double2 *const st = stage_base + (size_t)s * a.stage_stride2;
            if (lane == 0) mbar_arrive_expect_tx(&full[s], bytes);
            __syncwarp();
            for (int r = lane; r <= Ft + 1; r += 32) {
                if (r == 0) {
                    tma_load_1d(st, a.ty + (size_t)c * P, (uint32_t)P * 16u, &full[s]);
                } else if (r == 1) {
                    tma_load_1d(st + a.y_off2, a.Yt + ((size_t)sb * a.n_pad + (size_t)c * P) * NS,
                                (uint32_t)P * NS * 8u, &full[s]);
                } else {
                    const size_t row = (size_t)(k_first + r - 2);
                    tma_load_1d(st + P + (size_t)(r - 2) * a.row_stride2, a.basis + row * (size_t)a.n_pad + (size_t)c * P,
                                (uint32_t)P * 16u, &full[s]);
                }
            }
            if (++s == kStages) { s = 0; ++use; }
        }
        return;
    }

    const long long q = q0 + tid;
    const bool ok = q < a.ncell;
    int k = k_first, j = 0;
    if (ok) {
        k = (int)(q / a.nt);
        j = (int)(q - (long long)k * a.nt);
    }
    const double kap = a.kappa[k];
    const double tj = a.tau[j];
    const int row_off = P + (k - k_first) * a.row_stride2;

    double W = 0.0, Q = 0.0, Ws = 0.0, Wc = 0.0;
    double Wy[NS], Wys[NS], Wyc[NS];
#pragma unroll
    for (int m = 0; m < NS; ++m) Wy[m] = Wys[m] = Wyc[m] = 0.0;

    {
        int s = 0, use = 0;
        for (int c = 0; c < a.nchunks; ++c) {
            mbar_wait(&full[s], (uint32_t)use & 1u);
            const double2 *const st = stage_base + (size_t)s * a.stage_stride2;
            const double2 *const row = st + row_off;
            const double2 *const yv = st + a.y_off2;
            const int cnt = (int)min((long long)P, a.n - (long long)c * P);
            for (int p = 0; p < cnt; ++p) {
                const double t = st[p].x;
                const double2 sc = row[p];
                const double w = gauss_weight(kap * (t - tj), tbl);
                const double ws = w * sc.x, wc = w * sc.y;
                W += w;
                Q = fma(w, w, Q);
                Ws += ws;
                Wc += wc;
#pragma unroll
                for (int m = 0; m < NS / 2; ++m) {
                    const double2 y2 = yv[p * (NS / 2) + m];
                    Wy[2 * m] = fma(w, y2.x, Wy[2 * m]);
                    Wys[2 * m] = fma(ws, y2.x, Wys[2 * m]);
                    Wyc[2 * m] = fma(wc, y2.x, Wyc[2 * m]);
                    Wy[2 * m + 1] = fma(w, y2.y, Wy[2 * m + 1]);
                    Wys[2 * m + 1] = fma(ws, y2.y, Wys[2 * m + 1]);
                    Wyc[2 * m + 1] = fma(wc, y2.y, Wyc[2 * m + 1]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == kStages) { s = 0; ++use; }
        }
    }

    if (!ok) return;
    const size_t cell = (size_t)j * (size_t)a.nf + (size_t)k;   // output order: tau-major rows
    const double neff = W * W / Q;
    if (sb == 0) {
        a.neffs[cell] = neff;
        a.wq[cell] = W;
        a.wq[(size_t)a.ncell + cell] = Q;
    }
    const bool valid = neff > a.thr;
    const double S1 = Ws / W, C1 = Wc / W;
    const double2 rt = a.rot[cell];
    const size_t o = cell * (size_t)a.S_pad + (size_t)sb * NS;
#pragma unroll
    for (int m = 0; m < NS; ++m) {
        double c0 = CUDART_NAN, c1 = CUDART_NAN, c2 = CUDART_NAN;
        if (valid) {
            const double Yb = Wy[m] / W;
            const double cyc = Wyc[m] / W - Yb * C1;
            const double cys = Wys[m] / W - Yb * S1;
            c0 = Yb;
            c1 = 2.0 * (rt.y * cyc + rt.x * cys);
            c2 = 2.0 * (rt.y * cys - rt.x * cyc);
        }
        a.amp[o + m] = sqrt(c1 * c1 + c2 * c2);
        if (a.phase) a.phase[o + m] = atan2(c2, c1);
        if (a.a0) a.a0[o + m] = c0;
        if (a.a1) a.a1[o + m] = c1;
        if (a.a2) a.a2[o + m] = c2;
    }
}

// (sin, cos)(omega_k * tau_j) per cell, same residual folding as finalize_kernel.
__global__ void rotation_kernel(const double *__restrict__ tau, const double *__restrict__ omega, int64_t nt,
                                int64_t nf, double2 *__restrict__ rot)
{
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= nt * nf) return;
    const int64_t j = idx / nf, k = idx - j * nf;
    const double om = omega[k], tj = tau[j];
    const double th = om * tj;
    const double eps = fma(om, tj, -th);
    double s0, c0;
    sincos(th, &s0, &c0);
    rot[idx] = make_double2(fma(eps, c0, s0), fma(-eps, s0, c0));
}

// {sin, cos}(omega_k t_i) records for the shared-axis kernel (y does not enter the basis there).
__global__ void basis_sc_kernel(const double *__restrict__ ts, int64_t n, int64_t n_pad,
                                const double *__restrict__ omega, double2 *__restrict__ ty, double2 *__restrict__ basis)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    const int64_t k = blockIdx.y;
    const double om = omega[k];
    double t = 0.0, s = 0.0, c = 0.0;
    if (i < n) {
        t = ts[i];
        if (!isnan(om)) {
            const double th = om * t;
            const double eps = fma(om, t, -th);
            double s0, c0;
            sincos(th, &s0, &c0);
            s = fma(eps, c0, s0);
            c = fma(-eps, s0, c0);
        }
    }
    basis[(size_t)k * (size_t)n_pad + (size_t)i] = make_double2(s, c);
    if (k == 0) ty[i] = make_double2(t, 0.0);
}

cudaError_t launch_basis_sc(const double *d_ts, int64_t n, int64_t n_pad, const double *d_omega, int64_t nf,
                            double2 *d_ty, double2 *d_basis, cudaStream_t st)
{
    if (nf <= 0 || n_pad <= 0) return cudaSuccess;
    dim3 grid((unsigned)((n_pad + 255) / 256), (unsigned)nf);
    basis_sc_kernel<<<grid, 256, 0, st>>>(d_ts, n, n_pad, d_omega, d_ty, d_basis);
    bump_launch_counter();
    return cudaGetLastError();
}

// Deep-underflow cells of a batch: flagged from the shared W, Q, then every (cell, series) pair is
// recomputed by one warp with the reference's two-pass expressions.
__global__ void flag_cells_kernel(const double *__restrict__ wq, int64_t ncell, unsigned int *__restrict__ count,
                                  unsigned int *__restrict__ list)
{
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= ncell) return;
    const double W = wq[idx], Q = wq[ncell + idx];
    if (W < 0x1p-200 || Q < 0x1p-400) list[atomicAdd(count, 1u)] = (unsigned int)idx;
}

template <int NS>
__global__ void __launch_bounds__(128)
    faithful_batch_kernel(const double *__restrict__ ts, const double *__restrict__ Yt, int64_t n, int64_t n_pad,
                          int64_t S, int64_t S_pad, const double *__restrict__ tau, const double *__restrict__ omega,
                          int64_t nf, double c, double thr, const unsigned int *__restrict__ count,
                          const unsigned int *__restrict__ list, double *__restrict__ neffs, double *__restrict__ amp,
                          double *__restrict__ phase, double *__restrict__ a0o, double *__restrict__ a1o,
                          double *__restrict__ a2o)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = (int64_t)(*count) * S;
    for (int64_t item = warp0; item < total; item += nwarps) {
        const int64_t cell = list[item / S], s = item % S;
        const int64_t j = cell / nf, k = cell - j * nf;
        const double *ys = Yt + ((s / NS) * n_pad) * NS + (s % NS);
        double neff, a0, a1, a2;
        faithful_cell_warp(ts, ys, NS, n, tau[j], omega[k], c, thr, lane, neff, a0, a1, a2);
        if (lane == 0) {
            const size_t o = (size_t)cell * (size_t)S_pad + (size_t)s;
            if (s == 0) neffs[cell] = neff;
            amp[o] = sqrt(a1 * a1 + a2 * a2);
            if (phase) phase[o] = atan2(a2, a1);
            if (a0o) a0o[o] = a0;
            if (a1o) a1o[o] = a1;
            if (a2o) a2o[o] = a2;
        }
    }
}

cudaError_t launch_shared_fixup(const double *d_ts, const double *d_Yt, int64_t n, int64_t n_pad, int64_t S,
                                int64_t S_pad, const double *d_tau, const double *d_omega, int64_t nt, int64_t nf,
                                double c, double thr, const double *d_wq, unsigned int *d_fix_count,
                                unsigned int *d_fix_list, double *d_neffs, double *d_amp, double *d_phase,
                                double *d_a0, double *d_a1, double *d_a2, int sms, cudaStream_t st)
{
    const int64_t ncell = nt * nf;
    if (ncell <= 0 || S <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(d_fix_count, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    flag_cells_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, st>>>(d_wq, ncell, d_fix_count, d_fix_list);
    faithful_batch_kernel<kSurrBlock><<<(unsigned)(sms * 8), 128, 0, st>>>(d_ts, d_Yt, n, n_pad, S, S_pad, d_tau, d_omega,
                                                                            nf, c, thr, d_fix_count, d_fix_list, d_neffs,
                                                                            d_amp, d_phase, d_a0, d_a1, d_a2);
    bump_launch_counter(2);
    return cudaGetLastError();
}

SharedPlan plan_shared(int64_t n, int64_t nt, int64_t nf)
{
    SharedPlan p{};
    int64_t ft = (kConsumerThreads + nt - 1) / nt + 1;
    if (ft > nf) ft = nf;
    if (ft < 1) ft = 1;
    p.ft_max = (int)ft;
    // stage = P*16 (t,y) + ft*(P+1)*16 ({s,c} rows) + P*NS*8 (series tile)
    int64_t P = (kStageBudget - 16 * ft) / (16 + 16 * ft + 8 * kSurrBlock);
    P &= ~(int64_t)1;   // keep the series tile 16-byte aligned
    if (P > 256) P = 256;
    if (P > n) P = ((n + 7) / 8) * 8;
    if (P < 8) P = 8;
    p.points = (int)P;
    p.n_pad = ((n + P - 1) / P) * P;
    p.row_stride = (P + 1) * 16;
    int64_t y_off = P * 16 + ft * p.row_stride;
    y_off = (y_off + 15) & ~(int64_t)15;
    p.y_off = y_off;
    p.stage_bytes = (int)(y_off + P * kSurrBlock * 8);
    p.smem_bytes = kStages * p.stage_bytes + 2 * kStages * 8 + 16 * 8 + 16;
    return p;
}

cudaError_t launch_rotation(const double *d_tau, const double *d_omega, int64_t nt, int64_t nf, double2 *d_rot,
                            cudaStream_t st)
{
    if (nt * nf <= 0) return cudaSuccess;
    rotation_kernel<<<(unsigned)((nt * nf + 255) / 256), 256, 0, st>>>(d_tau, d_omega, nt, nf, d_rot);
    bump_launch_counter();
    return cudaGetLastError();
}

cudaError_t launch_shared(const SharedPlan &plan, const double2 *d_ty, const double2 *d_basis, const double *d_Yt,
                          int64_t n, int64_t S_pad, const double *d_tau, int64_t nt, const double *d_kappa,
                          int64_t nf, const double2 *d_rot, double thr, double *d_neffs, double *d_wq, double *d_amp,
                          double *d_phase, double *d_a0, double *d_a1, double *d_a2, cudaStream_t st)
{
    const int64_t ncell = nt * nf;
    if (ncell <= 0 || S_pad <= 0) return cudaSuccess;
    SharedArgs a{};
    a.ty = d_ty;
    a.basis = d_basis;
    a.Yt = d_Yt;
    a.tau = d_tau;
    a.kappa = d_kappa;
    a.rot = d_rot;
    a.neffs = d_neffs;
    a.wq = d_wq;
    a.amp = d_amp;
    a.phase = d_phase;
    a.a0 = d_a0;
    a.a1 = d_a1;
    a.a2 = d_a2;
    a.n = n;
    a.n_pad = plan.n_pad;
    a.ncell = ncell;
    a.S_pad = S_pad;
    a.nt = (int)nt;
    a.nf = (int)nf;
    a.P = plan.points;
    a.nchunks = (int)((n + plan.points - 1) / plan.points);
    a.row_stride2 = (int)(plan.row_stride / 16);
    a.stage_stride2 = plan.stage_bytes / 16;
    a.y_off2 = (int)(plan.y_off / 16);
    a.thr = thr;
    static int configured = -1;
    if (plan.smem_bytes > configured) {
        cudaError_t e = cudaFuncSetAttribute(wwz_shared_kernel<kSurrBlock>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             plan.smem_bytes);
        if (e != cudaSuccess) return e;
        configured = plan.smem_bytes;
    }
    dim3 grid((unsigned)((ncell + kConsumerThreads - 1) / kConsumerThreads), (unsigned)(S_pad / kSurrBlock));
    wwz_shared_kernel<kSurrBlock><<<grid, kCellThreads, plan.smem_bytes, st>>>(a);
    bump_launch_counter();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// [ncell][S_pad] (cell-major) -> [S][ncell] (series-major, the layout of amps[ne, ntau, nf],
// core/scalograms.py:635)
// ------------------------------------------------------------------------------------------------
__global__ void transpose_kernel(const double *__restrict__ in, int64_t rows, int64_t cols, int64_t in_stride,
                                 double *__restrict__ out, int64_t out_stride)
{
    __shared__ double tile[32][33];
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int64_t r = r0 + y, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[y][threadIdx.x] = in[r * in_stride + c];
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int64_t c = c0 + y, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[c * out_stride + r] = tile[threadIdx.x][y];
    }
}

// out[c][r] = in[r][c] for r < rows, c < cols
cudaError_t launch_transpose(const double *d_in, int64_t rows, int64_t cols, int64_t in_stride, double *d_out,
                             int64_t out_stride, cudaStream_t st)
{
    if (rows <= 0 || cols <= 0) return cudaSuccess;
    const int64_t gy = (rows + 31) / 32;
    const int64_t gy_max = 65535;
    for (int64_t y0 = 0; y0 < gy; y0 += gy_max) {
        const int64_t ny = std::min<int64_t>(gy_max, gy - y0);
        dim3 grid((unsigned)((cols + 31) / 32), (unsigned)ny);
        transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(d_in + y0 * 32 * in_stride, rows - y0 * 32, cols, in_stride,
                                                       d_out + y0 * 32, out_stride);
        bump_launch_counter();
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Batched wwa2psd: psd[s][k] from the cell-major amplitudes amp[(j*nf+k)][S_pad] and the shared
// Neffs[j][k].  One thread per (k, s), rows added in order; s is the fast index (coalesced).
// ------------------------------------------------------------------------------------------------
__global__ void psd_batch_kernel(const double *__restrict__ amp, const double *__restrict__ neffs, int64_t nt,
                                 int64_t nf, int64_t S, int64_t S_pad, double tspan, double n_d, double thr,
                                 double *__restrict__ psd)
{
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t k = blockIdx.y;
    if (s >= S) return;
    double sp = 0.0, se = 0.0;
    for (int64_t j = 0; j < nt; ++j) {
        const double ne = neffs[j * nf + k];
        const double a = amp[(j * nf + k) * S_pad + s];
        const double power = a * a * 0.5 * tspan / n_d * ne;
        double d = ne - thr;
        if (d < 0.0) d = 0.0;
        const double pd = power * d;
        if (!isnan(pd)) sp += pd;
        if (!isnan(d)) se += d;
    }
    psd[s * nf + k] = sp / se;
}

cudaError_t launch_psd_batch(const double *d_amp, const double *d_neffs, int64_t nt, int64_t nf, int64_t S,
                             int64_t S_pad, double tspan, int64_t n, double thr, double *d_psd, cudaStream_t st)
{
    if (S <= 0 || nf <= 0) return cudaSuccess;
    dim3 grid((unsigned)((S + 127) / 128), (unsigned)nf);
    psd_batch_kernel<<<grid, 128, 0, st>>>(d_amp, d_neffs, nt, nf, S, S_pad, tspan, (double)n, thr, d_psd);
    bump_launch_counter();
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Per-cell quantiles, scipy.stats.mstats.mquantiles(alphap=betap=0.4) semantics on the sorted
// sample (NaN sorts last, like np.sort):
//   aleph = n*p + m, m = alphap + p*(1-alphap-betap), k = floor(clip(aleph, 1, n-1)),
//   gamma = clip(aleph-k, 0, 1), q = (1-gamma)*x[k-1] + gamma*x[k].
// One CTA per cell; the S values are sorted in shared memory by a bitonic network on keys that
// map NaN above +inf.  X is cell-major [ncell][stride].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long sort_key(double v)
{
    if (isnan(v)) return 0xFFFFFFFFFFFFFFFFull;
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ double key_value(unsigned long long k)
{
    if (k == 0xFFFFFFFFFFFFFFFFull) return CUDART_NAN;
    const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(1024)
    quantile_kernel(const double *__restrict__ X, int64_t S, int64_t stride, int64_t ncell, const double *__restrict__ probs,
                    int nq, int npow2, double *__restrict__ q, unsigned long long *__restrict__ gscratch)
{
    extern __shared__ unsigned long long keys_sh[];
    for (int64_t cell = blockIdx.x; cell < ncell; cell += gridDim.x) {
        unsigned long long *keys = gscratch ? gscratch + (size_t)blockIdx.x * npow2 : keys_sh;
        for (int i = threadIdx.x; i < npow2; i += blockDim.x)
            keys[i] = (i < S) ? sort_key(X[cell * stride + i]) : 0xFFFFFFFFFFFFFFFFull;
        __syncthreads();
        for (int size = 2; size <= npow2; size <<= 1) {
            for (int str = size >> 1; str > 0; str >>= 1) {
                for (int i = threadIdx.x; i < (npow2 >> 1); i += blockDim.x) {
                    const int lo = 2 * i - (i & (str - 1));   // index with bit `str` cleared
                    const int hi = lo + str;
                    const bool up = ((lo & size) == 0);
                    const unsigned long long a = keys[lo], b = keys[hi];
                    if ((a > b) == up) {
                        keys[lo] = b;
                        keys[hi] = a;
                    }
                }
                __syncthreads();
            }
        }
        // padding keys equal the NaN key; NaNs of the sample and padding are indistinguishable and both
        // sort last, so positions < S hold the sorted sample with its NaNs at the end.
        for (int iq = threadIdx.x; iq < nq; iq += blockDim.x) {
            // every operation rounded separately, in scipy's order, so that gamma (a cancellation of
            // n*p + m against k) comes out bit-identical to scipy.stats.mstats.mquantiles
            const double p = probs[iq];
            const double m = __dadd_rn(0.4, __dmul_rn(p, (1.0 - 0.4) - 0.4));
            const double aleph = __dadd_rn(__dmul_rn((double)S, p), m);
            const double kc = fmin(fmax(aleph, 1.0), (double)(S - 1));
            const long long kk = (long long)floor(kc);
            const double gamma = fmin(fmax(__dadd_rn(aleph, -(double)kk), 0.0), 1.0);
            double r;
            if (S == 1) {
                r = key_value(keys[0]);
            } else {
                const double x0 = key_value(keys[kk - 1]), x1 = key_value(keys[kk]);
                r = __dadd_rn(__dmul_rn(__dadd_rn(1.0, -gamma), x0), __dmul_rn(gamma, x1));
            }
            q[(size_t)iq * (size_t)ncell + (size_t)cell] = r;
        }
        __syncthreads();
    }
}

cudaError_t launch_quantiles(const double *d_X, int64_t S, int64_t stride, int64_t ncell, const double *d_probs,
                             int64_t nq, double *d_q, unsigned long long *d_scratch, int64_t scratch_ctas, int sms,
                             cudaStream_t st)
{
    if (ncell <= 0 || nq <= 0 || S <= 0) return cudaSuccess;
    int npow2 = 2;
    while (npow2 < S) npow2 <<= 1;
    const size_t smem = (size_t)npow2 * 8;
    if (smem <= 200 * 1024) {
        static size_t configured = 0;
        if (smem > configured) {
            cudaError_t e = cudaFuncSetAttribute(quantile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured = smem;
        }
        const int64_t grid = std::min<int64_t>(ncell, (int64_t)sms * 8);
        quantile_kernel<<<(unsigned)grid, 1024, smem, st>>>(d_X, S, stride, ncell, d_probs, (int)nq, npow2, d_q, nullptr);
    } else {
        if (!d_scratch || scratch_ctas < 1) return cudaErrorInvalidValue;
        const int64_t grid = std::min<int64_t>(ncell, scratch_ctas);
        quantile_kernel<<<(unsigned)grid, 1024, 0, st>>>(d_X, S, stride, ncell, d_probs, (int)nq, npow2, d_q, d_scratch);
    }
    bump_launch_counter();
    return cudaGetLastError();
}

}  // namespace wwz
