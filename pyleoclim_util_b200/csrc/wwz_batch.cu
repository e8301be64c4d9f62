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
#include <cstring>
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

// ar1_model (uar1 == 0):  rho_i = exp(-(t_i - t_{i-1})/tau), q_i = sqrt(1 - rho_i^2)*sigma, y_0 = 0
//                         (utils/tsmodel.py:61-66)
// uar1_sim  (uar1 == 1):  phi_i = exp(-delta_i/tau), q_i = sqrt(sigma_2*(1 - phi_i^2)), y_0 = z_0
//                         (utils/tsmodel.py:702-711; `sigma` carries sigma_2, the innovation variance)
__global__ void ar1_coeff_kernel(const double *__restrict__ ts, int64_t n, double tau, double sigma, int uar1,
                                 double *__restrict__ rho, double *__restrict__ q)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) {
        rho[0] = 0.0;
        q[0] = uar1 ? 1.0 : 0.0;
        return;
    }
    if (uar1) {
        const double delta = ts[i] - ts[i - 1];
        const double phi = exp(-delta / tau);
        rho[i] = phi;
        q[i] = sqrt(sigma * (1.0 - phi * phi));
        return;
    }
    const double scaled_dt = (ts[i] - ts[i - 1]) / tau;
    const double r = exp(-scaled_dt);
    rho[i] = r;
    q[i] = sqrt(1.0 - r * r) * sigma;
}

// Generation in two kernels.  (1) ar1_normals_kernel: the standard-normal draws, fully parallel -- one thread per
// (series, Philox pair) writes z[2p], z[2p+1] of its series straight into Y.  (2) ar1_recur_kernel: the recurrence
// in place.  One warp = 8 series (the kernel is latency bound: 1,250 narrow warps hide DRAM round trips better than
// 313 wide ones); 32 steps at a time, the 8x32 tile goes through shared memory so that global accesses are
// row-contiguous (256 B per series) while lanes 0..7 walk their rows; the next tile is prefetched into registers
// while the current one is processed.
// y[i] = y[i-1]*rho_i + q_i*z_i (:67, uar1: :711) with rho_0 = 0 and q_0 = 0 (ar1_model: y[0] = 0, :61) or
// q_0 = 1 (uar1_sim: y[0] = z_0, :704-705); stored value is y + mu (core/surrogateseries.py:139, :156).
// Step i consumes z0 of Philox pair i/2 when i is even and z1 when i is odd.
__global__ void __launch_bounds__(256)
    ar1_normals_kernel(int64_t n, int64_t S, uint64_t first_series, uint64_t seed, double *__restrict__ Y)
{
    const int64_t npair = (n + 1) / 2;
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= S * npair) return;
    const int64_t s = idx / npair, pr = idx - s * npair;
    double z0, z1;
    normal_pair((uint64_t)pr, first_series + (uint64_t)s, seed, z0, z1);   // the stream of a series is keyed by its global index
    double *row = Y + s * n + 2 * pr;
    row[0] = z0;
    if (2 * pr + 1 < n) row[1] = z1;
}

constexpr int kRecurSeries = 8;   // series per warp of ar1_recur_kernel (latency bound: more warps, not wider ones)

__global__ void __launch_bounds__(128)
    ar1_recur_kernel(const double *__restrict__ rho, const double *__restrict__ q, int64_t n, int64_t S, double mu,
                     double *__restrict__ Y)
{
    __shared__ double tile[4][kRecurSeries][33];
    __shared__ double coef[4][2][32];            // rho, q of the 32 steps of the current tile
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t s0 = ((int64_t)blockIdx.x * 4 + wib) * kRecurSeries;
    if (s0 >= S) return;
    const int nrows = (int)min((int64_t)kRecurSeries, S - s0);
    double (*const tl)[33] = tile[wib];
    double nxt[kRecurSeries];
    auto fetch = [&](int64_t i0) {
#pragma unroll
        for (int r = 0; r < kRecurSeries; ++r)
            nxt[r] = (r < nrows && i0 + lane < n) ? Y[(s0 + r) * n + i0 + lane] : 0.0;
    };
    fetch(0);
    double y = 0.0;
    for (int64_t i0 = 0; i0 < n; i0 += 32) {
        const int cnt = (int)min((int64_t)32, n - i0);
#pragma unroll
        for (int r = 0; r < kRecurSeries; ++r) tl[r][lane] = nxt[r];
        coef[wib][0][lane] = lane < cnt ? rho[i0 + lane] : 0.0;      // steps beyond n: never stored
        coef[wib][1][lane] = lane < cnt ? q[i0 + lane] : 0.0;
        __syncwarp();
        if (i0 + 32 < n) fetch(i0 + 32);
        if (lane < kRecurSeries) {
            // fully unrolled: the 96 shared-memory loads are issued ahead of the dependent chain, one FMA per step
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                y = y * coef[wib][0][j] + coef[wib][1][j] * tl[lane][j];
                tl[lane][j] = y + mu;
            }
        }
        __syncwarp();
        for (int r = 0; r < nrows; ++r)
            if (lane < cnt) Y[(s0 + r) * n + i0 + lane] = tl[r][lane];
        __syncwarp();
    }
}

cudaError_t launch_ar1(const double *d_ts, int64_t n, double tau, double sigma, double mu, int64_t S, uint64_t seed,
                       bool uar1, double *d_rho, double *d_q, double *d_Y, cudaStream_t st, uint64_t first_series)
{
    if (n <= 0 || S <= 0) return cudaSuccess;
    ar1_coeff_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_ts, n, tau, sigma, uar1 ? 1 : 0, d_rho, d_q);
    const int64_t npair = (n + 1) / 2;
    ar1_normals_kernel<<<(unsigned)((S * npair + 255) / 256), 256, 0, st>>>(n, S, first_series, seed, d_Y);
    const int64_t per_block = 4 * kRecurSeries;
    ar1_recur_kernel<<<(unsigned)((S + per_block - 1) / per_block), 128, 0, st>>>(d_rho, d_q, n, S, mu, d_Y);
    bump_launch_counter(3);
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
__global__ void __launch_bounds__(256)
    retile_kernel(const double *__restrict__ Y, int64_t n, int64_t n_pad, int64_t S, const double *__restrict__ mu,
                  const double *__restrict__ sig, double *__restrict__ Yt)
{
    // block (b, 32 points): Yt[(b*n_pad + i)*NS + j] = (Y[(b*NS+j)*n + i] - mu_j)/sig_j, zero beyond n and S
    __shared__ double tile[NS][33];
    const int64_t b = blockIdx.y;
    const int64_t i0 = (int64_t)blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // blockDim = 32 x 8
    for (int jj = ty; jj < NS; jj += 8) {
        const int64_t s = b * NS + jj;
        const int64_t i = i0 + tx;
        double v = 0.0;
        if (s < S && i < n) v = (Y[s * n + i] - mu[s]) / sig[s];
        tile[jj][tx] = v;
    }
    __syncthreads();
    // write: NS contiguous doubles per point
    for (int idx = threadIdx.x; idx < 32 * NS; idx += blockDim.x) {
        const int p = idx / NS, j = idx - p * NS;
        const int64_t ip = i0 + p;
        if (ip < n_pad) Yt[(b * n_pad + ip) * NS + j] = tile[j][p];
    }
}

cudaError_t launch_standardize_retile(const double *d_Y, int64_t n, int64_t n_pad, int64_t S, int series_block,
                                      bool standardize, double *d_mu, double *d_sig, double *d_Yt, cudaStream_t st)
{
    if (S <= 0 || n <= 0) return cudaSuccess;
    row_stats_kernel<<<(unsigned)S, 256, 0, st>>>(d_Y, n, standardize ? 1 : 0, d_mu, d_sig);
    const int64_t blocks = (S + series_block - 1) / series_block;
    dim3 grid((unsigned)((n_pad + 31) / 32), (unsigned)blocks);
    if (series_block == 64)
        retile_kernel<64><<<grid, 256, 0, st>>>(d_Y, n, n_pad, S, d_mu, d_sig, d_Yt);
    else
        retile_kernel<16><<<grid, 256, 0, st>>>(d_Y, n, n_pad, S, d_mu, d_sig, d_Yt);
    bump_launch_counter(2);
    return cudaGetLastError();
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

__global__ void __launch_bounds__(128)
    faithful_batch_kernel(const double *__restrict__ ts, const double *__restrict__ Yt, int64_t n, int64_t n_pad,
                          int64_t S, int64_t S_pad, int NS, const double *__restrict__ tau, const double *__restrict__ omega,
                          int64_t nf, double c, double thr, const unsigned int *__restrict__ count,
                          const unsigned int *__restrict__ list, double *__restrict__ neffs, double *__restrict__ amp,
                          double *__restrict__ phase, double *__restrict__ a0o, double *__restrict__ a1o,
                          double *__restrict__ a2o, const AmpScatterDev sc)
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
            amp_row(sc, amp, (size_t)cell, S_pad)[s] = sqrt(a1 * a1 + a2 * a2);
            if (phase) phase[o] = atan2(a2, a1);
            if (a0o) a0o[o] = a0;
            if (a1o) a1o[o] = a1;
            if (a2o) a2o[o] = a2;
        }
    }
}

cudaError_t launch_shared_fixup(const double *d_ts, const double *d_Yt, int64_t n, int64_t n_pad, int64_t S,
                                int64_t S_pad, int series_block, const double *d_tau, const double *d_omega, int64_t nt, int64_t nf,
                                double c, double thr, const double *d_wq, unsigned int *d_fix_count,
                                unsigned int *d_fix_list, double *d_neffs, double *d_amp, double *d_phase,
                                double *d_a0, double *d_a1, double *d_a2, int sms, cudaStream_t st,
                                const AmpScatter *scatter)
{
    const int64_t ncell = nt * nf;
    if (ncell <= 0 || S <= 0) return cudaSuccess;
    AmpScatterDev sc{};
    if (scatter) memcpy(&sc, scatter, sizeof sc);
    cudaError_t e = cudaMemsetAsync(d_fix_count, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    flag_cells_kernel<<<(unsigned)((ncell + 255) / 256), 256, 0, st>>>(d_wq, ncell, d_fix_count, d_fix_list);
    faithful_batch_kernel<<<(unsigned)(sms * 8), 128, 0, st>>>(d_ts, d_Yt, n, n_pad, S, S_pad, series_block, d_tau, d_omega,
                                                                            nf, c, thr, d_fix_count, d_fix_list, d_neffs,
                                                                            d_amp, d_phase, d_a0, d_a1, d_a2, sc);
    bump_launch_counter(2);
    return cudaGetLastError();
}

cudaError_t launch_rotation(const double *d_tau, const double *d_omega, int64_t nt, int64_t nf, double2 *d_rot,
                            cudaStream_t st)
{
    if (nt * nf <= 0) return cudaSuccess;
    rotation_kernel<<<(unsigned)((nt * nf + 255) / 256), 256, 0, st>>>(d_tau, d_omega, nt, nf, d_rot);
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

// Row `cell` of the sample matrix: contiguous (seg_len == 0: X[cell*stride + i]) or cut into segments of seg_len
// values that lie seg_stride apart (the receive buffer of an all-to-all: segment r of every row comes from rank r).
__device__ __forceinline__ int64_t segmented_index(int64_t cell, int64_t i, int64_t stride, int64_t seg_len, int64_t seg_stride)
{
    if (seg_len == 0) return cell * stride + i;
    const int64_t r = i / seg_len;
    return r * seg_stride + cell * stride + (i - r * seg_len);
}

__global__ void __launch_bounds__(1024)
    quantile_kernel(const double *__restrict__ X, int64_t S, int64_t stride, int64_t ncell, const double *__restrict__ probs,
                    int nq, int npow2, double *__restrict__ q, unsigned long long *__restrict__ gscratch, int64_t seg_len,
                    int64_t seg_stride)
{
    extern __shared__ unsigned long long keys_sh[];
    for (int64_t cell = blockIdx.x; cell < ncell; cell += gridDim.x) {
        unsigned long long *keys = gscratch ? gscratch + (size_t)blockIdx.x * npow2 : keys_sh;
        for (int i = threadIdx.x; i < npow2; i += blockDim.x)
            keys[i] = (i < S) ? sort_key(X[segmented_index(cell, i, stride, seg_len, seg_stride)]) : 0xFFFFFFFFFFFFFFFFull;
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

// Few quantiles of many values (the significance test: one to three probabilities of 10,000 surrogates per
// cell): radix select instead of a sort.  One CTA per cell; the keys sit in shared memory (or are re-read from
// global memory above 25,600 values); per probability the order statistic k-1 is found by an 8-pass MSD radix
// select on the 64-bit keys (256-bin histogram per pass, warp-aggregated shared-memory atomics), then one more
// pass yields the order statistic k (the same key if it is repeated, else the smallest larger key).
// Same key order (NaN last) and the same interpolation arithmetic as quantile_kernel: results are identical.
template <bool SMEM_KEYS>
__global__ void __launch_bounds__(512)
    quantile_select_kernel(const double *__restrict__ X, int64_t S, int64_t stride, int64_t ncell,
                           const double *__restrict__ probs, int nq, double *__restrict__ q, int64_t seg_len,
                           int64_t seg_stride)
{
    extern __shared__ unsigned long long keys_sh[];
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long sel_prefix;
    __shared__ unsigned int sel_rank;
    __shared__ unsigned int cnt_le;
    __shared__ unsigned long long min_gt;
    const int tid = threadIdx.x, lane = tid & 31;
    const int N = (int)S;
    for (int64_t cell = blockIdx.x; cell < ncell; cell += gridDim.x) {
        if (SMEM_KEYS) {
            for (int i = tid; i < N; i += blockDim.x) keys_sh[i] = sort_key(X[segmented_index(cell, i, stride, seg_len, seg_stride)]);
        }
        __syncthreads();
        auto key_at = [&](int i) -> unsigned long long {
            return SMEM_KEYS ? keys_sh[i] : sort_key(X[segmented_index(cell, i, stride, seg_len, seg_stride)]);
        };
        for (int iq = 0; iq < nq; ++iq) {
            // every operation rounded separately, in scipy's order (see quantile_kernel)
            const double p = probs[iq];
            const double m = __dadd_rn(0.4, __dmul_rn(p, (1.0 - 0.4) - 0.4));
            const double aleph = __dadd_rn(__dmul_rn((double)S, p), m);
            const double kc = fmin(fmax(aleph, 1.0), (double)(S - 1));
            const long long kk = (long long)floor(kc);
            const double gamma = fmin(fmax(__dadd_rn(aleph, -(double)kk), 0.0), 1.0);
            const unsigned int r0 = (N == 1) ? 0u : (unsigned int)(kk - 1);   // 0-based rank of x[k-1]
            if (tid == 0) {
                sel_prefix = 0ull;
                sel_rank = r0;
            }
            unsigned long long mask = 0ull;
            for (int shift = 56; shift >= 0; shift -= 8) {
                if (tid < 256) hist[tid] = 0u;
                __syncthreads();
                const unsigned long long prefix = sel_prefix;
                for (int i0 = 0; i0 < N; i0 += blockDim.x) {
                    const int i = i0 + tid;
                    unsigned int digit = 0xFFFFFFFFu;                   // not a candidate
                    if (i < N) {
                        const unsigned long long k = key_at(i);
                        if ((k & mask) == prefix) digit = (unsigned int)(k >> shift) & 255u;
                    }
                    const unsigned int peers = __match_any_sync(0xffffffffu, digit);
                    if (digit != 0xFFFFFFFFu && lane == __ffs(peers) - 1) atomicAdd(&hist[digit], (unsigned int)__popc(peers));
                }
                __syncthreads();
                if (tid < 32) {
                    // bins 8*lane .. 8*lane+7: find the bin that holds the wanted rank
                    const unsigned int want = sel_rank;                  // read by all lanes before the shuffles
                    unsigned int c[8], tot = 0;
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        c[b] = hist[8 * lane + b];
                        tot += c[b];
                    }
                    unsigned int incl = tot;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += v;
                    }
                    const unsigned int excl = incl - tot;
                    if (want >= excl && want < incl) {
                        unsigned int acc = excl;
#pragma unroll
                        for (int b = 0; b < 8; ++b) {
                            if (want >= acc && want < acc + c[b]) {
                                sel_prefix = prefix | ((unsigned long long)(8 * lane + b) << shift);
                                sel_rank = want - acc;
                            }
                            acc += c[b];
                        }
                    }
                }
                mask |= 0xFFull << shift;
                __syncthreads();
            }
            const unsigned long long k0 = sel_prefix;                    // key of rank r0
            if (tid == 0) {
                cnt_le = 0u;
                min_gt = 0xFFFFFFFFFFFFFFFFull;
            }
            __syncthreads();
            {
                unsigned int le = 0;
                unsigned long long mg = 0xFFFFFFFFFFFFFFFFull;
                for (int i = tid; i < N; i += blockDim.x) {
                    const unsigned long long k = key_at(i);
                    if (k <= k0) ++le;
                    else if (k < mg) mg = k;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    le += __shfl_xor_sync(0xffffffffu, le, o);
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, mg, o);
                    mg = other < mg ? other : mg;
                }
                if (lane == 0) {
                    atomicAdd(&cnt_le, le);
                    atomicMin(&min_gt, mg);
                }
            }
            __syncthreads();
            if (tid == 0) {
                double r;
                if (N == 1) {
                    r = key_value(k0);
                } else {
                    const unsigned long long k1 = (cnt_le >= r0 + 2u) ? k0 : min_gt;   // key of rank r0 + 1
                    const double x0 = key_value(k0), x1 = key_value(k1);
                    r = __dadd_rn(__dmul_rn(__dadd_rn(1.0, -gamma), x0), __dmul_rn(gamma, x1));
                }
                q[(size_t)iq * (size_t)ncell + (size_t)cell] = r;
            }
            __syncthreads();
        }
    }
}

cudaError_t launch_quantiles(const double *d_X, int64_t S, int64_t stride, int64_t ncell, const double *d_probs,
                             int64_t nq, double *d_q, unsigned long long *d_scratch, int64_t scratch_ctas, int sms,
                             cudaStream_t st, int64_t seg_len, int64_t seg_stride)
{
    if (ncell <= 0 || nq <= 0 || S <= 0) return cudaSuccess;
    if (nq <= 16 && S >= 64 && S < ((int64_t)1 << 31)) {
        // few probabilities: radix select
        const size_t kb = (size_t)S * 8;
        if (kb <= 200 * 1024) {
            static SmemOptIn done_sel;
            cudaError_t e = ensure_dynamic_smem(quantile_select_kernel<true>, (int)kb, done_sel);
            if (e != cudaSuccess) return e;
            const int64_t grid = std::min<int64_t>(ncell, (int64_t)sms * 16);
            quantile_select_kernel<true><<<(unsigned)grid, 512, kb, st>>>(d_X, S, stride, ncell, d_probs, (int)nq, d_q, seg_len, seg_stride);
        } else {
            const int64_t grid = std::min<int64_t>(ncell, (int64_t)sms * 4);
            quantile_select_kernel<false><<<(unsigned)grid, 512, 0, st>>>(d_X, S, stride, ncell, d_probs, (int)nq, d_q, seg_len, seg_stride);
        }
        bump_launch_counter();
        return cudaGetLastError();
    }
    int npow2 = 2;
    while (npow2 < S) npow2 <<= 1;
    const size_t smem = (size_t)npow2 * 8;
    if (smem <= 200 * 1024) {
        static SmemOptIn done;
        cudaError_t e = ensure_dynamic_smem(quantile_kernel, (int)smem, done);
        if (e != cudaSuccess) return e;
        const int64_t grid = std::min<int64_t>(ncell, (int64_t)sms * 8);
        quantile_kernel<<<(unsigned)grid, 1024, smem, st>>>(d_X, S, stride, ncell, d_probs, (int)nq, npow2, d_q, nullptr, seg_len, seg_stride);
    } else {
        if (!d_scratch || scratch_ctas < 1) return cudaErrorInvalidValue;
        const int64_t grid = std::min<int64_t>(ncell, scratch_ctas);
        quantile_kernel<<<(unsigned)grid, 1024, 0, st>>>(d_X, S, stride, ncell, d_probs, (int)nq, npow2, d_q, d_scratch, seg_len, seg_stride);
    }
    bump_launch_counter();
    return cudaGetLastError();
}

}  // namespace wwz
