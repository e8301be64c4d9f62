// Foster's weighted wavelet Z-transform (the reference's method='Foster', wwz_basic, utils/wavelet.py:228-381):
// per cell a weighted projection of y on {1, cos w(t-tau), sin w(t-tau)} -- the 3x3 normal equations of
// :359-366 solved with a pseudo-inverse (np.linalg.pinv, rcond 1e-15) -- instead of Kirchner's centred
// covariances.
//
// Same weights, same hot kernels.  With s = sin(w t), c = cos(w t) the normal matrix in the unrotated basis
// {1, c, s} needs, beyond the seven sums of the Kirchner path, only  sum(w c^2)  and  sum(w s c)
// (sum(w s^2) = W - sum(w c^2)).  They are the "y*cos" and "y*sin" sums of a second pass of the unchanged cell
// kernels over basis records built with y := cos(w t).  The basis {1, cos w(t-tau), sin w(t-tau)} is the
// unrotated one turned by the angle w*tau in its last two members; a rotation is orthogonal, so the
// pseudo-inverse (same singular values, same cut-off) commutes with it: the system is solved unrotated and the
// solution (b_c, b_s) is turned into (ywave_2, ywave_3) = (C b_c + S b_s, C b_s - S b_c), C, S = cos, sin(w tau).
//
// Cells whose weight sums are deep in the underflow range (the same criterion as the Kirchner epilogue) are
// re-evaluated by one warp each with the reference's own expressions (:343-374).
#include <math_constants.h>

#include <algorithm>

#include "wwz_device.cuh"
#include "wwz_internal.h"

namespace wwz {

// Solution x = pinv(S) p for a symmetric 3x3 S (entries s00, s01, s02, s11, s12, s22): cyclic Jacobi
// eigen-decomposition, eigenvalues at or below rcond * max|lambda| dropped (np.linalg.pinv, rcond = 1e-15:
// singular values of a symmetric matrix are |lambda|).
__device__ __forceinline__ void sym3_pinv_apply(double s00, double s01, double s02, double s11, double s12, double s22,
                                                const double p[3], double x[3])
{
    double A[3][3] = {{s00, s01, s02}, {s01, s11, s12}, {s02, s12, s22}};
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 16; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        const double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (!(off > 1e-300) || off <= 1e-22 * diag) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int a = pq == 2 ? 1 : 0, b = pq == 0 ? 1 : 2;
            const double apq = A[a][b];
            if (apq == 0.0) continue;
            const double theta = (A[b][b] - A[a][a]) / (2.0 * apq);
            const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
            const double cs = rsqrt(fma(t, t, 1.0)), sn = t * cs;
            // A <- J^T A J, V <- V J with J the rotation in the (a, b) plane
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double ka = A[k][a], kb = A[k][b];
                A[k][a] = cs * ka - sn * kb;
                A[k][b] = sn * ka + cs * kb;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double ak = A[a][k], bk = A[b][k];
                A[a][k] = cs * ak - sn * bk;
                A[b][k] = sn * ak + cs * bk;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double ka = V[k][a], kb = V[k][b];
                V[k][a] = cs * ka - sn * kb;
                V[k][b] = sn * ka + cs * kb;
            }
        }
    }
    const double lmax = fmax(fabs(A[0][0]), fmax(fabs(A[1][1]), fabs(A[2][2])));
    const double cutoff = 1e-15 * lmax;
    x[0] = x[1] = x[2] = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double lam = A[i][i];
        const double proj = V[0][i] * p[0] + V[1][i] * p[1] + V[2][i] * p[2];
        // NaN input: the comparison is false for NaN, keep the NaN flowing like the arithmetic of the reference would
        const double coef = (fabs(lam) > cutoff || isnan(lam)) ? proj / lam : 0.0;
        x[0] = fma(coef, V[0][i], x[0]);
        x[1] = fma(coef, V[1][i], x[1]);
        x[2] = fma(coef, V[2][i], x[2]);
    }
}

__device__ __forceinline__ void store_foster(int64_t o, double neff, double y1, double y2, double y3,
                                             double *__restrict__ neffs, double *__restrict__ a0o,
                                             double *__restrict__ a1o, double *__restrict__ a2o,
                                             double *__restrict__ wwa, double *__restrict__ phase)
{
    neffs[o] = neff;
    if (a0o) a0o[o] = y1;
    if (a1o) a1o[o] = y2;
    if (a2o) a2o[o] = y3;
    if (wwa) wwa[o] = sqrt(y2 * y2 + y3 * y3);   // utils/wavelet.py:376
    if (phase) phase[o] = atan2(y3, y2);         // :377
}

// sums: pass 1 (y = series): W, Q, Ws, Wc, Wy, Wys, Wyc.  sums2: pass 2 (y := cos): planes 5, 6 = sum(w c s), sum(w c c).
__global__ void finalize_foster_kernel(const double *__restrict__ sums, const double *__restrict__ sums2,
                                       const double *__restrict__ tau, const double *__restrict__ omega, int64_t nt,
                                       int64_t nf, double thr, double *__restrict__ neffs, double *__restrict__ a0o,
                                       double *__restrict__ a1o, double *__restrict__ a2o, double *__restrict__ wwa,
                                       double *__restrict__ phase, int64_t row_stride,
                                       unsigned int *__restrict__ fix_count, unsigned int *__restrict__ fix_list,
                                       unsigned int fix_cap, int segs)
{
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t nc = nt * nf;
    if (idx >= nc) return;
    const int64_t j = idx / nf, k = idx - j * nf;
    double acc[kNumSums], Wsc = 0.0, Wcc = 0.0;
#pragma unroll
    for (int x = 0; x < kNumSums; ++x) acc[x] = 0.0;
    for (int sg = 0; sg < segs; ++sg) {
        const double *slab = sums + (size_t)sg * kNumSums * nc;
        const double *slab2 = sums2 + (size_t)sg * kNumSums * nc;
#pragma unroll
        for (int x = 0; x < kNumSums; ++x) acc[x] += slab[(size_t)x * nc + idx];
        Wsc += slab2[(size_t)5 * nc + idx];
        Wcc += slab2[(size_t)6 * nc + idx];
    }
    const double W = acc[0], Q = acc[1];
    if (fix_list && (W < 0x1p-200 || Q < 0x1p-400)) {
        const unsigned int slot = atomicAdd(fix_count, 1u);
        if (slot < fix_cap) fix_list[slot] = (unsigned int)idx;
    }
    const double neff = W * W / Q;
    double y1 = CUDART_NAN, y2 = CUDART_NAN, y3 = CUDART_NAN;
    if (neff > thr) {
        const double Ws = acc[2], Wc = acc[3], Wy = acc[4], Wys = acc[5], Wyc = acc[6];
        const double p[3] = {Wy / W, Wyc / W, Wys / W};
        double b[3];
        sym3_pinv_apply(1.0, Wc / W, Ws / W, Wcc / W, Wsc / W, (W - Wcc) / W, p, b);
        const double om = omega[k], tj = tau[j];
        const double th = om * tj;
        const double eps = fma(om, tj, -th);
        double s0, c0;
        sincos(th, &s0, &c0);
        const double st = fma(eps, c0, s0), ct = fma(-eps, s0, c0);
        y1 = b[0];
        y2 = ct * b[1] + st * b[2];
        y3 = ct * b[2] - st * b[1];
    }
    store_foster(j * row_stride + k, neff, y1, y2, y3, neffs, a0o, a1o, a2o, wwa, phase);
}

// One warp per cell with the reference's own expressions (utils/wavelet.py:343-374).
__global__ void __launch_bounds__(128)
    foster_faithful_kernel(const double *__restrict__ ts, const double *__restrict__ ys, int64_t n,
                           const double *__restrict__ tau, const double *__restrict__ omega, int64_t nt, int64_t nf,
                           double c, double thr, const unsigned int *__restrict__ fix_count,
                           const unsigned int *__restrict__ fix_list, unsigned int fix_cap,
                           double *__restrict__ neffs, double *__restrict__ a0o, double *__restrict__ a1o,
                           double *__restrict__ a2o, double *__restrict__ wwa, double *__restrict__ phase,
                           int64_t row_stride)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int64_t total = nt * nf;
    if (fix_list) {
        const unsigned int cnt = *fix_count;
        total = cnt < fix_cap ? cnt : fix_cap;
    }
    for (int64_t item = warp0; item < total; item += nwarps) {
        const int64_t idx = fix_list ? (int64_t)fix_list[item] : item;
        const int64_t j = idx / nf, k = idx - j * nf;
        const double tj = tau[j], om = omega[k];
        double sw = 0.0, sw2 = 0.0;
        for (int64_t i = lane; i < n; i += 32) {
            const double dz = om * (ts[i] - tj);                 // :343
            const double w = exp(-c * (dz * dz));                // :344
            sw += w;                                             // :346
            sw2 = __dadd_rn(sw2, __dmul_rn(w, w));               // :347
        }
        sw = warp_sum(sw);
        sw2 = warp_sum(sw2);
        const double neff = sw * sw / sw2;
        double y1 = CUDART_NAN, y2 = CUDART_NAN, y3 = CUDART_NAN;
        if (neff > thr) {                                        // :349 (a NaN Neff stays NaN)
            double c1 = 0, s1 = 0, cc = 0, ss = 0, cs = 0, yw = 0, yc = 0, ysn = 0;
            for (int64_t i = lane; i < n; i += 32) {
                const double dz = om * (ts[i] - tj);
                const double w = exp(-c * (dz * dz));
                double phi3, phi2;
                sincos(dz, &phi3, &phi2);                        // :354-355
                const double y = ys[i];
                c1 += w * phi2;
                s1 += w * phi3;
                cc += w * phi2 * phi2;
                ss += w * phi3 * phi3;
                cs += w * phi2 * phi3;
                yw += w * y;
                yc += w * phi2 * y;
                ysn += w * phi3 * y;
            }
            const double p[3] = {warp_sum(yw) / sw, warp_sum(yc) / sw, warp_sum(ysn) / sw};   // :368-370
            double x[3];
            sym3_pinv_apply(1.0, warp_sum(c1) / sw, warp_sum(s1) / sw, warp_sum(cc) / sw, warp_sum(cs) / sw,
                            warp_sum(ss) / sw, p, x);                                          // :357-366
            y1 = x[0];
            y2 = x[1];
            y3 = x[2];
        }
        if (lane == 0) store_foster(j * row_stride + k, neff, y1, y2, y3, neffs, a0o, a1o, a2o, wwa, phase);
    }
}

cudaError_t launch_finalize_foster(const double *d_sums, const double *d_sums2, const double *d_tau,
                                   const double *d_omega, int64_t nt, int64_t nf, double neff_thr, double *d_neffs,
                                   double *d_a0, double *d_a1, double *d_a2, double *d_wwa, double *d_phase,
                                   int64_t out_row_stride, unsigned int *d_fix_count, unsigned int *d_fix_list,
                                   unsigned int fix_cap, int segs, cudaStream_t st)
{
    const int64_t nc = nt * nf;
    if (nc <= 0) return cudaSuccess;
    finalize_foster_kernel<<<(unsigned)((nc + 127) / 128), 128, 0, st>>>(d_sums, d_sums2, d_tau, d_omega, nt, nf, neff_thr,
                                                                         d_neffs, d_a0, d_a1, d_a2, d_wwa, d_phase,
                                                                         out_row_stride, d_fix_count, d_fix_list,
                                                                         fix_cap, segs);
    bump_launch_counter();
    return cudaGetLastError();
}

cudaError_t launch_foster_faithful(const double *d_ts, const double *d_ys, int64_t n, const double *d_tau,
                                   const double *d_omega, int64_t nt, int64_t nf, double c, double neff_thr,
                                   const unsigned int *d_fix_count, const unsigned int *d_fix_list,
                                   unsigned int fix_cap, double *d_neffs, double *d_a0, double *d_a1, double *d_a2,
                                   double *d_wwa, double *d_phase, int64_t out_row_stride, int sms, cudaStream_t st)
{
    if (nt * nf <= 0) return cudaSuccess;
    int64_t blocks = (int64_t)sms * 8;
    if (!d_fix_list) blocks = std::min<int64_t>(blocks * 2, (nt * nf + 3) / 4);
    if (blocks < 1) blocks = 1;
    foster_faithful_kernel<<<(unsigned)blocks, 128, 0, st>>>(d_ts, d_ys, n, d_tau, d_omega, nt, nf, c, neff_thr,
                                                             d_fix_count, d_fix_list, fix_cap, d_neffs, d_a0, d_a1,
                                                             d_a2, d_wwa, d_phase, out_row_stride);
    bump_launch_counter();
    return cudaGetLastError();
}

}  // namespace wwz
