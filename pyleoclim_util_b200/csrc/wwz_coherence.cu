// Wavelet transform coherency of WWZ coefficient pairs, batched over surrogate pairs (sm_100a).
//
// Reference: pyleoclim/utils/wavelet.py:2240-2364 (wtc) and :2201-2238 (xwt), called by wwz_coherence (:1667-1672)
// for one pair of series and by Coherence.signif_test (core/coherences.py:748-963) once per surrogate pair through a
// process pool.  With coeff = a1 - i*a2 (:1664-1665) of both series on a common (tau, f) grid:
//
//   xwt = coeff1 * conj(coeff2),  power_k = |coeff_k|^2,  each divided by the scale 1/f (:2350-2352),
//   smoothing(.) = Gaussian in time -- multiplication by F_j(k) = exp(-smooth_factor * snorm_j^2 * k^2) between an FFT and
//                  an inverse FFT over npad = 2^ceil(log2 ntau) zero-padded points (:2318-2326) -- then a normalised
//                  boxcar over int(round(0.6 / dj * 2)) scales (:2331-2334, convolve2d 'same'),
//   xw_coherence = |S12|^2 / (S1 * S2),  xw_phase = angle(S12 / (sqrt(S1) sqrt(S2))),  xw_amplitude = |xwt|.
//
// The FFT / multiply / inverse FFT is the circular convolution of the zero-padded row with g_j = ifft(F_j) (real and
// even).  Here g is tabulated once per frequency (coh_kernel_table: a 2*npad-term cosine sum per entry) and the
// convolution is evaluated directly -- ntau is 50 by default, so a row costs 2,500 multiply-adds per smoothed
// quantity and nothing has to leave the SM.  A NaN coefficient (a cell masked by Neff) poisons its whole row of the
// time smoothing and then the boxcar neighbourhood, exactly as it does through the reference's FFT.
// Layouts: coefficients and outputs [S][ntau][nf] (tau-major rows, as every batched entry point returns them).
#include <math_constants.h>

#include <algorithm>

#include "wwz_internal.h"

namespace wwz {

// g[d][k] = (1/npad) * sum_q exp(-smooth * snorm_k^2 * (2 pi fftfreq_q)^2) * cos(2 pi q d / npad), table [npad][nf]
__global__ void coh_kernel_table(const double *__restrict__ freq, int64_t nf, double dt_tau, double smooth, int npad,
                                 double *__restrict__ g)
{
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= (int64_t)npad * nf) return;
    const int d = (int)(idx / nf);
    const int64_t k = idx - (int64_t)d * nf;
    const double snorm = (1.0 / freq[k]) / dt_tau;       // scales / dt (:2345)
    double acc = 0.0;
    for (int q = 0; q < npad; ++q) {
        const int qs = q < (npad + 1) / 2 ? q : q - npad;                  // np.fft.fftfreq ordering
        const double kk = 2.0 * 3.141592653589793 * ((double)qs / (double)npad);
        const double F = exp(-smooth * (snorm * snorm) * (kk * kk));
        // cos(2 pi q d / npad) with the argument reduced exactly (q*d mod npad is an integer)
        const int r = (int)(((long long)q * d) % npad);
        acc += F * cospi(2.0 * (double)r / (double)npad);
    }
    g[idx] = acc / (double)npad;
}

// time smoothing: T[s][x][i][k] = sum_l W_x[s][l][k] * g[(i - l) mod npad][k],  x = Re S12, Im S12, S1, S2 (before the
// boxcar), W = (xwt, power1, power2) / scale.  One thread per (s, i, k); k fastest (coalesced).
__global__ void __launch_bounds__(256)
    coh_time_kernel(const double *__restrict__ a1, const double *__restrict__ a2, const double *__restrict__ b1,
                    const double *__restrict__ b2, const double *__restrict__ freq, const double *__restrict__ g, int64_t S,
                    int nt, int64_t nf, int npad, double *__restrict__ T, double *__restrict__ xwa)
{
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t per = (int64_t)nt * nf;
    if (idx >= S * per) return;
    const int64_t s = idx / per;
    const int64_t rem = idx - s * per;
    const int i = (int)(rem / nf);
    const int64_t k = rem - (int64_t)i * nf;
    const double f = freq[k];                 // 1 / scale
    const double *pa1 = a1 + s * per + k, *pa2 = a2 + s * per + k, *pb1 = b1 + s * per + k, *pb2 = b2 + s * per + k;
    double xr = 0.0, xi = 0.0, p1 = 0.0, p2 = 0.0;
    const int mask = npad - 1;
    for (int l = 0; l < nt; ++l) {
        // coeff = a1 - i a2 (:1664-1665); xwt = c1 * conj(c2) (:2229)
        const double c1r = pa1[(int64_t)l * nf], c1i = -pa2[(int64_t)l * nf];
        const double c2r = pb1[(int64_t)l * nf], c2i = -pb2[(int64_t)l * nf];
        const double wr = (c1r * c2r + c1i * c2i) * f, wi = (c1i * c2r - c1r * c2i) * f;
        const double w1 = (c1r * c1r + c1i * c1i) * f, w2 = (c2r * c2r + c2i * c2i) * f;
        const double gg = g[(int64_t)((i - l) & mask) * nf + k];
        xr = fma(wr, gg, xr);
        xi = fma(wi, gg, xi);
        p1 = fma(w1, gg, p1);
        p2 = fma(w2, gg, p2);
        if (l == i && xwa) {                  // cross-wavelet amplitude of this cell (:2230), unsmoothed, unscaled
            const double pr = c1r * c2r + c1i * c2i, pi = c1i * c2r - c1r * c2i;
            xwa[idx] = sqrt(pr * pr + pi * pi);
        }
    }
    double *t = T + s * 4 * per + rem;
    t[0] = xr;
    t[per] = xi;
    t[2 * per] = p1;
    t[3 * per] = p2;
}

// boxcar over the scale (frequency) axis, then coherence and phase.  win = rect(L) normalised: 0.5 at both ends, 1 inside
// (L >= 2); 'same' alignment of scipy.signal.convolve2d: out[k] = sum_q T[q] * win[k + (L-1)/2 - q], zero outside.
__global__ void __launch_bounds__(256)
    coh_scale_kernel(const double *__restrict__ T, int64_t S, int nt, int64_t nf, int L, double *__restrict__ coh,
                     double *__restrict__ phase)
{
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t per = (int64_t)nt * nf;
    if (idx >= S * per) return;
    const int64_t s = idx / per;
    const int64_t rem = idx - s * per;
    const int64_t row = rem / nf, k = rem - row * nf;
    const double *t = T + s * 4 * per + row * nf;
    const double norm = L >= 2 ? 1.0 / (double)(L - 1) : 1.0;       // sum of rect(L) = L - 1
    const int off = (L - 1) / 2;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    // numpy/scipy order of the direct sum: q ascending (the value is the same to rounding either way)
    for (int m = L - 1; m >= 0; --m) {
        const int64_t q = k + off - m;
        if (q < 0 || q >= nf) continue;
        const double w = L == 1 ? 1.0 : ((m == 0 || m == L - 1) ? 0.5 : 1.0) * norm;
#pragma unroll
        for (int x = 0; x < 4; ++x) v[x] = fma(t[x * per + q], w, v[x]);
    }
    const double s12r = v[0], s12i = v[1], s1 = v[2], s2 = v[3];
    coh[idx] = (s12r * s12r + s12i * s12i) / (s1 * s2);             // |S12|^2 / (S1 * S2)  (:2355)
    if (phase) {
        const double den = sqrt(s1) * sqrt(s2);                      // (:2356-2357)
        phase[idx] = atan2(s12i / den, s12r / den);
    }
}

cudaError_t launch_wtc(const double *d_a1, const double *d_a2, const double *d_b1, const double *d_b2, int64_t S, int64_t nt,
                       int64_t nf, const double *d_freq, double dt_tau, double smooth, int npad, int L, double *d_g,
                       double *d_T, double *d_coh, double *d_phase, double *d_xwa, cudaStream_t st)
{
    if (S <= 0 || nt <= 0 || nf <= 0) return cudaSuccess;
    const int64_t ng = (int64_t)npad * nf, tot = S * nt * nf;
    coh_kernel_table<<<(unsigned)((ng + 255) / 256), 256, 0, st>>>(d_freq, nf, dt_tau, smooth, npad, d_g);
    coh_time_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_a1, d_a2, d_b1, d_b2, d_freq, d_g, S, (int)nt, nf, npad,
                                                                  d_T, d_xwa);
    coh_scale_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_T, S, (int)nt, nf, L, d_coh, d_phase);
    bump_launch_counter(3);
    return cudaGetLastError();
}

}  // namespace wwz
