// Device-side building blocks shared by the WWZ kernels (sm_100a).
//
//  * mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS: UBLKCP) wrappers,
//  * the Gaussian weight  w = exp(-c*omega^2*(t-tau)^2) = 2^(-u*u)  evaluated on the FP64 pipe.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "exp2_poly.inc"

namespace wwz {

// ----------------------------------------------------------------------------------------------
// mbarrier / TMA (PTX ISA: mbarrier.*, cp.async.bulk.shared::cluster.global)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// Make freshly initialised mbarriers visible to the async (TMA) proxy.
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 1-D bulk copy global -> shared, completion reported as bytes on `bar`.
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ----------------------------------------------------------------------------------------------
// Gaussian weight on the FP64 pipe.
//
//   w = exp(-c*omega^2*(t-tau)^2) = 2^(-u*u),  u = omega*sqrt(c*log2(e))*(t - tau)
//
// The kernels carry v = 4u (kappa4 = 4*omega*sqrt(c*log2 e)), so that -v*v = -16 u*u = N + g with
// N = rint(-v*v) = 16 n + j (j = 0..15) and |g| <= 1/2:
//
//   w = 2^n * 2^(j/16) * 2^(g/16)
//
// N and g come from the 1.5*2^52 shift trick; both the shift and the remainder are single FMAs
// on the *unrounded* product v*v, so g carries no error from squaring.  2^(g/16) is a degree-6
// minimax polynomial (rel. error 7e-18, tools/gen_exp2_poly.py) in Horner form with its
// coefficients in the constant bank; 2^(j/16) is a 16-entry table in shared memory (16 doubles =
// exactly the 32 banks: distinct entries never conflict); 2^n is one integer add on the exponent
// field.  FP64-pipe cost: DFMA + DADD + DFMA + 6 DFMA + DMUL = 10 (plus DADD + DMUL for v).
// ----------------------------------------------------------------------------------------------
constexpr double kShift = 6755399441055744.0;  // 1.5 * 2^52

__constant__ double c_exp2t[7] = {WWZ_EXP2T_C0, WWZ_EXP2T_C1, WWZ_EXP2T_C2, WWZ_EXP2T_C3,
                                  WWZ_EXP2T_C4, WWZ_EXP2T_C5, WWZ_EXP2T_C6};
__constant__ double c_exp2_table[16] = WWZ_EXP2_TABLE16;

// bits(kShift - 16336): z below this (as a signed 64-bit pattern) means N < -16336, i.e. n < -1021
// and the weight is below the normal range; the same test catches v*v beyond the int32 range and
// v = +-inf.
constexpr long long kTinyBits = 0x4337FFFFFFFFC030LL;

// Branch-free: weights below 2^-1021 are flushed to +0.0 here.  Cells whose weight sums are small
// enough for that to matter (or for w*w to underflow) are detected in the epilogue and recomputed
// by the faithful warp-per-cell kernel, which keeps subnormals like the reference's libm exp.
__device__ __forceinline__ double gauss_weight(double v, const double *__restrict__ tbl)
{
    const double z = fma(-v, v, kShift);   // kShift + rint(-v*v)
    const double Nf = z - kShift;          // rint(-v*v) as a double
    const double g = fma(-v, v, -Nf);      // (-v*v) - N, |g| <= 1/2
    const int N = __double2loint(z);
    const double T = tbl[N & 15];
    double p = c_exp2t[6];
    p = fma(p, g, c_exp2t[5]);
    p = fma(p, g, c_exp2t[4]);
    p = fma(p, g, c_exp2t[3]);
    p = fma(p, g, c_exp2t[2]);
    p = fma(p, g, c_exp2t[1]);
    p = fma(p, g, c_exp2t[0]);
    p *= T;
    const bool tiny = __double_as_longlong(z) < kTinyBits;
    const int hi = __double2hiint(p) + ((N >> 4) << 20);
    return __hiloint2double(tiny ? 0 : hi, tiny ? 0 : __double2loint(p));
}

// Same weight with the polynomial in Estrin form (dependency depth 3 instead of 6, two more multiplies): for
// kernels whose FP64 pipe is kept busy by other work (the DMMAs of wwz_mma.cu) and where the length of the
// dependent chain, not the instruction count, is what shows.  Agrees with gauss_weight to ~1 ulp.
__device__ __forceinline__ double gauss_weight_estrin(double v, const double *__restrict__ tbl)
{
    const double z = fma(-v, v, kShift);
    const double Nf = z - kShift;
    const double g = fma(-v, v, -Nf);
    const int N = __double2loint(z);
    const double T = tbl[N & 15];
    const double g2 = g * g;
    const double p01 = fma(c_exp2t[1], g, c_exp2t[0]);
    const double p23 = fma(c_exp2t[3], g, c_exp2t[2]);
    const double p45 = fma(c_exp2t[5], g, c_exp2t[4]);
    const double g4 = g2 * g2;
    const double lo = fma(p23, g2, p01);
    const double hi6 = fma(c_exp2t[6], g2, p45);
    double p = fma(hi6, g4, lo);
    p *= T;
    const bool tiny = __double_as_longlong(z) < kTinyBits;
    const int hi = __double2hiint(p) + ((N >> 4) << 20);
    return __hiloint2double(tiny ? 0 : hi, tiny ? 0 : __double2loint(p));
}

// Weight of the first of M tau-adjacent cells and the ratio between neighbours (tau-recurrence).
//
// For equally spaced tau_m = tau_0 + m*dtau:  v_m = a - m*D  with a = kappa4*(t - tau_0), D = kappa4*dtau,
//   w_m = 2^(-v_m^2/16) = w_0 * rho^m * h_m,   rho = 2^(a*D/8) = 2^(b/16) with b = 2*D*a,
//   h_m = 2^(-m^2 D^2/16)  (depends on the cell only: applied once in the epilogue).
// rho costs a second exp2 per (lane, point) but then every further cell of the lane costs one DMUL
// instead of an exp2.  When w_0 is flushed (|a| > 127.8) rho is forced to 0 as well, so that the
// products stay 0 instead of 0*inf; with D <= 8 and M <= 8 every weight dropped that way is below
// 2^-320 (the row classification in setup_kernel, wwz_kernels.cu).
__device__ __forceinline__ void gauss_weight_ratio(double a, double step16, const double *__restrict__ tbl,
                                                   double &w0, double &rho)
{
    const double z = fma(-a, a, kShift);
    const double Nf = z - kShift;
    const double g = fma(-a, a, -Nf);
    const int N = __double2loint(z);
    const double b = a * step16;
    const double zr = b + kShift;
    const double Nrf = zr - kShift;
    const double gr = b - Nrf;
    const int Nr = __double2loint(zr);
    const double T = tbl[N & 15], Tr = tbl[Nr & 15];
    double p = c_exp2t[6], pr = c_exp2t[6];
    p = fma(p, g, c_exp2t[5]);  pr = fma(pr, gr, c_exp2t[5]);
    p = fma(p, g, c_exp2t[4]);  pr = fma(pr, gr, c_exp2t[4]);
    p = fma(p, g, c_exp2t[3]);  pr = fma(pr, gr, c_exp2t[3]);
    p = fma(p, g, c_exp2t[2]);  pr = fma(pr, gr, c_exp2t[2]);
    p = fma(p, g, c_exp2t[1]);  pr = fma(pr, gr, c_exp2t[1]);
    p = fma(p, g, c_exp2t[0]);  pr = fma(pr, gr, c_exp2t[0]);
    p *= T;
    pr *= Tr;
    const bool tiny = __double_as_longlong(z) < kTinyBits;
    const int hi = __double2hiint(p) + ((N >> 4) << 20);
    const int hir = __double2hiint(pr) + ((Nr >> 4) << 20);
    w0 = __hiloint2double(tiny ? 0 : hi, tiny ? 0 : __double2loint(p));
    rho = __hiloint2double(tiny ? 0 : hir, tiny ? 0 : __double2loint(pr));
}

// The two halves of gauss_weight_ratio on their own, for the loop that advances rho from point to point
// (rho_{i+1} = rho_i * E_i, E tabulated per (row, point)) and evaluates it afresh once per chunk.
// gauss_weight_flag: w_0 and whether it was flushed; exp2_sixteenth: 2^(b/16) for |b| < 2^15.
__device__ __forceinline__ double gauss_weight_flag(double a, const double *__restrict__ tbl, bool &tiny)
{
    const double z = fma(-a, a, kShift);
    const double Nf = z - kShift;
    const double g = fma(-a, a, -Nf);
    const int N = __double2loint(z);
    const double T = tbl[N & 15];
    double p = c_exp2t[6];
    p = fma(p, g, c_exp2t[5]);
    p = fma(p, g, c_exp2t[4]);
    p = fma(p, g, c_exp2t[3]);
    p = fma(p, g, c_exp2t[2]);
    p = fma(p, g, c_exp2t[1]);
    p = fma(p, g, c_exp2t[0]);
    p *= T;
    tiny = __double_as_longlong(z) < kTinyBits;
    const int hi = __double2hiint(p) + ((N >> 4) << 20);
    return __hiloint2double(tiny ? 0 : hi, tiny ? 0 : __double2loint(p));
}

__device__ __forceinline__ double exp2_sixteenth(double b, const double *__restrict__ tbl)
{
    const double zr = b + kShift;
    const double Nrf = zr - kShift;
    const double gr = b - Nrf;
    const int Nr = __double2loint(zr);
    const double Tr = tbl[Nr & 15];
    double pr = c_exp2t[6];
    pr = fma(pr, gr, c_exp2t[5]);
    pr = fma(pr, gr, c_exp2t[4]);
    pr = fma(pr, gr, c_exp2t[3]);
    pr = fma(pr, gr, c_exp2t[2]);
    pr = fma(pr, gr, c_exp2t[1]);
    pr = fma(pr, gr, c_exp2t[0]);
    pr *= Tr;
    const int hir = __double2hiint(pr) + ((Nr >> 4) << 20);
    return __hiloint2double(hir, __double2loint(pr));
}

// The table lives in shared memory; call from the first 16 threads of a CTA before a barrier.
__device__ __forceinline__ void load_exp2_table(double *tbl, int tid)
{
    if (tid < 16) tbl[tid] = c_exp2_table[tid];
}

// Row of cell `cell` in the cell-major amplitude output of a shared-axis batch (AmpScatter, wwz_internal.h).
struct AmpScatterDev {
    double *peer[16];
    long long cells_per, dst_stride, col0;
};
__device__ __forceinline__ double *amp_row(const AmpScatterDev &sc, double *amp, size_t cell, long long S_pad)
{
    if (sc.cells_per == 0) return amp + cell * (size_t)S_pad;
    const size_t owner = cell / (size_t)sc.cells_per;
    return sc.peer[owner] + (cell - owner * (size_t)sc.cells_per) * (size_t)sc.dst_stride + sc.col0;
}

// ----------------------------------------------------------------------------------------------
// Per-cell epilogue shared by finalize_kernel (wwz_kernels.cu) and the fused epilogue of the recurrence flavour
// (wwz_rec.cu): Neff, the mask with numba's NaN semantics, and a0, a1, a2 from the seven sums
// acc = {W, Q, Ws, Wc, Wy, Wys, Wyc} (utils/wavelet.py:991-1032 reduced as in DESIGN.md section 2).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void cell_from_sums(const double (&acc)[7], double om, double tj, double thr, double &neff,
                                               double &a0, double &a1, double &a2)
{
    const double W = acc[0], Q = acc[1];
    neff = W * W / Q;
    a0 = a1 = a2 = __longlong_as_double(0x7ff8000000000000LL);
    if (neff > thr) {
        const double Ws = acc[2], Wc = acc[3], Wy = acc[4], Wys = acc[5], Wyc = acc[6];
        const double Y = Wy / W, S = Ws / W, C = Wc / W, YS = Wys / W, YC = Wyc / W;
        const double cyc = YC - Y * C;
        const double cys = YS - Y * S;
        const double th = om * tj;
        const double eps = fma(om, tj, -th);
        double s0, c0;
        sincos(th, &s0, &c0);
        const double st = fma(eps, c0, s0), ct = fma(-eps, s0, c0);
        a0 = Y;
        a1 = 2.0 * (ct * cyc + st * cys);
        a2 = 2.0 * (ct * cys - st * cyc);
    }
}

__device__ __forceinline__ void store_cell(int64_t o, double neff, double a0, double a1, double a2,
                                           double *__restrict__ neffs, double *__restrict__ a0o,
                                           double *__restrict__ a1o, double *__restrict__ a2o,
                                           double *__restrict__ wwa, double *__restrict__ phase)
{
    neffs[o] = neff;
    if (a0o) a0o[o] = a0;
    if (a1o) a1o[o] = a1;
    if (a2o) a2o[o] = a2;
    if (wwa) wwa[o] = sqrt(a1 * a1 + a2 * a2);   // utils/wavelet.py:1044
    if (phase) phase[o] = atan2(a2, a1);         // utils/wavelet.py:1045
}

// ----------------------------------------------------------------------------------------------
// Warp-cooperative search on the sorted time axis (point windows of a tile of cells).
// ----------------------------------------------------------------------------------------------
// First index i in [a, b) with ts[i] >= x (UPPER = false) or ts[i] > x (UPPER = true); b if there is none.
// All 32 lanes take part: 32 probes per round.
template <bool UPPER>
__device__ __forceinline__ long long warp_bound(const double *__restrict__ ts, long long a, long long b, double x,
                                                int lane)
{
    while (b - a > 32) {
        const long long step = (b - a + 32) / 33;
        const long long pidx = a + (long long)(lane + 1) * step - 1;
        bool below = false;
        if (pidx < b) {
            const double v = ts[pidx];
            below = UPPER ? (v <= x) : (v < x);
        }
        const int cnt = __popc(__ballot_sync(0xffffffffu, below));
        const long long na = a + (long long)cnt * step;                 // everything up to probe cnt-1 is below
        const long long nb = a + (long long)(cnt + 1) * step - 1;       // probe cnt is not below (if it exists)
        a = na;
        if (cnt < 32 && nb < b) b = nb;
    }
    bool below = false;
    if (a + lane < b) {
        const double v = ts[a + lane];
        below = UPPER ? (v <= x) : (v < x);
    }
    return a + __popc(__ballot_sync(0xffffffffu, below));
}

// ----------------------------------------------------------------------------------------------
// Faithful evaluation of one cell by one warp: lanes stride over the points, warp-shuffle
// reductions; the reference's own two-pass expressions (utils/wavelet.py:988-1032) with exp()
// keeping subnormal weights and w*w rounded separately.  All lanes return the same values.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void faithful_cell_warp(const double *__restrict__ ts, const double *__restrict__ ys,
                                                   long long ystride, long long n, double tj, double om, double c,
                                                   double thr, int lane, double &neff, double &a0, double &a1,
                                                   double &a2)
{
    double sw = 0.0, sw2 = 0.0;
    for (long long i = lane; i < n; i += 32) {
        const double dz = om * (ts[i] - tj);                 // :988
        const double w = exp(-c * (dz * dz));                // :989
        sw += w;                                             // :991
        sw2 = __dadd_rn(sw2, __dmul_rn(w, w));               // :992
    }
    sw = warp_sum(sw);
    sw2 = warp_sum(sw2);
    neff = sw * sw / sw2;
    a0 = a1 = a2 = __longlong_as_double(0x7ff8000000000000LL);
    if (!(neff > thr)) return;                               // :994 with numba's NaN semantics
    double s1 = 0, c1 = 0, sc = 0, ss = 0, cc = 0;
    for (long long i = lane; i < n; i += 32) {
        const double t = ts[i];
        const double dz = om * (t - tj);
        const double w = exp(-c * (dz * dz));
        double sb, cb;
        sincos(om * t, &sb, &cb);                            // :1002-1003
        s1 += w * sb;
        c1 += w * cb;
        sc += w * sb * cb;
        ss += w * sb * sb;
        cc += w * cb * cb;
    }
    const double sin_one = warp_sum(s1) / sw, cos_one = warp_sum(c1) / sw, sin_cos = warp_sum(sc) / sw;
    const double sin_sin = warp_sum(ss) / sw, cos_cos = warp_sum(cc) / sw;
    const double numerator = 2 * (sin_cos - sin_one * cos_one);                               // :1012
    const double denominator = (cos_cos - cos_one * cos_one) - (sin_sin - sin_one * sin_one); // :1013
    const double time_shift = atan2(numerator, denominator) / (2 * om);                       // :1014
    double ycs = 0, yss = 0, y1 = 0, cs1 = 0, ss1 = 0;
    for (long long i = lane; i < n; i += 32) {
        const double t = ts[i], y = ys[i * ystride];
        const double dz = om * (t - tj);
        const double w = exp(-c * (dz * dz));
        double sh, ch;
        sincos(om * (t - time_shift), &sh, &ch);             // :1016-1017
        ycs += w * y * ch;
        yss += w * y * sh;
        y1 += w * y;
        cs1 += w * ch;
        ss1 += w * sh;
    }
    const double ys_cos_shift = warp_sum(ycs) / sw, ys_sin_shift = warp_sum(yss) / sw;
    const double ys_one = warp_sum(y1) / sw;
    const double cos_shift_one = warp_sum(cs1) / sw, sin_shift_one = warp_sum(ss1) / sw;
    double stc, ctc;
    sincos(om * (time_shift - tj), &stc, &ctc);              // :1018-1019
    const double A = 2 * (ys_cos_shift - ys_one * cos_shift_one);   // :1027
    const double B = 2 * (ys_sin_shift - ys_one * sin_shift_one);   // :1028
    a0 = ys_one;                                             // :1030
    a1 = ctc * A - stc * B;                                  // :1031
    a2 = stc * A + ctc * B;                                  // :1032
}

}  // namespace wwz
