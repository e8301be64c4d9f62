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
//   w = 2^(-u*u),  u = kappa*(t - tau),  kappa = omega*sqrt(c*log2(e))
//
// -u*u is split as n + f with n = rint(-u*u), |f| <= 1/2, using the 1.5*2^52 shift trick; both
// the shift and the remainder are single FMAs on the *unrounded* product u*u, so f carries no
// error from squaring.  2^f is a degree-10 minimax polynomial (rel. error 2.1e-16, see
// tools/gen_exp2_poly.py) in Horner form: 10 DFMA.  2^n is applied with one integer add on the
// exponent field.
// FP64-pipe cost: 1 DFMA + 1 DADD + 1 DFMA + 10 DFMA = 13 (plus DADD + DMUL for u).
// ----------------------------------------------------------------------------------------------
constexpr double kShift = 6755399441055744.0;  // 1.5 * 2^52

// Polynomial coefficients live in the constant bank so that every Horner step is a single DFMA
// with a c[][] operand (literal doubles would cost two extra UMOV issue slots per step).
__constant__ double c_exp2[11] = {WWZ_EXP2_C0, WWZ_EXP2_C1, WWZ_EXP2_C2, WWZ_EXP2_C3, WWZ_EXP2_C4, WWZ_EXP2_C5,
                                  WWZ_EXP2_C6, WWZ_EXP2_C7, WWZ_EXP2_C8, WWZ_EXP2_C9, WWZ_EXP2_C10};

__device__ __forceinline__ double exp2_frac_poly(double f)
{
    double p = c_exp2[10];
    p = fma(p, f, c_exp2[9]);
    p = fma(p, f, c_exp2[8]);
    p = fma(p, f, c_exp2[7]);
    p = fma(p, f, c_exp2[6]);
    p = fma(p, f, c_exp2[5]);
    p = fma(p, f, c_exp2[4]);
    p = fma(p, f, c_exp2[3]);
    p = fma(p, f, c_exp2[2]);
    p = fma(p, f, c_exp2[1]);
    p = fma(p, f, c_exp2[0]);
    return p;
}

// bits(kShift - 1021): z below this (as a signed 64-bit pattern) means n < -1021, i.e. the weight is
// below the normal range; the same test catches u*u beyond the int32 range and u = +-inf.
constexpr long long kTinyBits = 0x4337FFFFFFFFFC03LL;

// Branch-free: weights below 2^-1021 are flushed to +0.0 here.  Cells whose weight sums are small
// enough for that to matter (or for w*w to underflow) are detected in the epilogue and recomputed
// by the faithful warp-per-cell kernel, which keeps subnormals like the reference's libm exp.
__device__ __forceinline__ double gauss_weight(double u)
{
    const double z = fma(-u, u, kShift);   // kShift + rint(-u*u)
    const double nf = z - kShift;          // rint(-u*u) as a double
    const double f = fma(-u, u, -nf);      // (-u*u) - n, |f| <= 1/2
    const double p = exp2_frac_poly(f);
    const int n = __double2loint(z);
    const bool tiny = __double_as_longlong(z) < kTinyBits;
    const int hi = __double2hiint(p) + (n << 20);
    return __hiloint2double(tiny ? 0 : hi, tiny ? 0 : __double2loint(p));
}

}  // namespace wwz
