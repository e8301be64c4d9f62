// Developer microbenchmark (round 2): what does the FP64 pipe of sm_100a sustain on the accumulation pattern of the
// tau-recurrence kernel (8-cell multiply chain + 56 accumulations per point), from registers only, at 1..4 warps per SMSP?
#include <cstdio>
#include <cuda_runtime.h>
#include "../pyleoclim_util_b200/csrc/exp2_poly.inc"
__constant__ double cpoly[7] = {WWZ_EXP2T_C0, WWZ_EXP2T_C1, WWZ_EXP2T_C2, WWZ_EXP2T_C3, WWZ_EXP2T_C4, WWZ_EXP2T_C5, WWZ_EXP2T_C6};
constexpr double kShift = 6755399441055744.0;
// V = 0: groups of seven sharing w (cell-major); V = 1: sum-major (groups of eight sharing the point operand);
// V = 2: cell-major, two points per iteration; V = 3: accumulations only, all weights given (no multiply chain)
template <int V> __global__ void __launch_bounds__(128, 3) kacc(double* out, int iters, const double* g) {
    constexpr int M = 8;
    double W[M], Q[M], Ws[M], Wc[M], Wy[M], Wys[M], Wyc[M];
    for (int r = 0; r < M; ++r) W[r] = Q[r] = Ws[r] = Wc[r] = Wy[r] = Wys[r] = Wyc[r] = g[r] * threadIdx.x;
    double w0 = g[8] + 1e-9 * threadIdx.x, rho = g[9], y = g[10], s = g[11], c = g[12], ys = g[13], yc = g[14];
    const double d = g[15];
    for (int i = 0; i < iters; ++i) {
        w0 += d; rho -= d; s = -s; c += d; y += d; ys -= d; yc += d;   // stand-in for the loaded point (7 cheap FP64 ops)
        double wv[M];
        wv[0] = w0;
#pragma unroll
        for (int r = 1; r < M; ++r) wv[r] = (V == 3) ? w0 + r : wv[r - 1] * rho;
        if (V == 1) {
#pragma unroll
            for (int r = 0; r < M; ++r) W[r] += wv[r];
#pragma unroll
            for (int r = 0; r < M; ++r) Q[r] = fma(wv[r], wv[r], Q[r]);
#pragma unroll
            for (int r = 0; r < M; ++r) Ws[r] = fma(wv[r], s, Ws[r]);
#pragma unroll
            for (int r = 0; r < M; ++r) Wc[r] = fma(wv[r], c, Wc[r]);
#pragma unroll
            for (int r = 0; r < M; ++r) Wy[r] = fma(wv[r], y, Wy[r]);
#pragma unroll
            for (int r = 0; r < M; ++r) Wys[r] = fma(wv[r], ys, Wys[r]);
#pragma unroll
            for (int r = 0; r < M; ++r) Wyc[r] = fma(wv[r], yc, Wyc[r]);
        } else {
#pragma unroll
            for (int r = 0; r < M; ++r) {
                const double w = wv[r];
                W[r] += w; Q[r] = fma(w, w, Q[r]); Ws[r] = fma(w, s, Ws[r]); Wc[r] = fma(w, c, Wc[r]);
                Wy[r] = fma(w, y, Wy[r]); Wys[r] = fma(w, ys, Wys[r]); Wyc[r] = fma(w, yc, Wyc[r]);
            }
        }
    }
    double res = 0; for (int r = 0; r < M; ++r) res += W[r] + Q[r] + Ws[r] + Wc[r] + Wy[r] + Wys[r] + Wyc[r];
    if (res == 123.456) out[0] = res;
}
// the exp2 pair alone, U independent points per iteration
template <int U> __global__ void __launch_bounds__(128, 3) kexp(double* out, int iters, const double* g) {
    __shared__ double tbl[16];
    if (threadIdx.x < 16) tbl[threadIdx.x] = 1.0 + threadIdx.x / 16.0;
    __syncthreads();
    double a = g[0] + 1e-3 * threadIdx.x, acc = 0;
    const double step = g[1], d = g[15];
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double aa = a + u * d;
            const double z = fma(-aa, aa, kShift), Nf = z - kShift, gg = fma(-aa, aa, -Nf);
            const int N = __double2loint(z);
            const double b = aa * step, zr = b + kShift, Nrf = zr - kShift, gr = b - Nrf;
            const int Nr = __double2loint(zr);
            double p = cpoly[6], pr = cpoly[6];
#pragma unroll
            for (int k = 5; k >= 0; --k) { p = fma(p, gg, cpoly[k]); pr = fma(pr, gr, cpoly[k]); }
            p *= tbl[N & 15]; pr *= tbl[Nr & 15];
            const int hi = __double2hiint(p) + ((N >> 4) << 20), hir = __double2hiint(pr) + ((Nr >> 4) << 20);
            acc += __hiloint2double(hi, __double2loint(p)) + __hiloint2double(hir, __double2loint(pr));
        }
        a += 1e-7;
    }
    if (acc == 123.456) out[0] = acc;
}
template <typename K> void run(const char* name, K kern, double* d_out, double* d_g, int bps, double fp64_per_iter) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * bps, iters = 1 << 13;
    cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(t0); kern<<<blocks, 128>>>(d_out, iters, d_g); cudaEventRecord(t1); cudaEventSynchronize(t1);
        float t; cudaEventElapsedTime(&t, t0, t1); if (rep >= 1 && t < best) best = t;
    }
    // SMSP cycles per warp iteration: one warp per SMSP per block
    const double cyc = best * 1e-3 * 1.965e9 / iters / bps;
    printf("%-44s warps/SMSP=%d  %7.3f ms  %7.1f cyc/warp-iter (ideal %5.0f)  pipe %.1f%%\n", name, bps, best, cyc, 2 * fp64_per_iter,
           100.0 * 2 * fp64_per_iter / cyc);
}
int main() {
    double h[16] = {0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 0.999, 0.3, 0.6, 0.8, 0.18, 0.24, 1e-9};
    double *d_g, *d_out; cudaMalloc(&d_g, sizeof h); cudaMalloc(&d_out, 8); cudaMemcpy(d_g, h, sizeof h, cudaMemcpyHostToDevice);
    for (int bps : {1, 2, 3}) {
        run("acc cell-major (7 share w) + chain", kacc<0>, d_out, d_g, bps, 7 + 7 + 56);
        run("acc sum-major (8 share x) + chain", kacc<1>, d_out, d_g, bps, 7 + 7 + 56);
        run("acc cell-major, no chain", kacc<3>, d_out, d_g, bps, 7 + 7 + 56);
        run("exp pair x1", kexp<1>, d_out, d_g, bps, 22 + 4);
        run("exp pair x4", kexp<4>, d_out, d_g, bps, 4 * (22 + 3) + 1);
    }
    return 0;
}
