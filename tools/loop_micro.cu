// Developer microbenchmark: the inner arithmetic of wwz_cells_kernel in registers (no shared memory), with parts
// switched off, to see what each part costs on the B200 FP64 path.
#include <cstdio>
#include <cuda_runtime.h>
#include "../pyleoclim_util_b200/csrc/exp2_poly.inc"
__constant__ double cpoly[7] = {WWZ_EXP2T_C0, WWZ_EXP2T_C1, WWZ_EXP2T_C2, WWZ_EXP2T_C3, WWZ_EXP2T_C4, WWZ_EXP2T_C5, WWZ_EXP2T_C6};
constexpr double kShift = 6755399441055744.0;
constexpr long long kTinyBits = 0x4337FFFFFFFFC030LL;
// MODE bits: 1 = polynomial, 2 = int scaling + tiny select, 4 = accumulates (7), 8 = table lookup from smem
template <int MODE, int M> __global__ void __launch_bounds__(256) k(double* out, int iters, const double* g) {
    __shared__ double tbl[16];
    if (threadIdx.x < 16) tbl[threadIdx.x] = 1.0 + threadIdx.x / 16.0;
    __syncthreads();
    double tau[M], W[M], Q[M], Ws[M], Wc[M], Wy[M], Wys[M], Wyc[M];
    for (int r = 0; r < M; ++r) { tau[r] = g[r] + threadIdx.x * 0.01; W[r] = Q[r] = Ws[r] = Wc[r] = Wy[r] = Wys[r] = Wyc[r] = 0; }
    const double kap = g[8];
    double t = g[9], y = g[10], s = g[11], c = g[12], ys = g[13], yc = g[14];
    const double dt = g[15];
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            t += dt; s = -s; c += 1e-9; y += 1e-9; ys -= 1e-9; yc += 1e-9;    // stand-in for the LDS'd point
#pragma unroll
            for (int r = 0; r < M; ++r) {
                const double v = kap * (t - tau[r]);
                const double z = fma(-v, v, kShift);
                const double Nf = z - kShift;
                const double gg = fma(-v, v, -Nf);
                const int N = __double2loint(z);
                double p = gg;
                if (MODE & 1) {
                    p = cpoly[6];
                    p = fma(p, gg, cpoly[5]); p = fma(p, gg, cpoly[4]); p = fma(p, gg, cpoly[3]);
                    p = fma(p, gg, cpoly[2]); p = fma(p, gg, cpoly[1]); p = fma(p, gg, cpoly[0]);
                }
                if (MODE & 8) p *= tbl[N & 15];
                double w = p;
                if (MODE & 2) {
                    const bool tiny = __double_as_longlong(z) < kTinyBits;
                    const int hi = __double2hiint(p) + ((N >> 4) << 20);
                    w = __hiloint2double(tiny ? 0 : hi, tiny ? 0 : __double2loint(p));
                }
                if (MODE & 16) {       // y-shared form: products first, then three FMAs that share the operand y
                    const double ws = w * s, wc = w * c;
                    W[r] += w; Q[r] = fma(w, w, Q[r]); Ws[r] += ws; Wc[r] += wc;
                    Wy[r] = fma(w, y, Wy[r]); Wys[r] = fma(ws, y, Wys[r]); Wyc[r] = fma(wc, y, Wyc[r]);
                } else if (MODE & 4) {
                    W[r] += w; Q[r] = fma(w, w, Q[r]); Ws[r] = fma(w, s, Ws[r]); Wc[r] = fma(w, c, Wc[r]);
                    Wy[r] = fma(w, y, Wy[r]); Wys[r] = fma(w, ys, Wys[r]); Wyc[r] = fma(w, yc, Wyc[r]);
                } else {
                    W[r] += w;
                }
            }
        }
    }
    double res = 0; for (int r = 0; r < M; ++r) res += W[r] + Q[r] + Ws[r] + Wc[r] + Wy[r] + Wys[r] + Wyc[r];
    if (res == 123.456) out[0] = res;
}
template <int MODE, int M> void run(const char* name, double* d_out, double* d_g, int bps) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * bps, iters = 1 << 11;
    cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(t0); k<MODE, M><<<blocks, 256>>>(d_out, iters, d_g); cudaEventRecord(t1); cudaEventSynchronize(t1);
        float t; cudaEventElapsedTime(&t, t0, t1); if (rep >= 1 && t < best) best = t;
    }
    double cp = (double)iters * 2 * M * blocks * 256;     // cell-points
    double cyc = best * 1e-3 * 1.965e9 * sms * 4 / (cp / 32);   // SMSP-cycles per warp-cell-point
    printf("%-52s M=%d bps=%d  %7.3f ms  %6.2f cyc/warp-cell-pt  %.3e cell-pts/s\n", name, M, bps, best, cyc, cp / (best * 1e-3));
}
int main() {
    double h[16] = {0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.37, 0.0, 0.3, 0.6, 0.8, 0.18, 0.24, 1e-3};
    double *d_g, *d_out; cudaMalloc(&d_g, sizeof h); cudaMalloc(&d_out, 8); cudaMemcpy(d_g, h, sizeof h, cudaMemcpyHostToDevice);
    for (int bps : {2, 4, 8}) {
        run<15, 2>("full (poly+table+int+acc)", d_out, d_g, bps);
        run<7, 2>("no table", d_out, d_g, bps);
        run<13, 2>("no int scaling/select", d_out, d_g, bps);
        run<11, 2>("no accumulates (W only)", d_out, d_g, bps);
        run<14, 2>("no polynomial", d_out, d_g, bps);
        run<4, 2>("accumulates only (+v,z,Nf,g)", d_out, d_g, bps);
        run<0, 2>("v,z,Nf,g + W only", d_out, d_g, bps);
        run<15 + 16, 2>("full, y-shared accumulates", d_out, d_g, bps);
        run<15 + 16, 4>("full, y-shared accumulates", d_out, d_g, bps);
        run<15 + 16, 1>("full, y-shared accumulates", d_out, d_g, bps);
        run<15, 1>("full", d_out, d_g, bps);
        run<15, 4>("full", d_out, d_g, bps);
    }
    return 0;
}
