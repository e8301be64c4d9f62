// Developer microbenchmark: cost of one Horner step p = fma(p, g, C) on sm_100a for different sources of C.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cc[8] = {1.0, 0.043321698784996574, 0.00093839, 1.355e-5, 1.46e-7, 1.27e-9, 9.1e-12, 5e-14};
#define CH 8
template <int V> __global__ void __launch_bounds__(256) k(double* out, int iters, const double* gm) {
    double p[CH], g[CH];
    for (int j = 0; j < CH; ++j) { p[j] = gm[j] + threadIdx.x * 1e-9; g[j] = gm[8 + j] - threadIdx.x * 1e-9; }
    const double r0 = gm[16], r1 = gm[17], r2 = gm[18], r3 = gm[19];
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            if (V == 0) { p[j] = fma(p[j], g[j], 0.5); p[j] = fma(p[j], g[j], 0.25); p[j] = fma(p[j], g[j], 2.0); p[j] = fma(p[j], g[j], 1.0); }
            if (V == 1) { p[j] = fma(p[j], g[j], cc[1]); p[j] = fma(p[j], g[j], cc[1]); p[j] = fma(p[j], g[j], cc[1]); p[j] = fma(p[j], g[j], cc[1]); }
            if (V == 2) { p[j] = fma(p[j], g[j], cc[3]); p[j] = fma(p[j], g[j], cc[2]); p[j] = fma(p[j], g[j], cc[1]); p[j] = fma(p[j], g[j], cc[0]); }
            if (V == 3) { p[j] = fma(p[j], g[j], r3); p[j] = fma(p[j], g[j], r2); p[j] = fma(p[j], g[j], r1); p[j] = fma(p[j], g[j], r0); }
            if (V == 4) { p[j] = fma(p[j], 0.5, g[j]); p[j] = fma(p[j], 0.25, g[j]); p[j] = fma(p[j], 2.0, g[j]); p[j] = fma(p[j], 1.0000001, g[j]); }
            if (V == 5) { p[j] = fma(p[j], cc[3], g[j]); p[j] = fma(p[j], cc[2], g[j]); p[j] = fma(p[j], cc[1], g[j]); p[j] = fma(p[j], cc[0], g[j]); }
            if (V == 6) { p[j] = p[j] * g[j]; p[j] = p[j] + g[j]; p[j] = p[j] * g[j]; p[j] = p[j] + g[j]; }
            if (V == 7) { p[j] = p[j] * cc[3]; p[j] = p[j] + cc[2]; p[j] = p[j] * cc[1]; p[j] = p[j] + cc[0]; }
        }
    }
    double r = 0; for (int j = 0; j < CH; ++j) r += p[j];
    if (r == 123.456) out[0] = r;
}
template <int V> void run(const char* name, double* d_out, double* d_g) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, iters = 1 << 12;
    cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(t0); k<V><<<blocks, 256>>>(d_out, iters, d_g); cudaEventRecord(t1); cudaEventSynchronize(t1);
        float t; cudaEventElapsedTime(&t, t0, t1); if (rep >= 1 && t < best) best = t;
    }
    double inst = (double)iters * 4 * CH * blocks * 256 / 32;   // warp instructions
    printf("%-44s %8.3f ms  %5.2f cycles per warp-instr per SMSP\n", name, best, best * 1e-3 * 1.965e9 * sms * 4 / inst);
}
int main() {
    double h[20]; for (int i = 0; i < 20; ++i) h[i] = 0.5 + i * 1e-3;
    double *d_g, *d_out; cudaMalloc(&d_g, sizeof h); cudaMalloc(&d_out, 8); cudaMemcpy(d_g, h, sizeof h, cudaMemcpyHostToDevice);
    run<0>("fma(p,g,imm)      2 RF + immediate", d_out, d_g);
    run<1>("fma(p,g,c[1])     2 RF + one constant", d_out, d_g);
    run<2>("fma(p,g,c[k])     2 RF + 4 constants", d_out, d_g);
    run<3>("fma(p,g,r)        3 RF", d_out, d_g);
    run<4>("fma(p,imm,g)      2 RF + immediate (B slot)", d_out, d_g);
    run<5>("fma(p,c[k],g)     2 RF + constant (B slot)", d_out, d_g);
    run<6>("p*g, p+g          2 RF", d_out, d_g);
    run<7>("p*c, p+c          1 RF + constant", d_out, d_g);
    return 0;
}
