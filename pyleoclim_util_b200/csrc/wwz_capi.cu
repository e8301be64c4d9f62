// C ABI of libwwz_b200.so (declared in include/wwz_b200.h).  Host-side orchestration only:
// argument checks, a grow-only per-process device workspace, H2D/D2H staging, frequency batching
// of the basis scratch, kernel sequencing.  All arithmetic lives in wwz_kernels.cu /
// wwz_batch.cu.  There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/wwz_b200.h"
#include "wwz_internal.h"

namespace {

thread_local std::string g_err;
std::mutex g_mu;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(e_ == cudaErrorMemoryAllocation ? WWZB200_ERR_NOMEM : WWZB200_ERR_CUDA,       \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);  \
    } while (0)

// Optional in-library timing of the hot kernel (wwzb200_profile): CUDA events recorded on the
// launching stream around every cell-kernel launch, read back by wwzb200_profile_read.
bool g_profile = false;
struct EventPair {
    cudaEvent_t a, b;
    double cell_points;
};
std::vector<EventPair> g_events;

// Grow-only named device buffers, one set per device.
struct Workspace {
    std::map<std::string, std::pair<void *, size_t>> bufs;
    cudaStream_t stream = nullptr;
    int sms = 0;

    int get(const char *name, size_t bytes, void **out)
    {
        auto &b = bufs[name];
        if (b.second < bytes || b.first == nullptr) {
            if (b.first) cudaFree(b.first);
            b.first = nullptr;
            b.second = 0;
            size_t want = bytes < 256 ? 256 : bytes;
            cudaError_t e = cudaMalloc(&b.first, want);
            if (e != cudaSuccess) {
                b.first = nullptr;
                return fail(WWZB200_ERR_NOMEM, "cudaMalloc(%zu bytes) for '%s' failed: %s", want, name,
                            cudaGetErrorString(e));
            }
            b.second = want;
        }
        *out = b.first;
        return 0;
    }
    void release()
    {
        for (auto &kv : bufs)
            if (kv.second.first) cudaFree(kv.second.first);
        bufs.clear();
    }
};

std::map<int, Workspace> g_ws;

int workspace(Workspace **out)
{
    int dev = 0;
    CU(cudaGetDevice(&dev));
    Workspace &w = g_ws[dev];
    if (!w.stream) {
        CU(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
        CU(cudaDeviceGetAttribute(&w.sms, cudaDevAttrMultiProcessorCount, dev));
    }
    *out = &w;
    return 0;
}

template <typename T>
int wsbuf(Workspace *w, const char *name, size_t count, T **out)
{
    void *p = nullptr;
    int rc = w->get(name, count * sizeof(T), &p);
    *out = static_cast<T *>(p);
    return rc;
}

size_t basis_budget_bytes()
{
    const char *e = getenv("WWZB200_BASIS_BYTES");
    if (e && *e) {
        double v = atof(e);
        if (v >= 1 << 20) return (size_t)v;
    }
    return (size_t)2 << 30;
}

int cells_per_lane_hint(int64_t nt, int64_t nf, int sms)
{
    const char *e = getenv("WWZB200_CELLS_PER_LANE");
    if (e && *e) return atoi(e);
    // two tau-adjacent cells per lane unless that leaves fewer than ~4 CTAs per SM
    const int64_t tiles2 = (nf * ((nt + 1) / 2) + wwz::kConsumerThreads - 1) / wwz::kConsumerThreads;
    return tiles2 >= 4 * (int64_t)sms ? 2 : 1;
}

// The transform on device-resident inputs.  Outputs may alias caller memory (row stride given).
// d_psd != NULL adds the wwa2psd reduction; tspan < 0 means "take max-min of d_ts on the device".
int transform_device(Workspace *w, const double *d_ts, const double *d_ys, int64_t n, const double *d_tau,
                     int64_t nt, const double *d_omega, int64_t nf, double c, int64_t thr, int64_t psd_thr,
                     uint32_t flags, double *d_neffs, double *d_a0, double *d_a1, double *d_a2, double *d_wwa,
                     double *d_phase, double *d_psd, int64_t row_stride, double tspan, cudaStream_t st)
{
    using namespace wwz;
    const int64_t ncell = nt * nf;
    if (ncell == 0) return 0;
    if (ncell >= ((int64_t)1 << 32)) return fail(WWZB200_ERR_ARG, "nt*nf = %lld exceeds 2^32 cells", (long long)ncell);

    double *d_kappa, *d_dcut, *d_sums;
    int *d_unsorted;
    unsigned int *d_fix_count, *d_fix_list;
    int rc;
    if ((rc = wsbuf(w, "kappa", (size_t)nf, &d_kappa))) return rc;
    if ((rc = wsbuf(w, "dcut", (size_t)nf, &d_dcut))) return rc;
    if ((rc = wsbuf(w, "flags", 4, &d_unsorted))) return rc;
    d_fix_count = reinterpret_cast<unsigned int *>(d_unsorted + 1);
    if ((rc = wsbuf(w, "fix_list", (size_t)ncell, &d_fix_list))) return rc;

    if (d_psd && !d_wwa) {
        if ((rc = wsbuf(w, "psd_wwa", (size_t)ncell, &d_wwa))) return rc;
        if (row_stride != nf) return fail(WWZB200_ERR_ARG, "psd without wwa output needs row_stride == nf");
    }

    if (flags & WWZB200_FAITHFUL) {
        CU(launch_faithful(d_ts, d_ys, n, d_tau, d_omega, nt, nf, c, (double)thr, nullptr, nullptr, 0, d_neffs, d_a0,
                           d_a1, d_a2, d_wwa, d_phase, row_stride, w->sms, st));
    } else {
        const CellPlan plan = plan_cells(n, nt, nf, cells_per_lane_hint(nt, nf, w->sms));
        size_t per_row = (size_t)plan.n_pad * 32;
        int64_t nf_batch = (int64_t)std::max<size_t>(1, basis_budget_bytes() / per_row);
        nf_batch = std::min<int64_t>(std::min<int64_t>(nf_batch, nf), 65535);
        double2 *d_ty, *d_basis;
        if ((rc = wsbuf(w, "ty", (size_t)plan.n_pad, &d_ty))) return rc;
        if ((rc = wsbuf(w, "basis", (size_t)nf_batch * (size_t)plan.n_pad * 2, &d_basis))) return rc;
        if ((rc = wsbuf(w, "sums", (size_t)kNumSums * (size_t)ncell, &d_sums))) return rc;

        CU(launch_freq_setup(d_omega, nf, c, d_kappa, d_dcut, st));
        CU(launch_unsorted_check(d_ts, n, d_unsorted, st));
        CU(cudaMemsetAsync(d_fix_count, 0, sizeof(unsigned int), st));
        const bool dense = (flags & WWZB200_DENSE) != 0;
        for (int64_t k0 = 0; k0 < nf; k0 += nf_batch) {
            const int64_t nb = std::min<int64_t>(nf_batch, nf - k0);
            CU(launch_basis(d_ts, d_ys, n, plan.n_pad, d_omega + k0, nb, d_ty, d_basis, k0 == 0, st));
            EventPair ev{nullptr, nullptr, (double)n * (double)nt * (double)nb};
            if (g_profile) {
                CU(cudaEventCreate(&ev.a));
                CU(cudaEventCreate(&ev.b));
                CU(cudaEventRecord(ev.a, st));
            }
            CU(launch_cells(plan, d_ty, d_basis, d_ts, n, d_tau, nt, d_kappa, d_dcut, d_unsorted, k0, nb, nf, dense,
                            d_sums, st));
            if (g_profile) {
                CU(cudaEventRecord(ev.b, st));
                g_events.push_back(ev);
            }
        }
        CU(launch_finalize(d_sums, d_tau, d_omega, nt, nf, (double)thr, d_neffs, d_a0, d_a1, d_a2, d_wwa, d_phase,
                           row_stride, d_fix_count, d_fix_list, (unsigned int)ncell, st));
        CU(launch_faithful(d_ts, d_ys, n, d_tau, d_omega, nt, nf, c, (double)thr, d_fix_count, d_fix_list,
                           (unsigned int)ncell, d_neffs, d_a0, d_a1, d_a2, d_wwa, d_phase, row_stride, w->sms, st));
    }
    if (d_psd) {
        if (tspan >= 0) {
            CU(launch_psd(d_wwa, d_neffs, nt, nf, row_stride, tspan, n, (double)psd_thr, d_psd, st));
        } else {
            double *d_mm;
            if ((rc = wsbuf(w, "minmax", 2, &d_mm))) return rc;
            CU(launch_minmax(d_ts, n, d_mm, st));
            CU(launch_psd_devspan(d_wwa, d_neffs, nt, nf, row_stride, d_mm, n, (double)psd_thr, d_psd, st));
        }
    }
    return 0;
}

int check_grid_args(const void *ts, const void *ys, int64_t n, const void *tau, int64_t nt, const void *om,
                    int64_t nf, double c, int64_t thr)
{
    if (n < 1 || nt < 0 || nf < 0) return fail(WWZB200_ERR_ARG, "sizes must satisfy nts>=1, nt>=0, nf>=0");
    if (!ts || !ys || (nt && !tau) || (nf && !om)) return fail(WWZB200_ERR_ARG, "null input pointer");
    if (thr < 1) return fail(WWZB200_ERR_ARG, "Neff_threshold must be a positive integer (utils/wavelet.py:971)");
    if (!(c > 0) || !std::isfinite(c)) return fail(WWZB200_ERR_ARG, "decay constant c must be positive and finite");
    return 0;
}

int host_transform(const double *ts, const double *ys, int64_t n, const double *tau, int64_t nt, const double *om,
                   int64_t nf, double c, int64_t thr, int64_t psd_thr, uint32_t flags, double *psd, double *Neffs,
                   double *a0, double *a1, double *a2, double *wwa, double *phase)
{
    int rc = check_grid_args(ts, ys, n, tau, nt, om, nf, c, thr);
    if (rc) return rc;
    if (psd && psd_thr < 0) return fail(WWZB200_ERR_ARG, "psd_Neff_threshold must be >= 0");
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    if ((rc = workspace(&w))) return rc;
    cudaStream_t st = w->stream;
    const int64_t ncell = nt * nf;
    double *d_in, *d_out;
    const size_t in_count = (size_t)(2 * n + nt + nf);
    if ((rc = wsbuf(w, "in", in_count, &d_in))) return rc;
    if ((rc = wsbuf(w, "out", (size_t)(6 * ncell + nf), &d_out))) return rc;
    double *d_ts = d_in, *d_ys = d_in + n, *d_tau = d_ys + n, *d_om = d_tau + nt;
    CU(cudaMemcpyAsync(d_ts, ts, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_ys, ys, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    if (nt) CU(cudaMemcpyAsync(d_tau, tau, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
    if (nf) CU(cudaMemcpyAsync(d_om, om, sizeof(double) * nf, cudaMemcpyHostToDevice, st));
    double *o_neff = d_out, *o_a0 = d_out + ncell, *o_a1 = d_out + 2 * ncell, *o_a2 = d_out + 3 * ncell;
    double *o_wwa = d_out + 4 * ncell, *o_ph = d_out + 5 * ncell, *o_psd = d_out + 6 * ncell;
    double tspan = 0;
    if (psd) {
        const auto mm = std::minmax_element(ts, ts + n);
        tspan = *mm.second - *mm.first;   // utils/wavelet.py:1270 (np.max(ts)-np.min(ts))
    }
    rc = transform_device(w, d_ts, d_ys, n, d_tau, nt, d_om, nf, c, thr, psd_thr, flags, o_neff, o_a0, o_a1, o_a2, o_wwa, o_ph, psd ? o_psd : nullptr, nf, tspan, st);
    if (rc) return rc;
    const size_t cb = sizeof(double) * (size_t)ncell;
    if (ncell) {
        if (Neffs) CU(cudaMemcpyAsync(Neffs, o_neff, cb, cudaMemcpyDeviceToHost, st));
        if (a0) CU(cudaMemcpyAsync(a0, o_a0, cb, cudaMemcpyDeviceToHost, st));
        if (a1) CU(cudaMemcpyAsync(a1, o_a1, cb, cudaMemcpyDeviceToHost, st));
        if (a2) CU(cudaMemcpyAsync(a2, o_a2, cb, cudaMemcpyDeviceToHost, st));
        if (wwa) CU(cudaMemcpyAsync(wwa, o_wwa, cb, cudaMemcpyDeviceToHost, st));
        if (phase) CU(cudaMemcpyAsync(phase, o_ph, cb, cudaMemcpyDeviceToHost, st));
    }
    if (psd && nf) CU(cudaMemcpyAsync(psd, o_psd, sizeof(double) * nf, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

}  // namespace

extern "C" {

const char *wwzb200_last_error(void) { return g_err.c_str(); }
int wwzb200_version(void) { return 100; }

int wwzb200_device_count(int *count)
{
    if (!count) return fail(WWZB200_ERR_ARG, "null pointer");
    CU(cudaGetDeviceCount(count));
    return 0;
}

int wwzb200_set_device(int device)
{
    CU(cudaSetDevice(device));
    return 0;
}

int wwzb200_get_device(int *device)
{
    if (!device) return fail(WWZB200_ERR_ARG, "null pointer");
    CU(cudaGetDevice(device));
    return 0;
}

int wwzb200_sm_count(int *sms)
{
    if (!sms) return fail(WWZB200_ERR_ARG, "null pointer");
    int dev = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
    return 0;
}

int wwzb200_release(void)
{
    std::lock_guard<std::mutex> lock(g_mu);
    int dev = 0;
    CU(cudaGetDevice(&dev));
    auto it = g_ws.find(dev);
    if (it != g_ws.end()) {
        CU(cudaDeviceSynchronize());
        it->second.release();
    }
    return 0;
}

int wwzb200_alloc_pinned(size_t bytes, void **ptr)
{
    if (!ptr) return fail(WWZB200_ERR_ARG, "null pointer");
    CU(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return 0;
}

int wwzb200_free_pinned(void *ptr)
{
    if (ptr) CU(cudaFreeHost(ptr));
    return 0;
}

int wwzb200_kirchner(const double *ts, const double *pd_ys, int64_t nts, const double *taus, int64_t nt,
                     const double *omegas, int64_t nf, double c, int64_t Neff_threshold, uint32_t flags,
                     double *Neffs, double *a0, double *a1, double *a2, double *wwa, double *phase)
{
    if (!Neffs && nt * nf) return fail(WWZB200_ERR_ARG, "Neffs output is required");
    return host_transform(ts, pd_ys, nts, taus, nt, omegas, nf, c, Neff_threshold, 0, flags, nullptr, Neffs, a0, a1,
                          a2, wwa, phase);
}

int wwzb200_kirchner_psd(const double *ts, const double *pd_ys, int64_t nts, const double *taus, int64_t nt,
                         const double *omegas, int64_t nf, double c, int64_t Neff_threshold,
                         int64_t psd_Neff_threshold, uint32_t flags, double *psd, double *Neffs, double *a0,
                         double *a1, double *a2, double *wwa, double *phase)
{
    if (!psd && nf) return fail(WWZB200_ERR_ARG, "psd output is required");
    return host_transform(ts, pd_ys, nts, taus, nt, omegas, nf, c, Neff_threshold, psd_Neff_threshold, flags, psd,
                          Neffs, a0, a1, a2, wwa, phase);
}

int wwzb200_wwa2psd(const double *wwa, const double *Neffs, int64_t nt, int64_t nf, double tspan, int64_t nts,
                    int64_t Neff_threshold, double *psd)
{
    if (nt < 0 || nf < 0 || nts < 1) return fail(WWZB200_ERR_ARG, "bad sizes");
    if (nf == 0) return 0;
    if (!wwa || !Neffs || !psd) return fail(WWZB200_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    int rc;
    if ((rc = workspace(&w))) return rc;
    const int64_t ncell = nt * nf;
    double *d;
    if ((rc = wsbuf(w, "out", (size_t)(6 * ncell + nf), &d))) return rc;
    cudaStream_t st = w->stream;
    if (ncell) {
        CU(cudaMemcpyAsync(d, wwa, sizeof(double) * ncell, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d + ncell, Neffs, sizeof(double) * ncell, cudaMemcpyHostToDevice, st));
    }
    CU(wwz::launch_psd(d, d + ncell, nt, nf, nf, tspan, nts, (double)Neff_threshold, d + 6 * ncell, st));
    CU(cudaMemcpyAsync(psd, d + 6 * ncell, sizeof(double) * nf, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int wwzb200_kirchner_dev(const double *d_ts, const double *d_pd_ys, int64_t nts, const double *d_taus, int64_t nt,
                         const double *d_omegas, int64_t nf, double c, int64_t Neff_threshold,
                         int64_t psd_Neff_threshold, uint32_t flags, double *d_Neffs, double *d_a0, double *d_a1,
                         double *d_a2, double *d_wwa, double *d_phase, double *d_psd, int64_t out_row_stride,
                         void *stream)
{
    int rc = check_grid_args(d_ts, d_pd_ys, nts, d_taus, nt, d_omegas, nf, c, Neff_threshold);
    if (rc) return rc;
    if (!d_Neffs && nt * nf) return fail(WWZB200_ERR_ARG, "d_Neffs output is required");
    if (out_row_stride == 0) out_row_stride = nf;
    if (out_row_stride < nf) return fail(WWZB200_ERR_ARG, "out_row_stride < nf");
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    if ((rc = workspace(&w))) return rc;
    return transform_device(w, d_ts, d_pd_ys, nts, d_taus, nt, d_omegas, nf, c, Neff_threshold, psd_Neff_threshold,
                            flags, d_Neffs, d_a0, d_a1, d_a2, d_wwa, d_phase, d_psd, out_row_stride, -1.0,
                            static_cast<cudaStream_t>(stream));
}

int wwzb200_wwa2psd_dev(const double *d_wwa, const double *d_Neffs, int64_t nt, int64_t nf, double tspan,
                        int64_t nts, int64_t Neff_threshold, double *d_psd, void *stream)
{
    if (nt < 0 || nf < 0 || nts < 1) return fail(WWZB200_ERR_ARG, "bad sizes");
    if (nf == 0) return 0;
    if (!d_wwa || !d_Neffs || !d_psd) return fail(WWZB200_ERR_ARG, "null pointer");
    CU(wwz::launch_psd(d_wwa, d_Neffs, nt, nf, nf, tspan, nts, (double)Neff_threshold, d_psd,
                       static_cast<cudaStream_t>(stream)));
    return 0;
}

int wwzb200_measure_fp64_peak(double *tflops, double *ms)
{
    if (!tflops || !ms) return fail(WWZB200_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lock(g_mu);
    CU(wwz::run_fp64_peak(tflops, ms));
    return 0;
}

int64_t wwzb200_launch_count(void) { return wwz::launch_counter(); }

int wwzb200_profile(int enable)
{
    std::lock_guard<std::mutex> lock(g_mu);
    g_profile = enable != 0;
    for (auto &e : g_events) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    g_events.clear();
    return 0;
}

int wwzb200_profile_read(double *cell_kernel_ms, int64_t *cell_kernel_launches, double *cell_points)
{
    if (!cell_kernel_ms || !cell_kernel_launches || !cell_points) return fail(WWZB200_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lock(g_mu);
    double ms = 0, cp = 0;
    for (auto &e : g_events) {
        CU(cudaEventSynchronize(e.b));
        float t = 0;
        CU(cudaEventElapsedTime(&t, e.a, e.b));
        ms += t;
        cp += e.cell_points;
    }
    *cell_kernel_ms = ms;
    *cell_kernel_launches = (int64_t)g_events.size();
    *cell_points = cp;
    return 0;
}

}  // extern "C"
