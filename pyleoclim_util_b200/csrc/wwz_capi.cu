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

// Counts the CUDA failures of the process: scratch that relies on a launch leaving its device counters zeroed (the
// fused epilogue) is cleared again after any of them.
unsigned int g_cuda_failures = 0;

int fail(int code, const char *fmt, ...)
{
    if (code == WWZB200_ERR_CUDA) ++g_cuda_failures;
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
constexpr int kLanes = 8;   // concurrent streams of the ragged batch, each with its own scratch set

struct Workspace {
    std::map<std::string, std::pair<void *, size_t>> bufs;
    cudaStream_t stream = nullptr;
    cudaStream_t lanes[kLanes] = {};
    cudaEvent_t lane_ev[kLanes] = {};
    cudaEvent_t fork_ev = nullptr;
    cudaStream_t side = nullptr;          // the direct cell kernel runs here, next to the recurrence kernel
    cudaEvent_t side_fork = nullptr, side_join = nullptr;
    cudaStream_t copy = nullptr;          // device-to-host copies of finished member groups, under the next group's kernels
    cudaEvent_t group_ev[2] = {};
    cudaEvent_t busy_ev = nullptr;        // end of the last piece of work that used this scratch set (any stream)
    cudaStream_t busy_stream = nullptr;
    bool busy = false;
    // the y-independent pass of a shared-axis batch that wwzb200_shared_axis_prepare_dev left in the scratch
    struct Prepared {
        bool valid = false;
        const void *ts = nullptr, *tau = nullptr, *omega = nullptr;
        int64_t n = 0, nt = 0, nf = 0, S_pad = 0;
        double c = 0;
        int64_t thr = 0;
        uint32_t flags = 0;
    } prepared;
    std::string prefix;   // scratch namespace of the stream currently being enqueued on
    int sms = 0;
    bool fuse_zeroed = false;         // counters of the fused epilogue ("tile_cnt", "blk_cnt") have been cleared
    int64_t fuse_zeroed_tiles = 0;
    unsigned int fuse_zeroed_gen = 0;
    double *h_in = nullptr;   // page-locked staging of the single-series inputs
    size_t h_in_count = 0;
    unsigned int *h_progress = nullptr;   // page-locked, mapped: progress flags of the fused epilogue + 2 summary words
    unsigned int progress_epoch = 0;

    int get(const char *name, size_t bytes, void **out)
    {
        auto &b = bufs[prefix + name];
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
        if (h_in) cudaFreeHost(h_in);
        h_in = nullptr;
        h_in_count = 0;
        if (h_progress) cudaFreeHost(h_progress);
        h_progress = nullptr;
        fuse_zeroed = false;
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
        CU(cudaEventCreateWithFlags(&w.fork_ev, cudaEventDisableTiming));
        CU(cudaStreamCreateWithFlags(&w.side, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&w.side_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&w.side_join, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&w.busy_ev, cudaEventDisableTiming));
        CU(cudaStreamCreateWithFlags(&w.copy, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&w.group_ev[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&w.group_ev[1], cudaEventDisableTiming));
        for (int i = 0; i < kLanes; ++i) {
            CU(cudaStreamCreateWithFlags(&w.lanes[i], cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&w.lane_ev[i], cudaEventDisableTiming));
        }
    }
    *out = &w;
    return 0;
}

// The scratch set of a device is shared by every entry point, whatever stream the caller passes: work enqueued on
// `st` first waits for the last work that used the scratch (if that was enqueued on another stream), and leaves its
// own end marker behind.  Calls on one stream keep their stream order for free; calls on different streams are
// serialised on the device instead of racing on kappa / basis / sums / ... (include/wwz_b200.h, "Streams").
int scratch_begin(Workspace *w, cudaStream_t st)
{
    if (w->busy && w->busy_stream != st) CU(cudaStreamWaitEvent(st, w->busy_ev, 0));
    return 0;
}

int scratch_end(Workspace *w, cudaStream_t st)
{
    CU(cudaEventRecord(w->busy_ev, st));
    w->busy_stream = st;
    w->busy = true;
    return 0;
}

// begin() on entry, end marker on every way out (also the error returns)
struct ScratchScope {
    Workspace *w;
    cudaStream_t st;
    ScratchScope(Workspace *w_, cudaStream_t st_) : w(w_), st(st_) {}
    int begin() { return scratch_begin(w, st); }
    ~ScratchScope() { scratch_end(w, st); }
};

template <typename T>
int wsbuf(Workspace *w, const char *name, size_t count, T **out)
{
    void *p = nullptr;
    int rc = w->get(name, count * sizeof(T), &p);
    *out = static_cast<T *>(p);
    return rc;
}

// Bytes of basis rows kept at once: grids whose rows need more are transformed in frequency batches.  8 GiB of the 180 GB:
// config 5 (6.5 GB of rows) then runs as ONE launch of the cell kernels instead of four (250.6 -> 241.2 ms: three ramps
// and tails fewer), and a sharded run of it (a rank's launches are short) gains more.
size_t basis_budget_bytes()
{
    const char *e = getenv("WWZB200_BASIS_BYTES");
    if (e && *e) {
        double v = atof(e);
        if (v >= 1 << 20) return (size_t)v;
    }
    return (size_t)8 << 30;
}

// Number of point-axis segments.  CTA-wide direct kernel: enough that (cells x segments) fills the machine for several
// waves -- 1.5e6 cell-slots (10 waves x 296 CTAs x 512 cells: measured best on config 2, dense).  Warp-tile kernels
// (persistent, work items = (warp tile, segment) pulled from a queue, longest first): the largest items are the
// low-frequency tiles that walk their whole segment, and the launch ends with its last item, so there should be a few
// dozen items per resident warp (measured: a 256-tau x 2047-f slab of config 5 takes 17.4 ms per launch with 2
// segments and 12.8 with 16) -- but every segment costs a slab of partial sums that the epilogue reads back, so a
// segment is at least 1000 points on large grids (config 2: 2 segments; 3 cost 7% on its c = 1e-3 transform), 256 on small ones.  The count depends on
// the problem size and on the flags only, never on the tile geometry, so every variant of a kernel adds the same
// partial sums.
int point_segments(int64_t n, int64_t nt, int64_t nf, bool recurrence, int64_t members = 1)
{
    const int64_t ncell = nt * nf;
    const char *e = getenv("WWZB200_POINT_SEGMENTS");
    int64_t s;
    if (e && *e) {
        s = atoll(e);
    } else if (recurrence) {
        const int64_t tiles = ((nt + 63) / 64) * ((std::min<int64_t>(nf, 65535) + 3) / 4) * members;   // 64 taus x 4 rows per warp tile
        s = (65536 + tiles - 1) / std::max<int64_t>(tiles, 1);
        s = std::min<int64_t>(s, n / (ncell >= 100000 ? 1000 : 256));   // (small grids: the slabs cost nothing)
    } else {
        s = (1500000 + ncell - 1) / std::max<int64_t>(ncell, 1);
        s = std::min<int64_t>(s, n / 256);
    }
    return (int)std::max<int64_t>(1, std::min<int64_t>(s, 4096));
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
// Progress reports of a transform whose epilogue is fused into the recurrence kernel: the tau axis in blocks of whole
// tiles; flag[b] (page-locked, mapped host memory) becomes `epoch` once every cell of block b is final in device memory.
struct Progress {
    volatile unsigned int *flag = nullptr;   // in: kMaxProgressBlocks entries
    unsigned int epoch = 0;                  // in
    int blocks = 0;                          // out: number of blocks in use (0: the transform does not report)
    int64_t block_taus = 0;                  // out: taus per block
};
constexpr int kMaxProgressBlocks = 32;

int transform_device(Workspace *w, const double *d_ts, const double *d_ys, int64_t n, const double *d_tau,
                     int64_t nt, const double *d_omega, int64_t nf, double c, int64_t thr, int64_t psd_thr,
                     uint32_t flags, double *d_neffs, double *d_a0, double *d_a1, double *d_a2, double *d_wwa,
                     double *d_phase, double *d_psd, int64_t row_stride, double tspan, cudaStream_t st,
                     Progress *progress = nullptr)
{
    using namespace wwz;
    const int64_t ncell = nt * nf;
    if (ncell == 0) return 0;
    if (ncell >= ((int64_t)1 << 32)) return fail(WWZB200_ERR_ARG, "nt*nf = %lld exceeds 2^32 cells", (long long)ncell);
    w->prepared.valid = false;   // (this call reuses the scratch a prepared shared-axis batch would rely on)

    double *d_kappa, *d_dcut, *d_sums;
    int *d_unsorted;
    unsigned int *d_fix_count, *d_fix_list;
    int rc;
    if ((rc = wsbuf(w, "kappa", (size_t)nf, &d_kappa))) return rc;
    if ((rc = wsbuf(w, "dcut", (size_t)nf, &d_dcut))) return rc;
    if ((rc = wsbuf(w, "flags", wwz::kFlagInts, &d_unsorted))) return rc;
    d_fix_count = reinterpret_cast<unsigned int *>(d_unsorted + 1);
    if ((rc = wsbuf(w, "fix_list", (size_t)ncell, &d_fix_list))) return rc;

    if (d_psd && !d_wwa) {
        if ((rc = wsbuf(w, "psd_wwa", (size_t)ncell, &d_wwa))) return rc;
        if (row_stride != nf) return fail(WWZB200_ERR_ARG, "psd without wwa output needs row_stride == nf");
    }

    const bool foster = (flags & WWZB200_FOSTER) != 0;
    if (flags & WWZB200_FAITHFUL) {
        if (foster)
            CU(launch_foster_faithful(d_ts, d_ys, n, d_tau, d_omega, nt, nf, c, (double)thr, nullptr, nullptr, 0, d_neffs,
                                      d_a0, d_a1, d_a2, d_wwa, d_phase, row_stride, w->sms, st));
        else
            CU(launch_faithful(d_ts, d_ys, n, d_tau, d_omega, nt, nf, c, (double)thr, nullptr, nullptr, 0, d_neffs, d_a0,
                               d_a1, d_a2, d_wwa, d_phase, row_stride, w->sms, st));
    } else {
        // rows with kappa4*dtau <= 8 on an equally spaced tau axis go to the tau-recurrence kernel; the
        // decision is taken per row on the device (freq_setup), both flavours are launched over all rows
        // and every CTA keeps only the rows that are its own
        bool allow_rec = !(flags & (WWZB200_DENSE | WWZB200_NO_RECURRENCE)) && nt >= 2 * kRecCells;
        {
            const char *e = getenv("WWZB200_NO_RECURRENCE");
            if (e && *e && atoi(e)) allow_rec = false;
        }
        const bool dense = (flags & WWZB200_DENSE) != 0;
        // small tau grids (the reference default is 50 taus): every row goes to a warp-tile kernel with tau-compact
        // tiles -- the recurrence flavour where it applies, the direct flavour elsewhere -- because a tile of the
        // CTA-wide kernel would span the whole tau axis and its point window the whole series
        bool tile_direct = !dense && nt <= kSmallTauGrid;
        {
            const char *e = getenv("WWZB200_NO_TILE_DIRECT");
            if (e && *e && atoi(e)) tile_direct = false;
        }
        int segs = point_segments(n, nt, nf, allow_rec);   // (not on tile_direct: windowed == dense bit for bit)
        // progress reports (streamed output): point segments of ~250 points, so that a work item is short against the
        // launch and whole tau blocks are final early (config 2: a 1000-point item takes 0.2-0.5 ms of a 0.55 ms launch)
        const bool want_progress = progress && progress->flag && allow_rec && !foster && w->prefix.empty();
        if (want_progress) {
            const char *e = getenv("WWZB200_POINT_SEGMENTS");
            if (!(e && *e)) segs = (int)std::max<int64_t>(segs, std::min<int64_t>(8, n / 250));
        }
        const int64_t seg_len = (((n + segs - 1) / segs + 15) / 16) * 16;
        CellPlan plan = plan_cells(n, nt, nf, cells_per_lane_hint(nt, nf, w->sms), seg_len);
        unsigned long long *d_walked = nullptr;   // profiling: warp-points evaluated by the recurrence kernel
        if (g_profile && (rc = wsbuf(w, "walked", 1, &d_walked))) return rc;
        RecPlan rplan{}, dplan{};
        // tabulated rho growth (one exp2 chain per point instead of two) where the windows are long: long series on
        // long tau axes (config 5: 266 -> 250 ms; config 2, n = 2000: 1.4%, left on the exact pair)
        bool rho_growth = nt >= 128 && n >= 8192 && w->prefix.empty();
        {
            const char *e = getenv("WWZB200_RHO_GROWTH");
            if (e && *e) rho_growth = atoi(e) != 0 && w->prefix.empty();
        }
        if (allow_rec) rplan = plan_rec(n, nt, nf, seg_len, kRecCells, rho_growth);
        if (tile_direct) dplan = plan_rec(n, nt, nf, seg_len, kTileDirectCells);
        {   // one row stride of the basis scratch for all kernels of the transform
            int64_t np = tile_direct ? dplan.n_pad : plan.n_pad;
            if (allow_rec) np = std::max(np, rplan.n_pad);
            plan.n_pad = rplan.n_pad = dplan.n_pad = np;
        }
        const double cut_bits = (flags & WWZB200_EXACT_WINDOW) ? 1081.0 : 100.0;
        double *d_grid;
        unsigned char *d_rec_row;
        if ((rc = wsbuf(w, "grid", 2, &d_grid))) return rc;
        if ((rc = wsbuf(w, "rec_row", (size_t)nf, &d_rec_row))) return rc;
        double *d_gap;
        if ((rc = wsbuf(w, "taugap", (size_t)nt, &d_gap))) return rc;
        size_t per_row = (size_t)plan.n_pad * 32;
        int64_t nf_batch = (int64_t)std::max<size_t>(1, basis_budget_bytes() / per_row);
        nf_batch = std::min<int64_t>(std::min<int64_t>(nf_batch, nf), 65535);
        double2 *d_ty, *d_basis;
        if ((rc = wsbuf(w, "ty", (size_t)plan.n_pad, &d_ty))) return rc;
        // (a device that cannot give the whole budget -- other tenants -- gets smaller frequency batches instead of an error)
        while ((rc = wsbuf(w, "basis", (size_t)nf_batch * (size_t)plan.n_pad * 2, &d_basis)) == WWZB200_ERR_NOMEM &&
               nf_batch > 64) {
            cudaGetLastError();
            nf_batch = (nf_batch + 1) / 2;
        }
        if (rc) return rc;
        double *d_estep = nullptr;    // recurrence flavour: growth of rho from point to point, per (row, point)
        if (allow_rec && rplan.growth && (rc = wsbuf(w, "estep", (size_t)rec_estep_count(rplan, nf_batch, 1), &d_estep))) return rc;
        if ((rc = wsbuf(w, "sums", (size_t)segs * (size_t)kNumSums * (size_t)ncell, &d_sums))) return rc;
        double *d_sums2 = nullptr;   // Foster: sums of the second pass (y := cos)
        if (foster && (rc = wsbuf(w, "sums2", (size_t)segs * (size_t)kNumSums * (size_t)ncell, &d_sums2))) return rc;

        // tspan < 0 with a psd request: the span of the time axis is taken on the device, in the same launch
        double *d_mm = nullptr;
        if (d_psd && tspan < 0 && (rc = wsbuf(w, "minmax", 2, &d_mm))) return rc;
        CU(launch_setup(d_omega, nf, d_tau, nt, d_ts, n, c, allow_rec, cut_bits, d_kappa, d_dcut, d_grid, d_rec_row, d_gap,
                        d_unsorted, d_mm, st));   // also zeroes the unsorted flag and d_fix_count (adjacent ints)
        unsigned int *d_queue = reinterpret_cast<unsigned int *>(d_unsorted) + 2;   // zeroed by launch_setup
        // Fused epilogue: with a single frequency batch (and Kirchner's projection) the recurrence kernel finishes its
        // own cells -- the warp that completes the last point segment of a tile adds the slabs and writes the outputs --
        // and finalize_kernel is left with the rows of the direct flavour.
        // It costs the kernel ~4% (measured on config 2) and saves the epilogue launch about half of that, so it is
        // used where it pays: when finished tau blocks are streamed to the host (WWZB200_FUSED_EPILOGUE=1 forces it,
        // =0 forbids it).
        bool fuse_on = want_progress && nf_batch == nf;
        {
            const char *e = getenv("WWZB200_FUSED_EPILOGUE");
            if (e && *e) fuse_on = atoi(e) != 0 && allow_rec && !foster && nf_batch == nf && w->prefix.empty();
        }
        RecFuse fuse{};
        if (fuse_on) {
            unsigned int *d_tile_cnt, *d_blk_cnt;
            if ((rc = wsbuf(w, "tile_cnt", (size_t)rec_tile_count(rplan, nf, 1), &d_tile_cnt))) return rc;
            if ((rc = wsbuf(w, "blk_cnt", (size_t)kMaxProgressBlocks, &d_blk_cnt))) return rc;
            if (!w->fuse_zeroed || w->fuse_zeroed_tiles < rec_tile_count(rplan, nf, 1) || w->fuse_zeroed_gen != g_cuda_failures) {
                // the counters are left zeroed by every launch: cleared when the buffer is new (or has grown)
                CU(cudaMemsetAsync(d_tile_cnt, 0, sizeof(unsigned int) * (size_t)rec_tile_count(rplan, nf, 1), st));
                CU(cudaMemsetAsync(d_blk_cnt, 0, sizeof(unsigned int) * kMaxProgressBlocks, st));
                w->fuse_zeroed = true;
                w->fuse_zeroed_tiles = rec_tile_count(rplan, nf, 1);
                w->fuse_zeroed_gen = g_cuda_failures;
            }
            fuse.omega = d_omega;
            fuse.m_omega = 0;
            fuse.row_stride = row_stride;
            fuse.thr = (double)thr;
            fuse.neffs = d_neffs;
            fuse.a0 = d_a0;
            fuse.a1 = d_a1;
            fuse.a2 = d_a2;
            fuse.wwa = d_wwa;
            fuse.phase = d_phase;
            fuse.tile_cnt = d_tile_cnt;
            fuse.fix_count = d_fix_count;
            fuse.fix_list = d_fix_list;
            fuse.fix_cap = (unsigned int)ncell;
            fuse.unfused = reinterpret_cast<unsigned int *>(d_unsorted) + 6;   // zeroed by launch_setup
            if (want_progress && fuse_on) {
                // tau blocks of whole tiles, at least ~1.5 MB of output each
                const int tiles_g = rec_tau_tiles(rplan), tile_taus = rec_tile_taus(rplan);
                int want = (int)std::min<int64_t>(16, std::max<int64_t>(1, 6 * ncell * 8 / (3 << 19)));
                {
                    const char *e = getenv("WWZB200_PROGRESS_BLOCKS");
                    if (e && *e) want = std::max(1, std::min(kMaxProgressBlocks, atoi(e)));
                }
                const int blk_gt = std::max(1, (tiles_g + want - 1) / want);
                fuse.blk_cnt = d_blk_cnt;
                fuse.blk_flag = progress->flag;
                fuse.blk_epoch = progress->epoch;
                fuse.blk_gt = blk_gt;
                progress->blocks = (tiles_g + blk_gt - 1) / blk_gt;
                progress->block_taus = (int64_t)blk_gt * tile_taus;
            }
        }
        for (int64_t k0 = 0; k0 < nf; k0 += nf_batch) {
            const int64_t nb = std::min<int64_t>(nf_batch, nf - k0);
            CU(launch_basis(d_ts, d_ys, n, plan.n_pad, d_omega + k0, nb, d_ty, d_basis, k0 == 0, false, st));
            EventPair ev{nullptr, nullptr, (double)n * (double)nt * (double)nb};
            if (g_profile) {
                CU(cudaEventCreate(&ev.a));
                CU(cudaEventCreate(&ev.b));
                CU(cudaEventRecord(ev.a, st));
            }
            // every frequency row belongs to one flavour (decided on the device), the two kernels write disjoint
            // cells: they run side by side, so a launch that finds no row of its own costs nothing
            const bool fork = allow_rec && w->prefix.empty();   // (the ragged batch already runs one stream per member)
            cudaStream_t st_direct = fork ? w->side : st;
            if (fork) {
                CU(cudaEventRecord(w->side_fork, st));
                CU(cudaStreamWaitEvent(w->side, w->side_fork, 0));
            }
            if (allow_rec)
                CU(launch_rec(rplan, d_ty, d_basis, d_ts, n, d_tau, nt, d_kappa, d_dcut, d_gap, d_unsorted, k0, nb, nf, segs,
                              seg_len, d_rec_row, d_grid, d_sums, d_queue, d_walked, w->sms, st, 1, 0, 0,
                              fuse_on ? &fuse : nullptr, d_estep));
            if (tile_direct)
                CU(launch_rec(dplan, d_ty, d_basis, d_ts, n, d_tau, nt, d_kappa, d_dcut, d_gap, d_unsorted, k0, nb, nf, segs,
                              seg_len, allow_rec ? d_rec_row : nullptr, d_grid, d_sums, d_queue + 2, nullptr, w->sms, st_direct));
            else
                CU(launch_cells(plan, d_ty, d_basis, d_ts, n, d_tau, nt, d_kappa, d_dcut, d_gap, d_unsorted, k0, nb, nf, dense, segs,
                                seg_len, allow_rec ? d_rec_row : nullptr, d_grid, d_sums, st_direct));
            if (fork) {
                CU(cudaEventRecord(w->side_join, w->side));
                CU(cudaStreamWaitEvent(st, w->side_join, 0));
            }
            if (g_profile) {
                CU(cudaEventRecord(ev.b, st));
                g_events.push_back(ev);
            }
            if (foster) {
                // second pass of the same kernels over records built with y := cos(omega t): its "y sin" and
                // "y cos" sums are sum(w c s) and sum(w c c), the two extra entries of Foster's normal matrix
                CU(launch_basis(d_ts, d_ys, n, plan.n_pad, d_omega + k0, nb, d_ty, d_basis, false, true, st));
                if (fork) {
                    CU(cudaEventRecord(w->side_fork, st));
                    CU(cudaStreamWaitEvent(w->side, w->side_fork, 0));
                }
                if (allow_rec)
                    CU(launch_rec(rplan, d_ty, d_basis, d_ts, n, d_tau, nt, d_kappa, d_dcut, d_gap, d_unsorted, k0, nb, nf, segs,
                                  seg_len, d_rec_row, d_grid, d_sums2, d_queue, nullptr, w->sms, st, 1, 0, 0, nullptr, d_estep));
                if (tile_direct)
                    CU(launch_rec(dplan, d_ty, d_basis, d_ts, n, d_tau, nt, d_kappa, d_dcut, d_gap, d_unsorted, k0, nb, nf,
                                  segs, seg_len, allow_rec ? d_rec_row : nullptr, d_grid, d_sums2, d_queue + 2, nullptr, w->sms,
                                  st_direct));
                else
                    CU(launch_cells(plan, d_ty, d_basis, d_ts, n, d_tau, nt, d_kappa, d_dcut, d_gap, d_unsorted, k0, nb, nf, dense,
                                    segs, seg_len, allow_rec ? d_rec_row : nullptr, d_grid, d_sums2, st_direct));
                if (fork) {
                    CU(cudaEventRecord(w->side_join, w->side));
                    CU(cudaStreamWaitEvent(st, w->side_join, 0));
                }
            }
        }
        if (foster) {
            CU(launch_finalize_foster(d_sums, d_sums2, d_tau, d_omega, nt, nf, (double)thr, d_neffs, d_a0, d_a1, d_a2,
                                      d_wwa, d_phase, row_stride, d_fix_count, d_fix_list, (unsigned int)ncell, segs, st));
            CU(launch_foster_faithful(d_ts, d_ys, n, d_tau, d_omega, nt, nf, c, (double)thr, d_fix_count, d_fix_list,
                                      (unsigned int)ncell, d_neffs, d_a0, d_a1, d_a2, d_wwa, d_phase, row_stride,
                                      w->sms, st));
        } else {
            CU(launch_finalize(d_sums, d_tau, d_omega, nt, nf, (double)thr, d_neffs, d_a0, d_a1, d_a2, d_wwa, d_phase,
                               row_stride, d_fix_count, d_fix_list, (unsigned int)ncell, segs, st,
                               fuse_on ? d_rec_row : nullptr));
            CU(launch_faithful(d_ts, d_ys, n, d_tau, d_omega, nt, nf, c, (double)thr, d_fix_count, d_fix_list,
                               (unsigned int)ncell, d_neffs, d_a0, d_a1, d_a2, d_wwa, d_phase, row_stride, w->sms, st));
        }
    }
    if (d_psd) {
        if (tspan >= 0) {
            CU(launch_psd(d_wwa, d_neffs, nt, nf, row_stride, tspan, n, (double)psd_thr, d_psd, st));
        } else {
            double *d_mm;
            if ((rc = wsbuf(w, "minmax", 2, &d_mm))) return rc;
            if (flags & WWZB200_FAITHFUL) CU(launch_minmax(d_ts, n, d_mm, st));   // (otherwise filled by launch_setup)
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
    ScratchScope scope(w, st);
    if ((rc = scope.begin())) return rc;
    const int64_t ncell = nt * nf;
    double *d_in, *d_out;
    const size_t in_count = (size_t)(2 * n + nt + nf);
    if ((rc = wsbuf(w, "in", in_count, &d_in))) return rc;
    if ((rc = wsbuf(w, "out", (size_t)(6 * ncell + nf), &d_out))) return rc;
    double *d_ts = d_in, *d_ys = d_in + n, *d_tau = d_ys + n, *d_om = d_tau + nt;
    // the four input vectors go up as ONE DMA from a page-locked staging buffer (caller memory is usually pageable:
    // four separate copies would each be staged by the driver)
    if (w->h_in_count < in_count) {
        if (w->h_in) cudaFreeHost(w->h_in);
        w->h_in = nullptr;
        w->h_in_count = 0;
        CU(cudaMallocHost(&w->h_in, sizeof(double) * std::max<size_t>(in_count, 4096)));
        w->h_in_count = std::max<size_t>(in_count, 4096);
    }
    memcpy(w->h_in, ts, sizeof(double) * n);
    memcpy(w->h_in + n, ys, sizeof(double) * n);
    if (nt) memcpy(w->h_in + 2 * n, tau, sizeof(double) * nt);
    if (nf) memcpy(w->h_in + 2 * n + nt, om, sizeof(double) * nf);
    CU(cudaMemcpyAsync(d_in, w->h_in, sizeof(double) * in_count, cudaMemcpyHostToDevice, st));
    double *o_neff = d_out, *o_a0 = d_out + ncell, *o_a1 = d_out + 2 * ncell, *o_a2 = d_out + 3 * ncell;
    double *o_wwa = d_out + 4 * ncell, *o_ph = d_out + 5 * ncell, *o_psd = d_out + 6 * ncell;
    double tspan = 0;
    if (psd) {
        const auto mm = std::minmax_element(ts, ts + n);
        tspan = *mm.second - *mm.first;   // utils/wavelet.py:1270 (np.max(ts)-np.min(ts))
    }
    const size_t cb = sizeof(double) * (size_t)ncell;
    // six outputs at one pitch in caller memory (one [6, nt, nf] block, as the Python layer allocates them, or the tau
    // rows of a rank inside a larger [6, NT, nf] block): one (2-D) DMA instead of six
    const ptrdiff_t pitch = (Neffs && a0) ? a0 - Neffs : 0;
    const bool one_block = ncell && Neffs && a0 && a1 && a2 && wwa && phase && pitch >= (ptrdiff_t)ncell &&
                           a1 - a0 == pitch && a2 - a1 == pitch && wwa - a2 == pitch && phase - wwa == pitch;
    // Streaming: when the recurrence kernel finishes its own cells (fused epilogue), finished blocks of tau rows leave
    // for the host on the copy engine while later tiles are still being computed.  The kernel reports a block through
    // a flag in page-locked host memory; this thread polls the flags and enqueues the copies.
    Progress pg;
    const int n_out = (Neffs != nullptr) + (a0 != nullptr) + (a1 != nullptr) + (a2 != nullptr) + (wwa != nullptr) + (phase != nullptr);
    bool stream_out = ncell && n_out * cb >= ((size_t)4 << 20);
    {
        const char *e = getenv("WWZB200_NO_STREAMED_OUTPUT");
        if (e && *e && atoi(e)) stream_out = false;
    }
    if (stream_out) {
        if (!w->h_progress) {
            CU(cudaHostAlloc(&w->h_progress, sizeof(unsigned int) * (kMaxProgressBlocks + 2), cudaHostAllocMapped));
            memset(w->h_progress, 0, sizeof(unsigned int) * (kMaxProgressBlocks + 2));
        }
        unsigned int *d_flag = nullptr;
        CU(cudaHostGetDevicePointer(&d_flag, w->h_progress, 0));
        pg.flag = d_flag;
        pg.epoch = ++w->progress_epoch;
        if (pg.epoch == 0) pg.epoch = ++w->progress_epoch;
    }
    rc = transform_device(w, d_ts, d_ys, n, d_tau, nt, d_om, nf, c, thr, psd_thr, flags, o_neff, o_a0, o_a1, o_a2, o_wwa, o_ph,
                          psd ? o_psd : nullptr, nf, tspan, st, stream_out ? &pg : nullptr);
    if (rc) return rc;
    bool copied = false;
    if (stream_out && pg.blocks > 0) {
        // the two words that decide whether the streamed copies are the final ones: tiles the recurrence kernel left to
        // finalize_kernel, cells that faithful_cells_kernel rewrote
        int *d_flags_dev;
        if ((rc = wsbuf(w, "flags", wwz::kFlagInts, &d_flags_dev))) return rc;
        volatile unsigned int *const hp = w->h_progress;
        unsigned int *const h_sum = w->h_progress + kMaxProgressBlocks;
        CU(cudaMemcpyAsync(&h_sum[0], d_flags_dev + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));   // deep-underflow cells
        CU(cudaMemcpyAsync(&h_sum[1], d_flags_dev + 6, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));   // unfused tiles
        bool all_reported = true;
        for (int b = 0; b < pg.blocks; ++b) {
            unsigned int spins = 0;
            bool seen = true;
            while (hp[b] != pg.epoch) {
                if ((++spins & 0x3ff) == 0) {
                    const cudaError_t q = cudaStreamQuery(st);
                    if (q == cudaSuccess) {          // everything has run: the flag is there now, or will never be
                        seen = hp[b] == pg.epoch;
                        break;
                    }
                    if (q != cudaErrorNotReady) CU(q);
                }
            }
            if (!seen) {
                all_reported = false;
                break;
            }
            const int64_t j0 = (int64_t)b * pg.block_taus, rows = std::min<int64_t>(pg.block_taus, nt - j0);
            const size_t bb = sizeof(double) * (size_t)(rows * nf);
            if (one_block) {
                CU(cudaMemcpy2DAsync(Neffs + j0 * nf, sizeof(double) * (size_t)pitch, o_neff + j0 * nf, cb, bb, 6,
                                     cudaMemcpyDeviceToHost, w->copy));
            } else {
                double *const dst[6] = {Neffs, a0, a1, a2, wwa, phase};
                const double *const src[6] = {o_neff, o_a0, o_a1, o_a2, o_wwa, o_ph};
                for (int i = 0; i < 6; ++i)
                    if (dst[i]) CU(cudaMemcpyAsync(dst[i] + j0 * nf, src[i] + j0 * nf, bb, cudaMemcpyDeviceToHost, w->copy));
            }
        }
        CU(cudaStreamSynchronize(st));
        copied = all_reported && h_sum[0] == 0 && h_sum[1] == 0;
        CU(cudaStreamSynchronize(w->copy));
    }
    if (copied) {
        // every block went out final
    } else if (one_block) {
        CU(cudaMemcpy2DAsync(Neffs, sizeof(double) * (size_t)pitch, o_neff, cb, cb, 6, cudaMemcpyDeviceToHost, st));
    } else if (ncell) {
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

// ---------------------------------------------------------------------------------------------
// Shared-time-axis batch on device-resident inputs.  d_Y is [S][n] row-major.
// Outputs: d_neffs [nt*nf]; d_wwa and the optional d_phase/d_a0/d_a1/d_a2 are [S][nt*nf];
// d_psd [S][nf].  amp_cells_out / S_pad_out (optional) receive the cell-major amplitudes
// [ncell][S_pad] left in the workspace (the layout the quantile kernel consumes).
// ---------------------------------------------------------------------------------------------
int shared_axis_device(Workspace *w, const double *d_ts, const double *d_Y, int64_t n, int64_t S, const double *d_tau,
                       int64_t nt, const double *d_omega, int64_t nf, double c, int64_t thr, int64_t psd_thr,
                       uint32_t flags, double *d_neffs, double *d_wwa, double *d_phase, double *d_a0, double *d_a1,
                       double *d_a2, double *d_psd, double tspan, double **amp_cells_out, int64_t *S_pad_out,
                       cudaStream_t st, double *d_amp_cells_ext = nullptr, int phase = 3,
                       const wwz::AmpScatter *scatter = nullptr)
{
    // phase 1: only what does not depend on the series (setup, basis rows, rotations, stored weights + Neffs);
    // phase 2: the rest, on a scratch that a phase-1 call with the same arguments has prepared; 3: everything
    using namespace wwz;
    const int64_t ncell = nt * nf;
    if (ncell == 0 || S == 0) return 0;
    if (ncell >= ((int64_t)1 << 32)) return fail(WWZB200_ERR_ARG, "nt*nf exceeds 2^32 cells");
    if (flags & WWZB200_FOSTER)
        return fail(WWZB200_ERR_ARG, "WWZB200_FOSTER is not available for the shared-time-axis batch (the 3x3 solve is "
                                     "not linear in the basis sums alone); use the single-series or ragged entry points");
    const SharedPlan plan = plan_shared(n, nt, nf, S);
    const int NS = plan.series_block;
    const int64_t S_pad = ((S + NS - 1) / NS) * NS;
    if (S_pad / NS > 65535) return fail(WWZB200_ERR_ARG, "more than 65535*%d series in one batch", NS);
    if (nf > 65535) return fail(WWZB200_ERR_ARG, "the shared-axis batch takes at most 65535 frequencies (one grid row of its "
                                                 "basis and psd kernels per frequency); split the frequency axis");
    if ((size_t)nf * (size_t)plan.n_pad * 16 > ((size_t)16 << 30))
        return fail(WWZB200_ERR_NOMEM, "basis rows of the shared-axis batch exceed 16 GiB");
    int rc;
    double *d_kappa, *d_dcut, *d_mu, *d_sig, *d_Yt, *d_wq, *d_amp_t, *d_ph_t = nullptr, *d_a0_t = nullptr,
           *d_a1_t = nullptr, *d_a2_t = nullptr, *d_ne;
    double2 *d_ty, *d_basis, *d_rot;
    unsigned int *d_fix;
    int *d_unsorted;
    if ((rc = wsbuf(w, "flags", wwz::kFlagInts, &d_unsorted))) return rc;
    if ((rc = wsbuf(w, "kappa", (size_t)nf, &d_kappa))) return rc;
    if ((rc = wsbuf(w, "dcut", (size_t)nf, &d_dcut))) return rc;
    if ((rc = wsbuf(w, "ty", (size_t)plan.n_pad, &d_ty))) return rc;
    if ((rc = wsbuf(w, "basis", (size_t)nf * (size_t)plan.n_pad, &d_basis))) return rc;
    if ((rc = wsbuf(w, "rot", (size_t)ncell, &d_rot))) return rc;
    if ((rc = wsbuf(w, "mu", (size_t)S_pad, &d_mu))) return rc;
    if ((rc = wsbuf(w, "sig", (size_t)S_pad, &d_sig))) return rc;
    if ((rc = wsbuf(w, "Yt", (size_t)S_pad * (size_t)plan.n_pad, &d_Yt))) return rc;
    if ((rc = wsbuf(w, "wq", (size_t)2 * ncell, &d_wq))) return rc;
    if (d_amp_cells_ext)
        d_amp_t = d_amp_cells_ext;       // the caller keeps the cell-major amplitudes [ncell][S_pad]
    else if (phase == 1 || scatter)
        d_amp_t = nullptr;               // (scatter: the amplitudes go to the owners of the cell blocks)
    else if ((rc = wsbuf(w, "amp_t", (size_t)ncell * (size_t)S_pad, &d_amp_t)))
        return rc;
    if (d_phase && (rc = wsbuf(w, "ph_t", (size_t)ncell * (size_t)S_pad, &d_ph_t))) return rc;
    if (d_a0 && (rc = wsbuf(w, "a0_t", (size_t)ncell * (size_t)S_pad, &d_a0_t))) return rc;
    if (d_a1 && (rc = wsbuf(w, "a1_t", (size_t)ncell * (size_t)S_pad, &d_a1_t))) return rc;
    if (d_a2 && (rc = wsbuf(w, "a2_t", (size_t)ncell * (size_t)S_pad, &d_a2_t))) return rc;
    if ((rc = wsbuf(w, "fix_list", (size_t)ncell + 1, &d_fix))) return rc;
    d_ne = d_neffs;
    if (!d_ne && (rc = wsbuf(w, "ne_t", (size_t)ncell, &d_ne))) return rc;
    double *d_wfrag = nullptr, *d_cellsums = nullptr;
    if (plan.stored_weights) {
        if ((rc = wsbuf(w, "wfrag", plan.weights_count, &d_wfrag))) return rc;
        if ((rc = wsbuf(w, "cellsums", (size_t)4 * ncell, &d_cellsums))) return rc;
    }

    double *d_grid;
    unsigned char *d_rec_row;
    if ((rc = wsbuf(w, "grid", 2, &d_grid))) return rc;
    if ((rc = wsbuf(w, "rec_row", (size_t)nf, &d_rec_row))) return rc;
    // point windows of the cell tiles: same relative 2^-100 cut as the single-series path; WWZB200_DENSE walks every point
    const double cut_bits = (flags & WWZB200_DENSE) ? HUGE_VAL : (flags & WWZB200_EXACT_WINDOW) ? 1081.0 : 100.0;
    double *d_gap;
    if ((rc = wsbuf(w, "taugap", (size_t)nt, &d_gap))) return rc;
    Workspace::Prepared key;
    key.valid = true;
    key.ts = d_ts;
    key.tau = d_tau;
    key.omega = d_omega;
    key.n = n;
    key.nt = nt;
    key.nf = nf;
    key.S_pad = S_pad;
    key.c = c;
    key.thr = thr;
    key.flags = flags & (WWZB200_DENSE | WWZB200_EXACT_WINDOW);
    if (phase == 2) {
        const Workspace::Prepared &p = w->prepared;
        if (!(p.valid && p.ts == key.ts && p.tau == key.tau && p.omega == key.omega && p.n == n && p.nt == nt && p.nf == nf &&
              p.S_pad == S_pad && p.c == c && p.thr == thr && p.flags == key.flags && plan.stored_weights))
            return fail(WWZB200_ERR_ARG, "WWZB200_PREPARED: the scratch does not hold the y-independent pass of this batch "
                                         "(call wwzb200_shared_axis_prepare_dev with the same arguments first)");
    }
    w->prepared.valid = false;
    if (phase & 1) {
        CU(launch_setup(d_omega, nf, d_tau, nt, d_ts, n, c, false, cut_bits, d_kappa, d_dcut, d_grid, d_rec_row, d_gap,
                        d_unsorted, nullptr, st));
        CU(launch_basis_sc(d_ts, n, plan.n_pad, d_omega, nf, d_ty, d_basis, st));
        CU(launch_rotation(d_tau, d_omega, nt, nf, d_rot, st));
    }
    if (phase == 1) {
        if (plan.stored_weights) {
            CU(launch_shared(plan, d_ty, d_basis, d_Yt, d_ts, n, S_pad, d_tau, nt, d_kappa, d_dcut, d_gap, d_unsorted, nf, d_rot,
                             (double)thr, d_wfrag, d_cellsums, d_ne, d_wq, d_amp_t, d_ph_t, d_a0_t, d_a1_t, d_a2_t, st, 1));
            w->prepared = key;
        }
        return 0;
    }
    CU(launch_standardize_retile(d_Y, n, plan.n_pad, S, NS, (flags & WWZB200_STANDARDIZE) != 0, d_mu, d_sig, d_Yt, st));
    EventPair ev{nullptr, nullptr, (double)n * (double)ncell * (double)S};
    if (g_profile) {
        CU(cudaEventCreate(&ev.a));
        CU(cudaEventCreate(&ev.b));
        CU(cudaEventRecord(ev.a, st));
    }
    CU(launch_shared(plan, d_ty, d_basis, d_Yt, d_ts, n, S_pad, d_tau, nt, d_kappa, d_dcut, d_gap, d_unsorted, nf, d_rot,
                     (double)thr, d_wfrag, d_cellsums, d_ne, d_wq, d_amp_t, d_ph_t, d_a0_t, d_a1_t, d_a2_t, st, phase, scatter));
    if (g_profile) {
        CU(cudaEventRecord(ev.b, st));
        g_events.push_back(ev);
    }
    CU(launch_shared_fixup(d_ts, d_Yt, n, plan.n_pad, S, S_pad, NS, d_tau, d_omega, nt, nf, c, (double)thr, d_wq, d_fix,
                           d_fix + 1, d_ne, d_amp_t, d_ph_t, d_a0_t, d_a1_t, d_a2_t, w->sms, st, scatter));
    // cell-major [ncell][S_pad] -> series-major [S][ncell]
    if (d_wwa) CU(launch_transpose(d_amp_t, ncell, S, S_pad, d_wwa, ncell, st));
    if (d_phase) CU(launch_transpose(d_ph_t, ncell, S, S_pad, d_phase, ncell, st));
    if (d_a0) CU(launch_transpose(d_a0_t, ncell, S, S_pad, d_a0, ncell, st));
    if (d_a1) CU(launch_transpose(d_a1_t, ncell, S, S_pad, d_a1, ncell, st));
    if (d_a2) CU(launch_transpose(d_a2_t, ncell, S, S_pad, d_a2, ncell, st));
    if (d_psd) CU(launch_psd_batch(d_amp_t, d_ne, nt, nf, S, S_pad, tspan, n, (double)psd_thr, d_psd, st));
    if (amp_cells_out) *amp_cells_out = d_amp_t;
    if (S_pad_out) *S_pad_out = S_pad;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// M members of equal shape in one set of launches (age ensembles: core/ensembleseries.py:47-175 feeding
// core/multipleseries.py:1520-1532): per-member Nyquist frequency and z-score, setup, basis rows, the warp-tile
// kernels over (member, tile, segment) work items -- recurrence flavour where it applies, direct flavour elsewhere --
// epilogue, deep-underflow fix-up and wwa2psd, each as ONE launch for all members.  Inputs by member strides
// (0 = shared), outputs [M][nt][nf] (psd [M][nf]).
// ---------------------------------------------------------------------------------------------
int members_device(Workspace *w, const double *d_ts, int64_t m_ts, const double *d_ys, int64_t m_ys, int64_t M, int64_t n,
                   const double *d_tau, int64_t m_tau, int64_t nt, const double *d_om, int64_t m_om, int64_t nf, double c,
                   int64_t thr, int64_t psd_thr, uint32_t flags, double *d_neffs, double *d_a0, double *d_a1, double *d_a2,
                   double *d_wwa, double *d_phase, double *d_psd, double *d_stats, cudaStream_t st)
{
    using namespace wwz;
    const int64_t ncell = nt * nf;
    if (ncell == 0 || M == 0) return 0;
    if (ncell * M >= ((int64_t)1 << 32)) return fail(WWZB200_ERR_ARG, "M*nt*nf = %lld exceeds 2^32 cells", (long long)(ncell * M));
    if (flags & (WWZB200_FOSTER | WWZB200_FAITHFUL))
        return fail(WWZB200_ERR_ARG, "WWZB200_FOSTER / WWZB200_FAITHFUL are not available for member batches; use the ragged entry point");
    w->prepared.valid = false;
    if (M > 65535 || nf > 65535) return fail(WWZB200_ERR_ARG, "at most 65535 members and frequencies per member batch");
    const bool freq_hz = (flags & WWZB200_FREQ_HZ) != 0, zs = (flags & WWZB200_STANDARDIZE) != 0;
    if (freq_hz && n - 1 > kNyquistMaxPoints)
        return fail(WWZB200_ERR_ARG, "WWZB200_FREQ_HZ needs nts <= %d (pass omega instead)", kNyquistMaxPoints + 1);
    int rc;
    double *d_kappa, *d_dcut, *d_grid, *d_gap, *d_mm, *d_sums, *d_fnyq = nullptr, *d_ysz = nullptr, *d_omega = nullptr;
    unsigned char *d_rec_row;
    int *d_flags;
    unsigned int *d_fix_list;
    double2 *d_ty, *d_basis;
    bool allow_rec = !(flags & (WWZB200_DENSE | WWZB200_NO_RECURRENCE)) && nt >= 2 * kRecCells;
    {
        const char *e = getenv("WWZB200_NO_RECURRENCE");
        if (e && *e && atoi(e)) allow_rec = false;
    }
    const int segs = point_segments(n, nt, nf, true, M);
    const int64_t seg_len = (((n + segs - 1) / segs + 15) / 16) * 16;
    RecPlan rplan{}, dplan = plan_rec(n, nt, nf, seg_len, kTileDirectCells);
    if (allow_rec) {
        rplan = plan_rec(n, nt, nf, seg_len, kRecCells);
        rplan.n_pad = dplan.n_pad = std::max(rplan.n_pad, dplan.n_pad);
    }
    const int64_t n_pad = dplan.n_pad;
    if ((rc = wsbuf(w, "kappa", (size_t)(M * nf), &d_kappa))) return rc;
    if ((rc = wsbuf(w, "dcut", (size_t)(M * nf), &d_dcut))) return rc;
    if ((rc = wsbuf(w, "rec_row", (size_t)(M * nf), &d_rec_row))) return rc;
    if ((rc = wsbuf(w, "grid", (size_t)(2 * M), &d_grid))) return rc;
    if ((rc = wsbuf(w, "taugap", (size_t)(M * nt), &d_gap))) return rc;
    if ((rc = wsbuf(w, "flags", (size_t)(M * kFlagInts), &d_flags))) return rc;
    if ((rc = wsbuf(w, "minmax", (size_t)(2 * M), &d_mm))) return rc;
    if ((rc = wsbuf(w, "fix_list", (size_t)(M * ncell) + 1, &d_fix_list))) return rc;
    if ((rc = wsbuf(w, "ty", (size_t)(M * n_pad), &d_ty))) return rc;
    if ((rc = wsbuf(w, "basis", (size_t)M * (size_t)nf * (size_t)n_pad * 2, &d_basis))) return rc;
    if ((rc = wsbuf(w, "sums", (size_t)M * (size_t)segs * kNumSums * (size_t)ncell, &d_sums))) return rc;
    double *d_estep = nullptr;     // (no tabulated rho growth for member batches: a table per member costs more than it saves)
    if (freq_hz) {
        if ((rc = wsbuf(w, "mb_fnyq", (size_t)M, &d_fnyq))) return rc;
        if ((rc = wsbuf(w, "mb_omega", (size_t)(M * nf), &d_omega))) return rc;
    }
    if (zs && (rc = wsbuf(w, "mb_ysz", (size_t)(M * n), &d_ysz))) return rc;
    if (d_psd && !d_wwa && (rc = wsbuf(w, "psd_wwa", (size_t)(M * ncell), &d_wwa))) return rc;
    unsigned int *d_fix_count = reinterpret_cast<unsigned int *>(d_flags) + 1;   // of member 0: one list for the batch
    unsigned int *d_queue = reinterpret_cast<unsigned int *>(d_flags) + 2;

    if (freq_hz || zs) CU(launch_member_prep(d_ts, m_ts, n, d_ys, m_ys, M, zs, d_fnyq, d_ysz, zs ? d_stats : nullptr, st));
    const double *ys_use = zs ? d_ysz : d_ys;
    const int64_t m_ys_use = zs ? n : m_ys;
    const double cut_bits = (flags & WWZB200_DENSE) ? HUGE_VAL : (flags & WWZB200_EXACT_WINDOW) ? 1081.0 : 100.0;
    CU(launch_setup_members(d_om, m_om, nf, d_tau, m_tau, nt, d_ts, m_ts, n, M, c, allow_rec, cut_bits, d_kappa, d_dcut, d_grid,
                            d_rec_row, d_gap, d_flags, d_psd ? d_mm : nullptr, d_fnyq, d_omega, st));
    const double *om_use = freq_hz ? d_omega : d_om;
    const int64_t m_om_use = freq_hz ? nf : m_om;
    CU(launch_basis_members(d_ts, m_ts, ys_use, m_ys_use, n, n_pad, om_use, m_om_use, nf, M, d_ty, d_basis, st));
    EventPair ev{nullptr, nullptr, (double)n * (double)ncell * (double)M};
    if (g_profile) {
        CU(cudaEventCreate(&ev.a));
        CU(cudaEventCreate(&ev.b));
        CU(cudaEventRecord(ev.a, st));
    }
    // the two flavours write disjoint cells: side by side, as in transform_device
    const bool fork = allow_rec;
    cudaStream_t st_direct = fork ? w->side : st;
    if (fork) {
        CU(cudaEventRecord(w->side_fork, st));
        CU(cudaStreamWaitEvent(w->side, w->side_fork, 0));
        CU(launch_rec(rplan, d_ty, d_basis, d_ts, n, d_tau, nt, d_kappa, d_dcut, d_gap, d_flags, 0, nf, nf, segs, seg_len,
                      d_rec_row, d_grid, d_sums, d_queue, nullptr, w->sms, st, M, m_ts, m_tau, nullptr, d_estep));
    }
    CU(launch_rec(dplan, d_ty, d_basis, d_ts, n, d_tau, nt, d_kappa, d_dcut, d_gap, d_flags, 0, nf, nf, segs, seg_len,
                  allow_rec ? d_rec_row : nullptr, d_grid, d_sums, d_queue + 2, nullptr, w->sms, st_direct, M, m_ts, m_tau));
    if (fork) {
        CU(cudaEventRecord(w->side_join, w->side));
        CU(cudaStreamWaitEvent(st, w->side_join, 0));
    }
    if (g_profile) {
        CU(cudaEventRecord(ev.b, st));
        g_events.push_back(ev);
    }
    CU(launch_finalize_members(d_sums, d_tau, m_tau, om_use, m_om_use, nt, nf, M, (double)thr, d_neffs, d_a0, d_a1, d_a2,
                               d_wwa, d_phase, nf, d_fix_count, d_fix_list, (unsigned int)(M * ncell), segs, st));
    CU(launch_faithful_members(d_ts, m_ts, ys_use, m_ys_use, n, d_tau, m_tau, om_use, m_om_use, nt, nf, M, c, (double)thr,
                               d_fix_count, d_fix_list, (unsigned int)(M * ncell), d_neffs, d_a0, d_a1, d_a2, d_wwa,
                               d_phase, nf, w->sms, st));
    if (d_psd) CU(launch_psd_devspan(d_wwa, d_neffs, nt, nf, nf, d_mm, n, (double)psd_thr, d_psd, st, M));
    return 0;
}

int quantiles_device(Workspace *w, const double *d_Xcells, int64_t S, int64_t stride, int64_t ncell,
                     const double *d_probs, int64_t nq, double *d_q, cudaStream_t st, int64_t seg_len = 0,
                     int64_t seg_stride = 0)
{
    int npow2 = 2;
    while (npow2 < S) npow2 <<= 1;
    unsigned long long *d_scr = nullptr;
    int64_t ctas = 0;
    if ((size_t)npow2 * 8 > 200 * 1024) {
        ctas = (int64_t)w->sms * 2;
        int rc = wsbuf(w, "qscratch", (size_t)ctas * (size_t)npow2, &d_scr);
        if (rc) return rc;
    }
    CU(wwz::launch_quantiles(d_Xcells, S, stride, ncell, d_probs, nq, d_q, d_scr, ctas, w->sms, st, seg_len, seg_stride));
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

int wwzb200_copy_to_host(void *dst, const void *d_src, size_t bytes, void *stream)
{
    if (bytes == 0) return 0;
    if (!dst || !d_src) return fail(WWZB200_ERR_ARG, "null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CU(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int wwzb200_copy_to_host_async(void *dst, const void *d_src, size_t bytes, void *stream)
{
    if (bytes == 0) return 0;
    if (!dst || !d_src) return fail(WWZB200_ERR_ARG, "null pointer");
    CU(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
    return 0;
}

int wwzb200_stream_synchronize(void *stream)
{
    CU(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return 0;
}

int wwzb200_host_register(void *ptr, size_t bytes)
{
    if (!ptr || !bytes) return fail(WWZB200_ERR_ARG, "null pointer");
    CU(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return 0;
}

int wwzb200_host_unregister(void *ptr)
{
    if (ptr) CU(cudaHostUnregister(ptr));
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
    cudaStream_t st = w->stream;
    ScratchScope scope(w, st);
    if ((rc = scope.begin())) return rc;
    if ((rc = wsbuf(w, "out", (size_t)(6 * ncell + nf), &d))) return rc;
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
    ScratchScope scope(w, static_cast<cudaStream_t>(stream));
    if ((rc = scope.begin())) return rc;
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

int wwzb200_wwa2psd_parts_dev(const double *d_wwa, const double *d_Neffs, int64_t nt, int64_t nf, double tspan,
                              int64_t nts, int64_t Neff_threshold, double *d_parts, void *stream)
{
    if (nt < 0 || nf < 0 || nts < 1) return fail(WWZB200_ERR_ARG, "bad sizes");
    if (nf == 0) return 0;
    if (!d_parts || (nt && (!d_wwa || !d_Neffs))) return fail(WWZB200_ERR_ARG, "null pointer");
    if (Neff_threshold < 0) return fail(WWZB200_ERR_ARG, "Neff_threshold must be >= 0");
    CU(wwz::launch_psd_parts(d_wwa, d_Neffs, nt, nf, nf, tspan, nts, (double)Neff_threshold, d_parts,
                             static_cast<cudaStream_t>(stream)));
    return 0;
}

int wwzb200_kirchner_shared_axis_dev(const double *d_ts, const double *d_Y, int64_t nts, int64_t S,
                                     const double *d_taus, int64_t nt, const double *d_omegas, int64_t nf, double c,
                                     int64_t Neff_threshold, int64_t psd_Neff_threshold, uint32_t flags,
                                     double *d_Neffs, double *d_wwa, double *d_phase, double *d_a0, double *d_a1,
                                     double *d_a2, double *d_psd, void *stream)
{
    int rc = check_grid_args(d_ts, d_Y, nts, d_taus, nt, d_omegas, nf, c, Neff_threshold);
    if (rc) return rc;
    if (S < 0) return fail(WWZB200_ERR_ARG, "S must be >= 0");
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    if ((rc = workspace(&w))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScratchScope scope(w, st);
    if ((rc = scope.begin())) return rc;
    double tspan = 0;
    if (d_psd) {   // span of the (device) time axis, needed by wwa2psd
        double *d_mm, mm[2];
        if ((rc = wsbuf(w, "minmax", 2, &d_mm))) return rc;
        CU(wwz::launch_minmax(d_ts, nts, d_mm, st));
        CU(cudaMemcpyAsync(mm, d_mm, sizeof mm, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        tspan = mm[1] - mm[0];
    }
    return shared_axis_device(w, d_ts, d_Y, nts, S, d_taus, nt, d_omegas, nf, c, Neff_threshold, psd_Neff_threshold,
                              flags, d_Neffs, d_wwa, d_phase, d_a0, d_a1, d_a2, d_psd, tspan, nullptr, nullptr, st);
}

int wwzb200_kirchner_shared_axis(const double *ts, const double *Y, int64_t nts, int64_t S, const double *taus,
                                 int64_t nt, const double *omegas, int64_t nf, double c, int64_t Neff_threshold,
                                 int64_t psd_Neff_threshold, uint32_t flags, double *Neffs, double *wwa,
                                 double *phase, double *a0, double *a1, double *a2, double *psd)
{
    int rc = check_grid_args(ts, Y, nts, taus, nt, omegas, nf, c, Neff_threshold);
    if (rc) return rc;
    if (S < 0) return fail(WWZB200_ERR_ARG, "S must be >= 0");
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    if ((rc = workspace(&w))) return rc;
    cudaStream_t st = w->stream;
    ScratchScope scope(w, st);
    if ((rc = scope.begin())) return rc;
    const int64_t ncell = nt * nf;
    double *d_in, *d_Y, *d_ne, *d_psd = nullptr;
    if ((rc = wsbuf(w, "in", (size_t)(2 * nts + nt + nf), &d_in))) return rc;
    if ((rc = wsbuf(w, "Y", (size_t)S * (size_t)nts, &d_Y))) return rc;
    if ((rc = wsbuf(w, "ne_out", (size_t)ncell, &d_ne))) return rc;
    if (psd && (rc = wsbuf(w, "psd_out", (size_t)S * (size_t)nf, &d_psd))) return rc;
    double *d_ts = d_in, *d_tau = d_in + 2 * nts, *d_om = d_tau + nt;
    CU(cudaMemcpyAsync(d_ts, ts, sizeof(double) * nts, cudaMemcpyHostToDevice, st));
    if (nt) CU(cudaMemcpyAsync(d_tau, taus, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
    if (nf) CU(cudaMemcpyAsync(d_om, omegas, sizeof(double) * nf, cudaMemcpyHostToDevice, st));
    if (S) CU(cudaMemcpyAsync(d_Y, Y, sizeof(double) * (size_t)S * (size_t)nts, cudaMemcpyHostToDevice, st));
    const auto mm = std::minmax_element(ts, ts + nts);
    const size_t ob = (size_t)S * (size_t)ncell;
    double *d_o[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    double *h_o[5] = {wwa, phase, a0, a1, a2};
    const char *names[5] = {"o_wwa", "o_ph", "o_a0", "o_a1", "o_a2"};
    for (int i = 0; i < 5; ++i)
        if (h_o[i] && (rc = wsbuf(w, names[i], ob, &d_o[i]))) return rc;
    rc = shared_axis_device(w, d_ts, d_Y, nts, S, d_tau, nt, d_om, nf, c, Neff_threshold, psd_Neff_threshold, flags,
                            d_ne, d_o[0], d_o[1], d_o[2], d_o[3], d_o[4], d_psd, *mm.second - *mm.first, nullptr,
                            nullptr, st);
    if (rc) return rc;
    if (Neffs && ncell) CU(cudaMemcpyAsync(Neffs, d_ne, sizeof(double) * ncell, cudaMemcpyDeviceToHost, st));
    for (int i = 0; i < 5; ++i)
        if (h_o[i] && ob) CU(cudaMemcpyAsync(h_o[i], d_o[i], sizeof(double) * ob, cudaMemcpyDeviceToHost, st));
    if (psd && S * nf) CU(cudaMemcpyAsync(psd, d_psd, sizeof(double) * (size_t)S * nf, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

// (caller holds g_mu and has opened a ScratchScope on the stream)
static int surrogates_nolock(Workspace *w, const double *d_ts, int64_t nts, double tau_est, double sigma, double mu,
                             int64_t S, uint64_t seed, bool uar1, double *d_Y, cudaStream_t st, uint64_t first_series)
{
    double *d_rq;
    int rc;
    if ((rc = wsbuf(w, "ar1_rq", (size_t)2 * nts, &d_rq))) return rc;
    CU(wwz::launch_ar1(d_ts, nts, tau_est, sigma, mu, S, seed, uar1, d_rq, d_rq + nts, d_Y, st, first_series));
    return 0;
}

static int check_surrogate_args(const void *ts, int64_t nts, double tau_est, double sigma, int64_t S, bool uar1,
                                const void *Y)
{
    if (nts < 1 || S < 0) return fail(WWZB200_ERR_ARG, "bad sizes");
    if (!ts || (S && !Y)) return fail(WWZB200_ERR_ARG, "null pointer");
    if (!(tau_est > 0)) return fail(WWZB200_ERR_ARG, "tau must be positive");
    if (uar1 && !(sigma >= 0)) return fail(WWZB200_ERR_ARG, "sigma_2 must be non-negative");
    return 0;
}

static int surrogates_dev(const double *d_ts, int64_t nts, double tau_est, double sigma, double mu, int64_t S,
                          uint64_t seed, bool uar1, double *d_Y, void *stream, uint64_t first_series = 0)
{
    int rc = check_surrogate_args(d_ts, nts, tau_est, sigma, S, uar1, d_Y);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    if ((rc = workspace(&w))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScratchScope scope(w, st);
    if ((rc = scope.begin())) return rc;
    return surrogates_nolock(w, d_ts, nts, tau_est, sigma, mu, S, seed, uar1, d_Y, st, first_series);
}

// Host buffers in and out.  The lock is held from the first workspace request to the final synchronisation: the
// named buffers ("in", "Y") belong to this call until its results have left the device.
static int surrogates_host(const double *ts, int64_t nts, double tau_est, double sigma, double mu, int64_t S,
                           uint64_t seed, bool uar1, double *Y)
{
    int rc = check_surrogate_args(ts, nts, tau_est, sigma, S, uar1, Y);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    double *d_ts, *d_Y;
    if ((rc = workspace(&w))) return rc;
    cudaStream_t st = w->stream;
    ScratchScope scope(w, st);
    if ((rc = scope.begin())) return rc;
    if ((rc = wsbuf(w, "in", (size_t)nts, &d_ts))) return rc;
    if ((rc = wsbuf(w, "Y", (size_t)S * (size_t)nts, &d_Y))) return rc;
    CU(cudaMemcpyAsync(d_ts, ts, sizeof(double) * nts, cudaMemcpyHostToDevice, st));
    if ((rc = surrogates_nolock(w, d_ts, nts, tau_est, sigma, mu, S, seed, uar1, d_Y, st, 0))) return rc;
    if (S) CU(cudaMemcpyAsync(Y, d_Y, sizeof(double) * (size_t)S * (size_t)nts, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int wwzb200_ar1_surrogates_dev(const double *d_ts, int64_t nts, double tau_est, double sigma, double mu, int64_t S,
                               uint64_t seed, double *d_Y, void *stream)
{
    return surrogates_dev(d_ts, nts, tau_est, sigma, mu, S, seed, false, d_Y, stream);
}

int wwzb200_ar1_surrogates(const double *ts, int64_t nts, double tau_est, double sigma, double mu, int64_t S,
                           uint64_t seed, double *Y)
{
    return surrogates_host(ts, nts, tau_est, sigma, mu, S, seed, false, Y);
}

int wwzb200_uar1_surrogates_dev(const double *d_ts, int64_t nts, double tau, double sigma_2, double mu, int64_t S,
                                uint64_t seed, double *d_Y, void *stream)
{
    return surrogates_dev(d_ts, nts, tau, sigma_2, mu, S, seed, true, d_Y, stream);
}

int wwzb200_uar1_surrogates(const double *ts, int64_t nts, double tau, double sigma_2, double mu, int64_t S,
                            uint64_t seed, double *Y)
{
    return surrogates_host(ts, nts, tau, sigma_2, mu, S, seed, true, Y);
}

static int check_quantile_args(const void *X, int64_t S, int64_t ncell, const void *probs, int64_t nq, const void *q)
{
    if (S < 1 || ncell < 0 || nq < 0) return fail(WWZB200_ERR_ARG, "bad sizes");
    if (ncell == 0 || nq == 0) return 0;
    if (!X || !probs || !q) return fail(WWZB200_ERR_ARG, "null pointer");
    if (S > ((int64_t)1 << 24)) return fail(WWZB200_ERR_ARG, "S too large for the quantile kernel");
    return 0;
}

int wwzb200_surrogates_block_dev(const double *d_ts, int64_t nts, double tau, double sigma, double mu, int64_t S,
                                 int64_t first_series, uint64_t seed, uint32_t flags, double *d_Y, void *stream)
{
    if (first_series < 0) return fail(WWZB200_ERR_ARG, "first_series must be >= 0");
    return surrogates_dev(d_ts, nts, tau, sigma, mu, S, seed, (flags & WWZB200_UAR1) != 0, d_Y, stream,
                          (uint64_t)first_series);
}

int wwzb200_series_pad(int64_t S, int64_t *S_pad)
{
    if (S < 0 || !S_pad) return fail(WWZB200_ERR_ARG, "bad arguments");
    const int64_t NS = wwz::shared_series_block(S);
    *S_pad = ((S + NS - 1) / NS) * NS;
    return 0;
}

int wwzb200_shared_axis_cells_dev(const double *d_ts, const double *d_Y, int64_t nts, int64_t S, const double *d_taus,
                                  int64_t nt, const double *d_omegas, int64_t nf, double c, int64_t Neff_threshold,
                                  uint32_t flags, double *d_Neffs, double *d_amp_cells, void *stream)
{
    int rc = check_grid_args(d_ts, d_Y, nts, d_taus, nt, d_omegas, nf, c, Neff_threshold);
    if (rc) return rc;
    if (S < 0 || (S && nt * nf && !d_amp_cells)) return fail(WWZB200_ERR_ARG, "bad batch arguments");
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    if ((rc = workspace(&w))) return rc;
    ScratchScope scope(w, static_cast<cudaStream_t>(stream));
    if ((rc = scope.begin())) return rc;
    return shared_axis_device(w, d_ts, d_Y, nts, S, d_taus, nt, d_omegas, nf, c, Neff_threshold, 0,
                              flags & ~WWZB200_PREPARED, d_Neffs, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0,
                              nullptr, nullptr, static_cast<cudaStream_t>(stream), d_amp_cells,
                              (flags & WWZB200_PREPARED) ? 2 : 3);
}

int wwzb200_shared_axis_cells_scatter_dev(const double *d_ts, const double *d_Y, int64_t nts, int64_t S,
                                          const double *d_taus, int64_t nt, const double *d_omegas, int64_t nf, double c,
                                          int64_t Neff_threshold, uint32_t flags, double *d_Neffs, void *const *peer_bufs,
                                          int64_t n_peers, int64_t cells_per, int64_t dst_stride, int64_t col0,
                                          void *stream)
{
    int rc = check_grid_args(d_ts, d_Y, nts, d_taus, nt, d_omegas, nf, c, Neff_threshold);
    if (rc) return rc;
    int64_t S_pad = 0;
    if (S < 1 || (rc = wwzb200_series_pad(S, &S_pad))) return fail(WWZB200_ERR_ARG, "bad batch arguments");
    if (!peer_bufs || n_peers < 1 || n_peers > wwz::kMaxPeers || cells_per < 1 || cells_per * n_peers < nt * nf ||
        col0 < 0 || dst_stride < col0 + S_pad)
        return fail(WWZB200_ERR_ARG, "bad scatter layout (peers <= %d, cells_per * peers >= nt * nf, dst_stride >= col0 + "
                                     "padded series count)", wwz::kMaxPeers);
    wwz::AmpScatter sc{};
    for (int64_t i = 0; i < n_peers; ++i) {
        if (!peer_bufs[i]) return fail(WWZB200_ERR_ARG, "null peer buffer");
        sc.peer[i] = static_cast<double *>(peer_bufs[i]);
    }
    sc.cells_per = cells_per;
    sc.dst_stride = dst_stride;
    sc.col0 = col0;
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    if ((rc = workspace(&w))) return rc;
    ScratchScope scope(w, static_cast<cudaStream_t>(stream));
    if ((rc = scope.begin())) return rc;
    return shared_axis_device(w, d_ts, d_Y, nts, S, d_taus, nt, d_omegas, nf, c, Neff_threshold, 0,
                              flags & ~WWZB200_PREPARED, d_Neffs, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0,
                              nullptr, nullptr, static_cast<cudaStream_t>(stream), nullptr,
                              (flags & WWZB200_PREPARED) ? 2 : 3, &sc);
}

// Device buffers that the other ranks of the box can map (CUDA IPC): cudaMalloc'ed here, exported as a 64-byte handle.
int wwzb200_ipc_alloc(size_t bytes, void **d_ptr, unsigned char *handle64)
{
    if (!d_ptr || !handle64 || !bytes) return fail(WWZB200_ERR_ARG, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaMalloc(d_ptr, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *d_ptr);
    if (e != cudaSuccess) {
        cudaFree(*d_ptr);
        *d_ptr = nullptr;
        return fail(WWZB200_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(handle64, &h, 64);
    return 0;
}

int wwzb200_ipc_open(const unsigned char *handle64, void **d_ptr)
{
    if (!d_ptr || !handle64) return fail(WWZB200_ERR_ARG, "bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int wwzb200_ipc_close(void *d_ptr)
{
    if (d_ptr) CU(cudaIpcCloseMemHandle(d_ptr));
    return 0;
}

int wwzb200_ipc_free(void *d_ptr)
{
    if (d_ptr) CU(cudaFree(d_ptr));
    return 0;
}

int wwzb200_shared_axis_prepare_dev(const double *d_ts, int64_t nts, int64_t S, const double *d_taus, int64_t nt,
                                    const double *d_omegas, int64_t nf, double c, int64_t Neff_threshold, uint32_t flags,
                                    double *d_Neffs, int *prepared, void *stream)
{
    int rc = check_grid_args(d_ts, d_ts, nts, d_taus, nt, d_omegas, nf, c, Neff_threshold);
    if (rc) return rc;
    if (S < 0 || !prepared) return fail(WWZB200_ERR_ARG, "bad batch arguments");
    *prepared = 0;
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    if ((rc = workspace(&w))) return rc;
    ScratchScope scope(w, static_cast<cudaStream_t>(stream));
    if ((rc = scope.begin())) return rc;
    rc = shared_axis_device(w, d_ts, nullptr, nts, S, d_taus, nt, d_omegas, nf, c, Neff_threshold, 0,
                            flags & ~WWZB200_PREPARED, d_Neffs, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0,
                            nullptr, nullptr, static_cast<cudaStream_t>(stream), nullptr, 1);
    if (rc) return rc;
    *prepared = w->prepared.valid ? 1 : 0;
    return 0;
}

int wwzb200_quantiles_segments_dev(const double *d_X, int64_t S, int64_t seg_len, int64_t seg_stride, int64_t stride,
                                   int64_t ncell, const double *d_probs, int64_t nq, double *d_q, void *stream)
{
    if (seg_len < 1 || stride < seg_len || seg_stride < 0) return fail(WWZB200_ERR_ARG, "bad segment layout");
    int rc = check_quantile_args(d_X, S, ncell, d_probs, nq, d_q);
    if (rc || ncell == 0 || nq == 0) return rc;
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    if ((rc = workspace(&w))) return rc;
    ScratchScope scope(w, static_cast<cudaStream_t>(stream));
    if ((rc = scope.begin())) return rc;
    return quantiles_device(w, d_X, S, stride, ncell, d_probs, nq, d_q, static_cast<cudaStream_t>(stream), seg_len,
                            seg_stride);
}


int wwzb200_quantiles_cells_dev(const double *d_Xcells, int64_t S, int64_t stride, int64_t ncell, const double *d_probs,
                                int64_t nq, double *d_q, void *stream)
{
    if (stride < S) return fail(WWZB200_ERR_ARG, "bad sizes");
    int rc = check_quantile_args(d_Xcells, S, ncell, d_probs, nq, d_q);
    if (rc || ncell == 0 || nq == 0) return rc;
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    if ((rc = workspace(&w))) return rc;
    ScratchScope scope(w, static_cast<cudaStream_t>(stream));
    if ((rc = scope.begin())) return rc;
    return quantiles_device(w, d_Xcells, S, stride, ncell, d_probs, nq, d_q, static_cast<cudaStream_t>(stream));
}

// series-major [S][ncell] input (caller holds g_mu and a ScratchScope on the stream)
static int quantiles_series_major_nolock(Workspace *w, const double *d_X, int64_t S, int64_t ncell, const double *d_probs,
                                         int64_t nq, double *d_q, cudaStream_t st)
{
    double *d_Xc;
    int rc;
    if ((rc = wsbuf(w, "q_cells", (size_t)S * (size_t)ncell, &d_Xc))) return rc;
    CU(wwz::launch_transpose(d_X, S, ncell, ncell, d_Xc, S, st));   // [S][ncell] -> [ncell][S]
    return quantiles_device(w, d_Xc, S, S, ncell, d_probs, nq, d_q, st);
}

int wwzb200_quantiles_dev(const double *d_X, int64_t S, int64_t ncell, const double *d_probs, int64_t nq, double *d_q,
                          void *stream)
{
    int rc = check_quantile_args(d_X, S, ncell, d_probs, nq, d_q);
    if (rc || ncell == 0 || nq == 0) return rc;
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    if ((rc = workspace(&w))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScratchScope scope(w, st);
    if ((rc = scope.begin())) return rc;
    return quantiles_series_major_nolock(w, d_X, S, ncell, d_probs, nq, d_q, st);
}

// Host buffers in and out: the lock is held until the quantiles have left the device (see surrogates_host).
int wwzb200_quantiles(const double *X, int64_t S, int64_t ncell, const double *probs, int64_t nq, double *q)
{
    int rc = check_quantile_args(X, S, ncell, probs, nq, q);
    if (rc || ncell == 0 || nq == 0) return rc;
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    double *d_X, *d_p, *d_q;
    if ((rc = workspace(&w))) return rc;
    cudaStream_t st = w->stream;
    ScratchScope scope(w, st);
    if ((rc = scope.begin())) return rc;
    if ((rc = wsbuf(w, "q_in", (size_t)S * (size_t)ncell, &d_X))) return rc;
    if ((rc = wsbuf(w, "q_out", (size_t)nq * (size_t)ncell + (size_t)nq, &d_q))) return rc;
    d_p = d_q + (size_t)nq * (size_t)ncell;
    CU(cudaMemcpyAsync(d_X, X, sizeof(double) * (size_t)S * (size_t)ncell, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_p, probs, sizeof(double) * nq, cudaMemcpyHostToDevice, st));
    if ((rc = quantiles_series_major_nolock(w, d_X, S, ncell, d_p, nq, d_q, st))) return rc;
    CU(cudaMemcpyAsync(q, d_q, sizeof(double) * (size_t)nq * (size_t)ncell, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

static int check_member_args(const void *ts, const void *ys, int64_t ys_stride, int64_t M, int64_t nts, const void *taus,
                             int64_t nt, const void *oms, int64_t nf, double c, int64_t thr, const void *Neffs)
{
    if (M < 0 || nts < 1 || nt < 0 || nf < 0) return fail(WWZB200_ERR_ARG, "sizes must satisfy M>=0, nts>=1, nt>=0, nf>=0");
    if (ys_stride != 0 && ys_stride < nts) return fail(WWZB200_ERR_ARG, "ys_stride must be 0 (shared values) or >= nts");
    if (M && (!ts || !ys || (nt && !taus) || (nf && !oms))) return fail(WWZB200_ERR_ARG, "null input pointer");
    if (thr < 1) return fail(WWZB200_ERR_ARG, "Neff_threshold must be a positive integer (utils/wavelet.py:971)");
    if (!(c > 0) || !std::isfinite(c)) return fail(WWZB200_ERR_ARG, "decay constant c must be positive and finite");
    if (M * nt * nf && !Neffs) return fail(WWZB200_ERR_ARG, "Neffs output is required");
    return 0;
}

int wwzb200_kirchner_members_dev(const double *d_ts, const double *d_ys, int64_t ys_stride, int64_t M, int64_t nts,
                                 const double *d_taus, int64_t nt, const double *d_omegas, int64_t nf, double c,
                                 int64_t Neff_threshold, int64_t psd_Neff_threshold, uint32_t flags, double *d_Neffs,
                                 double *d_a0, double *d_a1, double *d_a2, double *d_wwa, double *d_phase, double *d_psd,
                                 double *d_mu_sig, void *stream)
{
    int rc = check_member_args(d_ts, d_ys, ys_stride, M, nts, d_taus, nt, d_omegas, nf, c, Neff_threshold, d_Neffs);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    if ((rc = workspace(&w))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScratchScope scope(w, st);
    if ((rc = scope.begin())) return rc;
    return members_device(w, d_ts, nts, d_ys, ys_stride, M, nts, d_taus, nt, nt, d_omegas, nf, nf, c, Neff_threshold,
                          psd_Neff_threshold, flags, d_Neffs, d_a0, d_a1, d_a2, d_wwa, d_phase, d_psd, d_mu_sig, st);
}

// Host buffers.  The inputs go up once; the members are transformed in groups (scratch bounded by the basis budget,
// at least four groups when there are enough members) and the outputs of a finished group cross PCIe on a copy
// stream while the next group is computed.
int wwzb200_kirchner_members(const double *ts, const double *ys, int64_t ys_stride, int64_t M, int64_t nts,
                             const double *taus, int64_t nt, const double *omegas, int64_t nf, double c,
                             int64_t Neff_threshold, int64_t psd_Neff_threshold, uint32_t flags, double *Neffs,
                             double *a0, double *a1, double *a2, double *wwa, double *phase, double *psd, double *mu_sig)
{
    int rc = check_member_args(ts, ys, ys_stride, M, nts, taus, nt, omegas, nf, c, Neff_threshold, Neffs);
    if (rc) return rc;
    if (M == 0) return 0;
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    if ((rc = workspace(&w))) return rc;
    cudaStream_t st = w->stream;
    ScratchScope scope(w, st);
    if ((rc = scope.begin())) return rc;
    const int64_t ncell = nt * nf;
    const int64_t ny = ys_stride ? M * ys_stride : nts;
    double *d_in, *d_out;
    if ((rc = wsbuf(w, "mb_in", (size_t)(M * nts + ny + M * nt + M * nf), &d_in))) return rc;
    if ((rc = wsbuf(w, "mb_out", (size_t)(6 * M * ncell + M * nf + 2 * M), &d_out))) return rc;
    double *d_ts = d_in, *d_ys = d_ts + M * nts, *d_tau = d_ys + ny, *d_om = d_tau + M * nt;
    CU(cudaMemcpyAsync(d_ts, ts, sizeof(double) * M * nts, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_ys, ys, sizeof(double) * ny, cudaMemcpyHostToDevice, st));
    if (nt) CU(cudaMemcpyAsync(d_tau, taus, sizeof(double) * M * nt, cudaMemcpyHostToDevice, st));
    if (nf) CU(cudaMemcpyAsync(d_om, omegas, sizeof(double) * M * nf, cudaMemcpyHostToDevice, st));
    double *o[6], *h[6] = {Neffs, a0, a1, a2, wwa, phase};
    for (int i = 0; i < 6; ++i) o[i] = d_out + (size_t)i * M * ncell;
    double *o_psd = d_out + (size_t)6 * M * ncell, *o_stats = o_psd + M * nf;
    // group size: the basis rows of a group fit the basis budget; at least four groups once there are 64 members
    const size_t per_member = (size_t)nf * (size_t)(nts + 64) * 32;
    int64_t gsize = (int64_t)std::max<size_t>(1, basis_budget_bytes() / std::max<size_t>(per_member, 1));
    if (M >= 64) gsize = std::min<int64_t>(gsize, (M + 3) / 4);
    gsize = std::max<int64_t>(1, std::min<int64_t>(gsize, M));
    // the copy stream must not start before earlier work on it has finished reading mb_out (stream order), nor read a
    // group before its kernels are done (event)
    int slot = 0;
    for (int64_t m0 = 0; m0 < M; m0 += gsize, slot ^= 1) {
        const int64_t mg = std::min<int64_t>(gsize, M - m0);
        rc = members_device(w, d_ts + m0 * nts, nts, d_ys + m0 * ys_stride, ys_stride, mg, nts, d_tau + m0 * nt, nt, nt,
                            d_om + m0 * nf, nf, nf, c, Neff_threshold, psd_Neff_threshold, flags, o[0] + m0 * ncell,
                            o[1] + m0 * ncell, o[2] + m0 * ncell, o[3] + m0 * ncell, o[4] + m0 * ncell, o[5] + m0 * ncell,
                            psd ? o_psd + m0 * nf : nullptr, mu_sig ? o_stats + 2 * m0 : nullptr, st);
        if (rc) break;
        CU(cudaEventRecord(w->group_ev[slot], st));
        CU(cudaStreamWaitEvent(w->copy, w->group_ev[slot], 0));
        for (int i = 0; i < 6; ++i)
            if (h[i] && ncell)
                CU(cudaMemcpyAsync(h[i] + m0 * ncell, o[i] + m0 * ncell, sizeof(double) * mg * ncell, cudaMemcpyDeviceToHost,
                                   w->copy));
        if (psd && nf)
            CU(cudaMemcpyAsync(psd + m0 * nf, o_psd + m0 * nf, sizeof(double) * mg * nf, cudaMemcpyDeviceToHost, w->copy));
    }
    if (!rc && mu_sig && (flags & WWZB200_STANDARDIZE))
        cudaMemcpyAsync(mu_sig, o_stats, sizeof(double) * 2 * M, cudaMemcpyDeviceToHost, st);
    cudaError_t e1 = cudaStreamSynchronize(w->copy), e2 = cudaStreamSynchronize(st);
    if (rc) return rc;
    CU(e1);
    CU(e2);
    return 0;
}

// Ragged batch: members are independent transforms on their own axes and grids.  The packed inputs
// go to the device once, every member is transformed in stream order with the workspace pre-sized
// for the largest member, the packed outputs come back once.
int wwzb200_kirchner_ragged(const double *ts, const double *pd_ys, const int64_t *ts_off, const double *taus,
                            const int64_t *tau_off, const double *omegas, const int64_t *om_off,
                            const int64_t *out_off, int64_t M, double c, int64_t Neff_threshold,
                            int64_t psd_Neff_threshold, uint32_t flags, double *Neffs, double *a0, double *a1,
                            double *a2, double *wwa, double *phase, double *psd)
{
    if (M < 0) return fail(WWZB200_ERR_ARG, "M must be >= 0");
    if (M == 0) return 0;
    if (!ts || !pd_ys || !ts_off || !taus || !tau_off || !omegas || !om_off || !out_off)
        return fail(WWZB200_ERR_ARG, "null pointer");
    if (Neff_threshold < 1) return fail(WWZB200_ERR_ARG, "Neff_threshold must be a positive integer");
    if (!(c > 0) || !std::isfinite(c)) return fail(WWZB200_ERR_ARG, "decay constant c must be positive and finite");
    int64_t max_cells = 0, max_n = 0, max_nf = 0, max_nt = 0, max_basis = 0, max_sums = 0;
    for (int64_t m = 0; m < M; ++m) {
        const int64_t n = ts_off[m + 1] - ts_off[m], nt = tau_off[m + 1] - tau_off[m], nf = om_off[m + 1] - om_off[m];
        if (n < 1 || nt < 0 || nf < 0 || out_off[m + 1] - out_off[m] != nt * nf)
            return fail(WWZB200_ERR_ARG, "member %lld: inconsistent offsets", (long long)m);
        max_cells = std::max(max_cells, nt * nf);
        max_n = std::max(max_n, n);
        max_nf = std::max(max_nf, nf);
        max_nt = std::max(max_nt, nt);
        max_basis = std::max(max_basis, nf * (n + wwz::kMaxPointsPerStage + 16));
        max_sums = std::max(max_sums, (int64_t)std::max(point_segments(n, nt, nf, false), point_segments(n, nt, nf, true)) * wwz::kNumSums * nt * nf);
    }
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    int rc;
    if ((rc = workspace(&w))) return rc;
    cudaStream_t st = w->stream;
    ScratchScope scope(w, st);
    if ((rc = scope.begin())) return rc;
    const int64_t NT = ts_off[M] - ts_off[0], NTAU = tau_off[M] - tau_off[0], NOM = om_off[M] - om_off[0];
    const int64_t NOUT = out_off[M] - out_off[0];
    double *d_in, *d_out;
    if ((rc = wsbuf(w, "rag_in", (size_t)(2 * NT + NTAU + NOM), &d_in))) return rc;
    if ((rc = wsbuf(w, "rag_out", (size_t)(6 * NOUT + NOM), &d_out))) return rc;
    double *d_ts = d_in, *d_ys = d_in + NT, *d_tau = d_ys + NT, *d_om = d_tau + NTAU;
    CU(cudaMemcpyAsync(d_ts, ts + ts_off[0], sizeof(double) * NT, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_ys, pd_ys + ts_off[0], sizeof(double) * NT, cudaMemcpyHostToDevice, st));
    if (NTAU) CU(cudaMemcpyAsync(d_tau, taus + tau_off[0], sizeof(double) * NTAU, cudaMemcpyHostToDevice, st));
    if (NOM) CU(cudaMemcpyAsync(d_om, omegas + om_off[0], sizeof(double) * NOM, cudaMemcpyHostToDevice, st));
    double *o[6];
    for (int i = 0; i < 6; ++i) o[i] = d_out + (size_t)i * NOUT;
    double *o_psd = d_out + (size_t)6 * NOUT;
    // Members are independent: they are enqueued round-robin on kLanes streams, each with its own scratch
    // set, pre-sized for the largest member (every buffer transform_device asks for, with upper bounds for the
    // padded lengths) so that nothing is re-allocated -- cudaFree synchronises the device -- while work is in flight.
    const int nl = (int)std::min<int64_t>(kLanes, M);
    for (int l = 0; l < nl; ++l) {
        double *dd;
        double2 *d2;
        unsigned int *du;
        int *di;
        w->prefix = "lane" + std::to_string(l) + "_";
        if ((rc = wsbuf(w, "kappa", (size_t)max_nf, &dd))) break;
        if ((rc = wsbuf(w, "dcut", (size_t)max_nf, &dd))) break;
        if ((rc = wsbuf(w, "flags", wwz::kFlagInts, &di))) break;
        if ((rc = wsbuf(w, "grid", 2, &dd))) break;
        { unsigned char *dc; if ((rc = wsbuf(w, "rec_row", (size_t)max_nf, &dc))) break; }
        if ((rc = wsbuf(w, "ty", (size_t)(max_n + wwz::kMaxPointsPerStage + 16), &d2))) break;
        if ((rc = wsbuf(w, "basis", (size_t)std::min<int64_t>(max_basis, (int64_t)(basis_budget_bytes() / 16)) * 2, &d2)))
            break;
        if ((rc = wsbuf(w, "sums", (size_t)max_sums, &dd))) break;
        if ((rc = wsbuf(w, "fix_list", (size_t)max_cells + 1, &du))) break;
        if (psd && (rc = wsbuf(w, "psd_wwa", (size_t)max_cells, &dd))) break;
        if ((rc = wsbuf(w, "taugap", (size_t)std::max<int64_t>(max_nt, 1), &dd))) break;
        if ((rc = wsbuf(w, "minmax", 2, &dd))) break;
        if ((flags & WWZB200_FOSTER) && (rc = wsbuf(w, "sums2", (size_t)max_sums, &dd))) break;
        if (g_profile) {
            unsigned long long *dw;
            if ((rc = wsbuf(w, "walked", 1, &dw))) break;
        }
    }
    w->prefix.clear();
    if (rc) return rc;
    CU(cudaEventRecord(w->fork_ev, st));
    for (int l = 0; l < nl; ++l) CU(cudaStreamWaitEvent(w->lanes[l], w->fork_ev, 0));
    for (int64_t m = 0; m < M && !rc; ++m) {
        const int l = (int)(m % nl);
        const int64_t n = ts_off[m + 1] - ts_off[m], nt = tau_off[m + 1] - tau_off[m], nf = om_off[m + 1] - om_off[m];
        const int64_t t0 = ts_off[m] - ts_off[0], u0 = tau_off[m] - tau_off[0], f0 = om_off[m] - om_off[0];
        const int64_t c0 = out_off[m] - out_off[0];
        double tspan = 0;
        if (psd) {
            const auto mm = std::minmax_element(ts + ts_off[m], ts + ts_off[m + 1]);
            tspan = *mm.second - *mm.first;
        }
        w->prefix = "lane" + std::to_string(l) + "_";
        rc = transform_device(w, d_ts + t0, d_ys + t0, n, d_tau + u0, nt, d_om + f0, nf, c, Neff_threshold,
                              psd_Neff_threshold, flags, o[0] + c0, o[1] + c0, o[2] + c0, o[3] + c0, o[4] + c0,
                              o[5] + c0, psd ? o_psd + f0 : nullptr, nf, tspan, w->lanes[l]);
    }
    w->prefix.clear();
    if (rc) return rc;
    for (int l = 0; l < nl; ++l) {
        CU(cudaEventRecord(w->lane_ev[l], w->lanes[l]));
        CU(cudaStreamWaitEvent(st, w->lane_ev[l], 0));
    }
    double *h[6] = {Neffs, a0, a1, a2, wwa, phase};
    for (int i = 0; i < 6; ++i)
        if (h[i] && NOUT) CU(cudaMemcpyAsync(h[i] + out_off[0], o[i], sizeof(double) * NOUT, cudaMemcpyDeviceToHost, st));
    if (psd && NOM) CU(cudaMemcpyAsync(psd + om_off[0], o_psd, sizeof(double) * NOM, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

// AR(1)-surrogate significance test, everything on the device: generate S surrogates on the target's
// time axis, z-score each (Series.wavelet standardises by default), transform them as one
// shared-axis batch, reduce to per-cell quantiles.  Only Neffs, the [nq, nt, nf] quantiles and the
// optional [nq, nf] PSD quantiles return to the host.
int wwzb200_ar1_signif(const double *ts, int64_t nts, double tau_est, double sigma, double mu, int64_t S,
                       uint64_t seed, const double *taus, int64_t nt, const double *omegas, int64_t nf, double c,
                       int64_t Neff_threshold, int64_t psd_Neff_threshold, uint32_t flags, const double *probs,
                       int64_t nq, double *Neffs, double *q_wwa, double *q_psd)
{
    int rc = check_grid_args(ts, ts, nts, taus, nt, omegas, nf, c, Neff_threshold);
    if (rc) return rc;
    if (S < 1 || nq < 1 || !probs) return fail(WWZB200_ERR_ARG, "need S >= 1 surrogates and nq >= 1 probabilities");
    if (!(tau_est > 0)) return fail(WWZB200_ERR_ARG, "tau_est must be positive");
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    if ((rc = workspace(&w))) return rc;
    cudaStream_t st = w->stream;
    ScratchScope scope(w, st);
    if ((rc = scope.begin())) return rc;
    const int64_t ncell = nt * nf;
    double *d_in, *d_Y, *d_rq, *d_ne, *d_q, *d_psd = nullptr, *d_psd_t = nullptr;
    if ((rc = wsbuf(w, "in", (size_t)(2 * nts + nt + nf + nq), &d_in))) return rc;
    if ((rc = wsbuf(w, "Y", (size_t)S * (size_t)nts, &d_Y))) return rc;
    if ((rc = wsbuf(w, "ar1_rq", (size_t)2 * nts, &d_rq))) return rc;
    if ((rc = wsbuf(w, "ne_out", (size_t)ncell, &d_ne))) return rc;
    if ((rc = wsbuf(w, "q_out", (size_t)nq * (size_t)(ncell + nf), &d_q))) return rc;
    if (q_psd) {
        if ((rc = wsbuf(w, "psd_out", (size_t)S * (size_t)nf, &d_psd))) return rc;
        if ((rc = wsbuf(w, "psd_t", (size_t)S * (size_t)nf, &d_psd_t))) return rc;
    }
    double *d_ts = d_in, *d_tau = d_in + 2 * nts, *d_om = d_tau + nt, *d_probs = d_om + nf;
    CU(cudaMemcpyAsync(d_ts, ts, sizeof(double) * nts, cudaMemcpyHostToDevice, st));
    if (nt) CU(cudaMemcpyAsync(d_tau, taus, sizeof(double) * nt, cudaMemcpyHostToDevice, st));
    if (nf) CU(cudaMemcpyAsync(d_om, omegas, sizeof(double) * nf, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_probs, probs, sizeof(double) * nq, cudaMemcpyHostToDevice, st));
    CU(wwz::launch_ar1(d_ts, nts, tau_est, sigma, mu, S, seed, (flags & WWZB200_UAR1) != 0, d_rq, d_rq + nts, d_Y, st));
    const auto mm = std::minmax_element(ts, ts + nts);
    double *d_amp_cells = nullptr;
    int64_t S_pad = 0;
    rc = shared_axis_device(w, d_ts, d_Y, nts, S, d_tau, nt, d_om, nf, c, Neff_threshold, psd_Neff_threshold,
                            flags | WWZB200_STANDARDIZE, d_ne, nullptr, nullptr, nullptr, nullptr, nullptr, d_psd,
                            *mm.second - *mm.first, &d_amp_cells, &S_pad, st);
    if (rc) return rc;
    if (ncell && q_wwa) {
        if ((rc = quantiles_device(w, d_amp_cells, S, S_pad, ncell, d_probs, nq, d_q, st))) return rc;
        CU(cudaMemcpyAsync(q_wwa, d_q, sizeof(double) * (size_t)nq * ncell, cudaMemcpyDeviceToHost, st));
    }
    if (q_psd && nf) {
        CU(wwz::launch_transpose(d_psd, S, nf, nf, d_psd_t, S, st));   // [S][nf] -> [nf][S]
        double *d_qp = d_q + (size_t)nq * ncell;
        if ((rc = quantiles_device(w, d_psd_t, S, S, nf, d_probs, nq, d_qp, st))) return rc;
        CU(cudaMemcpyAsync(q_psd, d_qp, sizeof(double) * (size_t)nq * nf, cudaMemcpyDeviceToHost, st));
    }
    if (Neffs && ncell) CU(cudaMemcpyAsync(Neffs, d_ne, sizeof(double) * ncell, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

static int check_wtc_args(const void *a1, const void *a2, const void *b1, const void *b2, int64_t S, int64_t nt, int64_t nf,
                          const void *freq, double dt_tau, double smooth, int64_t boxcar, const void *coh, int *npad)
{
    if (S < 0 || nt < 1 || nf < 1) return fail(WWZB200_ERR_ARG, "bad sizes");
    if (S && (!a1 || !a2 || !b1 || !b2 || !freq || !coh)) return fail(WWZB200_ERR_ARG, "null pointer");
    if (!(dt_tau > 0) || !(smooth >= 0)) return fail(WWZB200_ERR_ARG, "dt_tau must be positive, smooth_factor non-negative");
    if (boxcar < 1) return fail(WWZB200_ERR_ARG, "boxcar length must be >= 1 (the reference's rect() needs int(round(0.6/dj*2)) >= 1)");
    int p = 1;
    while (p < nt) p <<= 1;                 // int(2 ** ceil(log2(ntau))), utils/wavelet.py:2310
    if (p > (1 << 20)) return fail(WWZB200_ERR_ARG, "tau axis too long");
    *npad = p;
    return 0;
}

int wwzb200_wtc_dev(const double *d_a1, const double *d_a2, const double *d_b1, const double *d_b2, int64_t S, int64_t nt,
                    int64_t nf, const double *d_freq, double dt_tau, double smooth_factor, int64_t boxcar, double *d_coh,
                    double *d_phase, double *d_xw_amplitude, void *stream)
{
    int npad = 0;
    int rc = check_wtc_args(d_a1, d_a2, d_b1, d_b2, S, nt, nf, d_freq, dt_tau, smooth_factor, boxcar, d_coh, &npad);
    if (rc || S == 0) return rc;
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    if ((rc = workspace(&w))) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScratchScope scope(w, st);
    if ((rc = scope.begin())) return rc;
    double *d_g, *d_T;
    if ((rc = wsbuf(w, "wtc_g", (size_t)npad * (size_t)nf, &d_g))) return rc;
    if ((rc = wsbuf(w, "wtc_T", (size_t)S * 4 * (size_t)nt * (size_t)nf, &d_T))) return rc;
    CU(wwz::launch_wtc(d_a1, d_a2, d_b1, d_b2, S, nt, nf, d_freq, dt_tau, smooth_factor, npad, (int)boxcar, d_g, d_T, d_coh,
                       d_phase, d_xw_amplitude, st));
    return 0;
}

int wwzb200_wtc(const double *a1, const double *a2, const double *b1, const double *b2, int64_t S, int64_t nt, int64_t nf,
                const double *freq, double dt_tau, double smooth_factor, int64_t boxcar, double *coh, double *phase,
                double *xw_amplitude)
{
    int npad = 0;
    int rc = check_wtc_args(a1, a2, b1, b2, S, nt, nf, freq, dt_tau, smooth_factor, boxcar, coh, &npad);
    if (rc || S == 0) return rc;
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    if ((rc = workspace(&w))) return rc;
    cudaStream_t st = w->stream;
    ScratchScope scope(w, st);
    if ((rc = scope.begin())) return rc;
    const size_t cnt = (size_t)S * (size_t)nt * (size_t)nf;
    double *d_io, *d_g, *d_T;
    if ((rc = wsbuf(w, "wtc_io", 7 * cnt + (size_t)nf, &d_io))) return rc;
    if ((rc = wsbuf(w, "wtc_g", (size_t)npad * (size_t)nf, &d_g))) return rc;
    if ((rc = wsbuf(w, "wtc_T", 4 * cnt, &d_T))) return rc;
    double *d_in[4] = {d_io, d_io + cnt, d_io + 2 * cnt, d_io + 3 * cnt};
    double *d_coh = d_io + 4 * cnt, *d_ph = d_io + 5 * cnt, *d_xa = d_io + 6 * cnt, *d_freq = d_io + 7 * cnt;
    const double *h_in[4] = {a1, a2, b1, b2};
    for (int i = 0; i < 4; ++i) CU(cudaMemcpyAsync(d_in[i], h_in[i], sizeof(double) * cnt, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_freq, freq, sizeof(double) * nf, cudaMemcpyHostToDevice, st));
    CU(wwz::launch_wtc(d_in[0], d_in[1], d_in[2], d_in[3], S, nt, nf, d_freq, dt_tau, smooth_factor, npad, (int)boxcar, d_g,
                       d_T, d_coh, phase ? d_ph : nullptr, xw_amplitude ? d_xa : nullptr, st));
    CU(cudaMemcpyAsync(coh, d_coh, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st));
    if (phase) CU(cudaMemcpyAsync(phase, d_ph, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st));
    if (xw_amplitude) CU(cudaMemcpyAsync(xw_amplitude, d_xa, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
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
    if (g_profile) {   // the walked-points counter of the recurrence kernel restarts with the events
        Workspace *w;
        unsigned long long *d_walked;
        int rc;
        if ((rc = workspace(&w))) return rc;
        if ((rc = wsbuf(w, "walked", 1, &d_walked))) return rc;
        CU(cudaMemset(d_walked, 0, sizeof(unsigned long long)));
    }
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

int wwzb200_profile_walked(double *warp_points)
{
    if (!warp_points) return fail(WWZB200_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lock(g_mu);
    Workspace *w;
    unsigned long long *d_walked, h = 0;
    int rc;
    if ((rc = workspace(&w))) return rc;
    if ((rc = wsbuf(w, "walked", 1, &d_walked))) return rc;
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(&h, d_walked, sizeof h, cudaMemcpyDeviceToHost));
    *warp_points = (double)h;
    return 0;
}

}  // extern "C"
