// Internal host-side interface between the C ABI (wwz_capi.cu) and the kernel launchers
// (wwz_kernels.cu).  Not part of the public boundary; see include/wwz_b200.h for that.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wwz {

constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kCellThreads = kConsumerThreads + 32;  // + one TMA producer warp
constexpr int kStages = 3;
constexpr int kNumSums = 7;  // W, Q, Ws, Wc, Wy, Wys, Wyc
constexpr int kStageBudget = 32 * 1024;
constexpr int kMaxPointsPerStage = 512;
constexpr int kRecCells = 8;     // tau-adjacent cells per lane in the tau-recurrence flavour of the warp-tile kernel
constexpr int kTileDirectCells = 2;     // ... in its direct flavour (small tau grids)
constexpr int kSmallTauGrid = 64;       // at most this many taus: tau-compact warp tiles for every row
constexpr int kFlagInts = 8;     // per-transform device ints: [0] unsorted flag, [1] deep-underflow counter,
                                 // [2..3] work queue of the recurrence launch, [4..5] of the direct warp-tile launch
constexpr int kMaxWarps = 16;

// Geometry of one launch of the cell kernel, derived on the host by plan_cells().
struct CellPlan {
    int cells_per_lane;   // M: tau-adjacent cells owned by one lane (1, 2, 4; 8 = recurrence kernel)
    int consumer_warps;   // consumer warps per CTA
    int groups;           // G = ceil(nt / M) lane slots per frequency row
    int ft_max;           // max number of frequency rows a 256-slot tile can touch
    int points;           // P: points per smem stage
    int64_t n_pad;        // series length rounded up to a multiple of P
    int64_t row_stride;   // smem basis row stride in bytes ((P+1)*32)
    int stage_bytes;      // bytes of one smem stage
    int smem_bytes;       // dynamic smem of the kernel
    int64_t tiles;        // grid size for nf rows
};

CellPlan plan_cells(int64_t n, int64_t nt, int64_t nf, int cells_per_lane_hint, int64_t seg_len);

// Geometry of the warp-tile kernels (wwz_rec.cu), derived on the host by plan_rec().
struct RecPlan {
    int cells_per_lane;   // kRecCells = tau-recurrence flavour, kTileDirectCells = direct flavour
    int groups;           // G = ceil(nt / cells_per_lane) tau-groups per frequency row
    int tg_log2;          // a warp tile is 2^tg_log2 tau-groups x (32 >> tg_log2) frequency rows
    int growth;           // 1: stages carry the [P][TR] block of rho growth factors (recurrence flavour)
    int stages;           // stages of a warp's ring
    int points;           // P: points per stage of a warp's ring
    int64_t n_pad;
    int64_t row_stride;   // smem basis row stride in bytes (P*32 + 16)
    int stage_bytes;      // bytes of one stage
    int warp_bytes;       // shared memory per warp (ring + mbarriers)
    int warps, nreg;      // CTA shape: warps per CTA, register budget per thread
    int smem_bytes;       // dynamic shared memory of the CTA
};
// rho_growth: recurrence flavour with the tabulated growth of rho (launch_rec then needs the d_estep scratch)
RecPlan plan_rec(int64_t n, int64_t nt, int64_t nf, int64_t seg_len, int cells_per_lane = kRecCells, bool rho_growth = false);
// Fused epilogue of the recurrence flavour (launch_rec, optional): the warp that finishes the last point segment of a
// tile writes the tile's finished cells itself; finalize_kernel then only takes the rows of the direct flavour
// (launch_finalize*, skip_rec_rows).  tile_cnt: one zeroed counter per (tile, member) of the launch (rec_tile_count;
// left zeroed).  unfused: counts the tiles whose rows all belong to the direct flavour.  blk_cnt / blk_flag (optional):
// progress per block of blk_gt tiles along tau -- blk_flag[b] (page-locked host memory, mapped) is set to blk_epoch once
// every cell of the block's taus is final in device memory; blk_cnt: zeroed device counters, one per block.
struct RecFuse {
    const double *omega;
    int64_t m_omega, row_stride;
    double thr;
    double *neffs, *a0, *a1, *a2, *wwa, *phase;
    unsigned int *tile_cnt, *fix_count, *fix_list;
    unsigned int fix_cap;
    unsigned int *unfused;
    unsigned int *blk_cnt;
    volatile unsigned int *blk_flag;
    unsigned int blk_epoch;
    int blk_gt;
};
int64_t rec_tile_count(const RecPlan &plan, int64_t nf_batch, int64_t members);
int rec_tau_tiles(const RecPlan &plan);   // tiles along the tau axis
int rec_tile_taus(const RecPlan &plan);   // taus covered by one tile
// d_queue: two zeroed ints (work queue; left zeroed by the launch).  d_rec_row: rows marked 1 belong to the recurrence
// flavour, rows marked 0 to the direct flavour (null: the launch takes every row).
cudaError_t launch_rec(const RecPlan &plan, const double2 *d_ty, const double2 *d_basis, const double *d_ts, int64_t n,
                       const double *d_tau, int64_t nt, const double *d_kappa, const double *d_dcut,
                       const double *d_taugap, const int *d_unsorted, int64_t k0, int64_t nf_batch, int64_t nf_total, int segs,
                       int64_t seg_len, const unsigned char *d_rec_row, const double *d_grid, double *d_sums,
                       unsigned int *d_queue, unsigned long long *d_walked, int sms, cudaStream_t st, int64_t members = 1,
                       int64_t m_ts = 0, int64_t m_tau = 0, const RecFuse *fuse = nullptr, double *d_estep = nullptr);
// doubles of the optional d_estep scratch of a recurrence-flavour launch (per point and row: the growth of rho to the next
// point; with it the kernel advances rho by one multiplication per point)
int64_t rec_estep_count(const RecPlan &plan, int64_t nf_batch, int64_t members);

// Per-frequency constants kappa4 = 4*omega*sqrt(c*log2 e) and dcut = distance beyond which a weight is below
// 2^-cut_bits (1081 = exactly zero weights only; windowing); which rows the tau-recurrence flavour may take
// (d_rec_row; allow_rec = false marks every row for the direct flavour) and the tau spacing (d_grid[0]), decided on
// the device; gap[j] = distance from tau[j] to the nearest sample of the sorted time axis; the sortedness flag; the
// min / max of the time axis when d_minmax != NULL; and the reset of the kFlagInts ints at d_flags.
// One launch (+ one memset) for the whole preamble of a transform.
cudaError_t launch_setup(const double *d_omega, int64_t nf, const double *d_tau, int64_t nt, const double *d_ts, int64_t n,
                         double c, bool allow_rec, double cut_bits, double *d_kappa, double *d_dcut, double *d_grid,
                         unsigned char *d_rec_row, double *d_gap, int *d_flags, double *d_minmax, cudaStream_t st);

// ---- batches of equally shaped members (ensembles).  m_* = element stride between consecutive members of an input
// (0 = one vector shared by all members); scratch and outputs are laid out [member][...].
constexpr int kNyquistMaxPoints = 16384;   // member_prep sorts the sampling steps of a member in shared memory
// setup of every member; with d_omega_out != NULL, d_omega holds FREQUENCIES and make_omega (utils/wavelet.py:1210-1234)
// is applied here with the per-member Nyquist frequencies d_fnyq
cudaError_t launch_setup_members(const double *d_omega, int64_t m_omega, int64_t nf, const double *d_tau, int64_t m_tau,
                                 int64_t nt, const double *d_ts, int64_t m_ts, int64_t n, int64_t members, double c,
                                 bool allow_rec, double cut_bits, double *d_kappa, double *d_dcut, double *d_grid,
                                 unsigned char *d_rec_row, double *d_gap, int *d_flags, double *d_minmax,
                                 const double *d_fnyq, double *d_omega_out, cudaStream_t st);
// d_fnyq (nullable): 0.5 / median(diff(ts)) per member; d_ysz (nullable): [member][n] values, z-scored when asked;
// d_stats (nullable): [member][2] the mean and standard deviation applied ((0, 1) for a nearly constant member)
cudaError_t launch_member_prep(const double *d_ts, int64_t m_ts, int64_t n, const double *d_ys, int64_t m_ys,
                               int64_t members, bool standardize, double *d_fnyq, double *d_ysz, double *d_stats,
                               cudaStream_t st);
cudaError_t launch_basis_members(const double *d_ts, int64_t m_ts, const double *d_ys, int64_t m_ys, int64_t n,
                                 int64_t n_pad, const double *d_omega, int64_t m_omega, int64_t nf, int64_t members,
                                 double2 *d_ty, double2 *d_basis, cudaStream_t st);
cudaError_t launch_finalize_members(const double *d_sums, const double *d_tau, int64_t m_tau, const double *d_omega,
                                    int64_t m_omega, int64_t nt, int64_t nf, int64_t members, double neff_thr,
                                    double *d_neffs, double *d_a0, double *d_a1, double *d_a2, double *d_wwa,
                                    double *d_phase, int64_t out_row_stride, unsigned int *d_fix_count,
                                    unsigned int *d_fix_list, unsigned int fix_cap, int segs, cudaStream_t st,
                                    const unsigned char *d_skip_rows = nullptr);
cudaError_t launch_faithful_members(const double *d_ts, int64_t m_ts, const double *d_ys, int64_t m_ys, int64_t n,
                                    const double *d_tau, int64_t m_tau, const double *d_omega, int64_t m_omega, int64_t nt,
                                    int64_t nf, int64_t members, double c, double neff_thr,
                                    const unsigned int *d_fix_count, const unsigned int *d_fix_list, unsigned int fix_cap,
                                    double *d_neffs, double *d_a0, double *d_a1, double *d_a2, double *d_wwa,
                                    double *d_phase, int64_t out_row_stride, int sms, cudaStream_t st);

// (t, y) interleaved + per-frequency basis rows (sin, cos, y sin, y cos), zero padded to n_pad.
// cos_as_y: records (sin, cos, cos*sin, cos*cos) -- the second pass of Foster's method (wwz_foster.cu).
cudaError_t launch_basis(const double *d_ts, const double *d_ys, int64_t n, int64_t n_pad,
                         const double *d_omega, int64_t nf_batch, double2 *d_ty, double2 *d_basis,
                         bool write_ty, bool cos_as_y, cudaStream_t st);

// The hot kernel: weighted sums for every (tau, f) cell of a frequency batch.
// sums is [segs][7][nt * nf_total], each slab laid out like the outputs (cell = j*nf_total + k);
// segment s covers points [s*seg_len, (s+1)*seg_len).
cudaError_t launch_cells(const CellPlan &plan, const double2 *d_ty, const double2 *d_basis, const double *d_ts,
                         int64_t n, const double *d_tau, int64_t nt, const double *d_kappa,
                         const double *d_dcut, const double *d_taugap, const int *d_unsorted, int64_t k0, int64_t nf_batch,
                         int64_t nf_total, bool dense, int segs, int64_t seg_len, const unsigned char *d_rec_row,
                         const double *d_grid, double *d_sums, cudaStream_t st);

// Per-cell epilogue: Neff, mask, a0, a1, a2, amplitude, phase from the seven sums.
cudaError_t launch_finalize(const double *d_sums, const double *d_tau, const double *d_omega, int64_t nt,
                            int64_t nf, double neff_thr, double *d_neffs, double *d_a0, double *d_a1,
                            double *d_a2, double *d_wwa, double *d_phase, int64_t out_row_stride,
                            unsigned int *d_fix_count, unsigned int *d_fix_list, unsigned int fix_cap, int segs,
                            cudaStream_t st, const unsigned char *d_skip_rows = nullptr);

// Faithful two-pass evaluation (one warp per cell) of the cells in fix_list (count read on the
// device), or of every cell when fix_list is NULL.
cudaError_t launch_faithful(const double *d_ts, const double *d_ys, int64_t n, const double *d_tau,
                            const double *d_omega, int64_t nt, int64_t nf, double c, double neff_thr,
                            const unsigned int *d_fix_count, const unsigned int *d_fix_list, unsigned int fix_cap,
                            double *d_neffs, double *d_a0, double *d_a1, double *d_a2, double *d_wwa,
                            double *d_phase, int64_t out_row_stride, int sms, cudaStream_t st);

// Foster's method (wwz_foster.cu): epilogue over the sums of both passes (3x3 pseudo-inverse solve per cell) and
// the warp-per-cell evaluator with the reference's own expressions (utils/wavelet.py:343-374).
cudaError_t launch_finalize_foster(const double *d_sums, const double *d_sums2, const double *d_tau,
                                   const double *d_omega, int64_t nt, int64_t nf, double neff_thr, double *d_neffs,
                                   double *d_a0, double *d_a1, double *d_a2, double *d_wwa, double *d_phase,
                                   int64_t out_row_stride, unsigned int *d_fix_count, unsigned int *d_fix_list,
                                   unsigned int fix_cap, int segs, cudaStream_t st);
cudaError_t launch_foster_faithful(const double *d_ts, const double *d_ys, int64_t n, const double *d_tau,
                                   const double *d_omega, int64_t nt, int64_t nf, double c, double neff_thr,
                                   const unsigned int *d_fix_count, const unsigned int *d_fix_list,
                                   unsigned int fix_cap, double *d_neffs, double *d_a0, double *d_a1, double *d_a2,
                                   double *d_wwa, double *d_phase, int64_t out_row_stride, int sms, cudaStream_t st);

// wwa2psd tau-reduction (one thread per frequency column, rows added in order).
cudaError_t launch_psd(const double *d_wwa, const double *d_neffs, int64_t nt, int64_t nf, int64_t row_stride,
                       double tspan, int64_t n, double neff_thr, double *d_psd, cudaStream_t st);

// partial column sums of wwa2psd over a block of tau rows: d_parts[0][k] = numerator, d_parts[1][k] = denominator
cudaError_t launch_psd_parts(const double *d_wwa, const double *d_neffs, int64_t nt, int64_t nf, int64_t row_stride,
                             double tspan, int64_t n, double neff_thr, double *d_parts, cudaStream_t st);

// min / max of a device vector (for tspan in the device-pointer entry points) -> d_out[0..1].
cudaError_t launch_minmax(const double *d_x, int64_t n, double *d_out, cudaStream_t st);
cudaError_t launch_psd_devspan(const double *d_wwa, const double *d_neffs, int64_t nt, int64_t nf,
                               int64_t row_stride, const double *d_minmax, int64_t n, double neff_thr,
                               double *d_psd, cudaStream_t st, int64_t members = 1);

// ---- batched paths (wwz_batch.cu) -------------------------------------------------------------
// Where the cell-major amplitudes of a shared-axis batch go.  cells_per == 0: one local buffer [ncell][S_pad].
// Otherwise the cells are dealt in blocks of cells_per to up to kMaxPeers destinations -- buffers in the HBM of the
// GPUs of this box, mapped through CUDA IPC -- and the contraction kernel stores every amplitude straight into the
// owner's buffer over NVLink: row = cell inside the owner's block, row stride dst_stride, this batch's series at
// columns col0 ...  (the exchange of the surrogate-sharded significance test, fused into the kernel's epilogue).
constexpr int kMaxPeers = 16;
struct AmpScatter {
    double *peer[kMaxPeers];
    long long cells_per, dst_stride, col0;
};

struct SharedPlan {
    int series_block;     // NS: series per CTA tile of the shared-time-axis kernel (64, or 16 for small batches)
    int points, stage_bytes, smem_bytes;
    int64_t n_pad, row_stride, y_off, w_off;
    bool stored_weights;  // the A fragments (weights) are evaluated once and streamed to every series block
    size_t weights_count; // doubles of the fragment store (tiles * n_pad * 64)
};
SharedPlan plan_shared(int64_t n, int64_t nt, int64_t nf, int64_t S);

// AR(1) surrogates Y[S][n] (row-major); d_rho, d_q are n-element scratch.  uar1: uar1_sim semantics
// (y_0 = z_0, sigma = innovation variance sigma_2) instead of ar1_model's.
// first_series: global index of row 0 (the Philox stream of a series is keyed by its global index, so a block of
// surrogates generated on its own equals the same rows of the whole batch).
cudaError_t launch_ar1(const double *d_ts, int64_t n, double tau, double sigma, double mu, int64_t S, uint64_t seed,
                       bool uar1, double *d_rho, double *d_q, double *d_Y, cudaStream_t st, uint64_t first_series = 0);
int shared_series_block(int64_t S);
// per-series mean/std (identity when !standardize) and the tiled copy Yt[S_pad/NS][n_pad][NS]
cudaError_t launch_standardize_retile(const double *d_Y, int64_t n, int64_t n_pad, int64_t S, int series_block,
                                      bool standardize, double *d_mu, double *d_sig, double *d_Yt, cudaStream_t st);
// {sin, cos}(omega_k t_i) records (16 B) + (t, y) pairs
cudaError_t launch_basis_sc(const double *d_ts, int64_t n, int64_t n_pad, const double *d_omega, int64_t nf,
                            double2 *d_ty, double2 *d_basis, cudaStream_t st);
cudaError_t launch_rotation(const double *d_tau, const double *d_omega, int64_t nt, int64_t nf, double2 *d_rot,
                            cudaStream_t st);
// the contraction on the FP64 tensor pipe (wwz_mma.cu)
cudaError_t launch_shared(const SharedPlan &plan, const double2 *d_ty, const double2 *d_basis, const double *d_Yt,
                          const double *d_ts, int64_t n, int64_t S_pad, const double *d_tau, int64_t nt,
                          const double *d_kappa, const double *d_dcut, const double *d_taugap, const int *d_unsorted,
                          int64_t nf, const double2 *d_rot, double thr, double *d_wfrag, double *d_cellsums, double *d_neffs,
                          double *d_wq, double *d_amp, double *d_phase, double *d_a0, double *d_a1, double *d_a2,
                          cudaStream_t st, int phase = 3,    // phase: 1 = y-independent pass (stored weights), 2 = contraction
                          const AmpScatter *scatter = nullptr);
// deep-underflow cells of a shared-axis batch (flagged from W, Q), recomputed faithfully for every series
cudaError_t launch_shared_fixup(const double *d_ts, const double *d_Yt, int64_t n, int64_t n_pad, int64_t S,
                                int64_t S_pad, int series_block, const double *d_tau, const double *d_omega, int64_t nt,
                                int64_t nf, double c, double thr, const double *d_wq, unsigned int *d_fix_count,
                                unsigned int *d_fix_list, double *d_neffs, double *d_amp, double *d_phase,
                                double *d_a0, double *d_a1, double *d_a2, int sms, cudaStream_t st,
                                const AmpScatter *scatter = nullptr);
cudaError_t launch_transpose(const double *d_in, int64_t rows, int64_t cols, int64_t in_stride, double *d_out,
                             int64_t out_stride, cudaStream_t st);
cudaError_t launch_psd_batch(const double *d_amp, const double *d_neffs, int64_t nt, int64_t nf, int64_t S,
                             int64_t S_pad, double tspan, int64_t n, double thr, double *d_psd, cudaStream_t st);
cudaError_t launch_quantiles(const double *d_X, int64_t S, int64_t stride, int64_t ncell, const double *d_probs,
                             int64_t nq, double *d_q, unsigned long long *d_scratch, int64_t scratch_ctas, int sms,
                             cudaStream_t st, int64_t seg_len = 0, int64_t seg_stride = 0);

// Wavelet transform coherency of coefficient pairs (wwz_coherence.cu): coefficients and outputs [S][nt][nf];
// d_g [npad][nf] and d_T [S][4][nt][nf] are scratch; L = boxcar length over the scale axis (>= 1).
cudaError_t launch_wtc(const double *d_a1, const double *d_a2, const double *d_b1, const double *d_b2, int64_t S, int64_t nt,
                       int64_t nf, const double *d_freq, double dt_tau, double smooth, int npad, int L, double *d_g,
                       double *d_T, double *d_coh, double *d_phase, double *d_xwa, cudaStream_t st);

// FP64 FMA peak microbenchmark.
cudaError_t run_fp64_peak(double *tflops, double *ms);

// Opt-in dynamic shared memory is a per-device attribute of a kernel: remember, per device, the largest size set.
constexpr int kMaxDevices = 64;
struct SmemOptIn {
    int bytes[kMaxDevices];
    SmemOptIn() { for (int i = 0; i < kMaxDevices; ++i) bytes[i] = -1; }
};
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, SmemOptIn &done)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const bool tracked = dev >= 0 && dev < kMaxDevices;
    if (tracked && bytes <= done.bytes[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && tracked) done.bytes[dev] = bytes;
    return e;
}

int64_t launch_counter();
void bump_launch_counter(int64_t k = 1);

}  // namespace wwz
