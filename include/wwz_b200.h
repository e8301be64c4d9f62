/*
 * wwz_b200.h -- C ABI of the B200-native weighted wavelet Z-transform (Kirchner variant).
 *
 * This is the drop-in boundary for Pyleoclim's WWZ kernels.  The only native interface the
 * reference has for this path is the f2py module
 *
 *     f2py_wwz.f2py_wwz.wwa(taus, omegas, c, Neff, ts, pd_ys, nthread, nts, nt, nf)
 *         -> Neffs, a0, a1, a2                      pyleoclim/f2py/f2py_wwz.f90:13-21
 *     (call site pyleoclim/utils/wavelet.py:1154, NaN sentinel handling :1156-1159,
 *      amplitude / phase :1160-1161)
 *
 * and every entry point below keeps its conventions: inputs are borrowed read-only
 * `const double*` vectors, `omegas` is the angular-frequency vector already produced by
 * make_omega (utils/wavelet.py:1210-1234; NaN above Nyquist), `pd_ys` is the already
 * preprocessed series (utils/wavelet.py:976), outputs are caller-allocated and laid out
 * [nt, nf] row-major (C order, tau-major rows -- the layout of np.ndarray(shape=(nt, nf)),
 * utils/wavelet.py:980-983).  Masked cells hold NaN (not the Fortran -99999 sentinel).
 *
 * No Python, NumPy or torch type appears in any signature.  All functions return 0 on success
 * and a negative WWZB200_ERR_* code otherwise; wwzb200_last_error() returns a thread-local
 * human-readable message for the last failure.  There is no CPU fallback: without a CUDA
 * device every compute entry point fails with WWZB200_ERR_CUDA.
 *
 * Host-pointer entry points copy host<->device internally (pinned staging, one stream).
 * The *_dev entry points take device pointers and a cudaStream_t (passed as void*); they
 * enqueue work and return without synchronising.
 *
 * Threads and streams.  Every entry point may be called from any host thread: one process-wide lock is held from
 * the first workspace request of a call to its last enqueue (host-pointer entry points: to the final
 * synchronisation).  All calls on one device share one grow-only scratch set; a call whose stream differs from the
 * stream of the previous call first waits (cudaStreamWaitEvent) for the end of that call's work, so calls on
 * different streams are ordered on the device instead of racing on the scratch -- results are always those of the
 * serial order of the calls, and concurrency between two library calls on one device is not available.  Growing a
 * scratch buffer frees the old one (cudaFree synchronises the device).
 *
 * Streamed output.  When a host-pointer transform returns at least 4 MB of per-cell outputs, finished blocks of tau rows
 * are copied to the caller's buffers while later cells are still being computed: the recurrence kernel finishes its own
 * cells (its warps add the partial sums of a tile and write Neff, coefficients, amplitude and phase themselves) and
 * reports every finished tau block through a flag in page-locked host memory that the calling thread polls.  If another
 * kernel turns out to own cells of the grid (rows beyond the recurrence, deep-underflow cells), everything is copied
 * again after the last kernel -- the values returned are the same either way.  Page-locked output buffers
 * (wwzb200_alloc_pinned / wwzb200_host_register) make the copies DMAs.  WWZB200_NO_STREAMED_OUTPUT=1 turns it off.
 */
#ifndef WWZ_B200_H
#define WWZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define WWZB200_OK 0
#define WWZB200_ERR_ARG (-1)   /* bad argument (the reference asserts, utils/wavelet.py:971) */
#define WWZB200_ERR_CUDA (-2)  /* CUDA runtime / launch failure, or no device */
#define WWZB200_ERR_NOMEM (-3) /* host or device allocation failure */

/* flags */
#define WWZB200_DENSE 1u        /* evaluate every (cell, point) pair with its own exp2: no windowing  \
                                   and no tau-recurrence.  The default walks, per tile of cells, only \
                                   the points whose weight reaches 2^-100 of the largest weight of    \
                                   some cell of the tile (the window widens with the distance from    \
                                   tau to the nearest sample): skipped terms are < 2^-100 of sum(w) */
#define WWZB200_EXACT_WINDOW 16u /* window only on weights that are exactly 0.0 in the kernel (below   \
                                   2^-1081): bit-identical to WWZB200_DENSE for the direct kernel */
#define WWZB200_STANDARDIZE 2u  /* batched entry points: z-score every series on the device       \
                                   (utils/tsutils.py:1134-1154, ddof=0, eps=1e-3 guard) */
#define WWZB200_NO_RECURRENCE 8u /* keep one exp2 per (cell, point): do not use the tau-recurrence kernel \
                                   (equally spaced tau, frequency rows with kappa*dtau small) */
#define WWZB200_UAR1 32u        /* wwzb200_ar1_signif: surrogates follow uar1_sim (utils/tsmodel.py:673-716) instead  \
                                   of ar1_model; the `sigma` argument then carries the innovation variance sigma_2 */
#define WWZB200_FOSTER 64u      /* single-series and ragged entry points: Foster's weighted projection on            \
                                   {1, cos w(t-tau), sin w(t-tau)} with the 3x3 pseudo-inverse solve of the           \
                                   reference's method='Foster' (wwz_basic, utils/wavelet.py:341-374) instead of        \
                                   Kirchner's centred covariances; a0,a1,a2 then hold ywave_1,2,3 (:372-374) */
#define WWZB200_FREQ_HZ 128u    /* wwzb200_kirchner_members(_dev): the `omegas` argument holds FREQUENCIES; make_omega   \
                                   (utils/wavelet.py:1210-1234: NaN above 0.5/median(diff(ts)), omega = 2 pi f) is        \
                                   applied per member on the device; needs nts <= 16385 */
#define WWZB200_PREPARED 256u   /* wwzb200_shared_axis_cells_dev: the scratch already holds the y-independent pass of     \
                                   this batch (wwzb200_shared_axis_prepare_dev with the same arguments reported            \
                                   *prepared = 1 and no other transform ran on the device since) */
#define WWZB200_FAITHFUL 4u     /* validation mode: evaluate every cell with the reference's own    \
                                   two-pass expressions (utils/wavelet.py:988-1032), one warp per   \
                                   cell; ~10x slower, agrees with kirchner_numba to ~1e-12 */

const char *wwzb200_last_error(void);
int wwzb200_version(void);

/* (6) device selection.  One CUDA context per process; one process per GPU under torchrun. */
int wwzb200_device_count(int *count);
int wwzb200_set_device(int device);
int wwzb200_get_device(int *device);
int wwzb200_sm_count(int *sms);
/* Release every cached device workspace of the current device. */
int wwzb200_release(void);
/* Page-locked host buffers (so that the host-pointer entry points move data by DMA). */
int wwzb200_alloc_pinned(size_t bytes, void **ptr);
int wwzb200_free_pinned(void *ptr);
/* Device-to-host copy on `stream` (a cudaStream_t), synchronised before returning; a plain DMA when
 * dst is page-locked (wwzb200_alloc_pinned).  For callers that keep results in HBM (the *_dev entry
 * points) and bring an assembled result to the host once. */
int wwzb200_copy_to_host(void *dst, const void *d_src, size_t bytes, void *stream);
/* The same without the synchronisation, and the synchronisation on its own: several blocks of a sharded result are
 * queued back to back and awaited once. */
int wwzb200_copy_to_host_async(void *dst, const void *d_src, size_t bytes, void *stream);
int wwzb200_stream_synchronize(void *stream);
/* Page-lock (and release) host memory the caller owns -- e.g. a POSIX shared-memory segment that the ranks of one box
 * map, so that every GPU copies its block of a sharded result over its own PCIe link into the consumer's buffer. */
int wwzb200_host_register(void *ptr, size_t bytes);
int wwzb200_host_unregister(void *ptr);

/* (1) Single-series Kirchner transform.  Replaces f2py_wwz.wwa (f2py_wwz.f90:13-44) and the cell
 * loop of kirchner_numba (utils/wavelet.py:985-1042), plus wwa/phase (:1044-1045).
 * Any of a0, a1, a2, wwa, phase may be NULL (not returned).  Neffs is required. */
int wwzb200_kirchner(const double *ts, const double *pd_ys, int64_t nts, const double *taus,
                     int64_t nt, const double *omegas, int64_t nf, double c, int64_t Neff_threshold,
                     uint32_t flags, double *Neffs, double *a0, double *a1, double *a2, double *wwa,
                     double *phase);

/* (2) Same transform with the wwa2psd tau-reduction fused behind it (utils/wavelet.py:1270-1278,
 * reached from utils/spectral.py:889-896).  `psd_Neff_threshold` is the threshold wwa2psd uses;
 * the transform itself uses `Neff_threshold` (the reference's wwz_psd does not forward its own
 * threshold to wwz, utils/spectral.py:889-891).  The per-cell outputs may all be NULL. */
int wwzb200_kirchner_psd(const double *ts, const double *pd_ys, int64_t nts, const double *taus,
                         int64_t nt, const double *omegas, int64_t nf, double c,
                         int64_t Neff_threshold, int64_t psd_Neff_threshold, uint32_t flags,
                         double *psd, double *Neffs, double *a0, double *a1, double *a2, double *wwa,
                         double *phase);

/* wwa2psd alone on an existing scalogram (utils/wavelet.py:1270-1278): the reference skips the
 * transform when a scalogram is supplied (utils/spectral.py:888). */
int wwzb200_wwa2psd(const double *wwa, const double *Neffs, int64_t nt, int64_t nf, double tspan,
                    int64_t nts, int64_t Neff_threshold, double *psd);

/* (3) Batched transform of S series that share one time axis -- the AR(1)-surrogate significance
 * loop (core/scalograms.py:515-517, core/psds.py:246-282: every surrogate of
 * SurrogateSeries.from_series keeps the target's time axis, core/surrogateseries.py:170-174).
 * Y is [S, nts] row-major.  The weights, Neffs and basis sums are evaluated once and contracted
 * against all S series.  Outputs: Neffs[nt,nf] (shared), wwa[S,nt,nf]; optional a0/a1/a2/phase
 * [S,nt,nf]; optional psd[S,nf] (wwa2psd per series with psd_Neff_threshold). */
int wwzb200_kirchner_shared_axis(const double *ts, const double *Y, int64_t nts, int64_t S,
                                 const double *taus, int64_t nt, const double *omegas, int64_t nf,
                                 double c, int64_t Neff_threshold, int64_t psd_Neff_threshold,
                                 uint32_t flags, double *Neffs, double *wwa, double *phase, double *a0,
                                 double *a1, double *a2, double *psd);

/* (4) Batched transform of M series with their own time axes and grids -- the age-ensemble case
 * (core/ensembleseries.py:169-173 -> core/multipleseries.py:1520-1532).  Series m occupies
 * ts/ys[ts_off[m] .. ts_off[m+1]), taus[tau_off[m] .. tau_off[m+1]), omegas[om_off[m] ..
 * om_off[m+1]); its outputs occupy [out_off[m] .. out_off[m+1]) with out_off[m+1]-out_off[m] =
 * nt_m*nf_m, each block [nt_m, nf_m] row-major.  Optional psd is packed by om_off. */
int wwzb200_kirchner_ragged(const double *ts, const double *pd_ys, const int64_t *ts_off,
                            const double *taus, const int64_t *tau_off, const double *omegas,
                            const int64_t *om_off, const int64_t *out_off, int64_t M, double c,
                            int64_t Neff_threshold, int64_t psd_Neff_threshold, uint32_t flags,
                            double *Neffs, double *a0, double *a1, double *a2, double *wwa,
                            double *phase, double *psd);

/* (4b) Batched transform of M members of EQUAL shape: the age ensemble of core/ensembleseries.py:47-175 (one value
 * vector on M time axes; default grids of equal size), which the reference hands member by member to a process pool
 * (core/multipleseries.py:1520-1532).  ts [M][nts]; pd_ys [M][ys_stride] or, with ys_stride = 0, one [nts] vector
 * shared by all members; taus [M][nt]; omegas [M][nf] (frequencies with WWZB200_FREQ_HZ).  WWZB200_STANDARDIZE
 * z-scores every member on the device.  Every kernel of the transform is launched once for all members of a group;
 * the host variant moves finished groups over PCIe while the next group is computed.  Outputs [M][nt][nf] each
 * (Neffs required, the others optional), psd [M][nf] optional (wwa2psd per member, time span taken on the device),
 * mu_sig [M][2] optional: the mean and standard deviation applied by WWZB200_STANDARDIZE ((0, 1): member left as it is,
 * the case the reference warns about, utils/tsutils.py:1146-1149).
 * WWZB200_FOSTER / WWZB200_FAITHFUL are refused here (use the ragged entry point). */
int wwzb200_kirchner_members(const double *ts, const double *pd_ys, int64_t ys_stride, int64_t M, int64_t nts,
                             const double *taus, int64_t nt, const double *omegas, int64_t nf, double c,
                             int64_t Neff_threshold, int64_t psd_Neff_threshold, uint32_t flags, double *Neffs,
                             double *a0, double *a1, double *a2, double *wwa, double *phase, double *psd,
                             double *mu_sig);

/* (5) AR(1) surrogates on an uneven time axis: S realisations of ar1_model
 * (utils/tsmodel.py:59-69) as driven by ar1_sim (:159-165) and SurrogateSeries.from_series
 * (core/surrogateseries.py:137-139):  y[0]=0; rho_i=exp(-(t_i-t_{i-1})/tau_est);
 * y_i = y_{i-1} rho_i + sigma sqrt(1-rho_i^2) z_i;  Y[s,i] = y_i + mu.
 * z ~ N(0,1) from a counter-based Philox4x32-10 stream keyed by (seed, s, i); the reference
 * draws from NumPy's unseeded legacy global generator (its `seed` argument is ineffective,
 * core/surrogateseries.py:132-133), so streams are statistically, not bitwise, equivalent.
 * Y is [S, nts] row-major. */
int wwzb200_ar1_surrogates(const double *ts, int64_t nts, double tau_est, double sigma, double mu,
                           int64_t S, uint64_t seed, double *Y);

/* The 'uar1' surrogates of SurrogateSeries.from_series (core/surrogateseries.py:147-156): S realisations of
 * uar1_sim (utils/tsmodel.py:673-716) on the time axis ts with the maximum-likelihood parameters of
 * uar1_fit (:646-671; a Nelder-Mead fit on the host):  y_0 = z_0; phi_i = exp(-(t_i - t_{i-1})/tau);
 * y_i = phi_i y_{i-1} + sqrt(sigma_2 (1 - phi_i^2)) z_i;  Y[s,i] = y_i + mu.  Same Philox stream as
 * wwzb200_ar1_surrogates (the reference draws one normal(size=(n,p)) matrix from NumPy's global generator). */
int wwzb200_uar1_surrogates(const double *ts, int64_t nts, double tau, double sigma_2, double mu, int64_t S,
                            uint64_t seed, double *Y);

/* Per-cell quantiles over a batch: scipy.stats.mstats.mquantiles(alphap=betap=0.4) semantics as
 * used by MultipleScalogram.quantiles (core/scalograms.py:635-641) and MultiplePSD.quantiles
 * (core/psds.py:913).  X is [S, ncell] row-major, q is [nq, ncell]. */
int wwzb200_quantiles(const double *X, int64_t S, int64_t ncell, const double *probs, int64_t nq,
                      double *q);

/* The whole AR(1)-surrogate significance test of a scalogram / PSD on the device
 * (Scalogram.signif_test core/scalograms.py:515-522, PSD.signif_test core/psds.py:246-282):
 * (5) S surrogates on the target's time axis -> z-score each (Series.wavelet / .spectral standardise
 * by default, core/series.py:3559, utils/wavelet.py:976) -> (3) shared-axis transform -> per-cell
 * quantiles.  Returns Neffs[nt,nf], q_wwa[nq,nt,nf] and, when q_psd != NULL, the quantiles of the
 * surrogate spectra q_psd[nq,nf] (wwa2psd with psd_Neff_threshold).  Only these leave the device. */
int wwzb200_ar1_signif(const double *ts, int64_t nts, double tau_est, double sigma, double mu, int64_t S,
                       uint64_t seed, const double *taus, int64_t nt, const double *omegas, int64_t nf,
                       double c, int64_t Neff_threshold, int64_t psd_Neff_threshold, uint32_t flags,
                       const double *probs, int64_t nq, double *Neffs, double *q_wwa, double *q_psd);

/* (7) Wavelet transform coherency of S pairs of WWZ coefficient sets on one (tau, f) grid -- utils/wavelet.py:2240-2364
 * (wtc) and :2201-2238 (xwt), what wwz_coherence (:1667-1672) computes for one pair and Coherence.signif_test
 * (core/coherences.py:748-963) for every surrogate pair.  a1, a2: coefficients of the first series, b1, b2 of the
 * second, each [S][nt][nf] (coeff = a1 - i a2, :1664-1665); freq [nf] (scale = 1/freq); dt_tau = median(diff(tau));
 * smooth_factor as the reference's (0.25); boxcar = int(round(0.6 / dj * 2)) with dj = log2(scale[0]/scale[-1]) / nf
 * (:2346-2349, >= 1).  Outputs [S][nt][nf]: xw_coherence (required), xw_phase and xw_amplitude (optional).  The
 * Gaussian smoothing in time is evaluated as the circular convolution over npad = 2^ceil(log2 nt) points that the
 * reference's FFT / multiply / inverse FFT computes; NaN cells propagate as they do there. */
int wwzb200_wtc(const double *a1, const double *a2, const double *b1, const double *b2, int64_t S, int64_t nt, int64_t nf,
                const double *freq, double dt_tau, double smooth_factor, int64_t boxcar, double *xw_coherence,
                double *xw_phase, double *xw_amplitude);

/* ---- device-pointer entry points (inputs/outputs resident in HBM; async on `stream`) ---- */
/* (4b) on device-resident members: same layouts as wwzb200_kirchner_members, all M members in one group. */
int wwzb200_kirchner_members_dev(const double *d_ts, const double *d_pd_ys, int64_t ys_stride, int64_t M, int64_t nts,
                                 const double *d_taus, int64_t nt, const double *d_omegas, int64_t nf, double c,
                                 int64_t Neff_threshold, int64_t psd_Neff_threshold, uint32_t flags, double *d_Neffs,
                                 double *d_a0, double *d_a1, double *d_a2, double *d_wwa, double *d_phase, double *d_psd,
                                 double *d_mu_sig, void *stream);

int wwzb200_wtc_dev(const double *d_a1, const double *d_a2, const double *d_b1, const double *d_b2, int64_t S, int64_t nt,
                    int64_t nf, const double *d_freq, double dt_tau, double smooth_factor, int64_t boxcar,
                    double *d_xw_coherence, double *d_xw_phase, double *d_xw_amplitude, void *stream);

int wwzb200_kirchner_dev(const double *d_ts, const double *d_pd_ys, int64_t nts, const double *d_taus,
                         int64_t nt, const double *d_omegas, int64_t nf, double c,
                         int64_t Neff_threshold, int64_t psd_Neff_threshold, uint32_t flags,
                         double *d_Neffs, double *d_a0, double *d_a1, double *d_a2, double *d_wwa,
                         double *d_phase, double *d_psd, int64_t out_row_stride, void *stream);

int wwzb200_wwa2psd_dev(const double *d_wwa, const double *d_Neffs, int64_t nt, int64_t nf, double tspan,
                        int64_t nts, int64_t Neff_threshold, double *d_psd, void *stream);

/* wwa2psd over a block of tau rows, before the divide: d_parts[0..nf) = nansum_j(power * Neff_diff),
 * d_parts[nf..2nf) = nansum_j(Neff_diff) (utils/wavelet.py:1275-1278).  tau-sharded runs add the parts of all
 * blocks (an all-reduce of [2, nf]) and divide; tspan and nts are those of the whole series (:1270). */
int wwzb200_wwa2psd_parts_dev(const double *d_wwa, const double *d_Neffs, int64_t nt, int64_t nf, double tspan,
                              int64_t nts, int64_t Neff_threshold, double *d_parts, void *stream);

int wwzb200_kirchner_shared_axis_dev(const double *d_ts, const double *d_Y, int64_t nts, int64_t S,
                                     const double *d_taus, int64_t nt, const double *d_omegas,
                                     int64_t nf, double c, int64_t Neff_threshold,
                                     int64_t psd_Neff_threshold, uint32_t flags, double *d_Neffs,
                                     double *d_wwa, double *d_phase, double *d_a0, double *d_a1,
                                     double *d_a2, double *d_psd, void *stream);

int wwzb200_ar1_surrogates_dev(const double *d_ts, int64_t nts, double tau_est, double sigma,
                               double mu, int64_t S, uint64_t seed, double *d_Y, void *stream);

int wwzb200_uar1_surrogates_dev(const double *d_ts, int64_t nts, double tau, double sigma_2, double mu,
                                int64_t S, uint64_t seed, double *d_Y, void *stream);

/* ---- building blocks of the surrogate-sharded significance test (one rank = one block of surrogates) ----
 * wwzb200_surrogates_block_dev: rows [first_series, first_series + S) of the surrogate batch that
 * wwzb200_ar1_surrogates / wwzb200_uar1_surrogates (flags & WWZB200_UAR1) would generate with the same seed.
 * wwzb200_series_pad: the padded series count S_pad the batch kernels use for S series.
 * wwzb200_shared_axis_cells_dev: entry (3) with the amplitudes left cell-major, d_amp_cells[nt*nf][S_pad]
 * (what the per-cell quantiles read); the caller exchanges cell blocks between ranks (all-to-all) and calls
 * wwzb200_quantiles_cells_dev: quantiles of X[ncell][stride] over the first S entries of every row -> q[nq][ncell]. */
int wwzb200_surrogates_block_dev(const double *d_ts, int64_t nts, double tau, double sigma, double mu, int64_t S,
                                 int64_t first_series, uint64_t seed, uint32_t flags, double *d_Y, void *stream);
int wwzb200_series_pad(int64_t S, int64_t *S_pad);
int wwzb200_shared_axis_cells_dev(const double *d_ts, const double *d_Y, int64_t nts, int64_t S, const double *d_taus,
                                  int64_t nt, const double *d_omegas, int64_t nf, double c, int64_t Neff_threshold,
                                  uint32_t flags, double *d_Neffs, double *d_amp_cells, void *stream);
int wwzb200_quantiles_cells_dev(const double *d_Xcells, int64_t S, int64_t stride, int64_t ncell,
                                const double *d_probs, int64_t nq, double *d_q, void *stream);

/* wwzb200_shared_axis_cells_dev with the exchange of the surrogate-sharded significance test fused into the kernel:
 * the (tau, f) cells are dealt in blocks of cells_per to n_peers GPUs of the box (cell block i belongs to peer i), and
 * the contraction kernel stores every amplitude straight into the owner's buffer peer_bufs[i] over NVLink (peer memory
 * mapped through CUDA IPC, wwzb200_ipc_*): row = cell inside the block, row stride dst_stride doubles, this batch's
 * series at columns col0 .. col0 + S (padding columns up to the padded series count are written too and must lie
 * inside the row).  No staging buffer, no separate all-to-all: the transfer overlaps the contraction tile by tile.  The
 * caller orders "all peers have finished writing" before it reads its own buffer (any collective on the stream). */
int wwzb200_shared_axis_cells_scatter_dev(const double *d_ts, const double *d_Y, int64_t nts, int64_t S,
                                          const double *d_taus, int64_t nt, const double *d_omegas, int64_t nf, double c,
                                          int64_t Neff_threshold, uint32_t flags, double *d_Neffs, void *const *peer_bufs,
                                          int64_t n_peers, int64_t cells_per, int64_t dst_stride, int64_t col0,
                                          void *stream);
/* Device memory the other ranks of the box can map: allocate + export a 64-byte handle; map a peer's handle; unmap;
 * free.  (cudaIpcGetMemHandle / cudaIpcOpenMemHandle: one process per GPU, all on one node.) */
int wwzb200_ipc_alloc(size_t bytes, void **d_ptr, unsigned char *handle64);
int wwzb200_ipc_open(const unsigned char *handle64, void **d_ptr);
int wwzb200_ipc_close(void *d_ptr);
int wwzb200_ipc_free(void *d_ptr);
/* The part of wwzb200_shared_axis_cells_dev that does not depend on the series -- setup, basis rows, rotations, the
 * stored weight fragments and Neffs (utils/wavelet.py:988-1019 depend on ts, tau, omega only) -- enqueued on its own, so
 * that a caller can run it under host work that the series depend on (the tau_estimation fit of the surrogates,
 * utils/tsmodel.py:243-275).  S = series the batch will hold.  *prepared = 1 if the following cells call may pass
 * WWZB200_PREPARED (the batch is large enough for stored weights), 0 if it has to run in full.  d_Neffs optional. */
int wwzb200_shared_axis_prepare_dev(const double *d_ts, int64_t nts, int64_t S, const double *d_taus, int64_t nt,
                                    const double *d_omegas, int64_t nf, double c, int64_t Neff_threshold, uint32_t flags,
                                    double *d_Neffs, int *prepared, void *stream);
/* wwzb200_quantiles_cells_dev on rows that are cut into segments: value i of row `cell` lies at
 * X[(i / seg_len) * seg_stride + cell * stride + i % seg_len] -- the receive buffer [rank][cell][stride] of an
 * all-to-all in which every rank contributed seg_len surrogates (the last one possibly fewer). */
int wwzb200_quantiles_segments_dev(const double *d_X, int64_t S, int64_t seg_len, int64_t seg_stride, int64_t stride,
                                   int64_t ncell, const double *d_probs, int64_t nq, double *d_q, void *stream);
int wwzb200_quantiles_dev(const double *d_X, int64_t S, int64_t ncell, const double *d_probs,
                          int64_t nq, double *d_q, void *stream);

/* FP64 FMA peak of the current device, measured by a register-resident DFMA-chain kernel
 * (the roofline denominator; MEASURED_PEAKS.json has no FP64 entry).  Returns TFLOP/s. */
int wwzb200_measure_fp64_peak(double *tflops, double *ms);

/* Kernel launch counter (incremented by every kernel this library launches). */
int64_t wwzb200_launch_count(void);

/* In-library timing of the hot kernel (wwz_cells_kernel): while enabled, CUDA events are recorded
 * on the launching stream around every launch.  wwzb200_profile(1) starts (and clears), (0) stops;
 * wwzb200_profile_read waits for the recorded events and returns the summed kernel time, the
 * number of launches and the (cell, point) pairs they covered. */
int wwzb200_profile(int enable);
int wwzb200_profile_read(double *cell_kernel_ms, int64_t *cell_kernel_launches, double *cell_points);
/* While profiling: the number of (warp, point) pairs the tau-recurrence kernel actually evaluated since
 * wwzb200_profile(1) (point windows skip the rest).  Times the 88 FP64 instructions of its loop body per pair, this
 * is the executed FP64 instruction count -- the live counterpart of ncu's sm__inst_executed_pipe_fp64. */
int wwzb200_profile_walked(double *warp_points);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* WWZ_B200_H */
