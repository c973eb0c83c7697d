/*
 * hepmc_b200 -- C ABI of the B200-native hot path of hepmc (mathisgerdes/hep-monte-carlo).
 *
 * The reference is pure Python/NumPy: it has NO plugin/FFI layer.  The
 * drop-in boundary is therefore its Python class API (hepmc_b200/ mirrors it) and THIS
 * header is what that mirror -- or any other host language -- binds.  Each entry point
 * names the reference function (path relative to /root/reference/src/hepmc/, file:line)
 * whose NumPy body it replaces.
 *
 * Conventions
 *   - every pointer named *_dev is DEVICE memory owned by the caller (e.g. a torch CUDA
 *     tensor's data_ptr()); nothing is allocated, retained or freed by the library
 *     except the explicitly documented workspaces the caller also passes in;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every
 *     call only ENQUEUES work on it and returns without synchronising, unless stated;
 *   - all arithmetic is IEEE float64; indices are int64 unless stated;
 *   - return value: 0 on success, a negative HEPMC_E_* code for argument errors, or a
 *     positive cudaError_t; hepmc_last_error() returns a thread-local message;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * RNG: Philox4x32-10 (Random123).  key = 64-bit seed; counter = (item index lo, item
 * index hi, draw-block index inside the item, stream id).  An "item" is a sample point
 * (integrators, RAMBO) or a chain (samplers) identified by its GLOBAL index, so any
 * partition of the index range over GPUs reproduces the same draws.
 */
#ifndef HEPMC_B200_H
#define HEPMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HEPMC_B200_ABI_VERSION 2
#define HEPMC_MAX_DIM 32          /* max dimensionality of a built-in density      */
#define HEPMC_MAX_DIVISIONS 255   /* max VEGAS bins per axis (bin ids are bytes)    */
#define HEPMC_N_SLOTS 8           /* fixed reduction slots: GPU-count invariance    */
#define HEPMC_MAX_CHANNELS 16      /* multi-channel importance sampling                */
#define HEPMC_MAX_CHANNEL_DIM 16
#define HEPMC_CHANNEL_ROW (8 + 3 * HEPMC_MAX_CHANNEL_DIM)

/* error codes (negative); positive values are cudaError_t */
#define HEPMC_E_ARG (-1)          /* bad argument (null pointer, size, kind)        */
#define HEPMC_E_UNSUPPORTED (-2)  /* shape outside what the kernels are built for   */
#define HEPMC_E_NODEVICE (-3)     /* no CUDA device                                 */

/* built-in analytic densities: core/densities/{camel,banana,gaussian,uniform}.py */
enum hepmc_density_kind {
    HEPMC_CAMEL = 0,    /* Camel: two-Gaussian mixture restricted to the open unit cube (camel.py:83-100) */
    HEPMC_UCAMEL = 1,   /* UnconstrainedCamel (camel.py:9-45)                                           */
    HEPMC_BANANA = 2,   /* Banana (banana.py:8-51)                                                      */
    HEPMC_GAUSSIAN = 3, /* Gaussian with diagonal covariance (gaussian.py:8-76)                         */
    HEPMC_UNIFORM = 4,  /* Uniform on the open box (low, high)^ndim (uniform.py:7-34)                   */
    HEPMC_NONE = 5      /* no fused integrand: sample/weight only (two-kernel path)                     */
};

/* Parameter block of a built-in density (host struct, passed by pointer, copied into
 * kernel parameters).  Filled by the host mirror exactly the way SciPy's
 * multivariate_normal prepares its constants:
 *   CAMEL/UCAMEL: a[d]=mu_a, b[d]=mu_b, s0=sqrt(1/cov), s1=n*log(2pi)+n*log(cov), s2=cov
 *   BANANA:       s0=bananicity b, a[d]=sqrt(1/cov_d) with cov=(100,1,..,1), s1=n*log(2pi)+log(100)
 *   GAUSSIAN:     a[d]=mean, b[d]=sqrt(1/cov_d), c[d]=1/cov_d, s1=n*log(2pi)+sum(log cov_d)
 *   UNIFORM:      s0=low, s1=high, s2=volume
 */
typedef struct hepmc_density {
    int32_t kind;
    int32_t ndim;
    double a[HEPMC_MAX_DIM];
    double b[HEPMC_MAX_DIM];
    double c[HEPMC_MAX_DIM];
    double s0, s1, s2, s3;
} hepmc_density_t;

enum hepmc_eval_what {
    HEPMC_PDF = 0,          /* Density.pdf           density.py:70   -> (N,)      */
    HEPMC_POT = 1,          /* Density.pot           density.py:48   -> (N,)      */
    HEPMC_POT_GRADIENT = 2, /* Density.pot_gradient  density.py:55   -> (N,ndim)  */
    HEPMC_PDF_GRADIENT = 3  /* Density.pdf_gradient  density.py:73   -> (N,ndim)  */
};

/* ---- library / device ------------------------------------------------------------ */
int hepmc_abi_version(void);
const char* hepmc_last_error(void);
/* out[0]=SM count, out[1]=cc major, out[2]=cc minor, out[3]=max smem/block optin, out[4]=SM clock kHz */
int hepmc_device_info(int device, int64_t* out5);
/* FP64 DFMA micro-benchmark used as the FP64-SIMT roofline denominator.  Runs
 * `iters` dependent-chain rounds of 8 independent DFMAs per thread on a full grid, and
 * returns the elapsed milliseconds (CUDA events, synchronises) and the flop count. */
int hepmc_fp64_probe(int64_t iters, double* ms_out, double* flops_out, void* stream);
/* MUFU.EX2 micro-benchmark (8 independent ex2.approx chains per thread, full grid): elapsed milliseconds
 * and the number of exponentials; the special-function roofline of the float32 / tensor-core surrogate
 * kernels (one exponential per chain, node and gradient evaluation). */
int hepmc_mufu_probe(int64_t iters, double* ms_out, double* ops_out, void* stream);

/* ---- native generator --------------------------------------------------------------
 * The integrators' native draws (the counterpart of np.random.randint / np.random.rand in
 * stratified_volume.py:110-121 and np.random.random in integration.py:41) come in three
 * modes, chosen per call (`rng_mode`):
 *   HEPMC_RNG_U52_R10 (default): 52-bit uniforms / 64 random bits per (bin, offset) pair,
 *                                Philox4x32-10, two axes per Philox call;
 *   HEPMC_RNG_U32_R10:           32 random bits per axis: uniform (w + 1/2) 2^-32; VEGAS
 *                                t = K w: bin = t >> 32, offset = (t mod 2^32 + 1/2) 2^-32
 *                                (bin probabilities within 2^-32 of 1/K); four axes per call;
 *   HEPMC_RNG_U32_R7:            the same with Philox4x32-7 (the smallest round count that
 *                                passes BigCrush in Salmon et al., SC'11; Random123 KATs).
 * Counters are the same in every mode (item index, call index, stream id), so results stay
 * independent of the GPU count.  The modes apply to the integrators and to RAMBO; the chain kernels
 * (normal proposals, momenta, accept uniforms) always use 52-bit uniforms with 10 rounds. */
enum hepmc_rng_mode { HEPMC_RNG_U52_R10 = 0, HEPMC_RNG_U32_R10 = 1, HEPMC_RNG_U32_R7 = 2, HEPMC_RNG_MODES = 3 };

/* ---- densities: replaces Density.pdf/pot/pot_gradient/pdf_gradient ---------------- */
/* xs_dev: (n, ndim) row-major.  out_dev: (n,) or (n, ndim).
 * Replaces camel.py:20-45,87-100; banana.py:13-51; gaussian.py:30-47; uniform.py:15-31;
 * generic rules density.py:48-68 (pot=-log pdf, NaN->inf; pot_gradient=inf where pdf==0).
 * BANANA pot follows the reference's sign (returns +log pdf, banana.py:35). */
int hepmc_density_eval(const hepmc_density_t* dens, int what, const double* xs_dev,
                       int64_t n, double* out_dev, void* stream);

/* ---- plain Monte Carlo: replaces PlainMC.__call__ integration.py:28-50 ------------ */
/* Deterministic reduction unit: one CTA processes one CHUNK of consecutive points with a
 * fixed point->thread mapping and writes one partial row.  The chunk size must be a
 * function of the call's TOTAL point count only (never of the GPU count); this is the
 * library's choice.  Always a multiple of 256. */
int64_t hepmc_default_chunk(int64_t n_total);

/* Points first_index .. first_index+n-1 of the global stream.  Native RNG unless
 * u_replay_dev (n, ndim) is given (the reference's own np.random.random draws).
 * Optional outputs data_dev (n, ndim), f_dev (n,).  Per-chunk deterministic partials
 * (count, mean, M2) are written to partials_dev[(n_chunks) x 3], n_chunks =
 * ceil(n / chunk); chunk boundaries are relative to first_index, which must be a
 * multiple of chunk.  dens->kind == HEPMC_NONE: draw data only (two-kernel path); no partial rows
 * are written, so first_index may then be any global index (point i always uses counter
 * first_index + i). */
int hepmc_plain_mc_sample(const hepmc_density_t* dens, int64_t n, int64_t first_index, int64_t chunk,
                          uint64_t seed, uint32_t stream_id, int rng_mode, const double* u_replay_dev,
                          double* data_dev, double* f_dev, double* partials_dev, void* stream);

/* Generic ordered reduction of (count, mean, M2) rows: rows [row_begin[s], row_begin[s+1])
 * -> out_dev[s*3 .. s*3+2] for s < n_out, merged in row order (Chan et al.).  Used for the
 * slot partials of every integrator and for the final combine (n_out = 1).
 * Replaces the np.mean / np.var(ddof=0) pair of integration.py:47-48, importance.py:47-50,
 * vegas.py:110-111. */
int hepmc_reduce_stats(const double* rows_dev, const int64_t* row_begin_host, int n_out,
                       double* out_dev, void* stream);

/* mean/var of values[i] (optionally values[i]/weights[i]) for the two-kernel path:
 * per-chunk partials like hepmc_plain_mc_sample.  importance.py:47-50. */
int hepmc_stats_partials(const double* values_dev, const double* weights_dev, double weight_scalar,
                         int64_t n, int64_t chunk, double* partials_dev, void* stream);

/* ---- VEGAS: replaces GridVolumes.random_bins/sample_indices (stratified_volume.py:
 * 104-122), VegasMC.weights_at (vegas.py:68-72) and the loop body vegas.py:100-111 ---- */
/* grid_dev: (ndim, 2K+1) = per axis [bounds(K+1) | sizes(K)].
 * Samples points first_index..first_index+n-1 of iteration stream `stream_id`.
 * Replay: idx_replay_dev (ndim, n) int64 + u_replay_dev (ndim, n) (reference layout).
 * Optional outputs (reference layouts): x_dev (ndim, n), f_dev (n,), w_dev (n,),
 * idx_dev (ndim, n) int64.  If dens->kind == HEPMC_NONE no integrand is evaluated and
 * no partials are produced (two-kernel path: follow with hepmc_vegas_accumulate).
 * Partials: stats_partials_dev (n_chunks,3) and hist_partials_dev (n_chunks, ndim*K),
 * hist[d*K+k] = sum of f*w over the chunk's points with bin k on axis d. */
int hepmc_vegas_sample(const hepmc_density_t* dens, const double* grid_dev, int ndim, int divisions,
                       int64_t n, int64_t first_index, int64_t chunk, uint64_t seed, uint32_t stream_id, int rng_mode,
                       const int64_t* idx_replay_dev, const double* u_replay_dev,
                       double* x_dev, double* f_dev, double* w_dev, int64_t* idx_dev,
                       double* stats_partials_dev, double* hist_partials_dev, void* stream);

/* Two-kernel path, second half: f_dev (n,) user integrand values, w_dev (n,), idx_dev
 * (ndim, n) int64 -> the same partials as hepmc_vegas_sample. */
int hepmc_vegas_accumulate(const double* f_dev, const double* w_dev, const int64_t* idx_dev,
                           int ndim, int divisions, int64_t n, int64_t chunk,
                           double* stats_partials_dev, double* hist_partials_dev, void* stream);

/* Ordered column sums of hist partial rows [row_begin[s], row_begin[s+1]) -> out (n_out, ncols). */
int hepmc_reduce_hist(const double* rows_dev, const int64_t* row_begin_host, int n_out, int ncols,
                      double* out_dev, void* stream);

/* Grid adaptation: replaces vegas.py:110-118 + damped_update (helper.py:56-72) + the
 * sizes setter (stratified_volume.py:161-169, with bounds[d][0] = 0).
 * slot_stats_dev (n_slots,3), slot_hist_dev (n_slots, ndim*K) are merged in slot order;
 * est_var_dev[2*j], [2*j+1] receive est_j and var_j; grid_dev is updated in place. */
int hepmc_vegas_update_grid(const double* slot_stats_dev, const double* slot_hist_dev, int n_slots,
                            double* grid_dev, int ndim, int divisions, double c, int j,
                            double* est_var_dev, void* stream);

/* The slot reduction of one VEGAS iteration in one call: chunk partial rows (stats (n_chunks,3), hist
 * (n_chunks, ncols)) -> PACKED slot rows packed_dev (n_out, 3 + ncols) = [count, mean, M2, hist...], the
 * layout that is all-gathered between GPUs (vegas.py:108-111).  Two levels with fixed row groups and a
 * fixed combine order (one launch: the last level-1 CTA of a slot runs its level 2): bit-identical for
 * any partition of the slots over GPUs.  scratch_dev: hepmc_reduce_slots_scratch_bytes(n_out, ncols)
 * bytes of device memory whose LAST 256 bytes (arrival counters) are zero before the first call; the
 * kernel leaves them zero, so the same scratch serves every iteration. */
int64_t hepmc_reduce_slots_scratch_bytes(int n_out, int ncols);
int hepmc_reduce_slots(const double* stats_rows_dev, const double* hist_rows_dev, const int64_t* row_begin_host,
                       int n_out, int ncols, double* packed_dev, void* scratch_dev, void* stream);
/* hepmc_vegas_update_grid on packed slot rows (n_slots, 3 + ndim*K), e.g. the all-gather receive buffer. */
int hepmc_vegas_update_grid_packed(const double* packed_dev, int n_slots, double* grid_dev, int ndim, int divisions,
                                   double c, int j, double* est_var_dev, void* stream);

/* ---- the same exchange over NVLink peer memory, fused into the two kernels (multi-GPU VEGAS) ----------
 * Every rank owns an EXCHANGE buffer of hepmc_slot_exchange_bytes(ncols) bytes, zero-initialised, mapped
 * into all peers of the node (e.g. torch.distributed._symmetric_memory): [2 parities][8 slots][3 + ncols]
 * doubles followed by [2][8] 64-bit flags.  hepmc_reduce_slots_p2p = hepmc_reduce_slots whose last CTA of
 * each slot also stores the packed row into EVERY rank's buffer (peer_buffers_host[r], r < world, pointers
 * valid on this GPU) at the slot's global position (first_slot + local index) and then releases
 * flag = epoch there.  hepmc_vegas_update_grid_p2p = the grid update on this rank's own buffer, waiting
 * (bounded spin; *timeout_flag_dev != 0 afterwards means a peer never arrived) until all 8 flags of the
 * parity carry `epoch`.  epoch: iteration counter > 0, equal on all ranks, increasing by one per
 * iteration over the life of the buffer.  No NCCL call, no host synchronisation: replaces the all-gather
 * of vegas.py:108-111's sums across GPUs. */
int64_t hepmc_slot_exchange_bytes(int ncols);
int hepmc_reduce_slots_p2p(const double* stats_rows_dev, const double* hist_rows_dev, const int64_t* row_begin_host,
                           int n_out, int ncols, double* packed_dev, void* scratch_dev,
                           const int64_t* peer_buffers_host, int world, int first_slot, uint64_t epoch, void* stream);
int hepmc_vegas_update_grid_p2p(const double* exchange_dev, int ncols, uint64_t epoch, double* grid_dev, int ndim,
                                int divisions, double c, int j, double* est_var_dev, int* timeout_flag_dev, void* stream);

/* Workspace (bytes) hepmc_vegas_integrate needs for n points per iteration. */
int64_t hepmc_vegas_workspace_bytes(int ndim, int divisions, int64_t n);
/* Whole single-GPU VegasMC.__call__ (vegas.py:74-118): `iterations` x (sample, slot
 * reduce, grid update) enqueued back to back, no host synchronisation.  est_var_dev
 * (iterations, 2).  Iteration j uses stream id stream_id0 + j. */
int hepmc_vegas_integrate(const hepmc_density_t* dens, double* grid_dev, int ndim, int divisions,
                          int64_t n, int iterations, double c, uint64_t seed, uint32_t stream_id0, int rng_mode,
                          double* est_var_dev, void* workspace_dev, int64_t workspace_bytes,
                          void* stream);

/* GridVolumes.pdf (stratified_volume.py:74-83, default base counts): xs (n, ndim) -> (n,) */
int hepmc_grid_pdf(const double* grid_dev, int ndim, int divisions, const double* xs_dev,
                   int64_t n, double* out_dev, void* stream);

/* ---- multi-channel importance sampling ("next" row 1): replaces MultiChannel.rvs/pdf/sample
 * (core/integration/multi_channel.py:103-160) and the sampling + reduction half of
 * MultiChannelMC.iterate (importance.py:141-175) ------------------------------------------ */
/* channel_table_dev: (n_channels, HEPMC_CHANNEL_ROW) doubles, one row per channel:
 *   [0] kind (HEPMC_GAUSSIAN | HEPMC_UNIFORM), [1] channel weight alpha,
 *   GAUSSIAN: [2] n*log(2pi) + sum(log cov_d), [8+d] mean_d, [8+16+d] sqrt(cov_d), [8+32+d] sqrt(1/cov_d)
 *   UNIFORM : [2] low, [3] high, [4] volume.
 * Point i (global index first_index + i) picks its channel with probability alpha (i.i.d.
 * categorical = multinomial counts, multi_channel.py:141), is drawn from it, and the mixture
 * density p = sum_c alpha_c p_c(x) (:103-114) is evaluated.  Replay: xs_replay_dev (n, ndim)
 * and ch_replay_dev (n,) int64 (the reference's own sample and its channel of origin).
 * Optional outputs x_dev (n, ndim), f_dev (n,), p_dev (n,), ch_dev (n,) int64.
 * Partials per chunk: stats_partials_dev (n_chunks, 3) = (count, mean, M2) of f/p;
 * ch_partials_dev (n_chunks, 2*n_channels) = per-channel counts, then per-channel sums of
 * (f/p)^2 (importance.py:166-167).  chunk must be a multiple of 128. */
int hepmc_multichannel_sample(const hepmc_density_t* dens, const double* channel_table_dev, int n_channels, int ndim,
                              int64_t n, int64_t first_index, int64_t chunk, uint64_t seed, uint32_t stream_id,
                              const double* xs_replay_dev, const int64_t* ch_replay_dev,
                              double* x_dev, double* f_dev, double* p_dev, int64_t* ch_dev,
                              double* stats_partials_dev, double* ch_partials_dev, void* stream);
/* two-kernel path: user integrand values f, mixture densities p, channel ids -> the same partials */
int hepmc_multichannel_accumulate(const double* f_dev, const double* p_dev, const int64_t* ch_dev, int n_channels,
                                  int64_t n, int64_t chunk, double* stats_partials_dev, double* ch_partials_dev,
                                  void* stream);

/* ---- RAMBO: replaces phase_space.Rambo.map (rambo.py:38-69,162-173) ---------------- */
/* out_dev (n, 4*nparticles) as [E,px,py,pz] per particle.  Native RNG (point index
 * first_index + i; rng_mode as for the integrators: in the 32-bit modes the four uniforms of
 * particle p are the four words of Philox call p, in HEPMC_RNG_U52_R10 two calls per particle)
 * unless xs_replay_dev (n, 4*nparticles) is given. */
int hepmc_rambo_map(int nparticles, double e_cm, int64_t n, int64_t first_index, uint64_t seed,
                    uint32_t stream_id, int rng_mode, const double* xs_replay_dev, double* out_dev, void* stream);
/* Rambo.pdf (rambo.py:28-36): the constant 1/Phi_n, computed on the host. */
double hepmc_rambo_pdf(int nparticles, double e_cm);

/* ---- RamboOnDiet + accept/reject ("next" row 4): phase_space.RamboOnDiet.map / map_inverse
 * (core/phase_space/rambo.py:72-149, boost :180-189) and the acceptance test of
 * AcceptRejectSampler.sample (core/sampling.py:196-201) ------------------------------------- */
/* map: xs (n, 3*nparticles-4) unit-cube points (native RNG when xs_replay_dev is NULL) ->
 * out_dev (n, 4*nparticles).  The reference's per-point brentq root becomes an in-kernel
 * safeguarded Newton iteration (converged to 4e-16 relative; brentq stops at xtol = 2e-12). */
int hepmc_rambo_diet_map(int nparticles, double e_cm, int64_t n, int64_t first_index, uint64_t seed, uint32_t stream_id,
                         const double* xs_replay_dev, double* out_dev, void* stream);
/* inverse: p_dev (n, 4*nparticles) -> xs_out_dev (n, 3*nparticles-4) */
int hepmc_rambo_diet_inverse(int nparticles, double e_cm, int64_t n, const double* p_dev, double* xs_out_dev, void* stream);
/* flags[i] = (u_i * bound * q_i <= f_i) with u_i the native uniform of item first_index + i;
 * q_dev may be NULL (then q_scalar is used). */
int hepmc_accept_flags(const double* f_dev, const double* q_dev, double q_scalar, double bound, int64_t n,
                       int64_t first_index, uint64_t seed, uint32_t stream_id, unsigned char* flags_dev, void* stream);

/* out_dev[i, :] = src_dev[order_dev[i], :] for rows of row_len doubles: the by-channel grouping of a
 * multi-channel sample (MultiChannel.rvs / ChannelsSample return the points channel by channel,
 * multi_channel.py:141-150); order_dev = positions of channel 0's points, then channel 1's, ... */
int hepmc_gather_rows(const double* src_dev, const int64_t* order_dev, int64_t n, int row_len,
                      double* out_dev, void* stream);

/* ---- random-walk Metropolis ensembles: replaces MarkovUpdate.sample (base.py:44-97)
 * + MetropolisUpdate.next_state/accept (metropolis.py:37-50,76-95) + DefaultMetropolis.
 * proposal (:122-124) + proposals.Gaussian.proposal (gaussian.py:25-27), one chain per
 * thread, diagonal proposal covariance -------------------------------------------------- */
/* x0_dev (n_chains, ndim).  proposal_kind 0: Gaussian random walk, prop_scale_host[ndim] =
 * sqrt(diag(cov)); proposal_kind 1: independence proposal uniform on (prop_scale_host[0],
 * prop_scale_host[1])^ndim = DefaultMetropolis' default Uniform(ndim) (metropolis.py:114-115,
 * uniform.py:33-34; the replay tape then holds uniforms); proposal_kind 2: UniformLocal
 * (proposals/uniform_local.py:5-50), prop_scale_host = [delta_0 .. delta_{ndim-1}, low, high]:
 * candidate_d = min(max(low, x_d - delta_d/2), high - delta_d) + u delta_d with ONE uniform u per
 * step (column 0 of the replay tape), plain Metropolis ratio.  sample_size T >= 1
 * (chain[0] = x0, T-1 transitions).  Chain c has global index first_chain + c.
 * Replay tapes: z_replay_dev (n_chains, T-1, ndim), u_replay_dev (n_chains, T-1).
 * Outputs: chain_dev (n_chains, n_kept, ndim) with n_kept = ceil(T / thin), storing states
 * t = 0, thin, 2*thin, ... (NULL = do not store); last_dev (n_chains, ndim); accepted_dev
 * (n_chains,) int64 = number of transitions whose state changed (base.py:72-73);
 * total_accepted_dev (1,) int64 is atomically incremented by the grid total. */
int hepmc_rwm_chains(const hepmc_density_t* dens, const double* x0_dev, int proposal_kind,
                     const double* prop_scale_host, int64_t n_chains, int64_t first_chain, int64_t sample_size, int64_t thin,
                     uint64_t seed, uint32_t stream_id,
                     const double* z_replay_dev, const double* u_replay_dev,
                     double* chain_dev, double* last_dev, int64_t* accepted_dev,
                     int64_t* total_accepted_dev, void* stream);

/* ---- HMC ensembles: replaces HamiltonianUpdate (hmc.py:54-88), HamiltonLeapfrog
 * (simulation.py:26-46), with the momentum distribution Gaussian(0, diag(mass)) ---------- */
enum hepmc_grad_mode {
    HEPMC_GRAD_EXACT = 0,     /* target density's own pot_gradient                         */
    HEPMC_GRAD_ELM_GAUSS = 1, /* GaussianBasis surrogate (extreme_learning.py:222-230)      */
    HEPMC_GRAD_ELM_TRIG = 2   /* TrigBasis surrogate (extreme_learning.py:140-146,253-257)  */
};
/* Extreme-learning-machine surrogate parameters (device pointers):
 *   GAUSS: p0 = centers (nodes, ndim), p1 = widths (nodes,), out_w (nodes,)
 *   TRIG : p0 = in_weights (nodes, ndim), p1 = biases (nodes,), out_w (nodes,) */
typedef struct hepmc_elm {
    int32_t mode;       /* hepmc_grad_mode */
    int32_t nodes;
    const double* p0_dev;
    const double* p1_dev;
    const double* out_w_dev;
    double out_bias;
} hepmc_elm_t;

/* mass_host[ndim] = diag of the momentum covariance.  Momentum tape p_replay_dev
 * (n_chains, T-1, ndim) holds the momenta themselves.  Other arguments as hepmc_rwm_chains.
 * precision: 0 = float64 everywhere (parity mode); 1 = surrogate gradient in float32
 * SIMT; 2 = surrogate gradient on tcgen05 tensor cores (TF32 operands, FP32 accumulate) in both
 * GEMMs: every term of the gradient sum carries ~2^-11 relative error, fine while the trained
 * weights cancel by less than ~1e3; 3 = exponent GEMM on the tensor cores (3xTF32), contraction
 * with the output weights in FP32 on the FMA pipe (float32-level accuracy for any weights).
 * The Metropolis test always uses the float64 target pdf, so 1, 2 and 3 stay exact samplers.
 * coop (surrogate gradients, precision 0/1): 1 = one CTA per chain with the node sum of every gradient
 * split over its threads (small ensembles), 0 = one chain per thread, -1 = decide from n_chains of THIS
 * call.  The two layouts add the node terms in different orders (last-bit differences), so a caller
 * that shards one ensemble over several calls or GPUs passes 0 or 1 to keep the chains independent of
 * the sharding. */
int hepmc_hmc_chains(const hepmc_density_t* dens, const hepmc_elm_t* elm, int precision, int coop,
                     const double* x0_dev, const double* mass_host, double step_size, int steps,
                     int64_t n_chains, int64_t first_chain, int64_t sample_size, int64_t thin,
                     uint64_t seed, uint32_t stream_id,
                     const double* p_replay_dev, const double* u_replay_dev,
                     double* chain_dev, double* last_dev, int64_t* accepted_dev,
                     int64_t* total_accepted_dev, void* stream);

/* HamiltonLeapfrog.__call__ (hamiltonian/simulation.py:26-46), batched: n independent states (q, p),
 * q_dev / p_dev (n, ndim), -> the states after `steps` leapfrog steps of size step_size:
 *   g = dU(q);  steps x { p -= eps/2 g;  q += eps (p / mass);  g = dU(q);  p -= eps/2 g }
 * (half steps not merged, steps + 1 gradient evaluations, exactly the loop hepmc_hmc_chains runs).
 * dU = the density's pot_gradient or the ELM surrogate (elm non-NULL; precision 0 / 1; coop as in
 * hepmc_hmc_chains); kinetic term = Gaussian(0, diag mass_host).pot_gradient (gaussian.py:45-47).
 * Outputs may alias the inputs.  Non-finite results are returned as such (the reference returns
 * (None, None) on a NumPy RuntimeWarning; the host mirror maps all-non-finite states to that). */
int hepmc_leapfrog(const hepmc_density_t* dens, const hepmc_elm_t* elm, int precision, int coop,
                   const double* mass_host, double step_size, int steps, int64_t n,
                   const double* q_dev, const double* p_dev, double* q_out_dev, double* p_out_dev, void* stream);

/* ---- (MC)^3 mixing chains ("next" row 3): replaces MixingMarkovUpdate.next_state
 * (core/markov/base.py:139-179) over [DefaultMetropolis(target, MultiChannel) , local update],
 * i.e. the sampler mc3.BasicMC3 / interfaces/mc3_hmc*.py build ------------------------------ */
/* Per step and chain: with probability beta an importance-sampling Metropolis-Hastings update
 * (candidate drawn from the channel mixture, Hastings ratio pdf(c) q(x) / (pdf(x) q(c)),
 * metropolis.py:46-47), otherwise the local update: local_kind 0 = Gaussian random walk with
 * diag covariance local_scale_host[ndim]; local_kind 1 = HMC with diag mass local_scale_host[ndim],
 * step_size, steps and the exact or ELM-surrogate gradient (elm may be NULL; precision 0/1);
 * local_kind 2 = UniformLocal Metropolis (mc3.MC3Uniform, mc3.py:100-114), local_scale_host =
 * [delta_0 .. delta_{ndim-1}, low, high] as in hepmc_rwm_chains (uniform in column 0 of the draw tape).
 * channel_table_dev as in hepmc_multichannel_sample.  Replay tapes (all three or none):
 * sel_replay_dev (C, T-1) int64 (0 = IS update, 1 = local), draw_replay_dev (C, T-1, ndim) = the
 * candidate (IS) / standard normals (random walk) / momentum (HMC) / proposal uniform (UniformLocal),
 * u_replay_dev (C, T-1).  Outputs as hepmc_rwm_chains; local_steps_dev (optional, one int64, zeroed by
 * the caller) receives the number of local updates taken over all chains (the evaluation counter of
 * util.count_calls: every local HMC update costs steps + 1 gradient evaluations).  coop as in
 * hepmc_hmc_chains. */
int hepmc_mc3_chains(const hepmc_density_t* dens, const double* channel_table_dev, int n_channels, double beta,
                     int local_kind, const double* local_scale_host, double step_size, int steps,
                     const hepmc_elm_t* elm, int precision, int coop, const double* x0_dev,
                     int64_t n_chains, int64_t first_chain, int64_t sample_size, int64_t thin,
                     uint64_t seed, uint32_t stream_id,
                     const int64_t* sel_replay_dev, const double* draw_replay_dev, const double* u_replay_dev,
                     double* chain_dev, double* last_dev, int64_t* accepted_dev, int64_t* total_accepted_dev,
                     int64_t* local_steps_dev, void* stream);

/* ---- ELM surrogate: replaces FunctionBasis.eval (extreme_learning.py:18-29),
 * RadialBasis.output_matrix/eval_gradient (:190-207,222-230), AdditiveBasis.* (:120-146) */
/* xs_dev (n, ndim).  what: 0 = eval -> (n,), 1 = eval_gradient -> (n, ndim),
 * 2 = output_matrix -> (n, nodes).  precision as in hepmc_hmc_chains. */
int hepmc_elm_eval(const hepmc_elm_t* elm, int ndim, int what, int precision,
                   const double* xs_dev, int64_t n, double* out_dev, void* stream);

/* ---- sample statistics: replaces auto_cov / auto_corr / effective_sample_size / bin_wise_chi2
 * (core/util/stat_tools.py:11-129), the reductions Sample.__repr__ shows
 * (core/sampling.py:79-94).  data_dev is (n_chains, n_steps, ndim) row-major; mean_dev and
 * var_dev are (ndim,) -- the KNOWN target moments, as in the reference. -------------------- */
/* out_dev (n_chains, n_lags, ndim): lag-k autocovariance sum_t (x_t-mean)(x_{t+k}-mean) / n_steps
 * for k in [lag_begin, lag_begin + n_lags); lag 0 is var (stat_tools.py:17-37). */
int hepmc_autocov(const double* data_dev, int64_t n_chains, int64_t n_steps, int ndim,
                  const double* mean_dev, const double* var_dev, int64_t lag_begin, int64_t n_lags,
                  double* out_dev, void* stream);
/* ess_dev (n_chains, ndim) = n / (1 + 2 sum_lag (1 - lag/n) rho_lag), rho_lag the unbiased
 * autocorrelation, summed from lag 1 while rho >= 0.05 and lag < n - 1 (stat_tools.py:46-70).
 * lag_dev (n_chains, ndim) or NULL receives the lag at which each sum stopped.  n_steps >= 2. */
int hepmc_effective_sample_size(const double* data_dev, int64_t n_chains, int64_t n_steps, int ndim,
                                const double* mean_dev, const double* var_dev,
                                double* ess_dev, int64_t* lag_dev, void* stream);
/* np.histogramdd as bin_wise_chi2 calls it (stat_tools.py:108): edges_dev holds the ndim edge
 * arrays back to back (nbins_host[d] + 1 doubles each, increasing); counts_dev has
 * prod(nbins) entries in C order and is zeroed here.  A point on the last edge belongs to
 * the last bin; points outside (or NaN) are dropped. */
int hepmc_histogramdd(const double* data_dev, int64_t n, int ndim, const double* edges_dev,
                      const int* nbins_host, unsigned long long* counts_dev, void* stream);
/* int_steps uniform points inside each of the m bins flat_dev[j] (C-order flat indices):
 * out_dev (m * int_steps, ndim), x = low + (high - low) u (stat_tools.py:122-123).
 * u_replay_dev (m, int_steps, ndim) or NULL (native generator). */
int hepmc_bin_points(int ndim, const double* edges_dev, const int* nbins_host,
                     const int64_t* flat_dev, int64_t m, int int_steps,
                     uint64_t seed, uint32_t stream_id, const double* u_replay_dev,
                     double* out_dev, void* stream);
/* expected_dev[j] = mean_s pdf_dev[j, s] * sample_size * bin_volume; out_dev[0] =
 * sum over bins with expected >= min_count of (count - expected)^2 / expected (what
 * scipy.stats.chisquare returns, stat_tools.py:125-127), out_dev[1] = number of such bins. */
int hepmc_bin_chi2(const double* pdf_dev, const int64_t* flat_dev, const unsigned long long* counts_dev,
                   int64_t m, int int_steps, double sample_size, double bin_volume, double min_count,
                   double* expected_dev, double* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HEPMC_B200_H */
