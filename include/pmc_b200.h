/*
 * pmc_b200.h -- C ABI of libpmc_b200.so: the B200 (sm_100a) implementation of pocoMC's
 * data-parallel hot path (normalizing-flow preconditioner + vectorised SMC/MCMC inner loop).
 *
 * The reference (minaskar/pocomc v1.2.6) is pure Python and has no FFI; its boundary for this
 * path is Python duck typing at three seams.  Each entry point below
 * replaces the array maths of the cited reference lines; the host-side mirror in pocomc_b200/
 * keeps the reference's Python names and argument meaning and binds these symbols with ctypes.
 *
 * Conventions
 *  - plain C: raw device pointers + sizes; the caller owns every buffer (no hidden allocation);
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), returns 0 on
 *    success or a non-zero code with a message available from pmc_last_error();
 *  - SMC state is float64, flow tensors float32 (reference precision regime);
 *  - row-major [N, D] matrices; N = particles (int64), D = dimensions (int32).
 */
#ifndef PMC_B200_H
#define PMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* pmc_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------------ */
const char* pmc_last_error(void);
int pmc_version(void);
/* number of SMs / device ordinal the library sees (diagnostics, grid sizing) */
int pmc_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ---- normalizing flow: zuko MAF / NSF as reached from pocomc/flow.py:97-163 ----------------
 * `meta` (int32, host-built by pocomc_b200.made_layout) describes the degree-sorted slab layout;
 * `packed` is the flow's parameters in that layout.                                           */

/* packed[i] = gather[i] >= 0 ? raw[gather[i]] : 0   (re-run after every weight update) */
int pmc_flow_pack(const float* raw, const int32_t* gather, float* packed, int64_t n_packed,
                  pmc_stream_t stream);

/* Flow.forward (flow.py:99-114 -> transform.call_and_ladj): data -> latent, ladj = log|dz/dx|.
 * Flow.inverse (flow.py:116-132 -> transform.inv.call_and_ladj): latent -> data,
 * ladj = log|dx/dz|.  One degree-ordered sweep per transform replaces the reference's D+1
 * hyper-network passes.  in/out [N, D] f32 (may alias), ladj [N] f32.  `meta` is the device
 * copy of the table, `meta_host` the same table in host memory (launch geometry).              */
int pmc_flow_sweep(const float* packed, const int32_t* meta, const int32_t* meta_host,
                   int32_t meta_len, const float* in, float* out, float* ladj, int64_t n,
                   int32_t inverse, pmc_stream_t stream);

/* Flow.log_prob (flow.py:134-147): N(0,I).log_prob(z) + ladj from a forward sweep's outputs.   */
int pmc_flow_base_logprob(const float* z, const float* ladj, float* logprob, int64_t n, int32_t d,
                          pmc_stream_t stream);

/* ---- tensor-core (tcgen05) dense forward: Flow.forward / Flow.log_prob / training forward ----
 * (flow.py:99-114,134-147 -> zuko transform.call_and_ladj: one masked-MLP pass per transform.)
 * `packed` is the TF32 hi/lo weight image described by pocomc_b200.made_layout.build_tc; it is
 * produced from the flat parameter blob by pmc_flow_tc_pack (gather codes: g >= 0 hi(raw[g]),
 * -(g+2) lo(raw[g]), g | 2^30 plain copy, -1 zero).  passes = 3: 3xTF32 split, fp32 fidelity;
 * passes = 1: plain TF32.  in/out [N, D] f32, ladj [N] f32; `meta_host` is a HOST table.        */
int pmc_flow_tc_pack(const float* raw, const int32_t* gather, float* packed, int64_t n_packed,
                     pmc_stream_t stream);
int pmc_flow_forward_tc(const float* packed, const int32_t* meta_host, int32_t meta_len,
                        const float* in, float* out, float* ladj, int64_t n, int32_t passes,
                        pmc_stream_t stream);

/* ---- tensor-core (tcgen05) block-triangular sweep: Flow.inverse (flow.py:116-132 -> zuko
 * transform.inv.call_and_ladj, the hot call of pocomc/mcmc.py:88,256) and Flow.forward of affine (zuko MAF,
 * flow.py:54-63) and 8-bin spline (zuko NSF, flow.py:65-86: the reference's default presets) flows of any preset
 * width (D = 8 .. 200, H = 32 .. 1024); the table's TRI_KIND field selects the univariate head.
 * Order positions are processed in blocks of 4, blocks in windows whose per-unit accumulators fit tensor
 * memory: the dense dependence on earlier blocks runs as tcgen05.mma (right-looking updates inside a
 * window, a left-looking initialisation from a scratch area when a window starts; 3xTF32 split,
 * passes = 3; plain TF32, passes = 1), the dependence inside a block as fp32 FMAs.  `packed` is the image
 * described by pocomc_b200.tri_layout.build_tri, produced by pmc_flow_tc_pack; `meta_host` is the table in
 * HOST memory (launch geometry, validation), `meta_dev` the same table in device memory (read by the
 * kernel).  in/out [N, D] f32 (may alias), ladj [N] f32.  `workspace`: device scratch of at least
 * pmc_flow_sweep_tri_workspace(meta_host, meta_len, n) floats (0 for flows that fit one window; the
 * pointer may then be NULL); the query returns -1 on a bad table.                                   */
int64_t pmc_flow_sweep_tri_workspace(const int32_t* meta_host, int32_t meta_len, int64_t n);
int pmc_flow_sweep_tri(const float* packed, const int32_t* meta_host, const int32_t* meta_dev,
                       int32_t meta_len, const float* in, float* out, float* ladj, int64_t n,
                       int32_t inverse, int32_t passes, float* workspace, int64_t workspace_floats,
                       pmc_stream_t stream);

/* ---- Flow.fit optimiser step (flow.py:268,314-319) -------------------------------------------
 * torch.nn.utils.clip_grad_norm_(max_norm = hyper[5]; <= 0 disables) followed by
 * torch.optim.AdamW.step (amsgrad off) over the flat parameter blob, two launches.  `hyper` is a
 * DEVICE array of 6 doubles {lr, beta1, beta2, eps, weight_decay, clip}; `step` a device int64 that
 * the call increments (AdamW's t) -- both device-resident so the call can be replayed from a CUDA
 * graph while the learning rate changes.  scratch: pmc_adamw_scratch_size() doubles.
 * gnorm_out (may be NULL) receives the un-clipped gradient norm.                                 */
int64_t pmc_adamw_scratch_size(void);
int pmc_adamw_clip_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                        const double* hyper, int64_t* step, double* scratch, float* gnorm_out,
                        pmc_stream_t stream);
/* Same update with the per-batch bookkeeping of the training loop (flow.py:301-322) folded in, so that one
 * optimiser step is four launches: (i) *loss_acc += sum of loss_partials[0..n_loss) in index order and
 * *cursor += 1 (either may be NULL); (ii) every updated parameter i is also written to image[pos_a[i]] and
 * image[pos_b[i]] (negative = none) -- the forward and backward weight images of build_train -- which
 * replaces the pmc_flow_pack pass between steps (image may be NULL).                                    */
int pmc_adamw_clip_step_ex(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                           const double* hyper, int64_t* step, double* scratch, float* gnorm_out,
                           const double* loss_partials, int32_t n_loss, double* loss_acc, int64_t* cursor,
                           const int32_t* pos_a, const int32_t* pos_b, float* image, pmc_stream_t stream);

/* ---- fused training step of Flow.fit (flow.py:301-319) for MAF ---------------------------------
 * One mini-batch: weighted negative log-likelihood (flow.py:305-310) and, if `backward`, its gradient
 * with respect to every flow parameter, written to the flat blob `grad` (masked entries untouched:
 * keep the buffer zero-initialised).  `packed` = fp32 weight images described by
 * pocomc_b200.made_layout.build_train (built from the blob by pmc_flow_pack with that layout's gather
 * table); `meta_host`, HOST table of the same layout; `tiles` / `wmap` device copies of its tile list and
 * gradient scatter maps.  Batch b = *cursor takes rows idx[b*bp .. b*bp+bp) of the training matrix
 * xdata [rows, D] (wdata [rows] sample weights or NULL), mask = 0 marks padding rows; bp % 32 == 0.
 * loss_partials receives bp/8 partial sums, one per 8 consecutive rows whatever tile the kernel picks
 * (8, 16 or 32 rows per CTA); loss = their sum, in index order; logprob (may be NULL)
 * the per-row log-probability.  scratch: pmc_flow_train_scratch_size(meta_host, bp) floats.          */
int64_t pmc_flow_train_scratch_size(const int32_t* meta_host, int64_t bp);
int pmc_flow_train_step(const float* packed, const int32_t* meta_host, int32_t meta_len,
                        const float* xdata, const float* wdata, const int64_t* idx, const float* mask,
                        const int64_t* cursor, int64_t bp, float* scratch, double* loss_partials,
                        float* logprob, const int32_t* tiles, const int32_t* wmap, float* grad,
                        int32_t backward, pmc_stream_t stream);
/* Validation pass of one epoch (flow.py:326-342) in ONE launch: the weighted negative log-likelihood of the
 * n_batches consecutive batches *cursor .. *cursor + n_batches - 1 of the same idx / mask tables, every batch
 * normalised by its own weight sum like the reference's per-batch loss.  loss_partials receives
 * n_batches * bp/8 partial sums (batch-major; sum of batch b = sum of its bp/8 entries); logprob (may be NULL)
 * n_batches * bp per-row log-probabilities.  No scratch: nothing is kept for a backward pass.            */
int pmc_flow_eval_batches(const float* packed, const int32_t* meta_host, int32_t meta_len,
                          const float* xdata, const float* wdata, const int64_t* idx, const float* mask,
                          const int64_t* cursor, int64_t bp, int64_t n_batches, double* loss_partials,
                          float* logprob, pmc_stream_t stream);

/* ---- MCMC controller state ------------------------------------------------------------------
 * Device-resident f64 block shared by the step kernels so a whole MCMC step is host-sync free:
 * ctl[PMC_CTL_*] scalars followed by mu[D] at ctl[PMC_CTL_MU].                                 */
enum {
  PMC_CTL_SIGMA = 0,      /* proposal scale sigma (mcmc.py:54,152) */
  PMC_CTL_STEP = 1,       /* i, steps done */
  PMC_CTL_BEST = 2,       /* logp2_val (mcmc.py:70,170-173) */
  PMC_CTL_CNT = 3,        /* cnt, steps without improvement */
  PMC_CTL_STOP = 4,       /* 1 once the plateau rule or n_max fired */
  PMC_CTL_ACCEPT = 5,     /* mean(alpha) of the last step */
  PMC_CTL_CALLS = 6,      /* likelihood calls so far (sum of finite masks) */
  PMC_CTL_TRACK = 7,      /* mean(logl+logp[+logdetj]) of the last step */
  PMC_CTL_NACC = 8,       /* number accepted in the last step */
  PMC_CTL_MU = 16         /* mu[D] (tpCN mean, mcmc.py:63,156) */
};

/* MCMC kernel kinds (pocomc/mcmc.py) */
enum {
  PMC_KIND_TPCN_FLOW = 0, /* preconditioned_pcn  mcmc.py:8-183   */
  PMC_KIND_RWM_FLOW = 1,  /* preconditioned_rwm  mcmc.py:186-341 */
  PMC_KIND_TPCN = 2,      /* pcn                 mcmc.py:344-506 */
  PMC_KIND_RWM = 3        /* rwm                 mcmc.py:508-654 */
};

/* t-preconditioned Crank-Nicolson proposal (mcmc.py:77-85 / 409-417) plus both Mahalanobis
 * distances needed by the acceptance factors A, B (mcmc.py:124-129).
 *  pos      : current position, f32 theta [N,D] (pos_is_f32=1, flow kernels) or f64 u [N,D]
 *  ctl      : reads sigma and mu[D]
 *  inv_cov_t, chol_t : TRANSPOSED D x D f64 matrices (np.linalg.inv / cholesky of t_cov)
 *  g [N] standard gamma((D+nu)/2) draws, z [N,D] standard normals (explicit noise)
 *  out: prop64 [N,D], prop32 [N,D] (may be NULL), m_cur [N], m_prop [N]                        */
int pmc_tpcn_propose(int32_t pos_is_f32, const void* pos, const double* ctl, const double* inv_cov_t,
                     const double* chol_t, double nu, const double* g, const double* z,
                     double* prop64, float* prop32, double* m_cur, double* m_prop, int64_t n,
                     int32_t d, pmc_stream_t stream);

/* Random-walk proposal pos + sigma * chol z (mcmc.py:251-253 / 569-571). */
int pmc_rwm_propose(int32_t pos_is_f32, const void* pos, const double* ctl, const double* chol_t,
                    const double* z, double* prop64, float* prop32, int64_t n, int32_t d,
                    pmc_stream_t stream);

/* Reparameterize (pocomc/scaler.py). Per-dimension parameter block, all [D]:
 *  kind : 0 none, 1 left (low only), 2 right (high only), 3 both      (scaler.py:459-490)
 *  bc   : bit0 periodic, bit1 reflective                              (scaler.py:84-157)      */
typedef struct pmc_scaler {
  const int32_t* kind;
  const int32_t* bc;      /* may be NULL: no boundary conditions */
  const double* low;
  const double* high;
  const double* mu;
  const double* sigma;
  double log_sigma_sum;   /* np.sum(np.log(sigma)) (scaler.py:309) */
  int32_t logit;          /* 1: logit, 0: probit */
  int32_t scale;          /* 1: apply the affine part */
} pmc_scaler;

/* Reparameterize.inverse (scaler.py:204-226): u -> (x, log|dx/du|), with the optional boundary
 * wrap + re-forward + re-inverse of mcmc.py:94-97.  u_in f32 or f64 [N,D]; outputs u_out f64
 * (the possibly wrapped u), x f64 [N,D], logdetj f64 [N], finite u8 [N] =
 * isfinite(logdetj) & all(isfinite(x)) (mcmc.py:100-102).                                      */
int pmc_scaler_inverse(int32_t u_is_f32, const void* u_in, const pmc_scaler* sc, double* u_out,
                       double* x, double* logdetj, uint8_t* finite, int64_t n, int32_t d,
                       pmc_stream_t stream);

/* pmc_scaler_inverse followed by pmc_logprior (below) on the x it produced, in one launch: what an MCMC step needs between
 * the flow pull-back and the likelihood call (mcmc.py:91-109) when the prior is a product of norm / uniform factors.
 * logp [N] f64; finite also drops the rows whose log-prior is not finite.  Same numbers as the two separate calls.   */
int pmc_scaler_inverse_prior(int32_t u_is_f32, const void* u_in, const pmc_scaler* sc, const int32_t* prior_kind,
                             const double* prior_loc, const double* prior_scale, double* u_out, double* x,
                             double* logdetj, uint8_t* finite, double* logp, int64_t n, int32_t d,
                             pmc_stream_t stream);

/* Reparameterize.forward (scaler.py:180-202): x -> u (no bounds check; host validates). */
int pmc_scaler_forward(const double* x, const pmc_scaler* sc, double* u, int64_t n, int32_t d,
                       pmc_stream_t stream);

/* Reparameterize.apply_boundary_conditions_x (scaler.py:84-157), in place on x [N,D]:
 * periodic wrap (bc bit0) then reflection (bc bit1) with the reference's while-loop arithmetic. */
int pmc_apply_bc(double* x, const int32_t* bc, const double* low, const double* high, int64_t n,
                 int32_t d, pmc_stream_t stream);

/* Metropolis accept / masked state update / per-block partial sums
 * (mcmc.py:124-149 and the reductions of :152,156,170; same for the other three kernels).
 * State (updated in place): pos (f32 theta for flow kinds, else NULL -- u is the position),
 * u,x [N,D] f64, logdetj, logl, logp [N] f64, logdetj_flow [N] f32 (flow kinds).
 * Proposal: prop64, u_p, x_p, logdetj_p, logl_p, logp_p, logdetj_flow_p, m_cur, m_prop.
 * r [N] uniforms; finite_count_in = this step's likelihood calls (host-known) or -1 to count
 * finite_mask on device.  partials: [n_blocks, D+4] f64 scratch (pmc_mh_partials_size).       */
int64_t pmc_mh_partials_size(int64_t n, int32_t d);
int pmc_mh_accept_update(int32_t kind, double beta, double nu, float* pos32, double* u, double* x,
                         double* logdetj, double* logl, double* logp, float* logdetj_flow,
                         const double* prop64, const double* u_p, const double* x_p,
                         const double* logdetj_p, const double* logl_p, const double* logp_p,
                         const float* logdetj_flow_p, const double* m_cur, const double* m_prop,
                         const double* r, const uint8_t* finite, double* alpha_out,
                         double* partials, int64_t n, int32_t d, pmc_stream_t stream);

/* pmc_mh_accept_update and pmc_mcmc_finalize in ONE launch (single-GPU runs): the thread block that finishes last
 * reduces the block partials and adapts the controller.  The launch is a no-op once ctl[PMC_CTL_STOP] is set, so
 * the host may queue several MCMC steps before reading the controller back and still stop exactly where
 * mcmc.py:170-180 stops.  `ticket`: one device uint32, zero before the first launch (the kernel resets it).        */
int pmc_mh_accept_finalize(int32_t kind, double beta, double nu, float* pos32, double* u, double* x,
                           double* logdetj, double* logl, double* logp, float* logdetj_flow,
                           const double* prop64, const double* u_p, const double* x_p,
                           const double* logdetj_p, const double* logl_p, const double* logp_p,
                           const float* logdetj_flow_p, const double* m_cur, const double* m_prop,
                           const double* r, const uint8_t* finite, double* alpha_out, double* partials,
                           double* ctl, uint32_t* ticket, int32_t mean_mode, int32_t n_steps,
                           int32_t n_max, int64_t n, int32_t d, pmc_stream_t stream);

/* ---- layer-wise training step of Flow.fit (flow.py:301-319) for every flow shape: spline (NSF, the reference's default
 * presets) and affine heads, any width.  raw / mask / grad: flat parameter blob in module order (per transform W0 [H,D], b0,
 * W1 [H,H], b1, W2, b2, W3 [D*total,H], b3), its MADE mask laid out alike (1 for biases), its gradient.  x [*, D] f32 training
 * matrix, w (NULL: unweighted), idx_all / mask_all [*, B] batch tables, cursor: device int64 selecting the batch.  kind 0 =
 * affine heads (2 parameters per feature), 1 = spline heads (23 = 3 * 8 bins - 1).  partials:
 * pmc_flow_train_lw_partials(B) doubles whose sum is the batch loss (flow.py:305-310).  train = 0: loss only.  scratch:
 * pmc_flow_train_lw_scratch_size(D, H, T, total, numel, B) floats.  A fixed sequence of launches on `stream` (graph-capturable). */
int64_t pmc_flow_train_lw_scratch_size(int32_t D, int32_t H, int32_t T, int32_t total, int64_t numel, int64_t B);
int32_t pmc_flow_train_lw_partials(int64_t B);
int pmc_flow_train_step_lw(const float* raw, const float* mask, int32_t D, int32_t H, int32_t T, int32_t kind, int64_t numel,
                           const float* x, const float* w, const int64_t* idx_all, const float* mask_all, const int64_t* cursor,
                           int64_t B, float* scratch, double* partials, float* grad, int32_t train, pmc_stream_t stream);

/* ---- proposal geometry (geometry.py:31-59, student.py:5-85) ------------------------------------------------
 * The O(n D^2) reductions of Geometry.fit / fit_mvstud over a cloud x [n, d] f64 (row-major), f64, fixed-order two-stage sums
 * (independent of the SM count).  scratch: pmc_geometry_scratch_size(n, d) doubles.
 *   pmc_weighted_colsums : out[0..d) = sum_i w_i x_i, out[d] = sum w, out[d+1] = sum w^2, out[d+2] = max_i |x_i - center|^2
 *                          (np.average / np.mean numerators; w == NULL: unit weights; center == NULL: no max)
 *   pmc_weighted_scatter : C [d,d] = sum_i w_i (x_i - center)(x_i - center)^T  (np.cov numerators, the EM scatter update)
 *   pmc_mahalanobis      : delta [n] = (x_i - center)^T P (x_i - center), P = Sigma^-1 [d,d]        (student.py:40)
 *   pmc_student_weights  : w_i = (nu + dim) / (nu + delta_i); out2 = (sum log w_i, sum w_i); w_out may be NULL (student.py:42-56) */
int64_t pmc_geometry_scratch_size(int64_t n, int32_t d);
int pmc_weighted_colsums(const double* x, const double* w, const double* center, int64_t n, int32_t d, double* scratch,
                         double* out, pmc_stream_t stream);
int pmc_weighted_scatter(const double* x, const double* w, const double* center, int64_t n, int32_t d, double* scratch,
                         double* C, pmc_stream_t stream);
int pmc_mahalanobis(const double* x, const double* center, const double* P, int64_t n, int32_t d, double* delta,
                    pmc_stream_t stream);
int pmc_student_weights(const double* delta, int64_t n, double nu, double dim, double* w_out, double* scratch, double* out2,
                        pmc_stream_t stream);

/* ---- peer-memory exchange of a particle-sharded run (one process per GPU, all on one NVLink / NVSwitch node) --------
 * Replaces "all-gather the block partials, then pmc_mcmc_finalize" (mcmc.py:152-180 needs the mean acceptance, the mean
 * of theta and the tracked mean log-density over ALL particles every step) by stores into peer memory from inside the
 * accept kernel.  pmc_comm_create allocates this rank's exchange buffer (2 x capacity_doubles + flags) on the current
 * device and returns its CUDA IPC handle (PMC_COMM_HANDLE_BYTES bytes) -- exchange the handles of all ranks with any
 * host-side collective (torch.distributed.all_gather) -- pmc_comm_connect opens the peers' buffers; `block_off`
 * [world + 1] = first 256-row block of every rank in global particle order (pmc_comm_set_blocks changes it later).
 * pmc_comm_error: 1 after a peer failed to publish within 20 s (the kernel then sets the stop flag), -1 bad context. */
#define PMC_COMM_MAX_RANKS 8
#define PMC_COMM_HANDLE_BYTES 64
int pmc_comm_create(int32_t rank, int32_t world, int64_t capacity_doubles, void** comm_out, unsigned char* handle64);
int pmc_comm_connect(void* comm, const unsigned char* handles, const int32_t* block_off);
int pmc_comm_set_blocks(void* comm, const int32_t* block_off, pmc_stream_t stream);
int pmc_comm_error(void* comm);
int pmc_comm_destroy(void* comm);
/* pmc_mh_accept_finalize for a sharded run, ONE launch per MCMC step and rank: Metropolis update of this rank's n rows,
 * block partials, and in the block that finishes last: push the partials into every peer's buffer over NVLink, publish
 * an epoch flag, wait for all peers, adapt the controller from the rank-ordered partials of all n_global particles
 * (mean_mode 0: the GPU-count independent f64 block sums).  Every rank must launch it for every step.               */
int pmc_mh_accept_finalize_p2p(int32_t kind, double beta, double nu, float* pos32, double* u, double* x,
                               double* logdetj, double* logl, double* logp, float* logdetj_flow,
                               const double* prop64, const double* u_p, const double* x_p,
                               const double* logdetj_p, const double* logl_p, const double* logp_p,
                               const float* logdetj_flow_p, const double* m_cur, const double* m_prop,
                               const double* r, const uint8_t* finite, double* alpha_out, double* partials,
                               double* ctl, uint32_t* ticket, int32_t n_steps, int32_t n_max,
                               int64_t n, int32_t d, void* comm, int64_t n_global, pmc_stream_t stream);

/* Scalar adaptation + stop rule (mcmc.py:152-180 and the three siblings), on device:
 * reduces `partials` in fixed order, updates ctl (sigma, mu, step, best, cnt, stop, accept...).
 * mean_mode 1 reproduces np.mean(theta f32, axis=0)'s sequential f32 accumulation exactly
 * (needs pos32), 0 uses the f64 block partials.  `partials` holds n_blocks rows of D+4 and `n`
 * is the number of particles they cover: for a sharded run pass the rank-ordered all-gather of
 * every rank's partials and the GLOBAL particle count (n_blocks <= 0: pmc_mh_partials_size(n,d)/(d+4)). */
int pmc_mcmc_finalize(int32_t kind, double* ctl, const double* partials, int64_t n_blocks,
                      const float* pos32, int32_t mean_mode, int32_t n_steps, int32_t n_max, int64_t n,
                      int32_t d, pmc_stream_t stream);

/* Counter-based device RNG (Philox4x32-10) for throughput mode: fills the explicit noise tensors
 * the kernels above consume.  Keyed by (seed, step, particle, stream id) -> independent of the
 * number of GPUs.  gamma_shape <= 0 skips g.                                                   */
int pmc_rng_fill(uint64_t seed, uint64_t step, int64_t particle_offset, double gamma_shape,
                 double* g, double* z, double* r, int64_t n, int32_t d, pmc_stream_t stream);
/* same draws with the step counter read on the device: step = ctl[PMC_CTL_STEP] + 1 (lets the host queue steps) */
int pmc_rng_fill_ctl(uint64_t seed, const double* ctl, int64_t particle_offset, double gamma_shape,
                     double* g, double* z, double* r, int64_t n, int32_t d, pmc_stream_t stream);

/* ---- persistent-sampling weights (pocomc/particles.py:215-231, tools.py:56-93) --------------
 * History logl [T, N] f64 and the running log-denominator den [T, N] =
 * logaddexp_i(beta_i * logl - logz_i) (un-normalised).  ps_append folds iteration(s) into den in
 * the reference's summation order: rows [0,t_new) get the new terms i in [t_new, t_total);
 * rows [t_new, t_total) get all terms.  beta/logz are device arrays [t_total].               */
int pmc_ps_append(const double* logl, double* den, const double* beta, const double* logz,
                  int32_t t_new, int32_t t_total, int64_t n, pmc_stream_t stream);

/* One probe of Sampler._reweight's get_weights_and_ess (sampler.py:739-746): for
 * logw = beta_f*logl - (den - log T) over all M = T*N elements returns
 * out[0]=max logw, out[1]=sum e, out[2]=sum e^2 with e = exp(logw - max), out[3] = USS sum
 * (sum 1-(1-e/sum e)^k, only if uss_k > 0).  scratch: pmc_ps_scratch_size(M) doubles.          */
int64_t pmc_ps_scratch_size(int64_t m);
int pmc_ps_reduce(const double* logl, const double* den, double beta_f, int32_t t_total, int64_t n,
                  int64_t uss_k, double* scratch, double* out4, pmc_stream_t stream);

/* Normalised weights w = exp(logw - max)/sum and optionally logw (normalised like
 * compute_logw_and_logz(normalize=True)); stats4 = output of pmc_ps_reduce (device).           */
int pmc_ps_weights(const double* logl, const double* den, double beta_f, int32_t t_total, int64_t n,
                   const double* stats4, double* w, double* logw, pmc_stream_t stream);

/* effective_sample_size / unique_sample_size on an explicit weight vector (tools.py:56-93):
 * out3 = [sum w, sum w^2, sum 1-(1-w/sum w)^k (if uss_k > 0)]; ESS = out[0]^2 / out[1].
 * scratch: pmc_ps_scratch_size(m) doubles.                                                      */
int pmc_weight_stats(const double* w, int64_t m, int64_t uss_k, double* scratch, double* out3,
                     pmc_stream_t stream);

/* ---- resampling (sampler.py:702-713, tools.py:136-186) --------------------------------------
 * Sequential f64 inclusive cumsum in index order (bit-identical to np.cumsum), then
 * multinomial: idx = searchsorted(cdf/cdf[-1], r, 'right'); systematic: the reference's walk
 * over the same sequential cumsum, idx[i] = first j with (u0+i)/n_out <= cdf[j].               */
int pmc_cumsum_f64(const double* w, double* cdf, int64_t m, pmc_stream_t stream);
int pmc_resample_multinomial(const double* cdf, const double* r, int64_t* idx, int64_t m,
                             int64_t n_out, pmc_stream_t stream);
int pmc_resample_systematic(const double* cdf, double u0, int64_t* idx, int64_t m, int64_t n_out,
                            pmc_stream_t stream);
/* dst[i, :] = src[idx[i], :] for f64 rows of width `d` (d=1 for the scalar arrays). */
int pmc_gather_rows_f64(const double* src, const int64_t* idx, double* dst, int64_t n_out,
                        int32_t d, pmc_stream_t stream);

/* ---- weight trimming (tools.py:10-53) --------------------------------------------------------
 * Given weights SORTED ascending (ws) evaluates every percentile grid point linspace(0,99,bins)
 * with np.percentile's linear interpolation and suffix sums, and returns the threshold of the
 * first grid point (from the top) whose trimmed-ESS ratio >= ess_frac.
 * out3: [threshold, kept sum, grid index].                                                     */
int pmc_trim_threshold(const double* ws, int64_t m, double ess_frac, int32_t bins, double* scratch,
                       double* out3, pmc_stream_t stream);
int64_t pmc_trim_scratch_size(int64_t m);

/* ---- evidence (sampler.py:907-913) ----------------------------------------------------------- */
/* logw = logl + logp + logdetj - logq ; out2 = [logsumexp(logw) - log n, max]                  */
int pmc_lse(const double* logw, int64_t n, double* scratch, double* out2, pmc_stream_t stream);
/* bootstrap: for b < n_boot: out[b] = logsumexp(logw[idx[b, :]]) - log n                        */
int pmc_lse_bootstrap(const double* logw, const int64_t* idx, int64_t n, int64_t n_boot,
                      double* out, pmc_stream_t stream);
/* the same bootstrap with the resampling indices drawn on the device (Philox4x32-10 keyed by seed, row, position):
 * no [n_boot, n] index matrix in host or device memory; n < 2^32.                      */
int pmc_lse_bootstrap_rng(const double* logw, int64_t n, int64_t n_boot, uint64_t seed, double* out,
                          pmc_stream_t stream);

/* ---- synthetic likelihoods / priors evaluated on device --------------------------------------
 * Used by bench.py's device-resident throughput arm and by the opt-in device fast path for
 * scipy.stats uniform / norm priors.  The user's log_likelihood stays a
 * host black box on the reference-facing path.                                                  */
enum { PMC_LIKE_GAUSS = 0, PMC_LIKE_ROSENBROCK = 1, PMC_LIKE_MIXTURE = 2, PMC_LIKE_FUNNEL = 3 };
/* gauss: -0.5 x^T P x + c0 with P^T passed as mat_t [D,D]; rosenbrock (README.md:53-55);
 * mixture: logaddexp(N(x;+c,s), N(x;-c,s)) - log 2 with p0=c, p1=s; funnel: p0 = sd of x0.     */
int pmc_loglike(int32_t which, const double* x, const uint8_t* finite, const double* mat_t,
                double p0, double p1, double* logl, int64_t n, int32_t d, pmc_stream_t stream);
/* product prior: per dim kind 0 = norm(loc, scale), 1 = uniform(loc, loc+scale); -inf outside.
 * Also ANDs isfinite(logp) into `finite` (mcmc.py:108-109).                                     */
int pmc_logprior(const double* x, uint8_t* finite, const int32_t* kind, const double* loc,
                 const double* scale, double* logp, int64_t n, int32_t d, pmc_stream_t stream);

/* ---- host staging of an MCMC step (mcmc.py:111-121: the likelihood is a host callable) ---------
 * x' [N,D] f64 leaves the GPU every step and logl' [N] returns.  pmc_download_rows queues, on `stream`, the copy of the
 * finite flags (may be NULL) followed by the rows of x' in n_chunks contiguous chunks of ceil(N / n_chunks) rows, and
 * records events[k] behind chunk k: the host waits for chunk k with pmc_event_synchronize and evaluates the likelihood of
 * those rows while the later chunks are still crossing the PCIe bus.  Host pointers must be page-locked.
 * pmc_memcpy_async: cudaMemcpyAsync(cudaMemcpyDefault) for the small per-step arrays (logl', the controller block).     */
int pmc_event_create(void** event_out);
int pmc_event_destroy(void* event);
int pmc_event_synchronize(void* event);
int pmc_stream_synchronize(pmc_stream_t stream);
int pmc_memcpy_async(void* dst, const void* src, int64_t bytes, pmc_stream_t stream);
int pmc_download_rows(const double* x_dev, double* x_host, const uint8_t* flag_dev, uint8_t* flag_host,
                      int64_t n, int32_t d, int32_t n_chunks, void* const* events, pmc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PMC_B200_H */
