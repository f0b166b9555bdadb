/*
 * tapqir_b200 -- C ABI of the B200-native cosmos SVI hot path.
 *
 * The reference (gelles-brandeis/tapqir) is pure Python and has no FFI of its own; the entry
 * points below are what a binding for this path would call, one per reference function /
 * method they replace (cited as file:line in the reference tree).  INTEGRATION.md shows the
 * ctypes stub a maintainer would add on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless marked "host";
 *   - nothing is allocated or freed inside; work is enqueued on `stream` (a cudaStream_t,
 *     0 = legacy default stream) and the call returns without synchronising;
 *   - `dtype` selects the arithmetic type of all floating-point buffers of the call:
 *     TQ_F32 (production) or TQ_F64 (the reference CLI's dtype, main.py:428);
 *   - return value: 0 on success; TQ_ERR_* otherwise, with tq_last_error() giving the text.
 *     The Python layer turns TQ_ERR_OOM into CudaOutOfMemoryError (exceptions.py:33-39) and
 *     anything else into RuntimeError;
 *   - re-entrant per stream, no global mutable state except the last-error string (thread local).
 *
 * Patch indexing: a "unit" is one (AOI, frame, channel) PxP patch.  Minibatch unit
 *   u = (ni * fb + fi) * C + c   reads dataset patch (ndx[ni], fdx[fi], c);  ndx / fdx == NULL
 *   mean the identity (0..nb-1 / 0..fb-1).  Per-unit arrays are laid out (nb, fb, C) and
 *   per-spot arrays (K, nb, fb, C), K = 2.
 */
#ifndef TAPQIR_B200_H
#define TAPQIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TQ_K 2  /* spots per patch (cosmos K) */
#define TQ_M 4  /* enumerated spot-presence configurations, 2^K */

enum { TQ_F32 = 0, TQ_F64 = 1 };
enum { TQ_PIX_U16 = 0, TQ_PIX_F32 = 1, TQ_PIX_F64 = 2 };
enum {
    TQ_OK = 0,
    TQ_ERR_ARG = 1,      /* bad argument (shape, dtype, NULL) */
    TQ_ERR_CUDA = 2,     /* CUDA runtime error other than OOM */
    TQ_ERR_OOM = 3,      /* cudaErrorMemoryAllocation */
    TQ_ERR_UNSUPPORTED = 4
};

/* Device-resident dataset view + minibatch selection (host struct, passed by pointer).
 * Replaces CosmosDataset.fetch (utils/dataset.py:140-151) and OffsetData (dataset.py:18-37). */
typedef struct {
    int32_t nb, fb, C;          /* minibatch AOIs, minibatch frames, channels */
    int32_t F;                  /* frames in the stored dataset (row stride of the AOI axis) */
    int32_t P;                  /* patch edge, 2 <= P <= 32 */
    int32_t O;                  /* offset bins */
    int32_t pixtype;            /* TQ_PIX_* of `pixels` */
    const int32_t* ndx;         /* (nb,) AOI indices into the store, or NULL */
    const int32_t* fdx;         /* (fb,) frame indices, or NULL */
    const void* pixels;         /* (Nt, F, C, P, P) */
    const void* xy;             /* (Nt, F, C, 2) target x, y; dtype */
    const uint8_t* is_ontarget; /* (Nt,) */
    const uint8_t* mask;        /* (Nt,) */
    const void* offset_samples; /* (O,) dtype */
    const void* offset_logits;  /* (O,) dtype: log of eps-clamped weights (dataset.py:27-29) */
} tq_patch_view;

int tq_version(void);
const char* tq_last_error(void);

/* gaussian_spots (distributions/util.py:15-64) for U target locations with K spots each.
 * height,width,x,y: (K, U); target_xy: (U, 2); m: (K, U) or NULL; out: (U, K, P, P). */
int tq_gaussian_spots(int dtype, int64_t U, int K, int P, const void* height, const void* width,
                      const void* x, const void* y, const void* target_xy, const void* m,
                      void* out, void* stream);

/* KSMOGN.log_prob (distributions/ksmogn.py:187-238) for NM spot-presence configurations.
 * height,width,x,y: (K, U); background: (U,); gain: device scalar; mcfg: (NM, K) device table
 * (NM in {1, 4}); logp out: (NM, U).  mcfg == NULL (with NM == 4) selects the built-in enumerated
 * {0,1}^K table of cosmos.py:419-425 and, for TQ_F32, the production kernel (csrc/ksmogn_fast.cuh). */
int tq_ksmogn_fwd(int dtype, const tq_patch_view* view, const void* height, const void* width,
                  const void* x, const void* y, const void* background, const void* gain,
                  const void* mcfg, int NM, void* logp, void* stream);

/* Reverse mode of the above (what autograd derives in the reference): given W = dLoss/dlogp
 * (NM, U) writes g_height.. g_y (K, U), g_background (U,) and g_rate (U,), the per-unit
 * derivative w.r.t. 1/gain (sum it and multiply by -1/gain^2 for d/dgain).  logp may be NULL. */
int tq_ksmogn_fwd_bwd(int dtype, const tq_patch_view* view, const void* height, const void* width,
                      const void* x, const void* y, const void* background, const void* gain,
                      const void* mcfg, int NM, const void* W, void* logp, void* g_height,
                      void* g_width, void* g_x, void* g_y, void* g_background, void* g_rate,
                      void* stream);


/* ---------------------------------------------------------------------------------------------
 * One cosmos SVI step = what `svi.step()` does in models/model.py:212 around the model/guide of
 * models/cosmos.py:82-462.  Call order on one stream:
 *
 *   tq_subsample (x2, optional) -> tq_cosmos_globals_sample -> tq_cosmos_sites -> tq_ksmogn_fwd_bwd
 *   -> tq_cosmos_local_post -> [all-reduce of `acc` across ranks] -> tq_cosmos_globals_grad
 *      (or tq_cosmos_globals_prepare any time after the sample + tq_cosmos_globals_finish here)
 *   -> tq_adam_dense (local buffer, global buffer) -> tq_step_advance
 *
 * Parameter buffers (unconstrained values, what pyro's param store holds; models/cosmos.py:471-598):
 *   local  flat (Nt = AOIs of this rank): background_mean_loc (Nt,1,C) | background_std_loc (Nt,1,C)
 *          | b_loc (Nt,F,C) | b_beta (Nt,F,C) | m_probs | h_loc | h_beta | w_mean | w_size | x_mean
 *          | y_mean | size, each (K,Nt,F,C)
 *   global flat: gain_loc | gain_beta | proximity_loc | proximity_size | pi_mean (Q,2) | pi_size (Q)
 *          | lamda_loc (Q) | lamda_beta (Q)        (Q = C channels, <= 4)
 * `mc` points to a host `tq_model_const`; `state` to a device uint64 iteration counter (Philox
 * stream id and Adam bias correction), so a captured CUDA graph can be replayed unchanged.
 * --------------------------------------------------------------------------------------------- */

/* Prior hyper-parameters (models/cosmos.py:55-64) + eps/tiny of the reference dtype (clamps of
 * torch's Categorical/Bernoulli/sigmoid and pyro's AffineBeta follow finfo(dtype)). */
typedef struct {
    double bg_mean_std, bg_std_std, lamda_rate, height_std, width_min, width_max, proximity_rate, gain_std;
    double eps, tiny;
    double logit_lim;   /* log((1 - eps) / eps) */
    int32_t P;
} tq_model_const;

int tq_sizeof_model_const(void);   /* sizeof(tq_model_const) as compiled */
int tq_sizeof_tables(void);        /* bytes of the device-side global tables blob */
int tq_sizeof_gstate(void);        /* bytes of the device-side global variates+samples blob */
int tq_site_record_rows(void);     /* rows NREC of the per-site record buffer (NREC, U) */
/* scratch of tq_cosmos_local_post for a minibatch of nb AOIs x fb frames x C channels, in doubles:
 * block_partial (plain scratch) and tickets (must be ZERO before the first call; left zeroed by every call) */
int64_t tq_local_post_scratch(int nb, int fb, int C);
int64_t tq_local_post_tickets(int nb, int fb, int C);

/* The AOI-local part of a step -- tq_cosmos_sites_ws + tq_ksmogn_fwd_bwd + tq_cosmos_local_post -- as ONE persistent
 * kernel (csrc/cosmos_fused.cu): guide sites (cosmos.py:393-462), rendered likelihood forward + reverse
 * (ksmogn.py:187-238, util.py:15-64), priors / (z, theta) sums / chain rule (cosmos.py:216-327 under TraceEnum_ELBO),
 * the per-unit intermediates staying in shared memory.  Same inputs and outputs as the three calls it replaces (`gain`:
 * the float written by tq_cosmos_globals_sample; noise_in: (9, U) base variates or NULL); samples_out (9, U) and L_out
 * (4, U) are optional copies of the guide samples / configuration log-likelihoods.  tq_cosmos_fused_supported says
 * whether a view qualifies (dtype float, P = 14, uint16 pixels, O <= 512); scratch as for tq_cosmos_local_post, sized
 * by tq_cosmos_fused_scratch (doubles) / tq_cosmos_fused_tickets (8-byte slots, zero before the first call). */
int tq_cosmos_fused_supported(int dtype, const tq_patch_view* view);
int64_t tq_cosmos_fused_scratch(int nb, int fb, int C);
int64_t tq_cosmos_fused_tickets(int nb, int fb, int C);
int tq_cosmos_fused_step(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc, const void* lparams,
                         const void* tables, const void* gain, int64_t aoi_offset, uint64_t seed,
                         const void* state, const void* noise_in, double sN, double sF, void* lgrads,
                         double* tickets, double* block_partial, double* acc, void* samples_out, void* L_out,
                         void* stream);

/* pyro.plate(subsample_size=n) [third party: randperm(size)[:n]]: uniform sample without
 * replacement by a partial Fisher-Yates on the persistent permutation `perm` (n_total int32,
 * initialise to arange once).  out: (n_pick,) int32.  stream_id separates independent draws. */
int tq_subsample(int n_total, int n_pick, uint64_t seed, const void* state, uint64_t stream_id,
                 void* perm, void* out, void* stream);
/* The AOI and the frame draw of one step in one launch (same keys as two tq_subsample calls: identical results); both
 * axes <= 16384 entries (tq_subsample_pair_supported). */
int tq_subsample_pair_supported(int n_total0, int n_total1);
int tq_subsample_pair(int n_total0, int n_pick0, uint64_t stream_id0, void* out0, int n_total1, int n_pick1,
                      uint64_t stream_id1, void* out1, uint64_t seed, const void* state, void* stream);

/* Guide samples of the four global sites (cosmos.py:342-368) and the prior tables derived from
 * them (distributions/util.py:67-173).  gparams: global flat buffer, ALWAYS float64 (as are ggrads and the
 * global Adam moments; `dtype` here only types gain_out -- a dozen scalars whose concentration-like gradients
 * cancel ~1e3-fold in the base variate, so fp32 storage alone would cost 1e-4 of gain_beta's gradient).  noise_in: base
 * variates (double; gain, proximity, pi (Q,2), lamda (Q)) for replay, or NULL to draw them with
 * Philox(seed, *state).  Outputs: gstate blob, tables blob, gain_out (1 value, dtype). */
int tq_cosmos_globals_sample(int dtype, int Q, const void* gparams, const void* mc,
                             const double* noise_in, uint64_t seed, const void* state,
                             double* gstate, void* tables, void* gain_out, void* stream);

/* Guide sites of every unit (cosmos.py:397-462): background, height/width/x/y per spot.
 * noise_in: (9, U) base variates (standard-gamma draws / (0,1) Beta draws) for replay, or NULL.
 * Outputs: samples (9, U); qm (4, U) = q(m) weights of the enumerated spot-presence configs;
 * rec (NREC, U) per-site log q, d log q/d sample and the linear maps of the reparameterised
 * gradient (csrc/cosmos_local.cuh).  Nt: AOIs held by this rank; aoi_offset: global index of AOI 0. */
int tq_cosmos_sites(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc,
                    const void* lparams, int64_t aoi_offset, uint64_t seed, const void* state,
                    const void* noise_in, void* samples, void* qm, void* rec, void* stream);

/* The same with a workspace for the sites that leave the fp32 production forms (a percent or so of a trained model's):
 * worklist: 9 * U 32-bit slots of device scratch -- it may alias any buffer that is only written LATER in the step, e.g.
 * the likelihood kernel's gradient output --, work_count: one 32-bit device counter.  Those sites are then redone in
 * double by dense warps instead of block by block.  (dtype double ignores the workspace.) */
int tq_cosmos_sites_ws(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc,
                       const void* lparams, int64_t aoi_offset, uint64_t seed, const void* state,
                       const void* noise_in, void* samples, void* qm, void* rec, void* worklist,
                       void* work_count, void* stream);

/* tq_cosmos_sites_ws for a FULL-BATCH float32 step (no ndx / fdx, nb == Nt, fb == F) with the previous step's dense Adam
 * update of the AOI-local parameters folded in (the reference's optimiser step, models/model.py:168-171,212, moved from
 * the end of step t to the first read of the parameters in step t + 1: the same values reach the same arithmetic).  The
 * dense update is pure HBM traffic, this kernel is issue-bound with the memory system idle: when StepState says an update
 * is pending (tq_step_advance_deferred) every site's thread applies it -- same formula, same constants, same bits as
 * tq_adam_dense -- to the parameters it owns before reading them.  lparams is read AND written; lgrads / exp_avg /
 * exp_avg_sq are the flat buffers tq_adam_dense would get.  Covered: the flat range [begin, end) of
 * tq_local_deferred_range (b_loc, b_beta, m_probs, h_loc, h_beta, w_mean, w_size, x_mean, y_mean); the rest (per-AOI
 * background parameters in front of it, `size` -- shared by the x and y sites -- behind it) stays with tq_adam_dense at
 * the end of the step.  Call order of such a step:
 *   ... -> tq_cosmos_sites_adam -> tq_ksmogn_fwd_bwd -> tq_cosmos_local_post -> ... -> tq_adam_dense ([0, begin) and
 *   [end, numel)) -> tq_step_advance_deferred
 * and tq_adam_deferred_flush before anything else reads the parameters or the moments. */
int tq_cosmos_sites_adam(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc,
                         void* lparams, int64_t aoi_offset, uint64_t seed, const void* state,
                         const void* noise_in, void* samples, void* qm, void* rec, void* worklist,
                         void* work_count, const void* lgrads, void* exp_avg, void* exp_avg_sq,
                         double beta1, double beta2, double eps, void* stream);
int tq_local_deferred_range(int64_t Nt, int64_t F, int64_t C, int64_t* begin, int64_t* end);

/* Priors, (z, theta) log-sum-exp, q(m)-weighted ELBO summand and its reverse mode per unit
 * (TraceEnum_ELBO [third party] on cosmos.py:216-327).  Inputs: the buffers above plus L (4, U),
 * gs (9, U) and g_rate (U,) from tq_ksmogn_fwd_bwd with W = qm.  sN = Nt_total / nb_total and
 * sF = F / fb are the plate scales.  Outputs: lgrads (local flat layout; entries of the minibatch
 * units and of their AOIs), acc (C, 18) per-channel sums for the globals (double);
 * tickets (tq_local_post_tickets doubles, zero-initialised once by the caller) and block_partial
 * (tq_local_post_scratch doubles) are scratch: the cross-unit sums (channel accumulators, AOI-level gradients)
 * are reduced inside the launch in a fixed order by the last block to finish. */
int tq_cosmos_local_post(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc,
                         const void* lparams, const void* tables, const void* samples,
                         const void* rec, const void* L, const void* gs, const void* g_rate,
                         double sN, double sF, void* lgrads, double* tickets,
                         double* block_partial, double* acc, void* stream);

/* Posterior of the enumerated latents for one guide draw (cosmos.compute_probs, cosmos.py:609-672):
 * z_probs (nb, fb, C, 2) += weight * p(z | ...), theta_probs (K, nb, fb, C) += weight * p(theta = k+1 | ...),
 * from the samples of tq_cosmos_sites and the tables of tq_cosmos_globals_sample (data term hidden,
 * average over m with q(m)).  Call once per particle with weight = 1 / particles. */
int tq_cosmos_zprobs(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc,
                     const void* lparams, const void* tables, const void* samples, double weight,
                     void* z_probs, void* theta_probs, void* stream);

/* Global sites: ELBO terms and reverse mode from `acc` (summed over ranks) to the global flat
 * gradient.  elbo_parts: (2 + 2Q,) scratch; loss: 1 double = -ELBO of the step. */
int tq_cosmos_globals_grad(int dtype, int Q, const void* gparams, const void* mc,
                           const double* gstate, const double* acc, double sN, double sF,
                           void* ggrads, double* elbo_parts, double* loss, void* stream);

/* The same reverse mode in two halves, so that its expensive part leaves the critical path: the gradient
 * of a global site is affine in one or two linear functionals of `acc`; tq_cosmos_globals_prepare
 * (right after tq_cosmos_globals_sample, e.g. on a side stream) evaluates the densities and implicit
 * reparameterisation gradients into `gprep` (tq_sizeof_gprep() bytes of device memory),
 * tq_cosmos_globals_finish (after `acc` is complete) writes ggrads and loss. */
int tq_sizeof_gprep(void);
int tq_cosmos_globals_prepare(int dtype, int Q, const void* gparams, const void* mc,
                              const double* gstate, void* gprep, void* stream);
int tq_cosmos_globals_finish(int dtype, int Q, const void* mc, const double* gstate,
                             const void* gprep, const double* acc, double sN, double sF,
                             void* ggrads, double* loss, void* stream);

/* ---------------------------------------------------------------------------------------------
 * hmm variant of cosmos (models/hmm.py): the guide enumerates a Markov chain z_f with AOI-local
 * transition tables `z_trans` and spot presences conditional on z_f; every frame is used (fb == F, fdx == NULL).
 * Call order of one step:
 *   tq_hmm_globals_sample -> tq_hmm_globals_prepare, tq_cosmos_sites (continuous sites, unchanged) -> tq_hmm_forward
 *   (needs the tables of tq_hmm_globals_sample)
 *   -> tq_ksmogn_fwd_bwd (W = qm of tq_hmm_forward) -> tq_hmm_local_post -> tq_hmm_backward
 *   -> [all-reduce of acc and hacc] -> tq_hmm_globals_finish -> tq_adam_dense x2 -> tq_step_advance
 * Local flat buffer (tq_hmm_local_numel values): the cosmos layout (its m_probs slabs hold m_probs[z = 0]), then
 * m_probs[z = 1] (K, Nt, F, C), then z_trans (Nt, F, C, 2, 2) -- all unconstrained.
 * Global flat buffer: the cosmos layout with pi_* read as init_*, then trans_mean (Q, 2, 2), trans_size (Q, 2);
 * global noise: cosmos order, then trans (Q, 2, 2).
 * a: (2, U) double forward marginals; v: (2, U) emission values of the two states; hpartial: (nb * C,
 * tq_hmm_chain_sums()) and hacc: (C, tq_hmm_chain_sums()) double scratch / sums of the chain (ELBO terms, expected
 * initial-state and transition counts). */
int64_t tq_hmm_local_numel(int64_t Nt, int64_t F, int64_t C);
int tq_hmm_chain_sums(void);
int tq_hmm_globals_sample(int dtype, int Q, const void* gparams, const void* mc, const double* noise_in,
                          uint64_t seed, const void* state, double* gstate, void* tables,
                          void* gain_out, void* stream);
int tq_hmm_globals_prepare(int dtype, int Q, const void* gparams, const void* mc,
                           const double* gstate, void* gprep, void* stream);
int tq_hmm_globals_finish(int dtype, int Q, const void* mc, const double* gstate, const void* gprep,
                          const double* acc, const double* hacc, double sN, void* ggrads,
                          double* loss, void* stream);
int tq_hmm_chain_rows(void);   /* rows: (tq_hmm_chain_rows(), U) double per-frame terms of the chain, written by forward */
int tq_hmm_forward(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc,
                   const void* lparams, const void* tables, double* rows, double* a_out, void* qm,
                   void* stream);
int tq_hmm_local_post(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc,
                      const void* lparams, const void* tables, const void* samples, const void* rec,
                      const void* L, const void* gs, const void* g_rate, const double* a_in, double sN,
                      void* lgrads, void* v_out, double* tickets, double* block_partial, double* acc,
                      void* stream);
/* theta_probs (K, nb, F, C) += weight * p(theta = k+1 | z_map, m, x, y) averaged over m with q(m | z_map), for one guide
 * draw (samples of tq_cosmos_sites, tables of tq_hmm_globals_sample); z_map: (nb, F, C) uint8 (hmm.py:541-625). */
int tq_hmm_theta_probs(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc,
                       const void* lparams, const void* tables, const void* samples, const void* z_map,
                       double weight, void* theta_probs, void* stream);
int tq_hmm_backward(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc,
                    const void* lparams, const void* tables, const double* rows, const double* a_in,
                    const void* v_in, double sN, void* lgrads, double* hpartial, double* hacc,
                    void* stream);

/* ---------------------------------------------------------------------------------------------
 * Latency-bound sum of <= tq_p2p_max_values() doubles across the <= tq_p2p_max_ranks() GPUs of one box (the global
 * sites' accumulators), over NVLink peer memory instead of NCCL: tq_p2p_alloc gives a device buffer and
 * its 64-byte CUDA IPC handle; exchange the handles (any host channel), tq_p2p_open the peers', put the `world` buffer
 * pointers (own at index `rank`) in a device array.  Per step, on one stream: tq_p2p_push (writes this rank's values
 * into every peer's buffer, then a flag) ... tq_p2p_wait_sum (polls this rank's own buffer until every peer's flag has
 * arrived, adds the slots in rank order into `out`: bit-identical on all ranks).  Graph-capturable; a wait gives up
 * after seconds instead of hanging (tq_p2p_timed_out). */
int64_t tq_p2p_bytes(void);
int tq_p2p_max_values(void);
int tq_p2p_max_ranks(void);
int tq_p2p_alloc(void** buffer, unsigned char* handle64);
int tq_p2p_open(const unsigned char* handle64, void** buffer);
int tq_p2p_close(void* buffer);
int tq_p2p_free(void* buffer);
int tq_p2p_push(const double* values, int n, int rank, int world, const void* peers, void* stream);
int tq_p2p_wait_sum(void* own_buffer, int n, int world, double* out, void* stream);
int tq_p2p_timed_out(const void* own_buffer, unsigned long long* seq);

/* ---------------------------------------------------------------------------------------------
 * Ingestion of raw .glimpse frames (imscroll/glimpse_reader.py:168-186, 354-381).
 * frames_raw: (Fc, H, W) big-endian int16 exactly as stored in the file (device memory); pixel value =
 * int16 + 2^15.  Frame f0 + i of the movie is chunk frame i.
 *
 * tq_crop_aois: for every AOI n and chunk frame, raw = aoi_xy[n] + drift[f] (x, y order; double),
 * shift = round_half_even(raw - (P-1)/2), patches[n, f] = frame[shifty : shifty+P, shiftx : shiftx+P]
 * (uint16, layout (N, F, P, P)), target_xy[n, f] = raw - shift.  *status |= 1 if a window leaves the frame
 * (that patch is left untouched).
 * tq_offset_hist: counts[v] += #pixels equal to v in frame[oy : oy+oP, ox : ox+oP] over the chunk
 * (counts: 65536 uint64). */
int tq_crop_aois(const void* frames_raw, int H, int W, int Fc, int f0, const double* aoi_xy,
                 const double* drift, int N, int F, int P, void* patches, double* target_xy,
                 int* status, void* stream);
int tq_offset_hist(const void* frames_raw, int H, int W, int Fc, int offset_x, int offset_y,
                   int offset_P, void* counts, void* stream);

/* torch.optim.Adam step (pyro.optim.Adam({"lr", "betas"}), models/model.py:168-171), dense over
 * the whole buffer; *state = number of completed steps. */
int tq_adam_dense(int dtype, int64_t n, void* params, const void* grads, void* exp_avg,
                  void* exp_avg_sq, double lr, double beta1, double beta2, double eps,
                  const void* state, void* stream);

/* *state += 1 */
int tq_step_advance(void* state, void* stream);

/* Device-resident step state: { uint64 step; uint32 pending; float step_size, inv_sqrt_bc2 } (tq_sizeof_step_state
 * bytes, zero-initialised by the caller).  Entry points that only need the step count accept an 8-byte state.
 * tq_step_advance_deferred: end of a step whose local update is deferred -- stores the bias-corrected constants of
 * update number step + 1, sets `pending`, then step += 1.
 * tq_adam_deferred_flush: applies a pending update now over n entries (the tq_local_deferred_range part of the flat
 * buffers) and clears `pending`; a no-op when nothing is pending. */
int tq_sizeof_step_state(void);
int tq_step_advance_deferred(void* state, double lr, double beta1, double beta2, void* stream);
int tq_adam_deferred_flush(int dtype, int64_t n, void* params, const void* grads, void* exp_avg,
                           void* exp_avg_sq, double beta1, double beta2, double eps, void* state,
                           void* stream);

/* Measured-peak helpers for the roofline (register-resident FMA / MUFU loops); *ops receives the
 * operations issued by the launch (host pointer). */
int tq_peak_fma(int blocks, int iters, void* scratch, double* ops, void* stream);
int tq_peak_mufu(int blocks, int iters, void* scratch, double* ops, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Post-fit statistics.  Central credible intervals of the guide distributions: what
 * cosmos.compute_params (models/cosmos.py:711-784) obtains from scipy.stats.gamma / beta `.interval(CI)` through
 * stats.torch_to_scipy_dist (utils/stats.py:262-293), for whole parameter arrays at once: n elements, all arrays
 * DOUBLE device pointers, lo / hi = quantiles (1 -+ ci) / 2 of Gamma(conc, rate) resp. Beta(c1, c0) on [0, 1]
 * (an AffineBeta's interval is low + scale * these; a Dirichlet's marginals are Beta(c_i, sum c - c_i)). */
int tq_gamma_interval(int64_t n, const double* conc, const double* rate, double ci, double* lo, double* hi,
                      void* stream);
int tq_beta_interval(int64_t n, const double* c1, const double* c0, double ci, double* lo, double* hi,
                     void* stream);
/* stats.snr_and_chi2 (utils/stats.py:29-86) for U patches in store order: pixels (U, P, P) of `pixtype`
 * (TQ_PIX_U16 / TQ_PIX_F32), xy (U, 2), height / width / x / y (2, U), background (U) float -> snr (2, U), chi2 (U). */
int tq_snr_chi2(int64_t U, int P, int pixtype, const void* pixels, const void* xy, const void* height,
                const void* width, const void* x, const void* y, const void* background, double gain,
                double offset_mean, double offset_var, void* snr, void* chi2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TAPQIR_B200_H */
