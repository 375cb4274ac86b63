/*
 * mhentropy_b200 — C ABI of the B200-native multi-hypothesis hot path of GloryyrolG/MHEntropy.
 *
 * The reference (pure Python/PyTorch) has no FFI; its boundary for this path is the nn.Module API
 * of RealNVP (hand/flows.py:125-362), ManoLayer (hand/ManoLayer.py:10-165,
 * hand/manopth/manolayer.py:13-274) and the MHEnt loss glue (hand/network.py:455-831).  This
 * header is what a binding for that boundary binds: every entry point names the reference code
 * it replaces.  The Python drop-in (mhentropy_b200/*.py) calls exactly these symbols through
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 (unless noted) owned by the caller; the library
 *     never allocates, never synchronises and keeps no global state besides the last-error text;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it;
 *   - return value 0 = ok, otherwise an mhe_status; mhe_last_error_string() has the detail;
 *   - rows are hypothesis-major: row r = n*B + b  (reference network.py:734,747,793);
 *   - "accumulate" outputs are added to (+=), everything else is overwritten.
 */
#ifndef MHENTROPY_B200_H
#define MHENTROPY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum mhe_status {
    MHE_OK = 0,
    MHE_ERR_INVALID_ARG = 1,
    MHE_ERR_WORKSPACE = 2,
    MHE_ERR_CUDA = 3,
    MHE_ERR_UNSUPPORTED = 4
} mhe_status;

const char* mhe_last_error_string(void);
/* library version, and the compute capability the kernels were built for (100 = sm_100a) */
int mhe_version(void);
int mhe_built_for_sm(void);
/* number of kernel launches enqueued by this process so far (bench.py's gpu_launches) */
long long mhe_kernel_launch_count(void);
/* Timing probe for measurement only: CUDA events are recorded on the launching stream around every
 * GEMM launch whose label contains `tag` (e.g. "flow G1", "dgrad G1", "wgrad W1"), up to max_launches
 * pairs; mhe_probe_read() synchronises on them and returns the summed duration and the launch count.
 * mhe_probe_configure(NULL, 0) switches it off (the default).                                     */
int mhe_probe_configure(const char* tag, int max_launches);
int mhe_probe_reset(void);
int mhe_probe_read(float* total_ms, int* launches);

/* ------------------------------------------------------------------------------------------
 * Conditional RealNVP flow — reference hand/flows.py.
 *
 * Shape: dim D (45), hidden H (512, both hidden layers), cond C (512), layers L (12 = 2*num_steps).
 * Parameters live in ONE flat fp32 buffer ("flat layout"), every segment aligned to 64 floats:
 *   for layer i in [0,L), net n in {0 = s, 1 = t}:   block (i*2+n) at offset (i*2+n)*BLK
 *        W0 [H][D]  (l.0.weight)   b0 [H]   W1 [H][H] (l.1.weight)   b1 [H]   W2 [D][H] (l.2.weight)   b2 [D]
 *   then for idx = (i*2+n)*2+j, j in {0,1}:          Cw[idx] [H][C]  (c.j.weight)
 *   then for the same idx:                           Cb[idx] [H]     (c.j.bias)
 * (nn.Linear layouts, reference flows.py:83-93).  Gradients use the same layout.
 * mhe_flow_param_offset() returns the float offset of a tensor; which = 0..5 for W0,b0,W1,b1,W2,b2
 * and 6,7 / 8,9 for c.0.weight,c.0.bias / c.1.weight,c.1.bias.
 * ------------------------------------------------------------------------------------------ */
typedef struct mhe_flow_shape {
    int dim;
    int hidden;
    int cond;
    int layers;
    /* Largest number of transformed (mask == 0) or conditioning (mask == 1) dims of any coupling layer, computed by the caller from
     * its {0,1} mask (reference flows.py:152-155 builds dim/2 | dim - dim/2 splits; RealNVP also accepts a user mask, :131).
     * 0 = the reference's default alternating half masks.  The cluster-fused kernels exchange at most 24 dims per layer and side:
     * larger splits run on the per-GEMM tensor-core path.  Masks that are not exactly {0,1} are outside every kernel path.      */
    int max_split;
} mhe_flow_shape;

size_t mhe_flow_param_floats(mhe_flow_shape s);
size_t mhe_flow_param_offset(mhe_flow_shape s, int layer, int net, int which);
/* floats per image of the hoisted conditioning projections: L*4*H */
size_t mhe_flow_cp_floats_per_image(mhe_flow_shape s);
/* Two arithmetic paths share every flow entry point:
 *   packed == NULL : exact fp32 on the CUDA cores;
 *   packed != NULL : tcgen05 tensor cores in 3-pass split precision (hi*hi + hi*lo + lo*hi, fp32 accumulate): half planes
 *                    for weights / activations, bfloat16 planes for gradients;
 *                    `packed` holds the weights as split planes, refreshed by mhe_flow_pack_weights()
 *                    whenever the parameters change.  Needs dim <= 64, hidden % 64 == 0, cond % 8 == 0
 *                    (mhe_flow_packed_bytes() returns 0 otherwise).                                   */
size_t mhe_flow_packed_bytes(mhe_flow_shape s);
/* which (bit set): 1 = the half planes the forward GEMMs read, 2 = the bfloat16 planes the backward GEMMs read; 4 / 8 = only the
 * conditioning / only the coupling part of the half planes (so the conditioning GEMM can start while the rest is converted);
 * 16 / 32 = only the coupling / only the conditioning part of the bfloat16 planes */
int mhe_flow_pack_weights(mhe_flow_shape s, const float* params, void* packed, int which, void* stream);
/* bytes of scratch needed by the flow passes over R rows (forward and backward) on the chosen path */
size_t mhe_flow_workspace_bytes(mhe_flow_shape s, int R, int tensor_core);
/* bytes of the saved-for-backward block a forward pass over R rows fills: the layer inputs on the fp32 path (hidden
 * activations are recomputed), layer inputs + head outputs + hidden activations on the tensor-core path          */
size_t mhe_flow_saved_bytes(mhe_flow_shape s, int R, int tensor_core);
/* bytes of scratch the tensor-core conditioning GEMMs need for B images (0 when that path is unsupported) */
size_t mhe_flow_cond_workspace_bytes(mhe_flow_shape s, int B);

/* Hoisted conditioning projections (reference flows.py:107-109 "Can be pre-processed"):
 *   cp[b][idx][h] = sum_c feat[b][c]*Cw[idx][h][c] + Cb[idx][h] + b_j[h]      (b_j = l.j.bias folded in)
 * feat [B][C] -> cp [B][L*4][H].                                                               */
/* 1 when mhe_flow_cond_fwd on the tensor-core path reads the conditioning half planes of `packed` for B images (so
 * mhe_flow_pack_weights(which & 4) must have refreshed them); 0 when it streams the fp32 weights directly (B <= 128). */
int mhe_flow_cond_fwd_uses_planes(mhe_flow_shape s, int B);
int mhe_flow_cond_fwd(mhe_flow_shape s, const float* params, const void* packed, const float* feat, int B, float* cp,
                      void* workspace, size_t workspace_bytes, void* stream);
/* dcp [B][L*4][H] -> dparams (accumulate: Cw, Cb, b0, b1 slots), dfeat [B][C] (overwritten; may be NULL) */
int mhe_flow_cond_bwd(mhe_flow_shape s, const float* params, const void* packed, const float* feat, const float* dcp, int B,
                      float* dparams, float* dfeat, void* workspace, size_t workspace_bytes, void* stream);

/* One pass through the L coupling layers.
 *   direction 0: z -> x, layers 0..L-1,  x' = m x + (1-m)(x e^s + t),  logdet += sum s   (flows.py:210-217)
 *   direction 1: x -> z, layers L-1..0,  z' = (1-m)(z - t) e^-s + m z,  logdet -= sum s   (flows.py:219-227)
 * in [R][D] -> out [R][D], logdet [R] (may be NULL).  Row r uses cp of image r % B.
 * saved: NULL, or a block of mhe_flow_saved_bytes() receiving what the backward needs.            */
int mhe_flow_pass_fwd(mhe_flow_shape s, const float* params, const void* packed, const float* mask, const float* cp,
                      const float* in, int R, int B, int direction,
                      float* out, float* logdet, float* saved,
                      void* workspace, size_t workspace_bytes, void* stream);

/* The same pass when EVERY row has its own conditioning vector - feat [R][C], no hoisting possible: RealNVP.log_prob(x, logvar=feat) of the
 * reference's scoring / p_nf use (flows.py:271-331 with cond (R, C); hand/CrossModalHand.py:278-279, 293-295) - and nothing is saved for
 * a backward.  The projections c.j(feat) of flows.py:107-109 are not materialised (R x L*4*H fp32, 1.6 GB at 16,384 rows, written and read
 * back) but contracted inside the coupling GEMMs: h_j = lrelu([a | feat] [W_j | Cw_j]^T + b_j + Cb_j).  Tensor-core path only (`packed`
 * from mhe_flow_pack_weights); mhe_flow_rowcond_supported() says whether this entry applies and pays (C a multiple of 64; R >= 1536 by
 * measurement, MHE_ROWCOND_MIN_ROWS) - otherwise use mhe_flow_cond_fwd(B = R) + mhe_flow_pass_fwd.                                   */
int mhe_flow_rowcond_supported(mhe_flow_shape s, int R);
size_t mhe_flow_rowcond_workspace_bytes(mhe_flow_shape s, int R);
int mhe_flow_pass_fwd_rowcond(mhe_flow_shape s, const float* params, const void* packed, const float* mask, const float* feat,
                              const float* in, int R, int direction, float* out, float* logdet,
                              void* workspace, size_t workspace_bytes, void* stream);
/* dout [R][D], dlogdet [R] (NULL = 0; multiplied by dlogdet_scale, so -1 turns dL/dlog_q of the fused
 * sampler into dL/dlogdet) -> din [R][D]; dparams (accumulate; W0,W1,W2,b2 slots), dcp [B][L*4][H] (accumulate). */
int mhe_flow_pass_bwd(mhe_flow_shape s, const float* params, const void* packed, const float* mask, const float* cp,
                      const float* saved, int R, int B, int direction,
                      const float* dout, const float* dlogdet, float dlogdet_scale,
                      float* din, float* dparams, float* dcp,
                      void* workspace, size_t workspace_bytes, void* stream);

/* mhe_flow_pass_bwd followed by mhe_flow_cond_bwd as ONE call (same results).  On the fused tensor-core path the conditioning backward
 * is pipelined into the pass: the layers' dcp sums, conditioning weight / bias gradients and their share of dfeat are enqueued chunk by
 * chunk behind the data-gradient kernel of the chunk, on internal streams, so most of it overlaps the remaining layers instead of
 * following the last one.  With mhe_flow_set_async bit 0 it returns with that work pending (mhe_flow_join).  dfeat may be NULL.        */
int mhe_flow_pass_cond_bwd(mhe_flow_shape s, const float* params, const void* packed, const float* mask, const float* cp,
                           const float* saved, int R, int B, int direction,
                           const float* dout, const float* dlogdet, float dlogdet_scale,
                           float* din, float* dparams, float* dcp, const float* feat, float* dfeat,
                           void* workspace, size_t workspace_bytes, void* cond_workspace, size_t cond_workspace_bytes, void* stream);

/* The conditioning weight gradient alone, from its two factors: dparams Cw slots (+)= dcp[:, idx, :]^T feat over Bt rows (tensor-core
 * path; workspace of mhe_flow_cond_workspace_bytes(s, Bt)).  It is the same contraction mhe_flow_cond_bwd runs with the local batch.
 * Data parallelism: this gradient has rank <= (images) per weight matrix, so the ranks exchange the FACTORS (all-gather of dcp
 * [B][L*4*H] and feat [B][C]: 6.4 MB per rank at 64 images) and each computes the global gradient with Bt = world x B, instead of
 * all-reducing the dense 50 MB; the remaining 30 MB of the flat gradient are all-reduced as before.                               */
int mhe_flow_cond_wgrad(mhe_flow_shape s, const float* feat, const float* dcp, int Bt, float* dparams, void* workspace,
                        size_t workspace_bytes, void* stream);

/* Bucketed gradient exchange.  On the fused tensor-core path the backward pass runs as mhe_flow_bwd_chunk_count() chunks of consecutive
 * layers (1 on the other paths), chunk 0 first (it holds the layers the backward reaches first); mhe_flow_bwd_chunk_layers() gives
 * chunk c's layers.  After an asynchronous mhe_flow_pass_cond_bwd (mhe_flow_set_async bit 0), mhe_flow_join_chunk(stream, c) makes
 * `stream` wait until EVERY gradient of those layers (coupling and conditioning weights and biases) is complete, so that the caller can
 * all-reduce that part of dparams while the remaining chunks still run.  mhe_flow_join() still ends the pass.                         */
/* 1 when a tensor-core pass over R rows runs on the cluster-fused kernels (production shape, R <= MHE_FUSED_MAX_ROWS, splits <= 24 dims),
 * 0 when it runs GEMM by GEMM.  Matters to callers that use mhe_flow_set_async bit 1: only the fused path STORES its weight gradients;
 * the per-GEMM path always accumulates (its long contractions are split over CTAs), so its dparams must be zeroed by the caller.        */
int mhe_flow_pass_is_fused(mhe_flow_shape s, int R);
int mhe_flow_bwd_chunk_count(mhe_flow_shape s, int R);
int mhe_flow_bwd_chunk_layers(mhe_flow_shape s, int R, int direction, int chunk, int* first_layer, int* layers);
int mhe_flow_join_chunk(void* stream, int chunk);

/* Optional head start for mhe_flow_pass_bwd on the fused tensor-core path: everything its weight-gradient GEMMs need that depends only
 * on the forward pass (re-planed saved activations, masked inputs, zeroed scratch), enqueued on `stream`.  May run on any stream once
 * the forward pass that filled `saved` has completed - e.g. while the loss is computed; the caller orders it before
 * mhe_flow_pass_bwd and sets mhe_flow_set_async bit 3.  Returns MHE_ERR_UNSUPPORTED (and does nothing) when the pass would not take
 * the fused path.                                                                                                                   */
int mhe_flow_pass_bwd_prepare(mhe_flow_shape s, const float* mask, const float* saved, int R, int direction, void* workspace,
                              size_t workspace_bytes, void* stream);
/* Zero the slots of a gradient buffer that accumulate even with mhe_flow_set_async bit 1 (the biases); the weight slots are then
 * stored by the backward entry points and need no zeroing.                                                                         */
int mhe_flow_zero_bias_grads(mhe_flow_shape s, float* dparams, void* stream);

/* Gradient-accumulation options (bit set; default 0).
 *   bit 0  asynchronous weight gradients: mhe_flow_pass_bwd may return while its weight-gradient GEMMs (the dparams W0/W1/W2
 *          slots) still run on internal streams, so that the caller can enqueue independent work (the conditioning backward only
 *          needs dcp); mhe_flow_join(stream) makes `stream` wait for them and must be called before dparams is read or the
 *          captured graph ends.
 *   bit 1  the caller promises that the WEIGHT slots of dparams (W0, W1, W2, Cw) are zero when mhe_flow_pass_bwd /
 *          mhe_flow_cond_bwd run: their epilogues then store instead of read-modify-write (bias slots always accumulate).
 *   bit 4  mhe_flow_cond_bwd skips the conditioning WEIGHT gradient (Cw slots): the caller computes it with mhe_flow_cond_wgrad, e.g.
 *          from factors gathered over the data-parallel ranks (tensor-core path only).
 *   bit 3  the caller has run mhe_flow_pass_bwd_prepare on the same saved block / workspace and ordered it before mhe_flow_pass_bwd.
 *   bit 2  the caller promises that dfeat is zero when mhe_flow_cond_bwd runs on the tensor-core path (it is accumulated into with
 *          atomics there): the memset that would otherwise sit on the critical path is skipped.                                 */
int mhe_flow_set_async(int on);
int mhe_flow_join(void* stream);

/* Optimizer step on the flat buffer (reference hand/CrossModalHand.py:201 torch.optim.Adam with default betas / eps, :462-470
 * clip_grad_norm_ then optimizer.step()).  exp_avg / exp_avg_sq: Adam state, same layout as params.  step >= 1 is the step count AFTER
 * this update.  grad_scale multiplies the gradient first: the reference clips by the global norm over the whole encoder (CNN backbone
 * included), so the caller combines mhe_flow_grad_sqnorm() with its other modules' terms and passes min(1, max_norm / norm).  When
 * packed != NULL the split planes mhe_flow_pack_weights() would produce (half and bfloat16, every family) are refreshed in the same
 * call - the two dense families straight from the update's registers -, so the next step needs no re-pack.                         */
int mhe_flow_grad_sqnorm(mhe_flow_shape s, const float* dparams, double* sqnorm_out /* device, 1 double */, void* stream);
int mhe_flow_adam_step(mhe_flow_shape s, float* params, const float* dparams, float* exp_avg, float* exp_avg_sq, void* packed, int step,
                       double lr, double beta1, double beta2, double eps, double grad_scale, void* stream);

/* Reduction step of the peer-memory gradient exchange (mhentropy_b200/parallel.py PeerExchange; the reference has no distributed code,
 * SURVEY.md section 8e): acc[i] += sum over s in [0, n_src), s != skip, of land[s * stride + i] for i < n.  `land` holds the copies of this
 * rank's shard that the other ranks pushed over NVLink (slot s = source rank), `skip` is this rank's own (unused) slot.              */
int mhe_sum_shards(float* acc, const float* land, int n_src, int skip, size_t stride, size_t n, void* stream);

/* log N(z; 0, I) + logdet per row (flows.py:320) and its gradient seeds:
 *   fwd: logp[r] = -0.5|z_r|^2 - 0.5 D ln(2 pi) + logdet_sign*logdet[r]   (logdet may be NULL)
 *   bwd: dz[r][:] = -z[r][:] * dlogp[r]                                                          */
int mhe_std_normal_logp_fwd(const float* z, const float* logdet, float logdet_sign, int R, int D, float* logp, void* stream);
int mhe_std_normal_logp_bwd(const float* z, const float* dlogp, int R, int D, float* dz, void* stream);

/* ------------------------------------------------------------------------------------------
 * MANO layer — reference hand/manopth/manolayer.py:110-274 (use_pca, ncomps=45, axis-angle root,
 * right hand, center_idx=9) and the wrapper's second joint set hand/ManoLayer.py:141-148.
 * Constants (device, fp32), packed by the caller once:
 *   comps [45][45] (th_selected_comps), hands_mean [45], v_template [778][3], shapedirs [778][3][10],
 *   posedirs_t [135][2334] (th_posedirs transposed: k-major), jreg [16][778], weights [778][16],
 *   jt [16][3] = jreg*v_template, js [16][3][10] = jreg*shapedirs.
 * ------------------------------------------------------------------------------------------ */
typedef struct mhe_mano_consts {
    const float* comps;
    const float* hands_mean;
    const float* v_template;
    const float* shapedirs;
    const float* posedirs_t;
    const float* jreg;
    const float* weights;
    const float* jt;
    const float* js;
    const float* pose_tables;   /* optional (may be NULL): mhe_mano_pose_tables_floats() floats filled by mhe_mano_pack_pose_tables() */
    const void* posedirs_planes; /* optional (may be NULL): mhe_mano_posedirs_planes_bytes() bytes filled by mhe_mano_pack_posedirs_planes() */
} mhe_mano_consts;

/* The pose / joints kernels stage the small tables they need (PCA basis, joint regressors, the blend-shape rows and skinning
 * weights of the five tip vertices) in shared memory.  pose_tables is that set gathered once into one contiguous array, so the
 * staging is a coalesced copy; with pose_tables == NULL every block gathers it from the full-size constants instead.          */
size_t mhe_mano_pose_tables_floats(void);
int mhe_mano_pack_pose_tables(const mhe_mano_consts* c, float* pose_tables, void* stream);

/* Pose blend shapes on the tensor cores (reference manolayer.py:186-190: th_v_posed = th_v_shaped + posedirs . pose_map): with
 * posedirs_planes set - posedirs as split half planes, the B operand of a tcgen05 GEMM - the mesh forward computes the pose offsets of
 * all rows as ONE [R x 135] . [135 x 2334] contraction (3-pass split precision, fp32 accumulate) instead of a 135-term dot product per
 * vertex on the CUDA cores; with NULL it takes the latter path.                                                                    */
size_t mhe_mano_posedirs_planes_bytes(void);
int mhe_mano_pack_posedirs_planes(const mhe_mano_consts* c, void* posedirs_planes, void* stream);

#define MHE_MANO_VERTS 778
#define MHE_MANO_JOINTS 21
/* joint_order: 0 = manopth order (manolayer.py:260); 1 = RHD order (ManoLayer.py:54-56, utils.py:15).
 * mesh_grad != 0 sizes the scratch for a backward that receives dverts / djoints2.                 */
size_t mhe_mano_workspace_bytes(int R, int mesh_grad);

/* theta [R][ld_theta>=48], beta [R][ld_beta>=10] -> verts [R][778][3] mm (NULL = skip the mesh, only the
 * 5 tip vertices are skinned), jtr [R][21][3] mm, joints2 [R][21][3] mm (wrapper's regressed set; NULL = skip;
 * needs verts).  Nothing is saved for the backward: it recomputes from theta / beta.               */
int mhe_mano_fwd(const mhe_mano_consts* c, const float* theta, int ld_theta, const float* beta, int ld_beta,
                 int R, int joint_order, float* verts, float* jtr, float* joints2,
                 void* workspace, size_t workspace_bytes, void* stream);
/* dverts / djtr / djoints2 (any may be NULL = 0) -> dtheta [R][ld_dtheta] (48 written), dbeta [R][ld_dbeta]
 * (10 written).  accumulate != 0 adds into dtheta/dbeta instead of overwriting.                  */
int mhe_mano_bwd(const mhe_mano_consts* c, const float* theta, int ld_theta, const float* beta, int ld_beta,
                 int R, int joint_order,
                 const float* dverts, const float* djtr, const float* djoints2,
                 float* dtheta, int ld_dtheta, float* dbeta, int ld_dbeta, int accumulate,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Reprojection + loss reductions — reference hand/network.py:466-514 (root 12 / bone 11 normalise,
 * orthographic projection), :233-258 (Laplace, b const), :155-165 (box / ball priors),
 * :612-667, :760-831 (N-means, entropy) and criteria.py:55,173 (mean_B of -log_p).
 * z [R][61] = th3 | th45 | bt | logs | t (network.py:703-717); joints in RHD order.
 * ------------------------------------------------------------------------------------------ */
typedef struct mhe_loss_cfg {
    float laplace_b;      /* 0.03  ho3d.yaml:44 */
    float th45_box;       /* 2     network.py:427 */
    float th45_alpha;     /* 50    ho3d.yaml:41 */
    float th3_radius;     /* pi    network.py:431 */
    float th3_alpha;      /* 5 */
    float bt_box;         /* 0.03  network.py:433 */
    float bt_alpha;       /* 50 */
    int root_idx;         /* 12    network.py:478 */
    int norm_idx;         /* 11    network.py:479 */
} mhe_loss_cfg;

/* z = combine(x_flow [R][45], z_det [B][16] = th3|bt|logs|t)  (network.py:703-717, 747) */
int mhe_combine_z_fwd(const float* x_flow, const float* z_det, int R, int B, float* z, void* stream);
/* dz [R][61] -> dx_flow [R][45] (overwritten; may be NULL), dz_det [B][16] (overwritten: sum over the hypotheses) */
int mhe_combine_z_bwd(const float* dz, int R, int B, float* dx_flow, float* dz_det, void* stream);

/* joints [R][21][3] (RHD order, mm), z [R][61], crop_uv [B][42], vis [B][21], log_q [R] ->
 *   uv [R][42] (may be NULL), row_log_p [R] (Laplace + priors, network.py:662), and per image
 *   log_p [B] = mean_n(-log_q) + mean_n(row_log_p), h [B], q_log_p [B] (any may be NULL),
 *   loss [1] = mean_b(-log_p) (may be NULL).                                                    */
int mhe_reproj_loss_fwd(const mhe_loss_cfg* cfg, const float* joints, const float* z, const float* crop_uv,
                        const float* vis, const float* log_q, int R, int B,
                        float* uv, float* row_log_p, float* log_p, float* h, float* q_log_p, float* loss,
                        void* stream);
/* dlog_p [B] (NULL: use dloss), dloss [1] device scalar (NULL = 1.0) ->
 *   djoints [R][21][3], dz [R][61] (priors + camera terms; overwritten), dlog_q [R].               */
int mhe_reproj_loss_bwd(const mhe_loss_cfg* cfg, const float* joints, const float* z, const float* crop_uv,
                        const float* vis, int R, int B, const float* dlog_p, const float* dloss,
                        float* djoints, float* dz, float* dlog_q, void* stream);

/* The image-level reductions of mhe_reproj_loss_fwd alone (network.py:793-808, criteria.py:55,173): row_log_p [R], log_q [R] ->
 *   log_p [B], h [B], q_log_p [B] (any may be NULL), loss [1] (may be NULL; needs log_p).                                        */
int mhe_image_loss_reduce(const float* row_log_p, const float* log_q, int R, int B, float* log_p, float* h, float* q_log_p,
                          float* loss, void* stream);

/* The joints-only training path of every hypothesis in ONE launch: MANO forward (mhe_mano_fwd without the mesh; reference
 * manolayer.py:110-274), root / bone normalisation, projection, Laplace(visible) and priors (mhe_reproj_loss_fwd's row part;
 * utils.py:46-66, network.py:497-514, 233-258, 155-165), and the backward of both down to dz (mhe_reproj_loss_bwd + mhe_mano_bwd).
 * The loss is linear in the row terms, so with the criterion's loss = -mean_b log_p (criteria.py:55,173) the gradient seed of every
 * row is the constant -dloss / R and nothing waits for a reduction; dlog_p [B] != NULL supplies dL/dlog_p per image instead (the seed
 * of a row of image b is then dlog_p[b] / N: what an autograd caller with an arbitrary downstream loss needs).  z [R][61] (theta = z[:, 0:48], beta = z[:, 48:58]) - or z == NULL and its two sources x_flow [R][45], z_det [B][16]
 * (mhe_combine_z_fwd's inputs: the kernel assembles the row itself) -, crop_uv [B][42], vis [B][21] ->
 *   jtr [R][21][3] (may be NULL), uv [R][42] (may be NULL), row_log_p [R], dz [R][61] (overwritten), dx_flow [R][45] (may be NULL: the
 *   flow's columns of dz, what mhe_combine_z_bwd extracts), dlog_q [R] (may be NULL).                                               */
int mhe_hypothesis_rows_fwd_bwd(const mhe_mano_consts* c, const mhe_loss_cfg* cfg, const float* z, const float* x_flow,
                                const float* z_det, const float* crop_uv, const float* vis, int R, int B, int joint_order, float dloss,
                                const float* dlog_p,
                                float* jtr, float* uv, float* row_log_p, float* dz, float* dx_flow, float* dlog_q, void* stream);

/* xyz / verts normalisation and projection for MHEnt.sample (network.py:466-483, 497-514, 876-877):
 * joints [R][21][3], verts [R][778][3] (NULL ok), logs_t = z[:, 58:61] with row stride ld_z ->
 * xyz [R][21][3], verts_n [R][778][3], uv [R][21][2] (pixels when inv_norm != 0).                 */
int mhe_normalize_project(const mhe_loss_cfg* cfg, const float* joints, const float* verts, const float* z,
                          int ld_z, int R, int inv_norm, int image_size,
                          float* xyz, float* verts_n, float* uv, void* stream);

/* ------------------------------------------------------------------------------------------
 * Hypothesis selection and multi-hypothesis evaluation metrics.
 *
 * mhe_topk_hypotheses: reference hand/network.py:866-871 (torch.topk of log q over the hypothesis axis).
 *   log_q [N][B] -> idx [k][B] (int64): the k most likely hypotheses of every image, most likely first.
 * mhe_hypothesis_metrics: reference hand/criteria.py:91-168 (MHEntLoss metrics with aligned = False) and hand/utils.py:21-30.
 *   xyz [N][B][21][3] (normalised joints), uv [N][B][21][2] (pixels); targets pose3d [B][21][3], scale [B], crop_uv [B][42] in
 *   [-1, 1], vis [B][21] -> metrics [14][B], rows in the reference's key order:
 *     eucLoss_3d_rgb_{sample, sample_std, vis, vis_std, vis_mean, invis, invis_std}, then the same seven for 2d:
 *   best-hypothesis mean joint error per group (worst hypothesis for 2d / vis), spread of the hypotheses, mean over hypotheses.
 * ------------------------------------------------------------------------------------------ */
size_t mhe_hypothesis_metrics_workspace_bytes(int B);
int mhe_hypothesis_metrics(const float* xyz, const float* uv, const float* pose3d, const float* scale, const float* crop_uv,
                           const float* vis, int N, int B, int root_idx, float image_size, float* metrics,
                           void* workspace, size_t workspace_bytes, void* stream);
int mhe_topk_hypotheses(const float* log_q, int N, int B, int k, long long* idx, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tensor-core building blocks (tcgen05 + TMA), exposed for tests and tools.
 * fp32 values travel as split 16-bit planes x = hi + lo, [batches][planes][rows_p][cols_p]: IEEE half planes (f16 = 1,
 * ~22 significant bits, for range-safe data: weights, activations) or bfloat16 planes (f16 = 0, ~16 bits, full fp32
 * range: gradients).  The flow uses half planes for every forward operand and bfloat16 planes for gradients.
 * ------------------------------------------------------------------------------------------ */
/* src fp32 [batches][rows][cols] dense -> dst planes, zero padded to rows_p x cols_p (cols_p % 8 == 0). */
int mhe_split_planes(const float* src, int rows, int cols, void* dst, int rows_p, int cols_p, int planes, int batches, int f16, void* stream);
/* C [batches][M][N] fp32 = A * B from dense plane tensors.  a_mn / b_mn = 0: A is [M][K], B is [N][K] (K-major);
 * = 1: A is [K][M], B is [K][N] (MN-major).  planes = 1: one 16-bit pass; 2: three passes (hi*hi + hi*lo + lo*hi).  f16 selects half or bfloat16 planes.
 * bn in {64, 128} is the N tile; ksplit > 1 splits K across CTAs with atomic accumulation.          */
int mhe_tc_gemm_raw(const void* A, const void* B, float* C, int M, int N, int K, int batches, int planes, int a_mn, int b_mn, int bn,
                    int ksplit, int f16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MHENTROPY_B200_H */
