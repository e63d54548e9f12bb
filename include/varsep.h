/* varsep.h — C ABI of libvarsep_sm100a.so (B200 / sm_100a only).
 *
 * The reference (JeremieDona/spatiotemporal_variable_separation) has no native
 * code and no FFI: its hot path bottoms out in PyTorch ATen operators.  Each
 * entry point below replaces the ATen/cuDNN/cuBLAS work issued by one reference
 * call site; the citation after "replaces:" is that call site (paths relative
 * to /root/reference/var_sep/).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the name ends in _host;
 *   - nothing allocates, synchronises or touches the default stream: work is
 *     enqueued on the cudaStream_t passed as `stream`;
 *   - return 0 on success, non-zero on error; vs_last_error() gives the text
 *     (thread-local);
 *   - activations are NHWC ("pixel rows x channels"), element type `dtype`
 *     (VS_F32 or VS_BF16); accumulation is always fp32, BatchNorm statistics
 *     are accumulated in fp64;
 *   - "groups": the leading (sample) dimension is G equal consecutive blocks,
 *     one per *reference call* that was batched together; BatchNorm batch
 *     statistics are computed per (group, channel) so that one launch over G
 *     decoder time steps equals G sequential reference calls.
 */
#ifndef VARSEP_H_
#define VARSEP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { VS_F32 = 0, VS_BF16 = 1 };
enum { VS_ACT_NONE = 0, VS_ACT_RELU = 1, VS_ACT_LEAKY = 2, VS_ACT_ELU = 3, VS_ACT_SIGMOID = 4, VS_ACT_TANH = 5 };
enum { VS_CONV_DIRECT = 0, VS_CONV_TRANSPOSED = 1 };
enum { VS_FLAG_FORCE_SIMT = 1 };   /* bypass the tcgen05 path (debug / A-B parity) */

/* Geometry of a convolution, always stated for the DIRECT map
 *   big[N,H,W,C]  --(R x S, stride, pad)-->  small[N,P,Q,K],  P = (H + 2*pad - R)/stride + 1.
 * nn.Conv2d          : fprop = DIRECT (big->small), dgrad = TRANSPOSED.
 * nn.ConvTranspose2d : fprop = TRANSPOSED (small->big), dgrad = DIRECT.
 * nn.Linear          : H=W=P=Q=R=S=1.
 * The torch weight of both module kinds is w[K][C][R][S] in these terms. */
typedef struct {
    int32_t dtype;
    int32_t N, H, W, C;
    int32_t P, Q, K;
    int32_t R, S, stride, pad;
    int32_t groups;     /* BatchNorm statistic groups along N (>=1) */
    int32_t act;        /* activation fused in the epilogue (VS_ACT_*) */
    int32_t flags;
} vs_conv_geom;

int         vs_abi_version(void);
const char* vs_last_error(void);
/* number of kernels this library has launched in the calling process (for bench.py "gpu_launches") */
int64_t     vs_launch_count(void);

/* ---- weights ------------------------------------------------------------------------------
 * replaces: the implicit weight layout handling inside cuDNN/oneDNN for nn.Conv2d /
 * nn.ConvTranspose2d / nn.Linear (networks/conv.py:119-123,258-263; networks/mlp.py:40).
 * w is the fp32 torch weight [K][C][R][S];  swap=0 -> out[K][R*S][C] (DIRECT operand),
 * swap=1 -> out[C][R*S][K] (TRANSPOSED operand); out has element type `dtype`. */
int vs_pack_weight(const float* w, void* out, int32_t dtype, int32_t K, int32_t C, int32_t RS, int32_t swap,
                   void* stream);

/* All packed copies of a model in one launch (after the optimizer step).  table: DEVICE array of n rows
 * { const float* src; void* dst; int32 K, C, RS, swap, dtype, first_block; } (40 bytes, 8-byte aligned), rows ordered
 * by first_block; row i owns blocks [first_block_i, first_block_{i+1}) of VS_PACK_BLOCK_ELEMS elements each (four
 * 1024-element transpose tiles: their loads are issued together). */
#define VS_PACK_BLOCK_ELEMS 4096
int vs_pack_weights_multi(const void* table, int32_t n, int32_t total_blocks, void* stream);

/* ---- convolution / linear -------------------------------------------------------------------
 * replaces: aten::convolution (Conv2d / ConvTranspose2d forward; conv.py:119-123,147-170,258-263,
 * 295-318,327-343,363-417,435,517-533; resnet.py:57-63) and aten::addmm (mlp.py:40, conv.py:124).
 * mode = VS_CONV_DIRECT    : in=[N,H,W,C] -> out=[N,P,Q,K], wp=[K][R*S][C], bias[K]
 * mode = VS_CONV_TRANSPOSED: in=[N,P,Q,K] -> out=[N,H,W,C], wp=[C][R*S][K], bias[C]
 * out = act(conv + bias).  If stats != NULL (double[groups][OC][2], zeroed by the caller) the
 * per-(group, channel) sum and sum of squares of (conv + bias) are accumulated into it
 * (BatchNorm batch statistics fused into the producer; act must then be VS_ACT_NONE). */
int vs_conv_forward(const vs_conv_geom* g, int32_t mode, const void* in, const void* wp, const float* bias,
                    void* out, double* stats, void* stream);

/* Which kernel family vs_conv_forward would use for this geometry: 0 = CUDA-core gather GEMM, 1 = tcgen05/TMA tap
 * GEMM, 2 = thin streaming kernel, 3 = tcgen05 GEMM over an im2col tile the CTA builds itself (few image-side
 * channels), 4 = tcgen05 GEMM + col2im (last decoder up-convolution), -1 = invalid geometry.  Host-only (no launch); used by bench.py to attribute
 * the measured launch times to the tensor-core kernel for the roofline. */
int vs_conv_forward_path(const vs_conv_geom* g, int32_t mode);

/* For a geometry on path 1: which tap-GEMM kernel is launched (default switches, 16-byte aligned tensors): 1 = per-class
 * single-CTA kernel, 2 = CTA-pair kernel (cta_group::2), 3 = shifted-window CTA-pair kernel with resident weights; 0 on
 * every other path.  A function of the layer geometry and dtype only — never of the batch size g->N (a sample's result
 * must not depend on what else is in the launch).  Host-only (no launch). */
int vs_conv_forward_variant(const vs_conv_geom* g, int32_t mode);

/* replaces: aten::convolution_backward, weight part.  dw[K][C][R][S] (fp32, torch layout) +=
 * sum over pixels of small[n,p,q,k] * big[n, p*stride-pad+r, q*stride-pad+s, c].
 * For nn.Conv2d small = dy, big = x; for nn.ConvTranspose2d small = x, big = dy. */
int vs_conv_wgrad(const vs_conv_geom* g, const void* small, const void* big, float* dw, void* stream);

/* replaces: the bias part of aten::convolution_backward / the aten::sum of addmm's backward.
 * db[C] += column sums of a [rows, C] matrix. */
int vs_colsum(const void* a, int32_t dtype, int64_t rows, int32_t C, float* db, void* stream);

/* ---- BatchNorm (training, grouped) + activation ---------------------------------------------
 * replaces: aten::native_batch_norm (+ the running-stat EMA and the num_batches_tracked add_)
 * and the in-place activation that follows it (conv.py:56-59; nn.BatchNorm2d eps 1e-5, momentum 0.1).
 * stats: double[G][C][2] from vs_conv_forward; count = elements per (group, channel).
 * Writes mean/invstd [G][C]; applies G sequential EMA updates (unbiased variance) to running_*. */
int vs_bn_finalize(const double* stats, int32_t G, int32_t C, int64_t count, float eps, float momentum,
                   float* mean, float* invstd, float* running_mean, float* running_var,
                   int64_t* num_batches_tracked, void* stream);
/* vs_bn_finalize followed by vs_bn_act_forward in ONE launch where the column kernel applies (identical results: every
 * CTA derives mean / invstd of its channels from the fp64 sums; one CTA per group publishes them, one applies the EMA) */
int vs_bn_finalize_act_forward(const double* stats, int32_t G, int32_t C, int64_t count, float eps, float momentum,
                               float* mean, float* invstd, float* running_mean, float* running_var,
                               int64_t* num_batches_tracked, const void* y, void* out, int32_t dtype, int64_t rows,
                               const float* gamma, const float* beta, int32_t act, void* stream);
/* eval mode: mean = running_mean, invstd = rsqrt(running_var + eps) */
int vs_bn_eval_stats(const float* running_mean, const float* running_var, int32_t C, float eps, float* mean,
                     float* invstd, void* stream);
/* out = act(gamma * (y - mean[g]) * invstd[g] + beta),  y/out [rows, C], rows = G * rows_per_group */
int vs_bn_act_forward(const void* y, void* out, int32_t dtype, int64_t rows, int32_t C, int32_t G,
                      const float* mean, const float* invstd, const float* gamma, const float* beta, int32_t act,
                      void* stream);
/* replaces: aten::native_batch_norm_backward + the activation backward.
 * pass 1: sums[g][c] = { sum dz, sum dz * xhat } with dz = dout * act'(.)  (double[G][C][2], zeroed by caller) */
int vs_bn_act_backward_reduce(const void* dout, const void* y, int32_t dtype, int64_t rows, int32_t C, int32_t G,
                              const float* mean, const float* invstd, const float* gamma, const float* beta,
                              int32_t act, double* sums, void* stream);
/* pass 2: dy = gamma*invstd*(dz - s1/count - xhat*s2/count)  (train)   or gamma*invstd*dz (eval, train=0);
 * dgamma[c] += sum_g s2, dbeta[c] += sum_g s1 (may be NULL) */
int vs_bn_act_backward_apply(const void* dout, const void* y, void* dy, int32_t dtype, int64_t rows, int32_t C,
                             int32_t G, const float* mean, const float* invstd, const float* gamma,
                             const float* beta, int32_t act, const double* sums, int32_t train, float* dgamma,
                             float* dbeta, void* stream);

/* ---- element-wise --------------------------------------------------------------------------
 * replaces: aten::leaky_relu_/relu_/sigmoid/... backward (networks/utils.py:50-72). dx = dout * act'(out) */
int vs_act_backward(const void* dout, const void* out, void* dx, int32_t dtype, int64_t n, int32_t act,
                    void* stream);
/* replaces: aten::add (resnet.py:29,69; conv.py:465) [+ relu_ of conv.py:466]: out = act(a + b) */
int vs_add_act(const void* a, const void* b, void* out, int32_t dtype, int64_t n, int32_t act, void* stream);
/* replaces: aten::cat on channels with the S-code / skip broadcast over groups (conv.py:221,228,388-394;
 * model.py:74,82).  dst[r][dst_off + c] = src[r % src_rows_mod][c] for c < src_C; rows = dst rows.
 * src_rows_mod = rows of src: pixel rows repeat every src_rows_mod (broadcast of one call's S over G groups). */
int vs_copy_channels(const void* src, int32_t src_C, int64_t src_rows_mod, void* dst, int32_t dst_C,
                     int32_t dst_off, int64_t rows, int32_t dtype, void* stream);
/* backward of the above: dsrc[r][c] (+)= sum over repeats of ddst[r + j*src_rows][dst_off + c]; fp32 accumulate */
int vs_slice_channels_reduce(const void* ddst, int32_t dst_C, int32_t dst_off, int64_t rows, void* dsrc,
                             int32_t src_C, int64_t src_rows, int32_t dtype, void* stream);
/* replaces: aten::mul mixing (conv.py:223, mlp_encdec.py:47): out[r][c] = s[r % s_rows][c] * t[r][c] */
int vs_mul_bcast(const void* s, int64_t s_rows, const void* t, void* out, int64_t rows, int32_t C, int32_t dtype,
                 void* stream);
/* ds[r][c] = sum_j dout[r + j*s_rows][c] * t[r + j*s_rows][c];  dt[r][c] = dout[r][c] * s[r % s_rows][c] */
int vs_mul_bcast_backward(const void* dout, const void* s, int64_t s_rows, const void* t, void* ds, void* dt,
                          int64_t rows, int32_t C, int32_t dtype, void* stream);
/* replaces: aten::max_pool2d_with_indices (+backward) (conv.py:151-169,330-335,520). NHWC, k x k, stride, pad.
 * Backward recomputes the arg-max (first maximum in window scan order, as ATen) instead of storing indices. */
int vs_maxpool_forward(const void* x, void* y, int32_t dtype, int32_t N, int32_t H, int32_t W, int32_t C,
                       int32_t k, int32_t stride, int32_t pad, void* stream);
int vs_maxpool_backward(const void* x, const void* dy, void* dx, int32_t dtype, int32_t N, int32_t H, int32_t W,
                        int32_t C, int32_t k, int32_t stride, int32_t pad, void* stream);
/* replaces: aten::upsample_nearest2d x2 (+backward = 2x2 sum) (conv.py:296-314,371-413). */
int vs_upsample2_forward(const void* x, void* y, int32_t dtype, int32_t N, int32_t H, int32_t W, int32_t C,
                         void* stream);
int vs_upsample2_backward(const void* dy, void* dx, int32_t dtype, int32_t N, int32_t H, int32_t W, int32_t C,
                          void* stream);

/* ---- layout boundary (reference tensors are fp32 NCHW frames) ----------------------------
 * replaces: x.view(B, T*C, H, W) time folding (conv.py:90, mlp_encdec.py:31) and the frame slicing of
 * train.py:63-65,76.  out[n][h][w][t*Cf + c] = frames[n][t0 + t][c][h][w];  frames is [B][T][Cf][H][W] fp32. */
int vs_frames_to_nhwc(const float* frames, int32_t B, int32_t T, int32_t Cf, int32_t H, int32_t W, int32_t t0,
                      int32_t nt, void* out, int32_t dtype, void* stream);
/* NHWC(dtype) <-> NCHW(fp32) for decoder outputs, codes that are feature maps, and their gradients */
int vs_nhwc_to_nchw(const void* in, int32_t dtype, float* out, int32_t N, int32_t C, int32_t H, int32_t W,
                    void* stream);
int vs_nchw_to_nhwc(const float* in, void* out, int32_t dtype, int32_t N, int32_t C, int32_t H, int32_t W,
                    void* stream);

/* ---- losses ---------------------------------------------------------------------------------
 * replaces: F.mse_loss / (a-b).pow(2).mean() / t_codes[:,0].pow(2) reductions (train.py:42,86,139,145-149).
 * acc[0] += sum over b<B, t<T, l<L of (a[b*a_sb + t*a_st + l] - b[b*b_sb + t*b_st + l])^2  (b may be NULL: a^2) */
int vs_sqdiff_sum(const float* a, int64_t a_sb, int64_t a_st, const float* b, int64_t b_sb, int64_t b_st,
                  int64_t B, int64_t T, int64_t L, double* acc, void* stream);
/* da[...] (+)= scale * (g_term[0] + lamb * g_total[0]) * (a - b)  with the same indexing; g_term / g_total are
 * device scalars (upstream gradients of this loss term and of the lamb-weighted total; either may be NULL) */
int vs_sqdiff_backward(const float* a, int64_t a_sb, int64_t a_st, const float* b, int64_t b_sb, int64_t b_st,
                       int64_t B, int64_t T, int64_t L, float scale, const float* g_term, const float* g_total,
                       float lamb, float* da, int32_t accumulate, void* stream);
/* terms[i] = (float)(acc[i] * coef[i]) for i<n;  terms[n] = sum_i lamb[i]*terms[i]  (the total loss) */
int vs_loss_combine(const double* acc, const double* coef_host, const double* lamb_host, int32_t n, float* terms,
                    void* stream);

/* ---- fused decoder tail ---------------------------------------------------------------------
 * replaces: the last BatchNorm + LeakyReLU of the DCGAN decoder followed by ConvTranspose2d(nf, nc, 4, 2, 1) + output
 * activation (conv.py:255-263: `conv[2]` = make_conv_block(ConvTranspose2d(2nf, nf)), `conv[3]` = ConvTranspose2d(nf, nc)),
 * i.e. aten::native_batch_norm + leaky_relu_ + convolution (+ sigmoid) forward and their three backward kernels, WITHOUT
 * materialising the normalised tensor act = act_bn(gamma * (y - mean) * invstd + beta) or its gradient:
 *   forward : out = act_out(convT(act) + bias), act built from y on the operand path of the GEMM + col2im kernel;
 *   wgrad   : dw[K][C][4][4] += sum_pix act[pix][k] * dout[...]          (act rebuilt from y on the operand path);
 *   backward: the gradient w.r.t. act is a direct convolution of `dout` (the gradient w.r.t. the convT's pre-activation
 *             output, [N,H,W,C]) with wp_direct = vs_pack_weight(w, K, C, 16, swap=0); it is produced tile by tile on the
 *             tensor cores and consumed in the epilogue: phase 0 accumulates sums[G][K][2] += {sum dz, sum dz * xhat}
 *             (zeroed by the caller), phase 1 writes dy = gamma * invstd * (dz - mean(dz) - xhat * mean(dz * xhat))
 *             (train != 0) or gamma * invstd * dz (train == 0) and adds the affine gradients from `sums`.
 * g is the geometry of the THIN transposed convolution: big = out [N,H,W,C], small = y [N,P,Q,K]; mean / invstd are
 * [bn_groups][K] as written by vs_bn_finalize.  vs_tail_eligible: 1 when all four kernels accept g (bf16, K == 64,
 * C <= 2, k4 s2 p1, whole 128-pixel tiles), 0 otherwise; the other entry points fail for a non-eligible g. */
int vs_tail_eligible(const vs_conv_geom* g);
int vs_tail_forward(const vs_conv_geom* g, const void* y, const float* mean, const float* invstd, const float* gamma,
                    const float* beta, int32_t bn_groups, int32_t bn_act, const void* wp, const float* bias, void* out,
                    void* stream);
int vs_tail_wgrad(const vs_conv_geom* g, const void* y, const float* mean, const float* invstd, const float* gamma,
                  const float* beta, int32_t bn_groups, int32_t bn_act, const void* dout, float* dw, void* stream);
int vs_tail_bn_backward(const vs_conv_geom* g, const void* y, const float* mean, const float* invstd, const float* gamma,
                        const float* beta, int32_t bn_groups, int32_t bn_act, const void* dout, const void* wp_direct,
                        int32_t phase, int32_t train, double* sums, void* dy, float* dgamma, float* dbeta, void* stream);

/* ---- synthetic Moving-MNIST sequences (the step before the hot path) ----
 * replaces: MovingMNIST.__getitem__ (train branch) + _compute_trajectory + _process_collision of
 * data/moving_mnist.py:112-253, deterministic variant: frames[b][t][0] = min(255, sum over the n_obj objects of the glyph
 * blitted at its bounced position) / 255.  objs [B][n_obj][5] int32 = {glyph index, sx, sy, dx, dy} drawn by the host in
 * the reference's order (sx: first row, sy: first column); glyphs [n_glyphs][gh][gw] uint8; frames [B][T][1][F][F] fp32. */
int vs_moving_sequences(const uint8_t* glyphs, int32_t n_glyphs, int32_t gh, int32_t gw, const int32_t* objs,
                        int32_t n_obj, int32_t B, int32_t T, int32_t F, float* frames, void* stream);

/* ---- gradient exchange over NVLink peer memory (data-parallel training) ----
 * replaces: ncclAllReduce of the gradient arena (the reference has no multi-GPU path; this is the exchange step of the
 * batch-sharded design).  Every rank's arena (and a `world`-word flag array) is mapped into every process (CUDA IPC);
 * *_ptrs_host are HOST arrays of `world` device pointers, entry r = rank r's buffer as addressable from this process.
 * vs_peer_allreduce: rank `rank` sums slice `rank` of all arenas in rank order and writes the sum into all arenas
 * (fused reduce-scatter + all-gather; n fp32 elements, n % 4 == 0; max_blocks caps the grid, 0 = 2 per SM).
 * vs_peer_barrier : all ranks' preceding work on their streams is complete and visible before any rank's following
 * work starts; *epoch_dev (device, zero-initialised, same history on all ranks) counts the barriers. */
int vs_peer_barrier(void* const* flag_ptrs_host, int32_t rank, int32_t world, uint32_t* epoch_dev, void* stream);
int vs_peer_allreduce(void* const* arena_ptrs_host, int32_t rank, int32_t world, int64_t n, int32_t max_blocks,
                      void* stream);
/* bf16 transport of the same exchange (half the NVLink bytes; the usual compressed gradient all-reduce of bf16 training):
 * vs_grad_compress rounds n fp32 gradients to bf16 into a peer-mapped buffer, vs_peer_allreduce_bf16 sums slice `rank` of
 * the `world` bf16 buffers in fp32 and writes the bf16-rounded sum into all of them, vs_grad_expand widens the result
 * back into the fp32 arena.  n % 8 == 0. */
int vs_grad_compress(const float* grad, void* out_bf16, int64_t n, void* stream);
int vs_grad_expand(const void* in_bf16, float* grad, int64_t n, void* stream);
int vs_peer_allreduce_bf16(void* const* buf_ptrs_host, int32_t rank, int32_t world, int64_t n, int32_t max_blocks,
                           void* stream);

/* ---- optimizer -----------------------------------------------------------------------------
 * replaces: torch.optim.Adam.step (main.py:145, train.py:162): eps-outside-sqrt, bias-corrected, no decay.
 * One launch over a flat, 16-byte aligned parameter arena.  The 1-based step count is step_host, or
 * *step_dev when step_dev != NULL (device-resident counter: keeps a captured CUDA graph replayable).
 * grad is multiplied by grad_scale first (1/world_size after a sum all-reduce).  The learning rate is lr, or
 * *lr_dev when lr_dev != NULL (device-resident: MultiStepLR, main.py:146-147, then needs no graph re-capture). */
int vs_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, float lr, float beta1,
                 float beta2, float eps, float grad_scale, int32_t step_host, const int32_t* step_dev,
                 const float* lr_dev, void* stream);

/* ---- latent rollout ------------------------------------------------------------------------
 * replaces: the loop of model.py:78-83 over MLPResnet.forward (resnet.py:42-50, mlp.py:66-71):
 * per time step and block  x <- x + L3(relu(L2(relu(L1 x)))),  one launch for the whole rollout (each CTA owns
 * a few batch rows for all T-1 steps).  fp32 throughout.
 * codes [T][B][d]: codes[0] given, codes[1..] written.  w_host: HOST array of 6*n_blocks device pointers
 * {W1[h][d], b1[h], W2[h][h], b2[h], W3[d][h], b3[d]} per block (torch nn.Linear layout).
 * Saved for backprop-through-time (each may be NULL when no backward follows):
 *   hidden [n_blocks][2][T-1][B][h]  post-ReLU activations of L1 and L2,
 *   xin    [n_blocks][T-1][B][d]     input of every block,   res [n_blocks][T-1][B][d]  its residual. */
int vs_latent_rollout_forward(float* codes, const float* const* w_host, int32_t T, int32_t B, int32_t d, int32_t h,
                              int32_t n_blocks, float* hidden, float* xin, float* res, void* stream);
/* Adjoint recurrence.  dcodes [T][B][d] holds dL/dcodes[t] (from the decoder / losses) on entry; on exit
 * dcodes[0] is the gradient w.r.t. the initial code.  Emits the pre-activation gradients needed for the
 * weight gradients, which are then (T-1)*B-row GEMMs (vs_conv_wgrad / vs_colsum).  w_host here holds the
 * TRANSPOSED weights {W1^T[d][h], -, W2^T[h][h], -, W3^T[h][d], -} per block (bias slots ignored), so that the
 * adjoint products stream rows exactly like the forward ones:
 *   dres [n_blocks][T-1][B][d] = dL/d(residual),  dhidden [n_blocks][2][T-1][B][h] = dL/d(pre-ReLU of L1, L2). */
int vs_latent_rollout_backward(float* dcodes, const float* const* w_host, int32_t T, int32_t B, int32_t d, int32_t h,
                               int32_t n_blocks, const float* hidden, float* dres, float* dhidden, void* stream);

/* ---- evaluation metrics ---------------------------------------------------------------------
 * replaces: var_sep/utils/ssim.py:95-116 (_ssim: five grouped Gaussian conv2d + the SSIM formula, reduction over the
 * map by the caller, test/utils.py:19-24) and the per-frame F.mse_loss(..., 'none').mean([3,4]) of
 * test/mnist/test.py:137.  pred / target: `planes` contiguous fp32 H x W planes (NCHW images, one plane per (image,
 * channel)); kernel: fs x fs window weights (ssim.py:84-92), fs <= 15, H*W <= 4096.  Per plane:
 *   mse_mean[p]  = mean((pred - target)^2),
 *   ssim_mean[p] = mean over the (H-fs+1) x (W-fs+1) valid window positions of the SSIM index,
 *   ssim_map (may be NULL): that map, [planes][H-fs+1][W-fs+1]. */
int vs_ssim_mse_planes(const float* pred, const float* target, int64_t planes, int32_t H, int32_t W, const float* kernel,
                       int32_t fs, float c1, float c2, float* ssim_map, float* ssim_mean, float* mse_mean, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VARSEP_H_ */
