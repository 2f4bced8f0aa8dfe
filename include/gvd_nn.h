/*
 * gvd_nn.h -- C ABI of the B200-native building blocks of the ViewCrafter U-Net denoiser and DDIM sampler.
 *
 * The reference runs these layers through PyTorch library kernels (cuBLAS / cuDNN / ATen) from
 * third_party/ViewCrafter/lvdm/modules/{attention.py,networks/openaimodel3d.py} and
 * lvdm/models/samplers/ddim.py; there is no reference FFI for them, so the boundary is the operator level:
 * each entry point names the reference module/lines it serves.  All pointers are CUDA device pointers, all
 * buffers caller-owned, every call takes the stream; return 0 on success (message: gvd_nn_last_error()).
 */
#ifndef GVD_NN_H_
#define GVD_NN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GVD_NN_API __attribute__((visibility("default")))
#else
#define GVD_NN_API
#endif

typedef struct CUstream_st* gvd_nn_stream_t; /* == cudaStream_t */

enum {
    GVD_ACT_NONE = 0, GVD_ACT_SILU = 1, GVD_ACT_GELU = 2,
    GVD_ACT_ROUND_SCALE = 3, /* C = bf16(bf16(acc) * alpha): the rounding points of `einsum(q,k) * scale` under autocast */
    GVD_ACT_GEGLU = 4        /* fused GEGLU (attention.py:415-423): B's rows (and bias) interleaved in blocks of 16 --
                                rows 32b..32b+15 = value rows 16b.., rows 32b+16..32b+31 = the gate rows of the same outputs;
                                C[m, j] = bf16(value_j) * bf16(gelu(bf16(gate_j))) has N/2 columns (ldc >= N/2).  N % 32 == 0,
                                bf16 output, no residual / bias2 */
};

/* Strided-batched bf16 GEMM on tcgen05 tensor cores (fp32 accumulation in TMEM):
 *     C[b,h,m,n] = act( alpha * sum_k A[b,h,m,k] * B[b,h,n,k] + bias[n] ) (+ bias2[n]) + residual[b,h,m,n]
 * A and B are K-major (k contiguous); every row/batch stride is in ELEMENTS and must be a multiple of 8.
 * Serves nn.Linear / 1x1 conv / im2col'ed 3x3 and (3,1,1) convs (openaimodel3d.py:155-236,255-279), the
 * q/k/v/out projections, GEGLU/FF linears and the QK^T / PV products of CrossAttention (attention.py:81-144). */
typedef struct GvdGemmArgs {
    int M, N, K;
    int batch_h, batch_b;              /* two batch levels (e.g. heads, batch); use 1 when unused */
    const void* A; long long lda, a_stride_h, a_stride_b;   /* bf16 */
    const void* B; long long ldb, b_stride_h, b_stride_b;   /* bf16 */
    void* C;       long long ldc, c_stride_h, c_stride_b;   /* bf16, or fp32 when out_fp32 */
    const float* bias;                 /* [N] fp32 or NULL */
    const float* bias2;                /* [N] fp32 or NULL: added AFTER the bf16 rounding of the layer output
                                          (ResBlock `h + emb_out`, openaimodel3d.py:228) */
    const void* residual;              /* same layout/dtype as C, or NULL; added after bf16 rounding as well */
    float alpha;
    int act;                           /* GVD_ACT_* applied before the residual add */
    int out_fp32;
    int b_mn_major;                    /* 1: B is given as B[k][n] (N contiguous, ldb = elements between consecutive k): the
                                          product C = A * B without a transposed copy of B (tcgen05 MN-major operand).
                                          Needs bf16 output with 16-byte aligned rows; tiles are 128 x 64.  Used for
                                          dK^T = Q^T dS and dV^T = dO^T P in the attention backward. */
} GvdGemmArgs;
GVD_NN_API int gvd_gemm_bf16(const GvdGemmArgs* args, gvd_nn_stream_t stream);

/* Implicit-GEMM convolution on the same tensor-core kernel: no im2col buffer.  The K loop walks (tap, 64-channel block);
 * the A tile of a tap is the activation tile itself, fetched by TMA at shifted coordinates of a 4-D tensor map over
 * x (channels, then the convolved axes), out-of-bounds rows zero-filled by the copy engine (= the zero padding).
 *   kind 1: 3x3 / pad 1 / stride 1 over x[F, H, W, Cin]  (ResBlock / Downsample-free convs, openaimodel3d.py:155-236;
 *           the VAE decoder's convs, ae_modules.py:466-579): y[F, H*W, Cout], weight [Cout, 9*Cin] in (ky, kx, cin) order.
 *           An output tile is 128 consecutive pixels of one frame: needs W % 128 == 0 or 128 % W == 0, and frames
 *           whose pixel count fills its tiles to at least 8/9 (smaller frames stay on the im2col route).
 *   kind 2: (3,1,1) / pad (1,0,0) over x[B, T, S, Cin] (TemporalConvBlock, openaimodel3d.py:246-266): weight [Cout, 3*Cin]
 *           in (kt, cin) order.  Rows are (frame, pixel) flattened, a tap is a shift by S rows.
 * Same epilogue as gvd_gemm_bf16 (bias, activation, bias2, residual with the reference's rounding points).  With the
 * weight re-packed as [Cin, taps*Cout] and its taps reversed the same call computes the data gradient dX of the layer
 * (vc_b200.ops.conv3x3_dx).  Cin % 64 == 0 and Cout % 8 == 0; gvd_conv_bf16_supported() tells (1 / 0) whether a geometry
 * is served -- callers keep the im2col route for the rest (stride 2, fused upsampling, 4-channel ends of the networks). */
typedef struct GvdConvArgs {
    int kind;
    int F, H, W;                       /* kind 1 */
    int B, T; long long S;             /* kind 2 */
    int Cin, Cout;
    const void* x;                     /* bf16, channels-last, pixel stride Cin */
    const void* weight;                /* bf16 [Cout, taps*Cin] */
    void* y;                           /* bf16 [.., Cout] */
    const float* bias;
    const float* bias2;
    const void* residual;              /* same layout as y, or NULL */
    int act;
} GvdConvArgs;
GVD_NN_API int gvd_conv_bf16_supported(int kind, int H, int W, int Cin, int Cout);
GVD_NN_API int gvd_conv_bf16(const GvdConvArgs* args, gvd_nn_stream_t stream);

/* GroupNorm(groups) [+ SiLU] on channels-last activations x[F, S, C] (bf16 in/out, fp32 statistics over S x C/groups
 * per (frame, group)).  Serves normalization()/nn.GroupNorm(32) in ResBlock.in_layers/out_layers
 * (openaimodel3d.py:155-181; per frame: F = b*t, S = h*w), TemporalConvBlock (openaimodel3d.py:255-266; statistics
 * span t*h*w: F = b, S = t*h*w) and the Spatial/TemporalTransformer input norms (attention.py:268,331).
 * do_silu: 0 = none, 1 = SiLU applied to the bf16-rounded norm output (GroupNormSpecific semantics),
 *          2 = SiLU in fp32 before the single bf16 rounding (plain nn.GroupNorm under autocast).
 * tmp: gvd_groupnorm_tmp_floats(F, S, groups) floats of scratch. */
GVD_NN_API size_t gvd_groupnorm_tmp_floats(int F, long long S, int groups);
GVD_NN_API int gvd_groupnorm_cl(const void* x, void* y, const float* gamma, const float* beta, int F, long long S, int C,
                                int groups, float eps, int do_silu, float* tmp, size_t tmp_floats, gvd_nn_stream_t stream);

/* The same normalisation split at its reduction, for a group whose rows are sharded across GPUs (frame- or
 * pixel-sharded TemporalConvBlock / TemporalTransformer norms): `_stats` writes (sum, sum of squares)
 * per (frame, group) of the LOCAL rows into stats[F, groups, 2]; the caller adds the shards' stats (one all-reduce of
 * 2*F*groups floats) and `_apply` normalises the local rows with them, stat_rows = total rows behind the sums. */
GVD_NN_API int gvd_groupnorm_cl_stats(const void* x, float* stats, int F, long long S, int C, int groups, float* tmp,
                                      size_t tmp_floats, gvd_nn_stream_t stream);
GVD_NN_API int gvd_groupnorm_cl_apply(const void* x, void* y, const float* gamma, const float* beta, const float* stats, int F,
                                      long long S, long long stat_rows, int C, int groups, float eps, int do_silu,
                                      gvd_nn_stream_t stream);

/* gvd_groupnorm_cl that also hands back the statistics it normalised with -- stats[F, groups, 2], bit-identical to what
 * gvd_groupnorm_cl_stats writes for the same x -- so a caller that needs them for the backward (the guided sampler's
 * tape) does not pay a separate fold launch per layer. */
GVD_NN_API int gvd_groupnorm_cl_keep_stats(const void* x, void* y, const float* gamma, const float* beta, float* stats, int F,
                                           long long S, int C, int groups, float eps, int do_silu, float* tmp, size_t tmp_floats,
                                           gvd_nn_stream_t stream);

/* LayerNorm over the last dimension of x[rows, C] (bf16 in/out) -- BasicTransformerBlock.norm1/2/3 (attention.py:236-238). */
GVD_NN_API int gvd_layernorm(const void* x, void* y, const float* gamma, const float* beta, long long rows, int C, float eps,
                             gvd_nn_stream_t stream);

/* GEGLU gate: out[r, j] = h[r, j] * gelu(h[r, D + j]), h[rows, 2D] bf16 (attention.py:415-423). */
GVD_NN_API int gvd_geglu(const void* h, void* out, long long rows, int D, gvd_nn_stream_t stream);

/* Row softmax of fp32 or bf16 scores x[rows, ldx] (first `cols` columns) into bf16 probabilities y[rows, ldy]; columns
 * cols..ldy-1 of y are zeroed so y can feed the PV GEMM with K = ldy (attention.py:118,138). */
GVD_NN_API int gvd_softmax_rows(const void* x, int x_is_bf16, long long ldx, void* y, long long ldy, long long rows,
                                int cols, gvd_nn_stream_t stream);

/* im2col for 3x3 / pad 1 convolutions on x[F, H, W, C] bf16 -> col[F, Ho, Wo, 9*C] with K order (ky, kx, c).
 * stride 1 or 2 (Downsample, openaimodel3d.py:61-75); upsample=1 convolves the nearest-neighbour 2x upsampling of x
 * without materialising it (Upsample, openaimodel3d.py:88-103). */
GVD_NN_API int gvd_im2col3x3_cl(const void* x, void* col, int F, int H, int W, int C, int stride, int upsample,
                                gvd_nn_stream_t stream);

/* im2col for the VAE encoder's Downsample (lvdm/modules/networks/ae_modules.py:93-106): zero padding on the right and
 * bottom only, stride 2, no further padding -- x[F, H, W, C] -> col[F, Ho, Wo, 9*C] with Ho = (H - 2) / 2 + 1. */
GVD_NN_API int gvd_im2col3x3_down_cl(const void* x, void* col, int F, int H, int W, int C, gvd_nn_stream_t stream);

/* Nearest-neighbour 2x upsampling of channels-last x[F, H, W, C] -> y[F, 2H, 2W, C] (Upsample.forward:
 * F.interpolate(scale_factor=2, mode="nearest"), openaimodel3d.py / ae_modules.py) and its adjoint, dx[h, w] = the fp32
 * sum of dy's 2 x 2 block rounded once.  They put the Upsample convolutions on the implicit-GEMM route
 * (gvd_conv_bf16 over the 4x tensor) instead of a 36x im2col matrix.  C % 8 == 0. */
GVD_NN_API int gvd_upsample2x_cl(const void* x, void* y, int F, int H, int W, int C, gvd_nn_stream_t stream);
GVD_NN_API int gvd_upsample2x_bwd_cl(const void* dy, void* dx, int F, int H, int W, int C, gvd_nn_stream_t stream);

/* im2col for the (3,1,1) / pad (1,0,0) temporal convolutions on x[B, T, S, C] -> col[B, T, S, 3*C], K order (kt, c)
 * (TemporalConvBlock, openaimodel3d.py:246-266). */
GVD_NN_API int gvd_im2col_t3_cl(const void* x, void* col, int B, int T, long long S, int C, gvd_nn_stream_t stream);

/* Temporal self-attention over T <= 32 frames per (pixel, head), head dim 64: q,k,v,out [B, T, S, H*64] bf16
 * (CrossAttention inside TemporalTransformer, attention.py:365-412 with '(b h w) t c' sequences). */
GVD_NN_API int gvd_temporal_attention(const void* q, const void* k, const void* v, void* out, int B, int T, long long S,
                                      int H, float scale, gvd_nn_stream_t stream);

/* Fused attention, head dim 64: out = softmax(q k^T * scale) v (fp32 logits) without materialising the scores
 * (CrossAttention.forward, attention.py:81-144; the reference's einsum path writes the full [b*h, Nq, Nk] matrix).
 * q, out: [B, Nq, H*64] bf16 (batch stride q_batch_stride elements); k, v: [B, Nk, H*64] bf16 (kv_batch_stride).
 * Shared keys for all batch items (text / image cross-attention): pass B = 1 and Nq = batch * tokens. */
GVD_NN_API int gvd_flash_attention(const void* q, const void* k, const void* v, void* out, int B, int Nq, int Nk, int H,
                                   long long q_batch_stride, long long kv_batch_stride, float scale, gvd_nn_stream_t stream);

/* The same forward, also writing the per-row statistic its backward needs: lse[b, h, i] = log2 sum_j exp2(s_ij * scale *
 * log2 e), fp32, [B, H, Nqp] with Nqp = Nq rounded up to 128 (rows beyond Nq hold finite filler). */
GVD_NN_API int gvd_flash_attention_lse(const void* q, const void* k, const void* v, void* out, float* lse, int B, int Nq, int Nk,
                                       int H, long long q_batch_stride, long long kv_batch_stride, float scale,
                                       gvd_nn_stream_t stream);

/* Backward of gvd_flash_attention without materialising scores (the autograd adjoint of CrossAttention.forward,
 * attention.py:81-144, as ddim_guidance.py:259-337 differentiates it): dq, and dk / dv unless both are NULL (keys and
 * values projected from the frozen context).  out / lse are the forward's results, dout the incoming gradient; q, out,
 * dout, dq share q_batch_stride, k, v, dk, dv share kv_batch_stride; delta is [B, H, Nqp] fp32 scratch (sum_d dO O).
 * lse and delta must be 16-byte aligned.  P and dS are rounded to bf16 before their products, as in the reference's
 * autocast backward; accumulation is fp32. */
typedef struct GvdFlashBwdArgs {
    const void* q;
    const void* k;
    const void* v;
    const void* out;
    const void* dout;
    const float* lse;
    float* delta;
    void* dq;
    void* dk;
    void* dv;
    int B, Nq, Nk, H;
    long long q_batch_stride, kv_batch_stride;
    float scale;
} GvdFlashBwdArgs;
GVD_NN_API int gvd_flash_attention_bwd(const GvdFlashBwdArgs* args, gvd_nn_stream_t stream);

/* One DDIM update, fused (lvdm/models/samplers/ddim.py:206-280 with v-prediction, classifier-free guidance,
 * rescale_noise_cfg (utils_diffusion.py:147-158) and dynamic rescale).  All tensors fp32 with n elements (one batch
 * item); e_uncond may be NULL (no guidance).  scratch: 32 + 4*n bytes. */
typedef struct GvdDdimArgs {
    long long n;
    const float* x;          /* x_t                                       */
    const float* e_cond;     /* model output with conditioning             */
    const float* e_uncond;   /* model output with unconditional cond, or NULL */
    const float* noise;      /* N(0,1) noise for this step                 */
    float* x_prev;           /* out                                        */
    float* pred_x0;          /* out                                        */
    void* scratch;
    float cfg_scale, guidance_rescale;
    float sqrt_alphas_cumprod_t, sqrt_one_minus_alphas_cumprod_t;   /* model schedule at timestep t */
    float ddim_alpha_prev, ddim_sigma, temperature;                 /* DDIM schedule at this index  */
    float scale_t, scale_prev;
    int use_dynamic_rescale;
} GvdDdimArgs;
GVD_NN_API int gvd_ddim_step(const GvdDdimArgs* args, gvd_nn_stream_t stream);

/* ---- input-gradient operators: the guided sampler (lvdm/models/samplers/ddim_guidance.py:259-337) differentiates
 * pred_x0 with respect to the latent x through both U-Net forwards (`pred_x0.backward(gradient=..., inputs=x)`, :309);
 * the reference leaves that to autograd over ATen/cuDNN.  Parameters are frozen there, so only activations get
 * gradients: dX of a linear / conv layer is the tensor-core GEMM above against the transposed weight; the entries
 * below are the adjoints of the memory-bound layers.  bf16 in/out, fp32 arithmetic, same layouts as the forwards. ---- */

/* GroupNorm backward.  stats[F, groups, 2] = (sum x, sum x^2) from gvd_groupnorm_cl_stats over the same rows;
 * do_silu as in the forward (the SiLU derivative is taken where the forward evaluated it).
 * tmp: gvd_groupnorm_bwd_tmp_bytes(F, S, groups) bytes, 8-byte aligned. */
GVD_NN_API size_t gvd_groupnorm_bwd_tmp_bytes(int F, long long S, int groups);
GVD_NN_API int gvd_groupnorm_cl_bwd(const void* x, const void* dy, void* dx, const float* gamma, const float* beta,
                                    const float* stats, int F, long long S, int C, int groups, float eps, int do_silu,
                                    void* tmp, size_t tmp_bytes, gvd_nn_stream_t stream);

/* The same backward split at its reduction, for a group whose rows are sharded across GPUs (the mirror image of
 * gvd_groupnorm_cl_stats / _apply): `_sums` writes (sum g, sum g*xh) per (frame, group) of the LOCAL rows into
 * sums[F, groups, 2] (float64), the caller adds the shards' sums (one all-reduce), `_apply` forms dx of the local rows.
 * stats = the already summed forward statistics, stat_rows = total rows behind them. */
GVD_NN_API int gvd_groupnorm_cl_bwd_sums(const void* x, const void* dy, const float* gamma, const float* beta, const float* stats,
                                         double* sums, int F, long long S, long long stat_rows, int C, int groups, float eps,
                                         int do_silu, void* tmp, size_t tmp_bytes, gvd_nn_stream_t stream);
GVD_NN_API int gvd_groupnorm_cl_bwd_apply(const void* x, const void* dy, void* dx, const float* gamma, const float* beta,
                                          const float* stats, const double* sums, int F, long long S, long long stat_rows, int C,
                                          int groups, float eps, int do_silu, gvd_nn_stream_t stream);

/* LayerNorm backward over the last dimension: x, dy, dx [rows, C] bf16. */
GVD_NN_API int gvd_layernorm_bwd(const void* x, const void* dy, void* dx, const float* gamma, long long rows, int C,
                                 float eps, gvd_nn_stream_t stream);

/* GEGLU backward: h [rows, 2D] (the forward's input), dout [rows, D] -> dh [rows, 2D]. */
GVD_NN_API int gvd_geglu_bwd(const void* h, const void* dout, void* dh, long long rows, int D, gvd_nn_stream_t stream);

/* Row softmax backward: ds = p o (dp - sum_j p_j dp_j) over the first `cols` columns of p, dp, ds [rows, ld] bf16
 * (ds may alias dp); columns cols..ld-1 of ds are zeroed so it can feed a GEMM with K = ld. */
GVD_NN_API int gvd_softmax_bwd_rows(const void* p, const void* dp, void* ds, long long ld, long long rows, int cols,
                                    gvd_nn_stream_t stream);

/* Adjoints of the two im2col layouts: dcol [F, Ho, Wo, 9*C] -> dx [F, H, W, C] (same stride / upsample meaning as
 * gvd_im2col3x3_cl; every input pixel gathers its taps, no atomics) and dcol [B, T, S, 3*C] -> dx [B, T, S, C]. */
GVD_NN_API int gvd_col2im3x3_cl(const void* dcol, void* dx, int F, int H, int W, int C, int stride, int upsample,
                                gvd_nn_stream_t stream);
GVD_NN_API int gvd_col2im_t3_cl(const void* dcol, void* dx, int B, int T, long long S, int C, gvd_nn_stream_t stream);

/* Backward of gvd_temporal_attention: q, k, v, dout -> dq, dk, dv, all [B, T, S, H*64] bf16 (the probabilities are
 * recomputed with the forward's rounding points). */
GVD_NN_API int gvd_temporal_attention_bwd(const void* q, const void* k, const void* v, const void* dout, void* dq,
                                          void* dk, void* dv, int B, int T, long long S, int H, float scale,
                                          gvd_nn_stream_t stream);

/* Vector-Jacobian product of the guided step's pred_x0 arithmetic (ddim_guidance.py:263-278: CFG mix,
 * rescale_noise_cfg, predict_start_from_z_and_v, dynamic rescale): given grad_pred_x0 = dL/dpred_x0 writes
 * dx = the gradient through the explicit x_t term and de_cond / de_uncond = the cotangents of the two U-Net outputs.
 * All tensors fp32 with n elements; e_uncond / de_uncond may be NULL together.  scratch: 64 bytes, 8-byte aligned. */
typedef struct GvdDdimVjpArgs {
    long long n;
    const float* e_cond;
    const float* e_uncond;
    const float* grad_pred_x0;
    float* dx;
    float* de_cond;
    float* de_uncond;
    void* scratch;
    size_t scratch_bytes;
    float cfg_scale, guidance_rescale;
    float sqrt_alphas_cumprod_t, sqrt_one_minus_alphas_cumprod_t;
    float scale_t, scale_prev;
    int use_dynamic_rescale;
} GvdDdimVjpArgs;
GVD_NN_API int gvd_ddim_pred_x0_vjp(const GvdDdimVjpArgs* args, gvd_nn_stream_t stream);

/* Kernel generation of the memory-bound operators: 0 = the round-1 kernels; 1 (default) = GEGLU with 16-byte vectors,
 * gvd_im2col3x3_cl / gvd_im2col_t3_cl with 32-bit index arithmetic (same results bit for bit, 1.3-2.0x faster on B200) and
 * temporal attention on mma.sync tiles (csrc/tattn_mma.cu: same rounding points, 4.8x faster); 2 = level 0's temporal
 * attention with K / V staged in fp32 (measured 10 % slower than level 0, kept for A/B timing).
 * Default: the environment variable GVD_NN_FAST ("0" / "1" / "2"), read at the first call.  on in 0..2 sets the level,
 * any other value only queries; returns the previous level. */
GVD_NN_API int gvd_nn_set_fast(int on);

GVD_NN_API const char* gvd_nn_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* GVD_NN_H_ */
