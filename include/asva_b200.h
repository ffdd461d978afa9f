/*
 * asva_b200.h — C ABI of libasva_b200.so: the B200 (sm_100a) kernels behind ASVA's denoising hot path.
 *
 * The reference (lzhangbj/ASVA) is pure Python/PyTorch and has no FFI of its own; its operator seams are
 * Python callables.  Each entry point below replaces the ATen/cuDNN/cuBLAS library calls
 * that one reference function dispatches; the reference file:line it stands in for is cited per function.
 * All pointers are DEVICE pointers owned by the caller (borrowed from torch tensors); the library never
 * allocates device memory, never synchronises, and is CUDA-graph capturable.  Every function returns 0 on
 * success or a negative asva_status; asva_last_error() gives the message (thread-local).
 *
 * Activation layout everywhere: channels-last tokens  x[b][f][y][x][c]  (bf16), i.e. the reference's
 * "(b f) (h w) c" token layout (ff_spatio_audio_temp_transformer_3d.py:121) kept for the whole UNet.
 */
#ifndef ASVA_B200_H
#define ASVA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* asva_stream_t; /* cudaStream_t */

enum asva_status {
  ASVA_OK = 0,
  ASVA_ERR_INVALID = -1, /* bad descriptor (shape/alignment/unsupported size) */
  ASVA_ERR_CUDA = -2,    /* CUDA runtime / driver error at launch */
  ASVA_ERR_DEVICE = -3   /* not an sm_100 device */
};

/* ------------------------------------------------------------------------------------------------------------
 * Generic tensor-core GEMM   out[M, N] = epilogue( A[M, K] * W[N, K]^T )        (tcgen05.mma, TMEM accumulators,
 * TMA-fed).  A is described as a 4-D view (c, d1, d2, d3) of up to two bf16 sources; a tile of <=128 output rows
 * is a box over (d1, d2, d3) and the K loop walks a table of segments, each a (source, channel offset, coordinate
 * offset) triple -> this one kernel is the plain linear layer, the 1x1 conv, the implicit-GEMM 3x3 conv
 * (9 segments = taps, zero padding by TMA out-of-bounds fill, stride 2 by TMA traversal strides), the two-source
 * skip-concat 1x1 conv and the temporal 3-tap "conv_temp" GEMM (segments = current / previous frame).
 * Replaces: nn.Conv2d + nn.Linear inside FFInflatedConv3d (avgen/models/unets/utils.py:22-57), the diffusers
 * Attention projections and FeedForward/GEGLU linears (ff_spatio_audio_temp_transformer_3d.py:199-276).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct asva_gemm_seg {
  int32_t src;      /* 0 -> a[0], 1 -> a[1] */
  int32_t c0;       /* first channel inside the source */
  int32_t off[3];   /* input-space coordinate offsets along d1..d3 (may be negative: zero fill) */
  int32_t num_kb;   /* number of 64-wide K blocks in this segment */
  int32_t wk;       /* first column of W this segment multiplies (multiple of 64) */
  int32_t wk_first; /* >= 0: column of W used instead of wk by tiles whose d2 origin is 0 (needs box[1] == 1) */
  int32_t fix2;     /* >= 0: absolute d2 coordinate of the rows read, whatever the tile's d2 position (box[1] == 1) */
  int32_t reserved;
} asva_gemm_seg;

typedef struct asva_rowadd { /* fp32 addend  ptr[(row / div) * ld + col]  (one addend row per block of div rows) */
  const float* ptr;          /* NULL = disabled */
  int64_t ld;
  int32_t div;
  int32_t reserved;
} asva_rowadd;

#define ASVA_GEMM_MAX_SEG 10

typedef struct asva_gemm_desc {
  /* A operand */
  const void* a[2];        /* bf16 sources; a[1] may be NULL */
  int64_t a_dims[2][4];    /* per source: extents (c, d1, d2, d3) in input space */
  int64_t a_strides[2][3]; /* per source: element strides of d1, d2, d3 (c is contiguous) */
  int32_t box[3];          /* tile extents along d1..d3 in output space; product <= 128 */
  int32_t trav[3];         /* traversal stride (1 or 2) along d1..d3: in = out * trav + off */
  int32_t out_dims[3];     /* output-space extents along d1..d3; M = product; row = (o3*D2 + o2)*D1 + o1 */
  int32_t nseg;
  asva_gemm_seg seg[ASVA_GEMM_MAX_SEG];
  /* W operand: [N, wcols] row-major bf16 (leading dimension ldw) */
  const void* w;
  int64_t ldw;
  int32_t N;
  int32_t K;     /* contraction length = 64 * sum(seg.num_kb) */
  int32_t wcols; /* columns of W that exist (>= every seg.wk + 64*num_kb) */
  int32_t cta_group; /* 0 = auto, 1 = one CTA per 128-row tile, 2 = CTA pairs (tcgen05 cta_group::2: 256-row tiles,
                        each CTA of the pair loads half of the W tile) */
  /* epilogue: out = acc + bias[col] + add + res[0] + res[1]   (GEGLU: (h + bias_h) * gelu_erf(g + bias_g));
   * with the LayerNorm fold (ln_cols below) acc is replaced by rstd_row * (acc - mean_row * ln_wsum[col]) */
  const float* bias; /* [N] fp32 or NULL */
  asva_rowadd add;
  const void* res[2]; /* bf16 residuals res[i][row * res_ld[i] + col]; NULL = disabled; may alias out */
  int64_t res_ld[2];
  int32_t geglu; /* 1: each 128-column tile holds [64 value | 64 gate] columns; writes N/2 columns */
  int32_t out_fp32;
  void* out; /* [M][ldo] bf16 (or fp32), row-major; written by TMA stores (tails clipped) */
  int64_t ldo;
  int32_t block_n; /* 0 = auto; 64 / 128 / 160 / 256 */
  int32_t split_k; /* 0 = auto, 1 = off, n = split the K loop n ways (fp32 partials in ws, then a reduce+epilogue
                      kernel); ignored (off) when ws is too small or for GEGLU */
  void* ws;        /* device scratch for split-K partial sums, or NULL */
  int64_t ws_bytes;
  int32_t epilogue; /* 0 = auto; 1 = panel epilogue (residual panels arrive by TMA, output leaves by TMA store);
                       2 = per-warp epilogue (each warp finishes its own 32 x 32 sub-panels: direct residual loads
                       and 16-byte stores; bf16, non-GEGLU, non-split outputs only - otherwise 1 is used);
                       3 = warp-private TMA epilogue (each of the eight epilogue warps moves the 32 rows of its TMEM
                       quadrant with its own TMA loads / stores, no block-level barrier; any output type; needs a
                       row box whose 32-row quadrants are themselves boxes - otherwise 1 is used);
                       4 = cluster split-K: with split_k = 2 / 4 / 8 the K splits of a tile are the CTAs of one
                       thread-block cluster - partials stay in tensor memory, column slices are exchanged through
                       distributed shared memory and every CTA finishes one slice (no workspace, no reduce kernel);
                       cta_group 1, block_n / split_k a multiple of 32, no GEGLU - otherwise 3 or 1 is used */
  /* Row statistics and the LayerNorm fold (ff_spatio_audio_temp_transformer_3d.py:288-362: every sub-block of a
   * BasicTransformerBlock is LayerNorm -> projection).  With W' = W * gamma (column scaling, folded into the packed
   * weight), wsum[n] = sum_k W'[n][k] and bias' = W beta + bias,
   *     LayerNorm(x) W^T + bias  =  rstd_r * (x W'^T - mean_r * wsum) + bias'
   * so the projection runs on the un-normalised rows and the epilogue applies the two per-row scalars - no separate
   * LayerNorm pass.  The per-row sums come from the GEMM that produced x: with stats_out set, a GEMM also writes, for
   * every output row and every 32-column slot, (sum, sum of squares) of the fp32 values it stores:
   *     stats[slot][row][2] fp32, slot = col / 32, row as in out_dims (N % 32 == 0; not with GEGLU / fp32 outputs).
   * A consumer names such a table in ln_stats; its output row r reads statistics row
   *     (r / ln_grp_rows) * ln_grp_stride + r % ln_grp_rows      (ln_grp_rows = 0: row r)
   * and sums the ln_cols / 32 slots in order (deterministic).  ln_cols = 0 turns the fold off.  Folded launches do
   * not split K and use epilogue form 1 or 3. */
  int32_t ln_cols;        /* width C of the rows the statistics describe (the K of this GEMM); 0 = off */
  float* stats_out;       /* [N/32][M][2] fp32 or NULL */
  const float* ln_stats;  /* [ln_cols/32][ln_stat_rows][2] fp32 */
  const float* ln_wsum;   /* [N] fp32 */
  int64_t ln_stat_rows;
  int32_t ln_grp_rows, ln_grp_stride;
  float ln_eps;
  int32_t reserved;
} asva_gemm_desc;

int asva_gemm(const asva_gemm_desc* d, asva_stream_t stream);

/* The tile plan asva_gemm would use for `d` (its cost model's choice where block_n / split_k / cta_group are 0). */
int asva_gemm_plan(const asva_gemm_desc* d, int32_t* block_n, int32_t* split_k, int32_t* cta_group, int32_t* stages,
                   int32_t* epilogue);

/* Measures every feasible (block_n, split_k, cta_group, epilogue) plan of `d` on the device (CUDA events on `stream`, `reps`
 * launches each after one warm-up) and returns the fastest.  Runs the GEMM repeatedly - the output (and anything that
 * aliases it) is scratch afterwards - and synchronises with the stream, so it must not be called during graph
 * capture.  Callers cache the result per problem shape and pass it in block_n / split_k / cta_group /
 * epilogue. */
int asva_gemm_tune(const asva_gemm_desc* d, asva_stream_t stream, int32_t reps, int32_t* block_n, int32_t* split_k,
                   int32_t* cta_group, int32_t* epilogue, float* best_us);

/* ------------------------------------------------------------------------------------------------------------
 * Fused softmax(Q K^T * scale [+ mask]) V on tcgen05 (flash-style online softmax, S and O in TMEM).
 * Replaces F.scaled_dot_product_attention in FFAttnProcessor (utils.py:151-153: first-frame spatial attention,
 * all frames of a clip share the keys/values of frame 0) and in diffusers AttnProcessor2_0 for attn_audio /
 * attn2 (ff_spatio_audio_temp_transformer_3d.py:315-341; bool mask, True = attend).
 *   q   : bf16 [G*R][ldq]  token-major projection output, head h at columns h*d .. h*d+d-1 (the kernel's TMA box
 *         zero-fills the padding up to dpad, so no head-split copy exists)
 *   kv  : bf16 rows of ldkv elements; key j of group g is row g*kv_rows_per_group + j; K at column
 *         k_col0 + head*d, V at column v_col0 + head*d
 *   mask: uint8 [G*R / mask_rows][mask_ld] (1 = attend) or NULL; query row r of group g uses mask row
 *         (g*R + r) / mask_rows
 *   out : bf16 [G*R][ldo], head h at columns h*d .. h*d+d-1
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct asva_attn_desc {
  const void* q;
  const void* kv;
  const uint8_t* mask;
  void* out;
  int64_t ldq, ldkv, ldo, mask_ld;
  int32_t G, heads, R, Nk, d, dpad;
  int32_t kv_rows_per_group, k_col0, v_col0, mask_rows;
  float scale;
  int32_t form; /* 0 = auto and 1 = tcgen05 flash kernel (attn_tc.cu); 2 = warp-MMA kernel for small key sets
                   (attn_mma.cu: Nk <= 128, d <= 160 - K and V of a group staged once in shared memory, 16 query rows
                   per warp on mma.sync).  Form 2 is an alternative kept for comparison: on B200 it measured slower
                   than the tcgen05 kernel on every cross-attention / low-resolution shape of the workload - its time
                   grows with the mma.sync count (22 us at 25 keys, 38 us at 77), profiles/r2_attn_mma.md */
} asva_attn_desc;

int asva_attention(const asva_attn_desc* d, asva_stream_t stream);

/* Temporal self-attention core over the frame axis (F x F per pixel and head).
 * Replaces the SDPA inside attn_temp (ff_spatio_audio_temp_transformer_3d.py:352-358).
 *   qkv : bf16 [B][F][N][3C] (q | k | v), out: bf16 [B][F][N][C]
 * The op moves 8 C F bytes per pixel for 4 F^2 C flops - memory-bound.  Forms (asva_temporal_attention_form):
 *   3  warp-MMA (misc.cu): a CTA streams the q | k | v rows of a few pixels into shared memory with bulk-async copies,
 *      one warp per (pixel, head) runs S = Q K^T, the softmax and P V on mma.sync m16n8k16 fragments, outputs leave
 *      by bulk-async stores.  F <= 32.  The default wherever it serves.
 *   1  tcgen05 (attn_tc.cu): blocks of floor(KV/F) pixels x F frames are queries and keys of one block-diagonal
 *      128-row tile gathered by 5-D TMA boxes.  F <= 64; the fallback for F > 32.
 *   2  one CUDA-core thread per (pixel, head, query frame), same data path as 3 (kept as the measured baseline).
 *   0  auto. */
int asva_temporal_attention(const void* qkv, void* out, int32_t B, int32_t F, int32_t N, int32_t heads,
                            int32_t d, float scale, asva_stream_t stream);
int asva_temporal_attention_form(const void* qkv, void* out, int32_t B, int32_t F, int32_t N, int32_t heads,
                                 int32_t d, float scale, int32_t form, asva_stream_t stream);

/* LayerNorm over C (eps, affine) of x[M][C] (+ optional positional rows pos[F][C] added BEFORE the norm, frame
 * index = (row / N) % F) -> bf16.  Replaces nn.LayerNorm norm1/norm_audio/norm2/norm_temp/norm3
 * (ff_spatio_audio_temp_transformer_3d.py:288-362; pos added at :352). */
int asva_layernorm(const void* x, const float* gamma, const float* beta, const float* pos, void* out, int64_t M,
                   int32_t C, float eps, int32_t N, int32_t F, asva_stream_t stream);

/* GroupNorm statistics over channels-last data.  Instance i covers rows [i*rows, (i+1)*rows) of the (virtually
 * concatenated) sources x0[.,C0] | x1[.,C1]; group g covers channels [g*(C0+C1)/groups, ...).  Writes the folded
 * per-channel affine  stats[i][c] = (scale, shift)  with scale = rstd_g * gamma[c], shift = beta[c] - mean_g * scale,
 * so that GroupNorm(x)[c] = x * scale + shift.  Replaces the statistics half of nn.GroupNorm in
 * FFSpatioTempResnetBlock3D (resnets/ff_spatio_temp_resnet_3d.py:164,175; instance = one clip, all frames), the
 * per-frame GroupNorm of the transformer (ff_spatio_audio_temp_transformer_3d.py:117) and conv_norm_out
 * (audio_cond_unet_3d_condition.py:791). */
int asva_groupnorm_stats(const void* x0, int32_t C0, const void* x1, int32_t C1, int32_t n_inst, int64_t rows,
                         int32_t groups, float eps, const float* gamma, const float* beta, float* stats,
                         float* partial_ws, asva_stream_t stream);
/* number of floats asva_groupnorm_stats needs in partial_ws */
int64_t asva_groupnorm_ws_floats(int32_t n_inst, int64_t rows, int32_t C /* C0 + C1 */);

/* GroupNorm apply (+ optional SiLU, + optional nearest 2x spatial upsample, + channel concat of two sources)
 * -> bf16 [n_img][h_out][w_out][C0+C1] with the (scale, shift) table of asva_groupnorm_stats.  upsample=1 replicates
 * each source pixel 2x2 (F.interpolate nearest, ff_spatio_temp_resnet_3d.py:47); stats==NULL skips the affine. */
int asva_groupnorm_apply(const void* x0, int32_t C0, const void* x1, int32_t C1, const float* stats, int32_t n_inst,
                         int32_t n_img, int32_t h, int32_t w, int32_t silu, int32_t upsample, void* out,
                         asva_stream_t stream);

/* Fused GroupNorm (+ optional SiLU, + channel concat of two sources): statistics and apply in ONE launch, the form the
 * engine uses wherever no upsample sits between the statistics and the apply (every nn.GroupNorm of the UNet:
 * ff_spatio_temp_resnet_3d.py:164,175, ff_spatio_audio_temp_transformer_3d.py:117,
 * audio_cond_unet_3d_condition.py:791).  Same operand meaning as asva_groupnorm_stats / asva_groupnorm_apply;
 * out bf16 [n_inst*rows][C0+C1].  The grid synchronises through sync_ws (asva_groupnorm_sync_bytes() bytes of device
 * memory, zeroed ONCE by the caller at allocation; the kernel leaves it zeroed), so launches that share a sync_ws
 * must be ordered on one stream.  Needs n_inst * groups <= 4096. */
int asva_groupnorm(const void* x0, int32_t C0, const void* x1, int32_t C1, int32_t n_inst, int64_t rows,
                   int32_t groups, float eps, const float* gamma, const float* beta, int32_t silu, void* out,
                   void* sync_ws, asva_stream_t stream);
int64_t asva_groupnorm_sync_bytes(void);
/* Which kernels asva_groupnorm runs for a shape: 0 = cluster kernel (DSMEM exchange, one launch), 1 = grid-barrier
 * kernel (one launch), 2 = two launches (full-row statistics to per-group partials, then reduce + apply): the
 * few-instances x many-rows norms of the top resolution.  -1: invalid shape. */
int asva_groupnorm_form(int32_t n_inst, int64_t rows, int32_t C, int32_t groups);

/* conv_in front end: fp32 latents [Bs][Cl][F][h][w] (Cl<=7) -> bf16 im2col rows [B*F*h*w][64] for the 3x3, pad 1
 * conv (column = tap*Cl + c, zero padded to 64); batch b reads latent b % Bs (CFG duplication,
 * pipeline_audio_cond_animation.py:331-336). */
int asva_conv_in_im2col(const float* latents, void* out, int32_t B, int32_t Bs, int32_t Cl, int32_t F, int32_t h,
                        int32_t w, asva_stream_t stream);

/* conv_temp operand gather (utils.py:43-52 builds exactly this with index + cat): y bf16 [B][F][N][C] ->
 * out bf16 [B*F*N][3C] = [y_f | y_max(f-1,0) | y_0].  Used only where a frame has fewer than 128 pixels, so that the
 * temporal conv runs as one plain asva_gemm with full tiles; the larger levels read the three taps in place. */
int asva_tconv_gather(const void* y, void* out, int32_t B, int32_t F, int32_t N, int32_t C, asva_stream_t stream);

/* conv_out back end: y fp32 [B*F*h*w][ldy] (first Co columns valid) -> conv_temp (Linear(3Co -> Co) over
 * [frame 0 | previous frame | current frame], utils.py:43-53) -> fp32 [B][Co][F][h][w]. */
int asva_conv_out_finish(const float* y, int32_t ldy, const float* wt, const float* bt, float* out, int32_t B,
                         int32_t Co, int32_t F, int32_t h, int32_t w, asva_stream_t stream);

/* Skinny linear for M <= 32 rows, fp32 activations, bf16 weights [N][K]: out = act_out(act_in(x) W^T + b).
 * act: 0 none, 1 SiLU.  Replaces TimestepEmbedding and the 22 time_emb_proj linears
 * (audio_cond_unet_3d_condition.py:673-681, ff_spatio_temp_resnet_3d.py:170-171) in three launches. */
int asva_small_linear(const float* x, const void* w, const float* bias, float* out, int32_t M, int32_t N,
                      int32_t K, int32_t act_in, int32_t act_out, asva_stream_t stream);

/* Sinusoidal timestep features (diffusers get_timestep_embedding, flip_sin_to_cos, shift 0): out[B][dim] fp32 =
 * [cos(t w_i) | sin(t w_i)], w_i = exp(-ln(10000) i / (dim/2)).  t is read from device memory. */
int asva_timestep_features(const float* t, float* out, int32_t B, int32_t dim, int32_t flip_sin_to_cos,
                           asva_stream_t stream);

/* Row softmax  probs[r][c] = softmax_c(scale * scores[r][c]): fp32 scores (an asva_gemm output) -> bf16 probabilities
 * (an asva_gemm operand).  Serves attention whose single head is wider than asva_attention tiles: the mid-block
 * attention of the SD-1.5 VAE decoder (diffusers Attention, heads 1, dim_head 512) behind
 * pipeline_audio_cond_animation.py:205-213 `decode_latents`.  cols % 4 == 0. */
int asva_softmax_rows(const float* scores, int64_t lds, void* probs, int64_t ldp, int64_t rows, int32_t cols,
                      float scale, asva_stream_t stream);

/* Classifier-free guidance + sampler update on frames 1..F-1 (frame 0 is the conditioning image and is never
 * written, pipeline_audio_cond_animation.py:363-364).  eps: fp32 [k][clips][C][F][h][w] UNet outputs (k <= 3 CFG
 * branches, branch-major like the pipeline's torch.cat([latents] * k), :331-336); latents: fp32 [clips][C][F][h][w]
 * updated in place; every clip of a launch is at the same sampler step.  All scalars live in DEVICE memory so the
 * launch can sit inside a replayed CUDA graph:  coef fp32[9] = {w0, w1, w2, c_sample, c_eps, a0, a1, a2, a3}
 *   e = w0*eps[0] + w1*eps[1] + w2*eps[2]                                   (CFG combine, :349-361)
 *   DDIM (eta = 0, epsilon prediction):  x <- c_sample * x + c_eps * e
 *   PLMS (PNDMScheduler.step_plms):      e_hat = a0*e + a1*hist[slots[1]] + a2*hist[slots[2]] + a3*hist[slots[3]];
 *                                        if slots[0] >= 0: hist[slots[0]] <- e;   x <- c_sample * x + c_eps * e_hat
 * hist: fp32 [4][clips][C][F][h][w] ring of past e; slots: int32[4] in device memory. */
int asva_cfg_ddim_step(const float* eps, int32_t k, int32_t clips, float* latents, const float* coef, int32_t C,
                       int32_t F, int32_t hw, asva_stream_t stream);
int asva_cfg_plms_step(const float* eps, int32_t k, int32_t clips, float* latents, float* hist, const float* coef,
                       const int32_t* slots, int32_t C, int32_t F, int32_t hw, asva_stream_t stream);

const char* asva_last_error(void);
int asva_version(void);
/* 0 if the current device is sm_100, ASVA_ERR_DEVICE otherwise */
int asva_device_check(void);

#ifdef __cplusplus
}
#endif
#endif /* ASVA_B200_H */
