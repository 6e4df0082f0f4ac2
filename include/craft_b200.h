/* craft_b200 -- C ABI of the B200-native CRAFT hot path (libcraft_b200.so).
 *
 * Every entry point takes plain device pointers, sizes and a cudaStream_t (as void*); no torch
 * types cross this boundary.  All return 0 on success or a negative code; the message is
 * available from craft_b200_last_error().  Pointers are device pointers unless noted.
 *
 * Layout convention ("padded-flat token grid"): a feature map of h x w
 * tokens is stored token-major with rows p = y*(w+2) + x; the two trailing cells of every grid
 * row are zero halo cells.  Mp = h*(w+2).  bf16 buffers are row-major [Mp, ld].
 *
 * Each function names the reference code it replaces (paths relative to askerlee/craft).
 */
#ifndef CRAFT_B200_H_
#define CRAFT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRAFT_B200_ABI_VERSION 2
#define CRAFT_MAX_TAPS 49

int craft_b200_abi_version(void);
const char* craft_b200_last_error(void);
/* number of kernels this library has launched so far in this process (host-side counter) */
long long craft_b200_launch_count(void);
/* GEMM launches that took the experimental big-box TMA path (CRAFT_GEMM_BIGBOX=1|2); 0 otherwise.   */
long long craft_b200_bigbox_gemm_count(void);
/* device properties the host uses for grid sizing: [0]=sm count, [1]=cc major, [2]=cc minor */
int craft_b200_device_info(int* out3);

/* ---- layout ------------------------------------------------------------------------------ */
/* NCHW f32 -> token rows.  mode: 0 copy, 1 LayerNorm over C (eps 1e-12, no affine), 2 tanh,
 * 3 relu.  C in {128, 256}.  Replaces SETransInputFeatEncoder.forward
 * core/setrans.py:791-795 (mode 1) and the tanh/relu split core/network.py:209-211.       */
int craft_pack_tokens(const float* src_nchw, int C, int H, int W, int mode, void* out_bf16,
                      int ldb, int colb, float* out_f32, int ldf, int colf, void* stream);
/* Same packer for a channels-last source [H][W][ldc] (f16 if src_is_half else f32 -- what the fused encoders
 * produce): channels [c0, c0+C) of every token, C in {128, 256}; mode 4 = relu then LayerNorm.             */
int craft_pack_tokens_nhwc(const void* src, int src_is_half, int ldc, int c0, int C, int H, int W, int mode,
                           void* out_16, int ldb, int colb, float* out_f32, int ldf, int colf, void* stream);
/* token rows (bf16 if is_bf16 else f32) -> NCHW f32 */
int craft_unpack_tokens(const void* src, int is_bf16, int ld, int col, int C, int H, int W,
                        float* dst_nchw, void* stream);

/* ---- tcgen05 shift-GEMM -------------------------------------------------------------------
 * D[m,n] = sum_t sum_k A[m + tap_off[t], a_koff + k] * B[t*Npad + n, b_koff + k]
 * Replaces every nn.Linear / nn.Conv2d on the hot path: query/key projections
 * core/setrans.py:507-508, first_linear :373, BasicMotionEncoder core/update.py:80-86,
 * SepConvGRU :49-64, FlowHead :16, mask head :124-127,161.                                 */
typedef struct craft_gemm_args {
  const void* A;      /* bf16 [a_rows, lda]                                                  */
  int a_rows, lda, a_koff;
  const void* B;      /* bf16 [T*Npad, ldb_]                                                  */
  int b_rows, ldb_, b_koff;
  int b_blocked;      /* !=0: B rows are grid tokens (b_H x b_W); n-tile j = 8 x BN/8 spatial block */
  int b_H, b_W;
  int M, Npad, K, T;
  int BN;             /* CTA tile width: 32, 64, 128 or 256 (Npad % BN == 0)                 */
  int cluster;        /* M tiles per thread-block cluster sharing the weight tile: 0 auto,1,2,4 */
  int stages;         /* TMA pipeline depth: 0 = maximum that fits                            */
  int tap_off[CRAFT_MAX_TAPS];
  int H, W;           /* >0: rows are a padded-flat grid, halo rows are not written; 0: plain */
  int epilogue;       /* 0 store, 1 gru_zr, 2 gru_q, 3 motion, 4 flow (delta + coords1/flow update) */
  float alpha;
  int act;            /* 0 none, 1 relu                                                       */
  const float* bias;  /* [Npad] or NULL                                                       */
  void* out_bf16;     /* row-major, may be NULL                                               */
  int ldo_b, colo_b;
  float* out_f32;
  int ldo_f, colo_f;
  float* aux0;        /* gru: Z [M,128]; flow: coords1 [M,2]                                  */
  float* aux1;        /* gru: Hm [M,128]; motion / flow: flow [M,2]                           */
  int a_share;        /* experimental, default 0.  1: the taps come in groups of consecutive row offsets (the kw
                         taps of one kernel row); the A rows of a group are loaded once and every tap reads
                         them through a row-shifted descriptor.  Correct, but currently slower */
} craft_gemm_args;
int craft_shift_gemm(const craft_gemm_args* a, void* stream);

/* ---- scores: correlation volume build + attention LSE ------------------------------------ */
typedef struct craft_scores_args {
  const void* Q;      /* bf16 [Mp, C] projected queries (token rows)                          */
  const void* K;      /* bf16 [Mp, C] projected keys                                          */
  int C, M, d;        /* channels, modes, per-mode dim (C == M*d)                             */
  int H, W;
  float scale;        /* 1/sqrt(d)                                                            */
  float w_pos;        /* pos_code_weight                                                      */
  const float* pos_table;   /* [(2R+1)^2] or NULL                                             */
  int R;
  const float* clip;  /* device scalar: +inf or attn_clip                                     */
  const int* run_flag;/* NULL, or device int: kernel is a no-op when *run_flag == 0           */
  int ksplit;         /* 0 = auto                                                             */
  /* corr build */
  float w_agg;        /* attn_softaggr.feat2score.weight                                      */
  double* stat_sum;   /* [2] sum, sum of squares (must be zeroed by the caller)               */
  float* stat_max;    /* [1] running max of raw scaled scores (caller initialises to -inf)    */
  float* lvl[4];      /* pyramid volumes [Mp][h_l*w_l]; lvl[0] may be NULL                    */
  /* lse */
  void* lse_part;     /* float2 [ksplit][M][Mp] scratch                                       */
  float* lse2;        /* [M][Mp] out                                                          */
  int mask_radius;    /* attn_lse only; > 0: keys with max(|dy|,|dx|) > mask_radius are masked out
                         (--f2radius, SelfAttVisPosTrans.forward core/setrans.py:580-584); <= 0: off */
  void* lvl0_h16;     /* corr build, optional: level 0 in 16 bits.  Mp*nblk*64 IEEE fp16 (both precision tiers)
                         [Mp][nblk = ceil(H/8)*ceil(W/8)][64]: 8x8 key block (by,bx) -> block by*ceil(W/8)+bx,
                         cell (y%8)*8 + x%8, each value a DELTA against the fp32 mean of its 4x8 half block;
                         followed (same allocation, at fp16 element Mp*nblk*64) by those means, f32
                         [Mp][nblk][2].  Mp*nblk*136 bytes.  Read back by craft_corr_lookup(lvl0_h16). */
} craft_scores_args;
/* TransCorrBlock.update core/corr.py:148-207 (+ CorrBlock.__init__ :16-45 with M=1).         */
int craft_corr_build(const craft_scores_args* a, void* stream);
/* softmax statistics of CrossAttFeatTrans.forward core/setrans.py:514-553 / gma.Attention.   */
int craft_attn_lse(const craft_scores_args* a, void* stream);
/* grid heuristics (host-side): key splits chosen so the grid fills whole waves of SMs         */
int craft_scores_auto_ksplit(int H, int W);
int craft_pv_auto_ksplit(int H, int W, int M);
/* keys per tile (8 x BK/8 spatial block) the P.V kernel uses for (d, F); V^T must be laid out in
 * that block order: column = block*BK + (y%8)*(BK/8) + x%(BK/8) (craft_shift_gemm b_blocked). */
int craft_pv_block_keys(int d, int F);
/* {sum,sumsq} -> {mean,rstd} (F.layer_norm, core/corr.py:200-204); n = number of elements.
 * flag (optional device int): when set, stat_sum + 2 (the clamped re-pass) is used instead.   */
int craft_corr_stats_finalize(const double* stat_sum, const int* flag, double n, float* mean_rstd, void* stream);
/* clip = (max > attn_clip) ? attn_clip : +inf ; flag = hit (core/setrans.py:527-529).  diag
 * (optional, f32[2]) accumulates the module diagnostics {max_attn, clamp_count} (:524-529).   */
int craft_clip_gate(const float* stat_max, float attn_clip, float* clip, int* flag, float* diag, void* stream);

/* ---- flash P.V ---------------------------------------------------------------------------- */
typedef struct craft_pv_args {
  const void* Q;      /* bf16 [Mp, C]                                                          */
  const void* K;      /* bf16 [Mp, C]                                                          */
  const void* Vt;     /* bf16 [M*F, ldv]  (V transposed, keys in 8 x BK/8 block order)         */
  int ldv;
  int C, M, d, F;
  int H, W;
  float scale, w_pos;
  const float* pos_table;
  int R;
  const float* clip;
  const float* lse2;  /* [M][Mp]                                                               */
  float* out;         /* f32 [ksplit][M][F/8][Mp][8]: partial sums, one slot per CTA sharing a unit */
  int ksplit;         /* slots available in out (>= craft_pv_auto_ksplit)                        */
  int zero_fill;      /* != 0: slots a (query tile, mode) unit does not use are written as zeros */
  int mask_radius;    /* as in craft_scores_args                                                  */
} craft_pv_args;
/* ExpandedFeatTrans.forward core/setrans.py:373-383, gma.Aggregate.forward core/gma.py:131-134 */
int craft_attn_pv(const craft_pv_args* a, void* stream);

/* mode soft-pooling + input skip + LayerNorm (core/setrans.py:395-407, :289-300);
 * gma != 0: y = x + gamma*O (core/gma.py:140).  O: [nsum][M][F/8][Mp][8] partials are summed.  */
int craft_modes_finalize(const float* O, int nsum, int M, int F, const float* w_score,
                         const float* b_score, const float* coeff, int gma, const void* x_bf16,
                         int ldx, int colx, const float* x_f32, int ldxf, int colxf, int H, int W,
                         void* out_bf16, int ldb, int colb, float* out_f32, int ldf, int colf,
                         int pv_bk /* 0: all nsum slots valid; 64/128: O from craft_attn_pv with that key block */,
                         void* stream);

/* ---- standalone forms of operations that CRAFT.forward runs fused -------------------------- */
/* LearnedSoftAggregate.forward core/setrans.py:289-300 on a dense f32 tensor whose group (mode) axis
 * leads: F == 1: x,basis [M][n] -> out [n] (p_m = softmax_m(w*basis_m + b));  F > 1: x,basis
 * [M][n][F] -> out [n][F] (p_m = softmax_m(<w, basis_m[r,:]> + b)).  M <= 8; basis may alias x.  */
int craft_soft_aggregate(const float* x, const float* basis, int M, long long n, int F, const float* w,
                         const float* b, float* out, void* stream);
/* Dense [M][U][U] attention matrix over the real tokens from projected rows -- the tensor the
 * production path never forms; for standalone CrossAttFeatTrans.forward on small grids and tests.
 * lse2 == NULL: scores after clamp + bias + mask (core/setrans.py:514-542); else softmax
 * probabilities exp(S - lse) (core/setrans.py:553).                                              */
typedef struct craft_dense_attn_args {
  const void* Q;
  const void* K;      /* bf16 [Mp, C]                                                           */
  int C, M, d, H, W;
  float scale, w_pos;
  const float* pos_table;
  int R;
  const float* clip;
  const float* lse2;  /* [M][Mp] or NULL                                                        */
  int mask_radius;
  float* out;         /* f32 [M][U][U]                                                          */
} craft_dense_attn_args;
int craft_attn_dense(const craft_dense_attn_args* a, void* stream);

/* ---- correlation lookup (CorrBlock.__call__ core/corr.py:47-71) --------------------------- */
/* lvl0_h16 (optional): level 0 as written by craft_corr_build(lvl0_h16) -- used instead of lvl[0].   */
int craft_corr_lookup(const float* const* lvl /*host array of 4 device ptrs*/, const void* lvl0_h16, int H, int W,
                      const float* coords /*[Mp,2]*/, const float* mean_rstd, void* out_bf16,
                      int ldb, float* out_nchw, int first_level, void* stream);

/* Level 0 of the same lookup computed on demand from the projected query/key rows [Mp,256] (the
 * U x U level-0 volume is then never stored): channels 0..80 of the output.  Pair it with
 * craft_corr_lookup(first_level = 1).                                                             */
int craft_corr_lookup0(const void* Q, const void* K, int M, int d, float scale, float w_agg, float w_pos,
                       const float* pos_table, int R, const float* clip, int H, int W, const float* coords,
                       const float* mean_rstd, void* out_bf16, int ldb, float* out_nchw, void* stream);

/* ---- small HBM-bound pieces ---------------------------------------------------------------- */
/* BasicMotionEncoder.convf1 (7x7, 2->128) + ReLU, core/update.py:75,82. wt [98][128], bias[128] */
int craft_convf1(const float* flow, const float* wt, const float* bias, int H, int W, void* out_bf16,
                 int ldo, int colo, void* stream);
/* coords1 += delta; flow = coords1 - coords0 (core/network.py:236,247). delta may be NULL.     */
int craft_flow_update(float* coords1, float* flow, const float* delta, int ldd, int H, int W,
                      void* stream);
/* coords1 = grid (+ flow_init NCHW) (core/network.py:142-149,221-222)                          */
int craft_init_coords(float* coords1, const float* flow_init_nchw, int H, int W, void* stream);
/* convex 8x upsampling (CRAFT.upsample_flow core/network.py:151-162). mask rows [Mp, ldm].      */
int craft_upsample_flow(const void* mask, int mask_is_bf16, int ldm, const float* flow, int H, int W,
                        float* out_nchw, void* stream);

/* ---- host I/O around the path ---------------------------------------- */
/* forward_interpolate core/utils/utils.py:34-62 (warm start, evaluate.py:146-147): out[g] = flow of the
 * nearest point among {p + flow(p)} that landed inside (0,W)x(0,H).  flow/out: [2,H,W] f32 planes.       */
int craft_forward_interpolate(const float* flow, int H, int W, float* out, void* stream);
/* byte re-packing for the writers of core/utils/frame_utils.py: mode 0 = .flo payload, interleaved f32
 * [H][W][2] (writeFlow :70-99); mode 1 = KITTI png pixels, u16 [H][W][3] = (1, 64v+2^15, 64u+2^15) in the
 * B,G,R order cv2.imwrite takes (writeFlowKITTI :116-120).                                              */
int craft_flow_encode(const float* flow, int H, int W, int mode, void* out, void* stream);

/* ---- encoder glue (core/extractor.py; first row outside the named hot path) ------------------- */
/* InstanceNorm2d statistics of a channels-last tensor [N,HW,C] (f32, or f16 when is_half) ->
 * ab[N,C,2] = (rstd, -mean*rstd) (eps, biased variance, no affine: extractor.py:136-137).
 * part: f32 scratch for per-block partial sums (part_capacity floats, >= 1024*N*C*2 is always
 * enough); the partials are reduced in a fixed order, so the result is run-to-run deterministic. */
int craft_nhwc_instnorm_stats(const void* x, int is_half, int N, int HW, int C, float eps, float* part,
                              long long part_capacity, float* ab, void* stream);
/* InstanceNorm2d + ReLU (+ residual) of a channels-last tensor in ONE cooperative launch (extractor.py:55-64 with
 * norm_fn='instance'):  out = relu_out( [ra*res+rb | res] + relu_in( (x-mean)*rstd ) ).  Every CTA keeps its slab
 * of x in shared memory across a grid barrier, so x is read from memory once (craft_nhwc_instnorm_stats +
 * craft_nhwc_affine read it twice, in three launches).  x/res/out: [N,HW,C] f32 or f16 (is_half); rab: f32
 * [N or 1][C][2] or NULL; part: f32 scratch, >= 64 + 2*C floats per SM, ZERO before its first use (the first 64
 * floats are barrier state the kernel leaves reusable); ab_out: optional [N,C,2] (rstd, -mean*rstd).
 * N <= number of SMs, C <= 256 (f16) / 128 (f32).  Deterministic (fixed reduction order).                        */
int craft_nhwc_instnorm_apply(const void* x, int is_half, int N, int HW, int C, float eps, const void* res,
                              const float* rab, int rab_nstride, int relu_in, int relu_out, float* part,
                              long long part_capacity, float* ab_out, void* out, void* stream);
/* 3x3 stride-1 convolution, 64 -> 64 channels, fp16, of the encoders' first residual layer (extractor.py:24-26,
 * 142-146) as a persistent tcgen05 implicit GEMM.  x / out: "padded-flat" activations [N*(H+1)*(W+2)][64] f16 (row
 * (n*(H+1)+y)*(W+2)+x; the cells x >= W and the row y = H of every image hold zeros: they are the padding); w: the
 * weights as [9*64][64] f16, tap-major (ky,kx row-major) blocks of [cout][cin]; bias: f32 [64] or NULL, relu: the
 * folded eval BatchNorm + ReLU of the context encoder.  With part/ab non-NULL the kernel also accumulates the
 * InstanceNorm2d statistics of its fp32 output over the valid cells and ab[N][64][2] = (rstd, -mean*rstd) is
 * written (part: f32 scratch, >= 128 floats per SM; deterministic).                                              */
int craft_conv3x3_c64(const void* x, const void* w, const float* bias, int relu, int N, int H, int W, void* out,
                      float* part, long long part_capacity, float* ab, float eps, void* stream);
/* craft_nhwc_affine between the dense channels-last layout [N][H][W][C] and the padded-flat one of
 * craft_conv3x3_c64 (*_pad != 0), in any combination; halo cells / gap rows of a padded output are set to zero. */
int craft_nhwc_affine_pad(const void* v, int is_half, int v_pad, const float* ab, int ab_nstride, const void* res,
                          int res_pad, const float* rab, int rab_nstride, int relu_in, int relu_out, int N, int H,
                          int W, int C, void* out, int out_pad, void* stream);
/* Input transform of both encoders: f32 frames [N,3,H,W] in 0..255 -> 2*(x/255)-1 (core/network.py:170-171), 2x2
 * space-to-depth, channels-last, zero border (2 cells before, 1 after): out [N][H/2+3][W/2+3][16] (f16 if
 * out_is_half else f32), channel (py*2+px)*3 + c, channels 12..15 zero.  The 7x7 stride-2 convolution
 * core/extractor.py:129 becomes a 4x4 stride-1 convolution over it with no padding.                       */
int craft_image_s2d(const float* img, int N, int H, int W, void* out, int out_is_half, void* stream);
/* out = relu_out( [ra*res+rb] + relu_in(a*v+b) ): norm + ReLU + residual of ResidualBlock.forward
 * (extractor.py:55-64).  v/res/out: f32 or f16 (is_half); ab / rab: f32 [N or 1][C][2];
 * *_nstride = 2*C per image or 0 when shared.                                                       */
int craft_nhwc_affine(const void* v, int is_half, const float* ab, int ab_nstride, const void* res,
                      const float* rab, int rab_nstride, int relu_in, int relu_out, int N, int HW, int C,
                      void* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRAFT_B200_H_ */
