/*
 * relpose_b200 -- C ABI of the B200 (sm_100a) compute library behind the rel_pose hot path.
 *
 * The reference (crockwell/rel_pose @35d1352) is pure Python: its "operator interface" for this
 * path is the nn.Module `ViTEss` (src/model.py:11-191) plus the third-party `lietorch.SE3`
 * type (environment.yml:20).  There is no FFI in the reference; this header is the boundary a
 * maintainer would bind with ctypes.  Each entry point cites the reference
 * lines whose arithmetic it replaces.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*;
 *   - tensors are dense row-major float32 unless stated otherwise, 16-byte aligned;
 *   - `device` is the CUDA ordinal that owns the buffers, `stream` a cudaStream_t on it
 *     (passed as void*); the call enqueues work and returns, it never synchronises, allocates
 *     or frees;
 *   - the caller owns inputs, outputs and workspaces and keeps them alive until the stream has
 *     passed the call;
 *   - return 0 on success; a negative RP_E* code for bad arguments; a positive cudaError_t for
 *     CUDA failures.  rp_last_error() returns a thread-local message.  Nothing throws.
 *   - stateless and re-entrant (autograd calls the *_bwd entry points from its own threads).
 */
#ifndef RELPOSE_B200_H
#define RELPOSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RP_OK 0
#define RP_EINVAL (-1)    /* bad shape / null pointer / unsupported size */
#define RP_EALIGN (-2)    /* pointer not 16-byte aligned */
#define RP_EWORKSPACE (-3) /* workspace too small */

/* epilogue activations of rp_linear_f32 */
#define RP_ACT_NONE 0
#define RP_ACT_GELU 1     /* exact erf GELU, vit_layers/mlp.py:22 */
#define RP_ACT_RELU 2     /* src/model.py:93,95 */

/* fixed geometry of the path (src/model.py:19-23, vision_transformer.py:409-424) */
#define RP_GRID 24
#define RP_NTOK 576
#define RP_EMBED 192
#define RP_HEADS 3
#define RP_HDIM 64
#define RP_NPOS 6
#define RP_EMW 70         /* HDIM + NPOS */

const char* rp_last_error(void);
int rp_version(void);
/* Compute capability major*10+minor of `device` (100 on B200), or a negative code. */
int rp_device_arch(int device);

/* ---- input pipeline step before the path -----------------------------------
 * Copies from PINNED host memory only the rows that A1's nearest resize (src row = floor(d*H/out_rows),
 * model.py:125) reads: src [planes][H][row_bytes] -> dst (device) [planes][out_rows][row_bytes], as
 * out_rows/gcd strided 2-D DMAs on `stream`.  The compact tensor is a valid input of the rp_preprocess_*
 * kernels (rows already selected, columns untouched): same pixels, 224/H of the PCIe traffic. */
int rp_copy_rows_h2d(void* dst, const void* src_pinned, int64_t planes, int H, int64_t row_bytes, int out_rows,
                     int device, void* stream);

/* ---- A1  src/model.py:114-125 -------------------------------------------------------------
 * images [n_img,3,H,W] BGR 0..255  ->  out [n_img,3,224,224] RGB, (x/255-mean)/std, legacy
 * nearest resize (src = floor(dst*in/out)).  Bit-exact w.r.t. the reference's float32 ops. */
int rp_preprocess_f32(const float* images, float* out, int n_img, int H, int W, int device, void* stream);
/* same from uint8 pixels (cv2.imread layout converted to NCHW by the caller) */
int rp_preprocess_u8(const uint8_t* images, float* out, int n_img, int H, int W, int device, void* stream);

/* ---- src/model.py:100-109 + vision_transformer.py:117-145 -----------------------------------
 * In place: fx,cx *= 24/W ; fy,cy *= 24/H on intrinsics [B,2,4]; writes kxy [B,2] =
 * (1/(fx/cx), 1/(fy/cy)) of view 0 (the diagonal of the reference's K^-1) and sets
 * flags[0] |= 1 if any pair has intrinsics[:,0] != intrinsics[:,1], |= 2 if cx*cy == 0 for
 * pair 0 (the two conditions the reference asserts on).  flags must be zeroed by the caller. */
int rp_intrinsics_prepare_f32(float* intrinsics, float* kxy, int* flags, int B, int H, int W, int device, void* stream);


/* ---- A2/A3 CNN front end  src/model.py:127-134, src/modules/extractor.py:51-65 ----------------
 * Activations are NHWC float32.  A convolution is an implicit GEMM (rows = output pixels,
 * K = KH*KW*C with the channel innermost) whose epilogue applies the folded eval-mode BatchNorm:
 *   y = act(conv(x,w)*scale[o] + shift[o] + res_pre) + res_post[row % res_post_rows]
 * w is [O][KH][KW][C] (rp_permute_conv_weight_f32), C % 4 == 0; scale/shift/res_* may be NULL;
 * res_post_rows = 0 means "same rows as y" (576 adds the [576,192] pos_embed to every image, A4).
 * With O = 192 and a 24x24 output, y IS the token matrix [n_img,576,192] of src/model.py:136-141. */
size_t rp_conv2d_workspace_bytes(int n_img, int H, int W, int C, int O, int KH, int KW, int stride, int pad);
int rp_conv2d_nhwc_f32(const float* x, const float* w, const float* scale, const float* shift,
                       const float* res_pre, const float* res_post, int res_post_rows, float* y,
                       int n_img, int H, int W, int C, int O, int KH, int KW, int stride, int pad, int act,
                       void* workspace, size_t workspace_bytes, int device, void* stream);
/* A1 with NHWC output, channels padded 3 -> 4 (4th = 0): [n_img,3,H,W] BGR -> [n_img,224,224,4] */
int rp_preprocess_nhwc4_f32(const float* images, float* out, int n_img, int H, int W, int device, void* stream);
int rp_preprocess_nhwc4_u8(const uint8_t* images, float* out, int n_img, int H, int W, int device, void* stream);
/* nn.MaxPool2d(3,2,1) on NHWC (torchvision resnet stem), C % 4 == 0; output [(H+1)/2... ] = floor((H-1)/2)+1 */
int rp_maxpool3x3s2_nhwc_f32(const float* x, float* y, int n_img, int H, int W, int C, int device, void* stream);
/* parameter preparation (once per weight version): [O][C][KH][KW] -> [O][KH][KW][Cp], zero padded */
int rp_permute_conv_weight_f32(const float* w, float* out, int O, int C, int KH, int KW, int Cp, int device, void* stream);
/* scale = gamma/sqrt(var+eps), shift = (conv_bias - mean)*scale + beta   (conv_bias may be NULL) */
int rp_bn_fold_f32(const float* gamma, const float* beta, const float* mean, const float* var,
                   const float* conv_bias, float eps, float* scale, float* shift, int C, int device, void* stream);

/* ---- A4  src/model.py:136-141,172 ---------------------------------------------------------
 * fmap [n_img,192,576] (NCHW feature map, 24x24 flattened) -> x [n_img,576,192] + pos_embed[576,192] */
int rp_tokens_posembed_f32(const float* fmap, const float* pos_embed, float* x, int n_img, int device, void* stream);

/* ---- LayerNorm  vision_transformer.py:396 (eps 1e-6), rows x cols, cols <= 1024 ------------ */
int rp_layernorm_f32(const float* x, const float* gamma, const float* beta, float* y, int rows, int cols,
                     float eps, int device, void* stream);

/* ---- nn.Linear with fused epilogue  (vision_transformer.py:323,331; mlp.py:21-24; model.py:91-98)
 * C[M,N] = act(A[M,K] * W[N,K]^T + bias[N]) + residual[M,N]    (bias, residual may be NULL;
 * residual may alias C).  K % 4 == 0.  workspace: rp_linear_workspace_bytes(M,N,K) bytes (may be 0). */
size_t rp_linear_workspace_bytes(int M, int N, int K);
int rp_linear_f32(const float* A, const float* W, const float* bias, const float* residual, float* C,
                  int M, int N, int K, int act, void* workspace, size_t workspace_bytes, int device, void* stream);


/* ---- tensor-core (tcgen05 + TMA + TMEM) nn.Linear on split-bf16 operands ----------------------
 * A float32 tensor is carried as P bf16 "planes" laid out [P][rows][K]: plane 0 = bf16(x), plane 1 =
 * bf16(x - plane0).  P = 1 is plain bf16; P = 2 evaluates a0*b0 + a0*b1 + a1*b0 with fp32
 * accumulation in tensor memory ("bf16x3", fp32-class products: holds the 1e-4 parity bar).
 *   out = act(A W^T + bias) + residual, written as float32 [M,N] (out_f32) and/or as P_out planes
 *   [P_out][M][N] (out_planes) ready to be the A operand of the next GEMM.   K % 8 == 0. */
int rp_split_planes_bf16(const float* x, void* planes, int64_t n, int P, int device, void* stream);
int rp_layernorm_planes_bf16(const float* x, const float* gamma, const float* beta, void* planes, int rows, int cols,
                             float eps, int P, int device, void* stream);
int rp_linear_tc(const void* A_planes, const void* W_planes, const float* bias, const float* residual,
                 float* out_f32, void* out_planes, int M, int N, int K, int P, int P_out, int act, int device,
                 void* stream);
/* pose_regressor[2:5] (src/model.py:93-97) in one launch: out [B,14] = W2 relu(W1 h + b1) + b2 for h [B,512]
 * (the ReLU'ed output of layer 0).  W1T is pose_regressor.2.weight TRANSPOSED ([in][out], contiguous); W2 is
 * pose_regressor.4.weight as stored ([14][512]). */
int rp_regressor_tail_f32(const float* h, const float* W1T, const float* b1, const float* W2, const float* b2, float* out,
                          int B, int hidden, int n_out, int device, void* stream);

/* Split-K variant of rp_linear_tc for skinny, weight-bandwidth-bound layers (pose_regressor.0: K = 26 880,
 * src/model.py:91-98,189): out_f32 = act(A W^T + bias), float32 partials of K / ksplit blocks in `workspace`
 * (rp_linear_tc_splitk_workspace_bytes; ksplit_out may be NULL), added in a fixed order.  N % 4 == 0, K % 8 == 0. */
size_t rp_linear_tc_splitk_workspace_bytes(int M, int N, int K, int* ksplit_out);
int rp_linear_tc_splitk(const void* A_planes, const void* W_planes, const float* bias, float* out_f32, int M, int N, int K,
                        int P, int act, void* workspace, size_t workspace_bytes, int device, void* stream);

/* 3x3 / stride 1 / pad 1 convolution with 64 input channels (ResNet layer1, src/model.py:127-131), "halo" variant of
 * rp_conv2d_tc: one TMA box per tile brings the activation halo, the nine filter taps read it through shifted
 * shared-memory descriptors (2x less L2 -> SM traffic on layers that are bound by it).  Same operands and epilogue
 * as rp_conv2d_tc (no res_post).  rp_conv3x3_halo_supported tells whether a shape qualifies. */
int rp_conv3x3_halo_supported(int H, int W, int C, int O, int KH, int KW, int stride, int pad);
int rp_conv3x3_halo_tc(const void* x_planes, const void* w_planes, const float* scale, const float* shift,
                       const float* res_pre, float* out_f32, void* out_planes, int n_img, int H, int W, int C, int O,
                       int P, int P_out, int act, int device, void* stream);

/* LayerNorm fused into a projection  out = LayerNorm(x) W^T + bias  in one launch: `self.qkv(self.norm1(x))` of
 * Block / Attention.forward (vision_transformer.py:350,323) and of CrossBlock / CrossAttention.forward
 * (vision_transformer.py:288-289,191-194).  x float32 [M,K]; W_planes bf16 [P][N][K]; outputs as in rp_linear_tc
 * (float32 [M,N] and/or P_out bf16 planes [P_out][M][N]).  Built for K = 192; N % 4 == 0 for the vector path. */
int rp_ln_linear_tc(const float* x, const float* ln_gamma, const float* ln_beta, float eps, const void* W_planes,
                    const float* bias, float* out_f32, void* out_planes, int M, int N, int K, int P, int P_out,
                    int device, void* stream);
/* Generalised form of rp_ln_linear_tc for a chain of Blocks in which every producer of the residual stream also
 * emits the NEXT LayerNorm's output as bf16 planes, so that no consumer computes a LayerNorm on its critical path:
 *   xn_planes != NULL     the A operand is supplied as bf16 planes [P][M][192] (x / ln_gamma / ln_beta unused), loaded
 *                         by TMA: `self.qkv(...)` of vision_transformer.py:323 on the planes of norm1(x);
 *   residual  != NULL     (N = 192) out_f32 = A W^T + bias + residual: `x = x + self.attn.proj(...)`
 *                         (vision_transformer.py:331,351); out_planes must be NULL;
 *   out_ln_planes != NULL (with residual) additionally LayerNorm(out_f32; ln2_gamma, ln2_beta, eps2) as bf16 planes
 *                         [P][M][192]: the Block's norm2 (vision_transformer.py:352), consumed by rp_mlp_tc_ex. */
int rp_ln_linear_tc_ex(const float* x, const void* xn_planes, const float* ln_gamma, const float* ln_beta, float eps,
                       const void* W_planes, const float* bias, const float* residual, float* out_f32, void* out_planes,
                       void* out_ln_planes, const float* ln2_gamma, const float* ln2_beta, float eps2, int M, int N, int K,
                       int P, int P_out, int device, void* stream);

/* Fused MLP half-block  out = x + fc2(GELU(fc1(LayerNorm(x))))  in one launch: Block.forward's
 * `x = x + self.mlp(self.norm2(x))` (vision_transformer.py:352-353; mlp.py:20-26) and CrossBlock.forward's
 * `out = f + mlp(norm2(f))` (vision_transformer.py:295-296).  x, out float32 [M,dim]; W1_planes bf16
 * [P][hidden][dim], W2_planes bf16 [P][dim][hidden] (as written by rp_split_planes_bf16 from the nn.Linear
 * weights); LayerNorm eps as in vision_transformer.py:396.  The [M,hidden] activation never reaches HBM.
 * Built for dim = 192, hidden = 768 (the only widths the reference instantiates).  out may alias x. */
int rp_mlp_tc(const float* x, const float* ln_gamma, const float* ln_beta, float eps, const void* W1_planes,
              const float* b1, const void* W2_planes, const float* b2, float* out, int M, int dim, int hidden,
              int P, int device, void* stream);
/* Same with the LayerNorms moved to the producers (see rp_ln_linear_tc_ex): xn_planes != NULL supplies norm2(x) as
 * bf16 planes [P][M][192] (fc1's A operand by TMA; ln_gamma / ln_beta unused); out_ln_planes != NULL additionally
 * receives LayerNorm(out; ln2_gamma, ln2_beta, eps2) -- norm1 of the NEXT Block (vision_transformer.py:350) -- as bf16
 * planes [P][M][192]. */
int rp_mlp_tc_ex(const float* x, const void* xn_planes, const float* ln_gamma, const float* ln_beta, float eps,
                 const void* W1_planes, const float* b1, const void* W2_planes, const float* b2, float* out,
                 void* out_ln_planes, const float* ln2_gamma, const float* ln2_beta, float eps2, int M, int dim, int hidden,
                 int P, int device, void* stream);

/* Tensor-core convolution (A2/A3), same epilogue contract as rp_conv2d_nhwc_f32.  x_planes is the NHWC
 * activation as bf16 planes [P][n_img][H][W][C] (C % 64 == 0), w_planes [P][O][KH*KW*C] is the split of
 * the [O][KH][KW][C] weight, O in {64,128,192}, stride 1 or 2.  Every (tap, 64-channel block) K step is
 * one 5-D TMA box; padding is TMA zero fill.  Outputs: float32 [n,Ho,Wo,O] and/or P_out planes. */
int rp_conv2d_tc(const void* x_planes, const void* w_planes, const float* scale, const float* shift,
                 const float* res_pre, const float* res_post, int res_post_rows, float* out_f32, void* out_planes,
                 int n_img, int H, int W, int C, int O, int KH, int KW, int stride, int pad, int P, int P_out, int act,
                 int device, void* stream);
/* Stem on tensor cores: resnet.conv1 (7x7/2, pad 3; src/model.py:127) as a 4x4 stride-1 convolution over
 * the 2x2 space-to-depth image.  rp_preprocess_stem_windows_* is A1 (src/model.py:114-125: BGR->RGB, /255,
 * mean/std, legacy-nearest resize to 224) fused with that layout: it writes bf16 planes
 * [P][n_img][115][112][64] where [yp][ox][b*16 + (dy*2+dx)*3 + c] = pixel (2(yp-2)+dy, 2(ox-2+b)+dx) channel c
 * of the normalised 224x224 image (0 outside, channels 12..15 of each group 0).  rp_stem_weight_windows_f32
 * re-lays conv1.weight [O][3][7][7] as [O][4][64] to match; the convolution itself is
 * rp_conv2d_tc(H=115, W=112, C=64, KH=4, KW=1, stride 1, pad 0). */
int rp_preprocess_stem_windows_f32(const float* images, void* planes, int n_img, int H, int W, int P, int device,
                                   void* stream);
int rp_preprocess_stem_windows_u8(const uint8_t* images, void* planes, int n_img, int H, int W, int P, int device,
                                  void* stream);
int rp_stem_weight_windows_f32(const float* w, float* out, int O, int device, void* stream);
/* The whole stem (src/model.py:127-130: conv1 -> bn1 -> relu -> maxpool 3x3/2/1) in one launch: tcgen05 convolution whose
 * epilogue keeps the vertical 3-row maximum in registers and finishes the pooling through shared memory; only the pooled
 * [n,56,56,64] map is written (float32 and/or P_out bf16 planes).  z_planes is either the window tensor of
 * rp_preprocess_stem_windows_* (compact = 0) or the COMPACT space-to-depth image of rp_preprocess_stem_compact_*
 * (compact = 1): bf16 planes [P][n_img][115][116][16], [yp][xs][(dy*2+dx)*3 + c] = pixel (2(yp-2)+dy, 2(xs-2)+dx) channel c
 * of the normalised 224x224 image -- 4x smaller; the convolution reads its overlapping 64-element windows through a
 * tensor map whose window stride (32 B) is smaller than the window (128 B).  rp_stem_compact_supported: 1 if the driver
 * encodes that map.  w_planes [P][64][256] (rp_stem_weight_windows_f32 + rp_split_planes_bf16), scale/shift = folded bn1.
 * Bit-identical to rp_conv2d_tc + rp_maxpool3x3s2_planes on the window tensor. */
int rp_preprocess_stem_compact_f32(const float* images, void* planes, int n_img, int H, int W, int P, int device, void* stream);
int rp_preprocess_stem_compact_u8(const uint8_t* images, void* planes, int n_img, int H, int W, int P, int device, void* stream);
int rp_stem_compact_supported(int device);
int rp_stem_pool_tc(const void* z_planes, int compact, const void* w_planes, const float* scale, const float* shift,
                    float* out_f32, void* out_planes, int n_img, int P, int P_out, int device, void* stream);
/* nn.MaxPool2d(3,2,1) on NHWC float32 writing float32 (y_f32, may be NULL) and/or P bf16 planes */
int rp_maxpool3x3s2_planes(const float* x, float* y_f32, void* y_planes, int P, int n_img, int H, int W, int C,
                           int device, void* stream);

/* ---- A5 attention core  vision_transformer.py:323-329 --------------------------------------
 * qkv [n_img,576,576] (column = s*192+h*64+d)  ->  out [n_img,576,192] (column = h*64+d),
 * out = softmax(q k^T * 0.125) v per image and head; nothing is materialised in HBM. */
int rp_self_attention_f32(const float* qkv, float* out, int n_img, int device, void* stream);
/* Same contract on tcgen05 tensor cores (flash style: QK^T -> online softmax -> PV inside one kernel, S and
 * O accumulators in tensor memory, operands by TMA).  qkv_planes = bf16 planes [P][n_img][576][576] as
 * written by rp_linear_tc (P = 1 bf16, P = 2 split bf16 = fp32-class); outputs float32 [n_img,576,192]
 * (out_f32, may be NULL) and/or P_out bf16 planes [P_out][n_img][576][192] (A operand of attn.proj). */
int rp_self_attention_tc(const void* qkv_planes, float* out_f32, void* out_planes, int n_img, int P, int P_out,
                         int device, void* stream);
/* --noess ablation, CrossAttention.forward vision_transformer.py:239-253: plain cross attention
 * between the two views of a pair.  Same layouts as above; image n's queries attend to the keys / values of image
 * n^1 (n_img must be even): out[2b] = softmax(q1 k2^T * 0.125) v2, out[2b+1] = softmax(q2 k1^T * 0.125) v1, i.e.
 * already in the flipped order the reference returns (:262). */
int rp_cross_attention_f32(const float* qkv, float* out, int n_img, int device, void* stream);
int rp_cross_attention_tc(const void* qkv_planes, float* out_f32, void* out_planes, int n_img, int P, int P_out,
                          int device, void* stream);

/* ---- A6  vision_transformer.py:90-158 ------------------------------------------------------
 * pos [B,576,6] = [p3^2,p4^2,p3*p4,p3,p4,1], p3 = ys[i%24]*ky, p4 = xs[i/24]*kx (transposed grid).
 * lin24: host pointer to the 24 floats of torch.linspace(-1,1,24); kxy [B,2] from
 * rp_intrinsics_prepare_f32 or NULL (intrinsics=None: kx = ky = 1 without multiplication). */
int rp_posenc_f32(const float* kxy, const float* host_lin24, float* pos, int B, int device, void* stream);
/* l1 != 0: --l1_pos_encoding, get_l1_positional_encodings (vision_transformer.py:36-87): channels [1,1,1,p3,p4,1] */
int rp_posenc_ex_f32(const float* kxy, const float* host_lin24, float* pos, int B, int l1, int device, void* stream);

/* ---- A7 Essential Matrix Module core  vision_transformer.py:198-223 -------------------------
 * qkv [2B,576,576] with the two views of pair b at rows 2b, 2b+1; pos [B,576,6] or NULL
 * (no_pos_encoding); out bil [B,2,3,W,W], W = 70 (64 if pos == NULL):
 *   bil[b,0,h] = V1^T A1 V1,  A1 = softmax(S1,-1)*softmax(S1,-2),  S1 = q2 k1^T * 0.125
 *   bil[b,1,h] = V2^T A2 V2,  A2 likewise from S2 = q1 k2^T * 0.125,   V = [v | pos].
 * The 576x576 affinities never leave the chip.  workspace: rp_essential_workspace_bytes(B). */
size_t rp_essential_workspace_bytes(int B);
int rp_essential_f32(const float* qkv, const float* pos, float* bil, int B, void* workspace,
                     size_t workspace_bytes, int device, void* stream);
/* Ablation branches of the same module, fp32 SIMT kernels: flags = RP_EM_SINGLE_SOFTMAX
 * (--use_single_softmax, vision_transformer.py:201-203: A = softmax(S,-1) only) | RP_EM_CROSS_FEATURES
 * (--cross_features, :219-220: F1 = V2^T A1 V1, F2 = V1^T A2 V2). */
#define RP_EM_SINGLE_SOFTMAX 1
#define RP_EM_CROSS_FEATURES 2
int rp_essential_ex_f32(const float* qkv, const float* pos, float* bil, int B, int flags, void* workspace,
                        size_t workspace_bytes, int device, void* stream);
/* Same contract on tcgen05 tensor cores.  qkv_planes = bf16 planes [P][2B][576][576] of the cross block's QKV GEMM
 * (P = 1 bf16, P = 2 split bf16 = fp32 class).  Two kernels: row/column log-sum-exp of S, then the fused
 * dual-softmax -> A [v|pos] -> [v|pos]^T T accumulation with F resident in tensor memory (no partials, no
 * atomics, bit-reproducible).  workspace: rp_essential_tc_workspace_bytes(B, P). */
size_t rp_essential_tc_workspace_bytes(int B, int P);
int rp_essential_tc(const void* qkv_planes, const float* pos, float* bil, int B, int P, void* workspace,
                    size_t workspace_bytes, int device, void* stream);
/* The same two kernels with the ablation flags of rp_essential_ex_f32 (vision_transformer.py:201-203,219-220):
 * RP_EM_SINGLE_SOFTMAX drops the column term of the exponent, RP_EM_CROSS_FEATURES takes the left factor of F from
 * the other view.  flags = 0 is rp_essential_tc. */
int rp_essential_ex_tc(const void* qkv_planes, const float* pos, float* bil, int B, int P, int flags, void* workspace,
                       size_t workspace_bytes, int device, void* stream);


/* ---- A7 tail  vision_transformer.py:229-238 + :292-294 --------------------------------------
 * Z[b,c,h*70+a] = bil[b,dir,h,a,c];  Y = Z W^T + bias (proj_fundamental, W [192,210]);
 * out [2B,70,192] with out[2b] = Y from bil[b,1] and out[2b+1] = Y from bil[b,0] (the flip). */
int rp_em_project_f32(const float* bil, const float* W, const float* bias, float* out, int B, int device, void* stream);

/* ---- A10  src/model.py:145-152 --------------------------------------------------------------
 * raw [B,2,7], Gs [B,2,7] -> out [B,2,7]: out[:,0] = Gs[:,0]; out[:,1] = raw[:,1] with the
 * quaternion divided by max(||q||, 0.01). */
int rp_normalize_pose_f32(const float* raw, const float* Gs, float* out, int B, int device, void* stream);

/* ---- A12 lietorch SE3 (lietorch==0.2; call sites src/geom/losses.py:8-10) --------------------
 * Elements are 7 floats [t, q_xyzw]; tangents 6 floats [tau, phi].  n = number of elements.
 * Quaternions are re-normalised on load, as lietorch's SO3 constructor does.
 * Backward = lietorch's convention: gradient of a LEFT perturbation exp(d)X, written into the
 * first 6 of 7 slots (slot 7 = 0). */
int rp_se3_mul_fwd_f32(const float* X, const float* Y, float* Z, int64_t n, int device, void* stream);
int rp_se3_mul_bwd_f32(const float* dZ, const float* X, const float* Y, float* dX, float* dY, int64_t n, int device, void* stream);
int rp_se3_inv_fwd_f32(const float* X, float* Y, int64_t n, int device, void* stream);
int rp_se3_inv_bwd_f32(const float* dY, const float* X, float* dX, int64_t n, int device, void* stream);
int rp_se3_log_fwd_f32(const float* X, float* a, int64_t n, int device, void* stream);
int rp_se3_log_bwd_f32(const float* da, const float* X, float* dX, int64_t n, int device, void* stream);
int rp_se3_exp_fwd_f32(const float* a, float* X, int64_t n, int device, void* stream);
int rp_se3_exp_bwd_f32(const float* dX, const float* a, float* da, int64_t n, int device, void* stream);

/* ---- training path (A11 / config 5): building blocks of the backward pass and of the train-mode forward ----------
 * fp32, deterministic (two-stage fixed-order reductions, no atomics).  Row-major everywhere.
 * rp_gemm_f32: strided-batched C = alpha op(A) op(B) + beta C with a two-level batch index
 *   (b = bo * batch_inner + bi; X_b = X + bo * sXo + bi * sXi): dX = dY W, dW = dY^T X, and the attention / Essential
 *   Matrix Module products on materialised 576x576 matrices (autograd of vision_transformer.py:198-223,325-329).
 * rp_im2col / rp_col2im: NHWC convolution gradients as GEMMs (column = (ky*KW+kx)*C + c).
 * rp_bn_*: nn.BatchNorm2d in training mode on [M][C] (batch statistics, biased variance, running-stat update with
 *   the unbiased variance; torchvision resnet18 / extractor.py:24-28), C % 4 == 0.  rp_bn_bwd: y_relu (may be NULL) = the
 *   block's output when a ReLU followed (dz = dy .* (y > 0) is formed on the fly); dz_out (may be NULL) receives dz, the
 *   gradient of the residual branch.   rp_layernorm_*: eps inside the sqrt.
 * rp_softmax_{rows,cols}: softmax(scale * S) over dim -1 / dim -2 of [mats][n][n]; *_bwd adds into ds if accumulate.
 * rp_maxpool3x3s2_bwd: gradient to the FIRST maximum of each window (PyTorch's tie-break). */
/* float32 [R][C] -> bf16 planes of the transpose [P][C][R] (R even): operands of the weight-gradient GEMMs
 * dW = dY^T X on the tensor-core engine (rp_linear_tc_splitk contracts over what were the rows). */
int rp_transpose_split_planes_bf16(const float* x, void* planes, int R, int C, int P, int device, void* stream);
int rp_gemm_f32(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, const float* B,              int ldb, float beta, float* C, int ldc, int batch_outer, int batch_inner, int64_t sAo, int64_t sAi,              int64_t sBo, int64_t sBi, int64_t sCo, int64_t sCi, int device, void* stream);
int rp_gelu_fwd_f32(const float* z, float* y, int64_t n, int device, void* stream);
int rp_gelu_bwd_f32(const float* dy, const float* z, float* dz, int64_t n, int device, void* stream);
int rp_relu_bwd_f32(const float* dy, const float* y, float* dx, int64_t n, int device, void* stream);
int rp_mul_f32(const float* a, const float* b, float* c, int64_t n, int device, void* stream);
int rp_axpby_f32(float alpha, const float* x, float beta, const float* y, float* out, int64_t n, int device, void* stream);
int rp_add_bcast_rows_f32(const float* a, const float* b, float* out, int64_t rows, int cols, int period, int device,                   void* stream);
int rp_sum_over_period_f32(const float* a, float* out, int reps, int period, int cols, int device, void* stream);
size_t rp_colsum_workspace_bytes(int64_t rows, int cols);
int rp_colsum_f32(const float* A, const float* B, float* out, int64_t rows, int cols, void* workspace,               size_t workspace_bytes, int device, void* stream);
int rp_layernorm_train_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,                      int rows, int cols, float eps, int device, void* stream);
size_t rp_layernorm_bwd_workspace_bytes(int rows, int cols);
int rp_layernorm_bwd_f32(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,                   float* dx, float* dgamma, float* dbeta, int rows, int cols, void* workspace,                   size_t workspace_bytes, int device, void* stream);
int rp_softmax_rows_fwd_f32(const float* s, float* p, int64_t rows, int n, float scale, int device, void* stream);
int rp_softmax_rows_bwd_f32(const float* dp, const float* p, float* ds, int64_t rows, int n, float scale, int accumulate,                    int device, void* stream);
int rp_softmax_cols_fwd_f32(const float* s, float* p, int mats, int n, float scale, int device, void* stream);
int rp_softmax_cols_bwd_f32(const float* dp, const float* p, float* ds, int mats, int n, float scale, int accumulate,                    int device, void* stream);
size_t rp_bn_workspace_bytes(int64_t M, int C);
int rp_bn_train_stats_f32(const float* x, float* mean, float* var, float* running_mean, float* running_var,                   float momentum, int64_t M, int C, void* workspace, size_t workspace_bytes, int device,                   void* stream);
int rp_bn_apply_f32(const float* x, const float* mean, const float* var, const float* gamma, const float* beta,                const float* residual, float* y, int64_t M, int C, float eps, int relu, int device, void* stream);
int rp_bn_bwd_f32(const float* dy, const float* y_relu, const float* x, const float* mean, const float* var, const float* gamma,
                  float eps, float* dx, float* dz_out, float* dgamma, float* dbeta, int64_t M, int C, void* workspace,
                  size_t workspace_bytes, int device, void* stream);
int rp_im2col_nhwc_f32(const float* x, float* cols, int n, int H, int W, int C, int KH, int KW, int stride, int pad,                  int device, void* stream);
/* im2col written as the bf16 planes of cols^T, planes [P][KH*KW*C][n*Ho*Wo] (the B operand of the split-K weight-gradient
 * GEMM), for the convolutions rp_conv_dw_tc does not take (C % 64 != 0: the stem). */
int rp_im2col_t_planes_bf16(const float* x, void* planes, int P, int n, int H, int W, int C, int KH, int KW, int stride, int pad,
                            int device, void* stream);
int rp_col2im_nhwc_f32(const float* dcols, float* dx, int n, int H, int W, int C, int KH, int KW, int stride, int pad,                  int device, void* stream);
size_t rp_maxpool3x3s2_bwd_workspace_bytes(int n, int H, int W, int C);
int rp_maxpool3x3s2_bwd_f32(const float* dy, const float* x, float* dx, int n, int H, int W, int C, void* workspace,
                            size_t workspace_bytes, int device, void* stream);
int rp_normalize_pose_bwd_f32(const float* dout, const float* raw, float* draw, int B, int device, void* stream);
/* Flash-style attention for the training step (autograd of vision_transformer.py:321-329 without the 576x576 tensors):
 * rp_self_attention_tc_lse = rp_self_attention_tc that also writes lse [n_img][3][576], the log2-sum-exp of every
 *   scaled score row (P_ij = 2^(s_ij 0.125 log2 e - lse_i)).
 * rp_attention_bwd_prep: d_out, out float32 [n_img,576,192] -> bf16 planes of d_out [2][n_img][576][192] and
 *   delta [n_img][3][576] = <d_out, out> per row and head.
 * rp_attention_bwd_tc: qkv planes [2][n_img][576][576], d_out planes, lse, delta -> d_qkv float32 [n_img,576,576]
 *   (every element written once).  tcgen05: S, dP recomputed per tile; dv = P^T dO, dk = dS^T q, dq = dS k. */
int rp_self_attention_tc_lse(const void* qkv_planes, float* out_f32, float* lse, int n_img, int P, int device, void* stream);
int rp_attention_bwd_prep(const float* d_out, const float* out, void* d_out_planes, float* delta, int n_img, int device,
                          void* stream);
int rp_attention_bwd_tc(const void* qkv_planes, const void* d_out_planes, const float* lse, const float* delta,
                        float* d_qkv, int n_img, int device, void* stream);
/* Weight gradient of nn.Conv2d (stride 1 or 2) as an implicit GEMM on tcgen05 (no im2col matrix, no transposed copies):
 * x_planes bf16 [2][n][H][W][C], dy_planes bf16 [2][n][OH][OW][O] (OH = (H + 2 pad - KH) / stride + 1) -> dw float32 [O][KH][KW][C].
 * Both operands are read as MN-major tiles straight from the NHWC planes; the taps are shifted views of one halo
 * tile (stride 1; one box per tap for stride 2); per-pixel-slice partial sums are reduced in a fixed order (deterministic).  C % 64 == 0, O in {64,128,192}. */
int rp_conv_dw_tc_supported(int C, int O, int KH, int KW, int stride);
size_t rp_conv_dw_tc_workspace_bytes(int n_img, int H, int W, int C, int O, int KH, int KW, int pad, int stride, int device);
int rp_conv_dw_tc(const void* x_planes, const void* dy_planes, float* dw, int n_img, int H, int W, int C, int O, int KH,
                  int KW, int pad, int stride, void* workspace, size_t workspace_bytes, int device, void* stream);
/* Weight gradient of nn.Linear (autograd of vision_transformer.py:323,331, vit_layers/mlp.py:21,24), dW [N][K] = dY^T X,
 * through the same implicit-GEMM kernel: the M rows are the "pixels", the K / 64 column blocks of X the "taps" of a
 * 1 x (K/64) convolution over 64 channels, so x_planes bf16 [2][M][K] and dy_planes bf16 [2][M][N] are read in place as
 * MN-major operands (no transposed copies).  M % 8 == 0, M >= 64, K % 64 == 0, K <= 4096, N in {64,128,192} or a
 * multiple of 192 / 128.  Deterministic (fixed-order slice reduction). */
int rp_linear_dw_tc_supported(int M, int N, int K);
size_t rp_linear_dw_tc_workspace_bytes(int M, int N, int K, int device);
int rp_linear_dw_tc(const void* x_planes, const void* dy_planes, float* dw, int M, int N, int K, void* workspace,
                    size_t workspace_bytes, int device, void* stream);
/* Flash-style backward of the Essential Matrix Module core (autograd of vision_transformer.py:198-223 without the
 * 576x576 tensors).  Forward = rp_essential_tc, whose workspace begins with lse2 [B][2][2][3][576] (row / column
 * log2-sum-exp of the scaled scores): keep it.  d_bil [B,2,3,70,70] -> d_qkv float32 [2B,576,576] (every element written
 * once; the positional encodings receive no gradient).  Four launches: dT = V' dF, pass A (row / column sums of A .* dA,
 * T = A V', A^T dT), dv, pass B (dq = dS k, dk = dS^T q); S and A are recomputed per tile on tcgen05. */
size_t rp_em_bwd_tc_workspace_bytes(int B);
int rp_em_bwd_tc(const void* qkv_planes, const float* pos, const float* lse2, const float* d_bil, float* d_qkv, int B,
                 void* workspace, size_t workspace_bytes, int device, void* stream);
int rp_concat_vpos_f32(const float* qkv, const float* pos, float* vp, int n_img, int device, void* stream);
int rp_scatter_dv_f32(const float* dvp, float* dqkv, int n_img, int device, void* stream);

/* ---- BASELINE.json config 3 (no reference counterpart; oracle = LAPACK, parity unpinned) -----
 * E [n,3,3] -> U [n,3,3], S [n,3] (descending, >= 0), V [n,3,3] with E = U diag(S) V^T. */
int rp_svd3_f32(const float* E, float* U, float* S, float* V, int64_t n, int device, void* stream);
/* E [n,3,3] -> R1 = U W V^T, R2 = U W^T V^T (det +1), t = U[:,2]   ([n,3,3],[n,3,3],[n,3]) */
int rp_essential_to_rt_f32(const float* E, float* R1, float* R2, float* t, int64_t n, int device, void* stream);

/* ---- training step after the path: clip_grad_norm_ + Adam + this step's learning rate ----
 * train.py:161-165 (torch.nn.utils.clip_grad_norm_(params, 2.5); optimizer.step(); scheduler.step()) with
 * torch.optim.Adam(lr, weight_decay) of train.py:69, as two multi-tensor launches and no host synchronisation.
 * descs: device array of {float* param; const float* grad; float* exp_avg; float* exp_avg_sq; int64 numel}
 * (40 bytes each); blk_tensor / blk_off: for every thread block the tensor it works on and its first element
 * (chunks of `chunk` elements, chunk % 4 == 0).  rp_grad_norm_multi writes the global L2 norm of all gradients to
 * norm_out[0] (partial: nblk floats of scratch; fixed summation order).  rp_adam_clip_step_multi applies
 * clip_coef = min(1, max_norm / (norm + 1e-6)) (max_norm <= 0: no clipping), L2 weight decay, and the Adam update
 * of torch/optim/adam.py::_single_tensor_adam for step number `step` (>= 1) with learning rate lr. */
int rp_grad_norm_multi(const void* descs, const int* blk_tensor, const int64_t* blk_off, int nblk, int chunk,
                       float* partial, float* norm_out, int device, void* stream);
int rp_adam_clip_step_multi(const void* descs, const int* blk_tensor, const int64_t* blk_off, int nblk, int chunk,
                            const float* total_norm, double max_norm, double lr, double beta1, double beta2, double eps,
                            double weight_decay, int step, int device, void* stream);
/* rp_adam_clip_step_multi for CUDA-graph replays: the two scalars that change every step, step_hyper[0] = lr_t /
 * (1 - beta1^t) and step_hyper[1] = sqrt(1 - beta2^t) (float32, DEVICE memory), are read by the kernel instead of
 * being passed by value, so one captured launch serves every optimizer step (train.py:161-165 as a replayed graph). */
int rp_adam_clip_step_multi_dev(const void* descs, const int* blk_tensor, const int64_t* blk_off, int nblk, int chunk,
                                const float* total_norm, double max_norm, const float* step_hyper, double beta1,
                                double beta2, double eps, double weight_decay, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RELPOSE_B200_H */
