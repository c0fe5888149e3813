/* kmunet.h -- C ABI of libkmunet.so: the B200 (sm_100a) hot path of KM-UNet.
 *
 * Every entry point replaces the arithmetic of one reference PyTorch module method (cited per function as
 * <reference file>:<lines>, paths relative to the KM-UNet tree).  The host side that binds these symbols is
 * km_unet_b200/_lib.py (ctypes); INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - All tensors are dense fp32, NCHW / (B,C,L) contiguous exactly as the reference modules hold them.
 *   - Every pointer is a CUDA device pointer owned by the caller (the PyTorch caching allocator).  The library
 *     never allocates or frees device memory, keeps no per-tensor state, and never synchronises: each call
 *     only enqueues kernels on `stream` (a cudaStream_t) and is CUDA-graph capturable.
 *   - Scratch space is passed in as `workspace` (>= the matching *_workspace_bytes(); may be uninitialised).
 *   - Return value: KMU_OK (0) or a negative kmu_status; kmu_last_error() returns a thread-local message.
 *     No exception crosses the boundary, nothing calls exit().  There is NO CPU path: host pointers are UB.
 *   - Thread-safe: no global mutable state besides the thread-local error string.
 */
#ifndef KMUNET_H_
#define KMUNET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMU_VERSION 100 /* 0.1.0 */

typedef enum {
  KMU_OK = 0,
  KMU_ERR_BAD_ARG = -1,     /* null pointer, non-positive size, inconsistent shape */
  KMU_ERR_UNSUPPORTED = -2, /* shape/precision outside what the kernel family implements */
  KMU_ERR_WORKSPACE = -3,   /* workspace_bytes too small */
  KMU_ERR_LAUNCH = -4,      /* cudaGetLastError() after a launch */
  KMU_ERR_DEVICE = -5       /* current device is not sm_100 */
} kmu_status;

typedef enum {
  KMU_PREC_FP32 = 0, /* CUDA-core fp32 FMA: matches the reference's fp32 modules to ~1e-6 (parity gate 1e-4)   */
  KMU_PREC_BF16 = 1  /* tcgen05 tensor cores, bf16 operands / fp32 TMEM accumulation (parity gate 2e-2)        */
} kmu_precision;

typedef void* kmu_stream; /* cudaStream_t */

int kmu_version(void);
const char* kmu_last_error(void);
/* 1 when the current CUDA device is compute capability 10.x (tcgen05/TMEM present), else 0. */
int kmu_device_supported(void);
/* number of kernels this library has launched from the calling thread (bench.py's gpu_launches). */
uint64_t kmu_launch_count(void);
/* Bit-reproducible mode (default off, or KMU_DETERMINISTIC=1 in the environment): every libkmunet reduction has a fixed order except
 * DySample's dX scatter (fp32 atomics, as torch's grid_sampler backward); with the mode on that scatter accumulates in 64-bit fixed
 * point scaled by max|dout| (integer atomics commute) at 1.75x the cost of the backward. */
void kmu_set_deterministic(int on);
int kmu_get_deterministic(void);

/* ------------------------------------------------------------------------------------------------------------
 * K: KANConv2d / KANLinear         convKAN/KANConv2Dlayers.py:15-37, convKAN/KANlayers.py:577-610,644-660
 *   y[b,o,ho,wo] = sum_f SiLU(p_f) Wb[o,f] + sum_f sum_j B_j(p_f; grid[f]) Ws[o,f,j] s[o,f]
 *   p_f = zero-padded x[b, c, ho*stride - padding + ki, wo*stride - padding + kj],  f = c*k*k + ki*k + kj.
 * KANLinear on (M,in) is the same call with B=M, Cin=in, H=W=1, ksize=1, stride=1, padding=0.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, Cin, H, W, Cout;
  int32_t ksize, stride, padding;
  int32_t grid_size, spline_order; /* n_basis = grid_size + spline_order; knots per feature = grid_size + 2*order + 1 */
  int32_t precision;               /* kmu_precision */
  int32_t has_scaler;              /* enable_standalone_scale_spline */
  /* Caller-asserted fact about `grid` (checked once per module on the host, KANLinear never changes it outside
   * update_grid): every feature row equals the same uniform knot vector t_j = grid_t0 + j*grid_h.  Required by the
   * tcgen05 family (Phi is then a function of the pixel alone); 0 routes the call to the fp32 family. */
  int32_t grid_uniform;
  float grid_t0, grid_h;
} kmu_kanconv2d_desc;

typedef struct {
  kmu_kanconv2d_desc d;
  const float* x;             /* (B,Cin,H,W) */
  const float* base_weight;   /* (Cout, Cin*k*k) */
  const float* spline_weight; /* (Cout, Cin*k*k, n_basis) */
  const float* spline_scaler; /* (Cout, Cin*k*k) or NULL when !has_scaler */
  const float* grid;          /* (Cin*k*k, n_knots) */
  float* y;                   /* (B,Cout,Ho,Wo) */
  void* workspace;
  size_t workspace_bytes;
} kmu_kanconv2d_fwd_args;

typedef struct {
  kmu_kanconv2d_desc d;
  const float* x;
  const float* dy; /* (B,Cout,Ho,Wo) */
  const float* base_weight;
  const float* spline_weight;
  const float* spline_scaler;
  const float* grid;
  float* dx;             /* (B,Cin,H,W), overwritten; may be NULL to skip */
  float* d_base_weight;  /* (Cout,Cin*k*k), overwritten; NULL skips all weight gradients */
  float* d_spline_weight;
  float* d_spline_scaler; /* NULL when !has_scaler */
  void* workspace;
  size_t workspace_bytes;
} kmu_kanconv2d_bwd_args;

size_t kmu_kanconv2d_fwd_workspace_bytes(const kmu_kanconv2d_desc* d);
size_t kmu_kanconv2d_bwd_workspace_bytes(const kmu_kanconv2d_desc* d);
int kmu_kanconv2d_fwd(const kmu_kanconv2d_fwd_args* a, kmu_stream stream);
int kmu_kanconv2d_bwd(const kmu_kanconv2d_bwd_args* a, kmu_stream stream);
/* Which kernel family (forward AND backward) a descriptor resolves to: 0 = fp32 CUDA-core, 1 = tcgen05 implicit GEMM.  The
 * tensor path needs precision KMU_PREC_BF16, ksize 3, stride 1, padding 1, cubic splines with 8 basis functions,
 * Cin % 16 == 0, Cout in {16,32,64} and grid_uniform; anything else runs the fp32 family (still on the GPU). */
int kmu_kanconv2d_path(const kmu_kanconv2d_desc* d);
/* Bring-up / test hook, not part of the drop-in surface: process-wide kernel debug switches (0 = production). */
void kmu_debug_flags(int flags);

/* ------------------------------------------------------------------------------------------------------------
 * S: LayerNorm1D + HSMSSD          vim_block_init/vim_utils_init.py:50-59, vim_block_init/efficient_vim_init.py:33-61
 *   x (B,C,L), L = H*H.  P = dw3x3(Wp x) ; Bm,Cm,dt = split(P) ; A = softmax_L(dt + A_param) ;
 *   hs = x (A*Bm)^T ; [hh;z] = Whz hs ; ho = Wo (hh*SiLU(z) + hh*D) ; y = ho Cm.   Returns y and h = ho.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, C, L, H; /* L == H*H */
  int32_t N;          /* state_dim (64 in KM-UNet) */
  int32_t precision;  /* kmu_precision of the BCdt projection + depthwise conv in the FORWARD call: KMU_PREC_BF16 runs them as
                         one dense 3x3 convolution on tcgen05 (bf16 operands, fp32 TMEM accumulation, 2e-2 gate); everything
                         else, and the whole backward call, is fp32 either way */
} kmu_hsmssd_desc;

typedef struct {
  kmu_hsmssd_desc d;
  const float* x;      /* (B,C,L) */
  const float* w_bcdt; /* (3N,C) */
  const float* w_dw;   /* (3N,3,3) */
  const float* w_hz;   /* (2C,C) */
  const float* w_out;  /* (C,C) */
  const float* A;      /* (N) */
  const float* D;      /* (1) */
  float* y;            /* (B,C,L) */
  float* h;            /* (B,C,N) = ho */
  /* saved for backward (caller-owned, may all be NULL for inference) */
  float* P;     /* (B,3N,L) */
  float* stats; /* (B,2,N): row 0 = max_L(dt), row 1 = sum_L exp(dt - max)  (the A shift cancels) */
  float* hs;    /* (B,C,N) */
  float* hz;    /* (B,2C,N) */
  void* workspace;
  size_t workspace_bytes;
} kmu_hsmssd_fwd_args;

typedef struct {
  kmu_hsmssd_desc d;
  const float* x;
  const float* dy; /* (B,C,L) gradient w.r.t. y */
  const float* dh; /* (B,C,N) gradient w.r.t. h, or NULL (EfficientViMBlock discards h) */
  const float* w_bcdt;
  const float* w_dw;
  const float* w_hz;
  const float* w_out;
  const float* A;
  const float* D;
  const float* P; /* saved by forward */
  const float* stats;
  const float* hs;
  const float* hz;
  const float* h; /* ho */
  float* dx;      /* (B,C,L) overwritten */
  float* d_w_bcdt;
  float* d_w_dw;
  float* d_w_hz;
  float* d_w_out;
  float* d_A; /* written with zeros: the over-L softmax cancels the shift */
  float* d_D;
  void* workspace;
  size_t workspace_bytes;
} kmu_hsmssd_bwd_args;

size_t kmu_hsmssd_fwd_workspace_bytes(const kmu_hsmssd_desc* d);
size_t kmu_hsmssd_bwd_workspace_bytes(const kmu_hsmssd_desc* d);
int kmu_hsmssd_fwd(const kmu_hsmssd_fwd_args* a, kmu_stream stream);
int kmu_hsmssd_bwd(const kmu_hsmssd_bwd_args* a, kmu_stream stream);

/* LayerNorm1D over the channel axis of (B,C,L): y = (x-mean_c)/sqrt(var_c+eps)*w + b  (biased variance). */
int kmu_layernorm1d_fwd(const float* x, const float* weight, const float* bias, float* y, float* rstd /* (B,L) or NULL */,
                        int32_t B, int32_t C, int32_t L, float eps, kmu_stream stream);
/* dweight / dbias (C) are OVERWRITTEN; per-CTA partials in the workspace are summed in a fixed order (bit-reproducible). */
size_t kmu_layernorm1d_bwd_workspace_bytes(int32_t B, int32_t C, int32_t L);
int kmu_layernorm1d_bwd(const float* x, const float* weight, const float* dy, float* dx, float* dweight, float* dbias,
                        int32_t B, int32_t C, int32_t L, float eps, void* workspace, size_t workspace_bytes, kmu_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * D: DySample ('lp' style, dyscope off)        DySample_md.py:49-68
 *   off = (conv1x1(x; Wo, bo)) * 0.25 + init_pos                                  (B, 2*G*s*s, H, W)
 *   out[b, g*Cg+c, s*h+i, s*w+j] = bilinear(x[b, g*Cg+c]; row = clamp(h + off_y), col = clamp(w + off_x))
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, C, H, W;
  int32_t scale, groups;
} kmu_dysample_desc;

typedef struct {
  kmu_dysample_desc d;
  const float* x;        /* (B,C,H,W) */
  const float* w_offset; /* (2*G*s*s, C) */
  const float* b_offset; /* (2*G*s*s) */
  const float* init_pos; /* (2*G*s*s) */
  float* offset;         /* (B,2*G*s*s,H,W) written (saved for backward) */
  float* out;            /* (B,C,s*H,s*W) */
} kmu_dysample_fwd_args;

typedef struct {
  kmu_dysample_desc d;
  const float* x;
  const float* w_offset;
  const float* offset; /* saved by forward */
  const float* dout;   /* (B,C,s*H,s*W) */
  float* dx;           /* (B,C,H,W) overwritten */
  float* d_w_offset;   /* (2*G*s*s, C) overwritten */
  float* d_b_offset;   /* (2*G*s*s) overwritten */
  void* workspace;
  size_t workspace_bytes;
} kmu_dysample_bwd_args;

size_t kmu_dysample_bwd_workspace_bytes(const kmu_dysample_desc* d);
int kmu_dysample_fwd(const kmu_dysample_fwd_args* a, kmu_stream stream);
int kmu_dysample_bwd(const kmu_dysample_bwd_args* a, kmu_stream stream);
/* The reference's DySample.sample(x, offset) alone (DySample_md.py:49-61), for styles that build the offset
 * differently ('pl', dyscope). */
int kmu_dysample_sample_fwd(const kmu_dysample_desc* d, const float* x, const float* offset, float* out, kmu_stream stream);
int kmu_dysample_sample_bwd(const kmu_dysample_desc* d, const float* x, const float* offset, const float* dout,
                            float* dx /* zero-initialised, accumulated */, float* doffset /* overwritten */,
                            kmu_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * G': DAGEM's deformable 3x3 convolution   DAGEM_md.py:46 (DeformConv2d(C, C, kernel_size=3, padding=1)), :98-101
 *   (offset = offset_conv(x); deformed = deform_conv(x, offset)).  torchvision.ops.deform_conv2d semantics with one offset
 *   group, stride 1, dilation 1: out[b,o,p] = bias[o] + sum_{c,t} W[o,c,t] bilinear(x[b,c], p + tap_t - 1 + offset[b,2t:2t+2,p]),
 *   zero outside (-1,H) x (-1,W).  Every kernel runs on the given stream (torchvision's CUDA op uses the legacy default stream
 *   and is therefore NOT CUDA-graph capturable) and all reductions have a fixed order (no atomics).  H*W <= 4096.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, C, H, W, Cout;
} kmu_deform_desc;

size_t kmu_deformconv3x3_bwd_workspace_bytes(const kmu_deform_desc* d);
int kmu_deformconv3x3_fwd(const kmu_deform_desc* d, const float* x /* (B,C,H,W) */, const float* offset /* (B,18,H,W): (dy,dx) per tap */,
                          const float* weight /* (Cout,C,3,3) */, const float* bias /* (Cout) or NULL */, float* out /* (B,Cout,H,W) */,
                          kmu_stream stream);
int kmu_deformconv3x3_bwd(const kmu_deform_desc* d, const float* x, const float* offset, const float* weight,
                          const float* dout /* (B,Cout,H,W) */, float* dx, float* doffset, float* dweight,
                          float* dbias /* may be NULL */, void* workspace, size_t workspace_bytes, kmu_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * G: DAGEM attention-gated fusion        DAGEM_md.py:62-92 (edge / vertex gating), :104-110 (final aggregation)
 *   s = x . sum_k a_k nb_k + a_b ; agg = ReLU(BN0(s)) ; ue_k = We [x; x . nb_k] + e_b ; uvp = Wv [x; agg] + v_b ;
 *   r = sum_k r_k ReLU(BN1(ue_k)) + r_b ; z = Wf [deformed; ReLU(BN2(uvp)) . ReLU(BN3(r))] ; out = ReLU(BN4(z))
 *   nb_k = circular neighbours x[h-1], x[h+1], x[w-1], x[w+1] (:64-67).  `deformed` = deform_conv(x, offset_conv(x)) + x
 *   (:95-104) is the caller's tensor.  BatchNorm order bn[0..4] = edge_aggregation_func.1, edge_update_func.1,
 *   vertex_update_func.1, update_edge_reduce_func.1, final_aggregation_layer.1.  training != 0: batch statistics
 *   (biased variance) and running-stat update with `momentum` (unbiased variance), as torch.nn.BatchNorm does.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, C, H, W; /* C = input_channels in {8,16,32,64} */
  int32_t training;
  float momentum, eps;
} kmu_dagem_desc;

typedef struct {
  const float* weight;  /* gamma */
  const float* bias;    /* beta */
  float* running_mean;  /* updated in place when training; read when not */
  float* running_var;
} kmu_dagem_bn;

typedef struct {
  kmu_dagem_desc d;
  const float* x;        /* (B,C,H,W) */
  const float* deformed; /* (B,C,H,W) */
  const float* ea_w;     /* edge_aggregation_func.0.weight (1,4) */
  const float* ea_b;     /* (1) */
  const float* vu_w;     /* vertex_update_func.0.weight (C/2, 2C) */
  const float* vu_b;     /* (C/2) */
  const float* eu_w;     /* edge_update_func.0.weight (C/2, 2C) */
  const float* eu_b;     /* (C/2) */
  const float* er_w;     /* update_edge_reduce_func.0.weight (1,4) */
  const float* er_b;     /* (1) */
  const float* wf;       /* final_aggregation_layer.0.weight (C, C + C/2) */
  kmu_dagem_bn bn[5];
  float* out;            /* (B,C,H,W) */
  float* saved;          /* kmu_dagem_saved_bytes(): pre-BN activations + BN statistics, input of the backward call */
  void* workspace;
  size_t workspace_bytes;
} kmu_dagem_fwd_args;

typedef struct {
  kmu_dagem_desc d;
  const float* x;
  const float* deformed;
  const float* dout; /* (B,C,H,W) */
  const float* saved;
  const float* ea_w;
  const float* vu_w;
  const float* eu_w;
  const float* er_w;
  const float* wf;
  float* dx;         /* gradient through the gating path only (the caller adds the deform/residual path) */
  float* d_deformed; /* (B,C,H,W) */
  float* d_ea_w;
  float* d_ea_b;
  float* d_vu_w;
  float* d_vu_b;
  float* d_eu_w;
  float* d_eu_b;
  float* d_er_w;
  float* d_er_b;
  float* d_wf;
  float* d_bn_weight[5];
  float* d_bn_bias[5];
  void* workspace;
  size_t workspace_bytes;
} kmu_dagem_bwd_args;

size_t kmu_dagem_saved_bytes(const kmu_dagem_desc* d);
size_t kmu_dagem_fwd_workspace_bytes(const kmu_dagem_desc* d);
size_t kmu_dagem_bwd_workspace_bytes(const kmu_dagem_desc* d);
int kmu_dagem_fwd(const kmu_dagem_fwd_args* a, kmu_stream stream);
int kmu_dagem_bwd(const kmu_dagem_bwd_args* a, kmu_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * S3: EfficientViMBlock shell       vim_block_init/efficient_vim_init.py:82-96, vim_utils_init.py:83-89,128-130
 *   kmu_bnmix:     y = BN2d(x) [-> ReLU] [-> (1 - sigmoid(alpha_c)) res + sigmoid(alpha_c) y]   (train or eval statistics)
 *   kmu_dwconv3x3: depthwise 3x3 convolution, stride 1, zero padding 1, optional bias (ConvLayer2D with groups = dim;
 *                  DirectionAttention.conv, KM_UNetV3_SH.py:223)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, C, HW;
  int32_t training; /* batch statistics + running-stat update, else running statistics */
  int32_t relu;     /* ReLU after the affine normalisation */
  int32_t mix;      /* layer-scale mix with `res` through sigmoid(alpha) */
  float momentum, eps;
} kmu_bnmix_desc;

typedef struct {
  kmu_bnmix_desc d;
  const float* x;      /* (B,C,HW) */
  const float* weight; /* (C) gamma */
  const float* bias;   /* (C) beta */
  float* running_mean; /* (C) */
  float* running_var;  /* (C) */
  const float* res;    /* (B,C,HW) or NULL */
  const float* alpha;  /* (C) pre-sigmoid layer scale or NULL */
  float* y;            /* (B,C,HW) */
  float* stat;         /* (C,2) mean, rstd: input of the backward call */
  void* workspace;
  size_t workspace_bytes;
} kmu_bnmix_fwd_args;

typedef struct {
  kmu_bnmix_desc d;
  const float* x;
  const float* dy;
  const float* weight;
  const float* bias;
  const float* stat;
  const float* res;
  const float* alpha;
  float* dx;
  float* d_weight;
  float* d_bias;
  float* d_res;   /* (B,C,HW) or NULL when !mix */
  float* d_alpha; /* (C) or NULL when !mix */
  void* workspace;
  size_t workspace_bytes;
} kmu_bnmix_bwd_args;

size_t kmu_bnmix_workspace_bytes(const kmu_bnmix_desc* d);
int kmu_bnmix_fwd(const kmu_bnmix_fwd_args* a, kmu_stream stream);
int kmu_bnmix_bwd(const kmu_bnmix_bwd_args* a, kmu_stream stream);

typedef struct {
  int32_t B, C, H, W;
} kmu_dwconv3x3_desc;

size_t kmu_dwconv3x3_bwd_workspace_bytes(const kmu_dwconv3x3_desc* d);
int kmu_dwconv3x3_fwd(const kmu_dwconv3x3_desc* d, const float* x, const float* w /* (C,9) */, const float* bias /* (C) or NULL */,
                      float* y, kmu_stream stream);
/* dx, dw (C,9), dbias (C) are overwritten; dx == NULL skips the input gradient, dw == NULL the weight / bias gradients. */
int kmu_dwconv3x3_bwd(const kmu_dwconv3x3_desc* d, const float* x, const float* dy, const float* w, float* dx, float* dw,
                      float* dbias, void* workspace, size_t workspace_bytes, kmu_stream stream);
/* Backward with a residual gradient folded in: dx = conv^T(dy) + dx_add (dx may alias dx_add).  EfficientViMBlock feeds the same
 * tensor to the depthwise convolution and to the layer-scale mix behind it (efficient_vim_init.py:85,93): the mix's d(res) joins the
 * convolution's input gradient here instead of in a separate add kernel. */
int kmu_dwconv3x3_bwd_add(const kmu_dwconv3x3_desc* d, const float* x, const float* dy, const float* w, const float* dx_add, float* dx,
                          float* dw, float* dbias, void* workspace, size_t workspace_bytes, kmu_stream stream);
/* The same convolution followed by a per-(b, c) factor: y = scale[b, c] * (conv(x) + bias) -- DirectionAttention's
 * `self.conv(attn) * weight[:, :, None, None]` (KM_UNetV3_SH.py:130-151) without the broadcast multiply and its three backward
 * kernels.  scale is (B, C); dscale (B, C) = sum_hw dy * (conv + bias) falls out of the weight-gradient partials. */
int kmu_dwconv3x3_scaled_fwd(const kmu_dwconv3x3_desc* d, const float* x, const float* w, const float* bias, const float* scale,
                             float* y, kmu_stream stream);
int kmu_dwconv3x3_scaled_bwd(const kmu_dwconv3x3_desc* d, const float* x, const float* dy, const float* w, const float* bias,
                             const float* scale, float* dx, float* dw, float* dbias, float* dscale, void* workspace,
                             size_t workspace_bytes, kmu_stream stream);

/* Pointwise (1x1) convolution on NCHW: y[b,o,p] = bias[o] + sum_c w[o,c] x[b,c,p]   (vim_utils_init.py:122-130 FFN,
 * KM_UNetV3_SH.py:59,118-122,178,221).  The weight gradient kernel needs Cin a power of two in [16,1024] and
 * Cout <= 16*1024/Cin (kmu_pwconv_wgrad_supported); forward and input gradient take any channel counts. */
typedef struct {
  int32_t B, Cin, Cout, HW;
} kmu_pwconv_desc;

int kmu_pwconv_wgrad_supported(const kmu_pwconv_desc* d);
size_t kmu_pwconv_bwd_workspace_bytes(const kmu_pwconv_desc* d);
int kmu_pwconv_fwd(const kmu_pwconv_desc* d, const float* x, const float* w /* (Cout,Cin) */, const float* bias /* (Cout) or NULL */,
                   float* y, kmu_stream stream);
/* dx, dw (Cout,Cin), dbias (Cout) are overwritten; NULL dx skips the input gradient, NULL dw the weight / bias gradients. */
int kmu_pwconv_bwd(const kmu_pwconv_desc* d, const float* x, const float* dy, const float* w, float* dx, float* dw, float* dbias,
                   void* workspace, size_t workspace_bytes, kmu_stream stream);

/* ------------------------------------------------------------------------------------------------------------
 * Caller-side fusions
 *   kmu_triplenorm: TripleNorm.forward, KM_UNetV3_SH.py:277-284:
 *       y = (GroupNorm(1,C; gh,bh)(x) + GroupNorm(1,C; gw,bw)(x) + LayerNorm_C(gc,bc)(x)) / 3     x (B,C,HW), C in {16,32,64}
 *   kmu_qkv_gate:   DirectionAttention.forward, KM_UNetV3_SH.py:259-261:  out = sigmoid(q k) v, qkv (B,3C,HW) -> (B,C,HW)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t B, C, HW;
  float eps_gn, eps_ln;
} kmu_triplenorm_desc;

typedef struct {
  kmu_triplenorm_desc d;
  const float* x;
  const float* gh; /* norm_h.weight (C) */
  const float* bh;
  const float* gw; /* norm_w.weight */
  const float* bw;
  const float* gc; /* norm_c.weight */
  const float* bc;
  float* y;
  float* gstat; /* (B,2) per-sample mean, rstd: input of the backward call */
  void* workspace;
  size_t workspace_bytes;
} kmu_triplenorm_fwd_args;

typedef struct {
  kmu_triplenorm_desc d;
  const float* x;
  const float* dy;
  const float* gstat;
  const float* gh;
  const float* gw;
  const float* gc;
  float* dx;
  float* d_gh;
  float* d_bh;
  float* d_gw;
  float* d_bw;
  float* d_gc;
  float* d_bc;
  void* workspace;
  size_t workspace_bytes;
} kmu_triplenorm_bwd_args;

size_t kmu_triplenorm_workspace_bytes(const kmu_triplenorm_desc* d);
int kmu_triplenorm_fwd(const kmu_triplenorm_fwd_args* a, kmu_stream stream);
int kmu_triplenorm_bwd(const kmu_triplenorm_bwd_args* a, kmu_stream stream);
int kmu_qkv_gate_fwd(const float* qkv, float* out, int32_t B, int32_t C, int32_t HW, kmu_stream stream);
int kmu_qkv_gate_bwd(const float* qkv, const float* dout, float* dqkv, int32_t B, int32_t C, int32_t HW, kmu_stream stream);

/* kmu_combine3: EnhancedViMBlock.forward, KM_UNetV3_SH.py:349-368: x + DropPath(g0 f0 + g1 f1 + g2 f2) as ONE pass,
 *     out[b] = x[b] + sum_i coef[b][i] f_i[b],   coef (B,3) = softmax gate weight x per-sample DropPath factor (built by the caller);
 * backward: df_i = coef[b][i] dy, dcoef[b][i] = sum dy . f_i (per-CTA partials reduced in fixed order); dx = dy is the caller's.
 * n_per_b = C*H*W elements per sample, a multiple of 4. */
/* kmu_lerpmix: EfficientViMBlock.forward, vim_block_init/efficient_vim_init.py:89-90:
 *     y = (1 - sigmoid(alpha_c)) x + sigmoid(alpha_c) m,  alpha (C) raw;  backward: dx, dm, dalpha (fixed-order reduction).  HW % 4 == 0. */
size_t kmu_lerpmix_bwd_workspace_bytes(int32_t B, int32_t C, int64_t HW);
int kmu_lerpmix_fwd(const float* x, const float* m, const float* alpha, float* y, int32_t B, int32_t C, int64_t HW, kmu_stream stream);
int kmu_lerpmix_bwd(const float* x, const float* m, const float* dy, const float* alpha, float* dx, float* dm, float* dalpha, int32_t B,
                    int32_t C, int64_t HW, void* workspace, size_t workspace_bytes, kmu_stream stream);
/* HybridLoss (train_shanghai.py:298-326, train_LAPS.py:347-375) as four streaming passes around the two
 * banded GEMMs of the SSIM filter (which the caller runs): stats -> scal[8] (sums, minima, 1/range), stack -> the five maps
 * (p_n, t_n, p_n^2, t_n^2, p_n t_n), ssim -> loss value + the three derivative maps that reach the prediction (d/dmu_p, d/dE[pp],
 * d/dE[pt], already scaled by -(1 - alpha)/n_valid), bwd -> dpred from the back-filtered maps and grad_out (device scalar).
 * n = elements of pred (multiple of 4), n_valid = elements of one filtered map. */
size_t kmu_hybridloss_workspace_bytes(void);
int kmu_hybridloss_stats(const float* pred, const float* target, int64_t n, float* scal, void* workspace, size_t workspace_bytes,
                         kmu_stream stream);
int kmu_hybridloss_stack(const float* pred, const float* target, const float* scal, float* stack5, int64_t n, kmu_stream stream);
int kmu_hybridloss_ssim(const float* filtered5, const float* scal, float* gm3, float* loss_out, int64_t n_valid, int64_t n_full,
                        float alpha, float c1, float c2, void* workspace, size_t workspace_bytes, kmu_stream stream);
int kmu_hybridloss_bwd(const float* pred, const float* target, const float* scal, const float* dstack3, const float* grad_out,
                       float* dpred, int64_t n, float alpha, kmu_stream stream);
/* kmu_groupnorm_fwd: nn.GroupNorm forward (StableHybridKANConv.pre_norm KM_UNetV3_SH.py:72-94, MultiScaleFusion :292, output norm
 * :455): y = (x - mean_g) rstd_g gamma_c + beta_c, statistics over (C/G, HW) per sample and group; mean / rstd (B*G) are outputs for
 * the backward (ATen's native_group_norm_backward takes them).  HW a multiple of 4. */
size_t kmu_groupnorm_fwd_workspace_bytes(int32_t B, int32_t C, int64_t HW, int32_t G);
int kmu_groupnorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int32_t B, int32_t C,
                      int64_t HW, int32_t G, float eps, void* workspace, size_t workspace_bytes, kmu_stream stream);
/* kmu_resize_bilinear_ac: F.interpolate(x, size, mode='bilinear', align_corners=True) of the skip connections
 * (KM_UNetV3_SH.py:493-512), forward; x (planes, H, W) -> out (planes, OH, OW).  The backward stays ATen's. */
int kmu_resize_bilinear_ac_fwd(const float* x, float* out, int64_t planes, int32_t H, int32_t W, int32_t OH, int32_t OW,
                               kmu_stream stream);
size_t kmu_combine3_bwd_workspace_bytes(int32_t B, int64_t n_per_b);
int kmu_combine3_fwd(const float* x, const float* f0, const float* f1, const float* f2, const float* coef, float* out, int32_t B,
                     int64_t n_per_b, kmu_stream stream);
int kmu_combine3_bwd(const float* dy, const float* f0, const float* f1, const float* f2, const float* coef, float* df0, float* df1,
                     float* df2, float* dcoef, int32_t B, int64_t n_per_b, void* workspace, size_t workspace_bytes, kmu_stream stream);

/* Dense "same" convolution with at most 9 taps (1x1, 1x3, 3x1, 3x3; stride 1, zero padding k/2) on NCHW:
 * DirectionViM.proj (KM_UNetV3_SH.py:172-176), the decoder / fusion 3x3 convs (:292,299,427,437,439).
 * w (Cout,Cin,kh,kw).  kmu_smallconv_supported: the weight-gradient kernel needs Cin a power of two in [16,1024],
 * Cin*Cout <= 16384, and Cin*kh*kw*128 bytes of weights must fit shared memory. */
typedef struct {
  int32_t B, Cin, Cout, H, W, kh, kw;
} kmu_smallconv_desc;

int kmu_smallconv_supported(const kmu_smallconv_desc* d);
size_t kmu_smallconv_bwd_workspace_bytes(const kmu_smallconv_desc* d);
int kmu_smallconv_fwd(const kmu_smallconv_desc* d, const float* x, const float* w, const float* bias /* or NULL */, float* y,
                      kmu_stream stream);
int kmu_smallconv_bwd(const kmu_smallconv_desc* d, const float* x, const float* dy, const float* w, float* dx, float* dw,
                      float* dbias, void* workspace, size_t workspace_bytes, kmu_stream stream);

/* IntelligentWaveletPoolingModule.forward (WPL/iwp.py:116-132, Haar DWT of :9-113): x (B,C,H,W), H and W even ->
 * out (B,C,H/2,W/2) = fusion_conv([LL ; mean over channels of (LH,HL,HH)]), last high-pass row / column zero (:79-82).
 * wf = fusion_conv.weight (C, C+1), bias (C).  C in {16,32,64}. */
size_t kmu_iwp_bwd_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W);
int kmu_iwp_fwd(const float* x, const float* wf, const float* bias, float* out, int32_t B, int32_t C, int32_t H, int32_t W,
                kmu_stream stream);
int kmu_iwp_bwd(const float* x, const float* wf, const float* dout, float* dx, float* d_wf, float* d_bias, int32_t B, int32_t C,
                int32_t H, int32_t W, void* workspace, size_t workspace_bytes, kmu_stream stream);

/* tcgen05 path of the same pointwise convolution (bf16 operands, fp32 TMEM accumulation, 2e-2 gate): Cin, Cout multiples of
 * 16 in [16,256]; the weight-gradient GEMM additionally needs Cin <= 240 (kmu_pwconv_tc_wgrad_supported) -- callers route the
 * remaining pairs' weight gradient through kmu_pwconv_bwd.  One workspace size serves forward and backward. */
int kmu_pwconv_tc_supported(const kmu_pwconv_desc* d);
int kmu_pwconv_tc_wgrad_supported(const kmu_pwconv_desc* d);
size_t kmu_pwconv_tc_workspace_bytes(const kmu_pwconv_desc* d);
int kmu_pwconv_tc_fwd(const kmu_pwconv_desc* d, const float* x, const float* w, const float* bias, float* y, void* workspace,
                      size_t workspace_bytes, kmu_stream stream);
int kmu_pwconv_tc_bwd(const kmu_pwconv_desc* d, const float* x, const float* dy, const float* w, float* dx, float* dw, float* dbias,
                      void* workspace, size_t workspace_bytes, kmu_stream stream);

/* Fused backward of the same pointwise convolution (replaces the autograd pair of vim_utils_init.py:122-130 FFN /
 * KM_UNetV3_SH.py:59,118-122,178,221 1x1 convolutions): ONE persistent TMA -> tcgen05 kernel reads x and dy once and writes dx,
 * dW and (if dbias != NULL) db.  bf16 operands, fp32 TMEM accumulation (2e-2 gate).  Supported when Cin, Cout are multiples of
 * 16 (Cin <= 240, Cout <= 256), HW >= 128 and a multiple of 4, x / dy 16-byte aligned and the tile ring fits shared memory
 * (kmu_pwconv_fused_bwd_supported); dx and dw are both required. */
int kmu_pwconv_fused_bwd_supported(const kmu_pwconv_desc* d);
/* Forward of the same convolution on the same pipeline (TMA fp32 box -> bf16 planes -> tcgen05 -> NCHW stores); same shape rules
 * with Cin <= 256. */
int kmu_pwconv_tma_fwd_supported(const kmu_pwconv_desc* d);
size_t kmu_pwconv_tma_fwd_workspace_bytes(const kmu_pwconv_desc* d);
int kmu_pwconv_tma_fwd(const kmu_pwconv_desc* d, const float* x, const float* w, const float* bias, float* y, void* workspace,
                       size_t workspace_bytes, kmu_stream stream);
size_t kmu_pwconv_fused_bwd_workspace_bytes(const kmu_pwconv_desc* d);
int kmu_pwconv_fused_bwd(const kmu_pwconv_desc* d, const float* x, const float* dy, const float* w, float* dx, float* dw,
                         float* dbias, void* workspace, size_t workspace_bytes, kmu_stream stream);

/* kmu_adamw_step: the optimizer of train_shanghai.py:342,180 -- `optim.AdamW(params, lr, weight_decay)` / `optimizer.step()`
 * (torch.optim.AdamW semantics: decoupled weight decay, bias-corrected moments, no amsgrad) over every tensor of a parameter group
 * in one launch.  `entries` and `chunks` are DEVICE tables built by the caller: one entry per tensor (all fp32, contiguous), one
 * chunk per kmu_adamw_chunk_elems() elements of a tensor, encoded as (entry index << 32) | chunk number within the tensor.
 * `step` is a device scalar holding the number of updates applied so far; the call uses step + 1 for the bias corrections and
 * then increments it (so the call can be captured into a CUDA graph).  `lr_dev` (optional) overrides `lr` with a device scalar;
 * gradients are multiplied by `grad_scale` on the fly (1 / loss-scale, or 1). */
typedef struct kmu_adamw_entry {
  float* p;
  const float* g;
  float* m;
  float* v;
  int64_t n;
} kmu_adamw_entry;

typedef struct kmu_adamw_args {
  const kmu_adamw_entry* entries;
  const int64_t* chunks;
  int32_t n_entries;
  int32_t n_chunks;
  float* step;
  const float* lr_dev;
  double lr, beta1, beta2, eps, weight_decay, grad_scale;   /* doubles: 1 - beta2 must not inherit the fp32 rounding of beta2 */
} kmu_adamw_args;

int32_t kmu_adamw_chunk_elems(void);
int kmu_adamw_step(const kmu_adamw_args* a, kmu_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* KMUNET_H_ */
