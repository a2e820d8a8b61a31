/*
 * tdnet_b200 C-ABI: the drop-in boundary of the B200-native TDNet inference hot path.
 *
 * The reference (feinanshan/TDNet, /root/reference) is pure Python/PyTorch and has no FFI of its
 * own: on its hot path every numeric step is a torch library call.  The entry
 * points below are what a maintainer would bind in place of those calls; each one names the
 * reference call site (file:line under /root/reference/Testing/model/pspnet/) it replaces.
 * INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - Plain C: POD structs, raw device pointers, sizes.  No torch types, no exceptions.
 *   - Every function returns 0 on success or a negative tdn_status; tdn_strerror() names it and
 *     tdn_last_error() returns a thread-local detail string.
 *   - Work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises, nothing
 *     allocates: the caller owns every buffer including workspaces, so a whole frame can be
 *     captured in a CUDA graph.
 *   - Activations are NHWC ("pixel-major": channels contiguous), described by tdn_tensor with
 *     explicit element strides so that channel slices, stride-4 sub-sampled views
 *     (transformer.py:26 MaxPool2d(kernel 1, stride 4)) and token matrices [P, C] are all views.
 *   - There is no CPU path.  On a device that is not sm_100 the tensor-core entry points fail with
 *     TDN_ERR_ARCH; nothing falls back silently.
 */
#ifndef TDNET_B200_H_
#define TDNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDN_ABI_VERSION 2 /* 2: `flags` appended to tdn_tc_conv_desc and tdn_attention_desc */

typedef enum tdn_status {
  TDN_OK = 0,
  TDN_ERR_INVALID = -1,   /* bad argument (null pointer, misaligned, inconsistent dims) */
  TDN_ERR_UNSUPPORTED = -2, /* shape / option outside what the kernels implement */
  TDN_ERR_CUDA = -3,      /* a CUDA runtime/driver call failed; see tdn_last_error() */
  TDN_ERR_ARCH = -4,      /* device is not sm_100 (Blackwell B200) */
  TDN_ERR_WORKSPACE = -5  /* caller-provided workspace too small */
} tdn_status;

typedef enum tdn_dtype {
  TDN_F32 = 0,    /* one fp32 plane */
  TDN_SPLIT16 = 1 /* two fp16 planes hi/lo with value = hi + lo: the fp32-faithful operand format of
                     the tcgen05 kernels (22+ significant bits, see DESIGN.md "exact mode") */
} tdn_dtype;

typedef enum tdn_act {
  TDN_ACT_NONE = 0,
  TDN_ACT_RELU = 1,
  TDN_ACT_LEAKY_RELU = 2 /* slope in tdn_conv2d_desc.leaky_slope (reference: nn.LeakyReLU() = 0.01) */
} tdn_act;

/* NHWC view.  Element (n,y,x,ch) lives at data[n*stride_n + y*stride_h + x*stride_w + ch].
 * Strides are in elements.  For TDN_SPLIT16 `data` is the hi plane and `data_lo` the lo plane,
 * both with the same strides. */
typedef struct tdn_tensor {
  void* data;
  void* data_lo;
  int32_t dtype; /* tdn_dtype */
  int32_t n, h, w, c;
  int64_t stride_n, stride_h, stride_w;
} tdn_tensor;

/* ------------------------------------------------------------------------------------------------
 * tdn_conv2d: out = act( conv(in, weight) * scale[co] + bias[co] + residual )
 *
 * Replaces, with eval-mode BatchNorm folded into (scale, bias) by the host:
 *   resnet.py:43-59   BasicBlock  conv3x3 -> BN -> ReLU -> conv3x3 -> BN -> (+residual) -> ReLU
 *   resnet.py:91-111  Bottleneck  conv1x1/3x3/1x1 with BN/ReLU and the residual add
 *   resnet.py:122-137 stem convs (input first converted by tdn_image_to_nhwc)
 *   resnet.py:172-178 downsample conv1x1 + BN
 *   td4_psp18.py:255-266  PSP branch conv1x1 + BN + ReLU on the pooled bins
 *   transformer.py:18-24,142-161  Encoding w_qs / w_ks / w_vs 1x1 convs (+bias, BN+LeakyReLU)
 *   transformer.py:84-86  Attention.fc applied per token
 *   td4_psp18.py:295-299  FCNHead conv3x3 + BN + ReLU and the 1x1 classifier
 *   transformer.py:128,137  torch.bmm(q, k^T) and torch.bmm(attn, v) in the SIMT attention path
 *     (a GEMM is a 1x1 convolution over a [1,1,rows,K] view; `batch` handles the bmm batch).
 *
 * weight: fp32, "K-major" [cout][kh][kw][cin] when weight_kn == 0, or [kh*kw*cin][cout] when
 *         weight_kn == 1 (used for attn @ v where v is [P', d_v]).
 * scale/bias: fp32 [cout] or NULL (scale -> 1, bias -> 0).  residual: optional, same dims as out.
 * batch > 1: in/out/residual/weight advance by their *_batch_stride (elements) per batch entry.
 * ---------------------------------------------------------------------------------------------- */
typedef struct tdn_conv2d_desc {
  tdn_tensor in;
  tdn_tensor out;
  tdn_tensor residual; /* residual.data == NULL -> none */
  const float* weight;
  const float* scale;
  const float* bias;
  int32_t cout;
  int32_t kh, kw;
  int32_t stride, pad, dilation;
  int32_t act;         /* tdn_act */
  float leaky_slope;
  int32_t weight_kn;
  int32_t batch;       /* >= 1 */
  int64_t in_batch_stride, out_batch_stride, residual_batch_stride, weight_batch_stride;
} tdn_conv2d_desc;

int tdn_conv2d(const tdn_conv2d_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * tdn_conv2d_tc: the tcgen05 (5th-gen tensor core) implementation of the same operator for the
 * shapes that dominate the frame: "same"-padded convolutions (1x1 / 3x3, any dilation, stride 1 or 2) and
 * plain GEMMs with cin %% 64 == 0, on SPLIT16 operands.  fp32-faithful: three fp16 tensor-core products per
 * K step (hi*lo, lo*hi, hi*hi) accumulate into one fp32 TMEM accumulator ("exact mode").
 * Same call sites as tdn_conv2d; additionally transformer.py:128,137 (q k^T, attn v) where the
 * "weight" operand is itself an activation (weight_batched = 1: one [cout][K] matrix per image).
 *
 * weight_hi/lo: fp16 [cout][kh*kw*cin] K-major with row pitch weight_ld (elements).
 * bias_along_m: bias is indexed by the output pixel (row of the GEMM) instead of the channel; used to
 *   produce V'^T = W_fc V^T + b directly in the K-major layout the next GEMM wants.
 * out: SPLIT16 or F32; out_f32_copy (optional) additionally receives fp32 with out's strides.
 * range_flag (optional, device int): set to 1 if a SPLIT16 output exceeded the fp16 range guard.
 * Fails with TDN_ERR_ARCH on anything but sm_100.
 * ---------------------------------------------------------------------------------------------- */
typedef struct tdn_tc_conv_desc {
  tdn_tensor in;
  tdn_tensor out;
  tdn_tensor residual; /* residual.data == NULL -> none; SPLIT16 or F32 */
  float* out_f32_copy;
  const void* weight_hi;
  const void* weight_lo;
  int64_t weight_ld;
  int64_t weight_batch_stride;
  int32_t weight_batched;
  int32_t bias_along_m;
  const float* scale;
  const float* bias;
  int32_t cout;
  int32_t kh, kw;
  int32_t dilation;
  int32_t act;
  float leaky_slope;
  int32_t* range_flag;
  int32_t stride; /* 0 or 1: stride 1; 2: stride-2 conv (TMA element strides), out = ceil(in / 2) */
  int32_t variant; /* TDN_TC_AUTO (0): the library picks the kernel; otherwise force one (tests / tuning;
                      TDN_ERR_UNSUPPORTED if the geometry does not fit it).  All variants compute the same
                      products in the same order and are bit-identical. */
  int32_t flags;   /* TDN_TC_FLAG_* (0 = exact mode) */
} tdn_tc_conv_desc;

/* Opt-in FAST mode: one fp16 tensor-core product per K step (the hi planes only) instead of the three exact-mode
 * products -- roughly bf16/fp16-GEMM accuracy (11-bit operands), NOT the reference's fp32 arithmetic.  Never the parity
 * gate; bench.py reports its speed next to its measured arg-max mismatch rate. */
enum { TDN_TC_FLAG_FAST = 1 };

enum {
  TDN_TC_AUTO = 0,
  TDN_TC_BASE = 1, /* one CTA per 128-pixel x 64/128-channel tile, one A box per filter tap */
  TDN_TC_HALO = 2, /* 3x3 stride-1 dilation<=2: one halo-region load per channel block */
  TDN_TC_PAIR = 3, /* cout % 128 == 0, shared weights: 2-CTA clusters (tcgen05 cta_group::2), M 256 x N 256/128 */
  /* Two measured alternatives of the pair kernel (cout % 256 == 0), bit-identical to it, never picked by TDN_TC_AUTO
     (both leave the frame rate unchanged): */
  TDN_TC_PAIR_TAIL = 4, /* full rounds as N = 256 pair tiles, the ragged last round as N = 128 pair tiles (second launch) */
  TDN_TC_PAIR_QUAD = 5, /* clusters of two pairs that share the weight tile by TMA multicast */
  TDN_TC_HALO_SW = 6,   /* 3x3 stride-1 dilation<=2: one 128-byte-swizzled halo-region load per channel block */
  TDN_TC_BASE_TS = 7,   /* TDN_TC_BASE with the A tile copied to tensor memory per K block (tcgen05.cp) and read from there */
  TDN_TC_PAIR_BAND = 8  /* 3x3 stride-1 dilation<=4 cout%256==0 on CTA pairs: one activation band per channel block and filter row */
};

int tdn_conv2d_tc(const tdn_tc_conv_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * tdn_attention_tc: the fused attention-propagation kernel (tcgen05/TMEM/TMA, exact mode).
 *   out[b, q, :] = softmax_k( Q[b,q,:] . K[b,k,:] / sqrt(d_k) ) @ V'[b,k,:]  + residual[b,q,:]
 * Replaces ScaledDotProductAttention.forward (transformer.py:126-139: bmm, /temperature, softmax,
 * bmm) and -- with Attention.fc (transformer.py:84-86) folded into V' by the caller, which is exact
 * because softmax rows sum to one -- the whole of Attention.forward (transformer.py:71-92) plus the
 * `v + V_queue[j]` / `v_4_ + v_cur` adds of td4_psp18.py:146-151.  The [pq x pk] attention matrix is
 * never written to memory.
 *   q  : SPLIT16 token matrix [n][pq][64]   (row pitch q_ld, batch stride q_batch_stride; elements)
 *   k  : SPLIT16 token matrix [n][pk][64]
 *   vt : SPLIT16 V'^T [n][d_v][pk padded to 64] (keys contiguous; pad columns must be zero/finite)
 *   out / residual : [n,1,pq,d_v] token views, SPLIT16 or F32.  d_k must be 64, d_v % 128 == 0.
 * ---------------------------------------------------------------------------------------------- */
typedef struct tdn_attention_desc {
  const void* q_hi;
  const void* q_lo;
  int64_t q_ld, q_batch_stride;
  const void* k_hi;
  const void* k_lo;
  int64_t k_ld, k_batch_stride;
  const void* vt_hi;
  const void* vt_lo;
  int64_t vt_ld, vt_batch_stride;
  tdn_tensor out;
  tdn_tensor residual; /* residual.data == NULL -> none */
  int32_t n, pq, pk, d_k, d_v;
  int32_t* range_flag;
  int32_t flags; /* TDN_TC_FLAG_* (0 = exact mode); FAST: Qhi.Khi^T and Phi.V'hi^T only */
} tdn_attention_desc;

int tdn_attention_tc(const tdn_attention_desc* desc, void* stream);

/* Number of kernels tdn_attention_tc(desc) launches on the current device (1, or 2 when the shared-memory-operand kernel
 * family runs the ragged last round as a second launch); same validation and decisions, nothing is launched. */
int tdn_attention_tc_launches(const tdn_attention_desc* desc, int32_t* launches);

/* fp32 plane <-> SPLIT16 planes (hi = fp16(x), lo = fp16(x - hi)); views must have equal dims. */
int tdn_split16(const tdn_tensor* in_f32, const tdn_tensor* out_split16, void* stream);
int tdn_merge16(const tdn_tensor* in_split16, const tdn_tensor* out_f32, void* stream);

/* NCHW fp32 image [n,3,H,W] (Testing/dataloader.py:69-71) -> NHWC fp32 with channels zero-padded to
 * out.c (4), the layout every later kernel reads.  Replaces nothing numeric; it is the layout edge. */
int tdn_image_to_nhwc(const float* nchw, int32_t n, int32_t c, int32_t h, int32_t w,
                      const tdn_tensor* out, void* stream);

/* Fused ResNet-18/34 stem (resnet.py:133-137,205-208): conv 7x7 stride 2 pad 3 (3->64) + folded BatchNorm +
 * ReLU + max_pool2d(3, 2, 1), from the caller's NCHW fp32 image [n,3,h,w] straight to the pooled NHWC map
 * `out` [n, Hp, Wp, 64] (F32 or SPLIT16).  weight is fp32 [147][64] with k = (c*7 + ky)*7 + kx. */
int tdn_stem_conv_pool(const float* nchw, int32_t n, int32_t h, int32_t w, const float* weight,
                       const float* scale, const float* bias, const tdn_tensor* out, void* stream);

/* Device-side frame ingest: the same fused stem reading the camera frame as uint8 HWC
 * [n,h,w,3] and normalising on the fly through a host-built table lut[3][256] with
 * lut[c][v] = (float)((v / 255.0 - mean[c]) / std[c]) evaluated in fp64 -- bit-identical to
 * Testing/dataloader.py:66-71 ((img/255.0 - mean)/std, then .float()); H2D shrinks from 12 to 3 bytes/pixel. */
int tdn_stem_conv_pool_u8(const uint8_t* hwc, const float* lut, int32_t n, int32_t h, int32_t w, const float* weight,
                          const float* scale, const float* bias, const tdn_tensor* out, void* stream);

/* The same fused stem on the tensor cores (tcgen05, exact mode: split-fp16 operands, fp32 accumulation), reading
 * either the fp32 NCHW image (`nchw`, hwc_u8 = lut = NULL) or the uint8 HWC frame with its table (`hwc_u8`, `lut`,
 * nchw = NULL).  weight_tc is fp16 [2 planes (hi, lo)][7 ky][4 kx pairs][64 cout][2 kx][4 c] holding
 * w[cout][c][ky][kx] * 2^e(cout) (zero for kx = 7 and c = 3); `scale` is the folded BatchNorm scale times
 * 2^-e(cout).  Image magnitudes must stay below 65504 (fp16 operand range); *range_flag is OR-ed with 1 otherwise
 * (may be NULL).  Returns TDN_ERR_ARCH on a device that is not sm_100. */
int tdn_stem_conv_pool_tc(const float* nchw, const uint8_t* hwc_u8, const float* lut, int32_t n, int32_t h, int32_t w,
                          const void* weight_tc, const float* scale, const float* bias, const tdn_tensor* out,
                          int32_t* range_flag, void* stream);

/* Same kernel with the activation chosen by the caller: TDN_ACT_RELU, or TDN_ACT_LEAKY_RELU for the ResNet of the
 * td2_fanet tree (Training/ptsemseg/models/td2_fanet/resnet.py:116-117, 136-138: conv1 -> norm_layer(64,
 * activation='leaky_relu') -> maxpool).  With LeakyReLU the pool pads with -inf (outputs may be negative). */
int tdn_stem_conv_pool_tc_act(const float* nchw, const uint8_t* hwc_u8, const float* lut, int32_t n, int32_t h,
                              int32_t w, const void* weight_tc, const float* scale, const float* bias,
                              const tdn_tensor* out, int32_t act, float leaky_slope, int32_t* range_flag,
                              void* stream);

/* F.max_pool2d(kernel 3, stride 2, padding 1) of the stem (resnet.py:137,208), NHWC fp32. */
int tdn_maxpool3x3s2(const tdn_tensor* in, const tdn_tensor* out, void* stream);

/* The four nn.AdaptiveAvgPool2d(1,2,3,6) of PyramidPooling (td4_psp18.py:249-252,273-276) in one
 * pass: out is [n, 1, 50, c] with bins ordered 1x1, 2x2, 3x3, 6x6 (row-major inside each grid);
 * bin i of size o covers rows [floor(i*H/o), ceil((i+1)*H/o)). */
int tdn_psp_pool(const tdn_tensor* in, const tdn_tensor* out, void* workspace, uint64_t workspace_bytes,
                 void* stream);
/* Workspace the call above needs: n*h*12*c floats (per row, the sums over the 12 column ranges). */
uint64_t tdn_psp_pool_workspace_bytes(int32_t n, int32_t h, int32_t c);

/* F.interpolate(mode='bilinear', align_corners=True) of a small NHWC map into a (channel-slice)
 * view of a larger one (td4_psp18.py:273-276 + the slice/cat of :278-284). */
int tdn_bilinear_nhwc(const tdn_tensor* in, const tdn_tensor* out, void* stream);

/* The four PSP branch convolutions on the pooled bins in one launch (td4_psp18.py:255-266: conv1x1 without
 * bias -> BatchNorm -> ReLU, here only this path's output-channel slice): pooled is the dense [n,1,50,c4]
 * output of tdn_psp_pool; w[i] fp32 [eighth][c4], scale[i]/bias[i] the folded BN, out[i] dense fp32
 * [n, bins_i, bins_i, eighth] for bins = 1, 2, 3, 6. */
int tdn_psp_branch_convs(const tdn_tensor* pooled, const float* const* w, const float* const* scale,
                         const float* const* bias, int32_t eighth, float* const* out, void* stream);

/* tdn_psp_branch_convs plus the projection of the pyramid features into the weight matrices of the 1x1 convolutions
 * that consume z = cat(c4 slice, up(b1), up(b2), up(b3), up(b6)) (td4_psp18.py:278-284 feeding Encoding.w_qs / w_ks /
 * w_vs, transformer.py:53-55): a 1x1 convolution commutes with the bilinear resize, so with the interpolation weights
 * of the 50 bins as 64 extra input channels of every pixel (constant per map size) neither the resized branch maps
 * nor z are ever written.  For projection q and image i the call stores, as SPLIT16,
 *     dst[i * batch_stride + o * ld + bin] = sum_c w[lv(bin) * eighth + c][o] * b_lv[i][bin][c]      (bin = 0..49)
 * i.e. column `bin` of the dynamic K block of the consumer's K-major weight matrix (columns 50..63 of that block stay
 * as the caller initialised them: zero).  `w` must carry the same per-row scale as the consumer's static weights.
 * range_flag (may be NULL) is set to 1 when a stored value exceeds the fp16 range of a SPLIT16 plane. */
#define TDN_PSP_MAX_PROJECTIONS 4
typedef struct tdn_psp_projection {
  const float* w;       /* fp32 [4 * eighth][cout]: input-channel major, so that the kernel's loads coalesce over o */
  void* dst_hi;         /* fp16 planes: element (row 0, first dynamic column) of image 0 */
  void* dst_lo;
  int64_t ld;           /* row pitch of dst, elements */
  int64_t batch_stride; /* per-image stride of dst, elements (ignored for n = 1) */
  int32_t cout;
  int32_t reserved;
} tdn_psp_projection;
int tdn_psp_branch_project(const tdn_tensor* pooled, const float* const* w, const float* const* scale,
                           const float* const* bias, int32_t eighth, float* const* out,
                           const tdn_psp_projection* proj, int32_t n_proj, int32_t* range_flag, void* stream);

/* The whole PyramidPooling output in one pass (td4_psp18.py:273-284): z = cat(x, up(b1), up(b2), up(b3), up(b6))
 * where x is the (already sliced) channel view of c4 and small[i] are the four branch maps, dense fp32
 * [n, bins_i, bins_i, eighth] with bins = 1, 2, 3, 6, resized bilinearly (align_corners=True) on the fly. */
int tdn_psp_concat(const tdn_tensor* x, const float* const* small, int32_t eighth, const tdn_tensor* z,
                   void* stream);

/* Strided copy between NHWC views with equal dims; the two views may differ in dtype, which makes
 * this the F32 <-> SPLIT16 converter as well (the x[:, pid*c/2:...] part of the cat in
 * td4_psp18.py:278-284, and the FIFO snapshots of buffer_contral :123-134). */
int tdn_copy_nhwc(const tdn_tensor* in, const tdn_tensor* out, void* stream);

/* In-place softmax over the last dim of a [rows, cols] fp32 matrix after multiplying by `scale`
 * (transformer.py:128-134: attn / temperature, nn.Softmax(dim=2)).  ld = row pitch in elements. */
int tdn_softmax_rows(float* s, int64_t rows, int32_t cols, int64_t ld, float scale, void* stream);

/* Same softmax with the probabilities written as SPLIT16 fp16 planes (row pitch ld_out >= cols, the
 * pad columns are zeroed) multiplied by out_scale (a power of two), ready to be the K-major A operand
 * of tdn_conv2d_tc for attn @ v. */
int tdn_softmax_rows_split16(const float* s, int64_t rows, int32_t cols, int64_t ld, float scale, void* p_hi,
                             void* p_lo, int64_t ld_out, float out_scale, void* stream);

/* Layer_Norm over the (H8,W8) map of every (n, channel) (td4_psp18.py:306-312, nn.LayerNorm([H8,W8]),
 * biased variance, eps 1e-5) in two steps: statistics (mean/rstd are [n, c]; fixed-order fp64
 * two-stage reduction, so results are bit-reproducible), then normalise + affine gamma/beta[h*w]. */
int tdn_layernorm_hw_stats(const tdn_tensor* x, float* mean, float* rstd, float eps, void* workspace,
                           uint64_t workspace_bytes, void* stream);
uint64_t tdn_layernorm_hw_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t c);
int tdn_layernorm_hw_apply(const tdn_tensor* x, const float* mean, const float* rstd,
                           const float* gamma, const float* beta, const tdn_tensor* out, void* stream);

/* The nclass classifier, a 1x1 convolution to a handful of channels (td4_psp18.py:299 `nn.Conv2d(inter, out, 1)`,
 * pspnet.py:113, td2_fa.py:316): out[p][j] = (sum_c in[p][c] * weight[j][c]) * scale[j] + bias[j] with
 * out.c <= 32; weight is fp32 [out.c][in.c], scale / bias fp32 [out.c] or NULL.  Channels are summed in index
 * order by one thread (bit-reproducible).  Same operator as tdn_conv2d with kh = kw = 1, without its tile waste. */
int tdn_pointwise_linear(const tdn_tensor* in, const float* weight, const float* scale, const float* bias,
                         const tdn_tensor* out, void* stream);

/* Final F.interpolate(output, (H, W), bilinear, align_corners=True) (td4_psp18.py:227): NHWC fp32
 * low-resolution logits -> NCHW fp32 [n, c, H, W], the tensor test.py:53,61 consumes. */
int tdn_upsample_logits(const tdn_tensor* in, float* out_nchw, int32_t out_h, int32_t out_w,
                        void* stream);

/* Same interpolation with the arg-max over classes fused in (Testing/test.py:61: output.max(1)[1]; lowest
 * index wins ties): NHWC fp32 low-resolution logits -> uint8 labels [n, H, W].  Bit-consistent with
 * tdn_upsample_logits (labels == argmax of the logits that call would write).  SURVEY.md 8f rank 1. */
int tdn_upsample_argmax(const tdn_tensor* in, uint8_t* labels, int32_t out_h, int32_t out_w, void* stream);

/* What Testing/test.py:61-64 keeps of a frame: `output.max(1)[1]` resized with cv2.INTER_NEAREST to (W/4, H/4).
 * Nearest resampling selects full-resolution pixels; labels[n][oy][ox] = arg-max over classes of the logits
 * interpolated (bilinear, align_corners=True, to full_h x full_w) at pixel (ys[oy], xs[ox]).  ys / xs are device
 * int32 tables of out_h / out_w full-resolution coordinates (OpenCV resizeNN: min(floor(d * src / dst), src - 1));
 * bit-consistent with tdn_upsample_argmax at those pixels, 1/16 of its work for the quarter-size map. */
int tdn_upsample_argmax_sampled(const tdn_tensor* in, uint8_t* labels, int32_t full_h, int32_t full_w,
                                const int32_t* ys, const int32_t* xs, int32_t out_h, int32_t out_w, void* stream);

/* cv2.resize(frame, (out_w, out_h)) of uint8 HWC RGB frames [n,h,w,3] -> [n,out_h,out_w,3] (Testing/dataloader.py:63;
 * OpenCV INTER_LINEAR, 8-bit fixed-point path), bit-exact.  x_taps / y_taps are device int32 tables with four entries
 * per output column / row: {source offset 0, source offset 1, weight 0, weight 1} (11-bit weights, 16-byte aligned);
 * OpenCV derives them in float / double on the host, so does the caller (tdnet_b200/ingest.py). */
int tdn_resize_linear_u8(const uint8_t* src, int32_t n, int32_t h, int32_t w, const int32_t* x_taps,
                         const int32_t* y_taps, uint8_t* dst, int32_t out_h, int32_t out_w, void* stream);

/* ------------------------------------------------------------------------------------------------
 * TD2-FANet widening: the non-convolution pieces of FAModule.forward
 * (Training/ptsemseg/models/td2_fanet/td2_fa.py:350-395).
 *
 * tdn_fa_context: f[n][32][c] = sum over pixels of normalize(key)[p][j] * value[p][c]  -- td2_fa.py:361-366
 *   (`F.normalize(key_, p=2, dim=1, eps=1e-12)`, `torch.matmul(key, value)`).  key is a [n,h,w,32] view, value
 *   [n,h,w,c]; f is dense fp32.  Two-stage fixed-order reduction (bit-reproducible); the caller provides
 *   tdn_fa_context_workspace_bytes(n,h,w,c) bytes of scratch.
 * tdn_fa_apply:  out[n][p][c] = out_scale * sum_j normalize(query)[p][j] * f[n][j][c]  -- td2_fa.py:358-359, 367-371
 *   (query normalised over its 32 channels, `torch.matmul(query, f)`, permuted back to a map).  out is F32 or
 *   SPLIT16; *range_flag (optional) is OR-ed with 1 if a SPLIT16 output exceeded the fp16 range guard.  The sum
 *   grows with the pixel count of the map: out_scale (a power of two, undone exactly by the caller in the scale of
 *   the convolution that reads `out`) keeps SPLIT16 outputs in range at any image size.
 * tdn_add_upsampled: out = bilinear_align_corners(up -> out size) + (a + b)            -- td2_fa.py:373, 398-402
 *   (`p_feat = W_y + feat`, then `_upsample_add`); up may be NULL (out = a + b).  a, b, out have equal dims,
 *   up has the same n and c.
 * ---------------------------------------------------------------------------------------------- */
int tdn_fa_context(const tdn_tensor* key, const tdn_tensor* value, float* f, void* workspace,
                   uint64_t workspace_bytes, void* stream);
uint64_t tdn_fa_context_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t c);
int tdn_fa_apply(const tdn_tensor* query, const float* f, const tdn_tensor* out, float out_scale, int32_t* range_flag,
                 void* stream);
int tdn_add_upsampled(const tdn_tensor* a, const tdn_tensor* b, const tdn_tensor* up, const tdn_tensor* out,
                      void* stream);

/* Diagnostics: the SM clock the chip actually runs at while other kernels are executing.  `blocks` CTAs of one warp (no
 * shared memory, so they co-reside with the persistent tcgen05 kernels) each spin for at least min_ns nanoseconds of
 * %globaltimer and write {elapsed %clock64 cycles, elapsed nanoseconds, SM id} to out[3 * block]: cycles / ns is that SM's
 * clock in GHz during the interval.  NVML's clock reading lags the power management by far more than a kernel's run time
 * (tools/clock_probe.py). */
int tdn_sm_clock_probe(uint64_t* out, int32_t blocks, int64_t min_ns, void* stream);

/* Library info / errors. */
int tdn_abi_version(void);
const char* tdn_strerror(int status);
const char* tdn_last_error(void);
/* Compute capability of the current device as major*10+minor, or a negative tdn_status. */
int tdn_device_arch(void);

#ifdef __cplusplus
}
#endif
#endif /* TDNET_B200_H_ */
