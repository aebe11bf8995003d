/*
 * cmr_b200 -- C ABI of the B200-native Mask R-CNN (R50/R101-C4) hot path.
 *
 * Plain C entry points: pointers, sizes, a CUDA stream handle, int status.
 * No torch / C++ types cross this boundary.  All device entry points are
 * asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy
 * default stream), re-entrant, and never allocate or free user-visible
 * memory: the caller owns inputs, outputs and workspaces.
 *
 * Every function cites the reference interface (file:line under
 * wkentaro/chainer-mask-rcnn @ v0.5.24) that it replaces.  INTEGRATION.md
 * shows the ctypes binding a maintainer of the reference would add.
 */
#ifndef CMR_B200_H_
#define CMR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum cmr_status {
  CMR_OK = 0,
  CMR_ERR_INVALID_ARG = -1,  /* bad shape / null pointer / unsupported value  */
  CMR_ERR_CUDA = -2,         /* a CUDA runtime / driver call failed           */
  CMR_ERR_WORKSPACE = -3,    /* caller's workspace is too small               */
  CMR_ERR_UNSUPPORTED = -4   /* shape outside what the kernel was built for   */
} cmr_status;

/* Human-readable text for a status code (static storage). */
const char* cmr_status_string(int status);
/* Library ABI version (bumped when a signature changes). */
int cmr_version(void);
/* Last CUDA error string recorded by this library on the calling thread. */
const char* cmr_last_cuda_error(void);
/* Number of CUDA kernels this library has launched in this process so far. */
long long cmr_launch_count(void);
/* Measurement aid for bench.py: when enabled, every tensor-core and ROIAlign launch is
 * bracketed by a pair of CUDA events on its stream (not while the stream is captured).
 * cmr_prof_collect(kind) waits for the recorded launches of `kind`, returns their summed
 * device time (ms), summed algorithmic work and count, and forgets them.  Kinds:
 *   0 cmr_conv_gemm_tc[_ex]   work = FLOPs, 2*M*N*K per launch without padding
 *   1 cmr_conv_wgrad_tc       work = FLOPs
 *   2 cmr_roi_align_nhwc_fwd  work = bytes, 4*(R*C*oh*ow + N*C*H*W + 5R)
 *   3 cmr_roi_align_nhwc_bwd  work = bytes, same formula
 *   6 cmr_roi_align_fwd_cl    work = bytes, same formula (the drop-in operator's kernels)
 *   7 cmr_roi_align_bwd_cl    work = bytes, same formula (zero fill of gx included)
 *   4 / 5 the kind-0 launches again, split by what bounds them: 4 = tensor-bound (FLOPs
 *     >= 110 x algorithmic bytes; work = FLOPs), 5 = HBM-bound (work = algorithmic bytes:
 *     activations, filter, output, residual and mask operands once each).  Collect 4 and
 *     5 before 0. */
int cmr_prof_enable(int on);
int cmr_prof_collect(int kind, double* total_ms, double* total_work,
                     long long* launches);

/* ------------------------------------------------------------------------ *
 * ROIAlign, reference layout (x NCHW fp32, rois (R,5) = b,x1,y1,x2,y2).
 * Replaces ROIAlign2D.forward_gpu / backward_gpu
 * (chainer_mask_rcnn/functions/roi_align_2d.py:162-290, 391-524).
 * y: (R,C,outh,outw).  gx: (N,C,H,W), zero-filled by the callee.
 * sampling_ratio == 0 selects the adaptive grid ceil(roi/pooled).
 * ------------------------------------------------------------------------ */
int cmr_roi_align_fwd(const float* x, int N, int C, int H, int W,
                      const float* rois, int R, int outh, int outw,
                      float spatial_scale, int sampling_ratio, float* y,
                      void* stream);
int cmr_roi_align_bwd(const float* gy, const float* rois, int R, int N, int C,
                      int H, int W, int outh, int outw, float spatial_scale,
                      int sampling_ratio, float* gx, void* stream);

/* cmr_roi_align_nhwc_bwd without the zero fill: the RoI gradients are ADDED to what gx holds
 * (the train step starts gx with the RPN branch's gradient of the same feature map, which
 * was computed earlier on another stream). */
int cmr_roi_align_nhwc_bwd_accum(const float* gy, const float* rois, int R, int N,
                                 int H, int W, int C, int outh, int outw,
                                 int bin_stride, float spatial_scale,
                                 int sampling_ratio, float* gx, void* stream);

/* The same operator at speed: the feature map is read channels-last (vector loads of 4
 * channels, whole 128-byte lines per warp), the pooled tensor stays in the reference's
 * (R,C,outh,outw) layout -- each CTA stages its (64 channels x outh*outw) block, which is
 * contiguous in that layout, in shared memory and moves it as full lines.  Same products as
 * the reference, summed separably (a few ulp apart; cmr_roi_align_fwd / _bwd above keep the
 * reference's summation order).
 *   _cl  : x_nhwc (N,H,W,C) / gx_nhwc (N,H,W,C, zero-filled by the callee), C % 4 == 0;
 *          CMR_ERR_UNSUPPORTED when cmr_roi_align_cl_supported(...) == 0.
 *   _ws  : reference-layout x / gx plus a caller-owned workspace of
 *          cmr_roi_align_workspace_bytes(N,C,H,W) bytes for the re-laid map (2 % of the
 *          pooled tensor's bytes at R = 1000); falls back to cmr_roi_align_fwd / _bwd when
 *          the workspace is NULL / too small or the shape is not supported.
 * A RoI whose batch index is outside [0, N) produces zeros / no gradient in every variant. */
int cmr_roi_align_cl_supported(int N, int C, int H, int W, int R, int outh, int outw);
int cmr_roi_align_fwd_cl(const float* x_nhwc, int N, int H, int W, int C,
                         const float* rois, int R, int outh, int outw,
                         float spatial_scale, int sampling_ratio, float* y,
                         void* stream);
int cmr_roi_align_bwd_cl(const float* gy, const float* rois, int R, int N, int H,
                         int W, int C, int outh, int outw, float spatial_scale,
                         int sampling_ratio, float* gx_nhwc, void* stream);
size_t cmr_roi_align_workspace_bytes(int N, int C, int H, int W);
int cmr_roi_align_fwd_ws(const float* x, int N, int C, int H, int W,
                         const float* rois, int R, int outh, int outw,
                         float spatial_scale, int sampling_ratio, float* y,
                         void* workspace, size_t workspace_bytes, void* stream);
int cmr_roi_align_bwd_ws(const float* gy, const float* rois, int R, int N, int C,
                         int H, int W, int outh, int outw, float spatial_scale,
                         int sampling_ratio, float* gx, void* workspace,
                         size_t workspace_bytes, void* stream);
/* (batch, rows, cols) -> (batch, cols, rows), fp32: NCHW <-> NHWC re-layout
 * (rows = C, cols = H*W or the reverse). */
int cmr_transpose_batched(const float* in, int batch, int rows, int cols, float* out,
                          void* stream);

/* Same operator on channels-last tensors, used inside the model
 * (ResNetRoIHead.__call__, models/mask_rcnn_resnet.py:168-181):
 * x (N,H,W,C), y (R,outh/bin_stride,outw/bin_stride,C).  Only the bins
 * (ph, pw) with ph % bin_stride == 0 and pw % bin_stride == 0 are produced:
 * res5.a's stride-2 1x1 convolutions read no others when roi_size == 14.
 * C must be a multiple of 4.  round_tf32 != 0 rounds y to tf32 (it is the A operand
 * of res5.a's tensor-core GEMMs). */
int cmr_roi_align_nhwc_fwd(const float* x, int N, int H, int W, int C,
                           const float* rois, int R, int outh, int outw,
                           int bin_stride, float spatial_scale,
                           int sampling_ratio, int round_tf32, float* y,
                           void* stream);
int cmr_roi_align_nhwc_bwd(const float* gy, const float* rois, int R, int N,
                           int H, int W, int C, int outh, int outw,
                           int bin_stride, float spatial_scale,
                           int sampling_ratio, float* gx, void* stream);

/* ------------------------------------------------------------------------ *
 * AffineChannel2D as a stand-alone operator, reference layout (x (N,C,H,W) fp32
 * contiguous, HW = H*W; W, b (C,)).  Replaces AffineChannel2DFunction.forward /
 * backward (chainer_mask_rcnn/functions/affine_channel_2d.py:17-20, 48-55):
 *   fwd: y = W[c] * x + b[c]      (a multiplication and an addition, as NumPy rounds)
 *   bwd: gx = W[c] * gy;  gW[c] = sum over (n,h,w) of x * gy;  gb[c] = sum of gy
 * (inside the model the affine is the convolution kernels' epilogue instead).  The
 * backward's two-stage reduction is deterministic; workspace =
 * cmr_affine_channel_bwd_workspace_bytes(N, C) bytes.
 * ------------------------------------------------------------------------ */
int cmr_affine_channel_fwd(const float* x, const float* W, const float* b, int N, int C,
                           int HW, float* y, void* stream);
size_t cmr_affine_channel_bwd_workspace_bytes(int N, int C);
int cmr_affine_channel_bwd(const float* x, const float* W, const float* gy, int N, int C,
                           int HW, float* gx, float* gW, float* gb, void* workspace,
                           size_t workspace_bytes, void* stream);
/* BatchNormalization -> AffineChannel2D (_get_affine_from_bn,
 * chainer_mask_rcnn/models/resnet_extractor.py:16-29): W = gamma / sqrt(var + eps),
 * b = beta - mean * W, each step one IEEE fp32 operation (bit-exact with NumPy). */
int cmr_bn_fold(const float* gamma, const float* beta, const float* mean,
                const float* var, float eps, int C, float* W, float* b, void* stream);

/* ------------------------------------------------------------------------ *
 * Greedy NMS on score-sorted boxes (y1,x1,y2,x2), IoU >= thresh suppresses.
 * Replaces chainercv.utils.non_maximum_suppression as called at
 * chainer_mask_rcnn/models/mask_rcnn.py:193-194 and inside ProposalCreator
 * (models/region_proposal_network.py:136-138).
 * keep: int32[n] (first *n_keep entries valid, ascending = score order);
 * n_keep: device int32.  limit <= 0 means no limit.  The workspace also
 * receives the 64-bit suppression bitmask (n x ceil(n/64) words, row i bit j
 * set iff j > i and IoU(i,j) >= thresh) at offset 0 for the bit-exact test.
 * ------------------------------------------------------------------------ */
size_t cmr_nms_workspace_bytes(int n);
int cmr_nms(const float* boxes, int n, float thresh, int limit, int32_t* keep,
            int32_t* n_keep, void* workspace, size_t workspace_bytes,
            void* stream);

/* ------------------------------------------------------------------------ *
 * RPN proposal generation for a batch of B images: decode (loc2bbox), clip,
 * min-size filter, descending score sort, top n_pre, NMS, top n_post.
 * Replaces chainercv ProposalCreator.__call__ as invoked per image at
 * chainer_mask_rcnn/models/region_proposal_network.py:135-141.
 * loc (B,n_anchor,4), score (B,n_anchor), anchor (n_anchor,4) fp32.
 * rois_out (B,n_post,4) fp32, idx_out (B,n_post) int32 = anchor index of each
 * kept proposal, n_out (B) int32 (device).
 * ------------------------------------------------------------------------ */
size_t cmr_proposals_workspace_bytes(int B, int n_anchor, int n_pre);
int cmr_proposals(const float* loc, const float* score, const float* anchor,
                  int B, int n_anchor, float img_h, float img_w,
                  float min_size, int n_pre, int n_post, float nms_thresh,
                  float* rois_out, int32_t* idx_out, int32_t* n_out,
                  void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------ *
 * Implicit-GEMM convolution on the tcgen05 tensor cores (TF32 inputs, fp32
 * accumulate):  D[m,n] = epilogue(sum_k A[m,k] * W[n,k]).
 * Replaces the forward / data-gradient of chainer.links.Convolution2D,
 * Deconvolution2D and Linear as instantiated at
 * chainer_mask_rcnn/models/region_proposal_network.py:75-80,
 * models/mask_rcnn_resnet.py:131-143 and inside BuildingBlock
 * (models/resnet_extractor.py:47-90), with AffineChannel2D
 * (functions/affine_channel_2d.py:17-20), bias, residual add and ReLU fused.
 *
 * a : NHWC activations (batch, in_h, in_w, in_ld) of which in_c channels are
 *     used (in_c % 32 == 0; in_ld < in_c is allowed and makes one "channel" run
 *     span several pixels of a row -- the stem reads 8 RGB0 pixels = 32 floats per
 *     filter row this way); rows of the GEMM are output pixels (b, oy, ox),
 *     oy < out_h, ox < out_w, reading input pixel (oy*stride - pad + fr,
 *     ox*stride - pad + fs); out-of-image taps read zeros.
 * w : filter bank (n, kh, kw, in_c) fp32, i.e. K-major.
 * d : row (b, oy, ox) is written at pixel (oy*d_stride + d_oy, ox*d_stride + d_ox)
 *     of a (batch, d_h, d_w, d_ld) tensor, channels [0, n).
 * epilogue, in this order, each optional (NULL / 0 = skip):
 *     v *= scale[n];  v += bias[n];  v += addend[same address as d];
 *     v = max(v, 0) if relu;  v = 0 where mask[same address as d] <= 0;
 *     v = round-to-nearest tf32 if round_tf32.
 * tile_n: 0 = automatic, else 64 / 128 / 256.
 * ------------------------------------------------------------------------ */
typedef struct cmr_conv_desc {
  int batch, in_h, in_w, in_c, in_ld;
  int out_h, out_w;
  int kh, kw, stride, pad;
  int n;
  int d_h, d_w, d_ld, d_stride, d_oy, d_ox;
  int relu, round_tf32;
  int tile_n;
  int tap_cols;   /* > 0: fused Deconvolution2D(2, stride 2): n == 4*tap_cols, d_stride == 2;
                     GEMM column block t = column / tap_cols is filter tap (t >> 1, t & 1) and
                     is written to pixel (2*oy + d_oy + (t >> 1), 2*ox + d_ox + (t & 1)),
                     channels [0, tap_cols); scale / bias are indexed by the channel.
                     0 = plain layout. */
} cmr_conv_desc;

/* The activation operand of cmr_conv_gemm_tc / cmr_conv_wgrad_tc is fetched by im2col-mode
 * TMA whenever a tensor map can describe the layout (everything but the RGB0-packed
 * stem); 0 forces the cp.async gather path everywhere (kept for A/B measurements and
 * tests).  Process-wide; returns the previous setting. */
int cmr_set_im2col_tma(int on);

/* Measurement knob: tile-configuration override of cmr_conv_gemm_tc for same-box A/B runs of
 * single layers (tools/conv_shape_bench.py).  0 = the dispatch described above (default);
 * 1 = CTA pairs also for short reductions (5 stages, 3 epilogue groups); 2 = the long-reduction
 * configuration (pairs, 6 stages, 2 groups) for every 256-wide launch; 3 = single CTAs with
 * 4 stages and 2 groups; + 16 = the epilogue reads no addend / mask; + 32 = the epilogue does not
 * store (probes: the results are wrong).  Process-wide; returns the previous setting. */
int cmr_set_conv_variant(int variant);

/* Measurement knob: when `buf` (device memory, 8 int64 per CTA of the launch, i.e. 8 * 148
 * words) is non-NULL every following cmr_conv_gemm_tc launch writes, per CTA, the SM cycles its
 * roles spent waiting: [0] MMA issuer for a free accumulator, [1] MMA issuer for operand
 * stages, [2] MMA issuer total, [3] TMA producer for free stages, [4] first epilogue warp for
 * finished accumulators, [5] first epilogue warp total, [6] tiles.  NULL switches it off. */
int cmr_set_conv_debug(long long* buf);

int cmr_conv_gemm_tc(const cmr_conv_desc* desc, const float* a, const float* w,
                     float* d, const float* scale, const float* bias,
                     const float* addend, const float* mask, void* stream);

/* Same, with one more epilogue term between `addend` and `relu`:
 *     v += bcast[(row / bcast_group) * n + column] * bcast_scale
 * where row = (b, oy, ox) is the GEMM row.  With bcast_group = oh*ow and bcast_scale =
 * 1/(oh*ow) this is the backward of average_pooling_2d over the whole window
 * (chainer_mask_rcnn/models/mask_rcnn_resnet.py:187) fused into the data-gradient GEMM that
 * produces the other gradient flowing into the same tensor.  bcast == NULL disables it;
 * n % 4 == 0 and 16-byte alignment are required otherwise. */
int cmr_conv_gemm_tc_ex(const cmr_conv_desc* desc, const float* a, const float* w,
                        float* d, const float* scale, const float* bias,
                        const float* addend, const float* mask, const float* bcast,
                        int bcast_group, float bcast_scale, void* stream);

/* Same, with a caller-owned workspace that enables the K-split tail: when the launch's tiles
 * leave a last wave that fills at most half of the SM pairs (e.g. res5's 392 pair tiles on 74
 * SM pairs: 5.3 waves; CTA-pair launches with K >= 1024 only), every tile of that wave is
 * computed by up to 4 CTA pairs, one K range each; the partial sums go through `ws` and a small second kernel adds them in part order (a
 * fixed sum: results do not depend on timing) and runs the epilogue.  `ws`: device memory of
 * at least cmr_conv_gemm_ws_bytes() bytes, 16-byte aligned, used by one launch at a time --
 * one workspace per stream.  ws == NULL (or too small, or a launch the split does not apply
 * to): exactly cmr_conv_gemm_tc_ex.  (The Python engine passes a workspace only with
 * CMR_CONV_SPLIT_TAIL=1: the split shortens its launches by 1-4 % when they run alone but
 * measured 0.5 % slower inside the train step, whose side-stream weight gradients already
 * fill those tails -- DESIGN.md section 3.) */
size_t cmr_conv_gemm_ws_bytes(void);
int cmr_conv_gemm_tc_ws(const cmr_conv_desc* desc, const float* a, const float* w,
                        float* d, const float* scale, const float* bias,
                        const float* addend, const float* mask, const float* bcast,
                        int bcast_group, float bcast_scale, void* ws, size_t ws_bytes,
                        void* stream);

/* ------------------------------------------------------------------------ *
 * Weight gradient on the tcgen05 tensor cores (TF32 inputs, fp32 accumulate):
 *   gw[i, gw_col0 + j] += row_scale[i] * sum over pixels (b, oy, ox) of
 *        gy[b, oy*gy_stride + gy_off_y, ox*gy_stride + gy_off_x, gy_c0 + i]
 *      *  x[b, oy*x_stride  + x_off_y,  ox*x_stride  + x_off_x,  x_c0 + j]
 * for i < rows, j < cols, (oy, ox) in loop_h x loop_w; out-of-tensor pixels read
 * zeros.  One call covers one filter tap: for a (n, kh, kw, c) filter bank call it
 * with x_off = (fr - pad, fs - pad), gw_col0 = (fr*kw + fs)*c, gw_ld = kh*kw*c.
 * Replaces the gW of chainer's Convolution2D / Deconvolution2D / Linear backward
 * for the links at chainer_mask_rcnn/models/region_proposal_network.py:75-80 and
 * models/mask_rcnn_resnet.py:131-143.  gw is accumulated into (zero it first).
 * splits: 0 = automatic split of the pixel reduction across CTAs.
 * ------------------------------------------------------------------------ */
typedef struct cmr_wgrad_desc {
  int batch, loop_h, loop_w;
  int gy_h, gy_w, gy_ld, gy_stride, gy_off_y, gy_off_x, gy_c0;
  int x_h, x_w, x_ld, x_stride, x_off_y, x_off_x, x_c0;
  int rows, cols;
  int gw_ld, gw_col0;
  int splits;
  int taps_h, taps_w;   /* > 1: all filter taps in one launch; tap (fr, fs) reads x at
                           x_off + (fr, fs) and accumulates into column block
                           gw_col0 + (fr*taps_w + fs)*cols.  0 or 1 = a single tap. */
} cmr_wgrad_desc;

int cmr_conv_wgrad_tc(const cmr_wgrad_desc* desc, const float* gy, const float* x,
                      float* gw, const float* row_scale, void* stream);

/* ------------------------------------------------------------------------ *
 * Deterministic mode.  The three reductions of the train step whose arrival order is not
 * fixed -- split weight gradients, the ROIAlign backward scatter, bias column sums -- have
 * `_fixed` forms that accumulate into 64-bit fixed-point words (one unit = 2^-40) instead
 * of fp32: integer addition is associative, so the replayed step is bit-reproducible.  The
 * words use the element order of the fp32 tensor they stand for (same shapes, strides and
 * offsets, 8 bytes per element) and must be zero before the first accumulation.
 * cmr_fixed_to_float converts them back: out = (accumulate ? out : 0) + in * 2^-40, zeroing
 * `in` for the next step when zero_src != 0.
 * ------------------------------------------------------------------------ */
int cmr_conv_wgrad_tc_fixed(const cmr_wgrad_desc* desc, const float* gy, const float* x,
                            long long* gw_fixed, const float* row_scale, void* stream);
int cmr_col_sum_fixed(const float* g, long long M, int ld, int c0, int n,
                      long long* out_fixed, void* stream);   /* zeroes out_fixed[0..n) first */
int cmr_roi_align_nhwc_bwd_fixed(const float* gy, const float* rois, int R, int N,
                                 int H, int W, int C, int outh, int outw,
                                 int bin_stride, float spatial_scale,
                                 int sampling_ratio, long long* gx_fixed,
                                 void* stream);               /* adds to gx_fixed (no zero fill) */
int cmr_fixed_to_float(long long* in, float* out, size_t n, int accumulate,
                       int zero_src, void* stream);

/* out[i] = round-to-nearest-tf32(in[i]) (in == out allowed). */
int cmr_round_tf32(const float* in, float* out, size_t n, void* stream);

/* out[i] = mask[i] > 0 ? g[i] : 0 (the backward of F.relu as its own pass; out == g
 * allowed), rounded to tf32 when round_tf32 != 0.  n % 4 == 0, 16-byte aligned. */
int cmr_relu_mask(const float* g, const float* mask, float* out, size_t n,
                  int round_tf32, void* stream);

/* Operand split of the 3 x TF32 parity mode: every fp32 value is written as
 * hi = tf32(x) and lo = tf32(x - hi), x = hi + lo up to 2^-22 |x|.  x (rows, c_in) ->
 * out (rows, 3 * c_pad): three column blocks of c_pad >= c_in channels (zero padded),
 * order 0 = [hi | lo | hi] (activations), order 1 = [hi | hi | lo] (filters).  A GEMM over
 * the concatenated K axis, sum_k a3[k] * w3[k] = sum_c hi_a*hi_w + lo_a*hi_w + hi_a*lo_w,
 * is the fp32 product up to the dropped lo*lo term: the tensor-core convolution kernel run
 * on split operands reproduces an fp32 sgemm to ~1e-6 (three times the work; used to verify
 * the chained model against the fp32 reference, not for speed). */
int cmr_split_tf32x3(const float* x, size_t rows, int c_in, int c_pad, int order,
                     float* out, void* stream);

/* ------------------------------------------------------------------------ *
 * HBM-bound pieces of the graph around the convolutions (csrc/misc.cu).
 * ------------------------------------------------------------------------ */
/* (B,3,H,W) fp32 image planes -> (B,Hp,Wp,4) zero-padded RGB0 pixels, tf32-rounded;
 * the image sits at (pad_top, pad_left).  Input side of extractor.conv1
 * (chainer_mask_rcnn/models/resnet_extractor.py:63-66). */
int cmr_pack_image_nhwc4(const float* img, int B, int H, int W, int Hp, int Wp,
                         int pad_top, int pad_left, float* out, void* stream);
/* chainer.functions.max_pooling_2d(ksize, stride, pad) on NHWC (cover_all is
 * expressed through out_h/out_w); models/resnet_extractor.py:67-69. */
int cmr_max_pool_nhwc(const float* x, int B, int H, int W, int C, int ksize,
                      int stride, int pad, int out_h, int out_w, float* y,
                      void* stream);
/* average_pooling_2d over the whole HW window (models/mask_rcnn_resnet.py:187):
 * x (R,HW,C) -> y (R,C); backward adds g/HW to `out` (R,HW,C), then zeroes where
 * mask <= 0 (mask may be NULL). */
int cmr_avg_pool_nhwc_fwd(const float* x, int R, int HW, int C, float* y,
                          int round_tf32, void* stream);
int cmr_avg_pool_nhwc_bwd_accum(const float* g, int R, int HW, int C, float* out,
                                const float* mask, int round_tf32, void* stream);
/* out[j] = sum_m g[m*ld + c0 + j], j < n  (bias gradients). */
int cmr_col_sum(const float* g, long long M, int ld, int c0, int n, float* out,
                void* stream);
/* Filter bank of the data-gradient GEMM (row length ld_out >= col0 + O, so that
 * several layers can share one fused bank):
 * out[(i*T + (flip ? T-1-t : t))*ld_out + col0 + o] =
 *     tf32(w[o*stride_o + t*stride_t + i] * scale[o]);  scale may be NULL. */
int cmr_prep_dgrad_weight(const float* w, int O, int T, int I, long long stride_o,
                          long long stride_t, const float* scale, int flip,
                          float* out, int ld_out, int col0, void* stream);
/* The same re-layout for many layers in ONE launch.  descs_dev: DEVICE array of n_desc
 * descriptors whose tile_begin fields are the running sum of
 * ceil(I/32) * ceil(O/32) * T over the preceding descriptors (first = 0);
 * total_tiles = that sum over all of them. */
typedef struct cmr_prep_desc {
  const float* w;
  const float* scale;   /* may be NULL */
  float* out;
  long long stride_o, stride_t;
  int O, T, I, flip, ld_out, col0;
  int tile_begin;
  int reserved;
} cmr_prep_desc;
int cmr_prep_dgrad_weight_batch(const cmr_prep_desc* descs_dev, int n_desc,
                                int total_tiles, void* stream);
/* MomentumSGD + WeightDecay (examples/train_common.py:176-180) on a flat buffer:
 * g' = grad_scale*g + wd*p;  v = momentum*v - lr*g';  p += v.  n % 4 == 0.
 * param_tf32 (may be NULL): receives round-to-nearest-tf32(p), the copy the next
 * forward pass's GEMMs read. */
int cmr_sgd_momentum(float* param, const float* grad, float* velocity, size_t n,
                     float lr, float momentum, float weight_decay,
                     float grad_scale, float* param_tf32, void* stream);

/* ------------------------------------------------------------------------ *
 * Losses of MaskRCNNTrainChain.__call__ and their gradients
 * (chainer_mask_rcnn/models/mask_rcnn_train_chain.py:160-213), csrc/loss.cu.
 * losses: device float[8]; [0] rpn_loc [1] rpn_cls [2] roi_loc [3] roi_cls
 * [4] roi_mask, [5..7] scratch.  Gradient buffers are fully written.
 * ------------------------------------------------------------------------ */
/* loc (n_pixel, ld_loc) holds 4*A values per pixel, score (n_pixel, ld_score) A;
 * gt_loc (n_pixel*A, 4), gt_label (n_pixel*A) in {-1,0,1};
 * g (n_pixel, ld_g): [0,4A) d/dloc, [4A,5A) d/dscore, [5A,ld_g) zeros. */
int cmr_rpn_loss(const float* loc, int ld_loc, const float* score, int ld_score,
                 const float* gt_loc, const int32_t* gt_label, long long n_pixel,
                 int n_anchor, float sigma, float* g, int ld_g, float* losses,
                 void* stream);
/* cls_loc (R, ld_cls_loc) 4*n_class values, score (R, ld_score) n_class logits;
 * g (R, ld_g): [0,4*n_class) d/dcls_loc, [4*n_class,5*n_class) d/dscore, rest 0. */
int cmr_roi_loss(const float* cls_loc, int ld_cls_loc, const float* score,
                 int ld_score, const float* gt_loc, const int32_t* gt_label, int R,
                 int n_class, float sigma, float* g, int ld_g, float* losses,
                 void* stream);
/* masks (R,HW,ld_masks) logits in channels [0,n_fg), gt_mask (R,HW) in {-1,0,1};
 * g (R,HW,ld_g), ld_g % 4 == 0, 16-byte aligned: channel label-1 of each RoI gets the
 * gradient, the rest zeros. */
int cmr_mask_loss(const float* masks, int ld_masks, const int32_t* gt_label,
                  const int32_t* gt_mask, int R, int HW, int n_fg, float* g,
                  int ld_g, float* losses, void* stream);

/* ------------------------------------------------------------------------ *
 * Training targets on the device (csrc/targets.cu).  bbox (B,max_bbox,4) holds the
 * ground-truth boxes (y1,x1,y2,x2) of each image padded to max_bbox <= 256 rows,
 * n_bbox (B) their counts (device int32).  `seed` drives the random subsampling;
 * when seed_dev (device uint64, may be NULL) is given, *seed_dev is mixed in, so a
 * captured CUDA graph draws a new subset on every replay by advancing that word.
 * ------------------------------------------------------------------------ */
/* chainercv AnchorTargetCreator, called per image at
 * chainer_mask_rcnn/models/mask_rcnn_train_chain.py:151-158, for the whole batch:
 * gt_loc (B,n_anchor,4), gt_label (B,n_anchor) in {-1 ignore, 0, 1}. */
size_t cmr_anchor_targets_workspace_bytes(int B, int n_anchor, int max_bbox);
int cmr_anchor_targets(const float* anchor, int n_anchor, const float* bbox,
                       const int32_t* n_bbox, int B, int max_bbox, float img_h,
                       float img_w, int n_sample, float pos_iou_thresh,
                       float neg_iou_thresh, float pos_ratio,
                       unsigned long long seed, const unsigned long long* seed_dev,
                       float* gt_loc, int32_t* gt_label, void* workspace,
                       size_t workspace_bytes, void* stream);
/* ProposalTargetCreator.__call__ up to the mask targets
 * (chainer_mask_rcnn/models/utils/proposal_target_creator.py:115-161):
 * rois (B,max_roi,4) + n_roi (B) as produced by cmr_proposals; label (B,max_bbox)
 * foreground class ids; loc_mean / loc_std: HOST float[4].
 * Outputs, n_sample rows per image, foreground rows first: sample_roi, gt_roi_loc
 * (normalised), gt_roi_label (class+1, 0 = background, -1 = padding row when an
 * image has too few candidates), gt_assign (ground-truth index of each foreground
 * row, else -1), n_pos (B) foreground rows per image. */
size_t cmr_proposal_targets_workspace_bytes(int B, int max_roi, int max_bbox);
int cmr_proposal_targets(const float* rois, const int32_t* n_roi, int max_roi,
                         const float* bbox, const int32_t* label,
                         const int32_t* n_bbox, int B, int max_bbox, int n_sample,
                         float pos_ratio, float pos_iou_thresh,
                         float neg_iou_thresh_hi, float neg_iou_thresh_lo,
                         const float* loc_mean, const float* loc_std,
                         unsigned long long seed,
                         const unsigned long long* seed_dev, float* sample_roi,
                         float* gt_roi_loc, int32_t* gt_roi_label,
                         int32_t* gt_assign, int32_t* n_pos, void* workspace,
                         size_t workspace_bytes, void* stream);

/* The mask rasterisation of ProposalTargetCreator.__call__
 * (chainer_mask_rcnn/models/utils/proposal_target_creator.py:163-177) for the rows
 * cmr_proposal_targets produced: for every foreground row j < n_pos[b], round the RoI
 * to integers, crop instance mask gt_assign[b,j], and take argmax over the one-hot
 * planes of cv2.resize(INTER_LINEAR, float32) to (mask_size, mask_size); all other
 * rows are filled with -1.  masks: (B, max_bbox, H, W) device array of uint8
 * (mask_elem_bytes == 1) or int32 (== 4) labels >= 0, or (mask_elem_bytes == 0) binary
 * masks packed one bit per pixel: (B, max_bbox, H, ceil(W/8)) bytes, pixel x = bit x & 7
 * of byte x >> 3 (numpy.packbits(..., axis=-1, bitorder='little')).  gt_mask: (B,
 * n_sample, mask_size, mask_size) int32.  An empty crop yields zeros. */
int cmr_mask_targets(const void* masks, int mask_elem_bytes, int B, int max_bbox,
                     int H, int W, const float* sample_roi, const int32_t* gt_assign,
                     const int32_t* n_pos, int n_sample, int mask_size,
                     int32_t* gt_mask, void* stream);

/* ------------------------------------------------------------------------ *
 * Inference post-processing (csrc/detect.cu).
 * ------------------------------------------------------------------------ */
/* MaskRCNN._to_bboxes + _suppress (chainer_mask_rcnn/models/mask_rcnn.py:178-243) for a
 * batch: softmax of the class logits, boxes of every (RoI, class >= 1) pair whose
 * probability exceeds score_thresh decoded against roi / scale (loc * std + mean,
 * loc2bbox) and clipped to the image, then per-class NMS -- every (image, class) pair is
 * one row of a single batched NMS launch.
 * cls_loc (B*max_roi, ld_loc) holds 4*n_class offsets per RoI, score (B*max_roi,
 * ld_score) n_class logits; rois (B, max_roi, 4) with n_roi (B) valid rows (device);
 * img_info (B, 3) device floats = (scale, height, width) of each original image;
 * loc_mean / loc_std: HOST double[4].
 * Outputs per image in the reference's order -- class by class, descending score inside
 * a class (rows >= n_det[b] are padding with label -1): det_bbox (B, max_cand, 4),
 * det_label (B, max_cand) foreground class ids (class - 1), det_score (B, max_cand),
 * n_det (B).  max_cand is the row capacity of the outputs: max_roi * (n_class - 1) always
 * suffices (19 * max_roi when score_thresh >= 0.05); survivors beyond it are dropped. */
size_t cmr_detections_workspace_bytes(int B, int max_roi, int n_class, int max_cand);
int cmr_detections(const float* cls_loc, int ld_loc, const float* score, int ld_score,
                   const float* rois, const int32_t* n_roi, int B, int max_roi,
                   int n_class, const float* img_info, const double* loc_mean,
                   const double* loc_std, float score_thresh, float nms_thresh,
                   int max_cand, float* det_bbox, int32_t* det_label, float* det_score,
                   int32_t* n_det, void* workspace, size_t workspace_bytes,
                   void* stream);
/* segm_results (chainer_mask_rcnn/models/mask_rcnn.py:63-107): for detection k, the
 * (mask_size x mask_size) map of class label[k] in mask_prob -- element (k, c, y, x) at
 * mask_prob[k*stride_n + c*stride_c + y*stride_y + x*stride_x], so NCHW and channels-last
 * tensors both work; logits when apply_sigmoid != 0, else probabilities -- is zero-padded by
 * one pixel, resized (cv2 INTER_LINEAR, fp32) to the integer box grown by
 * (mask_size+2)/mask_size, thresholded at 0.5 and written into out (n, H, W) uint8
 * (fully written: zeros elsewhere).  bbox (n, 4) = (y1, x1, y2, x2). */
int cmr_paste_masks(const float* bbox, const int32_t* label, const float* mask_prob,
                    long long stride_n, long long stride_c, long long stride_y,
                    long long stride_x, int n, int mask_size, int H, int W,
                    int apply_sigmoid, uint8_t* out, void* stream);

/* MaskRCNN.prepare (chainer_mask_rcnn/models/mask_rcnn.py:152-176) on the device.
 * cmr_prepare_size: the size cv2.resize(img, None, fx=fx, fy=fy) produces,
 * (cvRound(H*fy), cvRound(W*fx)); host-only, no CUDA call.
 * cmr_prepare_image: img (3, H, W) fp32 planes on the device -> out (3, out_h, out_w) fp32
 * planes: cv2.resize (INTER_LINEAR, fp32; exact 2x decimation = 2x2 block mean as in
 * OpenCV) minus the per-channel mean inside (h, w), zeros outside (the padding
 * chainer.dataset.concat_examples(padding=0) adds, mask_rcnn.py:310-311). */
int cmr_prepare_size(int H, int W, double fx, double fy, int* h, int* w);
int cmr_prepare_image(const float* img, int H, int W, double fx, double fy, float mean0,
                      float mean1, float mean2, float* out, int out_h, int out_w,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CMR_B200_H_ */
