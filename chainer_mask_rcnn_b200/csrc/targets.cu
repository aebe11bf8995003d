// Training-target creation on the device (sm_100a), so that the train step has no
// host-side NumPy pass between the RPN and the RoI head:
//
//   cmr_anchor_targets    chainercv AnchorTargetCreator as called per image at
//                         chainer_mask_rcnn/models/mask_rcnn_train_chain.py:151-158
//   cmr_proposal_targets  ProposalTargetCreator.__call__ minus the mask rasterisation
//                         (models/utils/proposal_target_creator.py:115-161)
//
// Both are integer/index work around an IoU matrix that is never materialised: each
// thread owns one anchor / candidate RoI and walks the (few dozen) ground-truth boxes
// held in shared memory.  IoU follows chainercv.utils.bbox_iou in fp32 without FMA
// contraction (areas without +1).  Random subsampling ("np.random.choice(idx, k,
// replace=False)") is a sort of (class, hash(seed, index), index) keys: the first k keys
// of a class are a uniform random k-subset.  The reference draws from NumPy's global
// Mersenne Twister on the host; which subset is drawn is not part of the parity
// contract, its size and eligibility rules are.
#include <math.h>

#include "common.cuh"

namespace cmr {

// nms.cu: descending bitonic sort of `rows` arrays of n_pad (power of two) 64-bit keys.
int launch_sort_desc_u64(unsigned long long* keys, int n_pad, int rows, cudaStream_t st);

namespace {

constexpr int kMaxGt = 256;

__device__ __forceinline__ float iou_f(const float4 a, const float4 b) {  // (y1,x1,y2,x2)
  const float tl_y = fmaxf(a.x, b.x), tl_x = fmaxf(a.y, b.y);
  const float br_y = fminf(a.z, b.z), br_x = fminf(a.w, b.w);
  float inter = __fmul_rn(__fsub_rn(br_y, tl_y), __fsub_rn(br_x, tl_x));
  if (!(tl_y < br_y && tl_x < br_x)) inter = __fmul_rn(inter, 0.0f);
  const float area_a = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float area_b = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

// Effective seed = host seed + *seed_dev (seed_dev may be NULL): a captured CUDA graph
// replays with the same host arguments, the device word advances between replays.
__device__ __forceinline__ unsigned long long eff_seed(unsigned long long seed,
                                                       const unsigned long long* seed_dev) {
  return seed_dev ? seed + 0x9E3779B97F4A7C15ull * __ldg(seed_dev) : seed;
}

__device__ __forceinline__ unsigned int hash31(unsigned long long seed, unsigned long long i) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (i + 1);  // splitmix64
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (unsigned int)(z >> 33) | 1u;  // 31 bits, never zero
}

// chainercv bbox2loc(src, dst) in fp32.
__device__ __forceinline__ float4 box2loc(const float4 s, const float4 d) {
  float h = __fsub_rn(s.z, s.x), w = __fsub_rn(s.w, s.y);
  const float cy = __fadd_rn(s.x, __fmul_rn(0.5f, h)), cx = __fadd_rn(s.y, __fmul_rn(0.5f, w));
  const float bh = __fsub_rn(d.z, d.x), bw = __fsub_rn(d.w, d.y);
  const float bcy = __fadd_rn(d.x, __fmul_rn(0.5f, bh)), bcx = __fadd_rn(d.y, __fmul_rn(0.5f, bw));
  const float eps = 1.1920929e-07f;
  h = fmaxf(h, eps);
  w = fmaxf(w, eps);
  return make_float4(__fdiv_rn(__fsub_rn(bcy, cy), h), __fdiv_rn(__fsub_rn(bcx, cx), w),
                     logf(__fdiv_rn(bh, h)), logf(__fdiv_rn(bw, w)));
}

__device__ __forceinline__ void load_gt(const float4* __restrict__ bbox, int n, float4* sm) {
  for (int g = threadIdx.x; g < n; g += blockDim.x) sm[g] = bbox[g];
  __syncthreads();
}

// ------------------------------------------------------------ anchor targets --
// Pass 1: per-gt maximum IoU over the inside anchors (as uint bits: IoU >= 0).
__global__ void __launch_bounds__(256)
anchor_gtmax_kernel(const float4* __restrict__ anchor, int S, const float4* __restrict__ bbox,
                    const int* __restrict__ n_bbox, int G, float img_h, float img_w,
                    unsigned int* __restrict__ gt_max) {
  __shared__ float4 gt[kMaxGt];
  __shared__ unsigned int smax[kMaxGt];
  const int b = blockIdx.y;
  const int n = min(n_bbox[b], G);
  load_gt(bbox + (size_t)b * G, n, gt);
  for (int g = threadIdx.x; g < n; g += blockDim.x) smax[g] = 0u;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < S) {
    const float4 a = __ldg(anchor + i);
    if (a.x >= 0.f && a.y >= 0.f && a.z <= img_h && a.w <= img_w)
      for (int g = 0; g < n; ++g) {
        const float v = iou_f(a, gt[g]);
        if (v > 0.f) atomicMax(&smax[g], __float_as_uint(v));
      }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < n; g += blockDim.x)
    if (smax[g]) atomicMax(gt_max + (size_t)b * G + g, smax[g]);
}

// Pass 2: labels, box targets, sort keys, class counts.
__global__ void __launch_bounds__(256)
anchor_label_kernel(const float4* __restrict__ anchor, int S, int n_pad,
                    const float4* __restrict__ bbox, const int* __restrict__ n_bbox, int G,
                    float img_h, float img_w, float pos_iou, float neg_iou,
                    const unsigned int* __restrict__ gt_max, unsigned long long seed,
                    const unsigned long long* __restrict__ seed_dev, float4* __restrict__ gt_loc,
                    int* __restrict__ gt_label,
                    unsigned long long* __restrict__ keys, int* __restrict__ counts) {
  __shared__ float4 gt[kMaxGt];
  __shared__ float gmax[kMaxGt];
  const int b = blockIdx.y;
  const int n = min(n_bbox[b], G);
  load_gt(bbox + (size_t)b * G, n, gt);
  for (int g = threadIdx.x; g < n; g += blockDim.x)
    gmax[g] = __uint_as_float(gt_max[(size_t)b * G + g]);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int label = -1;
  unsigned long long key = 0ull;
  if (i < S) {
    const float4 a = __ldg(anchor + i);
    float4 loc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n > 0 && a.x >= 0.f && a.y >= 0.f && a.z <= img_h && a.w <= img_w) {
      float best = -1.f;
      int arg = 0;
      bool is_gt_best = false;
      for (int g = 0; g < n; ++g) {
        const float v = iou_f(a, gt[g]);
        if (v > best) {
          best = v;
          arg = g;
        }
        is_gt_best |= (v == gmax[g]);
      }
      if (best < neg_iou) label = 0;
      if (is_gt_best) label = 1;
      if (best >= pos_iou) label = 1;
      loc = box2loc(a, gt[arg]);
      const unsigned long long r =
          hash31(eff_seed(seed, seed_dev), (unsigned long long)b * S + i);
      if (label == 1) key = (1ull << 63) | (r << 32) | (unsigned int)i;
      if (label == 0) key = (r << 32) | (unsigned int)i;
    }
    gt_loc[(size_t)b * S + i] = loc;
    gt_label[(size_t)b * S + i] = label;
  }
  if (i < n_pad) keys[(size_t)b * n_pad + i] = key;
  const int np = __syncthreads_count(label == 1);
  const int nn = __syncthreads_count(label == 0);
  if (threadIdx.x == 0) {
    if (np) atomicAdd(counts + 2 * b, np);
    if (nn) atomicAdd(counts + 2 * b + 1, nn);
  }
}

// Pass 3 (after the sort): positives beyond the budget and negatives beyond what is
// left of n_sample are set to "ignore".
__global__ void __launch_bounds__(256)
anchor_subsample_kernel(const unsigned long long* __restrict__ keys, int n_pad, int S,
                        const int* __restrict__ counts, int n_sample, int max_pos,
                        int* __restrict__ gt_label) {
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_pad) return;
  const unsigned long long key = keys[(size_t)b * n_pad + k];
  if (key == 0ull) return;
  const int n_pos = counts[2 * b];
  const int kept_pos = min(n_pos, max_pos);
  const int idx = (int)(key & 0xffffffffull);
  const bool drop = (key >> 63) ? (k >= max_pos) : (k - n_pos >= n_sample - kept_pos);
  if (drop) gt_label[(size_t)b * S + idx] = -1;
}

// ---------------------------------------------------------- proposal targets --
// Candidates of image b: its proposals followed by its ground-truth boxes.
__device__ __forceinline__ float4 candidate(const float4* __restrict__ rois, int n_roi,
                                            const float4* gt, int c) {
  return c < n_roi ? __ldg(rois + c) : gt[c - n_roi];
}

__global__ void __launch_bounds__(256)
proposal_assign_kernel(const float4* __restrict__ rois, const int* __restrict__ n_roi_arr,
                       int roi_stride, const float4* __restrict__ bbox,
                       const int* __restrict__ n_bbox, int G, int n_pad, float pos_iou,
                       float neg_hi, float neg_lo, unsigned long long seed,
                       const unsigned long long* __restrict__ seed_dev, int* __restrict__ assign,
                       unsigned long long* __restrict__ keys,
                       int* __restrict__ counts) {
  __shared__ float4 gt[kMaxGt];
  const int b = blockIdx.y;
  const int n = min(n_bbox[b], G);
  const int n_roi = min(n_roi_arr[b], roi_stride);
  load_gt(bbox + (size_t)b * G, n, gt);
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  int cls = -1;
  unsigned long long key = 0ull;
  if (c < n_roi + n && n > 0) {
    const float4 box = candidate(rois + (size_t)b * roi_stride, n_roi, gt, c);
    float best = -INFINITY;
    int arg = 0;
    for (int g = 0; g < n; ++g) {
      const float v = iou_f(box, gt[g]);
      if (v > best || (v != v && best == best)) {  // NaN ranks highest, like numpy argmax
        best = v;
        arg = g;
      }
    }
    assign[(size_t)b * n_pad + c] = arg;
    if (best >= pos_iou) cls = 1;
    else if (best < neg_hi && best >= neg_lo) cls = 0;
    const unsigned long long r =
        hash31(eff_seed(seed, seed_dev), (unsigned long long)b * n_pad + c);
    if (cls == 1) key = (1ull << 63) | (r << 32) | (unsigned int)c;
    if (cls == 0) key = (r << 32) | (unsigned int)c;
  }
  if (c < n_pad) keys[(size_t)b * n_pad + c] = key;
  const int np = __syncthreads_count(cls == 1);
  const int nn = __syncthreads_count(cls == 0);
  if (threadIdx.x == 0) {
    if (np) atomicAdd(counts + 2 * b, np);
    if (nn) atomicAdd(counts + 2 * b + 1, nn);
  }
}

__global__ void __launch_bounds__(256)
proposal_emit_targets_kernel(const unsigned long long* __restrict__ keys, int n_pad,
                             const int* __restrict__ assign, const int* __restrict__ counts,
                             const float4* __restrict__ rois, const int* __restrict__ n_roi_arr,
                             int roi_stride, const float4* __restrict__ bbox,
                             const int* __restrict__ label, const int* __restrict__ n_bbox, int G,
                             int n_sample, int max_pos, float4 mean, float4 stdv,
                             float4* __restrict__ sample_roi, float4* __restrict__ gt_loc,
                             int* __restrict__ gt_label, int* __restrict__ gt_assign,
                             int* __restrict__ n_pos_out) {
  __shared__ float4 gt[kMaxGt];
  const int b = blockIdx.y;
  const int n = min(n_bbox[b], G);
  const int n_roi = min(n_roi_arr[b], roi_stride);
  load_gt(bbox + (size_t)b * G, n, gt);
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_sample) return;
  const int n_pos = counts[2 * b], n_neg = counts[2 * b + 1];
  const int kept_pos = min(n_pos, max_pos);
  const int kept_neg = min(n_sample - kept_pos, n_neg);
  if (j == 0) n_pos_out[b] = kept_pos;
  float4 box = make_float4(0.f, 0.f, 0.f, 0.f), loc = box;
  int lab = -1, asg = -1;
  if (j < kept_pos + kept_neg) {
    const int k = j < kept_pos ? j : n_pos + (j - kept_pos);
    const int c = (int)(keys[(size_t)b * n_pad + k] & 0xffffffffull);
    const int g = assign[(size_t)b * n_pad + c];
    box = candidate(rois + (size_t)b * roi_stride, n_roi, gt, c);
    const float4 l = box2loc(box, gt[g]);
    loc = make_float4(__fdiv_rn(__fsub_rn(l.x, mean.x), stdv.x), __fdiv_rn(__fsub_rn(l.y, mean.y), stdv.y),
                      __fdiv_rn(__fsub_rn(l.z, mean.z), stdv.z), __fdiv_rn(__fsub_rn(l.w, mean.w), stdv.w));
    if (j < kept_pos) {
      lab = __ldg(label + (size_t)b * G + g) + 1;
      asg = g;
    } else {
      lab = 0;
    }
  }
  const size_t o = (size_t)b * n_sample + j;
  sample_roi[o] = box;
  gt_loc[o] = loc;
  gt_label[o] = lab;
  gt_assign[o] = asg;
}

// ------------------------------------------------------------- mask targets --
// ProposalTargetCreator's mask rasterisation (models/utils/proposal_target_creator.py:
// 163-177): round the sampled RoI to integers, crop the assigned instance mask, one-hot
// it, cv2.resize each plane to (ms, ms) with INTER_LINEAR in float32, argmax over the
// planes.  cv2's float32 bilinear is restated exactly (OpenCV resize.cpp, the non-IPP
// path; checked bit for bit against cv2 in tests/test_oracle_mask_target.py):
//   scale = 1/(ms/size) in double;  f = float((d + 0.5)*scale - 0.5);  s = floor(f); f -= s
//   x: s < 0 -> (s, f) = (0, 0);  s >= size-1 -> (s, f) = (size-1, 0)
//   y: source rows clipped to [0, size-1], weights NOT clamped
//   horizontal pass first: r = S[x0]*(1-fx) + S[x1]*fx, then out = r0*(1-fy) + r1*fy,
//   every product and sum rounded to fp32 (no FMA).
// The argmax over one-hot planes only needs the (at most four) values under the taps.
struct Lin {
  int i0, i1;
  float w0, w1;
};

__device__ __forceinline__ Lin lin_coeff(int d, int size, int ms, bool clamp_weights) {
  const double scale = 1.0 / ((double)ms / (double)size);
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (clamp_weights) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= size - 1) { s = size - 1; f = 0.f; }
  }
  Lin l;
  l.i0 = min(max(s, 0), size - 1);
  l.i1 = min(max(s + 1, 0), size - 1);
  l.w0 = __fsub_rn(1.f, f);
  l.w1 = f;
  return l;
}

// Mask pixel readers: one value per element (uint8 / int32) or one bit per pixel
// (rows of ceil(W/8) bytes, bit x & 7 of byte x >> 3 = pixel x: numpy.packbits(...,
// bitorder='little') along the width).
template <typename T>
struct MaskElems {
  const T* base;
  size_t row_pitch;
  __device__ __forceinline__ int at(int y, int x) const { return (int)base[(size_t)y * row_pitch + x]; }
};
struct MaskBits {
  const uint8_t* base;
  size_t row_pitch;
  __device__ __forceinline__ int at(int y, int x) const {
    return (base[(size_t)y * row_pitch + (x >> 3)] >> (x & 7)) & 1;
  }
};

template <typename Reader, typename T>
__global__ void __launch_bounds__(256)
mask_targets_kernel(const T* __restrict__ masks, size_t row_pitch, int G, int H, int W,
                    const float4* __restrict__ sample_roi, const int* __restrict__ gt_assign,
                    const int* __restrict__ n_pos, int n_sample, int ms,
                    int* __restrict__ gt_mask) {
  const int b = blockIdx.y, j = blockIdx.x;
  const size_t row = (size_t)b * n_sample + j;
  int* out = gt_mask + row * ms * ms;
  const int g = j < n_pos[b] ? gt_assign[row] : -1;
  if (g < 0 || g >= G) {
    for (int t = threadIdx.x; t < ms * ms; t += blockDim.x) out[t] = -1;
    return;
  }
  const float4 r = sample_roi[row];
  // np.round(...).astype(int32) = round half to even; python slicing clips to the array
  const int y0 = max((int)rintf(r.x), 0), x0 = max((int)rintf(r.y), 0);
  const int y1 = min((int)rintf(r.z), H), x1 = min((int)rintf(r.w), W);
  const int h = y1 - y0, w = x1 - x0;
  if (h <= 0 || w <= 0) {
    for (int t = threadIdx.x; t < ms * ms; t += blockDim.x) out[t] = 0;
    return;
  }
  Reader m;
  m.base = masks + ((size_t)b * G + g) * H * row_pitch;
  m.row_pitch = row_pitch;
  for (int t = threadIdx.x; t < ms * ms; t += blockDim.x) {
    const int py = t / ms, px = t - py * ms;
    const Lin ly = lin_coeff(py, h, ms, false), lx = lin_coeff(px, w, ms, true);
    int v[4];
    v[0] = m.at(y0 + ly.i0, x0 + lx.i0);
    v[1] = m.at(y0 + ly.i0, x0 + lx.i1);
    v[2] = m.at(y0 + ly.i1, x0 + lx.i0);
    v[3] = m.at(y0 + ly.i1, x0 + lx.i1);
    int best_v = 0;
    float best_s = -1.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = v[k];
      if (c < 0) continue;   // negative labels match no one-hot plane
      const float top = __fadd_rn(v[0] == c ? lx.w0 : 0.f, v[1] == c ? lx.w1 : 0.f);
      const float bot = __fadd_rn(v[2] == c ? lx.w0 : 0.f, v[3] == c ? lx.w1 : 0.f);
      const float s = __fadd_rn(__fmul_rn(top, ly.w0), __fmul_rn(bot, ly.w1));
      if (s > best_s || (s == best_s && c < best_v)) {
        best_s = s;
        best_v = c;
      }
    }
    // planes of values absent under the taps score 0: they win only if every tapped
    // plane scores 0 too, and then np.argmax returns plane 0
    out[t] = best_s > 0.f ? best_v : 0;
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace
}  // namespace cmr

using namespace cmr;

extern "C" size_t cmr_anchor_targets_workspace_bytes(int B, int n_anchor, int max_bbox) {
  if (B <= 0 || n_anchor <= 0 || max_bbox <= 0) return 256;
  return align_up(sizeof(unsigned long long) * (size_t)B * next_pow2(n_anchor), 256) +
         align_up(sizeof(unsigned int) * (size_t)B * max_bbox, 256) + align_up(8 * (size_t)B, 256);
}

extern "C" int cmr_anchor_targets(const float* anchor, int n_anchor, const float* bbox,
                                  const int32_t* n_bbox, int B, int max_bbox, float img_h,
                                  float img_w, int n_sample, float pos_iou_thresh,
                                  float neg_iou_thresh, float pos_ratio, unsigned long long seed,
                                  const unsigned long long* seed_dev, float* gt_loc,
                                  int32_t* gt_label, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  CMR_REQUIRE(anchor && bbox && n_bbox && gt_loc && gt_label && workspace);
  CMR_REQUIRE(B > 0 && n_anchor > 0 && max_bbox > 0 && max_bbox <= kMaxGt && n_sample > 0);
  CMR_REQUIRE(((reinterpret_cast<uintptr_t>(anchor) | reinterpret_cast<uintptr_t>(bbox) |
                reinterpret_cast<uintptr_t>(gt_loc)) & 15) == 0);
  if (workspace_bytes < cmr_anchor_targets_workspace_bytes(B, n_anchor, max_bbox))
    return CMR_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int n_pad = next_pow2(n_anchor);
  char* ws = reinterpret_cast<char*>(workspace);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws);
  ws += align_up(sizeof(unsigned long long) * (size_t)B * n_pad, 256);
  unsigned int* gt_max = reinterpret_cast<unsigned int*>(ws);
  ws += align_up(sizeof(unsigned int) * (size_t)B * max_bbox, 256);
  int* counts = reinterpret_cast<int*>(ws);
  CMR_CUDA_TRY(cudaMemsetAsync(gt_max, 0, sizeof(unsigned int) * (size_t)B * max_bbox, st));
  CMR_CUDA_TRY(cudaMemsetAsync(counts, 0, 8 * (size_t)B, st));
  const float4* a4 = reinterpret_cast<const float4*>(anchor);
  const float4* b4 = reinterpret_cast<const float4*>(bbox);
  dim3 grid(ceil_div(n_pad, 256), B);
  anchor_gtmax_kernel<<<dim3(ceil_div(n_anchor, 256), B), 256, 0, st>>>(
      a4, n_anchor, b4, n_bbox, max_bbox, img_h, img_w, gt_max);
  CMR_LAUNCH_CHECK();
  anchor_label_kernel<<<grid, 256, 0, st>>>(a4, n_anchor, n_pad, b4, n_bbox, max_bbox, img_h,
                                           img_w, pos_iou_thresh, neg_iou_thresh, gt_max, seed,
                                           seed_dev, reinterpret_cast<float4*>(gt_loc), gt_label, keys,
                                           counts);
  CMR_LAUNCH_CHECK();
  int rc = launch_sort_desc_u64(keys, n_pad, B, st);
  if (rc != CMR_OK) return rc;
  anchor_subsample_kernel<<<grid, 256, 0, st>>>(keys, n_pad, n_anchor, counts, n_sample,
                                               (int)(pos_ratio * n_sample), gt_label);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" size_t cmr_proposal_targets_workspace_bytes(int B, int max_roi, int max_bbox) {
  if (B <= 0 || max_roi <= 0 || max_bbox <= 0) return 256;
  const size_t n_pad = next_pow2(max_roi + max_bbox);
  return align_up(sizeof(unsigned long long) * B * n_pad, 256) +
         align_up(sizeof(int) * B * n_pad, 256) + align_up(8 * (size_t)B, 256);
}

extern "C" int cmr_proposal_targets(const float* rois, const int32_t* n_roi, int max_roi,
                                    const float* bbox, const int32_t* label,
                                    const int32_t* n_bbox, int B, int max_bbox, int n_sample,
                                    float pos_ratio, float pos_iou_thresh, float neg_iou_thresh_hi,
                                    float neg_iou_thresh_lo, const float* loc_mean,
                                    const float* loc_std, unsigned long long seed,
                                    const unsigned long long* seed_dev, float* sample_roi,
                                    float* gt_roi_loc, int32_t* gt_roi_label,
                                    int32_t* gt_assign, int32_t* n_pos, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  CMR_REQUIRE(rois && n_roi && bbox && label && n_bbox && loc_mean && loc_std && workspace);
  CMR_REQUIRE(sample_roi && gt_roi_loc && gt_roi_label && gt_assign && n_pos);
  CMR_REQUIRE(B > 0 && max_roi > 0 && max_bbox > 0 && max_bbox <= kMaxGt && n_sample > 0);
  CMR_REQUIRE(((reinterpret_cast<uintptr_t>(rois) | reinterpret_cast<uintptr_t>(bbox) |
                reinterpret_cast<uintptr_t>(sample_roi) | reinterpret_cast<uintptr_t>(gt_roi_loc)) &
               15) == 0);
  if (workspace_bytes < cmr_proposal_targets_workspace_bytes(B, max_roi, max_bbox))
    return CMR_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int n_pad = next_pow2(max_roi + max_bbox);
  char* ws = reinterpret_cast<char*>(workspace);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws);
  ws += align_up(sizeof(unsigned long long) * (size_t)B * n_pad, 256);
  int* assign = reinterpret_cast<int*>(ws);
  ws += align_up(sizeof(int) * (size_t)B * n_pad, 256);
  int* counts = reinterpret_cast<int*>(ws);
  CMR_CUDA_TRY(cudaMemsetAsync(counts, 0, 8 * (size_t)B, st));
  const float4* r4 = reinterpret_cast<const float4*>(rois);
  const float4* b4 = reinterpret_cast<const float4*>(bbox);
  proposal_assign_kernel<<<dim3(ceil_div(n_pad, 256), B), 256, 0, st>>>(
      r4, n_roi, max_roi, b4, n_bbox, max_bbox, n_pad, pos_iou_thresh, neg_iou_thresh_hi,
      neg_iou_thresh_lo, seed, seed_dev, assign, keys, counts);
  CMR_LAUNCH_CHECK();
  int rc = launch_sort_desc_u64(keys, n_pad, B, st);
  if (rc != CMR_OK) return rc;
  const float4 mean = make_float4(loc_mean[0], loc_mean[1], loc_mean[2], loc_mean[3]);
  const float4 stdv = make_float4(loc_std[0], loc_std[1], loc_std[2], loc_std[3]);
  proposal_emit_targets_kernel<<<dim3(ceil_div(n_sample, 256), B), 256, 0, st>>>(
      keys, n_pad, assign, counts, r4, n_roi, max_roi, b4, label, n_bbox, max_bbox, n_sample,
      (int)nearbyintf(n_sample * pos_ratio), mean, stdv, reinterpret_cast<float4*>(sample_roi),
      reinterpret_cast<float4*>(gt_roi_loc), gt_roi_label, gt_assign, n_pos);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_mask_targets(const void* masks, int mask_elem_bytes, int B, int max_bbox,
                                int H, int W, const float* sample_roi,
                                const int32_t* gt_assign, const int32_t* n_pos, int n_sample,
                                int mask_size, int32_t* gt_mask, void* stream) {
  CMR_REQUIRE(masks && sample_roi && gt_assign && n_pos && gt_mask);
  CMR_REQUIRE(B > 0 && max_bbox > 0 && H > 0 && W > 0 && n_sample > 0 && mask_size > 0);
  CMR_REQUIRE(mask_elem_bytes == 0 || mask_elem_bytes == 1 || mask_elem_bytes == 4);
  CMR_REQUIRE((reinterpret_cast<uintptr_t>(sample_roi) & 15) == 0);
  cudaStream_t st = as_stream(stream);
  const float4* r4 = reinterpret_cast<const float4*>(sample_roi);
  dim3 grid(n_sample, B);
  if (mask_elem_bytes == 0)
    mask_targets_kernel<MaskBits, uint8_t><<<grid, 256, 0, st>>>(
        reinterpret_cast<const uint8_t*>(masks), (size_t)((W + 7) / 8), max_bbox, H, W, r4,
        gt_assign, n_pos, n_sample, mask_size, gt_mask);
  else if (mask_elem_bytes == 1)
    mask_targets_kernel<MaskElems<uint8_t>, uint8_t><<<grid, 256, 0, st>>>(
        reinterpret_cast<const uint8_t*>(masks), (size_t)W, max_bbox, H, W, r4, gt_assign, n_pos,
        n_sample, mask_size, gt_mask);
  else
    mask_targets_kernel<MaskElems<int32_t>, int32_t><<<grid, 256, 0, st>>>(
        reinterpret_cast<const int32_t*>(masks), (size_t)W, max_bbox, H, W, r4, gt_assign, n_pos,
        n_sample, mask_size, gt_mask);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}
