// AffineChannel2D as a stand-alone operator, the BatchNormalization -> AffineChannel2D
// fold, and the NCHW <-> NHWC re-layouts used by the reference-layout ROIAlign operator.
//
//   affine_channel_2d   chainer_mask_rcnn/functions/affine_channel_2d.py:17-20 (forward:
//                       y = W * x + b, W and b (1,C,1,1)) and :48-55 (backward: gx = W * gy,
//                       gW = sum over (n,h,w) of x * gy, gb = sum over (n,h,w) of gy)
//   bn fold             chainer_mask_rcnn/models/resnet_extractor.py:16-29
//                       (_get_affine_from_bn: W = gamma / sqrt(var + 1e-5),
//                       b = beta - mean * W)
//
// Inside the model the affine never runs alone (it is the epilogue of the convolution
// kernels); these kernels are the operator surface.  All HBM-bound: one read of every input,
// one write of every output, float4 lanes over the contiguous H*W axis.
#include "common.cuh"

namespace cmr {
namespace {

// One CTA walks a slice of one (n, c) plane; grid = (planes, slices).
__global__ void __launch_bounds__(256)
affine_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                  const float* __restrict__ b, float* __restrict__ y, int C, int HW) {
  const size_t plane = blockIdx.x;
  const int c = (int)(plane % C);
  const float w = __ldg(W + c), bb = __ldg(b + c);
  const float* __restrict__ xp = x + plane * HW;
  float* __restrict__ yp = y + plane * HW;
  const int per = ceil_div(HW, (int)gridDim.y);
  const int lo = blockIdx.y * per, hi = min(HW, lo + per);
  // planes start at arbitrary multiples of HW floats: vector lanes from the first 16-byte
  // boundary of the slice on
  int head = (int)(((16 - (reinterpret_cast<uintptr_t>(xp + lo) & 15)) & 15) >> 2);
  const bool same = ((reinterpret_cast<uintptr_t>(xp) ^ reinterpret_cast<uintptr_t>(yp)) & 15) == 0;
  if (!same) head = hi - lo;          // x and y misaligned against each other: scalar only
  head = min(head, hi - lo);
  for (int i = lo + threadIdx.x; i < lo + head; i += blockDim.x)
    yp[i] = __fadd_rn(__fmul_rn(w, xp[i]), bb);
  const int v0 = lo + head;
  const int nvec = (hi - v0) >> 2;
  const float4* __restrict__ xv = reinterpret_cast<const float4*>(xp + v0);
  float4* __restrict__ yv = reinterpret_cast<float4*>(yp + v0);
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const float4 v = __ldg(xv + i);
    // W * x + b as a multiplication and an addition (NumPy's two roundings)
    yv[i] = make_float4(__fadd_rn(__fmul_rn(w, v.x), bb), __fadd_rn(__fmul_rn(w, v.y), bb),
                        __fadd_rn(__fmul_rn(w, v.z), bb), __fadd_rn(__fmul_rn(w, v.w), bb));
  }
  for (int i = v0 + 4 * nvec + threadIdx.x; i < hi; i += blockDim.x)
    yp[i] = __fadd_rn(__fmul_rn(w, xp[i]), bb);
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (warp == 0) {
    t = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;   // valid in warp 0
}

// grid = (C, slices): CTA (c, s) covers images n = s, s + slices, ... of channel c: writes
// gx there and the partial sums part[(c * slices + s) * 2 + {0, 1}].
__global__ void __launch_bounds__(256)
affine_bwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                  const float* __restrict__ gy, float* __restrict__ gx,
                  float* __restrict__ part, int N, int C, int HW) {
  __shared__ float red[8];
  const int c = blockIdx.x, s = blockIdx.y, S = gridDim.y;
  const float w = __ldg(W + c);
  float sw = 0.f, sb = 0.f;
  for (int n = s; n < N; n += S) {
    const size_t off = ((size_t)n * C + c) * HW;
    const float* __restrict__ xp = x + off;
    const float* __restrict__ gp = gy + off;
    float* __restrict__ op = gx + off;
    const bool vec = ((reinterpret_cast<uintptr_t>(xp) | reinterpret_cast<uintptr_t>(gp) |
                       reinterpret_cast<uintptr_t>(op)) & 15) == 0;
    const int nvec = vec ? HW >> 2 : 0;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(xp) + i);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gp) + i);
      reinterpret_cast<float4*>(op)[i] = make_float4(w * g.x, w * g.y, w * g.z, w * g.w);
      sw += a.x * g.x + a.y * g.y + a.z * g.z + a.w * g.w;
      sb += g.x + g.y + g.z + g.w;
    }
    for (int i = 4 * nvec + threadIdx.x; i < HW; i += blockDim.x) {
      const float g = gp[i];
      op[i] = w * g;
      sw += xp[i] * g;
      sb += g;
    }
  }
  const float tw = block_sum(sw, red);
  const float tb = block_sum(sb, red);
  if (threadIdx.x == 0) {
    part[((size_t)c * S + s) * 2] = tw;
    part[((size_t)c * S + s) * 2 + 1] = tb;
  }
}

__global__ void __launch_bounds__(128)
affine_bwd_finish_kernel(const float* __restrict__ part, float* __restrict__ gW,
                         float* __restrict__ gb, int C, int S) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float sw = 0.f, sb = 0.f;
  for (int s = 0; s < S; ++s) {     // fixed order: deterministic
    sw += part[((size_t)c * S + s) * 2];
    sb += part[((size_t)c * S + s) * 2 + 1];
  }
  gW[c] = sw;
  gb[c] = sb;
}

int bwd_slices(int N, int C) {
  int s = ceil_div(2 * sm_count(), C);
  if (s > N) s = N;
  return s < 1 ? 1 : s;
}

__global__ void __launch_bounds__(128)
bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
               const float* __restrict__ mean, const float* __restrict__ var, float eps,
               float* __restrict__ W, float* __restrict__ b, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // resnet_extractor.py:23-25, each step one IEEE fp32 operation
  const float std_ = __fsqrt_rn(__fadd_rn(var[c], eps));
  const float w = __fdiv_rn(gamma[c], std_);
  W[c] = w;
  b[c] = __fsub_rn(beta[c], __fmul_rn(mean[c], w));
}

// (B, rows, cols) -> (B, cols, rows) through a 32 x 33 shared-memory tile: both sides move
// full 128-byte lines.  NCHW -> NHWC is rows = C, cols = H*W; NHWC -> NCHW the reverse.
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  __shared__ float tile[32][33];
  const size_t base = (size_t)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int r = r0 + ty + k, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + k][tx] = __ldg(in + base + (size_t)r * cols + c);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int c = c0 + ty + k, r = r0 + tx;
    if (r < rows && c < cols) out[base + (size_t)c * rows + r] = tile[tx][ty + k];
  }
}

}  // namespace
}  // namespace cmr

using namespace cmr;

extern "C" int cmr_affine_channel_fwd(const float* x, const float* W, const float* b, int N,
                                      int C, int HW, float* y, void* stream) {
  CMR_REQUIRE(N >= 0 && C > 0 && HW > 0);
  if (N == 0) return CMR_OK;
  CMR_REQUIRE(x && W && b && y);
  CMR_REQUIRE((long long)N * C < (1ll << 31));
  // enough CTAs to fill the machine when there are few planes
  int slices = ceil_div(4 * sm_count(), N * C);
  slices = slices < 1 ? 1 : (slices > ceil_div(HW, 1024) ? ceil_div(HW, 1024) : slices);
  if (slices > 65535) slices = 65535;
  affine_fwd_kernel<<<dim3(N * C, slices), 256, 0, as_stream(stream)>>>(x, W, b, y, C, HW);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" size_t cmr_affine_channel_bwd_workspace_bytes(int N, int C) {
  if (N <= 0 || C <= 0) return 0;
  return sizeof(float) * 2 * (size_t)C * bwd_slices(N, C);
}

extern "C" int cmr_affine_channel_bwd(const float* x, const float* W, const float* gy, int N,
                                      int C, int HW, float* gx, float* gW, float* gb,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  CMR_REQUIRE(N > 0 && C > 0 && HW > 0 && C <= 65535 * 32);
  CMR_REQUIRE(x && W && gy && gx && gW && gb && workspace);
  if (workspace_bytes < cmr_affine_channel_bwd_workspace_bytes(N, C)) return CMR_ERR_WORKSPACE;
  const int S = bwd_slices(N, C);
  float* part = static_cast<float*>(workspace);
  affine_bwd_kernel<<<dim3(C, S), 256, 0, as_stream(stream)>>>(x, W, gy, gx, part, N, C, HW);
  CMR_LAUNCH_CHECK();
  affine_bwd_finish_kernel<<<ceil_div(C, 128), 128, 0, as_stream(stream)>>>(part, gW, gb, C, S);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_bn_fold(const float* gamma, const float* beta, const float* mean,
                           const float* var, float eps, int C, float* W, float* b,
                           void* stream) {
  CMR_REQUIRE(C > 0 && gamma && beta && mean && var && W && b);
  bn_fold_kernel<<<ceil_div(C, 128), 128, 0, as_stream(stream)>>>(gamma, beta, mean, var, eps,
                                                                 W, b, C);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

extern "C" int cmr_transpose_batched(const float* in, int batch, int rows, int cols, float* out,
                                     void* stream) {
  CMR_REQUIRE(batch >= 0 && rows > 0 && cols > 0);
  if (batch == 0) return CMR_OK;
  CMR_REQUIRE(in && out && batch <= 65535 && ceil_div(rows, 32) <= 65535);
  transpose_kernel<<<dim3(ceil_div(cols, 32), ceil_div(rows, 32), batch), 256, 0,
                     as_stream(stream)>>>(in, out, rows, cols);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}
