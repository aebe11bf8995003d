// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   D[m, n] = epilogue( sum_k A[m, k] * B[n, k] )         TF32 inputs, fp32 accumulate
//
//   m = (image, oy, ox) output pixel, k = (fr, fs, c) filter tap x input channel,
//   A = NHWC activations gathered on the fly (implicit im2col, zero padding),
//   B = filter bank stored K-major as (n, fr, fs, c).
//
// This one kernel is the forward pass of every Convolution2D / Linear /
// Deconvolution2D of the reference graph (chainer.links.Convolution2D at
// models/region_proposal_network.py:75-80, models/mask_rcnn_resnet.py:131-143,
// BuildingBlock convs of models/resnet_extractor.py:47-90) and, fed with the
// transposed / flipped filter bank, their data gradient.  AffineChannel2D
// (functions/affine_channel_2d.py:17-20), bias, residual add, ReLU and the ReLU
// mask of the backward pass are fused in the epilogue.
//
// CTA = 192 threads:
//   warps 0-3  A producers: cp.async 16 B gathers straight into the 128B-swizzled
//              K-major tile the tensor core reads; afterwards the epilogue warps
//              (TMEM lane quarter = warp id)
//   warp  4    B producer: one thread issues TMA (cp.async.bulk.tensor.2d) loads
//   warp  5    one thread issues tcgen05.mma (128 x BN x 8 per instruction) and
//              tcgen05.commit; accumulator lives in TMEM (BN columns)
// full/empty mbarrier ring of STAGES k-blocks (32 fp32 = one 128 B swizzle row).
// Two CTAs are resident per SM, so one CTA's epilogue overlaps the other's MMAs.
#include "common.cuh"
#include "tc_common.cuh"

namespace cmr {
namespace {

using namespace tc;

struct ConvGemmParams {
  const float* a;
  int in_h, in_w, in_c, in_ld;
  int out_h, out_w;
  int kh, kw, stride, pad;
  int M, N, K;
  float* d;
  int d_h, d_w, d_ld, d_stride, d_oy, d_ox;
  const float* scale;
  const float* bias;
  const float* addend;
  const float* mask;
  int relu, round_out;
};

constexpr int kBM = 128;
constexpr int kBK = 32;                      // fp32 per k-block = 128 bytes
constexpr int kABytes = kBM * kBK * 4;       // 16 KB
constexpr int kProducerThreads = 128;
constexpr int kThreads = 192;

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int kBBytes = BN * kBK * 4;
  static constexpr int kAOff = 0;
  static constexpr int kBOff = STAGES * kABytes;
  static constexpr int kBarOff = kBOff + STAGES * kBBytes;
  static constexpr int kTotal = kBarOff + (2 * STAGES + 1) * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;  // slack for 1024 B alignment
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_b, const ConvGemmParams p) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = raw_addr + pad;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBM;
  const int n0 = blockIdx.y * BN;
  const int num_kb = p.K / kBK;

  if (warp == 4 && lane == 0) {
    prefetch_tensormap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], kProducerThreads + 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------ A producer (im2col gather)
    const int t = threadIdx.x;
    const int j = t & 7;
    const int r0 = t >> 3;
    int pix_base[8], iy0[8], ix0[8];
    const int ohw = p.out_h * p.out_w;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + r0 + 16 * i;
      if (m < p.M) {
        const int img = m / ohw;
        const int rem = m - img * ohw;
        const int oy = rem / p.out_w;
        const int ox = rem - oy * p.out_w;
        pix_base[i] = img * p.in_h * p.in_w;
        iy0[i] = oy * p.stride - p.pad;
        ix0[i] = ox * p.stride - p.pad;
      } else {
        pix_base[i] = 0;
        iy0[i] = -(1 << 28);  // never inside the image -> zero fill
        ix0[i] = -(1 << 28);
      }
    }
    const uint32_t dst_off =
        (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((j ^ (r0 & 7)) << 4));
    const int cpt = p.in_c / kBK;  // k-blocks per filter tap
    int fr = 0, fs = 0, cb = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t phase = (kb / STAGES) & 1;
      mbar_wait(&empty_bar[s], phase ^ 1);
      const uint32_t a_stage = smem_base + L::kAOff + s * kABytes + dst_off;
      const float* src_c = p.a + cb * kBK + j * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int iy = iy0[i] + fr, ix = ix0[i] + fs;
        const bool ok = (unsigned)iy < (unsigned)p.in_h && (unsigned)ix < (unsigned)p.in_w;
        const float* src =
            ok ? src_c + (size_t)(pix_base[i] + iy * p.in_w + ix) * p.in_ld : p.a;
        cp_async_16(a_stage + i * 2048, src, ok ? 16u : 0u);
      }
      cp_async_mbar_arrive_noinc(&full_bar[s]);
      if (++cb == cpt) {
        cb = 0;
        if (++fs == p.kw) {
          fs = 0;
          ++fr;
        }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");

    // ------------------------------------------------------------- epilogue
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int row = m0 + warp * 32 + lane;
    size_t doff = 0;
    const bool row_ok = row < p.M;
    if (row_ok) {
      const int img = row / ohw;
      const int rem = row - img * ohw;
      const int oy = rem / p.out_w;
      const int ox = rem - oy * p.out_w;
      doff = ((size_t)(img * p.d_h + oy * p.d_stride + p.d_oy) * p.d_w + ox * p.d_stride +
              p.d_ox) * p.d_ld;
    }
    const bool vec_ok = ((p.d_ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.d) & 15) == 0) &&
                        (!p.addend || (reinterpret_cast<uintptr_t>(p.addend) & 15) == 0) &&
                        (!p.mask || (reinterpret_cast<uintptr_t>(p.mask) & 15) == 0);
#pragma unroll 1
    for (int chunk = 0; chunk < BN / 32; ++chunk) {
      const int nc = n0 + chunk * 32;
      if (nc >= p.N) break;
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(chunk * 32), v);
      tmem_ld_wait();
      if (!row_ok) continue;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int n = nc + g * 4;
        if (n >= p.N) break;
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = __uint_as_float(v[g * 4 + e]);
        const bool full4 = vec_ok && (n + 3 < p.N);
        if (full4) {
          if (p.scale) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + n));
            o[0] *= sc.x; o[1] *= sc.y; o[2] *= sc.z; o[3] *= sc.w;
          }
          if (p.bias) {
            const float4 bi = __ldg(reinterpret_cast<const float4*>(p.bias + n));
            o[0] += bi.x; o[1] += bi.y; o[2] += bi.z; o[3] += bi.w;
          }
          if (p.addend) {
            const float4 ad = __ldg(reinterpret_cast<const float4*>(p.addend + doff + n));
            o[0] += ad.x; o[1] += ad.y; o[2] += ad.z; o[3] += ad.w;
          }
          if (p.relu) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
          }
          if (p.mask) {
            const float4 mk = __ldg(reinterpret_cast<const float4*>(p.mask + doff + n));
            o[0] = mk.x > 0.f ? o[0] : 0.f; o[1] = mk.y > 0.f ? o[1] : 0.f;
            o[2] = mk.z > 0.f ? o[2] : 0.f; o[3] = mk.w > 0.f ? o[3] : 0.f;
          }
          if (p.round_out) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = round_tf32(o[e]);
          }
          *reinterpret_cast<float4*>(p.d + doff + n) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
          for (int e = 0; e < 4 && n + e < p.N; ++e) {
            float x = o[e];
            if (p.scale) x *= __ldg(p.scale + n + e);
            if (p.bias) x += __ldg(p.bias + n + e);
            if (p.addend) x += __ldg(p.addend + doff + n + e);
            if (p.relu) x = fmaxf(x, 0.f);
            if (p.mask) x = __ldg(p.mask + doff + n + e) > 0.f ? x : 0.f;
            if (p.round_out) x = round_tf32(x);
            p.d[doff + n + e] = x;
          }
        }
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------ B producer (TMA)
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t phase = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], L::kBBytes);
        tma_load_2d(smem_base + L::kBOff + s * L::kBBytes, &tmap_b, &full_bar[s], kb * kBK, n0);
      }
    }
  } else {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_tf32(kBM, BN, 0, 0);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t phase = (kb / STAGES) & 1;
      mbar_wait(&full_bar[s], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t da = make_smem_desc_sw128(smem_base + L::kAOff + s * kABytes, 16, 1024);
        const uint64_t db =
            make_smem_desc_sw128(smem_base + L::kBOff + s * L::kBBytes, 16, 1024);
#pragma unroll
        for (int k = 0; k < kBK / 8; ++k)  // 8 tf32 = 32 bytes per MMA -> +2 (16 B units)
          umma_tf32(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[s]);
      }
      __syncwarp();
    }
    if (lane == 0) umma_commit(tmem_full_bar);
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
}

// ------------------------------------------------------------------ host ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D fp32 tensor (rows, cols) row-major, box = (box_rows, 32 cols), 128 B swizzle.
int make_tmap_2d(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols,
                 uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return CMR_ERR_CUDA;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)kBK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim,
                  gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CMR_OK : CMR_ERR_CUDA;
}

template <int BN, int STAGES>
int launch(const CUtensorMap& tmap, const ConvGemmParams& p, cudaStream_t st) {
  using L = SmemLayout<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    CMR_CUDA_TRY(cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, STAGES>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      L::kDynamic));
    configured = true;
  }
  dim3 grid(ceil_div(p.M, kBM), ceil_div(p.N, BN));
  prof_begin(kProfConvGemm, 2.0 * p.M * (double)p.N * p.K, st);
  conv_gemm_tc_kernel<BN, STAGES><<<grid, kThreads, L::kDynamic, st>>>(tmap, p);
  prof_end(st);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

}  // namespace
}  // namespace cmr

using namespace cmr;

extern "C" int cmr_conv_gemm_tc(const cmr_conv_desc* c, const float* a, const float* w, float* d,
                                const float* scale, const float* bias, const float* addend,
                                const float* mask, void* stream) {
  CMR_REQUIRE(c && a && w && d);
  CMR_REQUIRE(c->batch > 0 && c->in_h > 0 && c->in_w > 0 && c->out_h > 0 && c->out_w > 0);
  CMR_REQUIRE(c->kh > 0 && c->kw > 0 && c->stride > 0 && c->pad >= 0 && c->n > 0);
  if (c->in_c <= 0 || c->in_c % kBK != 0) return CMR_ERR_UNSUPPORTED;
  CMR_REQUIRE(c->in_ld > 0 && c->in_ld % 4 == 0);
  CMR_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0);
  CMR_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0);
  const long long M = (long long)c->batch * c->out_h * c->out_w;
  CMR_REQUIRE(M > 0 && M < (1ll << 31));
  CMR_REQUIRE((long long)c->batch * c->in_h * c->in_w < (1ll << 31));
  CMR_REQUIRE((long long)c->batch * c->d_h * c->d_w < (1ll << 31));
  ConvGemmParams p;
  p.a = a;
  p.in_h = c->in_h; p.in_w = c->in_w; p.in_c = c->in_c; p.in_ld = c->in_ld;
  p.out_h = c->out_h; p.out_w = c->out_w;
  p.kh = c->kh; p.kw = c->kw; p.stride = c->stride; p.pad = c->pad;
  p.M = (int)M; p.N = c->n; p.K = c->kh * c->kw * c->in_c;
  p.d = d;
  p.d_h = c->d_h; p.d_w = c->d_w; p.d_ld = c->d_ld; p.d_stride = c->d_stride;
  p.d_oy = c->d_oy; p.d_ox = c->d_ox;
  p.scale = scale; p.bias = bias; p.addend = addend; p.mask = mask;
  p.relu = c->relu; p.round_out = c->round_tf32;
  CMR_REQUIRE(p.d_ld >= p.N && p.d_stride >= 1);
  CMR_REQUIRE((c->out_h - 1) * c->d_stride + c->d_oy < c->d_h);
  CMR_REQUIRE((c->out_w - 1) * c->d_stride + c->d_ox < c->d_w);

  // Tile width: the widest N tile that does not waste more than half a tile.
  int bn = c->tile_n;
  if (bn == 0) bn = p.N > 128 ? 256 : (p.N > 64 ? 128 : 64);
  CUtensorMap tmap;
  int rc = make_tmap_2d(&tmap, w, (uint64_t)p.N, (uint64_t)p.K, (uint32_t)bn);
  if (rc != CMR_OK) return rc;
  cudaStream_t st = as_stream(stream);
  switch (bn) {
    case 64: return launch<64, 4>(tmap, p, st);
    case 128: return launch<128, 3>(tmap, p, st);
    case 256: return launch<256, 4>(tmap, p, st);
    default: return CMR_ERR_INVALID_ARG;
  }
}
