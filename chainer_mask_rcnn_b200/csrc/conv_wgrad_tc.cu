// Weight gradient of Convolution2D / Deconvolution2D / Linear on tcgen05 (sm_100a).
//
//   gW[i, j] += row_scale[i] * sum_pix  P[map_p(pix), p_c0 + i] * Q[map_q(pix), q_c0 + j]
//
// a "pixel-reduction GEMM": the reduction runs over output pixels (b, oy, ox), P is
// the upstream gradient gy (rows i = output channels), Q the layer input x at one
// filter tap (columns j = input channels).  Both operands are therefore MN-major
// for the tensor core (the contiguous axis is the channel axis, not the reduction
// axis): tiles are staged as [pixel][32 channels = 128 B] rows in shared memory with
// the SWIZZLE_128B_BASE32B pattern (the one layout tcgen05 reads MN-major 32-bit
// operands from) and read with MN-major UMMA descriptors.
// Each operand has its own pixel map (stride, offset), which covers stride-2 convs
// (x sampled at 2*oy - pad + fr), zero padding (out-of-image -> zeros) and the
// 2x2 stride-2 deconvolution (gy sampled at 2*y + dy).
//
// This replaces the gW computation of chainer's Convolution2DGradW /
// Deconvolution2D / Linear backward for the links built at
// models/region_proposal_network.py:75-80 and models/mask_rcnn_resnet.py:131-143.
//
// Two kernels share the tile layout, the MMA issue loop and the epilogue:
//   conv_wgrad_tma_kernel  (default) CTA = 192 threads: warps 0-3 epilogue, warp 4 = one
//       thread issuing im2col-mode TMA loads (32 pixels x 32 channels per instruction,
//       SWIZZLE_128B_ATOM_32B, padding and ragged channel ranges zero-filled by the copy
//       engine), warp 5 issues tcgen05.mma (128 x BN x 8, BN up to 256).
//   conv_wgrad_tc_kernel   (fallback for geometries a tensor map cannot describe)
//       CTA = 288 threads: warps 0-3 gather P, warps 4-7 gather Q with cp.async 16 B.
// grid = (row tiles, column tiles, taps x K splits); partial sums are added with vector
// red.global.add (gW must be zeroed by the caller once per step).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace cmr {
namespace {

using namespace tc;

// n / d and n % d for 0 <= n < 2^31 by a precomputed multiply-shift (d >= 1).
struct FastDiv {
  unsigned int d, mul, shr;
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const {
    q = d == 1 ? n : (int)(__umulhi((unsigned int)n, mul) >> shr);
    r = n - q * (int)d;
  }
};

struct PixelMap {
  const float* base;
  int h, w, ld;      // tensor (batch, h, w, ld)
  int stride, off_y, off_x;
  int c0;            // first channel of this operand's tile range
};

struct WgradParams {
  PixelMap p, q;
  int loop_h, loop_w;  // pixel loop space (batch, loop_h, loop_w)
  FastDiv div_hw, div_w;
  int M;               // number of pixels
  int rows, cols;      // gW tile space: rows = output channels, cols = channels of one tap
  float* gw;
  int gw_ld;           // floats between consecutive rows of gW
  int gw_col0;         // column offset of this tap inside a gW row
  const float* row_scale;
  int kb_per_split, num_kb;
  int splits, taps_w;  // grid.z = taps * splits; tap (fr, fs) shifts q by (fr, fs) pixels
  int p_tiled, q_tiled;  // operand is a plain (pixels x channels) matrix: tiled-mode TMA
  int probe;           // timing probes (CMR_WGRAD_PROBE): 1 = loads only for the first ring
                       // fill, 2 = no MMAs; results are garbage, 0 in normal operation
  int fixed;           // deterministic mode: gw points to int64 fixed-point words (common.cuh)
};

constexpr int kBM = 128;
constexpr int kPix = 32;  // pixels per k-block
constexpr int kThreads = 288;

template <int BN, int STAGES>
struct WSmem {
  static constexpr int kPBytes = kPix * kBM * 4;  // 16 KB
  static constexpr int kQBytes = kPix * BN * 4;
  static constexpr int kPOff = 0;
  static constexpr int kQOff = STAGES * kPBytes;
  static constexpr int kBarOff = kQOff + STAGES * kQBytes;
  static constexpr int kTotal = kBarOff + (2 * STAGES + 1) * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;
};

// Gathers one k-block (32 pixels) x (NBLK * 32 channels) of an operand.
// Lane layout per cp.async: 4 pixel rows x 128 B, i.e. 512 contiguous smem bytes.
template <int NBLK>
__device__ __forceinline__ void gather_tile(const PixelMap& pm, int pix0, int M,
                                            const FastDiv& div_hw, const FastDiv& div_w,
                                            int chan_limit, uint32_t stage_addr,
                                            int t /* 0..127 */) {
  const int jj = t & 7;
  const int rsub = (t >> 3) & 3;
  const int w = t >> 5;
  // A thread owns two pixel rows (decoded once each) x all NBLK channel blocks.
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int row = (2 * w + rr) * 4 + rsub;
    const int pix = pix0 + row;
    bool ok = pix < M;
    const float* src_row = pm.base;
    if (ok) {
      int img, rem, oy, ox;
      div_hw.divmod(pix, img, rem);
      div_w.divmod(rem, oy, ox);
      const int y = oy * pm.stride + pm.off_y, x = ox * pm.stride + pm.off_x;
      ok = (unsigned)y < (unsigned)pm.h && (unsigned)x < (unsigned)pm.w;
      if (ok) src_row = pm.base + ((size_t)(img * pm.h + y) * pm.w + x) * pm.ld;
    }
    // SWIZZLE_128B_BASE32B: 32-byte chunk (jj >> 1) lands at chunk ^ (row & 3).
    const uint32_t dst_row =
        stage_addr + row * 128 + ((((jj >> 1) ^ (row & 3)) << 5) | ((jj & 1) << 4));
#pragma unroll
    for (int cblk = 0; cblk < NBLK; ++cblk) {
      const int ch = pm.c0 + cblk * 32 + jj * 4;
      const bool okc = ok && ch < chan_limit;
      cp_async_16(dst_row + cblk * (kPix * 128), okc ? src_row + ch : pm.base, okc ? 16u : 0u);
    }
  }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads)
conv_wgrad_tc_kernel(const WgradParams p) {
  using L = WSmem<BN, STAGES>;
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
  const int i0 = blockIdx.x * kBM;
  const int j0 = blockIdx.y * BN;
  const int tap = blockIdx.z / p.splits;
  const int tap_fr = tap / p.taps_w, tap_fs = tap - tap_fr * p.taps_w;
  const int kb_begin = (blockIdx.z - tap * p.splits) * p.kb_per_split;
  const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
  const int nkb = kb_end - kb_begin;  // >= 1 by construction of the grid

  if (warp == 8 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 256);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 7) tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ---------------------------------------------------------- producers
    const bool is_p = warp < 4;
    const int t = threadIdx.x & 127;
    PixelMap pm = is_p ? p.p : p.q;
    pm.c0 += is_p ? i0 : j0;
    if (!is_p) {
      pm.off_y += tap_fr;
      pm.off_x += tap_fs;
    }
    const int chan_limit = (is_p ? p.p.c0 + p.rows : p.q.c0 + p.cols);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t phase = (kb / STAGES) & 1;
      mbar_wait(&empty_bar[s], phase ^ 1);
      const int pix0 = (kb_begin + kb) * kPix;
      if (is_p)
        gather_tile<kBM / 32>(pm, pix0, p.M, p.div_hw, p.div_w, chan_limit,
                              smem_base + L::kPOff + s * L::kPBytes, t);
      else
        gather_tile<BN / 32>(pm, pix0, p.M, p.div_hw, p.div_w, chan_limit,
                             smem_base + L::kQOff + s * L::kQBytes, t);
      cp_async_mbar_arrive_noinc(&full_bar[s]);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");

    if (warp < 4) {
      // -------------------------------------------------------- epilogue
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      const int row = i0 + warp * 32 + lane;
      const bool row_ok = row < p.rows;
      const float sc = (row_ok && p.row_scale) ? __ldg(p.row_scale + row) : 1.0f;
      float* out_row = p.gw + (size_t)row * p.gw_ld + p.gw_col0 + tap * p.cols;
      const bool vec_ok = ((p.gw_ld & 3) == 0) && ((p.gw_col0 & 3) == 0) &&
                          ((reinterpret_cast<uintptr_t>(p.gw) & 15) == 0);
#pragma unroll 1
      for (int chunk = 0; chunk < BN / 32; ++chunk) {
        const int jc = j0 + chunk * 32;
        if (jc >= p.cols) break;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(chunk * 32), v);
        tmem_ld_wait();
        if (!row_ok) continue;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int j = jc + g * 4;
          if (j >= p.cols) break;
          if (p.fixed) {
            long long* frow = reinterpret_cast<long long*>(p.gw) + (size_t)row * p.gw_ld +
                              p.gw_col0 + tap * p.cols;
            for (int e = 0; e < 4 && j + e < p.cols; ++e)
              red_fixed(frow + j + e, __uint_as_float(v[g * 4 + e]) * sc);
          } else if (vec_ok && j + 3 < p.cols) {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out_row + j),
                         "f"(__uint_as_float(v[g * 4 + 0]) * sc),
                         "f"(__uint_as_float(v[g * 4 + 1]) * sc),
                         "f"(__uint_as_float(v[g * 4 + 2]) * sc),
                         "f"(__uint_as_float(v[g * 4 + 3]) * sc)
                         : "memory");
          } else {
            for (int e = 0; e < 4 && j + e < p.cols; ++e)
              atomicAdd(out_row + j + e, __uint_as_float(v[g * 4 + e]) * sc);
          }
        }
      }
    }
  } else {
    // ---------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc = make_idesc_tf32(kBM, BN, 1, 1);  // both operands MN-major
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t phase = (kb / STAGES) & 1;
      mbar_wait(&full_bar[s], phase);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t pa = smem_base + L::kPOff + s * L::kPBytes;
        const uint32_t qa = smem_base + L::kQOff + s * L::kQBytes;
#pragma unroll
        for (int k = 0; k < kPix / 8; ++k) {
          // 8 pixels = two 4-row (512 B) swizzle atoms per 32-channel block (stride
          // byte offset); blocks are kPix * 128 bytes apart (leading byte offset).
          const uint64_t da = make_smem_desc(pa + k * 1024, kPix * 128, 512, 1);
          const uint64_t db = make_smem_desc(qa + k * 1024, kPix * 128, 512, 1);
          umma_tf32(tmem_base, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      __syncwarp();
    }
    if (lane == 0) umma_commit(tmem_full_bar);
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 7) tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
}

// ------------------------------------------------------------ TMA variant ----
constexpr int kTmaThreads = 192;
constexpr int kTPix = 48;   // pixels per k-block: TMA boxes of 6 KB

template <int BN, int STAGES, bool PAIR>
struct TSmem {
  static constexpr int kQCols = PAIR ? BN / 2 : BN;  // a CTA of a pair holds half the columns
  static constexpr int kPBytes = kTPix * kBM * 4;    // 24 KB
  static constexpr int kQBytes = kTPix * kQCols * 4;
  static constexpr int kPOff = 0;
  static constexpr int kQOff = STAGES * kPBytes;
  static constexpr int kBarOff = kQOff + STAGES * kQBytes;
  static constexpr int kTotal = kBarOff + (2 * STAGES + 1) * 8 + 16;
  static constexpr int kDynamic = kTotal + 1024;
};

// PAIR: clusters of two CTAs along grid.x (adjacent 128-row tiles, same column tile, same
// pixel range) compute ONE 256 x BN tile with tcgen05.mma.cta_group::2: each CTA loads its
// 128 rows of P and BN/2 columns of Q (48 KB instead of 72 KB per 48-pixel k-block and SM:
// the single-CTA kernel is bound by operand delivery), the leader issues the MMAs, each
// CTA's TMEM receives its 128 rows x BN columns and each CTA runs its own epilogue.
template <int BN, int STAGES, bool PAIR>
__global__ void __launch_bounds__(kTmaThreads)
conv_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmap_p,
                      const __grid_constant__ CUtensorMap tmap_q, const WgradParams p) {
  using L = TSmem<BN, STAGES, PAIR>;
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
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0;
  const int i0 = blockIdx.x * kBM;                 // this CTA's rows (pairs: 2 adjacent tiles)
  const int j0 = blockIdx.y * BN;                  // the tile's first column
  const int jq = j0 + (int)cta_rank * L::kQCols;   // first column this CTA loads
  const int tap = blockIdx.z / p.splits;
  const int tap_fr = tap / p.taps_w, tap_fs = tap - tap_fr * p.taps_w;
  const int kb_begin = (blockIdx.z - tap * p.splits) * p.kb_per_split;
  const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
  const int nkb = kb_end - kb_begin;  // >= 1 by construction of the grid
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;

  if (warp == 4 && lane == 0) {
    prefetch_tensormap(&tmap_p);
    prefetch_tensormap(&tmap_q);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    if (PAIR) tmem_alloc_pair(tmem_slot, kTmemCols);
    else tmem_alloc(tmem_slot, kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // both CTAs' barriers and TMEM exist before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // 32-channel blocks that hold data: blocks past the channel range would only load zeros,
  // their copies are skipped (the epilogue never writes rows >= rows or columns >= cols)
  auto blocks_in = [](int first, int limit, int max_blocks) {
    const int b = (limit - first + 31) / 32;
    return b < 0 ? 0 : (b > max_blocks ? max_blocks : b);
  };
  const int p_blocks = blocks_in(i0, p.rows, kBM / 32);
  const int q_blocks = blocks_in(jq, p.cols, L::kQCols / 32);

  if (warp <= 4) {
    // ------------------------------------------------------------ producers
    // One thread of each of warps 0-4 issues a share of the k-block's TMA loads (a copy
    // instruction occupies its issuing thread for ~50 cycles, which would otherwise sit on
    // the critical path of every stage); warp 4 (of the leader CTA) also arms the stage's
    // barrier with the byte count of all of them.  Warps 0-3 turn into the epilogue
    // afterwards.
    if (lane == 0) {
      const int n_blocks = p_blocks + q_blocks;
      // bytes of the peer CTA's boxes, which complete on the leader's barrier too
      int all_blocks = n_blocks;
      if (PAIR)
        all_blocks += blocks_in(i0 + kBM, p.rows, kBM / 32) +
                      blocks_in(j0 + L::kQCols, p.cols, L::kQCols / 32);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t phase = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], phase ^ 1);
        if (warp == 4 && cta_rank == 0) {
          if (p.probe == 1 && kb >= STAGES) {
            mbar_arrive(&full_bar[s]);
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[s], (uint32_t)all_blocks * (kTPix * 128));
        } else if (p.probe == 1 && kb >= STAGES) {
          continue;
        }
        int img, rem, oy, ox;
        const int pix0 = (kb_begin + kb) * kTPix;
        p.div_hw.divmod(pix0, img, rem);
        p.div_w.divmod(rem, oy, ox);
        const uint32_t pa = smem_base + L::kPOff + s * L::kPBytes;
        const uint32_t qa = smem_base + L::kQOff + s * L::kQBytes;
        const int ph = oy * p.p.stride + p.p.off_y, pw = ox * p.p.stride + p.p.off_x;
        const int qh = oy * p.q.stride + p.q.off_y, qw = ox * p.q.stride + p.q.off_x;
        const uint32_t bar_pair = PAIR ? mapa_cluster(smem_u32(&full_bar[s]), 0) : 0;
        for (int b = warp; b < n_blocks; b += 5) {
          const bool is_p = b < p_blocks;
          const int bb = is_p ? b : b - p_blocks;
          const uint32_t dst = (is_p ? pa : qa) + bb * (kTPix * 128);
          const CUtensorMap* map = is_p ? &tmap_p : &tmap_q;
          const int c = is_p ? p.p.c0 + i0 + bb * 32 : p.q.c0 + jq + bb * 32;
          const bool tiled = is_p ? p.p_tiled != 0 : p.q_tiled != 0;
          if (PAIR) {
            if (tiled) tma_load_2d_pair(dst, map, bar_pair, c, pix0);
            else if (is_p) tma_load_im2col_4d_pair(dst, map, bar_pair, c, pw, ph, img, 0, 0);
            else tma_load_im2col_4d_pair(dst, map, bar_pair, c, qw, qh, img, tap_fs, tap_fr);
          } else {
            if (tiled) tma_load_2d(dst, map, &full_bar[s], c, pix0);
            else if (is_p) tma_load_im2col_4d(dst, map, &full_bar[s], c, pw, ph, img, 0, 0);
            else tma_load_im2col_4d(dst, map, &full_bar[s], c, qw, qh, img, tap_fs, tap_fr);
          }
        }
      }
    }
    __syncwarp();
  }
  if (warp == 5) {
    // ---------------------------------------------------------- MMA issuer
    // both operands MN-major; a pair's MMA spans 256 rows (128 per CTA)
    constexpr uint32_t idesc = make_idesc_tf32(PAIR ? 2 * kBM : kBM, BN, 1, 1);
    if (!PAIR || cta_rank == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t phase = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t pa = smem_base + L::kPOff + s * L::kPBytes;
          const uint32_t qa = smem_base + L::kQOff + s * L::kQBytes;
#pragma unroll
          for (int k = 0; k < kTPix / 8; ++k) {
            // 8 pixels = two 4-row (512 B) swizzle atoms per 32-channel block (stride byte
            // offset); blocks are kTPix * 128 bytes apart (leading byte offset)
            const uint64_t da = make_smem_desc(pa + k * 1024, kTPix * 128, 512, 1);
            const uint64_t db = make_smem_desc(qa + k * 1024, kTPix * 128, 512, 1);
            const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
            if (p.probe == 2 && kb != 0) continue;
            if (PAIR) umma_tf32_pair(tmem_base, da, db, idesc, accum);
            else umma_tf32(tmem_base, da, db, idesc, accum);
          }
          if (PAIR) umma_commit_pair(&empty_bar[s], (uint16_t)3);
          else umma_commit(&empty_bar[s]);
        }
        __syncwarp();
      }
      if (lane == 0) {
        if (PAIR) umma_commit_pair(tmem_full_bar, (uint16_t)3);
        else umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------ epilogue
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int row = i0 + warp * 32 + lane;
    const bool row_ok = row < p.rows;
    const float sc = (row_ok && p.row_scale) ? __ldg(p.row_scale + row) : 1.0f;
    float* out_row = p.gw + (size_t)row * p.gw_ld + p.gw_col0 + tap * p.cols;
    const bool vec_ok = ((p.gw_ld & 3) == 0) && ((p.gw_col0 & 3) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.gw) & 15) == 0);
#pragma unroll 1
    for (int chunk = 0; chunk < BN / 32; ++chunk) {
      const int jc = j0 + chunk * 32;
      if (jc >= p.cols) break;
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(chunk * 32), v);
      tmem_ld_wait();
      if (!row_ok) continue;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int j = jc + g * 4;
        if (j >= p.cols) break;
        if (p.fixed) {
          long long* frow = reinterpret_cast<long long*>(p.gw) + (size_t)row * p.gw_ld +
                            p.gw_col0 + tap * p.cols;
          for (int e = 0; e < 4 && j + e < p.cols; ++e)
            red_fixed(frow + j + e, __uint_as_float(v[g * 4 + e]) * sc);
        } else if (vec_ok && j + 3 < p.cols) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out_row + j),
                       "f"(__uint_as_float(v[g * 4 + 0]) * sc),
                       "f"(__uint_as_float(v[g * 4 + 1]) * sc),
                       "f"(__uint_as_float(v[g * 4 + 2]) * sc),
                       "f"(__uint_as_float(v[g * 4 + 3]) * sc)
                       : "memory");
        } else {
          for (int e = 0; e < 4 && j + e < p.cols; ++e)
            atomicAdd(out_row + j + e, __uint_as_float(v[g * 4 + e]) * sc);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // no CTA leaves (or frees TMEM) while its peer still works
  if (warp == 5) {
    if (PAIR) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Round-up magic number for 31-bit dividends: q = (n * mul) >> (32 + shr).
FastDiv make_fast_div(int d) {
  FastDiv f;
  f.d = (unsigned int)d;
  f.mul = 0;
  f.shr = 0;
  if (d <= 1) return f;
  unsigned int l = 0;
  while ((1u << l) < (unsigned int)d) ++l;               // ceil(log2 d)
  const unsigned long long m = ((1ull << (31 + l)) + d - 1) / d;   // fits in 32 bits
  f.mul = (unsigned int)m;
  f.shr = l - 1;
  return f;
}

template <int BN, int STAGES>
int launch_wgrad(const WgradParams& p, int splits, int taps, cudaStream_t st) {
  using L = WSmem<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    CMR_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_tc_kernel<BN, STAGES>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      L::kDynamic));
    configured = true;
  }
  dim3 grid(ceil_div(p.rows, kBM), ceil_div(p.cols, BN), splits * taps);
  prof_begin(kProfWgrad, 2.0 * p.M * (double)p.rows * p.cols * taps, st);
  conv_wgrad_tc_kernel<BN, STAGES><<<grid, kThreads, L::kDynamic, st>>>(p);
  prof_end(st);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

template <int BN, int STAGES, bool PAIR>
int launch_wgrad_tma(const CUtensorMap& tp, const CUtensorMap& tq, const WgradParams& p,
                     int splits, int taps, cudaStream_t st) {
  using L = TSmem<BN, STAGES, PAIR>;
  static bool configured = false;
  if (!configured) {
    CMR_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_tma_kernel<BN, STAGES, PAIR>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      L::kDynamic));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ceil_div(p.rows, kBM), ceil_div(p.cols, BN), splits * taps);
  cfg.blockDim = dim3(kTmaThreads);
  cfg.dynamicSmemBytes = L::kDynamic;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  prof_begin(kProfWgrad, 2.0 * p.M * (double)p.rows * p.cols * taps, st);
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_wgrad_tma_kernel<BN, STAGES, PAIR>, tp, tq, p);
  prof_end(st);
  CMR_CUDA_TRY(e);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

// Splits of the pixel reduction: fill `slots` concurrent CTAs in as few full waves as
// possible (a partial last wave idles SMs), never fewer than 8 k-blocks per CTA.
int choose_splits(int tiles, int slots, int num_kb) {
  const int max_splits = ceil_div(num_kb, 8) < 1 ? 1 : ceil_div(num_kb, 8);
  int best = 1;
  double best_eff = 0.0;
  for (int waves = 1; waves <= 4; ++waves) {
    int s = (waves * slots) / tiles;
    if (s < 1) continue;
    if (s > max_splits) s = max_splits;
    const long long ctas = (long long)tiles * s;
    const double eff = (double)ctas / ((double)ceil_div_ll(ctas, slots) * slots);
    if (eff > best_eff + 0.04) {   // a later wave count must be clearly better
      best_eff = eff;
      best = s;
    }
    if (s == max_splits) break;
  }
  return best;
}

}  // namespace

// conv_tc.cu
int make_tmap_tiled_2d(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols,
                       uint64_t ld, uint32_t box_rows, bool atom32);
int make_tmap_im2col(CUtensorMap* map, const float* base, int batch, int h, int w, int ld,
                     int channels, int stride, int lower_h, int lower_w, int n_pos_h,
                     int n_pos_w, int pixels, bool mn_major);
extern int g_im2col_tma;
}  // namespace cmr

using namespace cmr;

namespace {
int wgrad_impl(const cmr_wgrad_desc* c, const float* gy, const float* x, float* gw, int fixed,
               const float* row_scale, void* stream);
}

extern "C" int cmr_conv_wgrad_tc(const cmr_wgrad_desc* c, const float* gy, const float* x,
                                 float* gw, const float* row_scale, void* stream) {
  return wgrad_impl(c, gy, x, gw, 0, row_scale, stream);
}

extern "C" int cmr_conv_wgrad_tc_fixed(const cmr_wgrad_desc* c, const float* gy, const float* x,
                                       long long* gw_fixed, const float* row_scale,
                                       void* stream) {
  return wgrad_impl(c, gy, x, reinterpret_cast<float*>(gw_fixed), 1, row_scale, stream);
}

namespace {
int wgrad_impl(const cmr_wgrad_desc* c, const float* gy, const float* x, float* gw, int fixed,
               const float* row_scale, void* stream) {
  CMR_REQUIRE(c && gy && x && gw);
  CMR_REQUIRE(c->batch > 0 && c->loop_h > 0 && c->loop_w > 0 && c->rows > 0 && c->cols > 0);
  CMR_REQUIRE(c->gy_h > 0 && c->gy_w > 0 && c->x_h > 0 && c->x_w > 0);
  CMR_REQUIRE(c->gy_stride >= 1 && c->x_stride >= 1);
  if (c->cols % 4 != 0 || c->x_ld % 4 != 0 || c->gy_ld % 4 != 0 || c->x_c0 % 4 != 0 ||
      c->gy_c0 % 4 != 0)
    return CMR_ERR_UNSUPPORTED;
  CMR_REQUIRE(((reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(x)) & 15) == 0);
  const long long M = (long long)c->batch * c->loop_h * c->loop_w;
  CMR_REQUIRE(M < (1ll << 31));
  CMR_REQUIRE((long long)c->batch * c->gy_h * c->gy_w < (1ll << 31));
  CMR_REQUIRE((long long)c->batch * c->x_h * c->x_w < (1ll << 31));
  WgradParams p;
  p.p.base = gy; p.p.h = c->gy_h; p.p.w = c->gy_w; p.p.ld = c->gy_ld;
  p.p.stride = c->gy_stride; p.p.off_y = c->gy_off_y; p.p.off_x = c->gy_off_x; p.p.c0 = c->gy_c0;
  p.q.base = x; p.q.h = c->x_h; p.q.w = c->x_w; p.q.ld = c->x_ld;
  p.q.stride = c->x_stride; p.q.off_y = c->x_off_y; p.q.off_x = c->x_off_x; p.q.c0 = c->x_c0;
  p.loop_h = c->loop_h; p.loop_w = c->loop_w; p.M = (int)M;
  p.div_hw = make_fast_div(c->loop_h * c->loop_w);
  p.div_w = make_fast_div(c->loop_w);
  p.rows = c->rows; p.cols = c->cols;
  p.gw = gw; p.gw_ld = c->gw_ld; p.gw_col0 = c->gw_col0;
  p.fixed = fixed;
  p.row_scale = row_scale;
  p.num_kb = ceil_div(p.M, kPix);
  {
    const char* e = getenv("CMR_WGRAD_PROBE");
    p.probe = e ? atoi(e) : 0;
  }
  const int taps_h = c->taps_h > 1 ? c->taps_h : 1, taps_w = c->taps_w > 1 ? c->taps_w : 1;
  const int taps = taps_h * taps_w;
  CMR_REQUIRE(taps == 1 || ((c->gw_col0 + taps * c->cols) <= c->gw_ld && c->cols % 4 == 0));
  cudaStream_t st = as_stream(stream);

  // Default path: both operands through im2col-mode tensor maps.
  if (g_im2col_tma) {
    CUtensorMap tp, tq;
    // An operand that is read at every pixel of its tensor, in order, is a plain
    // (pixels x channels) matrix: tiled-mode TMA (its rows stream ~1.5x faster through the
    // copy engine than im2col-mode positions).  Everything else goes through im2col mode.
    const bool multi_tap = taps > 1;
    p.p_tiled = c->gy_h == c->loop_h && c->gy_w == c->loop_w && c->gy_stride == 1 &&
                c->gy_off_y == 0 && c->gy_off_x == 0;
    p.q_tiled = !multi_tap && c->x_h == c->loop_h && c->x_w == c->loop_w && c->x_stride == 1 &&
                c->x_off_y == 0 && c->x_off_x == 0;
    const int rc_p =
        p.p_tiled ? make_tmap_tiled_2d(&tp, gy, (uint64_t)M, (uint64_t)(c->gy_c0 + c->rows),
                                       (uint64_t)c->gy_ld, kTPix, true)
                  : make_tmap_im2col(&tp, gy, c->batch, c->gy_h, c->gy_w, c->gy_ld,
                                     c->gy_c0 + c->rows, c->gy_stride, c->gy_off_y, c->gy_off_x,
                                     c->loop_h, c->loop_w, kTPix, true);
    const int rc_q =
        p.q_tiled ? make_tmap_tiled_2d(&tq, x, (uint64_t)M, (uint64_t)(c->x_c0 + c->cols),
                                       (uint64_t)c->x_ld, kTPix, true)
                  : make_tmap_im2col(&tq, x, c->batch, c->x_h, c->x_w, c->x_ld,
                                     c->x_c0 + c->cols, c->x_stride, c->x_off_y, c->x_off_x,
                                     c->loop_h, c->loop_w, kTPix, true);
    if (rc_p == CMR_OK && rc_q == CMR_OK) {
      const int bn = c->cols > 128 ? 256 : (c->cols > 64 ? 128 : 64);
      p.num_kb = ceil_div(p.M, kTPix);
      const int tiles = ceil_div(p.rows, kBM) * ceil_div(p.cols, bn) * taps;
      const int slots = sm_count();   // one CTA per SM (96 - 192 KB of operand stages)
      int splits = c->splits > 0 ? c->splits : choose_splits(tiles, slots, p.num_kb);
      CMR_REQUIRE((long long)splits * taps < 65536);
      if (splits > p.num_kb) splits = p.num_kb;
      p.kb_per_split = ceil_div(p.num_kb, splits);
      splits = ceil_div(p.num_kb, p.kb_per_split);
      p.splits = splits;
      p.taps_w = taps_w;
      // CTA pairs (tcgen05.mma.cta_group::2) over two adjacent row tiles
      static int pair_ok = -1;
      if (pair_ok < 0) {
        const char* e = getenv("CMR_WGRAD_PAIR");
        pair_ok = e ? atoi(e) : 1;
      }
      const bool pair = pair_ok && (ceil_div(p.rows, kBM) % 2 == 0);
      if (bn == 256)
        return pair ? launch_wgrad_tma<256, 4, true>(tp, tq, p, splits, taps, st)
                    : launch_wgrad_tma<256, 3, false>(tp, tq, p, splits, taps, st);
      if (bn == 128)
        return pair ? launch_wgrad_tma<128, 5, true>(tp, tq, p, splits, taps, st)
                    : launch_wgrad_tma<128, 4, false>(tp, tq, p, splits, taps, st);
      return launch_wgrad_tma<64, 5, false>(tp, tq, p, splits, taps, st);
    }
    p.num_kb = ceil_div(p.M, kPix);
  }

  // Fallback: cp.async gathers.  Split the pixel reduction so that the grid fills the
  // machine (~2 CTAs per SM).
  const int bn = c->cols > 64 ? 128 : 64;
  const int tiles = ceil_div(p.rows, kBM) * ceil_div(p.cols, bn) * taps;
  int splits = c->splits;
  if (splits <= 0) {
    // one full wave of the 2 CTAs per SM that fit: never spill a few CTAs into a second
    splits = (2 * sm_count()) / tiles;
    const int max_splits = ceil_div(p.num_kb, 8);  // at least 8 k-blocks per CTA
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  CMR_REQUIRE((long long)splits * taps < 65536);
  if (splits > p.num_kb) splits = p.num_kb;
  p.kb_per_split = ceil_div(p.num_kb, splits);
  splits = ceil_div(p.num_kb, p.kb_per_split);
  p.splits = splits;
  p.taps_w = taps_w;
  if (bn == 128) return launch_wgrad<128, 3>(p, splits, taps, st);
  return launch_wgrad<64, 4>(p, splits, taps, st);
}
}  // namespace
