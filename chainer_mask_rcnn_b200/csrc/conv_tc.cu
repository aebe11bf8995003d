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
// Persistent, warp-specialised CTA (448 threads, one per SM) walking output tiles
// (128 x BN) in n-fastest order:
//   warp  4     producer: one thread issues the TMA loads of a k-block -- the filter
//               tile (cp.async.bulk.tensor.2d) and the activation tile as ONE im2col-mode
//               TMA (cp.async.bulk.tensor.4d...im2col: 128 convolution positions x 32
//               channels of filter tap (fr, fs), padding zero-filled by the copy engine),
//               both landing in the 128B-swizzled K-major layout the tensor core reads
//   warps 0-3   fallback A producers for layouts the im2col tensor map cannot describe
//               (the RGB0-packed stem): cp.async 16 B gathers into the same layout
//   warp  5     one thread issues tcgen05.mma (128 x BN x 8 per instruction) and
//               tcgen05.commit; the accumulator lives in TMEM, double buffered
//               (2 x BN columns) so tile i+1's MMAs overlap tile i's epilogue
//   warps 6-13  epilogue (plus warps 0-3 when the A tiles come by TMA: 12 warps, three
//               per TMEM lane quarter, interleaved over the 32-column chunks):
//               tcgen05.ld (lane = row) of a 32 x 32 chunk -> per-warp XOR-swizzled
//               shared-memory transpose -> full 128 B row segments, 4 rows per store
//               instruction; the residual addend, the ReLU-mask operand and the affine
//               vectors are requested the same way before the accumulator is waited for.
//               One straight-line block per chunk, ~300 instructions (DESIGN.md section 3:
//               the epilogue was instruction-bound at ~850)
// full/empty mbarrier ring of STAGES k-blocks (32 fp32 = one 128 B swizzle row)
// shared by all tiles; tmem_full/tmem_empty barriers per accumulator buffer.
#include <stdlib.h>
#include <string.h>

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
  int tap_cols;   // > 0: column block t = n / tap_cols goes to pixel offset (t >> 1, t & 1)
  const float* scale;
  const float* bias;
  const float* addend;
  const float* mask;
  const float* bcast;       // (M / bcast_group, N): v += bcast[row / bcast_group][n] * bcast_scale
  int bcast_group;
  float bcast_scale;
  int relu, round_out;
  int m_tiles, n_tiles;
  int tma_a;   // A tiles by TMA: 1 = im2col-mode map, 2 = tiled-mode map (plain matrix);
               // 0 = cp.async gathers
  int epi_groups;   // epilogue warps per TMEM lane quarter: 3 with TMA A tiles, else 2
  // K-split tail (see cmr_conv_gemm_tc_ws): tiles >= tail_begin (a multiple of the slot
  // count) are each computed by tail_splits slots, one K-part of tail_kb k-blocks each
  int tail_begin, tail_rounds, tail_splits, tail_items, tail_kb;
  float4* ws;       // partial accumulators of the tail parts
  long long* dbg;   // measurement only: per-CTA wait-cycle counters (cmr_set_conv_debug)
  int probe;        // measurement only (cmr_set_conv_variant): 1 = the epilogue does not store
  double alg_bytes; // host only: algorithmic HBM bytes of the launch (profiling)
  void* ws_base;    // host only: caller's workspace (cmr_conv_gemm_tc_ws), may be NULL
  size_t ws_bytes;
};

constexpr int kBM = 128;
constexpr int kBK = 32;                      // fp32 per k-block = 128 bytes
constexpr int kABytes = kBM * kBK * 4;       // 16 KB
constexpr int kProducerThreads = 128;
constexpr int kEpiWarp0 = 6;                 // first epilogue warp
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kEpiWarp0 + kEpiWarps) * 32;  // 448

template <int BN, int STAGES, bool PAIR = false>
struct SmemLayout {
  static constexpr int kBBytes = (PAIR ? BN / 2 : BN) * kBK * 4;   // a CTA of a pair holds N/2 rows
  static constexpr int kAOff = 0;
  static constexpr int kBOff = STAGES * kABytes;
  static constexpr int kBarOff = kBOff + STAGES * kBBytes;
  static constexpr int kXposeOff = kBarOff + 256;   // (2 * STAGES + 4) barriers + TMEM slot
  // + groups x 4 warps x 4 KB of transpose buffers + slack for 1024 B alignment
  static constexpr int dynamic_bytes(int groups) { return kXposeOff + groups * 16384 + 1024; }
  static_assert((2 * STAGES + 4) * 8 + 16 <= 256, "barrier area");
};

// PAIR: clusters of two CTAs (one SM pair) run ONE 256 x BN tile with tcgen05.mma
// .cta_group::2: each CTA loads its 128 rows of A and BN/2 rows of B (32 KB instead of 48 KB
// per k-block and SM -- the kernel is bound by operand delivery, not by the tensor core),
// the leader CTA issues the MMAs for both, each CTA's TMEM receives its 128 rows and each
// CTA runs its own epilogue.  TMA loads of both CTAs complete on the leader's barriers;
// tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to both.
// EPI: 0 = the common epilogue; 1 = with the row-group broadcast term of
// cmr_conv_gemm_tc_ex; 2 = pixel-shuffle store of a fused 2x2 deconvolution (tap_cols).
// The rare forms are their own instantiations so that the common epilogue keeps its
// register budget (128 registers, no spills).
constexpr int kEpiPlain = 0, kEpiBcast = 1, kEpiTaps = 2;

// K-split tail (cmr_conv_gemm_tc_ws): at most kSplitMax parts per tile, only for reductions of
// at least kSplitMinKb k-blocks
constexpr int kSplitMax = 4;
constexpr int kSplitMinKb = 32;

// Wait-cycle counters and probe bits of cmr_set_conv_debug / cmr_set_conv_variant: compiled
// in only with -DCMR_CONV_INSTRUMENT=1 (build.py: CMR_CONV_INSTRUMENT=1 in the environment);
// the product kernel carries none of it.
#ifndef CMR_CONV_INSTRUMENT
#define CMR_CONV_INSTRUMENT 0
#endif
constexpr bool kInstr = CMR_CONV_INSTRUMENT != 0;


// SPLIT: the instantiation with the K-split tail (cmr_conv_gemm_tc_ws); the common one carries
// none of its code or registers.
template <int BN, int STAGES, bool PAIR, int EPI, bool SPLIT = false>
__global__ void __launch_bounds__(kThreads, 1)   // 4 warps per SM sub-partition: 128 registers
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_a, const ConvGemmParams p) {
  using L = SmemLayout<BN, STAGES, PAIR>;
  constexpr int kTileM = PAIR ? 2 * kBM : kBM;
  constexpr uint32_t kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = raw_addr + pad;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = p.K / kBK;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0;
  const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // i-th work item of this CTA (pair): whole tiles tile0 + i * tile_step below tail_begin,
  // then at most one K-part [kb0, kb1) of a tail tile (part >= 0).  Without a split tail
  // tail_begin == total_tiles.  Every role walks the same sequence.
  auto work_item = [&](int i, int& tile, int& kb0, int& kb1, int& part) -> bool {
    tile = tile0 + i * tile_step;
    kb0 = 0;
    kb1 = num_kb;
    part = -1;
    if (!SPLIT) return tile < total_tiles;
    if (tile < p.tail_begin) return true;
    if (p.tail_splits <= 1 || i != p.tail_rounds || tile0 >= p.tail_items) return false;
    tile = p.tail_begin + tile0 / p.tail_splits;
    part = tile0 % p.tail_splits;
    kb0 = part * p.tail_kb;
    kb1 = min(num_kb, kb0 + p.tail_kb);
    return true;
  };

  if (warp == 4 && lane == 0) {
    prefetch_tensormap(&tmap_b);
    if (p.tma_a) prefetch_tensormap(&tmap_a);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], p.tma_a ? 1 : kProducerThreads + 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], p.epi_groups * 4 * (PAIR ? 2 : 1));
    }
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
  const int ohw = p.out_h * p.out_w;

  if (warp < 4 && !p.tma_a) {
    // ------------------------------------- fallback A producer (cp.async gather)
    const int t = threadIdx.x;
    const int j = t & 7;
    const int r0 = t >> 3;
    const uint32_t dst_off =
        (uint32_t)((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((j ^ (r0 & 7)) << 4));
    const int cpt = p.in_c / kBK;  // k-blocks per filter tap
    uint32_t it = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step) {
      const int m0 = (tile / p.n_tiles) * kBM;
      int pix_base[8], iy0[8], ix0[8];
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
      int fr = 0, fs = 0, cb = 0;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const uint32_t s = it % STAGES;
        const uint32_t phase = (it / STAGES) & 1;
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
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == 4) {
    // ---------------------------------------------------------- TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      long long w_empty = 0;
      const int cpt = p.in_c / kBK;  // k-blocks per filter tap
      int tile, kb0, kb1, part;
      for (int wi = 0; work_item(wi, tile, kb0, kb1, part); ++wi) {
        const int n0 = (tile % p.n_tiles) * BN;
        // top-left input coordinate of the first convolution position of this CTA's rows
        const int m0 = (tile / p.n_tiles) * kTileM + (int)cta_rank * kBM;
        const int img = m0 / ohw;
        const int rem = m0 - img * ohw;
        const int oy = rem / p.out_w;
        const int h0 = oy * p.stride - p.pad, w0 = (rem - oy * p.out_w) * p.stride - p.pad;
        const int tap0 = kb0 / cpt;
        int fr = tap0 / p.kw, fs = tap0 - fr * p.kw, cb = kb0 - tap0 * cpt;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          const uint32_t phase = (it / STAGES) & 1;
          if (kInstr && p.dbg) {
            const long long c0 = clock64();
            mbar_wait(&empty_bar[s], phase ^ 1);
            w_empty += clock64() - c0;
          } else {
            mbar_wait(&empty_bar[s], phase ^ 1);
          }
          if (PAIR) {
            // both CTAs' boxes complete on the leader's barrier, armed with all their bytes
            const uint32_t bar = mapa_cluster(smem_u32(&full_bar[s]), 0);
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * (L::kBBytes + kABytes));
            if (p.tma_a == 2)
              tma_load_2d_pair(smem_base + L::kAOff + s * kABytes, &tmap_a, bar, cb * kBK, m0);
            else
              tma_load_im2col_4d_pair(smem_base + L::kAOff + s * kABytes, &tmap_a, bar, cb * kBK,
                                      w0, h0, img, fs, fr);
            if (++cb == cpt) {
              cb = 0;
              if (++fs == p.kw) {
                fs = 0;
                ++fr;
              }
            }
            tma_load_2d_pair(smem_base + L::kBOff + s * L::kBBytes, &tmap_b, bar, kb * kBK,
                             n0 + (int)cta_rank * (BN / 2));
            continue;
          }
          if (p.tma_a) {
            mbar_arrive_expect_tx(&full_bar[s], L::kBBytes + kABytes);
            if (p.tma_a == 2)   // 1x1 stride 1: the activations are a plain (M, C) matrix
              tma_load_2d(smem_base + L::kAOff + s * kABytes, &tmap_a, &full_bar[s], cb * kBK,
                          m0);
            else
              tma_load_im2col_4d(smem_base + L::kAOff + s * kABytes, &tmap_a, &full_bar[s],
                                 cb * kBK, w0, h0, img, fs, fr);
            if (++cb == cpt) {
              cb = 0;
              if (++fs == p.kw) {
                fs = 0;
                ++fr;
              }
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[s], L::kBBytes);
          }
          tma_load_2d(smem_base + L::kBOff + s * L::kBBytes, &tmap_b, &full_bar[s], kb * kBK,
                      n0);
        }
      }
      if (kInstr && p.dbg) p.dbg[blockIdx.x * 8 + 3] = w_empty;
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_tf32(kTileM, BN, 0, 0);
    uint32_t it = 0, tc_ = 0;
    long long w_tmem = 0, w_full = 0;
    const long long t_start = (kInstr && p.dbg) ? clock64() : 0;
    if (!PAIR || cta_rank == 0) {   // the leader CTA issues for the pair
    int tile, kb0, kb1, part;
    for (int wi = 0; work_item(wi, tile, kb0, kb1, part); ++wi, ++tc_) {
      const uint32_t buf = tc_ & 1;
      if (kInstr && p.dbg) {
        const long long c0 = clock64();
        mbar_wait(&tmem_empty_bar[buf], ((tc_ >> 1) & 1) ^ 1);
        w_tmem += clock64() - c0;
      } else {
        mbar_wait(&tmem_empty_bar[buf], ((tc_ >> 1) & 1) ^ 1);
      }
      tc_fence_after();
      const uint32_t acc = tmem_base + buf * BN;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const uint32_t s = it % STAGES;
        const uint32_t phase = (it / STAGES) & 1;
        if (kInstr && p.dbg) {
          const long long c0 = clock64();
          mbar_wait(&full_bar[s], phase);
          w_full += clock64() - c0;
        } else {
          mbar_wait(&full_bar[s], phase);
        }
        tc_fence_after();
        if (lane == 0) {
          const uint64_t da = make_smem_desc_sw128(smem_base + L::kAOff + s * kABytes, 16, 1024);
          const uint64_t db =
              make_smem_desc_sw128(smem_base + L::kBOff + s * L::kBBytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {  // 8 tf32 = 32 bytes per MMA -> +2 (16 B units)
            const uint32_t accumulate = (kb != kb0 || k != 0) ? 1u : 0u;
            if (PAIR) umma_tf32_pair(acc, da + 2 * k, db + 2 * k, idesc, accumulate);
            else umma_tf32(acc, da + 2 * k, db + 2 * k, idesc, accumulate);
          }
          if (PAIR) umma_commit_pair(&empty_bar[s], (uint16_t)3);
          else umma_commit(&empty_bar[s]);
        }
        __syncwarp();
      }
      if (lane == 0) {
        if (PAIR) umma_commit_pair(&tmem_full_bar[buf], (uint16_t)3);
        else umma_commit(&tmem_full_bar[buf]);
      }
      __syncwarp();
    }
    }
    if (kInstr && p.dbg && lane == 0) {
      p.dbg[blockIdx.x * 8 + 0] = w_tmem;
      p.dbg[blockIdx.x * 8 + 1] = w_full;
      p.dbg[blockIdx.x * 8 + 2] = clock64() - t_start;
      p.dbg[blockIdx.x * 8 + 6] = tc_;
    }
  } else {
    // ------------------------------------------------------------- epilogue
    // Written for instruction count: three epilogue warps per SM sub-partition cannot hide
    // much, so a 32 x 32 chunk is ONE straight-line block -- operand loads, accumulator
    // load, transpose through shared memory, then the arithmetic of all eight row groups
    // with every optional step behind a single warp-uniform branch per chunk.
    const int q = warp & 3;                   // TMEM lane quarter this warp may read
    // column group of this warp; n_groups warps share a lane quarter
    const int n_groups = p.epi_groups;
    const int grp = warp < 4 ? 0 : (warp - kEpiWarp0) / 4 + (p.tma_a ? 1 : 0);
    if (grp < n_groups) {
    float4* xp4 = reinterpret_cast<float4*>(smem + L::kXposeOff) + (grp * 4 + q) * (32 * 8);
    const int sub = lane >> 3;                // row within a 4-row store group
    const int c4 = (lane & 7) * 4;            // first of this lane's 4 columns in a chunk
    const float* __restrict__ scale_p = p.scale;
    const float* __restrict__ bias_p = p.bias;
    const float* __restrict__ addend_p = p.addend;
    const float* __restrict__ mask_p = p.mask;
    float* __restrict__ d_p = p.d;
    const bool relu = p.relu != 0, round_out = p.round_out != 0;
    const bool affine = scale_p != nullptr || bias_p != nullptr;
    const bool vec_ok = ((p.d_ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(d_p) & 15) == 0) &&
                        (!addend_p || (reinterpret_cast<uintptr_t>(addend_p) & 15) == 0) &&
                        (!mask_p || (reinterpret_cast<uintptr_t>(mask_p) & 15) == 0);
    // the output tensor is the (M, d_ld) matrix itself: row offsets need no division
    const bool linear = p.d_stride == 1 && p.d_oy == 0 && p.d_ox == 0 && p.d_h == p.out_h &&
                        p.d_w == p.out_w;
    // transpose buffer: this lane writes row `lane` (16-byte slots XOR-swizzled with the
    // row: conflict-free both ways, no padding) and reads rows 4 i + sub; the slot of an
    // even / odd row group differs in bit 2 only
    float4* const xw = xp4 + lane * 8;
    const float4* const xr0 = xp4 + sub * 8 + ((lane & 7) ^ sub);
    const float4* const xr1 = xp4 + (4 + sub) * 8 + ((lane & 7) ^ (4 + sub));
    constexpr int kChunks = BN / 32;          // 32-column chunks of the tile
    uint32_t tc_ = 0;
    long long w_acc = 0;
    const long long t_start = (kInstr && p.dbg) ? clock64() : 0;
    int tile, kb0, kb1, part;
    for (int wi = 0; work_item(wi, tile, kb0, kb1, part); ++wi, ++tc_) {
      const int m0 = (tile / p.n_tiles) * kTileM + (int)cta_rank * kBM;
      const int n0 = (tile % p.n_tiles) * BN;
      const uint32_t buf = tc_ & 1;
      // output element offsets of this lane's 8 rows (row = 32q + 4i + sub); rows past M
      // read the operands of row M - 1 and store nothing.  The host checks that the output
      // tensor has fewer than 2^31 elements.
      const int row0 = m0 + q * 32 + sub;
      const bool rows_ok = m0 + q * 32 + 32 <= p.M;        // warp-uniform
      auto row_offset = [&](int row) -> int {
        if (linear) return row * p.d_ld;
        const int img = row / ohw;
        const int rem = row - img * ohw;
        const int oy = rem / p.out_w;
        const int ox = rem - oy * p.out_w;
        return ((img * p.d_h + oy * p.d_stride + p.d_oy) * p.d_w + ox * p.d_stride + p.d_ox) *
               p.d_ld;
      };
      int doff[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) doff[i] = row_offset(min(row0 + 4 * i, p.M - 1));
      bool waited = false;
#pragma unroll 1
      for (int chunk = grp; chunk < kChunks; chunk += n_groups) {
        const int cbase = chunk * 32;                     // column of the tile
        if (n0 + cbase >= p.N) break;
        const int ng = n0 + cbase + c4;                   // this lane's first GEMM column
        // pixel-shuffle store of a fused 2x2 stride-2 deconvolution: column block t =
        // ng / tap_cols holds tap (t >> 1, t & 1); n / toff address the output tensor
        const int tap = EPI == kEpiTaps ? (n0 + cbase) / p.tap_cols : 0;
        const int cn = ((tap >> 1) * p.d_w + (tap & 1)) * p.d_ld + ng - tap * p.tap_cols;
        const int n = ng - tap * p.tap_cols;              // channel in the output tensor
        const bool full4 = vec_ok && (ng + 3 < p.N);
        // operands of the epilogue are requested before the accumulator is waited for
        // (coalesced: 8 lanes cover one 128 B row)
        float4 ad[8], mk[8];
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), bi = make_float4(0.f, 0.f, 0.f, 0.f);
        if (full4) {
          if (addend_p) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              ad[i] = __ldg(reinterpret_cast<const float4*>(addend_p + doff[i] + cn));
          }
          if (mask_p) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              mk[i] = __ldg(reinterpret_cast<const float4*>(mask_p + doff[i] + cn));
          }
          if (EPI == kEpiBcast) {
            // ad[] carries the broadcast rows: raw when there is no addend (scaled below),
            // else folded into the addend here
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int grp_row = min(row0 + 4 * i, p.M - 1) / p.bcast_group;
              const float4 bv =
                  __ldg(reinterpret_cast<const float4*>(p.bcast + (size_t)grp_row * p.N + n));
              if (addend_p) {
                ad[i].x = fmaf(bv.x, p.bcast_scale, ad[i].x); ad[i].y = fmaf(bv.y, p.bcast_scale, ad[i].y);
                ad[i].z = fmaf(bv.z, p.bcast_scale, ad[i].z); ad[i].w = fmaf(bv.w, p.bcast_scale, ad[i].w);
              } else {
                ad[i] = bv;
              }
            }
          }
          if (scale_p) sc = __ldg(reinterpret_cast<const float4*>(scale_p + n));
          if (bias_p) bi = __ldg(reinterpret_cast<const float4*>(bias_p + n));
        }
        if (!waited) {
          if (kInstr && p.dbg) {
            const long long c0 = clock64();
            mbar_wait(&tmem_full_bar[buf], (tc_ >> 1) & 1);
            w_acc += clock64() - c0;
          } else {
            mbar_wait(&tmem_full_bar[buf], (tc_ >> 1) & 1);
          }
          tc_fence_after();
          waited = true;
        }
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + (uint32_t)cbase, v);
        tmem_ld_wait();
        __syncwarp();   // previous chunk's reads of the transpose buffer are done
#pragma unroll
        for (int c = 0; c < 8; ++c)
          xw[c ^ (lane & 7)] =
              make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]),
                          __uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3]));
        __syncwarp();
        if (full4) {
          float4 o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = (i & 1) ? xr1[(i - 1) * 32] : xr0[i * 32];
          if (SPLIT && part >= 0) {
            // K-part of a tail tile: the raw partial sums go to the workspace in this layout
            // ((tail tile, part, CTA of the pair, lane quarter, chunk) blocks of 8 x 32 float4);
            // conv_tail_finish_kernel adds the parts and runs the epilogue
            float4* dst = p.ws + ((((size_t)(tile - p.tail_begin) * p.tail_splits + part) *
                                       (PAIR ? 2 : 1) + cta_rank) * 4 + q) * (kChunks * 256) +
                          chunk * 256 + lane;
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i * 32] = o[i];
            continue;
          }
          if (affine) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              o[i].x = fmaf(o[i].x, sc.x, bi.x); o[i].y = fmaf(o[i].y, sc.y, bi.y);
              o[i].z = fmaf(o[i].z, sc.z, bi.z); o[i].w = fmaf(o[i].w, sc.w, bi.w);
            }
          }
          if (addend_p) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              o[i].x += ad[i].x; o[i].y += ad[i].y; o[i].z += ad[i].z; o[i].w += ad[i].w;
            }
          }
          if (EPI == kEpiBcast && !addend_p) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              o[i].x = fmaf(ad[i].x, p.bcast_scale, o[i].x); o[i].y = fmaf(ad[i].y, p.bcast_scale, o[i].y);
              o[i].z = fmaf(ad[i].z, p.bcast_scale, o[i].z); o[i].w = fmaf(ad[i].w, p.bcast_scale, o[i].w);
            }
          }
          if (relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              o[i].x = fmaxf(o[i].x, 0.f); o[i].y = fmaxf(o[i].y, 0.f);
              o[i].z = fmaxf(o[i].z, 0.f); o[i].w = fmaxf(o[i].w, 0.f);
            }
          }
          if (mask_p) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              o[i].x = mk[i].x > 0.f ? o[i].x : 0.f; o[i].y = mk[i].y > 0.f ? o[i].y : 0.f;
              o[i].z = mk[i].z > 0.f ? o[i].z : 0.f; o[i].w = mk[i].w > 0.f ? o[i].w : 0.f;
            }
          }
          if (round_out) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              o[i].x = round_tf32(o[i].x); o[i].y = round_tf32(o[i].y);
              o[i].z = round_tf32(o[i].z); o[i].w = round_tf32(o[i].w);
            }
          }
          if (!(kInstr && (p.probe & 1))) {
            if (rows_ok) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4*>(d_p + doff[i] + cn) = o[i];
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (row0 + 4 * i < p.M) *reinterpret_cast<float4*>(d_p + doff[i] + cn) = o[i];
            }
          }
        } else {
          // ragged / unaligned tail columns: scalar path
          const float* xs = reinterpret_cast<const float*>(xp4);
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {
            if (row0 + 4 * i >= p.M) continue;
            const int r = 4 * i + sub;
            const int off = row_offset(row0 + 4 * i) + cn;
            for (int e = 0; e < 4 && ng + e < p.N; ++e) {
              float x = xs[r * 32 + ((((lane & 7) ^ (r & 7))) << 2) + e];
              if (scale_p) x *= __ldg(scale_p + n + e);
              if (bias_p) x += __ldg(bias_p + n + e);
              if (addend_p) x += __ldg(addend_p + off + e);
              if (EPI == kEpiBcast)
                x += __ldg(p.bcast + (size_t)((row0 + 4 * i) / p.bcast_group) * p.N + n + e) *
                     p.bcast_scale;
              if (relu) x = fmaxf(x, 0.f);
              if (mask_p) x = __ldg(mask_p + off + e) > 0.f ? x : 0.f;
              if (round_out) x = round_tf32(x);
              d_p[off + e] = x;
            }
          }
        }
      }
      if (!waited) {   // this warp had no columns to write: still consume the phase
        mbar_wait(&tmem_full_bar[buf], (tc_ >> 1) & 1);
        tc_fence_after();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(mapa_cluster(smem_u32(&tmem_empty_bar[buf]), 0));
        else mbar_arrive(&tmem_empty_bar[buf]);
      }
    }
    if (kInstr && p.dbg && warp == kEpiWarp0 && lane == 0) {
      p.dbg[blockIdx.x * 8 + 4] = w_acc;
      p.dbg[blockIdx.x * 8 + 5] = clock64() - t_start;
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

// K-split tail, second step: adds the K-parts of every tail tile in part order and runs the
// epilogue of conv_gemm_tc_kernel on the sums (same operations in the same order).  One thread
// per float4 of the workspace layout written there.
__global__ void __launch_bounds__(256)
conv_tail_finish_kernel(const ConvGemmParams p, int bn, int ranks, int n_tail_tiles) {
  const int chunks = bn / 32;
  const size_t total = (size_t)n_tail_tiles * ranks * 4 * chunks * 256;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int lane = (int)(e & 31);
  const int i = (int)((e >> 5) & 7);
  size_t r = e >> 8;
  const int chunk = (int)(r % chunks); r /= chunks;
  const int q = (int)(r & 3); r >>= 2;
  const int rank = (int)(r % ranks);
  const int tt = (int)(r / ranks);
  const int tile = p.tail_begin + tt;
  const int row = (tile / p.n_tiles) * (kBM * ranks) + rank * kBM + q * 32 + 4 * i + (lane >> 3);
  const int ng = (tile % p.n_tiles) * bn + chunk * 32 + (lane & 7) * 4;
  if (row >= p.M || ng >= p.N) return;
  const size_t part_stride = (size_t)ranks * 4 * chunks * 256;
  const float4* src = p.ws + (size_t)tt * p.tail_splits * part_stride +
                      ((size_t)(rank * 4 + q) * chunks + chunk) * 256 + i * 32 + lane;
  float4 o = __ldcg(src);
  for (int pp = 1; pp < p.tail_splits; ++pp) {
    const float4 u = __ldcg(src + pp * part_stride);
    o.x += u.x; o.y += u.y; o.z += u.z; o.w += u.w;
  }
  const int tap = p.tap_cols > 0 ? (ng - (lane & 7) * 4) / p.tap_cols : 0;
  const int n = ng - tap * p.tap_cols;
  int off;
  {
    const int ohw = p.out_h * p.out_w;
    const int img = row / ohw;
    const int rem = row - img * ohw;
    const int oy = rem / p.out_w;
    const int ox = rem - oy * p.out_w;
    off = ((img * p.d_h + oy * p.d_stride + p.d_oy) * p.d_w + ox * p.d_stride + p.d_ox) * p.d_ld +
          ((tap >> 1) * p.d_w + (tap & 1)) * p.d_ld + n;
  }
  if (p.scale || p.bias) {
    const float4 sc = p.scale ? __ldg(reinterpret_cast<const float4*>(p.scale + n))
                              : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 bi = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
    o.x = fmaf(o.x, sc.x, bi.x); o.y = fmaf(o.y, sc.y, bi.y);
    o.z = fmaf(o.z, sc.z, bi.z); o.w = fmaf(o.w, sc.w, bi.w);
  }
  float4 ad = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.addend) ad = __ldg(reinterpret_cast<const float4*>(p.addend + off));
  if (p.bcast) {
    const float4 bv = __ldg(reinterpret_cast<const float4*>(
        p.bcast + (size_t)(row / p.bcast_group) * p.N + n));
    if (p.addend) {
      ad.x = fmaf(bv.x, p.bcast_scale, ad.x); ad.y = fmaf(bv.y, p.bcast_scale, ad.y);
      ad.z = fmaf(bv.z, p.bcast_scale, ad.z); ad.w = fmaf(bv.w, p.bcast_scale, ad.w);
    } else {
      o.x = fmaf(bv.x, p.bcast_scale, o.x); o.y = fmaf(bv.y, p.bcast_scale, o.y);
      o.z = fmaf(bv.z, p.bcast_scale, o.z); o.w = fmaf(bv.w, p.bcast_scale, o.w);
    }
  }
  if (p.addend) { o.x += ad.x; o.y += ad.y; o.z += ad.z; o.w += ad.w; }
  if (p.relu) {
    o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
  }
  if (p.mask) {
    const float4 mk = __ldg(reinterpret_cast<const float4*>(p.mask + off));
    o.x = mk.x > 0.f ? o.x : 0.f; o.y = mk.y > 0.f ? o.y : 0.f;
    o.z = mk.z > 0.f ? o.z : 0.f; o.w = mk.w > 0.f ? o.w : 0.f;
  }
  if (p.round_out) {
    o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w);
  }
  *reinterpret_cast<float4*>(p.d + off) = o;
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

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeIm2colFn get_encode_im2col_fn() {
  static EncodeIm2colFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(ptr);
  }
  return fn;
}

}  // namespace

// Im2col-mode tensor map over an NHWC fp32 tensor (batch, h, w, ld) of which channels
// [0, channels) are visible.  One TMA instruction loads `pixels` consecutive positions of
// a walk with step `stride` whose first position has top-left input coordinate
// (lower_h, lower_w) and which visits n_pos_h x n_pos_w positions per image (w fastest,
// then h, then the image), x 32 channels, shifted by the im2col offsets (the filter tap).
// Returns CMR_ERR_UNSUPPORTED when the geometry does not fit the descriptor's fields.
int make_tmap_im2col(CUtensorMap* map, const float* base, int batch, int h, int w, int ld,
                     int channels, int stride, int lower_h, int lower_w, int n_pos_h,
                     int n_pos_w, int pixels, bool mn_major) {
  EncodeIm2colFn fn = get_encode_im2col_fn();
  if (!fn) return CMR_ERR_UNSUPPORTED;
  if (stride < 1 || stride > 8 || (ld & 3) != 0 || channels < 1 || channels > ld)
    return CMR_ERR_UNSUPPORTED;
  // extent of the bounding box along an axis so that ceil(extent / stride) == n_pos
  const int upper_h = (n_pos_h - 1) * stride + 1 - h + lower_h;
  const int upper_w = (n_pos_w - 1) * stride + 1 - w + lower_w;
  const int lim[4] = {lower_h, lower_w, upper_h, upper_w};
  for (int i = 0; i < 4; ++i)
    if (lim[i] < -128 || lim[i] > 127) return CMR_ERR_UNSUPPORTED;
  cuuint64_t gdim[4] = {(cuuint64_t)channels, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)batch};
  cuuint64_t gstride[3] = {(cuuint64_t)ld * 4, (cuuint64_t)w * ld * 4,
                           (cuuint64_t)h * w * ld * 4};
  int lower[2] = {lower_w, lower_h};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim,
                  gstride, lower, upper, 32, (cuuint32_t)pixels, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CMR_OK : CMR_ERR_UNSUPPORTED;
}

int g_im2col_tma = 1;   // cmr_set_im2col_tma
int g_conv_variant = 0; // cmr_set_conv_variant
long long* g_conv_dbg = nullptr;   // cmr_set_conv_debug

namespace {

}  // namespace

// 2-D fp32 matrix (rows, cols) with row pitch ld, box = (box_rows, 32 cols); 128 B swizzle
// of 16 B units (K-major operands) or of 32 B units (atom32: MN-major operands).
int make_tmap_tiled_2d(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols,
                       uint64_t ld, uint32_t box_rows, bool atom32) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn || (ld & 3) != 0 || cols > ld || box_rows > 256) return CMR_ERR_UNSUPPORTED;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)kBK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim,
                  gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CMR_OK : CMR_ERR_UNSUPPORTED;
}

namespace {

template <int BN, int STAGES, bool PAIR, int EPI>
int launch_b(const CUtensorMap& tmap, const CUtensorMap& tmap_a, const ConvGemmParams& p,
             cudaStream_t st) {
  using L = SmemLayout<BN, STAGES, PAIR>;
  static bool configured = false;
  if (!configured) {
    CMR_CUDA_TRY(cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, STAGES, PAIR, EPI>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      L::dynamic_bytes(3) <= 232448 ? L::dynamic_bytes(3)
                                                                    : L::dynamic_bytes(2)));
    configured = true;
  }
  CMR_REQUIRE(L::dynamic_bytes(p.epi_groups) <= 232448);
  ConvGemmParams q = p;
  q.m_tiles = ceil_div(p.M, PAIR ? 2 * kBM : kBM);
  q.n_tiles = ceil_div(p.N, BN);
  const long long tiles = (long long)q.m_tiles * q.n_tiles;
  const int slots = PAIR ? sm_count() / 2 : sm_count();
  const int grid = (int)(tiles < slots ? tiles : slots) * (PAIR ? 2 : 1);
  // K-split tail: when the tiles leave a last wave that fills at most half of the slots, each
  // of its tiles is computed by several slots (a K range each); the partial sums go through
  // the caller's workspace and conv_tail_finish_kernel adds them and runs the epilogue -- the
  // partial wave then lasts 1 / splits of a tile (+ the small second launch)
  q.tail_begin = (int)tiles;
  q.tail_rounds = 0; q.tail_splits = 1; q.tail_items = 0; q.tail_kb = 0;
  q.ws = nullptr;
  const int num_kb = p.K / kBK;
  const bool aligned = (p.d_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.d) & 15) == 0 &&
                       (!p.addend || (reinterpret_cast<uintptr_t>(p.addend) & 15) == 0) &&
                       (!p.mask || (reinterpret_cast<uintptr_t>(p.mask) & 15) == 0);
  long long rem = 0;
  constexpr bool kHasSplit = PAIR && STAGES == 6;   // (the long-reduction pair kernels)
  if (kHasSplit && p.ws_base && p.tma_a && aligned && p.N % 32 == 0 &&
      num_kb >= kSplitMinKb) {
    const long long full = tiles / slots;
    rem = tiles % slots;
    if (full >= 1 && rem > 0 && rem * 2 <= slots) {
      int splits = (int)(slots / rem) < kSplitMax ? (int)(slots / rem) : kSplitMax;
      const int tail_kb = ceil_div(num_kb, splits);
      splits = ceil_div(num_kb, tail_kb);
      const size_t need = (size_t)rem * splits * (PAIR ? 2 : 1) * kBM * BN * 4;
      if (splits > 1 && need <= p.ws_bytes) {
        q.tail_begin = (int)(full * slots);
        q.tail_rounds = (int)full;
        q.tail_splits = splits;
        q.tail_items = (int)rem * splits;
        q.tail_kb = tail_kb;
        q.ws = reinterpret_cast<float4*>(p.ws_base);
      }
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = L::dynamic_bytes(p.epi_groups);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const double flops = 2.0 * p.M * (double)p.N * p.K;
  prof_begin(kProfConvGemm, flops, st);
  // the launch is tensor-bound when its arithmetic intensity exceeds the machine balance
  // (TF32 ~706 TFLOP/s over ~6.45 TB/s measured on this pool: 110 FLOP per byte)
  if (flops >= 110.0 * p.alg_bytes) prof_tag(kProfConvTensorBound, flops);
  else prof_tag(kProfConvHbmBound, p.alg_bytes);
  cudaError_t e;
  if constexpr (kHasSplit) {
    if (q.tail_splits > 1) {
      static bool configured_split = false;
      if (!configured_split) {
        CMR_CUDA_TRY(cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, STAGES, PAIR, EPI, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          L::dynamic_bytes(3) <= 232448 ? L::dynamic_bytes(3)
                                                                        : L::dynamic_bytes(2)));
        configured_split = true;
      }
      e = cudaLaunchKernelEx(&cfg, conv_gemm_tc_kernel<BN, STAGES, PAIR, EPI, true>, tmap, tmap_a,
                             q);
    } else {
      e = cudaLaunchKernelEx(&cfg, conv_gemm_tc_kernel<BN, STAGES, PAIR, EPI, false>, tmap,
                             tmap_a, q);
    }
  } else {
    e = cudaLaunchKernelEx(&cfg, conv_gemm_tc_kernel<BN, STAGES, PAIR, EPI, false>, tmap, tmap_a,
                           q);
  }
  if (e == cudaSuccess && q.tail_splits > 1) {
    const int ranks = PAIR ? 2 : 1;
    const size_t total = (size_t)rem * ranks * 4 * (BN / 32) * 256;
    conv_tail_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(q, BN, ranks, (int)rem);
  }
  prof_end(st);
  CMR_CUDA_TRY(e);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

template <int BN, int STAGES, bool PAIR>
int launch(const CUtensorMap& tmap, const CUtensorMap& tmap_a, const ConvGemmParams& p,
           cudaStream_t st) {
  if (p.bcast) return launch_b<BN, STAGES, PAIR, kEpiBcast>(tmap, tmap_a, p, st);
  if (p.tap_cols > 0) return launch_b<BN, STAGES, PAIR, kEpiTaps>(tmap, tmap_a, p, st);
  return launch_b<BN, STAGES, PAIR, kEpiPlain>(tmap, tmap_a, p, st);
}

}  // namespace
}  // namespace cmr

using namespace cmr;

extern "C" int cmr_set_im2col_tma(int on) {
  const int old = g_im2col_tma;
  g_im2col_tma = on != 0;
  return old;
}

extern "C" int cmr_set_conv_debug(long long* buf) {
  g_conv_dbg = buf;
  return CMR_OK;
}

extern "C" int cmr_set_conv_variant(int v) {
  const int old = g_conv_variant;
  g_conv_variant = v;
  return old;
}

extern "C" int cmr_conv_gemm_tc(const cmr_conv_desc* c, const float* a, const float* w, float* d,
                                const float* scale, const float* bias, const float* addend,
                                const float* mask, void* stream) {
  return cmr_conv_gemm_tc_ex(c, a, w, d, scale, bias, addend, mask, nullptr, 1, 0.f, stream);
}

extern "C" int cmr_conv_gemm_tc_ex(const cmr_conv_desc* c, const float* a, const float* w,
                                   float* d, const float* scale, const float* bias,
                                   const float* addend, const float* mask, const float* bcast,
                                   int bcast_group, float bcast_scale, void* stream) {
  return cmr_conv_gemm_tc_ws(c, a, w, d, scale, bias, addend, mask, bcast, bcast_group,
                             bcast_scale, nullptr, 0, stream);
}

extern "C" size_t cmr_conv_gemm_ws_bytes(void) {
  return (size_t)sm_count() * kBM * 256 * 4;   // one 128 x 256 partial tile per SM
}

extern "C" int cmr_conv_gemm_tc_ws(const cmr_conv_desc* c, const float* a, const float* w,
                                   float* d, const float* scale, const float* bias,
                                   const float* addend, const float* mask, const float* bcast,
                                   int bcast_group, float bcast_scale, void* ws, size_t ws_bytes,
                                   void* stream) {
  CMR_REQUIRE(c && a && w && d);
  CMR_REQUIRE(!bcast || (bcast_group >= 1 && (reinterpret_cast<uintptr_t>(bcast) & 15) == 0 &&
                         c->n % 4 == 0));
  CMR_REQUIRE(c->batch > 0 && c->in_h > 0 && c->in_w > 0 && c->out_h > 0 && c->out_w > 0);
  CMR_REQUIRE(c->kh > 0 && c->kw > 0 && c->stride > 0 && c->pad >= 0 && c->n > 0);
  if (c->in_c <= 0 || c->in_c % kBK != 0) return CMR_ERR_UNSUPPORTED;
  CMR_REQUIRE(c->in_ld > 0 && c->in_ld % 4 == 0);
  CMR_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0);
  CMR_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0);
  const long long M = (long long)c->batch * c->out_h * c->out_w;
  CMR_REQUIRE(M > 0 && M < (1ll << 31));
  CMR_REQUIRE((long long)c->batch * c->in_h * c->in_w < (1ll << 31));
  CMR_REQUIRE((long long)c->batch * c->d_h * c->d_w * c->d_ld < (1ll << 31));
  ConvGemmParams p;
  p.a = a;
  p.in_h = c->in_h; p.in_w = c->in_w; p.in_c = c->in_c; p.in_ld = c->in_ld;
  p.out_h = c->out_h; p.out_w = c->out_w;
  p.kh = c->kh; p.kw = c->kw; p.stride = c->stride; p.pad = c->pad;
  p.M = (int)M; p.N = c->n; p.K = c->kh * c->kw * c->in_c;
  p.d = d;
  p.d_h = c->d_h; p.d_w = c->d_w; p.d_ld = c->d_ld; p.d_stride = c->d_stride;
  p.d_oy = c->d_oy; p.d_ox = c->d_ox;
  p.tap_cols = c->tap_cols;
  p.scale = scale; p.bias = bias; p.addend = addend; p.mask = mask;
  {
    // activations once (the pixels the filter touches), filter once, output once, plus the
    // epilogue operands
    const double in_px = (double)c->batch * c->in_h * c->in_w * c->in_c;
    const double mk = (double)p.M * p.K, mn = (double)p.M * p.N;
    p.alg_bytes = 4.0 * ((in_px < mk ? in_px : mk) + (double)p.N * p.K +
                         mn * (1 + (addend != nullptr) + (mask != nullptr)));
  }
  p.bcast = bcast; p.bcast_group = bcast_group; p.bcast_scale = bcast_scale;
  p.ws_base = ws; p.ws_bytes = ws ? ws_bytes : 0;
  CMR_REQUIRE(!ws || (reinterpret_cast<uintptr_t>(ws) & 15) == 0);
  p.relu = c->relu; p.round_out = c->round_tf32;
  CMR_REQUIRE(p.d_stride >= 1 && p.tap_cols >= 0);
  if (p.tap_cols > 0) {   // fused 2x2 stride-2 deconvolution: four column blocks of tap_cols
    CMR_REQUIRE(p.tap_cols % 32 == 0 && p.N == 4 * p.tap_cols && p.d_ld >= p.tap_cols);
    CMR_REQUIRE(p.d_stride == 2 && !bcast);
    CMR_REQUIRE((c->out_h - 1) * 2 + c->d_oy + 1 < c->d_h && (c->out_w - 1) * 2 + c->d_ox + 1 < c->d_w);
  } else {
    CMR_REQUIRE(p.d_ld >= p.N);
    CMR_REQUIRE((c->out_h - 1) * c->d_stride + c->d_oy < c->d_h);
    CMR_REQUIRE((c->out_w - 1) * c->d_stride + c->d_ox < c->d_w);
  }

  // Tile width: minimise (waves over the SMs) x (per-tile cost); narrower tiles move
  // more operand bytes per FLOP through L2 (relative tile rates 1 : 0.75 : 0.5).
  int bn = c->tile_n;
  if (bn == 0) {
    const int widths[3] = {256, 128, 64};
    const double rate[3] = {1.0, 0.75, 0.5};
    double best = 0.0;
    for (int i = 0; i < 3; ++i) {
      if (i < 2 && p.N <= widths[i] / 2) continue;   // more than half the tile would be padding
      const long long tiles = (long long)ceil_div(p.M, kBM) * ceil_div(p.N, widths[i]);
      const double cost = (double)ceil_div_ll(tiles, sm_count()) * widths[i] / rate[i];
      if (bn == 0 || cost < best) {
        bn = widths[i];
        best = cost;
      }
    }
  }
  // The activation operand through an im2col tensor map when the geometry is a plain
  // convolution over in_c <= in_ld channels (everything but the RGB0-packed stem).
  CUtensorMap tmap_a;
  memset(&tmap_a, 0, sizeof(tmap_a));
  p.tma_a = 0;
  if (g_im2col_tma && c->in_c <= c->in_ld &&
      c->out_h == (c->in_h + 2 * c->pad - c->kh) / c->stride + 1 &&
      c->out_w == (c->in_w + 2 * c->pad - c->kw) / c->stride + 1 &&
      c->in_h + 2 * c->pad >= c->kh && c->in_w + 2 * c->pad >= c->kw && c->kh < 256 &&
      c->kw < 256) {
    if (c->kh == 1 && c->kw == 1 && c->stride == 1 && c->pad == 0 &&
        make_tmap_tiled_2d(&tmap_a, a, (uint64_t)p.M, (uint64_t)c->in_c, (uint64_t)c->in_ld, kBM,
                           false) == CMR_OK)
      p.tma_a = 2;
    else if (make_tmap_im2col(&tmap_a, a, c->batch, c->in_h, c->in_w, c->in_ld, c->in_c,
                              c->stride, -c->pad, -c->pad, c->out_h, c->out_w, kBM,
                              false) == CMR_OK)
      p.tma_a = 1;
  }
  // Epilogue warps per TMEM lane quarter: with TMA A tiles warps 0-3 join the epilogue
  // (3 groups).  For 256-wide tiles the shared memory holds either 4 operand stages and 2
  // groups (long reductions: the main loop dominates) or 3 stages and 3 groups (short
  // reductions: the epilogue's HBM traffic dominates).
  p.epi_groups = p.tma_a ? 3 : 2;
  int variant = g_conv_variant & 15;
  {
    // A/B knob (CMR_CONV_PAIR_K512=1): CTA pairs also for the K = 512 layers with at most one
    // epilogue operand tensor (isolated: res5 conv3 + residual 228 -> 211 us, but 266 -> 304 us
    // with addend AND mask, hence the condition)
    static int pair512 = -1;
    if (pair512 < 0) {
      const char* e = getenv("CMR_CONV_PAIR_K512");
      pair512 = e ? atoi(e) : 0;
    }
    if (pair512 && variant == 0 && p.K >= 512 && p.K < 1024 && p.N >= 1024 && !(addend && mask))
      variant = 1;
  }
  p.dbg = g_conv_dbg;
  p.probe = (g_conv_variant >> 5) & 1;                   // 32: no stores (instrumented builds)
  if (g_conv_variant & 16) p.addend = p.mask = nullptr;  // 16: no epilogue operand loads
  const bool deep = p.K >= 1024 || !p.tma_a || variant == 3 || variant == 2;
  if (bn == 256 && deep) p.epi_groups = 2;
  // CTA pairs (tcgen05.mma.cta_group::2) for the long reductions with enough 256-row tiles
  // to fill the 74 SM pairs
  static int pair_ok = -1;
  if (pair_ok < 0) {
    const char* e = getenv("CMR_CONV_PAIR");
    pair_ok = e ? atoi(e) : 1;
  }
  // (128-wide pair tiles were measured on the res3 / res4 3x3 layers: no gain, those
  // single-wave launches are bound by their prologue / epilogue, not by operand delivery)
  const bool pair = pair_ok && bn == 256 && (deep || variant == 1) && p.tma_a && variant != 3 &&
                    (long long)ceil_div(p.M, 2 * kBM) * ceil_div(p.N, bn) >= sm_count() / 2;
  // (A first answer to the partial last wave of the 392-tile pair launches -- its images handed
  // to a second launch of 128 x 128 single-CTA tiles -- was measured slower, 616.6 / 619.7
  // against 626.2 / 628.9 TFLOP/s, and removed: such a tile is operand-bound and lasts as long
  // as a 256 x 256 pair tile.  The K-split tail of launch_b is what replaced it.)
  CUtensorMap tmap;
  int rc = make_tmap_2d(&tmap, w, (uint64_t)p.N, (uint64_t)p.K, (uint32_t)(pair ? bn / 2 : bn));
  if (rc != CMR_OK) return rc;
  cudaStream_t st = as_stream(stream);
  switch (bn) {
    case 64: return launch<64, 6, false>(tmap, tmap_a, p, st);
    case 128: return launch<128, 5, false>(tmap, tmap_a, p, st);
    case 256:
      if (pair && !deep) return launch<256, 5, true>(tmap, tmap_a, p, st);
      if (pair) return launch<256, 6, true>(tmap, tmap_a, p, st);
      return deep ? launch<256, 4, false>(tmap, tmap_a, p, st)
                  : launch<256, 3, false>(tmap, tmap_a, p, st);
    default: return CMR_ERR_INVALID_ARG;
  }
}
