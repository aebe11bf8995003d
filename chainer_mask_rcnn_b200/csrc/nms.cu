// Greedy NMS (warp-bitmask) and RPN proposal generation for sm_100a.
//
// Replaces, on device and without host round trips:
//   chainercv non_maximum_suppression  (called at models/mask_rcnn.py:193-194 and
//                                       inside ProposalCreator)
//   chainercv ProposalCreator.__call__ (models/region_proposal_network.py:135-141)
//
// Results are integers (keep lists / proposal indices) and must be bit-exact
// against the CPU algorithm, so every fp32 expression below is written with
// explicitly rounded intrinsics in the same order as the NumPy code (no FMA
// contraction):  area = (y2-y1)*(x2-x1);  inter = (br_y-tl_y)*(br_x-tl_x) when
// tl < br on both axes, else 0;  iou = inter / ((area_i + area_j) - inter);
// suppressed iff iou >= thresh.
#include "common.cuh"

namespace cmr {
namespace {

__device__ __forceinline__ float box_area(const float4 b) {  // (y1,x1,y2,x2) = (x,y,z,w)
  return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

__device__ __forceinline__ bool iou_ge(const float4 a, float area_a, const float4 b,
                                       float area_b, float thresh) {
  const float tl_y = fmaxf(a.x, b.x), tl_x = fmaxf(a.y, b.y);
  const float br_y = fminf(a.z, b.z), br_x = fminf(a.w, b.w);
  float inter = __fmul_rn(__fsub_rn(br_y, tl_y), __fsub_rn(br_x, tl_x));
  if (!(tl_y < br_y && tl_x < br_x)) inter = __fmul_rn(inter, 0.0f);
  const float denom = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  // The decision is that of the correctly rounded quotient.  An approximate reciprocal
  // (one MUFU, relative error < 2^-21) settles every pair whose quotient is not within
  // 2^-20 of the threshold; only those few take the IEEE division.  (0/0 and other
  // non-finite cases fail both quick tests and are decided by the division as before.)
  float rcp;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(denom));
  const float q = inter * rcp;
  if (q > thresh * 1.000001f) return true;
  if (q < thresh * 0.999999f) return false;
  return __fdiv_rn(inter, denom) >= thresh;
}

// mask[i][cb] bit k  <=>  j = 64*cb + k > i  and  IoU(i, j) >= thresh  (and, when class
// labels are given, label[i] == label[j]: per-class suppression in one pass).
// grid = (nb, nb, B); only the upper block triangle does work.
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float4* __restrict__ boxes, const int* __restrict__ labels,
                const int* __restrict__ n_arr, int n_max, float thresh,
                unsigned long long* __restrict__ mask, int nb_stride) {
  const int img = blockIdx.z;
  const int n = n_arr ? min(n_arr[img], n_max) : n_max;
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (cb < rb || rb * 64 >= n || cb * 64 >= n) return;
  boxes += (size_t)img * n_max;
  if (labels) labels += (size_t)img * n_max;
  mask += (size_t)img * n_max * nb_stride;
  __shared__ float4 cbox[64];
  __shared__ float carea[64];
  __shared__ int clabel[64];
  const int t = threadIdx.x;
  const int j = cb * 64 + t;
  if (j < n) {
    cbox[t] = boxes[j];
    carea[t] = box_area(cbox[t]);
    clabel[t] = labels ? labels[j] : 0;
  }
  __syncthreads();
  const int i = rb * 64 + t;
  if (i < n) {
    const float4 b = boxes[i];
    const float ai = box_area(b);
    const int li = labels ? labels[i] : 0;
    const int ncol = min(64, n - cb * 64);
    unsigned long long bits = 0ull;
    for (int k = (rb == cb) ? t + 1 : 0; k < ncol; ++k)
      if (clabel[k] == li && iou_ge(b, ai, cbox[k], carea[k], thresh)) bits |= (1ull << k);
    mask[(size_t)i * nb_stride + cb] = bits;
  }
}

// One CTA per image resolves the bitmask sequentially in 64-box chunks.  Per chunk, lane 0 of
// warp 0 resolves the diagonal 64x64 block (the only serial part: a chain over the removed
// mask and nothing else), the warp derives the keep list from the kept mask and ORs word cb+1
// of the kept rows (prefetched for all 64 rows) into the running "removed" mask -- all the
// next chunk needs.  The remaining words (>= cb+2) of the kept rows are ORed in by the other
// 15 warps one iteration later, concurrently with the next chunk's resolution; that part is a
// chain of L2 round trips, so each warp keeps four rows x 256 columns in flight.  Measured with
// cycle counters on 12000 clustered boxes (round 2): 1308 k -> 681 k cycles without a limit,
// 607 k -> 160 k with the train step's limit of 2000 kept boxes (the row fetches were 88 % of
// the old kernel: 7 warps, one row and 128 columns at a time).
// OR into a 64-bit word of shared memory as two native 32-bit atomics (a 64-bit atomicOr on
// shared memory compiles to a compare-and-swap spin loop; bits are only ever set, so the
// halves need not be updated together)
__device__ __forceinline__ void or_shared(unsigned long long* w, unsigned long long v) {
  unsigned int* h = reinterpret_cast<unsigned int*>(w);
  const unsigned int lo = (unsigned int)v, hi = (unsigned int)(v >> 32);
  if (lo) atomicOr(h, lo);
  if (hi) atomicOr(h + 1, hi);
}

constexpr int kSweepThreads = 512;
constexpr int kSweepRestWarps = kSweepThreads / 32 - 1;
constexpr int kSweepRows = 4;      // rows whose words one warp fetches together

__global__ void __launch_bounds__(kSweepThreads)
nms_sweep_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ n_arr,
                 int n_max, int nb_stride, int limit, int32_t* __restrict__ keep,
                 int32_t* __restrict__ n_keep, int keep_stride) {
  extern __shared__ unsigned long long remv[];  // nb_stride words
  __shared__ unsigned long long diag[64], nextw[64];
  __shared__ int kept_list[2][64];
  __shared__ int kept_n[2], count_s, done_s;
  const int img = blockIdx.x;
  const int n = n_arr ? min(n_arr[img], n_max) : n_max;
  const int nb = (n + 63) / 64;
  mask += (size_t)img * n_max * nb_stride;
  keep += (size_t)img * keep_stride;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int w = t; w < nb_stride; w += kSweepThreads) remv[w] = 0ull;
  if (t == 0) {
    count_s = 0;
    done_s = 0;
    kept_n[0] = kept_n[1] = 0;
  }
  // thread t < 64 fetches, for row cb*64 + t, the diagonal word (cb) and the next one
  unsigned long long next_diag = 0ull, next_next = 0ull;
  if (t < 64 && t < n) {
    next_diag = mask[(size_t)t * nb_stride + 0];
    if (nb > 1) next_next = mask[(size_t)t * nb_stride + 1];
  }
  __syncthreads();
  for (int cb = 0; cb < nb; ++cb) {
    const int cur = cb & 1, prev = cur ^ 1;
    if (t < 64) {
      diag[t] = next_diag;
      nextw[t] = next_next;
      const int i = (cb + 1) * 64 + t;  // prefetch the next chunk's two words
      const bool ok = cb + 1 < nb && i < n;
      next_diag = ok ? mask[(size_t)i * nb_stride + cb + 1] : 0ull;
      next_next = (ok && cb + 2 < nb) ? mask[(size_t)i * nb_stride + cb + 2] : 0ull;
    }
    __syncthreads();
    if (warp == 0) {
      // Lane 0 resolves the chunk: the only serial part of the algorithm, so its dependent
      // chain carries nothing but the removed mask r and the kept mask K (test bit j; if clear,
      // r |= row j's diagonal word).  Everything else -- the keep list, the list for the other
      // warps, the OR of the kept rows' next words -- is derived from K by the whole warp.
      const int lim = min(64, n - cb * 64);
      const unsigned long long valid = lim == 64 ? ~0ull : ((1ull << lim) - 1ull);
      unsigned long long K = 0ull;
      if (lane == 0) {
        unsigned long long r = remv[cb] | ~valid;
        unsigned long long alive = ~r;
        if (__popcll(alive) > 20) {
          // many survivors: the whole diagonal block in registers, fixed 64-step walk (the
          // loads are issued up front instead of one dependent shared-memory read per box)
          // (volatile asm: the compiler otherwise sinks each load into its conditional use,
          // i.e. back into the dependent chain)
          // The chain is written on 32-bit halves: test one bit of the half that holds box j,
          // OR the row's halves in if it is clear -- two dependent instructions per box.
          unsigned int dl[64], dh[64];
          const unsigned int diag_addr = (unsigned int)__cvta_generic_to_shared(diag);
#pragma unroll
          for (int j = 0; j < 64; ++j)
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                         : "=r"(dl[j]), "=r"(dh[j])
                         : "r"(diag_addr + 8 * j));
          unsigned int rl = (unsigned int)r, rh = (unsigned int)(r >> 32), kl = 0u, kh = 0u;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (!(rl & (1u << j))) {
              rl |= dl[j];
              rh |= dh[j];
              kl |= 1u << j;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (!(rh & (1u << j))) {
              rh |= dh[32 + j];
              kh |= 1u << j;
            }
          }
          K = ((unsigned long long)kh << 32) | kl;
        } else {
          // few survivors: find-first-set over the complement of the removed mask (a chunk
          // that earlier boxes suppressed entirely costs one test)
          // candidates = the boxes alive at the start of the chunk, four at a time: their
          // diagonal words are fetched together (independent of the chain), then each is
          // kept if no earlier kept box of the chunk removed it
          while (alive) {
            int j[4];
            unsigned long long d[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              j[e] = alive ? __ffsll((long long)alive) - 1 : -1;
              alive &= alive - 1ull;              // (0 stays 0)
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) d[e] = diag[j[e] < 0 ? 0 : j[e]];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (j[e] >= 0 && !((r >> j[e]) & 1ull)) {
                r |= d[e];
                K |= 1ull << j[e];
              }
            }
          }
        }
      }
      K = __shfl_sync(0xffffffffu, K, 0);
      const int cnt = count_s;
      int nk = __popcll(K);
      bool done = false;
      if (limit > 0 && cnt + nk >= limit) {     // the first (limit - cnt) kept boxes only
        done = true;
        while (cnt + nk > limit) {
          K &= ~(1ull << (63 - __clzll((long long)K)));
          --nk;
        }
      }
      unsigned long long rn = 0ull;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;
        if ((K >> j) & 1ull) {
          const int pos = __popcll(K & ((1ull << j) - 1ull));
          keep[cnt + pos] = cb * 64 + j;
          kept_list[cur][pos] = j;
          rn |= nextw[j];
        }
      }
      const unsigned int rn_lo = __reduce_or_sync(0xffffffffu, (unsigned int)rn);
      const unsigned int rn_hi = __reduce_or_sync(0xffffffffu, (unsigned int)(rn >> 32));
      rn = ((unsigned long long)rn_hi << 32) | rn_lo;
      __syncwarp();                 // every lane has read count_s before lane 0 replaces it
      if (lane == 0) {
        if (rn && cb + 1 < nb) or_shared(&remv[cb + 1], rn);
        kept_n[cur] = nk;
        count_s = cnt + nk;
        if (done) done_s = 1;
      }
    } else if (warp >= 1 && cb > 0) {
      // words >= cb+1 of the rows kept in chunk cb-1 (word cb went in last iteration)
      // (this is the chunk's critical path -- L2 round trips -- so a warp keeps the words of
      // kSweepRows rows x 8 x 32 columns in flight at once and ORs them before the atomics)
      const int nk = kept_n[prev];
      for (int q = warp - 1; q < nk; q += kSweepRestWarps * kSweepRows) {
        const unsigned long long* row[kSweepRows];
#pragma unroll
        for (int e = 0; e < kSweepRows; ++e) {
          const int qe = q + e * kSweepRestWarps;
          row[e] = qe < nk ? mask + (size_t)((cb - 1) * 64 + kept_list[prev][qe]) * nb_stride
                           : nullptr;
        }
        for (int w0 = cb + 1; w0 < nb; w0 += 256) {
          unsigned long long v[kSweepRows][8];
#pragma unroll
          for (int e = 0; e < kSweepRows; ++e) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int w = w0 + 32 * u + lane;
              v[e][u] = (row[e] && w < nb) ? row[e][w] : 0ull;
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            unsigned long long acc = v[0][u];
#pragma unroll
            for (int e = 1; e < kSweepRows; ++e) acc |= v[e][u];
            if (acc) or_shared(&remv[w0 + 32 * u + lane], acc);
          }
        }
      }
    }
    __syncthreads();
    if (done_s) break;
  }
  if (t == 0) n_keep[img] = count_s;
}

// --------------------------------------------------------------- proposals --
__device__ __forceinline__ unsigned int float_to_sortable(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// loc2bbox + clip + min-size filter.  One thread per anchor.
// key = (sortable(score) << 32 | anchor index) for surviving boxes, 0 otherwise.
__global__ void __launch_bounds__(256)
proposal_decode_kernel(const float4* __restrict__ loc, const float* __restrict__ score,
                       const float4* __restrict__ anchor, int n_anchor, int n_pad,
                       float img_h, float img_w, float min_size, float4* __restrict__ roi,
                       unsigned long long* __restrict__ keys, int* __restrict__ n_valid) {
  const int img = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_pad) return;
  unsigned long long key = 0ull;
  if (k < n_anchor) {
    const float4 a = __ldg(anchor + k);
    const float4 l = __ldg(loc + (size_t)img * n_anchor + k);
    const float h = __fsub_rn(a.z, a.x), w = __fsub_rn(a.w, a.y);
    const float cy = __fadd_rn(a.x, __fmul_rn(0.5f, h));
    const float cx = __fadd_rn(a.y, __fmul_rn(0.5f, w));
    const float ncy = __fadd_rn(__fmul_rn(l.x, h), cy);
    const float ncx = __fadd_rn(__fmul_rn(l.y, w), cx);
    // exp evaluated in fp64 and rounded once: the correctly rounded fp32 value.
    const float nh = __fmul_rn((float)exp((double)l.z), h);
    const float nw = __fmul_rn((float)exp((double)l.w), w);
    float y1 = __fsub_rn(ncy, __fmul_rn(0.5f, nh)), x1 = __fsub_rn(ncx, __fmul_rn(0.5f, nw));
    float y2 = __fadd_rn(ncy, __fmul_rn(0.5f, nh)), x2 = __fadd_rn(ncx, __fmul_rn(0.5f, nw));
    y1 = fminf(fmaxf(y1, 0.f), img_h);
    y2 = fminf(fmaxf(y2, 0.f), img_h);
    x1 = fminf(fmaxf(x1, 0.f), img_w);
    x2 = fminf(fmaxf(x2, 0.f), img_w);
    roi[(size_t)img * n_anchor + k] = make_float4(y1, x1, y2, x2);
    const bool ok = (__fsub_rn(y2, y1) >= min_size) && (__fsub_rn(x2, x1) >= min_size);
    if (ok) {
      key = ((unsigned long long)float_to_sortable(__ldg(score + (size_t)img * n_anchor + k))
             << 32) | (unsigned int)k;
      atomicAdd(n_valid + img, 1);
    }
  }
  keys[(size_t)img * n_pad + k] = key;
}

// Descending bitonic sort of `rows` independent arrays of n_pad (power of two) 64-bit
// keys, spread over many CTAs: chunks of 512 * KPT keys are sorted / merged by one CTA
// each, and the stages whose partner distance reaches across chunks run as grid-wide
// passes over the (L2-resident) array, two stages per pass.
//
// Inside a chunk thread t keeps the KPT consecutive keys KPT*t .. KPT*t + KPT-1 in
// registers: stages with partner distance j < KPT are compare-exchanges between a
// thread's own registers, KPT <= j <= 16*KPT exchange keys with lane (t ^ j/KPT) by warp
// shuffles, and only j >= 32*KPT goes through shared memory with a CTA barrier per stage
// (10 of the 78 stages of a 4096-key sort).
constexpr int kSortThreads = 512;

__device__ __forceinline__ void cmp_swap_desc(unsigned long long& a, unsigned long long& b,
                                              bool desc) {
  if ((a < b) == desc) {
    unsigned long long t = a;
    a = b;
    b = t;
  }
}

// Stages j = j_hi .. 1 of merge level k on the chunk held in r (position KPT*t + e of the
// chunk, global position base + that).
template <int KPT>
__device__ __forceinline__ void bitonic_level(unsigned long long (&r)[KPT],
                                              unsigned long long* sk, int base, int k,
                                              int j_hi) {
  constexpr int kChunk = kSortThreads * KPT;
  const int t = threadIdx.x;
  int j = j_hi;
  if (j >= 32 * KPT) {
#pragma unroll
    for (int e = 0; e < KPT; ++e) sk[KPT * t + e] = r[e];
    __syncthreads();
    for (; j >= 32 * KPT; j >>= 1) {
      for (int i = t; i < kChunk / 2; i += kSortThreads) {
        const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
        cmp_swap_desc(sk[lo], sk[lo | j], ((base + lo) & k) == 0);
      }
      __syncthreads();
    }
#pragma unroll
    for (int e = 0; e < KPT; ++e) r[e] = sk[KPT * t + e];
  }
  for (; j >= KPT; j >>= 1) {
    // k > j >= KPT: all keys of the thread share the direction and the side of the pair
    const bool take_max = (((KPT * t) & j) == 0) == (((base + KPT * t) & k) == 0);
#pragma unroll
    for (int e = 0; e < KPT; ++e) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, r[e], j / KPT);
      r[e] = ((other > r[e]) == take_max) ? other : r[e];
    }
  }
#pragma unroll
  for (int jj = KPT / 2; jj >= 1; jj >>= 1) {
    if (jj > j) continue;              // stages above j_hi belong to earlier calls
#pragma unroll
    for (int e = 0; e < KPT; ++e)
      if ((e & jj) == 0)
        cmp_swap_desc(r[e], r[e | jj], ((base + KPT * t + e) & k) == 0);
  }
}

// k_first == 2: full sort of every chunk (levels 2 .. chunk); otherwise the tail
// (j < chunk) of merge level k_first.  grid = (max(1, n_pad / chunk), rows); a row shorter
// than one chunk is padded with zero keys (they sort to the end).
template <int KPT>
__global__ void __launch_bounds__(kSortThreads)
bitonic_chunk_kernel(unsigned long long* __restrict__ keys_all, int n_pad, int k_first) {
  constexpr int kChunk = kSortThreads * KPT;
  __shared__ unsigned long long sk[kChunk];
  const int base = blockIdx.x * kChunk;
  const int t = threadIdx.x;
  unsigned long long* keys = keys_all + (size_t)blockIdx.y * n_pad + base;
  const int n = min(kChunk, n_pad - base);
  unsigned long long r[KPT];
#pragma unroll
  for (int e = 0; e < KPT; ++e) r[e] = KPT * t + e < n ? keys[KPT * t + e] : 0ull;
  if (k_first == 2) {
#pragma unroll 1
    for (int k = 2; k <= kChunk; k <<= 1) bitonic_level<KPT>(r, sk, base, k, k >> 1);
  } else {
    bitonic_level<KPT>(r, sk, base, k_first, kChunk >> 1);
  }
#pragma unroll
  for (int e = 0; e < KPT; ++e)
    if (KPT * t + e < n) keys[KPT * t + e] = r[e];
}

// Grid-wide pass: stages j and j/2 of merge level k (j/2 skipped when two == 0).  Each
// thread owns the four keys {q, q+j/2, q+j, q+3j/2} (or the pair {q, q+j}).
__global__ void __launch_bounds__(256)
bitonic_global_kernel(unsigned long long* __restrict__ keys_all, int n_pad, int k, int j,
                      int two) {
  unsigned long long* keys = keys_all + (size_t)blockIdx.y * n_pad;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (two) {
    if (i >= n_pad / 4) return;
    const int h = j >> 1;
    const int q = ((i & ~(h - 1)) << 2) | (i & (h - 1));
    const bool desc = (q & k) == 0;
    unsigned long long a = keys[q], b = keys[q + h], c = keys[q + j], d = keys[q + j + h];
    cmp_swap_desc(a, c, desc);
    cmp_swap_desc(b, d, desc);
    cmp_swap_desc(a, b, desc);
    cmp_swap_desc(c, d, desc);
    keys[q] = a;
    keys[q + h] = b;
    keys[q + j] = c;
    keys[q + j + h] = d;
  } else {
    if (i >= n_pad / 2) return;
    const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
    unsigned long long a = keys[lo], b = keys[lo | j];
    if ((a < b) == ((lo & k) == 0)) {
      keys[lo] = b;
      keys[lo | j] = a;
    }
  }
}

__global__ void __launch_bounds__(256)
proposal_gather_sorted_kernel(const unsigned long long* __restrict__ keys,
                              const float4* __restrict__ roi, const int* __restrict__ n_valid,
                              int n_anchor, int n_pad, int n_pre, float4* __restrict__ sorted,
                              int32_t* __restrict__ sorted_idx, int* __restrict__ n_cand) {
  const int img = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(n_valid[img], n_pre);
  if (k == 0) n_cand[img] = n;
  if (k >= n) return;
  const int idx = (int)(keys[(size_t)img * n_pad + k] & 0xffffffffull);
  sorted[(size_t)img * n_pre + k] = roi[(size_t)img * n_anchor + idx];
  sorted_idx[(size_t)img * n_pre + k] = idx;
}

__global__ void __launch_bounds__(256)
proposal_emit_kernel(const float4* __restrict__ sorted, const int32_t* __restrict__ sorted_idx,
                     const int32_t* __restrict__ keep, const int32_t* __restrict__ n_keep,
                     int n_pre, int n_post, float4* __restrict__ rois_out,
                     int32_t* __restrict__ idx_out, int32_t* __restrict__ n_out) {
  const int img = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(n_keep[img], n_post);
  if (q == 0) n_out[img] = n;
  if (q >= n_post) return;
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  int idx = -1;
  if (q < n) {
    const int s = keep[(size_t)img * n_pre + q];
    r = sorted[(size_t)img * n_pre + s];
    idx = sorted_idx[(size_t)img * n_pre + s];
  }
  rois_out[(size_t)img * n_post + q] = r;
  idx_out[(size_t)img * n_post + q] = idx;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

int launch_nms(const float4* boxes, const int* n_arr, int n_max, int B, float thresh, int limit,
               int32_t* keep, int32_t* n_keep, unsigned long long* mask, cudaStream_t st,
               const int* labels = nullptr) {
  const int nb = (n_max + 63) / 64;
  dim3 grid(nb, nb, B);
  nms_mask_kernel<<<grid, 64, 0, st>>>(boxes, labels, n_arr, n_max, thresh, mask, nb);
  CMR_LAUNCH_CHECK();
  const size_t smem = sizeof(unsigned long long) * nb;
  nms_sweep_kernel<<<B, kSweepThreads, smem, st>>>(mask, n_arr, n_max, nb, limit, keep, n_keep,
                                                   n_max);
  CMR_LAUNCH_CHECK();
  return CMR_OK;
}

}  // namespace

// Shared with detect.cu: greedy NMS of B score-sorted box lists (n_arr[b] <= n_max boxes
// each); with `labels` a box only suppresses boxes of its own class.
int launch_nms_batch(const float* boxes, const int* labels, const int* n_arr, int n_max, int B,
                     float thresh, int32_t* keep, int32_t* n_keep, unsigned long long* mask,
                     cudaStream_t st) {
  return launch_nms(reinterpret_cast<const float4*>(boxes), n_arr, n_max, B, thresh, 0, keep,
                    n_keep, mask, st, labels);
}

// Shared with targets.cu: sorts `rows` independent key arrays of n_pad (power of two).
template <int KPT>
int launch_sort_kpt(unsigned long long* keys, int n_pad, int rows, cudaStream_t st) {
  constexpr int chunk = kSortThreads * KPT;
  const dim3 cgrid(n_pad > chunk ? n_pad / chunk : 1, rows);
  bitonic_chunk_kernel<KPT><<<cgrid, kSortThreads, 0, st>>>(keys, n_pad, 2);
  CMR_LAUNCH_CHECK();
  for (int k = chunk << 1; k <= n_pad; k <<= 1) {
    int j = k >> 1;
    while (j >= chunk) {
      const int two = (j >> 1) >= chunk ? 1 : 0;
      const int work = two ? n_pad / 4 : n_pad / 2;
      bitonic_global_kernel<<<dim3(ceil_div(work, 256), rows), 256, 0, st>>>(keys, n_pad, k, j,
                                                                            two);
      CMR_LAUNCH_CHECK();
      j >>= two ? 2 : 1;
    }
    bitonic_chunk_kernel<KPT><<<cgrid, kSortThreads, 0, st>>>(keys, n_pad, k);
    CMR_LAUNCH_CHECK();
  }
  return CMR_OK;
}

int launch_sort_desc_u64(unsigned long long* keys, int n_pad, int rows, cudaStream_t st) {
  if (n_pad <= 0 || (n_pad & (n_pad - 1)) != 0 || rows <= 0 || rows > 65535)
    return CMR_ERR_INVALID_ARG;
  if (n_pad >= kSortThreads * 8) return launch_sort_kpt<8>(keys, n_pad, rows, st);
  if (n_pad >= kSortThreads * 4) return launch_sort_kpt<4>(keys, n_pad, rows, st);
  return launch_sort_kpt<2>(keys, n_pad, rows, st);
}
}  // namespace cmr

using namespace cmr;

extern "C" size_t cmr_nms_workspace_bytes(int n) {
  if (n <= 0) return 256;
  const size_t nb = (n + 63) / 64;
  return align_up(sizeof(unsigned long long) * (size_t)n * nb, 256);
}

extern "C" int cmr_nms(const float* boxes, int n, float thresh, int limit, int32_t* keep,
                       int32_t* n_keep, void* workspace, size_t workspace_bytes, void* stream) {
  CMR_REQUIRE(n >= 0 && n_keep);
  cudaStream_t st = as_stream(stream);
  if (n == 0) {
    CMR_CUDA_TRY(cudaMemsetAsync(n_keep, 0, sizeof(int32_t), st));
    return CMR_OK;
  }
  CMR_REQUIRE(boxes && keep && workspace);
  CMR_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0);
  CMR_REQUIRE(n <= 64 * 2048);  // sweep keeps ceil(n/64) words in shared memory
  if (workspace_bytes < cmr_nms_workspace_bytes(n)) return CMR_ERR_WORKSPACE;
  return launch_nms(reinterpret_cast<const float4*>(boxes), nullptr, n, 1, thresh, limit, keep,
                    n_keep, reinterpret_cast<unsigned long long*>(workspace), st);
}

namespace {
struct ProposalWs {
  size_t roi, keys, n_valid, sorted, sorted_idx, n_cand, keep, n_keep, mask, total;
  int n_pad;
};
ProposalWs proposal_ws(int B, int n_anchor, int n_pre) {
  ProposalWs w;
  w.n_pad = next_pow2(n_anchor);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const size_t nb = (n_pre + 63) / 64;
  w.roi = take(sizeof(float4) * (size_t)B * n_anchor);
  w.keys = take(sizeof(unsigned long long) * (size_t)B * w.n_pad);
  w.n_valid = take(sizeof(int) * B);
  w.sorted = take(sizeof(float4) * (size_t)B * n_pre);
  w.sorted_idx = take(sizeof(int32_t) * (size_t)B * n_pre);
  w.n_cand = take(sizeof(int) * B);
  w.keep = take(sizeof(int32_t) * (size_t)B * n_pre);
  w.n_keep = take(sizeof(int32_t) * B);
  w.mask = take(sizeof(unsigned long long) * (size_t)B * n_pre * nb);
  w.total = off;
  return w;
}
}  // namespace

extern "C" size_t cmr_proposals_workspace_bytes(int B, int n_anchor, int n_pre) {
  if (B <= 0 || n_anchor <= 0 || n_pre <= 0) return 256;
  return proposal_ws(B, n_anchor, n_pre).total;
}

extern "C" int cmr_proposals(const float* loc, const float* score, const float* anchor, int B,
                             int n_anchor, float img_h, float img_w, float min_size, int n_pre,
                             int n_post, float nms_thresh, float* rois_out, int32_t* idx_out,
                             int32_t* n_out, void* workspace, size_t workspace_bytes,
                             void* stream) {
  CMR_REQUIRE(B > 0 && n_anchor > 0 && n_post > 0);
  CMR_REQUIRE(loc && score && anchor && rois_out && idx_out && n_out && workspace);
  if (n_pre <= 0 || n_pre > n_anchor) n_pre = n_anchor;  // "no limit"
  CMR_REQUIRE(n_pre <= 64 * 2048 && n_anchor <= (1 << 22));
  const ProposalWs w = proposal_ws(B, n_anchor, n_pre);
  if (workspace_bytes < w.total) return CMR_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  char* ws = reinterpret_cast<char*>(workspace);
  float4* roi = reinterpret_cast<float4*>(ws + w.roi);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws + w.keys);
  int* n_valid = reinterpret_cast<int*>(ws + w.n_valid);
  float4* sorted = reinterpret_cast<float4*>(ws + w.sorted);
  int32_t* sorted_idx = reinterpret_cast<int32_t*>(ws + w.sorted_idx);
  int* n_cand = reinterpret_cast<int*>(ws + w.n_cand);
  int32_t* keep = reinterpret_cast<int32_t*>(ws + w.keep);
  int32_t* n_keep = reinterpret_cast<int32_t*>(ws + w.n_keep);
  unsigned long long* mask = reinterpret_cast<unsigned long long*>(ws + w.mask);

  CMR_CUDA_TRY(cudaMemsetAsync(n_valid, 0, sizeof(int) * B, st));
  {
    dim3 grid(ceil_div(w.n_pad, 256), B);
    proposal_decode_kernel<<<grid, 256, 0, st>>>(
        reinterpret_cast<const float4*>(loc), score, reinterpret_cast<const float4*>(anchor),
        n_anchor, w.n_pad, img_h, img_w, min_size, roi, keys, n_valid);
    CMR_LAUNCH_CHECK();
  }
  {
    int rc = launch_sort_desc_u64(keys, w.n_pad, B, st);
    if (rc != CMR_OK) return rc;
  }
  {
    dim3 grid(ceil_div(n_pre, 256), B);
    proposal_gather_sorted_kernel<<<grid, 256, 0, st>>>(keys, roi, n_valid, n_anchor, w.n_pad,
                                                        n_pre, sorted, sorted_idx, n_cand);
    CMR_LAUNCH_CHECK();
  }
  int rc = launch_nms(sorted, n_cand, n_pre, B, nms_thresh, n_post, keep, n_keep, mask, st);
  if (rc != CMR_OK) return rc;
  {
    dim3 grid(ceil_div(n_post, 256), B);
    proposal_emit_kernel<<<grid, 256, 0, st>>>(sorted, sorted_idx, keep, n_keep, n_pre, n_post,
                                               reinterpret_cast<float4*>(rois_out), idx_out,
                                               n_out);
    CMR_LAUNCH_CHECK();
  }
  return CMR_OK;
}
