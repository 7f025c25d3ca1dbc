// RPN proposal generation on device: anchor grid, sigmoid/top-k/decode/clip, and NMS with an
// on-device scan (no D2H mask copy, no host loop — the reference's nms.cu:99-123 does both).
//
// Reference semantics: rpn/anchor_generator.py:73-111, rpn/inference.py:87-115, box_coder.py:52-95,
// structures/bounding_box.py:214-225, structures/boxlist_ops.py:37-51, csrc/cuda/nms.cu:13-131.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kSortCap = 16384;     // max elements of the in-shared-memory bitonic sort (128 KB of u64)
constexpr int kTopkThreads = 1024;

// ----------------------------------------------------------------------------- anchors
__global__ void anchor_grid_kernel(const float* __restrict__ cell, int A, int FH, int FW, int stride, int img_w,
                                   int img_h, int straddle, float* __restrict__ anchors,
                                   uint8_t* __restrict__ vis) {
  const int total = FH * FW * A;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int a = i % A, x = (i / A) % FW, y = i / (A * FW);
    const float sx = (float)(x * stride), sy = (float)(y * stride);
    float4 b;
    b.x = sx + cell[a * 4 + 0];
    b.y = sy + cell[a * 4 + 1];
    b.z = sx + cell[a * 4 + 2];
    b.w = sy + cell[a * 4 + 3];
    reinterpret_cast<float4*>(anchors)[i] = b;
    bool inside = true;
    if (straddle >= 0)
      inside = b.x >= (float)(-straddle) && b.y >= (float)(-straddle) && b.z < (float)(img_w + straddle) &&
               b.w < (float)(img_h + straddle);
    vis[i] = inside ? 1 : 0;
  }
}

// ----------------------------------------------------------------------------- sort helpers
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// In-place bitonic sort, DESCENDING, of n_pow2 u64 keys in shared memory by the whole CTA.
__device__ void bitonic_sort_desc(unsigned long long* s, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n_pow2 >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // index with bit j cleared
        const int p = i | j;
        const bool up = (i & k) == 0;                          // "up" block => descending here
        const unsigned long long a = s[i], b = s[p];
        if ((a < b) == up) { s[i] = b; s[p] = a; }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// Exclusive block scan of 0/1 flags in index order; returns this thread's rank and the block total.
__device__ __forceinline__ int block_rank(bool flag, int* warp_tot, int& total) {
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int within = __popc(bal & ((1u << lane) - 1u));
  __syncthreads();
  if (lane == 0) warp_tot[wid] = __popc(bal);
  __syncthreads();
  int base = 0, tot = 0;
  const int nw = blockDim.x >> 5;
  for (int w = 0; w < nw; ++w) {
    const int c = warp_tot[w];
    if (w < wid) base += c;
    tot += c;
  }
  total = tot;
  return base + within;
}

// ----------------------------------------------------------------------------- top-k + decode
// One CTA per image.  dynamic smem: n_pow2 * 8 bytes.
__global__ void __launch_bounds__(kTopkThreads) rpn_topk_decode_kernel(
    const float* __restrict__ logits, const float* __restrict__ deltas, const float* __restrict__ anchors,
    int num_anchors, int k, int img_w, int img_h, float min_size, float* __restrict__ boxes,
    float* __restrict__ scores, int32_t* __restrict__ topk_idx, int32_t* __restrict__ valid) {
  extern __shared__ unsigned long long skeys[];
  __shared__ int hist[256];
  __shared__ int warp_tot[32];
  __shared__ uint32_t s_prefix;
  __shared__ int s_need, s_count;

  const int img = blockIdx.x;
  const float* lg = logits + (size_t)img * num_anchors;
  const int n_pow2 = next_pow2(k);

  // ---- radix select: exact k-th largest ordered key -------------------------------------------
  uint32_t prefix = 0, prefix_mask = 0;
  int need = k;   // how many more elements we must take from the "== prefix so far" population
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    // objectness logits share their leading bytes, so nearly every lane hits the same bin in the first passes:
    // aggregate equal bins inside the warp and issue ONE shared-memory atomic per distinct bin.  Four logits per
    // thread and iteration (one 128-bit load) keep enough loads in flight for a single CTA.
    const bool vec = (num_anchors & 3) == 0 && (reinterpret_cast<uintptr_t>(lg) & 15) == 0;
    constexpr int kUn = 4;                          // 128-bit loads in flight per thread
    const int step = blockDim.x * 4;
    for (int i0 = threadIdx.x * 4; i0 < num_anchors; i0 += step * kUn) {
      float v[kUn][4];
#pragma unroll
      for (int u = 0; u < kUn; ++u) {
        const int i = i0 + u * step;
        if (vec && i < num_anchors) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(lg + i));
          v[u][0] = f.x; v[u][1] = f.y; v[u][2] = f.z; v[u][3] = f.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[u][e] = i + e < num_anchors ? lg[i + e] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < kUn; ++u) {
        const int i = i0 + u * step;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t key = float_to_ordered(v[u][e]);
          const bool in = i + e < num_anchors && (key & prefix_mask) == prefix;
          const uint32_t bin = in ? ((key >> shift) & 0xFF) : 0xFFFFFFFFu;
          const unsigned peers = __match_any_sync(__activemask(), bin);
          if (in && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = 0, b = 255;
      for (; b > 0; --b) {
        if (acc + hist[b] >= need) break;
        acc += hist[b];
      }
      s_prefix = prefix | ((uint32_t)b << shift);
      s_need = need - acc;
    }
    __syncthreads();
    prefix = s_prefix;
    need = s_need;
    prefix_mask |= 0xFFu << shift;
    __syncthreads();
  }
  const uint32_t kth = prefix;   // exact key of the k-th largest; `need` of the ties at kth are taken

  // ---- compaction: everything > kth, plus the `need` lowest-index elements == kth ---------------
  if (threadIdx.x == 0) s_count = 0;
  for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) skeys[i] = 0ull;
  __syncthreads();
  int ties_taken = 0;
  for (int base = 0; base < num_anchors; base += blockDim.x) {
    const int i = base + threadIdx.x;
    uint32_t key = 0;
    bool gt = false, eq = false;
    if (i < num_anchors) {
      key = float_to_ordered(lg[i]);
      gt = key > kth;
      eq = key == kth;
    }
    int tot;
    const int r = block_rank(eq, warp_tot, tot);
    const bool take_eq = eq && (ties_taken + r) < need;
    ties_taken += tot;
    if (gt || take_eq) {
      const int slot = atomicAdd(&s_count, 1);
      skeys[slot] = ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
    }
  }
  __syncthreads();
  bitonic_sort_desc(skeys, n_pow2);

  // ---- decode, clip, min-size filter, ordered compaction ------------------------------------------
  const float clip = 4.135166556742356f;   // log(1000/16)
  const float xmax = (float)(img_w - 1), ymax = (float)(img_h - 1);
  int out_base = 0;
  for (int base = 0; base < k; base += blockDim.x) {
    const int r = base + threadIdx.x;
    bool ok = false;
    float4 b = make_float4(0, 0, 0, 0);
    float sc = 0.f;
    int idx = 0;
    if (r < k) {
      idx = (int)(0xFFFFFFFFu - (uint32_t)(skeys[r] & 0xFFFFFFFFull));
      const float x = lg[idx];
      sc = 1.0f / (1.0f + expf(-x));
      const float4 d = dd::ldg4(deltas + ((size_t)img * num_anchors + idx) * 4);
      const float4 a = dd::ldg4(anchors + (size_t)idx * 4);
      const float w = __fadd_rn(__fsub_rn(a.z, a.x), 1.0f), h = __fadd_rn(__fsub_rn(a.w, a.y), 1.0f);
      const float cx = __fadd_rn(a.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(a.y, __fmul_rn(0.5f, h));
      const float dw = fminf(d.z, clip), dh = fminf(d.w, clip);
      const float pcx = __fadd_rn(__fmul_rn(d.x, w), cx), pcy = __fadd_rn(__fmul_rn(d.y, h), cy);
      const float pw = __fmul_rn(expf(dw), w), phh = __fmul_rn(expf(dh), h);
      b.x = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
      b.y = __fsub_rn(pcy, __fmul_rn(0.5f, phh));
      b.z = __fsub_rn(__fadd_rn(pcx, __fmul_rn(0.5f, pw)), 1.0f);
      b.w = __fsub_rn(__fadd_rn(pcy, __fmul_rn(0.5f, phh)), 1.0f);
      b.x = fminf(fmaxf(b.x, 0.f), xmax);
      b.y = fminf(fmaxf(b.y, 0.f), ymax);
      b.z = fminf(fmaxf(b.z, 0.f), xmax);
      b.w = fminf(fmaxf(b.w, 0.f), ymax);
      const float ws = __fadd_rn(__fsub_rn(b.z, b.x), 1.0f), hs = __fadd_rn(__fsub_rn(b.w, b.y), 1.0f);
      ok = ws >= min_size && hs >= min_size;
    }
    int tot;
    const int rank = block_rank(ok, warp_tot, tot);
    if (ok) {
      const size_t o = (size_t)img * k + out_base + rank;
      reinterpret_cast<float4*>(boxes)[o] = b;
      scores[o] = sc;
      topk_idx[o] = idx;
    }
    out_base += tot;
  }
  if (threadIdx.x == 0) valid[img] = out_base;
}

// ----------------------------------------------------------------------------- top-k + decode, cluster version
// The one-CTA kernel above walks the 122 880 logits of an image five times through L2 with the memory parallelism of
// one SM and ranks its compaction 1024 candidates at a time (0.5 ms, on the critical path of the step).  Here a
// cluster of 8 CTAs owns one image.  Every CTA loads an interleaved eighth of the logits ONCE into shared memory
// (4096-logit chunks dealt round-robin, so that a region of high scores is spread over all CTAs); the exact k-th
// largest COMPOSITE key (ordered logit, then lower index first — no ties, so the selection needs no index-ordered
// ranking) is found by a radix select whose per-pass histograms are summed over the cluster through distributed
// shared memory; every CTA sorts its own survivors (bitonic, typically k/8 of them) and finds the global rank of
// each by binary searches in the seven other sorted lists; decode, clip and the min-size filter follow, the ordered
// compaction of the filter going through a k-bit map gathered from all CTAs.  Output identical to the kernel above.
constexpr int kTkCl = 8;
constexpr int kTkIter = 4;                         // float4 loads per thread: 8 CTAs x 1024 threads x 16 = 131 072 logits
constexpr int kTkStride = 4 * kTkIter + 1;         // odd stride between the chunks of consecutive threads

struct TopkShared {
  int hist[7][256];          // one histogram per radix pass (never re-used: one cluster barrier per exchange)
  int ghist[256];
  int sm32[32];
  uint32_t digit;
  int need;
  int n_local;
  uint32_t okbits[512];      // bit r: the candidate of global rank r (held by this CTA) passes the min-size filter
  uint32_t allbits[512];     // the same for all CTAs
  int okprefix[512];
};

__device__ __forceinline__ bool decode_clip_box(const float* __restrict__ deltas, const float* __restrict__ anchors,
                                                size_t delta_row, int idx, float xmax, float ymax, float min_size,
                                                float4& b) {
  const float clip = 4.135166556742356f;   // log(1000/16)
  const float4 d = dd::ldg4(deltas + delta_row * 4);
  const float4 a = dd::ldg4(anchors + (size_t)idx * 4);
  const float w = __fadd_rn(__fsub_rn(a.z, a.x), 1.0f), h = __fadd_rn(__fsub_rn(a.w, a.y), 1.0f);
  const float cx = __fadd_rn(a.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(a.y, __fmul_rn(0.5f, h));
  const float dw = fminf(d.z, clip), dh = fminf(d.w, clip);
  const float pcx = __fadd_rn(__fmul_rn(d.x, w), cx), pcy = __fadd_rn(__fmul_rn(d.y, h), cy);
  const float pw = __fmul_rn(expf(dw), w), phh = __fmul_rn(expf(dh), h);
  b.x = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
  b.y = __fsub_rn(pcy, __fmul_rn(0.5f, phh));
  b.z = __fsub_rn(__fadd_rn(pcx, __fmul_rn(0.5f, pw)), 1.0f);
  b.w = __fsub_rn(__fadd_rn(pcy, __fmul_rn(0.5f, phh)), 1.0f);
  b.x = fminf(fmaxf(b.x, 0.f), xmax);
  b.y = fminf(fmaxf(b.y, 0.f), ymax);
  b.z = fminf(fmaxf(b.z, 0.f), xmax);
  b.w = fminf(fmaxf(b.w, 0.f), ymax);
  const float ws = __fadd_rn(__fsub_rn(b.z, b.x), 1.0f), hs = __fadd_rn(__fsub_rn(b.w, b.y), 1.0f);
  return ws >= min_size && hs >= min_size;
}

__global__ void __launch_bounds__(kTopkThreads) rpn_topk_decode_cluster_kernel(
    const float* __restrict__ logits, const float* __restrict__ deltas, const float* __restrict__ anchors,
    int num_anchors, int k, int img_w, int img_h, float min_size, float* __restrict__ boxes,
    float* __restrict__ scores, int32_t* __restrict__ topk_idx, int32_t* __restrict__ valid) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ TopkShared sh;
  uint32_t* skey = reinterpret_cast<uint32_t*>(dyn);                               // kTopkThreads * kTkStride
  int* srank = reinterpret_cast<int*>(dyn);                                        // later: rank of list entry p
  unsigned long long* slist = reinterpret_cast<unsigned long long*>(dyn + kTopkThreads * kTkStride * 4 + 8);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int img = blockIdx.x / kTkCl;
  const int tid = threadIdx.x, lane = tid & 31;
  const float* lg = logits + (size_t)img * num_anchors;
  const bool vec = (num_anchors & 3) == 0 && (reinterpret_cast<uintptr_t>(lg) & 15) == 0;

  for (int i = tid; i < 7 * 256; i += kTopkThreads) (&sh.hist[0][0])[i] = 0;
  for (int i = tid; i < 512; i += kTopkThreads) sh.okbits[i] = 0u;
  // ---- the one pass over the logits: this CTA's chunks -> shared memory (ordered keys)
#pragma unroll
  for (int j = 0; j < kTkIter; ++j) {
    const int i0 = 4 * (tid + kTopkThreads * (rank + kTkCl * j));
    float v[4];
    if (vec && i0 < num_anchors) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(lg + i0));
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = i0 + e < num_anchors ? lg[i0 + e] : 0.f;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) skey[tid * kTkStride + j * 4 + e] = float_to_ordered(v[e]);
  }
  __syncthreads();

  // ---- radix select of the k-th largest composite key: 4 passes over the ordered logit, 3 over (2^32 - 1 - index)
  uint32_t pre_hi = 0u, mask_hi = 0u, pre_lo = 0xFF000000u, mask_lo = 0xFF000000u;
  int need = k;
  for (int pass = 0; pass < 7; ++pass) {
    const bool hi = pass < 4;
    const int shift = hi ? 24 - 8 * pass : 16 - 8 * (pass - 4);
#pragma unroll
    for (int j = 0; j < kTkIter; ++j) {
      const int i0 = 4 * (tid + kTopkThreads * (rank + kTkCl * j));
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t key = skey[tid * kTkStride + j * 4 + e];
        const uint32_t low = 0xFFFFFFFFu - (uint32_t)(i0 + e);
        const bool in = i0 + e < num_anchors &&
                        (hi ? (key & mask_hi) == pre_hi : (key == pre_hi && (low & mask_lo) == pre_lo));
        // objectness logits share their leading bytes: one shared-memory atomic per distinct bin of the warp
        if (__any_sync(0xffffffffu, in)) {
          const uint32_t bin = in ? (((hi ? key : low) >> shift) & 0xFF) : 0xFFFFFFFFu;
          const unsigned peers = __match_any_sync(0xffffffffu, bin);
          if (in && lane == __ffs(peers) - 1) atomicAdd(&sh.hist[pass][bin], __popc(peers));
        }
      }
    }
    cluster.sync();
    if (tid < 256) {
      int acc = 0;
#pragma unroll
      for (int q = 0; q < kTkCl; ++q) acc += cluster.map_shared_rank(&sh, q)->hist[pass][tid];
      sh.ghist[tid] = acc;
    }
    __syncthreads();
    if (tid < 32) {           // descending walk: lane l owns bins 255 - 8l .. 248 - 8l
      int b8[8], mine = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { b8[j] = sh.ghist[255 - (lane * 8 + j)]; mine += b8[j]; }
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const unsigned ballot = __ballot_sync(0xffffffffu, incl >= need);
      const int src = ballot ? __ffs(ballot) - 1 : 31;
      int acc = incl - mine, pos = lane * 8;
      if (lane == src) {
        for (int j = 0; j < 8 && pos < 255; ++j) {          // the serial rule stops at bin 0 without testing it
          if (acc + b8[j] >= need) break;
          acc += b8[j];
          ++pos;
        }
      }
      pos = __shfl_sync(0xffffffffu, pos, src);
      acc = __shfl_sync(0xffffffffu, acc, src);
      if (lane == 0) {
        sh.digit = (uint32_t)(255 - pos);
        sh.need = need - acc;
      }
    }
    __syncthreads();
    if (hi) { pre_hi |= sh.digit << shift; mask_hi |= 0xFFu << shift; }
    else    { pre_lo |= sh.digit << shift; mask_lo |= 0xFFu << shift; }
    need = sh.need;
  }
  const uint32_t kth_hi = pre_hi, kth_lo = pre_lo;           // composite key of the k-th largest candidate

  // ---- survivors of this CTA -> list, sorted descending
  int mine = 0;
#pragma unroll
  for (int j = 0; j < kTkIter; ++j) {
    const int i0 = 4 * (tid + kTopkThreads * (rank + kTkCl * j));
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t key = skey[tid * kTkStride + j * 4 + e];
      const uint32_t low = 0xFFFFFFFFu - (uint32_t)(i0 + e);
      mine += i0 + e < num_anchors && (key > kth_hi || (key == kth_hi && low >= kth_lo));
    }
  }
  int n_local;
  int slot = dd::block_exclusive_scan(mine, sh.sm32, n_local);
#pragma unroll
  for (int j = 0; j < kTkIter; ++j) {
    const int i0 = 4 * (tid + kTopkThreads * (rank + kTkCl * j));
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t key = skey[tid * kTkStride + j * 4 + e];
      const uint32_t low = 0xFFFFFFFFu - (uint32_t)(i0 + e);
      if (i0 + e < num_anchors && (key > kth_hi || (key == kth_hi && low >= kth_lo)))
        slist[slot++] = ((unsigned long long)key << 32) | (unsigned long long)low;
    }
  }
  const int n_pow2 = next_pow2(max(n_local, 1));
  for (int i = n_local + tid; i < n_pow2; i += kTopkThreads) slist[i] = 0ull;
  if (tid == 0) sh.n_local = n_local;
  __syncthreads();
  bitonic_sort_desc(slist, n_pow2);
  cluster.sync();

  // ---- global rank of every survivor: its local position + the number of larger keys in each other CTA's list
  const unsigned long long* peer_list[kTkCl];
  int peer_n[kTkCl];
#pragma unroll
  for (int q = 0; q < kTkCl; ++q) {
    peer_list[q] = cluster.map_shared_rank(slist, q);
    peer_n[q] = cluster.map_shared_rank(&sh, q)->n_local;
  }
  for (int p = tid; p < n_local; p += kTopkThreads) {
    const unsigned long long x = slist[p];
    int lo[kTkCl], hi2[kTkCl];
#pragma unroll
    for (int q = 0; q < kTkCl; ++q) { lo[q] = 0; hi2[q] = q == rank ? 0 : peer_n[q]; }
    for (int step = 0; step < 15; ++step) {
#pragma unroll
      for (int q = 0; q < kTkCl; ++q) {
        if (lo[q] < hi2[q]) {
          const int mid = (lo[q] + hi2[q]) >> 1;
          if (peer_list[q][mid] > x) lo[q] = mid + 1; else hi2[q] = mid;
        }
      }
    }
    int r = p;
#pragma unroll
    for (int q = 0; q < kTkCl; ++q) r += lo[q];
    srank[p] = r;
  }
  __syncthreads();

  // ---- decode, clip, min-size filter; ordered compaction through the rank bitmap of the whole cluster
  const float xmax = (float)(img_w - 1), ymax = (float)(img_h - 1);
  for (int p = tid; p < n_local; p += kTopkThreads) {
    const int idx = (int)(0xFFFFFFFFu - (uint32_t)(slist[p] & 0xFFFFFFFFull));
    float4 b;
    if (decode_clip_box(deltas, anchors, (size_t)img * num_anchors + idx, idx, xmax, ymax, min_size, b)) {
      const int r = srank[p];
      atomicOr(&sh.okbits[r >> 5], 1u << (r & 31));
    }
  }
  cluster.sync();
  if (tid < 512) {
    uint32_t w = 0u;
#pragma unroll
    for (int q = 0; q < kTkCl; ++q) w |= cluster.map_shared_rank(&sh, q)->okbits[tid];
    sh.allbits[tid] = w;
  }
  __syncthreads();
  if (tid < 32) {
    int cnt[16], mine16 = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) { cnt[j] = __popc(sh.allbits[lane * 16 + j]); mine16 += cnt[j]; }
    int incl = mine16;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    int acc = incl - mine16;
#pragma unroll
    for (int j = 0; j < 16; ++j) { sh.okprefix[lane * 16 + j] = acc; acc += cnt[j]; }
    if (lane == 31 && rank == 0) valid[img] = acc;
  }
  __syncthreads();
  for (int p = tid; p < n_local; p += kTopkThreads) {
    const int r = srank[p];
    const uint32_t w = sh.allbits[r >> 5];
    if (!((w >> (r & 31)) & 1u)) continue;
    const int idx = (int)(0xFFFFFFFFu - (uint32_t)(slist[p] & 0xFFFFFFFFull));
    float4 b;
    decode_clip_box(deltas, anchors, (size_t)img * num_anchors + idx, idx, xmax, ymax, min_size, b);
    const size_t o = (size_t)img * k + sh.okprefix[r >> 5] + __popc(w & ((1u << (r & 31)) - 1u));
    reinterpret_cast<float4*>(boxes)[o] = b;
    scores[o] = 1.0f / (1.0f + expf(-lg[idx]));
    topk_idx[o] = idx;
  }
  cluster.sync();             // no CTA leaves while another may still read its shared memory
}

// ----------------------------------------------------------------------------- NMS
// Reference predicate: inter / (sa + sb - inter) > thresh with an IEEE division (nms.cu:13-24).  The division is only
// executed inside a +-1e-6 band around the threshold; outside it the comparison inter vs thresh * union decides with
// a margin ~16x the rounding error of the quotient, so the result is identical.  Disjoint boxes leave after the
// width test.
__device__ __forceinline__ bool nms_suppresses(const float4 me, const float my_area, const float4 b, const float thresh) {
  const float left = fmaxf(me.x, b.x), right = fminf(me.z, b.z);
  const float top = fmaxf(me.y, b.y), bottom = fminf(me.w, b.w);
  const float width = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f);
  const float height = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
  const float inter = __fmul_rn(width, height);
  const float sb = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
  const float uni = __fsub_rn(__fadd_rn(my_area, sb), inter);
  const float tu = thresh * uni;
  if (!(uni > 0.f) || !(thresh > 0.f)) return __fdiv_rn(inter, uni) > thresh;   // degenerate boxes: as written
  if (inter > tu * 1.000001f) return true;
  if (inter < tu * 0.999999f) return false;
  return __fdiv_rn(inter, uni) > thresh;
}

// mask[i][cb] bit j: box (cb*64+j) is suppressed by box i (j > i only).  Upper triangle only.
// Batched over images (blockIdx.z): image g has n = n_dev ? n_dev[g] : n_cap boxes at boxes + g * n_cap and its own
// [n_cap x cb_cap] mask.
__global__ void __launch_bounds__(64) nms_mask_kernel(const float4* __restrict__ boxes, const int* __restrict__ n_dev,
                                                      int n_cap, float thresh, unsigned long long* __restrict__ mask,
                                                      int cb_cap) {
  const int img = blockIdx.z;
  const int n = n_dev ? min(n_dev[img], n_cap) : n_cap;
  const int rb = blockIdx.y, cb = blockIdx.x;
  if (cb < rb || cb * 64 >= n || rb * 64 >= n) return;
  boxes += (size_t)img * n_cap;
  mask += (size_t)img * n_cap * cb_cap;
  __shared__ float4 cols[64];
  const int col_size = min(n - cb * 64, 64);
  if ((int)threadIdx.x < col_size) cols[threadIdx.x] = boxes[cb * 64 + threadIdx.x];
  __syncthreads();
  const int i = rb * 64 + threadIdx.x;
  if (i < n) {
    const float4 me = boxes[i];
    unsigned long long bits = 0ull;
    const int start = (rb == cb) ? threadIdx.x + 1 : 0;
    const float my_area = __fmul_rn(__fadd_rn(__fsub_rn(me.z, me.x), 1.f), __fadd_rn(__fsub_rn(me.w, me.y), 1.f));
    for (int j = start; j < col_size; ++j) {
      const bool over = nms_suppresses(me, my_area, cols[j], thresh);
      if (over) bits |= 1ull << j;
    }
    mask[(size_t)i * cb_cap + cb] = bits;
  }
}

// Greedy scan over the bitmask, one CTA per image.  Per 64-box block: one thread walks the surviving boxes
// (find-first-set over the not-yet-suppressed bits, the block's diagonal words staged in shared memory and
// prefetched one block ahead); then the whole CTA ORs the mask rows of the boxes just kept into the running
// suppression words of all later blocks — rows are spread over four thread groups so that the L2 loads of one
// block are independent and in flight together.  keep_pos receives kept positions (ascending = score order).
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) nms_scan_kernel(const unsigned long long* __restrict__ mask,
                                                                const int* __restrict__ n_dev, int n_cap, int cb_cap,
                                                                int max_keep, int64_t* __restrict__ keep_pos,
                                                                int keep_stride, int* __restrict__ keep_count) {
  extern __shared__ unsigned long long remv[];   // cb_cap words
  __shared__ unsigned long long diag[2][64];
  __shared__ int kept_rows[64];
  __shared__ int s_cnt, s_nkeep, s_done;
  const int img = blockIdx.x;
  const int n = n_dev ? min(n_dev[img], n_cap) : n_cap;
  const int col_blocks = (n + 63) / 64;
  mask += (size_t)img * n_cap * cb_cap;
  keep_pos += (size_t)img * keep_stride;
  const int tid = threadIdx.x;
  for (int j = tid; j < col_blocks; j += blockDim.x) remv[j] = 0ull;
  if (tid == 0) { s_nkeep = 0; s_done = 0; }
  if (tid < 64 && tid < n) diag[0][tid] = mask[(size_t)tid * cb_cap];
  __syncthreads();
  for (int b = 0; b < col_blocks; ++b) {
    const int size = min(n - b * 64, 64);
    if (tid >= 64 && tid < 128 && b + 1 < col_blocks) {           // next block's diagonal words
      const int r = (b + 1) * 64 + tid - 64;
      if (r < n) diag[(b + 1) & 1][tid - 64] = mask[(size_t)r * cb_cap + (b + 1)];
    }
    if (tid == 0) {
      unsigned long long cur = remv[b];
      if (size < 64) cur |= ~0ull << size;
      int nk = s_nkeep, cnt = 0;
      unsigned long long avail = ~cur;
      while (avail) {
        if (max_keep > 0 && nk >= max_keep) break;
        const int i = __ffsll((long long)avail) - 1;
        kept_rows[cnt++] = i;
        keep_pos[nk++] = (int64_t)(b * 64 + i);
        cur |= diag[b & 1][i];
        avail = ~cur & ~((2ull << i) - 1ull);
      }
      s_cnt = cnt;
      s_nkeep = nk;
      s_done = (max_keep > 0 && nk >= max_keep) ? 1 : 0;
    }
    __syncthreads();
    if (s_done) break;
    const int cnt = s_cnt;
    if (cnt > 0) {
      const int slot = tid >> 8, j = b + 1 + (tid & 255);
      if (j < col_blocks) {
        unsigned long long acc = 0ull;
        const unsigned long long* base = mask + (size_t)(b * 64) * cb_cap + j;
#pragma unroll 4
        for (int r = slot; r < cnt; r += 4) acc |= base[(size_t)kept_rows[r] * cb_cap];
        if (acc) atomicOr(&remv[j], acc);
      }
    }
    __syncthreads();
  }
  if (tid == 0) keep_count[img] = s_nkeep;
}

// ----------------------------------------------------------------------------- NMS, cluster version (no mask)
// The two kernels above compute the whole upper triangle of the suppression matrix (72 M IoU tests for 12 000 boxes,
// on every SM of the GPU) although the greedy scan only ever consults the rows of the boxes it KEEPS, and stops at
// max_keep.  Here a cluster of 8 CTAs owns one image, the matrix never exists and the evaluation is LAZY: when the
// scan reaches a 64-box block, that block is tested against the boxes kept so far — the kept boxes are dealt
// round-robin to the CTAs (resident in shared memory), every CTA tests its share against the 64 candidates and the
// eight partial suppression words meet through distributed shared memory (one cluster barrier per block).  Every CTA
// then resolves the block redundantly (the diagonal 64 x 64 tests, the serial find-first-set walk: identical inputs,
// identical result, nothing to broadcast) and appends the newly kept boxes it owns.  Work = (boxes kept so far) x 64
// per visited block; blocks behind the stopping point cost nothing.
constexpr int kNmsCl = 8;
constexpr int kNmsThreads = 1024;
constexpr int kNmsLocalCap = 1024;                   // kept boxes per CTA held in shared memory (8192 per image)

struct NmsClusterShared {
  float4 kept_box[kNmsLocalCap];
  float kept_area[kNmsLocalCap];
  int kept_idx[kSortCap];                             // position of every kept box (all CTAs hold the full list)
  float4 cols[2][64];
  float cols_area[2][64];
  unsigned long long diag[64];
  uint32_t partial[2][2];                             // this CTA's share of the block's suppression word
  int kept[64];
  int cnt, nkeep, done;
};

__device__ __forceinline__ float box_area_plus1(const float4 b) {
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
}

__global__ void __launch_bounds__(kNmsThreads) nms_cluster_kernel(const float4* __restrict__ boxes,
                                                                  const int* __restrict__ n_dev, int n_cap, float thresh,
                                                                  int max_keep, int64_t* __restrict__ keep_pos,
                                                                  int keep_stride, int* __restrict__ keep_count) {
  extern __shared__ __align__(16) unsigned char dyn[];
  NmsClusterShared& sh = *reinterpret_cast<NmsClusterShared*>(dyn);
  cg::cluster_group cluster = cg::this_cluster();
  const int q = (int)cluster.block_rank();
  const int img = blockIdx.x / kNmsCl;
  const int n = n_dev ? min(n_dev[img], n_cap) : n_cap;
  const int col_blocks = (n + 63) / 64;
  boxes += (size_t)img * n_cap;
  keep_pos += (size_t)img * keep_stride;
  const int tid = threadIdx.x, lane = tid & 31;

  if (tid == 0) { sh.nkeep = 0; sh.done = 0; }
  if (tid < 64) {
    float4 b = make_float4(0.f, 0.f, -1.f, -1.f);
    if (tid < n) b = boxes[tid];
    sh.cols[0][tid] = b;
    sh.cols_area[0][tid] = box_area_plus1(b);
  }
  __syncthreads();
  for (int b = 0; b < col_blocks; ++b) {
    const int buf = b & 1;
    const int size = min(n - b * 64, 64);
    const int nkeep = sh.nkeep;
    if (tid < 2) sh.partial[buf][tid] = 0u;
    if (tid >= 64 && tid < 128 && b + 1 < col_blocks) {           // the next block's boxes
      const int r = (b + 1) * 64 + tid - 64;
      float4 bx = make_float4(0.f, 0.f, -1.f, -1.f);
      if (r < n) bx = boxes[r];
      sh.cols[buf ^ 1][tid - 64] = bx;
      sh.cols_area[buf ^ 1][tid - 64] = box_area_plus1(bx);
    }
    // diagonal tests: 16 threads per row, 4 columns each
    {
      const int i = tid >> 4, j0 = (tid & 15) * 4;
      const float4 me = sh.cols[buf][i];
      const float my_area = sh.cols_area[buf][i];
      unsigned long long bits = 0ull;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = j0 + e;
        if (j > i && j < size && i < size && nms_suppresses(me, my_area, sh.cols[buf][j], thresh)) bits |= 1ull << j;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, o);
      if ((tid & 15) == 0) sh.diag[i] = bits;
    }
    __syncthreads();
    // this CTA's kept boxes (ordinals q, q + 8, ...) against the 64 candidates: thread = (candidate, row group)
    {
      const int c = tid & 63, g = tid >> 6;
      const int n_local = nkeep > q ? (nkeep - q + kNmsCl - 1) / kNmsCl : 0;
      const float4 cb = sh.cols[buf][c];
      bool sup = false;
      if (c < size) {
        for (int l = g; l < n_local && !sup; l += kNmsThreads / 64) {
          float4 kb;
          float ka;
          if (l < kNmsLocalCap) { kb = sh.kept_box[l]; ka = sh.kept_area[l]; }
          else { kb = boxes[sh.kept_idx[l * kNmsCl + q]]; ka = box_area_plus1(kb); }
          sup = nms_suppresses(kb, ka, cb, thresh);
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, sup);
      if (lane == 0 && bal) atomicOr(&sh.partial[buf][(tid >> 5) & 1], bal);
    }
    cluster.sync();                        // every CTA's partial word of this block is final
    if (tid == 0) {
      unsigned long long cur = 0ull;
#pragma unroll
      for (int p = 0; p < kNmsCl; ++p) {
        const NmsClusterShared* r = cluster.map_shared_rank(&sh, p);
        cur |= (unsigned long long)r->partial[buf][0] | ((unsigned long long)r->partial[buf][1] << 32);
      }
      if (size < 64) cur |= ~0ull << size;
      int nk = nkeep, cnt = 0;
      unsigned long long avail = ~cur;
      while (avail) {
        if (max_keep > 0 && nk >= max_keep) break;
        const int i = __ffsll((long long)avail) - 1;
        sh.kept[cnt++] = i;
        ++nk;
        cur |= sh.diag[i];
        avail = ~cur & ~((2ull << i) - 1ull);
      }
      sh.cnt = cnt;
      sh.nkeep = nk;
      sh.done = (max_keep > 0 && nk >= max_keep) ? 1 : 0;
    }
    __syncthreads();
    if (tid < sh.cnt) {                    // append the boxes just kept
      const int ordinal = nkeep + tid, i = sh.kept[tid];
      sh.kept_idx[ordinal] = b * 64 + i;
      if (q == 0) keep_pos[ordinal] = (int64_t)(b * 64 + i);
      if (ordinal % kNmsCl == q && ordinal / kNmsCl < kNmsLocalCap) {
        sh.kept_box[ordinal / kNmsCl] = sh.cols[buf][i];
        sh.kept_area[ordinal / kNmsCl] = sh.cols_area[buf][i];
      }
    }
    __syncthreads();
    if (sh.done) break;
  }
  if (tid == 0 && q == 0) keep_count[img] = sh.nkeep;
  cluster.sync();             // no CTA leaves while another may still read its shared memory
}

// Sort (score desc, index asc) -> order; gather boxes.  Single CTA, n <= kSortCap.
__global__ void __launch_bounds__(1024) sort_scores_gather_kernel(const float* __restrict__ scores,
                                                                  const float4* __restrict__ boxes, int n,
                                                                  int32_t* __restrict__ order,
                                                                  float4* __restrict__ boxes_sorted) {
  extern __shared__ unsigned long long skeys[];
  const int n_pow2 = next_pow2(n);
  for (int i = threadIdx.x; i < n_pow2; i += blockDim.x)
    skeys[i] = i < n ? (((unsigned long long)float_to_ordered(scores[i]) << 32) |
                        (unsigned long long)(0xFFFFFFFFu - (uint32_t)i))
                     : 0ull;
  __syncthreads();
  bitonic_sort_desc(skeys, n_pow2);
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    const int idx = (int)(0xFFFFFFFFu - (uint32_t)(skeys[r] & 0xFFFFFFFFull));
    order[r] = idx;
    boxes_sorted[r] = boxes[idx];
  }
}

// keep positions -> original indices, ascending.  Single CTA.
__global__ void __launch_bounds__(1024) keep_to_sorted_indices_kernel(const int32_t* __restrict__ order,
                                                                      const int* __restrict__ keep_count,
                                                                      int64_t* __restrict__ keep, int cap_pow2) {
  extern __shared__ unsigned long long skeys[];
  const int m = *keep_count;
  const int n_pow2 = next_pow2(m > 1 ? m : 1);
  for (int i = threadIdx.x; i < n_pow2; i += blockDim.x)
    skeys[i] = i < m ? (0xFFFFFFFFFFFFFFFFull - (unsigned long long)order[keep[i]]) : 0ull;
  __syncthreads();
  bitonic_sort_desc(skeys, n_pow2);   // descending in (MAX - idx) == ascending in idx
  for (int i = threadIdx.x; i < m; i += blockDim.x) keep[i] = (int64_t)(0xFFFFFFFFFFFFFFFFull - skeys[i]);
  (void)cap_pow2;
}

int host_next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

struct NmsWs {
  unsigned long long* mask;
  float4* boxes_sorted;
  int32_t* order;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

NmsWs carve(void* ws, int n) {
  const int cb = (n + 63) / 64;
  char* p = (char*)ws;
  NmsWs w;
  w.mask = (unsigned long long*)p;
  p += align_up((size_t)n * cb * 8, 256);
  w.boxes_sorted = (float4*)p;
  p += align_up((size_t)n * 16, 256);
  w.order = (int32_t*)p;
  return w;
}

// images: number of images; n_dev: optional per-image box counts on the device (<= n_cap)
int nms_sorted_impl(const float* boxes_sorted, int images, const int* n_dev, int n_cap, float thresh, int max_keep,
                    int64_t* keep, int keep_stride, int* keep_count, unsigned long long* mask, cudaStream_t s) {
  const int cb = (n_cap + 63) / 64;
  DD_CHECK_ARG(cb <= 256 && images > 0 && images <= 65535);
  static int cluster_mode = -1;          // DD_NMS_CLUSTER=0 keeps the mask + scan kernels (A/B runs)
  if (cluster_mode < 0) {
    const char* e = getenv("DD_NMS_CLUSTER");
    cluster_mode = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  if (cluster_mode == 1 && n_cap >= 2048) {
    const size_t dyn = sizeof(NmsClusterShared);
    static bool configured = false;
    if (!configured) {
      DD_CUDA(cudaFuncSetAttribute(nms_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(images * kNmsCl));
    cfg.blockDim = dim3(kNmsThreads);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kNmsCl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DD_CUDA(cudaLaunchKernelEx(&cfg, nms_cluster_kernel, reinterpret_cast<const float4*>(boxes_sorted), n_dev, n_cap, thresh,
                               max_keep, keep, keep_stride, keep_count));
    DD_LAUNCHED();
    return 0;
  }
  dim3 grid(cb, cb, images);
  nms_mask_kernel<<<grid, 64, 0, s>>>(reinterpret_cast<const float4*>(boxes_sorted), n_dev, n_cap, thresh, mask, cb);
  DD_LAUNCHED();
  nms_scan_kernel<<<images, kScanThreads, cb * sizeof(unsigned long long), s>>>(mask, n_dev, n_cap, cb, max_keep, keep,
                                                                               keep_stride, keep_count);
  DD_LAUNCHED();
  return 0;
}

}  // namespace

extern "C" int dd_anchor_grid(const float* cell_anchors, int A, int FH, int FW, int stride, int img_w, int img_h,
                              int straddle_thresh, float* anchors, uint8_t* visibility, void* stream) {
  DD_CHECK_ARG(A > 0 && FH > 0 && FW > 0 && stride > 0);
  const int total = FH * FW * A;
  anchor_grid_kernel<<<dd::grid_for(total, 256), 256, 0, dd::S(stream)>>>(cell_anchors, A, FH, FW, stride, img_w,
                                                                          img_h, straddle_thresh, anchors, visibility);
  DD_LAUNCHED();
  return 0;
}

extern "C" size_t dd_rpn_topk_workspace_bytes(int N, int num_anchors) {
  (void)N;
  (void)num_anchors;
  return 256;   // everything lives in shared memory; kept for ABI stability
}

extern "C" int dd_rpn_topk_decode(const float* logits, const float* deltas, const float* anchors, int N, int FH,
                                  int FW, int A, int k, int img_w, int img_h, float min_size, float* boxes,
                                  float* scores, int32_t* topk_idx, int32_t* valid, void* workspace, void* stream) {
  (void)workspace;
  const int num_anchors = FH * FW * A;
  DD_CHECK_ARG(N > 0 && num_anchors > 0 && k > 0 && k <= num_anchors && k <= kSortCap);
  static int cluster_mode = -1;          // DD_TOPK_CLUSTER=0 keeps the one-CTA kernel (A/B runs)
  if (cluster_mode < 0) {
    const char* e = getenv("DD_TOPK_CLUSTER");
    cluster_mode = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  if (cluster_mode == 1 && num_anchors > 16384 && num_anchors <= kTkCl * kTopkThreads * 4 * kTkIter && k <= kSortCap) {
    const size_t dyn = (size_t)kTopkThreads * kTkStride * 4 + 8 + (size_t)kSortCap * sizeof(unsigned long long);
    static bool cl_configured = false;
    if (!cl_configured) {
      DD_CUDA(cudaFuncSetAttribute(rpn_topk_decode_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      cl_configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(N * kTkCl));
    cfg.blockDim = dim3(kTopkThreads);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = dd::S(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kTkCl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DD_CUDA(cudaLaunchKernelEx(&cfg, rpn_topk_decode_cluster_kernel, logits, deltas, anchors, num_anchors, k, img_w, img_h,
                               min_size, boxes, scores, topk_idx, valid));
    DD_LAUNCHED();
    return 0;
  }
  const size_t smem = (size_t)host_next_pow2(k) * sizeof(unsigned long long);
  static bool configured = false;
  if (!configured) {
    DD_CUDA(cudaFuncSetAttribute(rpn_topk_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kSortCap * (int)sizeof(unsigned long long)));
    configured = true;
  }
  rpn_topk_decode_kernel<<<N, kTopkThreads, smem, dd::S(stream)>>>(logits, deltas, anchors, num_anchors, k, img_w,
                                                                   img_h, min_size, boxes, scores, topk_idx, valid);
  DD_LAUNCHED();
  return 0;
}

extern "C" size_t dd_nms_workspace_bytes(int n) {
  if (n <= 0) return 256;
  const int cb = (n + 63) / 64;
  return align_up((size_t)n * cb * 8, 256) + align_up((size_t)n * 16, 256) + align_up((size_t)n * 4, 256) + 256;
}

extern "C" int dd_nms_sorted(const float* boxes_sorted, int n, float thresh, int max_keep, int64_t* keep_out,
                             int* keep_count, void* workspace, void* stream) {
  DD_CHECK_ARG(n >= 0 && workspace != nullptr);
  if (n == 0) {
    DD_CUDA(cudaMemsetAsync(keep_count, 0, sizeof(int), dd::S(stream)));
    return 0;
  }
  NmsWs w = carve(workspace, n);
  return nms_sorted_impl(boxes_sorted, 1, nullptr, n, thresh, max_keep, keep_out, 0, keep_count, w.mask, dd::S(stream));
}

extern "C" size_t dd_nms_batched_workspace_bytes(int images, int n_cap) {
  if (n_cap <= 0 || images <= 0) return 256;
  const int cb = (n_cap + 63) / 64;
  return align_up((size_t)images * n_cap * cb * 8, 256) + 256;
}

extern "C" int dd_nms_sorted_batched(const float* boxes_sorted, const int* n_dev, int images, int n_cap, float thresh,
                                     int max_keep, int64_t* keep_out, int keep_stride, int* keep_count,
                                     void* workspace, void* stream) {
  DD_CHECK_ARG(images > 0 && n_cap > 0 && n_cap <= kSortCap && workspace != nullptr && n_dev != nullptr);
  return nms_sorted_impl(boxes_sorted, images, n_dev, n_cap, thresh, max_keep, keep_out, keep_stride, keep_count,
                         (unsigned long long*)workspace, dd::S(stream));
}

extern "C" int dd_nms(const float* boxes, const float* scores, int n, float thresh, int64_t* keep_out,
                      int* keep_count, void* workspace, void* stream) {
  DD_CHECK_ARG(n >= 0 && n <= kSortCap && workspace != nullptr);
  cudaStream_t s = dd::S(stream);
  if (n == 0) {
    DD_CUDA(cudaMemsetAsync(keep_count, 0, sizeof(int), s));
    return 0;
  }
  static bool configured = false;
  if (!configured) {
    DD_CUDA(cudaFuncSetAttribute(sort_scores_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kSortCap * (int)sizeof(unsigned long long)));
    DD_CUDA(cudaFuncSetAttribute(keep_to_sorted_indices_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kSortCap * (int)sizeof(unsigned long long)));
    configured = true;
  }
  NmsWs w = carve(workspace, n);
  const int p2 = host_next_pow2(n);
  sort_scores_gather_kernel<<<1, 1024, (size_t)p2 * 8, s>>>(scores, reinterpret_cast<const float4*>(boxes), n,
                                                            w.order, w.boxes_sorted);
  DD_LAUNCHED();
  int rc = nms_sorted_impl(reinterpret_cast<const float*>(w.boxes_sorted), 1, nullptr, n, thresh, 0, keep_out, 0,
                           keep_count, w.mask, s);
  if (rc) return rc;
  keep_to_sorted_indices_kernel<<<1, 1024, (size_t)p2 * 8, s>>>(w.order, keep_count, keep_out, p2);
  DD_LAUNCHED();
  return 0;
}
