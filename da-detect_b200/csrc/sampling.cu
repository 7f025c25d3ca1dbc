// Device-resident proposal bookkeeping and fg/bg subsampling: fixed-capacity buffers with on-device counts,
// so that one training step has no host synchronisation between the RPN head and the box-head losses (the
// reference reads sizes back on the host in rpn/inference.py:87-127, boxlist_ops.py:30-34,
// balanced_positive_negative_sampler.py:35-76 and box_head/loss.py:55-130).
//
//   dd_proposals_gather   NMS survivors (+ ground-truth boxes for source images, inference.py:51-74) ->
//                         [N, cap, 4] / [N, cap] / count[N]
//   dd_balanced_sample    BalancedPositiveNegativeSampler: `num_pos = min(#pos, max_pos)` positives and
//                         `min(#neg, batch - num_pos)` negatives chosen uniformly at random.  The reference
//                         takes a prefix of randperm(#pos); here every candidate carries a random key and the
//                         smallest keys win (an exact radix select, ties broken by lower index) — the same
//                         distribution, and with keys = rank in a recorded permutation the identical choice.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 1024;
constexpr int kUn = 4;            // load pairs kept in flight per thread in the streaming sweeps

__global__ void __launch_bounds__(256) proposals_gather_kernel(
    const float4* __restrict__ boxes, const float* __restrict__ scores, const int64_t* __restrict__ keep,
    const int* __restrict__ keep_count, const float4* __restrict__ gt, const int* __restrict__ gt_off,
    const int* __restrict__ gt_counts, const uint8_t* __restrict__ append_gt, int k, int post, int cap,
    float4* __restrict__ out_boxes, float* __restrict__ out_obj, int* __restrict__ out_count) {
  const int img = blockIdx.x;
  const int nk = min(keep_count[img], post);
  const int g0 = gt_off[img];
  // gt_counts (optional): GT rows padded to a fixed capacity, the live count of each image on the device
  const int have = gt_counts ? min(max(gt_counts[img], 0), gt_off[img + 1] - g0) : gt_off[img + 1] - g0;
  const int ng = append_gt[img] ? have : 0;
  const int total = min(nk + ng, cap);
  for (int r = threadIdx.x; r < cap; r += blockDim.x) {
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    float s = 0.f;
    if (r < nk) {
      const int pos = (int)keep[(size_t)img * post + r];
      b = boxes[(size_t)img * k + pos];
      s = scores[(size_t)img * k + pos];
    } else if (r < total) {
      b = gt[g0 + r - nk];
      s = 1.f;
    }
    out_boxes[(size_t)img * cap + r] = b;
    out_obj[(size_t)img * cap + r] = s;
  }
  if (threadIdx.x == 0) out_count[img] = total;
}

__device__ __forceinline__ uint32_t ordered_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// class of a candidate: 0 = positive (label >= 1), 1 = negative (label == 0), 2 = ignored
__device__ __forceinline__ int cls_of(int label) { return label >= 1 ? 0 : (label == 0 ? 1 : 2); }

// Exclusive scan of one int per warp across the CTA (32 warps): returns the sum of the values of lower warps.
__device__ __forceinline__ int warp_totals_exclusive(int my_warp_total, int* sm32, int& grand_total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm32[wid] = my_warp_total;
  __syncthreads();
  int v = lane < (int)(blockDim.x >> 5) ? sm32[lane] : 0;
  int incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  grand_total = __shfl_sync(0xffffffffu, incl, 31);
  return __shfl_sync(0xffffffffu, incl - v, wid);
}

// Four consecutive candidates starting at i (i % 4 == 0): class (0 pos / 1 neg / 2 ignored or out of range) and
// ordered key.  128-bit loads when the rows are 16-byte aligned.
__device__ __forceinline__ void load4(const int* __restrict__ labels, const float* __restrict__ keys, int i, int n,
                                      bool vec, int (&c)[4], uint32_t (&k)[4]) {
  if (vec && i + 3 < n) {
    const int4 l = __ldg(reinterpret_cast<const int4*>(labels + i));
    const float4 f = __ldg(reinterpret_cast<const float4*>(keys + i));
    c[0] = cls_of(l.x); c[1] = cls_of(l.y); c[2] = cls_of(l.z); c[3] = cls_of(l.w);
    k[0] = ordered_key(f.x); k[1] = ordered_key(f.y); k[2] = ordered_key(f.z); k[3] = ordered_key(f.w);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool ok = i + e < n;
      c[e] = ok ? cls_of(labels[i + e]) : 2;
      k[e] = ok ? ordered_key(keys[i + e]) : 0u;
    }
  }
}

// One CTA per image.  Every warp owns a CONTIGUOUS slice of the candidates, walked 128 at a time (lane l holds
// candidates 4l..4l+3 of the step: coalesced 128-bit loads, and (lane, element) order IS index order); ordering
// across warps needs one scan of 32 warp totals per phase instead of a block-wide rank per 1024 candidates.
__global__ void __launch_bounds__(kThreads) balanced_sample_kernel(
    const int* __restrict__ labels, const int* __restrict__ n_dev, const float* __restrict__ keys, int n_cap,
    int batch, int max_pos, int64_t* __restrict__ sel_idx, int* __restrict__ counts) {
  __shared__ int hist[2][256];
  __shared__ int sm32[32];
  __shared__ int s_cnt[2];
  __shared__ uint32_t s_prefix[2];
  __shared__ int s_need[2];
  const int img = blockIdx.x;
  labels += (size_t)img * n_cap;
  keys += (size_t)img * n_cap;
  sel_idx += (size_t)img * batch;
  const int n = n_dev ? min(n_dev[img], n_cap) : n_cap;
  const bool vec = (n_cap & 3) == 0 && ((reinterpret_cast<uintptr_t>(labels) | reinterpret_cast<uintptr_t>(keys)) & 15) == 0;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  const int per_warp = ((n + nwarps - 1) / nwarps + 127) & ~127;     // slice length, multiple of 128
  const int w_begin = min(wid * per_warp, n), w_end = min(w_begin + per_warp, n);
  const int step = blockDim.x * 4;

  // ---- population of the two classes
  if (tid < 2) s_cnt[tid] = 0;
  __syncthreads();
  {
    int c0 = 0, c1 = 0;
    // kUn independent 128-bit load pairs per thread in flight: one CTA has to cover the L2 latency by itself
    for (int i0 = tid * 4; i0 < n; i0 += step * kUn) {
      int c[kUn][4]; uint32_t k[kUn][4];
#pragma unroll
      for (int u = 0; u < kUn; ++u) load4(labels, keys, i0 + u * step, n, vec, c[u], k[u]);
#pragma unroll
      for (int u = 0; u < kUn; ++u)
#pragma unroll
        for (int e = 0; e < 4; ++e) { c0 += c[u][e] == 0; c1 += c[u][e] == 1; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, o);
      c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    }
    if (lane == 0) { atomicAdd(&s_cnt[0], c0); atomicAdd(&s_cnt[1], c1); }
  }
  __syncthreads();
  const int n_pos = s_cnt[0], n_neg = s_cnt[1];
  const int num_pos = min(n_pos, max_pos);
  const int num_neg = min(n_neg, batch - num_pos);
  const int want[2] = {num_pos, num_neg};

  // ---- exact radix select (both classes in the same sweeps): key of the want[c]-th smallest candidate
  uint32_t prefix[2] = {0u, 0u}, mask = 0u;
  int need[2] = {want[0], want[1]};
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 512; i += blockDim.x) (&hist[0][0])[i] = 0;
    __syncthreads();
    for (int i0 = tid * 4; i0 < n; i0 += step * kUn) {
      int c[kUn][4]; uint32_t k[kUn][4];
#pragma unroll
      for (int u = 0; u < kUn; ++u) load4(labels, keys, i0 + u * step, n, vec, c[u], k[u]);
#pragma unroll
      for (int u = 0; u < kUn; ++u)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c[u][e] < 2 && (k[u][e] & mask) == prefix[c[u][e]])
            atomicAdd(&hist[c[u][e]][(k[u][e] >> shift) & 0xFF], 1);
    }
    __syncthreads();
    if (tid < 2) {
      const int c = tid;
      int acc = 0, b = 0;
      if (need[c] > 0) {
        for (; b < 255; ++b) {
          if (acc + hist[c][b] >= need[c]) break;
          acc += hist[c][b];
        }
      }
      s_prefix[c] = prefix[c] | ((uint32_t)b << shift);
      s_need[c] = need[c] - acc;
    }
    __syncthreads();
    prefix[0] = s_prefix[0]; prefix[1] = s_prefix[1];
    need[0] = s_need[0]; need[1] = s_need[1];
    mask |= 0xFFu << shift;
    __syncthreads();
  }
  // candidates with key < prefix[c] are taken, plus the first need[c] (by index) with key == prefix[c]

  // ---- phase A: ties (key == prefix) per warp slice and class -> tie rank base of every warp
  int eq_w[2] = {0, 0};
  for (int i0 = w_begin + lane * 4; i0 < w_end; i0 += 128 * kUn) {
    int c[kUn][4]; uint32_t k[kUn][4];
#pragma unroll
    for (int u = 0; u < kUn; ++u) load4(labels, keys, i0 + u * 128, w_end, vec, c[u], k[u]);
#pragma unroll
    for (int u = 0; u < kUn; ++u)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (c[u][e] < 2 && want[c[u][e]] > 0 && k[u][e] == prefix[c[u][e]]) eq_w[c[u][e]] += 1;
  }
  for (int o = 16; o > 0; o >>= 1) {
    eq_w[0] += __shfl_xor_sync(0xffffffffu, eq_w[0], o);
    eq_w[1] += __shfl_xor_sync(0xffffffffu, eq_w[1], o);
  }
  int dummy;
  int tie_base[2];
  tie_base[0] = warp_totals_exclusive(eq_w[0], sm32, dummy);
  tie_base[1] = warp_totals_exclusive(eq_w[1], sm32, dummy);

  // ---- phase B: number of selected candidates per warp slice -> output base of every warp
  // ---- phase C: write them in index order.  (B and C walk the slice identically; `write` switches.)
  int out_base = 0, total = 0;
  const unsigned below = (1u << lane) - 1u;
  for (int write = 0; write < 2; ++write) {
    int ties[2] = {tie_base[0], tie_base[1]};
    int taken = 0;
    int cn[4]; uint32_t kn[4];                       // the next step's candidates are loaded while this one is ranked
    load4(labels, keys, w_begin + lane * 4, w_end, vec, cn, kn);
    for (int i0 = w_begin; i0 < w_end; i0 += 128) {
      const int i = i0 + lane * 4;
      int c[4]; uint32_t k[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) { c[e] = cn[e]; k[e] = kn[e]; }
      load4(labels, keys, i + 128, w_end, vec, cn, kn);
      bool lt[4], eq[4];
      int eqc[2] = {0, 0};                       // ties of each class among this lane's four
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool live = c[e] < 2 && want[c[e] & 1] > 0;
        lt[e] = live && k[e] < prefix[c[e] & 1];
        eq[e] = live && k[e] == prefix[c[e] & 1];
      }
      // tie ranks: ties of lower lanes first, then this lane's earlier elements
      int lane_eq[2] = {0, 0};
#pragma unroll
      for (int e = 0; e < 4; ++e) { lane_eq[0] += eq[e] && c[e] == 0; lane_eq[1] += eq[e] && c[e] == 1; }
      int pre[2] = {lane_eq[0], lane_eq[1]};     // inclusive scan over lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t0 = __shfl_up_sync(0xffffffffu, pre[0], o), t1 = __shfl_up_sync(0xffffffffu, pre[1], o);
        if (lane >= o) { pre[0] += t0; pre[1] += t1; }
      }
      const int tot0 = __shfl_sync(0xffffffffu, pre[0], 31), tot1 = __shfl_sync(0xffffffffu, pre[1], 31);
      int rank[2] = {ties[0] + pre[0] - lane_eq[0], ties[1] + pre[1] - lane_eq[1]};
      bool take[4];
      int my_take = 0;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        take[e] = lt[e];
        if (eq[e]) {
          const int cc = c[e] & 1;
          take[e] = rank[cc] + eqc[cc] < need[cc];
          eqc[cc] += 1;
        }
        my_take += take[e];
      }
      ties[0] += tot0;
      ties[1] += tot1;
      int tpre = my_take;                        // inclusive scan of the take counts over lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, tpre, o);
        if (lane >= o) tpre += t;
      }
      const int ttot = __shfl_sync(0xffffffffu, tpre, 31);
      if (write) {
        int slot = out_base + taken + tpre - my_take;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (take[e]) {
            if (slot < batch) sel_idx[slot] = (int64_t)(i + e);
            ++slot;
          }
        }
      }
      taken += ttot;
    }
    if (!write) out_base = warp_totals_exclusive(taken, sm32, total);
  }
  (void)below;
  total = min(total, batch);
  for (int r = total + tid; r < batch; r += blockDim.x) sel_idx[r] = 0;   // padding rows point at candidate 0
  if (tid == 0) {
    counts[img * 2 + 0] = num_pos;
    counts[img * 2 + 1] = total;
  }
}


// ---- cluster version (large candidate sets: the RPN's 122 880 anchors) ---------------------------------------------
// The one-CTA kernel above walks its candidates eight times through L2 with the memory parallelism of a single SM
// (217 us for 122 880 anchors).  Here a cluster of 8 CTAs owns one image: every CTA loads ONE contiguous slice of the
// candidates into its shared memory (ordered key + class, a single pass over global memory), all later sweeps run out
// of shared memory, and the CTAs meet through distributed shared memory: per radix pass every CTA sums the eight
// partial histograms (so all CTAs walk the same prefix), then the tie / take counts of the lower-ranked CTAs give
// every CTA the output offset of its slice.  Same selection rule and output order as the kernel above.
constexpr int kCl = 8;                 // CTAs per image (portable cluster size)
constexpr int kSliceMax = 16384;       // candidates per CTA held in shared memory
constexpr int kChunkMax = kSliceMax / kThreads;       // 16 candidates per thread
constexpr int kStrideMax = kChunkMax | 1;             // odd stride between the chunks of consecutive threads

struct ClusterShared {
  int hist[4][2][256];       // one histogram per radix pass (never re-used: one cluster barrier per exchange)
  int cnt[2];                // class population of this slice
  int lt[2], eq[2];          // candidates below / equal to the selection key in this slice
  int ghist[2][256];
  int sm32[32];
  uint32_t s_prefix[2];
  int s_need[2];
};

// exclusive scan over the CTA of one int per thread (1024 threads); returns this thread's offset, total in `total`
__device__ __forceinline__ int block_exclusive(int v, int* sm32, int& total) {
  const int lane = threadIdx.x & 31;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
  const int base = warp_totals_exclusive(warp_total, sm32, total);
  return base + incl - v;
}

__global__ void __launch_bounds__(kThreads) balanced_sample_cluster_kernel(
    const int* __restrict__ labels, const int* __restrict__ n_dev, const float* __restrict__ keys, int n_cap,
    int batch, int max_pos, int64_t* __restrict__ sel_idx, int* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ ClusterShared sh;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int img = blockIdx.x / kCl;
  labels += (size_t)img * n_cap;
  keys += (size_t)img * n_cap;
  sel_idx += (size_t)img * batch;
  const int n = n_dev ? min(n_dev[img], n_cap) : n_cap;
  const int per = (((n + kCl - 1) / kCl) + 3) & ~3;                  // slice length, multiple of 4
  const int begin = min(rank * per, n), cnt = min(begin + per, n) - begin;
  const int C = (per + kThreads - 1) / kThreads;                    // candidates per thread (<= kChunkMax)
  const int stride = C | 1;
  uint32_t* skey = reinterpret_cast<uint32_t*>(dyn);
  uint8_t* scls = reinterpret_cast<uint8_t*>(skey + kThreads * kStrideMax);
  const int tid = threadIdx.x;
  const bool vec = (n_cap & 3) == 0 && ((reinterpret_cast<uintptr_t>(labels) | reinterpret_cast<uintptr_t>(keys)) & 15) == 0;

  for (int i = tid; i < 4 * 2 * 256; i += kThreads) (&sh.hist[0][0][0])[i] = 0;
  if (tid < 2) sh.cnt[tid] = 0;
  __syncthreads();
  // ---- the one pass over global memory: slice -> shared memory (thread-chunked layout), class populations
  {
    int c0 = 0, c1 = 0;
    for (int i0 = tid * 4; i0 < cnt; i0 += kThreads * 4) {
      int c[4]; uint32_t k[4];
      load4(labels, keys, begin + i0, begin + cnt, vec, c, k);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = i0 + e;
        if (i < cnt) {
          const int pos = (i / C) * stride + (i % C);
          skey[pos] = k[e];
          scls[pos] = (uint8_t)c[e];
          c0 += c[e] == 0;
          c1 += c[e] == 1;
        }
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      c0 += __shfl_xor_sync(0xffffffffu, c0, o);
      c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    }
    if ((tid & 31) == 0) { atomicAdd(&sh.cnt[0], c0); atomicAdd(&sh.cnt[1], c1); }
  }
  cluster.sync();
  int n_pos = 0, n_neg = 0;
  for (int q = 0; q < kCl; ++q) {
    const ClusterShared* r = cluster.map_shared_rank(&sh, q);
    n_pos += r->cnt[0];
    n_neg += r->cnt[1];
  }
  const int num_pos = min(n_pos, max_pos);
  const int num_neg = min(n_neg, batch - num_pos);
  const int want[2] = {num_pos, num_neg};
  const int my_lo = tid * C, my_hi = min(my_lo + C, cnt);         // this thread's chunk of the slice

  // ---- exact radix select over the whole image, both classes in the same sweeps
  uint32_t prefix[2] = {0u, 0u}, mask = 0u;
  int need[2] = {want[0], want[1]};
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = my_lo; i < my_hi; ++i) {
      const int pos = tid * stride + (i - my_lo);
      const int c = scls[pos];
      const uint32_t k = skey[pos];
      if (c < 2 && (k & mask) == prefix[c]) atomicAdd(&sh.hist[pass][c][(k >> shift) & 0xFF], 1);
    }
    cluster.sync();
    if (tid < 512) {
      int acc = 0;
      for (int q = 0; q < kCl; ++q) acc += (&cluster.map_shared_rank(&sh, q)->hist[pass][0][0])[tid];
      (&sh.ghist[0][0])[tid] = acc;
    }
    __syncthreads();
    if (tid < 64) {                            // warp 0: positives, warp 1: negatives; lane l owns bins 8l .. 8l+7
      const int c = tid >> 5, lane = tid & 31;
      int b8[8], mine = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { b8[j] = sh.ghist[c][lane * 8 + j]; mine += b8[j]; }
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      // first bin b (<= 254) with acc(b) + hist[b] >= need, else 255 (the serial rule of the one-CTA kernel)
      const bool crosses = need[c] > 0 && incl >= need[c];
      const unsigned ballot = __ballot_sync(0xffffffffu, crosses);
      int bsel = 255, acc_sel = 0;
      if (need[c] > 0) {
        if (ballot != 0u) {
          const int src = __ffs(ballot) - 1;
          int acc = incl - mine, b = lane * 8;
          if (lane == src) {
            for (int j = 0; j < 8; ++j) {
              if (acc + b8[j] >= need[c]) break;
              acc += b8[j];
              ++b;
            }
          }
          bsel = __shfl_sync(0xffffffffu, b, src);
          acc_sel = __shfl_sync(0xffffffffu, acc, src);
          if (bsel > 255) bsel = 255;
        }
        if (bsel == 255 && ballot == 0u) acc_sel = __shfl_sync(0xffffffffu, incl, 31) - sh.ghist[c][255];
        if (bsel == 255 && ballot != 0u) {
          // the crossing bin is 255 itself, or the rule stopped at 255: everything below bin 255 is accumulated
          acc_sel = __shfl_sync(0xffffffffu, incl, 31) - sh.ghist[c][255];
        }
      } else {
        bsel = 0;
      }
      if (lane == 0) {
        sh.s_prefix[c] = prefix[c] | ((uint32_t)bsel << shift);
        sh.s_need[c] = need[c] - acc_sel;
      }
    }
    __syncthreads();
    prefix[0] = sh.s_prefix[0]; prefix[1] = sh.s_prefix[1];
    need[0] = sh.s_need[0]; need[1] = sh.s_need[1];
    mask |= 0xFFu << shift;
  }

  // ---- candidates below / equal to the selection key: per thread chunk, per CTA, per cluster
  int lt_t[2] = {0, 0}, eq_t[2] = {0, 0};
  for (int i = my_lo; i < my_hi; ++i) {
    const int pos = tid * stride + (i - my_lo);
    const int c = scls[pos];
    if (c < 2 && want[c] > 0) {
      const uint32_t k = skey[pos];
      lt_t[c] += k < prefix[c];
      eq_t[c] += k == prefix[c];
    }
  }
  int tot_eq[2], tot_lt[2];
  const int eq_base0 = block_exclusive(eq_t[0], sh.sm32, tot_eq[0]);
  const int eq_base1 = block_exclusive(eq_t[1], sh.sm32, tot_eq[1]);
  int dummy_off = block_exclusive(lt_t[0], sh.sm32, tot_lt[0]);
  dummy_off = block_exclusive(lt_t[1], sh.sm32, tot_lt[1]);
  (void)dummy_off;
  if (tid == 0) { sh.eq[0] = tot_eq[0]; sh.eq[1] = tot_eq[1]; sh.lt[0] = tot_lt[0]; sh.lt[1] = tot_lt[1]; }
  cluster.sync();
  int tie_base[2] = {0, 0}, out_base = 0, total = 0;
  {
    int ties[2] = {0, 0};
    for (int q = 0; q < kCl; ++q) {
      const ClusterShared* r = cluster.map_shared_rank(&sh, q);
      int take = 0;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int e = r->eq[c];
        take += r->lt[c] + max(min(need[c] - ties[c], e), 0);
        if (q == rank) tie_base[c] = ties[c];
        ties[c] += e;
      }
      if (q == rank) out_base = total;
      total += take;
    }
  }
  // ---- ordered compaction of this slice: thread chunks are contiguous and ordered, so slots follow the index order
  int my_take = 0;
  {
    int r0 = tie_base[0] + eq_base0, r1 = tie_base[1] + eq_base1;
    for (int i = my_lo; i < my_hi; ++i) {
      const int pos = tid * stride + (i - my_lo);
      const int c = scls[pos];
      if (c < 2 && want[c] > 0) {
        const uint32_t k = skey[pos];
        if (k < prefix[c]) my_take += 1;
        else if (k == prefix[c]) {
          int& r = c == 0 ? r0 : r1;
          my_take += r < need[c];
          r += 1;
        }
      }
    }
  }
  int cta_take;
  int slot = out_base + block_exclusive(my_take, sh.sm32, cta_take);
  {
    int r0 = tie_base[0] + eq_base0, r1 = tie_base[1] + eq_base1;
    for (int i = my_lo; i < my_hi; ++i) {
      const int pos = tid * stride + (i - my_lo);
      const int c = scls[pos];
      if (c < 2 && want[c] > 0) {
        const uint32_t k = skey[pos];
        bool take = k < prefix[c];
        if (k == prefix[c]) {
          int& r = c == 0 ? r0 : r1;
          take = r < need[c];
          r += 1;
        }
        if (take) {
          if (slot < batch) sel_idx[slot] = (int64_t)(begin + i);
          ++slot;
        }
      }
    }
  }
  total = min(total, batch);
  if (rank == 0) {
    for (int r = total + tid; r < batch; r += kThreads) sel_idx[r] = 0;     // padding rows point at candidate 0
    if (tid == 0) {
      counts[img * 2 + 0] = num_pos;
      counts[img * 2 + 1] = total;
    }
  }
  cluster.sync();             // no CTA leaves while another may still read its shared memory
}

}  // namespace

extern "C" int dd_proposals_gather(const float* boxes, const float* scores, const int64_t* keep, const int* keep_count,
                                   const float* gt, const int* gt_offsets, const int* gt_counts,
                                   const uint8_t* append_gt, int N, int k, int post, int cap, float* out_boxes,
                                   float* out_objectness, int* out_count, void* stream) {
  DD_CHECK_ARG(N > 0 && k > 0 && post > 0 && cap >= post);
  proposals_gather_kernel<<<N, 256, 0, dd::S(stream)>>>(
      reinterpret_cast<const float4*>(boxes), scores, keep, keep_count, reinterpret_cast<const float4*>(gt), gt_offsets,
      gt_counts, append_gt, k, post, cap, reinterpret_cast<float4*>(out_boxes), out_objectness, out_count);
  DD_LAUNCHED();
  return 0;
}

extern "C" int dd_balanced_sample(const int* labels, const int* n_dev, const float* keys, int images, int n_cap,
                                  int batch, int max_pos, int64_t* sel_idx, int* counts, void* stream) {
  DD_CHECK_ARG(images > 0 && n_cap > 0 && batch > 0 && max_pos >= 0 && max_pos <= batch);
  // large candidate sets (the RPN's anchors): a cluster of 8 CTAs per image working out of shared memory
  static int cluster_mode = -1;          // DD_SAMPLER_CLUSTER=0 keeps the one-CTA kernel (A/B runs, tests)
  if (cluster_mode < 0) {
    const char* e = getenv("DD_SAMPLER_CLUSTER");
    cluster_mode = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  if (cluster_mode == 1 && n_cap > 8192 && n_cap <= kCl * kSliceMax) {
    const size_t dyn = (size_t)kThreads * kStrideMax * 5;
    static bool configured = false;
    if (!configured) {
      DD_CUDA(cudaFuncSetAttribute(balanced_sample_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
      configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(images * kCl));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = dd::S(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DD_CUDA(cudaLaunchKernelEx(&cfg, balanced_sample_cluster_kernel, labels, n_dev, keys, n_cap, batch, max_pos, sel_idx,
                               counts));
    DD_LAUNCHED();
    return 0;
  }
  balanced_sample_kernel<<<images, kThreads, 0, dd::S(stream)>>>(labels, n_dev, keys, n_cap, batch, max_pos, sel_idx,
                                                               counts);
  DD_LAUNCHED();
  return 0;
}
