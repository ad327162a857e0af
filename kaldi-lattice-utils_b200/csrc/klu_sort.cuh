// klu_sort.cuh -- batched segmented LSD radix sort (SURVEY.md K7/K8 building block).
//
// One CTA sorts one segment (one lattice's index entries) of (u64 key, u32 value)
// pairs, stable, 8 bits per pass, ping-ponging between two global buffers.  All
// digit histograms are taken in ONE read of the keys; passes whose digit is the
// same for every key of the segment are skipped, so keys may be laid out
// generously.  where[l] tells the consumer which buffer holds segment l.
#pragma once
#include "klu_common.cuh"

namespace klu {

constexpr int kSortItems = 4;

template <typename K>
struct SegSortArgsT {
  const int64_t* seg_base;  // [nseg] first element of each segment
  const int32_t* seg_cnt;   // [nseg] elements in each segment
  K* key_a;
  unsigned int* val_a;
  K* key_b;
  unsigned int* val_b;
  unsigned char* where;  // [nseg] out: 0 = result in a, 1 = result in b
  int lo_bit, hi_bit;
};
typedef SegSortArgsT<unsigned long long> SegSortArgs;
typedef SegSortArgsT<unsigned int> SegSortArgs32;  // 32-bit keys: a third less traffic per pass

#ifdef __CUDACC__
// kSortThreads = 512, or 1024 when there are too few segments to fill the SMs with 512-thread CTAs
template <typename K, int kSortThreads>
static __global__ void __launch_bounds__(kSortThreads) k_seg_radix_sort_t(SegSortArgsT<K> a) {
  constexpr int kSortWarps = kSortThreads / 32;
  constexpr int kSortTile = kSortThreads * kSortItems;
  __shared__ unsigned int hist[8][256];
  __shared__ unsigned int bin_base[256];
  __shared__ unsigned int warp_cnt[kSortWarps][256];
  __shared__ int skip_flag;
  const int seg = blockIdx.x;
  const int n = a.seg_cnt[seg];
  const int64_t base = a.seg_base[seg];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int npass = (a.hi_bit - a.lo_bit + 7) / 8;
  if (n <= 1 || npass <= 0) {
    if (tid == 0) a.where[seg] = 0;
    return;
  }
  K* kin = a.key_a + base;
  unsigned int* vin = a.val_a + base;
  K* kout = a.key_b + base;
  unsigned int* vout = a.val_b + base;
  for (int i = tid; i < 8 * 256; i += kSortThreads) (&hist[0][0])[i] = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 4 * kSortThreads) {
    K kk[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {  // four requests in flight per thread
      const int i = i0 + r * kSortThreads + tid;
      kk[r] = i < n ? kin[i] : (K)0;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (i0 + r * kSortThreads + tid >= n) break;
      const K k = kk[r] >> a.lo_bit;
      for (int p = 0; p < npass; ++p) atomicAdd(&hist[p][(unsigned int)(k >> (8 * p)) & 255u], 1u);
    }
  }
  __syncthreads();
  int executed = 0;
  for (int p = 0; p < npass; ++p) {
    const int shift = a.lo_bit + 8 * p;
    if (tid == 0) skip_flag = 0;
    __syncthreads();
    if (tid < 256 && hist[p][tid] == (unsigned)n) skip_flag = 1;
    __syncthreads();
    if (skip_flag) continue;
    // exclusive scan of hist[p] -> bin_base (first 256 threads = 8 warps)
    if (tid < 256) {
      unsigned int v = hist[p][tid];
      unsigned int x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      warp_cnt[0][tid] = x;  // inclusive within warp (scratch)
      __syncwarp();
      bin_base[tid] = x - v;
    }
    __syncthreads();
    if (tid < 256) {
      unsigned int add = 0;
      for (int w = 0; w < warp; ++w) add += warp_cnt[0][w * 32 + 31];
      __syncwarp();
      bin_base[tid] += add;
    }
    __syncthreads();
    // The loads of a tile are issued together (kSortItems independent requests per thread) and
    // one tile ahead: the next tile's keys travel while this one is ranked and scattered.
    K k[kSortItems], kn[kSortItems];
    unsigned int v[kSortItems], vn[kSortItems];
#pragma unroll
    for (int r = 0; r < kSortItems; ++r) {
      const int i = (warp * kSortItems + r) * 32 + lane;
      k[r] = i < n ? kin[i] : (K)0;
      v[r] = i < n ? vin[i] : 0u;
    }
    for (int tile = 0; tile < n; tile += kSortTile) {
      for (int i = lane; i < 256; i += 32) warp_cnt[warp][i] = 0;
      __syncwarp();
      int d[kSortItems];
#pragma unroll
      for (int r = 0; r < kSortItems; ++r) {
        const int i = tile + (warp * kSortItems + r) * 32 + lane;
        const bool valid = i < n;
        d[r] = valid ? (int)((k[r] >> shift) & 255) : 256 + lane;
        const unsigned int mask = __match_any_sync(0xffffffffu, d[r]);
        if (valid && lane == __ffs(mask) - 1) warp_cnt[warp][d[r]] += __popc(mask);
        __syncwarp();
      }
#pragma unroll
      for (int r = 0; r < kSortItems; ++r) {
        const int i = tile + kSortTile + (warp * kSortItems + r) * 32 + lane;
        kn[r] = i < n ? kin[i] : (K)0;
        vn[r] = i < n ? vin[i] : 0u;
      }
      __syncthreads();
      if (tid < 256) {
        unsigned int run = bin_base[tid];
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
          const unsigned int cnt = warp_cnt[w][tid];
          warp_cnt[w][tid] = run;
          run += cnt;
        }
        bin_base[tid] = run;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kSortItems; ++r) {
        const int i = tile + (warp * kSortItems + r) * 32 + lane;
        const bool valid = i < n;
        const unsigned int mask = __match_any_sync(0xffffffffu, d[r]);
        const int leader = __ffs(mask) - 1;
        unsigned int pos = 0;
        if (valid && lane == leader) {
          pos = warp_cnt[warp][d[r]];
          warp_cnt[warp][d[r]] = pos + __popc(mask);
        }
        pos = __shfl_sync(0xffffffffu, pos, leader);
        if (valid) {
          const unsigned int dst = pos + __popc(mask & ((1u << lane) - 1u));
          kout[dst] = k[r];
          vout[dst] = v[r];
        }
        __syncwarp();
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kSortItems; ++r) {
        k[r] = kn[r];
        v[r] = vn[r];
      }
    }
    // swap buffers
    K* tk = kin;
    kin = kout;
    kout = tk;
    unsigned int* tv = vin;
    vin = vout;
    vout = tv;
    ++executed;
    __syncthreads();
  }
  if (tid == 0) a.where[seg] = (unsigned char)(executed & 1);
}
template <typename K>
static inline void seg_sort_launch(const SegSortArgsT<K>& a, int nseg, int num_sms, cudaStream_t stream) {
  if (nseg <= 0) return;
  if (nseg < num_sms * 3)  // fewer CTAs than the SMs can hold at 512 threads: larger CTAs instead
    k_seg_radix_sort_t<K, 1024><<<nseg, 1024, 0, stream>>>(a);
  else
    k_seg_radix_sort_t<K, 512><<<nseg, 512, 0, stream>>>(a);
}

// After a sort that looked at the bits >= a.lo_bit only: every run of elements whose keys agree
// there is put in full-key order by a stable insertion sort, one thread per run (the runs are
// log-posteriors equal to ~1e-6 relative: a handful per segment).  Grid (segments, tiles);
// where[] as k_seg_radix_sort left it; seg0 = first segment of this launch.
static __global__ void __launch_bounds__(256) k_seg_order_fixup(SegSortArgs a, int seg0) {
  const int seg = seg0 + blockIdx.x;
  const int n = a.seg_cnt[seg];
  const int64_t base = a.seg_base[seg];
  unsigned long long* K = (a.where[seg] ? a.key_b : a.key_a) + base;
  unsigned int* V = (a.where[seg] ? a.val_b : a.val_a) + base;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i + 1 < n; i += gridDim.y * blockDim.x) {
    const unsigned long long t = K[i] >> a.lo_bit;
    if ((i > 0 && (K[i - 1] >> a.lo_bit) == t) || (K[i + 1] >> a.lo_bit) != t) continue;  // not the head of a run
    int j = i + 1;
    while (j < n && (K[j] >> a.lo_bit) == t) {  // insert element j into the ordered [i, j)
      const unsigned long long k = K[j];
      const unsigned int v = V[j];
      int q = j;
      while (q > i && K[q - 1] > k) {
        K[q] = K[q - 1];
        V[q] = V[q - 1];
        --q;
      }
      K[q] = k;
      V[q] = v;
      ++j;
    }
  }
}
#endif

}  // namespace klu
