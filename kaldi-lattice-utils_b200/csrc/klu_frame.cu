// klu_frame.cu -- frame-synchronous word posteriorgram (SURVEY.md F1, K6-K8 fused).
//
// Reference: latbin/lattice-to-word-frame-post.cc:94-135 -- for every word arc and
// every frame k it spans, acc[k][word] (a std::map per frame) is LogAdd-ed with
// fw[u] + bw[v] - (float)(g + a); then each frame is normalised, cast to float and
// sorted by (float log-posterior desc, word asc).
//
// Here one warp owns a run of consecutive frames of one lattice.  For a frame it
// walks the frame -> arc CSR built by the packer (coalesced 4-byte ids, 16-byte arc
// records and alpha/beta gathers that hit L1/L2 because neighbouring frames share
// arcs), groups by word in a per-warp shared-memory hash table, and forms each
// group's log-sum DETERMINISTICALLY: an atomic max over the order-preserving bits
// of the doubles, then an integer atomic add of exp(v - max) in 2^-40 fixed point
// (integer addition is associative, so the result does not depend on the order
// lanes hit the table).  The <= 256 survivors are bitonic-sorted in shared memory
// on ((~ordered float bits) << 32 | word) and written next to the frame's slot
// range; a compaction pass produces the dense (frame, word, logp) table.
// Frames with more distinct words than the table holds are split by word hash
// into several passes and sorted in global memory (slow path, same results).
#include <math.h>

#include <algorithm>

#include "klu_common.cuh"

namespace klu {

namespace {

constexpr int kFrameWarps = 4;          // warps per CTA
constexpr int kRegInst = 8;             // instances per lane kept in registers (fast path: n <= 256)
constexpr int kSlots = 256;             // hash slots per warp
constexpr int kFramesPerItem = 16;      // consecutive frames handled by one warp
constexpr double kFixScale = 1099511627776.0;  // 2^40

struct FrameArgs {
  BatchView b;
  CostParams cp;
  const double* alpha;
  const double* beta;
  const double* total;
  int4* arcv;                // [E] per out-order arc: {value f64 (2 words), label, 0}
  const int32_t* item_base;  // [L+1] first work item of each lattice
  int num_items;
  // sparse output: frame k of lattice l owns slots [fr_off[k], fr_off[k+1])
  unsigned long long* ent;   // sort key per slot: (~ord_f32(logp) << 32) | word
  int32_t* frame_cnt;        // per frame slot of fr_off: entries written
  // dense output
  const int64_t* frame_out;  // per frame: first dense entry
  int32_t *o_frame, *o_word;
  float* o_logp;
  int32_t* lat_cnt;          // [L]
  int64_t* res_off;          // [L+1]
};

__device__ __forceinline__ unsigned int hash_word(int w) {
  unsigned int x = (unsigned int)w * 0x9E3779B1u;
  return x ^ (x >> 15);
}

__device__ __forceinline__ float inv_ord_f32(unsigned int u) {
  const unsigned int b = (u & 0x80000000u) ? (u ^ 0x80000000u) : ~u;
  return __uint_as_float(b);
}

// per arc, once: fw[u] + bw[next] - (float)(g + a), latbin/lattice-to-word-frame-post.cc:102-104
__global__ void __launch_bounds__(256) k_arc_values(FrameArgs a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.b.E; e += stride) {
    const int4 r = ld_stream(a.b.out_rec + e);
    const int s = a.b.out_src[e];
    const double v = __dadd_rn(__dadd_rn(a.alpha[s], a.beta[r.x]), -rec_cost(r, a.cp));
    const long long bits = __double_as_longlong(v);
    a.arcv[e] = make_int4((int)(bits & 0xffffffffLL), (int)(bits >> 32), r.w, 0);
  }
}

__device__ __forceinline__ double instance_value(const FrameArgs& a, int e, int* word) {
  const int4 r = __ldg(a.arcv + e);
  *word = r.z;
  return __longlong_as_double(((long long)r.y << 32) | (unsigned int)r.x);
}

// order-preserving bits of v rounded UP to float: a 32-bit running maximum m' >= v
// is all the fixed-point sum needs (exp(v - m') <= 1)
__device__ __forceinline__ unsigned int ord_up_f32(double v) { return ord_f32(__double2float_ru(v)); }

// warp-wide bitonic sort of n_pow2 <= 256 keys in shared memory (ascending); every
// lane owns whole compare-exchange pairs, so no lane idles inside a stage
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long* keys, int n_pow2, int lane) {
  const int half = n_pow2 >> 1;
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int p = lane; p < half; p += 32) {
        const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
        const int q = i | j;
        const unsigned long long x = keys[i], y = keys[q];
        const bool up = (i & k) == 0;
        if ((x > y) == up) {
          keys[i] = y;
          keys[q] = x;
        }
      }
      __syncwarp();
    }
  }
}

// slow path: odd-even transposition sort of a frame's entries in global memory
__device__ void warp_sort_global(unsigned long long* keys, int n, int lane) {
  for (int round = 0; round < n; ++round) {
    for (int i = (round & 1) + 2 * lane; i + 1 < n; i += 64) {
      const unsigned long long x = keys[i], y = keys[i + 1];
      if (x > y) {
        keys[i] = y;
        keys[i + 1] = x;
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kFrameWarps * 32) k_frame_post(FrameArgs a) {
  __shared__ int s_key[kFrameWarps][kSlots];
  __shared__ unsigned int s_max[kFrameWarps][kSlots];
  __shared__ unsigned long long s_sum[kFrameWarps][kSlots];
  __shared__ unsigned long long s_sort[kFrameWarps][kSlots];
  __shared__ int s_list[kFrameWarps][kSlots];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* key = s_key[warp];
  unsigned int* vmax = s_max[warp];
  unsigned long long* vsum = s_sum[warp];
  unsigned long long* sortbuf = s_sort[warp];
  int* list = s_list[warp];
  const BatchView& b = a.b;
  const unsigned int kEmptyMax = 0u;  // below ord_f32 of every float, -inf included
  for (int i = lane; i < kSlots; i += 32) {
    key[i] = -1;
    vmax[i] = kEmptyMax;
    vsum[i] = 0ULL;
  }
  __syncwarp();
  for (int item = blockIdx.x * kFrameWarps + warp; item < a.num_items; item += gridDim.x * kFrameWarps) {
    // lattice of this item: last l with item_base[l] <= item
    int lo = 0, hi = b.L - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (a.item_base[mid] <= item) lo = mid;
      else hi = mid - 1;
    }
    const int l = lo;
    const int T = b.fr_base[l + 1] - b.fr_base[l] - 1;
    const int k0 = (item - a.item_base[l]) * kFramesPerItem;
    const int k1 = min(T, k0 + kFramesPerItem);
    const int64_t* fo = b.fr_off + b.fr_base[l];
    const double total = a.total[l];
    for (int k = k0; k < k1; ++k) {
      const int64_t f0 = fo[k];
      const int n = (int)(fo[k + 1] - f0);
      unsigned long long* out = a.ent + f0;
      int written = 0;
      bool done = false;
      if (n <= kRegInst * 32) {
        // ---- fast path: every instance lives in registers between the phases ----
        double v[kRegInst];
        int slot[kRegInst];
        int ovf = 0;
        int c = 0;  // occupied slots so far (warp-uniform)
#pragma unroll
        for (int r = 0; r < kRegInst; ++r) {
          if (r * 32 < n) {
            const int i = r * 32 + lane;
            slot[r] = -1;
            v[r] = 0.0;
            bool claimed = false;
            unsigned int h = 0;
            if (i < n) {
              int w;
              v[r] = instance_value(a, __ldg(b.frame_arc + f0 + i), &w);
              h = hash_word(w) & (kSlots - 1);
              int probes = 0;
              for (;; h = (h + 1) & (kSlots - 1)) {
                const int old = atomicCAS(&key[h], -1, w);
                if (old == -1) {
                  claimed = true;
                  break;
                }
                if (old == w) break;
                if (++probes >= kSlots) {
                  ovf = 1;
                  break;
                }
              }
              if (!ovf) {
                slot[r] = (int)h;
                atomicMax(&vmax[h], ord_up_f32(v[r]));
              }
            }
            // append the newly claimed slots to the occupied list (no atomics)
            const unsigned int bal = __ballot_sync(0xffffffffu, claimed);
            if (claimed) list[c + __popc(bal & ((1u << lane) - 1u))] = (int)h;
            c += __popc(bal);
          } else {
            slot[r] = -1;
            v[r] = 0.0;
          }
        }
        const bool overflow = __any_sync(0xffffffffu, ovf);
        __syncwarp();
        if (!overflow) {
#pragma unroll
          for (int r = 0; r < kRegInst; ++r) {
            if (slot[r] >= 0) {
              const unsigned int mo = vmax[slot[r]];
              const float mf = inv_ord_f32(mo);
              if (mf > -INFINITY) {
                // v - m' in double, the exponential in float (the output is float32)
                const float t = __expf((float)(v[r] - (double)mf));
                atomicAdd(&vsum[slot[r]], (unsigned long long)__float2ll_rn(t * (float)kFixScale));
              }
            }
          }
          __syncwarp();
        }
        if (!overflow) {
          int np2 = 1;
          while (np2 < c) np2 <<= 1;
          for (int i = lane; i < np2; i += 32) {
            unsigned long long sk = ~0ULL;
            if (i < c) {
              const int h = list[i];
              const float mf = inv_ord_f32(vmax[h]);
              double lse = neg_inf();
              if (mf > -INFINITY) lse = (double)mf + (double)__logf((float)vsum[h] * (float)(1.0 / kFixScale));
              const float f = (float)(lse - total) + 0.0f;
              sk = ((unsigned long long)(~ord_f32(f)) << 32) | (unsigned int)key[h];
            }
            sortbuf[i] = sk;
          }
          __syncwarp();
          warp_bitonic_sort(sortbuf, np2, lane);
          for (int i = lane; i < c; i += 32) out[i] = sortbuf[i];
          written = c;
          done = true;
        }
        // reset the touched slots for the next frame
        for (int i = lane; i < c; i += 32) {
          const int h = list[i];
          key[h] = -1;
          vmax[h] = kEmptyMax;
          vsum[h] = 0ULL;
        }
        __syncwarp();
      }
      // ---- slow path: word-hash partitions, values recomputed per phase, global sort ----
      for (int P = 1; !done; P <<= 1) {
        bool overflow = false;
        written = 0;
        for (int p = 0; p < P && !overflow; ++p) {
          int ovf = 0;
          for (int i = lane; i < n; i += 32) {
            int w;
            const double vv = instance_value(a, __ldg(b.frame_arc + f0 + i), &w);
            const unsigned int h0 = hash_word(w);
            if ((int)((h0 >> 8) & (unsigned)(P - 1)) != p) continue;
            unsigned int h = h0 & (kSlots - 1);
            int probes = 0;
            for (;; h = (h + 1) & (kSlots - 1)) {
              const int old = atomicCAS(&key[h], -1, w);
              if (old == -1 || old == w) break;
              if (++probes >= kSlots) {
                ovf = 1;
                break;
              }
            }
            if (!ovf) atomicMax(&vmax[h], ord_up_f32(vv));
          }
          overflow = __any_sync(0xffffffffu, ovf);
          __syncwarp();
          if (!overflow) {
            for (int i = lane; i < n; i += 32) {
              int w;
              const double vv = instance_value(a, __ldg(b.frame_arc + f0 + i), &w);
              const unsigned int h0 = hash_word(w);
              if ((int)((h0 >> 8) & (unsigned)(P - 1)) != p) continue;
              unsigned int h = h0 & (kSlots - 1);
              while (key[h] != w) h = (h + 1) & (kSlots - 1);
              const float mf = inv_ord_f32(vmax[h]);
              if (mf > -INFINITY)
                atomicAdd(&vsum[h], (unsigned long long)__double2ll_rn(exp(vv - (double)mf) * kFixScale));
            }
            __syncwarp();
            for (int base = 0; base < kSlots; base += 32) {
              const int i = base + lane;
              const int w = key[i];
              unsigned long long sk = 0;
              if (w != -1) {
                const float mf = inv_ord_f32(vmax[i]);
                double lse = neg_inf();
                if (mf > -INFINITY) lse = (double)mf + log((double)vsum[i] * (1.0 / kFixScale));
                const float f = (float)(lse - total) + 0.0f;
                sk = ((unsigned long long)(~ord_f32(f)) << 32) | (unsigned int)w;
              }
              const unsigned int bal = __ballot_sync(0xffffffffu, w != -1);
              if (w != -1) out[written + __popc(bal & ((1u << lane) - 1u))] = sk;
              written += __popc(bal);
            }
          }
          __syncwarp();
          for (int i = lane; i < kSlots; i += 32) {
            key[i] = -1;
            vmax[i] = kEmptyMax;
            vsum[i] = 0ULL;
          }
          __syncwarp();
        }
        if (!overflow) {
          warp_sort_global(out, written, lane);
          done = true;
        }
      }
      if (lane == 0) a.frame_cnt[b.fr_base[l] + k] = written;
      __syncwarp();
    }
  }
}

// One CTA per lattice: scan the per-frame counts into lattice-local dense offsets.
__global__ void __launch_bounds__(256) k_frame_scan(FrameArgs a, int64_t* frame_out_local) {
  __shared__ int warp_sum[8];
  __shared__ int carry_s;
  const int l = blockIdx.x;
  const int f0 = a.b.fr_base[l], T = a.b.fr_base[l + 1] - f0 - 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int tile = 0; tile < T; tile += 256) {
    const int k = tile + tid;
    const int c = k < T ? a.frame_cnt[f0 + k] : 0;
    int x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    int add = carry_s;
    for (int w = 0; w < warp; ++w) add += warp_sum[w];
    if (k < T) frame_out_local[f0 + k] = add + x - c;
    __syncthreads();
    if (tid == 255) carry_s = add + x;
    __syncthreads();
  }
  if (tid == 0) a.lat_cnt[l] = carry_s;
}

__global__ void __launch_bounds__(1024) k_frame_latscan(const int32_t* cnt, int L, int64_t* off) {
  __shared__ long long warp_sum[32];
  __shared__ long long carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int tile = 0; tile < L; tile += 1024) {
    const int i = tile + tid;
    const long long c = i < L ? cnt[i] : 0;
    long long x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    long long add = carry_s;
    for (int w = 0; w < warp; ++w) add += warp_sum[w];
    if (i < L) off[i] = add + x - c;
    __syncthreads();
    if (tid == 1023) carry_s = add + x;
    __syncthreads();
  }
  if (tid == 0) off[L] = carry_s;
}

// dense (frame, word, logp) rows: one warp per (lattice, frame run)
__global__ void __launch_bounds__(256) k_frame_compact(FrameArgs a, const int64_t* frame_out_local) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const BatchView& b = a.b;
  for (int item = blockIdx.x * 8 + warp; item < a.num_items; item += gridDim.x * 8) {
    int lo = 0, hi = b.L - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (a.item_base[mid] <= item) lo = mid;
      else hi = mid - 1;
    }
    const int l = lo;
    const int T = b.fr_base[l + 1] - b.fr_base[l] - 1;
    const int k0 = (item - a.item_base[l]) * kFramesPerItem;
    const int k1 = min(T, k0 + kFramesPerItem);
    const int64_t base = a.res_off[l];
    for (int k = k0; k < k1; ++k) {
      const int fs = b.fr_base[l] + k;
      const int n = a.frame_cnt[fs];
      const unsigned long long* src = a.ent + b.fr_off[fs];
      const int64_t dst = base + frame_out_local[fs];
      for (int i = lane; i < n; i += 32) {
        const unsigned long long sk = src[i];
        a.o_frame[dst + i] = k;
        a.o_word[dst + i] = (int32_t)(sk & 0xffffffffu);
        a.o_logp[dst + i] = inv_ord_f32(~(unsigned int)(sk >> 32));
      }
    }
  }
}

}  // namespace

int run_frame_post(klu_ctx* c, const klu_opts* o) {
  const int32_t L = c->L;
  for (int32_t l = 0; l < L; ++l)
    if (!c->h_times_ok[l]) {
      set_error("lattice " + std::to_string(l) + ": inconsistent state times (lattice is not aligned)");
      return 1;
    }
  CostParams cp = make_cost_params(o, false);
  KLU_TRY(run_log_sweeps(c, cp, false, 0.f));
  c->h_res_off.assign(L + 1, 0);
  c->last_entries = 0;
  if (L == 0) return 0;
  const int64_t N = std::max<int64_t>(c->frame_entries, 1);
  const int64_t F = std::max<int64_t>(c->h_fr_base[L], 1);
  std::vector<int32_t> item_base(L + 1, 0);
  for (int32_t l = 0; l < L; ++l)
    item_base[l + 1] = item_base[l] + (c->h_num_frames[l] + kFramesPerItem - 1) / kFramesPerItem;
  enum { F_ITEM = 0, F_ENT, F_CNT, F_FOUT, F_LCNT, F_ARCV };
  KLU_TRY(c->d_scratch[F_ARCV].reserve(16 * (size_t)std::max<int64_t>(c->E, 1)));
  KLU_TRY(c->d_scratch[F_ITEM].reserve(4 * (size_t)(L + 1)));
  KLU_TRY(c->d_scratch[F_ENT].reserve(8 * (size_t)N));
  KLU_TRY(c->d_scratch[F_CNT].reserve(4 * (size_t)F));
  KLU_TRY(c->d_scratch[F_FOUT].reserve(8 * (size_t)F));
  KLU_TRY(c->d_scratch[F_LCNT].reserve(4 * (size_t)L));
  KLU_TRY(c->d_res[5].reserve(8 * (size_t)(L + 1)));
  KLU_TRY(c->d_res[0].reserve(4 * (size_t)N));
  KLU_TRY(c->d_res[1].reserve(4 * (size_t)N));
  KLU_TRY(c->d_res[4].reserve(4 * (size_t)N));
  KLU_CUDA(cudaMemcpyAsync(c->d_scratch[F_ITEM].p, item_base.data(), 4 * (size_t)(L + 1), cudaMemcpyHostToDevice,
                           c->stream));
  KLU_CUDA(cudaStreamSynchronize(c->stream));
  FrameArgs a;
  a.b = c->view();
  a.cp = make_cost_params(o, true);  // the arc term is (float)(g + a)
  a.alpha = c->d_alpha.as<double>();
  a.beta = c->d_beta.as<double>();
  a.total = c->d_total.as<double>();
  a.arcv = c->d_scratch[F_ARCV].as<int4>();
  a.item_base = c->d_scratch[F_ITEM].as<int32_t>();
  a.num_items = item_base[L];
  a.ent = c->d_scratch[F_ENT].as<unsigned long long>();
  a.frame_cnt = c->d_scratch[F_CNT].as<int32_t>();
  a.frame_out = nullptr;
  a.o_frame = c->d_res[0].as<int32_t>();
  a.o_word = c->d_res[1].as<int32_t>();
  a.o_logp = c->d_res[4].as<float>();
  a.lat_cnt = c->d_scratch[F_LCNT].as<int32_t>();
  a.res_off = c->d_res[5].as<int64_t>();
  {
    KLU_LAUNCH(c, "k_arc_values");
    k_arc_values<<<c->num_sms * 8, 256, 0, c->stream>>>(a);
  }
  KLU_TRY(check_launch("k_arc_values"));
  if (a.num_items > 0) {
    const int grid = std::max(1, std::min((a.num_items + kFrameWarps - 1) / kFrameWarps, c->num_sms * 64));
    {
      KLU_LAUNCH(c, "k_frame_post");
      k_frame_post<<<grid, kFrameWarps * 32, 0, c->stream>>>(a);
    }
    KLU_TRY(check_launch("k_frame_post"));
  }
  {
    KLU_LAUNCH(c, "k_frame_scan");
    k_frame_scan<<<L, 256, 0, c->stream>>>(a, c->d_scratch[F_FOUT].as<int64_t>());
  }
  KLU_TRY(check_launch("k_frame_scan"));
  {
    KLU_LAUNCH(c, "k_scan_counts");
    k_frame_latscan<<<1, 1024, 0, c->stream>>>(a.lat_cnt, L, a.res_off);
  }
  KLU_TRY(check_launch("k_frame_latscan"));
  if (a.num_items > 0) {
    const int grid = std::max(1, std::min((a.num_items + 7) / 8, c->num_sms * 32));
    {
      KLU_LAUNCH(c, "k_frame_compact");
      k_frame_compact<<<grid, 256, 0, c->stream>>>(a, c->d_scratch[F_FOUT].as<int64_t>());
    }
    KLU_TRY(check_launch("k_frame_compact"));
  }
  c->last_entries = -1;
  return 0;
}

}  // namespace klu
