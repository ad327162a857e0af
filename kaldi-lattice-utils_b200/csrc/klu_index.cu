// klu_index.cu -- arc-posterior emit, sort-by-key, segmented log-add and output
// ordering for the word-level index tools:
//
//   KLU_SEGMENT     kwsbin2/lattice-word-index-segment.cc:134-177 (accumulate),
//                   :96-128 (flatten + sort by logp desc, word, t0, t1)
//   KLU_POSITION    kwsbin2/lattice-word-index-position.cc:135-190, :100-129
//   KLU_FRAME_POST  latbin/lattice-to-word-frame-post.cc:94-135
//
// The reference accumulates into std::map-of-maps; here every valid (arc[, frame
// or length]) pair emits one (key, value) entry in arc order, a stable per-lattice
// radix sort groups equal keys (so each group is folded in the reference's
// accumulation order up to the level renumbering), a segmented LogAdd reduces
// them, and a second stable sort on the ordered bits of the log-probability
// produces the reference's output order (ties fall back to key order because the
// reduced entries are already key-sorted).
#include <math.h>
#include <string.h>

#include <algorithm>

#include "klu_common.cuh"
#include "klu_sort.cuh"

namespace klu {

namespace {


struct IndexArgs {
  BatchView b;
  CostParams cp;
  int tool;
  int filter_mode, filter_n;
  const int32_t* filter;
  const double* alpha;
  const double* beta;
  const double* alpha2;
  const double* total;
  // --beam
  int use_beam;
  const double* vfwd;
  const double* vbwd;
  const double* best;
  double beam;
  // key layout; drop_key = one bit above every valid key (pruned / unreachable)
  int bits_label, bits_time, bits_len;
  unsigned long long drop_key;
  // entries
  const int64_t* ent_base;  // [L] first entry slot of each lattice
  int32_t* arc_ent_off;     // [E] lattice-local first entry of each out-order arc
  int32_t* ent_cnt;         // [L] emitted entries
  unsigned long long* key;
  unsigned int* idx;
  double* val;
  unsigned int* aux;
};

__device__ __forceinline__ bool label_valid(const IndexArgs& a, int label) {
  if (label == 0) return false;
  if (a.filter_mode == 0) return true;
  int lo = 0, hi = a.filter_n - 1;
  bool found = false;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const int v = a.filter[mid];
    if (v == label) {
      found = true;
      break;
    }
    if (v < label) lo = mid + 1;
    else hi = mid - 1;
  }
  return a.filter_mode == 1 ? found : !found;
}

// entries an out-order arc will emit
__device__ __forceinline__ int arc_entry_count(const IndexArgs& a, int e) {
  const int4 r = a.b.out_rec[e];
  if (a.tool == KLU_FRAME_POST) {
    if (r.w == 0) return 0;
    const int d = a.b.time[r.x] - a.b.time[a.b.out_src[e]];
    return d > 0 ? d : 0;
  }
  if (!label_valid(a, r.w)) return 0;
  if (a.tool == KLU_SEGMENT) return 1;
  const int s = a.b.out_src[e];
  return a.b.band_off[s + 1] - a.b.band_off[s];
}

// One CTA per lattice: exclusive scan of per-arc entry counts.
__global__ void __launch_bounds__(256) k_count_scan(IndexArgs a) {
  __shared__ int warp_sum[8];
  __shared__ int carry_s;
  const int l = blockIdx.x;
  const int e0 = a.b.e_off[l], e1 = a.b.e_off[l + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int tile = e0; tile < e1; tile += 256) {
    const int e = tile + tid;
    const int c = e < e1 ? arc_entry_count(a, e) : 0;
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
    if (e < e1) a.arc_ent_off[e] = add + x - c;
    __syncthreads();
    if (tid == 255) carry_s = add + x;
    __syncthreads();
  }
  if (tid == 0) a.ent_cnt[l] = carry_s;
}

__device__ __forceinline__ bool emit_arc_pruned(const IndexArgs& a, int l, int s, const int4& r) {
  CostParams cp = a.cp;
  cp.float_sum = 0;
  const double cost = rec_cost(r, cp);
  const double fb = __dadd_rn(a.vfwd[s], __dadd_rn(cost, a.vbwd[r.x]));
  return fb > __dadd_rn(a.best[l], a.beam);
}

// grid (tiles, lattices): one thread per out-order arc.
__global__ void __launch_bounds__(256) k_emit(IndexArgs a) {
  const int l = blockIdx.y;
  const int e0 = a.b.e_off[l], e1 = a.b.e_off[l + 1];
  const int64_t base = a.ent_base[l];
  for (int e = e0 + blockIdx.x * blockDim.x + threadIdx.x; e < e1; e += gridDim.x * blockDim.x) {
    const int4 r = a.b.out_rec[e];
    const int s = a.b.out_src[e];
    const int off = a.arc_ent_off[e];
    const unsigned int arc_local = (unsigned int)(e - e0);
    if (a.tool == KLU_SEGMENT) {
      if (!label_valid(a, r.w)) continue;
      const bool dead = a.use_beam && emit_arc_pruned(a, l, s, r);
      // fw[s] + arc_lkh + bw[next], kwsbin2/lattice-word-index-segment.cc:160-162
      const double v = __dadd_rn(__dadd_rn(a.alpha[s], -rec_cost(r, a.cp)), a.beta[r.x]);
      const unsigned long long t0 = (unsigned long long)a.b.time[s], t1 = (unsigned long long)a.b.time[r.x];
      const unsigned long long k = ((((unsigned long long)r.w << a.bits_time) | t0) << a.bits_time) | t1;
      a.key[base + off] = dead ? a.drop_key : k;
      a.val[base + off] = v;
      a.aux[base + off] = arc_local;
      a.idx[base + off] = (unsigned int)off;
    } else if (a.tool == KLU_FRAME_POST) {
      if (r.w == 0) continue;
      const int t0 = a.b.time[s], t1 = a.b.time[r.x];
      if (t1 <= t0) continue;
      // fw[u] + bw[next] - (float)(g + a), latbin/lattice-to-word-frame-post.cc:102-104
      const double v = __dadd_rn(__dadd_rn(a.alpha[s], a.beta[r.x]), -rec_cost(r, a.cp));
      for (int k = t0; k < t1; ++k) {
        const int o = off + (k - t0);
        a.key[base + o] = ((unsigned long long)k << a.bits_label) | (unsigned long long)r.w;
        a.val[base + o] = v;
        a.aux[base + o] = arc_local;
        a.idx[base + o] = (unsigned int)o;
      }
    } else {  // KLU_POSITION
      if (!label_valid(a, r.w)) continue;
      const int w = a.b.band_off[s + 1] - a.b.band_off[s];
      if (w <= 0) continue;
      const bool dead = a.use_beam && emit_arc_pruned(a, l, s, r);
      const int lo = a.b.band_lo[s];
      const double tail = __dadd_rn(-rec_cost(r, a.cp), 0.0);
      for (int i = 0; i < w; ++i) {
        const double al = a.alpha2[a.b.band_off[s] + i];
        // fw[(len,s)] + arc_lkh + bw[next]; beta of the unfolded lattice = beta[next]
        const double v = __dadd_rn(__dadd_rn(al, tail), a.beta[r.x]);
        const unsigned long long k = ((unsigned long long)r.w << a.bits_len) | (unsigned long long)(lo + i);
        const int o = off + i;
        a.key[base + o] = (dead || !(al > neg_inf())) ? a.drop_key : k;
        a.val[base + o] = v;
        a.aux[base + o] = arc_local;
        a.idx[base + o] = (unsigned int)o;
      }
    }
  }
}

struct ReduceArgs {
  BatchView b;
  int tool;
  const int64_t* ent_base;
  const int32_t* ent_cnt;
  const unsigned char* where;
  const unsigned long long *key_a, *key_b;
  const unsigned int *idx_a, *idx_b;
  const double* val;
  const unsigned int* aux;
  const double* total;
  unsigned long long* rkey;
  double* rval;
  unsigned int* raux;
  int32_t* rcnt;
  // second-sort inputs
  unsigned long long* key2;
  unsigned int* idx2;
  int bits_label;
  unsigned long long drop_key;
};

// One CTA per lattice: fold each run of equal keys with LogAdd (in sorted =
// emission order), subtract the lattice total, compact, and write the sort key of
// the output ordering.
__global__ void __launch_bounds__(256) k_reduce(ReduceArgs a) {
  __shared__ int warp_sum[8];
  __shared__ int carry_s;
  const int l = blockIdx.x;
  const int n = a.ent_cnt[l];
  const int64_t base = a.ent_base[l];
  const unsigned long long* key = (a.where[l] ? a.key_b : a.key_a) + base;
  const unsigned int* idx = (a.where[l] ? a.idx_b : a.idx_a) + base;
  const double* val = a.val + base;
  const unsigned int* aux = a.aux + base;
  const double total = a.total[l];
  const int e0 = a.b.e_off[l];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int tile = 0; tile < n; tile += 256) {
    const int i = tile + tid;
    unsigned long long k = a.drop_key;
    bool head = false;
    if (i < n) {
      k = key[i];
      head = k != a.drop_key && (i == 0 || key[i - 1] != k);
    }
    int x = head ? 1 : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    int add = carry_s;
    for (int w = 0; w < warp; ++w) add += warp_sum[w];
    if (head) {
      const int slot = add + x - 1;
      unsigned int j = idx[i];
      double sum = val[j];
      double bestv = sum;
      unsigned int besta = aux[j];
      for (int q = i + 1; q < n && key[q] == k; ++q) {
        j = idx[q];
        const double v = val[j];
        sum = log_add(sum, v);
        if (a.tool == KLU_POSITION) {
          // strict '>' in reference iteration order (input state, arc order):
          // kwsbin2/lattice-word-index-position.cc:178
          const unsigned int ar = aux[j];
          if (v > bestv || (v == bestv && a.b.out_orig[e0 + ar] < a.b.out_orig[e0 + besta])) {
            bestv = v;
            besta = ar;
          }
        }
      }
      double logp = sum - total;
      a.rkey[base + slot] = k;
      a.rval[base + slot] = logp;
      a.raux[base + slot] = besta;
      logp = logp + 0.0;  // -0.0 and +0.0 compare equal in the reference's sort
      if (a.tool == KLU_FRAME_POST) {
        const float f = (float)logp + 0.0f;
        a.key2[base + slot] = ((k >> a.bits_label) << 32) | (unsigned long long)(~ord_f32(f));
      } else {
        a.key2[base + slot] = ~ord_f64(logp);
      }
      a.idx2[base + slot] = (unsigned int)slot;
    }
    __syncthreads();
    if (tid == 255) carry_s = add + x;
    __syncthreads();
  }
  if (tid == 0) a.rcnt[l] = carry_s;
}

// res_off[l] = sum_{l' < l} rcnt[l'] (single block, L is small next to the arcs)
__global__ void __launch_bounds__(1024) k_scan_counts(const int32_t* cnt, int L, int64_t* off) {
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

struct GatherArgs {
  BatchView b;
  int tool;
  const int64_t* ent_base;
  const int32_t* rcnt;
  const int64_t* res_off;
  const unsigned char* where;
  const unsigned int *idx_a, *idx_b;
  const unsigned long long* rkey;
  const double* rval;
  const unsigned int* raux;
  int bits_label, bits_time, bits_len;
  int32_t *c0, *c1, *c2, *c3;
  double* v;
  float* vf;
};

__global__ void __launch_bounds__(256) k_gather(GatherArgs a) {
  const int l = blockIdx.y;
  const int n = a.rcnt[l];
  const int64_t base = a.ent_base[l];
  const int64_t out = a.res_off[l];
  const unsigned int* idx = (a.where[l] ? a.idx_b : a.idx_a) + base;
  const int e0 = a.b.e_off[l];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned int j = idx[i];
    const unsigned long long k = a.rkey[base + j];
    const double logp = a.rval[base + j];
    if (a.tool == KLU_SEGMENT) {
      const unsigned long long tm = (1ULL << a.bits_time) - 1ULL;
      a.c0[out + i] = (int32_t)(k >> (2 * a.bits_time));
      a.c1[out + i] = (int32_t)((k >> a.bits_time) & tm);
      a.c2[out + i] = (int32_t)(k & tm);
      a.v[out + i] = logp;
    } else if (a.tool == KLU_POSITION) {
      const unsigned long long lm = (1ULL << a.bits_len) - 1ULL;
      const int e = e0 + (int)a.raux[base + j];
      a.c0[out + i] = (int32_t)(k >> a.bits_len);
      a.c1[out + i] = (int32_t)(k & lm) + 1;  // 1-based position, :107
      a.c2[out + i] = a.b.time[a.b.out_src[e]];
      a.c3[out + i] = a.b.time[a.b.out_rec[e].x];
      a.v[out + i] = logp;
    } else {  // frame post
      const unsigned long long lm = (1ULL << a.bits_label) - 1ULL;
      a.c0[out + i] = (int32_t)(k >> a.bits_label);
      a.c1[out + i] = (int32_t)(k & lm);
      a.vf[out + i] = (float)logp;
    }
  }
}

int bits_for(int64_t maxv) {
  int b = 1;
  while (b < 63 && ((int64_t)1 << b) <= maxv) ++b;
  return b;
}

}  // namespace

int run_index_tool(klu_ctx* c, int tool, const klu_opts* o) {
  const int32_t L = c->L;
  const bool needs_times = tool != KLU_FWD_BWD;
  if (needs_times)
    for (int32_t l = 0; l < L; ++l)
      if (!c->h_times_ok[l]) {
        // CompactLatticeStateTimes [ext] KALDI_ASSERTs on this
        set_error("lattice " + std::to_string(l) + ": inconsistent state times (lattice is not aligned)");
        return 1;
      }
  const bool use_beam = tool != KLU_FRAME_POST && tool != KLU_FWD_BWD && o->beam != INFINITY;
  if (use_beam && !(o->beam > 0.0f)) {
    set_error("--beam must be positive");  // KALDI_ASSERT(beam > 0.0) in PruneLattice [ext]
    return 1;
  }
  CostParams cp = make_cost_params(o, false);
  if (use_beam) KLU_TRY(run_tropical_sweeps(c, cp));
  KLU_TRY(run_log_sweeps(c, cp, use_beam, o->beam));
  c->h_res_off.assign(L + 1, 0);
  c->last_entries = 0;
  if (tool == KLU_FWD_BWD || L == 0) return 0;
  if (tool == KLU_POSITION) KLU_TRY(run_banded_alpha(c, cp, use_beam, o->beam));

  // ---- entry slots ----
  std::vector<int64_t> ent_base(L + 1, 0);
  for (int32_t l = 0; l < L; ++l) {
    const int64_t cap = tool == KLU_SEGMENT ? (c->h_e_off[l + 1] - c->h_e_off[l])
                        : tool == KLU_FRAME_POST ? c->h_cap_frame[l] : c->h_cap_pos[l];
    if (cap >= ((int64_t)1 << 31)) {
      set_error("lattice " + std::to_string(l) + ": more than 2^31 index entries");
      return 1;
    }
    ent_base[l + 1] = ent_base[l] + cap;
  }
  const int64_t N = std::max<int64_t>(ent_base[L], 1);
  enum { S_BASE = 0, S_ARCOFF, S_CNT, S_KEYA, S_KEYB, S_IDXA, S_IDXB, S_VAL, S_AUX, S_WHERE, S_RCNT, S_R };
  KLU_TRY(c->d_scratch[S_BASE].reserve(sizeof(int64_t) * (L + 1)));
  KLU_TRY(c->d_scratch[S_ARCOFF].reserve(sizeof(int32_t) * std::max<int64_t>(c->E, 1)));
  KLU_TRY(c->d_scratch[S_CNT].reserve(sizeof(int32_t) * L));
  KLU_TRY(c->d_scratch[S_KEYA].reserve(sizeof(int64_t) * N));
  KLU_TRY(c->d_scratch[S_KEYB].reserve(sizeof(int64_t) * N));
  KLU_TRY(c->d_scratch[S_IDXA].reserve(sizeof(int32_t) * N));
  KLU_TRY(c->d_scratch[S_IDXB].reserve(sizeof(int32_t) * N));
  KLU_TRY(c->d_scratch[S_VAL].reserve(sizeof(double) * N));
  KLU_TRY(c->d_scratch[S_AUX].reserve(sizeof(int32_t) * N));
  KLU_TRY(c->d_scratch[S_WHERE].reserve(2 * (size_t)L));
  KLU_TRY(c->d_scratch[S_RCNT].reserve(sizeof(int32_t) * L));
  // reduced entries: key (8) + val (8) + aux (4) per slot
  KLU_TRY(c->d_scratch[S_R].reserve(20 * (size_t)N + 64));
  KLU_CUDA(cudaMemcpyAsync(c->d_scratch[S_BASE].p, ent_base.data(), sizeof(int64_t) * (L + 1),
                           cudaMemcpyHostToDevice, c->stream));
  KLU_CUDA(cudaStreamSynchronize(c->stream));  // ent_base is a stack object

  int fmode = 0, fn = 0;
  if (tool != KLU_FRAME_POST) KLU_TRY(upload_filter(c, o, &fmode, &fn));

  IndexArgs a;
  a.b = c->view();
  a.cp = make_cost_params(o, tool == KLU_FRAME_POST);  // F1 adds g + a in float
  a.tool = tool;
  a.filter_mode = fmode;
  a.filter_n = fn;
  a.filter = c->d_filter.as<int32_t>();
  a.alpha = c->d_alpha.as<double>();
  a.beta = c->d_beta.as<double>();
  a.alpha2 = c->d_alpha2.as<double>();
  a.total = c->d_total.as<double>();
  a.use_beam = use_beam ? 1 : 0;
  a.vfwd = c->d_vfwd.as<double>();
  a.vbwd = c->d_vbwd.as<double>();
  a.best = c->d_best.as<double>();
  a.beam = (double)o->beam;
  a.bits_label = bits_for(c->max_label);
  a.bits_time = bits_for(c->max_time);
  a.bits_len = bits_for(c->max_len);
  a.ent_base = c->d_scratch[S_BASE].as<int64_t>();
  a.arc_ent_off = c->d_scratch[S_ARCOFF].as<int32_t>();
  a.ent_cnt = c->d_scratch[S_CNT].as<int32_t>();
  a.key = c->d_scratch[S_KEYA].as<unsigned long long>();
  a.idx = c->d_scratch[S_IDXA].as<unsigned int>();
  a.val = c->d_scratch[S_VAL].as<double>();
  a.aux = c->d_scratch[S_AUX].as<unsigned int>();
  int key_bits = 0;
  if (tool == KLU_SEGMENT) key_bits = a.bits_label + 2 * a.bits_time;
  else if (tool == KLU_POSITION) key_bits = a.bits_label + a.bits_len;
  else key_bits = a.bits_time + a.bits_label;
  if (key_bits > 63 || (tool == KLU_FRAME_POST && a.bits_time > 31)) {
    set_error("index key does not fit 63 bits (labels/times too large)");
    return 1;
  }
  a.drop_key = 1ULL << key_bits;
  {
    KLU_LAUNCH(c, "k_count_scan");
    k_count_scan<<<L, 256, 0, c->stream>>>(a);
  }
  KLU_TRY(check_launch("k_count_scan"));
  int64_t max_arcs = 0;
  for (int32_t l = 0; l < L; ++l) max_arcs = std::max(max_arcs, c->h_e_off[l + 1] - c->h_e_off[l]);
  const int tiles = (int)std::max<int64_t>(1, std::min<int64_t>((max_arcs + 255) / 256, 64));
  {
    KLU_LAUNCH(c, "k_emit");
    k_emit<<<dim3(tiles, L), 256, 0, c->stream>>>(a);
  }
  KLU_TRY(check_launch("k_emit"));

  SegSortArgs s1;
  s1.seg_base = a.ent_base;
  s1.seg_cnt = a.ent_cnt;
  s1.key_a = c->d_scratch[S_KEYA].as<unsigned long long>();
  s1.val_a = c->d_scratch[S_IDXA].as<unsigned int>();
  s1.key_b = c->d_scratch[S_KEYB].as<unsigned long long>();
  s1.val_b = c->d_scratch[S_IDXB].as<unsigned int>();
  s1.where = c->d_scratch[S_WHERE].as<unsigned char>();
  s1.lo_bit = 0;
  s1.hi_bit = key_bits + 1;  // + the drop bit; degenerate digits are skipped per lattice
  {
    KLU_LAUNCH(c, "k_seg_radix_sort");
    k_seg_radix_sort<<<L, kSortThreads, 0, c->stream>>>(s1);
  }
  KLU_TRY(check_launch("k_seg_radix_sort(keys)"));

  ReduceArgs r;
  r.b = a.b;
  r.tool = tool;
  r.ent_base = a.ent_base;
  r.ent_cnt = a.ent_cnt;
  r.where = s1.where;
  r.key_a = s1.key_a;
  r.key_b = s1.key_b;
  r.idx_a = s1.val_a;
  r.idx_b = s1.val_b;
  r.val = a.val;
  r.aux = a.aux;
  r.total = a.total;
  char* rp = c->d_scratch[S_R].as<char>();
  r.rkey = reinterpret_cast<unsigned long long*>(rp);
  r.rval = reinterpret_cast<double*>(rp + 8 * (size_t)N);
  r.raux = reinterpret_cast<unsigned int*>(rp + 16 * (size_t)N);
  r.rcnt = c->d_scratch[S_RCNT].as<int32_t>();
  // the ordering sort's input pair is separate from the first sort's buffers
  // (the reduce reads those); its ping-pong partner is the then-free A side.
  KLU_TRY(c->d_res[6].reserve(sizeof(int64_t) * N));   // key2 a
  KLU_TRY(c->d_res[7].reserve(sizeof(int32_t) * N));   // idx2 a
  r.key2 = c->d_res[6].as<unsigned long long>();
  r.idx2 = c->d_res[7].as<unsigned int>();
  r.bits_label = a.bits_label;
  r.drop_key = a.drop_key;
  {
    KLU_LAUNCH(c, "k_reduce");
    k_reduce<<<L, 256, 0, c->stream>>>(r);
  }
  KLU_TRY(check_launch("k_reduce"));

  // ---- output ordering: stable sort on key2, ping-pong into the (now free)
  // first-sort buffers ----
  SegSortArgs s2;
  s2.seg_base = a.ent_base;
  s2.seg_cnt = r.rcnt;
  s2.key_a = r.key2;
  s2.val_a = r.idx2;
  s2.key_b = c->d_scratch[S_KEYA].as<unsigned long long>();
  s2.val_b = c->d_scratch[S_IDXA].as<unsigned int>();
  s2.where = c->d_scratch[S_WHERE].as<unsigned char>() + L;
  s2.lo_bit = 0;
  s2.hi_bit = 64;
  {
    KLU_LAUNCH(c, "k_seg_radix_sort");
    k_seg_radix_sort<<<L, kSortThreads, 0, c->stream>>>(s2);
  }
  KLU_TRY(check_launch("k_seg_radix_sort(order)"));

  KLU_TRY(c->d_res[5].reserve(sizeof(int64_t) * (L + 1)));
  {
    KLU_LAUNCH(c, "k_scan_counts");
    k_scan_counts<<<1, 1024, 0, c->stream>>>(r.rcnt, L, c->d_res[5].as<int64_t>());
  }
  KLU_TRY(check_launch("k_scan_counts"));
  for (int i = 0; i < 4; ++i) KLU_TRY(c->d_res[i].reserve(sizeof(int32_t) * N));
  KLU_TRY(c->d_res[4].reserve(sizeof(double) * N));
  GatherArgs g;
  g.b = a.b;
  g.tool = tool;
  g.ent_base = a.ent_base;
  g.rcnt = r.rcnt;
  g.res_off = c->d_res[5].as<int64_t>();
  g.where = s2.where;
  g.idx_a = s2.val_a;
  g.idx_b = s2.val_b;
  g.rkey = r.rkey;
  g.rval = r.rval;
  g.raux = r.raux;
  g.bits_label = a.bits_label;
  g.bits_time = a.bits_time;
  g.bits_len = a.bits_len;
  g.c0 = c->d_res[0].as<int32_t>();
  g.c1 = c->d_res[1].as<int32_t>();
  g.c2 = c->d_res[2].as<int32_t>();
  g.c3 = c->d_res[3].as<int32_t>();
  g.v = c->d_res[4].as<double>();
  g.vf = c->d_res[4].as<float>();
  {
    KLU_LAUNCH(c, "k_gather");
    k_gather<<<dim3(tiles, L), 256, 0, c->stream>>>(g);
  }
  KLU_TRY(check_launch("k_gather"));
  c->last_entries = -1;  // known after klu_result_offsets()
  return 0;
}

}  // namespace klu

using namespace klu;

// Lazily brings the per-lattice result offsets of the last run to the host.
static int ensure_offsets(klu_ctx* c) {
  if (c->last_tool < 0) {
    set_error("no results: klu_run has not succeeded on this batch");
    return 1;
  }
  KLU_CUDA(cudaSetDevice(c->device));
  if (c->last_entries >= 0) {
    KLU_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
  }
  c->h_res_off.resize(c->L + 1);
  KLU_CUDA(cudaMemcpyAsync(c->h_res_off.data(), c->d_res[5].p, sizeof(int64_t) * (c->L + 1), cudaMemcpyDeviceToHost,
                           c->stream));
  KLU_CUDA(cudaStreamSynchronize(c->stream));
  c->last_entries = c->h_res_off[c->L];
  return 0;
}

static int d2h(klu_ctx* c, void* dst, const void* src, size_t bytes) {
  if (!dst || !bytes) return 0;
  KLU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  return 0;
}

extern "C" {

int klu_result_offsets(klu_ctx* c, int64_t* entry_off) {
  KLU_TRY(ensure_offsets(c));
  memcpy(entry_off, c->h_res_off.data(), sizeof(int64_t) * (c->L + 1));
  return 0;
}

int klu_fetch_segment(klu_ctx* c, int32_t* word, int32_t* t0, int32_t* t1, double* logp) {
  if (c->last_tool != KLU_SEGMENT) {
    set_error("klu_fetch_segment: last run was not KLU_SEGMENT");
    return 1;
  }
  KLU_TRY(ensure_offsets(c));
  const size_t n = (size_t)c->last_entries;
  KLU_TRY(d2h(c, word, c->d_res[0].p, n * 4));
  KLU_TRY(d2h(c, t0, c->d_res[1].p, n * 4));
  KLU_TRY(d2h(c, t1, c->d_res[2].p, n * 4));
  KLU_TRY(d2h(c, logp, c->d_res[4].p, n * 8));
  KLU_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int klu_fetch_position(klu_ctx* c, int32_t* word, int32_t* pos, int32_t* t0, int32_t* t1, double* logp) {
  if (c->last_tool != KLU_POSITION) {
    set_error("klu_fetch_position: last run was not KLU_POSITION");
    return 1;
  }
  KLU_TRY(ensure_offsets(c));
  const size_t n = (size_t)c->last_entries;
  KLU_TRY(d2h(c, word, c->d_res[0].p, n * 4));
  KLU_TRY(d2h(c, pos, c->d_res[1].p, n * 4));
  KLU_TRY(d2h(c, t0, c->d_res[2].p, n * 4));
  KLU_TRY(d2h(c, t1, c->d_res[3].p, n * 4));
  KLU_TRY(d2h(c, logp, c->d_res[4].p, n * 8));
  KLU_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int klu_fetch_frame_post(klu_ctx* c, int32_t* num_frames, int32_t* frame, int32_t* word, float* logp) {
  if (c->last_tool != KLU_FRAME_POST) {
    set_error("klu_fetch_frame_post: last run was not KLU_FRAME_POST");
    return 1;
  }
  KLU_TRY(ensure_offsets(c));
  const size_t n = (size_t)c->last_entries;
  if (num_frames) memcpy(num_frames, c->h_num_frames.data(), sizeof(int32_t) * c->L);
  KLU_TRY(d2h(c, frame, c->d_res[0].p, n * 4));
  KLU_TRY(d2h(c, word, c->d_res[1].p, n * 4));
  KLU_TRY(d2h(c, logp, c->d_res[4].p, n * 4));
  KLU_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int klu_fetch_fwd_bwd(klu_ctx* c, double* alpha, double* beta, double* total) {
  if (c->last_tool < 0) {
    set_error("klu_fetch_fwd_bwd: no run");
    return 1;
  }
  KLU_CUDA(cudaSetDevice(c->device));
  std::vector<double> ha(c->S), hb(c->S);
  KLU_CUDA(cudaMemcpyAsync(ha.data(), c->d_alpha.p, sizeof(double) * c->S, cudaMemcpyDeviceToHost, c->stream));
  KLU_CUDA(cudaMemcpyAsync(hb.data(), c->d_beta.p, sizeof(double) * c->S, cudaMemcpyDeviceToHost, c->stream));
  if (total)
    KLU_CUDA(cudaMemcpyAsync(total, c->d_total.p, sizeof(double) * c->L, cudaMemcpyDeviceToHost, c->stream));
  KLU_CUDA(cudaStreamSynchronize(c->stream));
  for (int64_t s = 0; s < c->S; ++s) {  // back to input numbering
    const int32_t n = c->h_old2new[s];
    if (alpha) alpha[s] = ha[n];
    if (beta) beta[s] = hb[n];
  }
  return 0;
}

}  // extern "C"
