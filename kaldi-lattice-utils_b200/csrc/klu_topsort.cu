// klu_topsort.cu -- host-side helper of the C ABI mirroring
// TopSortCompactLatticeIfNeeded [ext Kaldi lat/lattice-functions.h], called by every
// tool of the reference before its sweeps (e.g. kwsbin2/lattice-word-index-position.cc:61).
//
// fst::TopSort [ext OpenFst]: depth-first search from the start state, then from
// every still unvisited state in id order; new ids = reverse finishing order.  Arcs
// keep their relative order inside each state and are regrouped by new source.
#include <algorithm>
#include <numeric>
#include <vector>

#include "klu_common.cuh"

using namespace klu;

extern "C" int klu_topsort(int32_t nstates, int64_t narcs, int32_t* arc_src, int32_t* arc_dst, int32_t* arc_label,
                           int32_t* arc_dur, float* arc_graph, float* arc_acoustic, float* fin_graph,
                           float* fin_acoustic, int32_t* fin_dur, int32_t* order_out) {
  if (nstates < 0 || narcs < 0) {
    set_error("klu_topsort: negative size");
    return 1;
  }
  bool sorted = true, grouped = true;
  for (int64_t e = 0; e < narcs; ++e) {
    if (arc_src[e] < 0 || arc_src[e] >= nstates || arc_dst[e] < 0 || arc_dst[e] >= nstates) {
      set_error("klu_topsort: arc " + std::to_string(e) + " references a state outside [0, nstates)");
      return 1;
    }
    if (arc_src[e] >= arc_dst[e]) sorted = false;
    if (e > 0 && arc_src[e] < arc_src[e - 1]) grouped = false;
  }
  if (sorted && grouped) {  // the "IfNeeded" part: nothing moves
    if (order_out) std::iota(order_out, order_out + nstates, 0);
    return 0;
  }
  // arcs of each state, in stored order
  std::vector<int64_t> first(nstates + 1, 0);
  for (int64_t e = 0; e < narcs; ++e) first[arc_src[e] + 1]++;
  for (int32_t s = 0; s < nstates; ++s) first[s + 1] += first[s];
  std::vector<int64_t> by_src(narcs), cursor(first.begin(), first.end() - 1);
  for (int64_t e = 0; e < narcs; ++e) by_src[cursor[arc_src[e]]++] = e;
  // iterative DFS
  std::vector<uint8_t> color(nstates, 0);  // 0 white, 1 grey, 2 black
  std::vector<int32_t> finish;
  finish.reserve(nstates);
  std::vector<std::pair<int32_t, int64_t> > stack;
  auto visit = [&](int32_t root) -> bool {
    if (color[root]) return true;
    color[root] = 1;
    stack.push_back(std::make_pair(root, first[root]));
    while (!stack.empty()) {
      const int32_t s = stack.back().first;
      int64_t& k = stack.back().second;
      if (k < first[s + 1]) {
        const int32_t d = arc_dst[by_src[k++]];
        if (color[d] == 1) return false;  // back edge
        if (color[d] == 0) {
          color[d] = 1;
          stack.push_back(std::make_pair(d, first[d]));
        }
      } else {
        color[s] = 2;
        finish.push_back(s);
        stack.pop_back();
      }
    }
    return true;
  };
  bool acyclic = nstates == 0 || visit(0);
  for (int32_t s = 0; acyclic && s < nstates; ++s) acyclic = visit(s);
  if (!acyclic) {
    set_error("klu_topsort: the lattice is cyclic");  // KALDI_ERR in the reference (fstext/fstext-utils2.h:118-121)
    return 1;
  }
  std::vector<int32_t> new_id(nstates);
  for (int32_t i = 0; i < nstates; ++i) new_id[finish[nstates - 1 - i]] = i;
  if (order_out) std::copy(new_id.begin(), new_id.end(), order_out);
  // permute the per-state arrays
  auto permute_states = [&](auto* arr) {
    if (!arr) return;
    std::vector<typename std::remove_pointer<decltype(arr)>::type> tmp(arr, arr + nstates);
    for (int32_t s = 0; s < nstates; ++s) arr[new_id[s]] = tmp[s];
  };
  permute_states(fin_graph);
  permute_states(fin_acoustic);
  permute_states(fin_dur);
  // arcs: stable by new source (the stored order inside each state survives)
  std::vector<int64_t> perm(narcs);
  auto permute_arcs = [&](auto* arr) {
    if (!arr) return;
    std::vector<typename std::remove_pointer<decltype(arr)>::type> tmp(arr, arr + narcs);
    for (int64_t e = 0; e < narcs; ++e) arr[e] = tmp[perm[e]];
  };
  for (int64_t e = 0; e < narcs; ++e) {
    arc_src[e] = new_id[arc_src[e]];
    arc_dst[e] = new_id[arc_dst[e]];
  }
  std::iota(perm.begin(), perm.end(), (int64_t)0);
  std::stable_sort(perm.begin(), perm.end(), [&](int64_t x, int64_t y) { return arc_src[x] < arc_src[y]; });
  permute_arcs(arc_src);
  permute_arcs(arc_dst);
  permute_arcs(arc_label);
  permute_arcs(arc_dur);
  permute_arcs(arc_graph);
  permute_arcs(arc_acoustic);
  return 0;
}
