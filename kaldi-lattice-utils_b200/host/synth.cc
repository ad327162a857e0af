// synth.cc -- seeded synthetic lattice generator.
//
// Deterministic in (seed, lattice id): every lattice is generated from its own
// splitmix64 stream, so any sub-range of a workload can be regenerated anywhere
// (tests, bench ranks) bit-identically.  Output is the klu_lattices SoA layout.
//
// Word lattices ("kind 0", configs 2-4): T frames; a few states per frame, ids in
// time order (=> topologically sorted, one consistent time per state, as
// CompactLatticeStateTimes [ext] demands); each state has ~arcs_per_state arcs to
// states 1..max_skip frames ahead (duration = frame distance), plus one guaranteed
// arc into every state so everything is reachable and co-reachable; labels come
// from a time-local pool of the vocabulary (Zipf over pool_size candidates per
// window of `window` frames), eps_prob of the arcs are epsilon; weights U(0,10)
// float32; one final state at frame T.
//
// Char lattices ("kind 1", config 5): HTR-like; one frame per arc; a chain of
// "slots"; each slot has 1..alts alternative character states; label 1 is the
// whitespace delimiter, emitted at word boundaries (word length 1..12).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

extern "C" {

typedef struct klu_synth_cfg {
  int32_t kind;            // 0 word lattice, 1 char lattice
  int32_t frames;          // T (+-10% jitter per lattice)
  float states_per_frame;  // word: mean states per frame; char: mean alternatives per slot
  float arcs_per_state;    // mean out-degree
  int32_t max_skip;        // arcs reach 1..max_skip frames ahead
  int32_t vocab;           // labels in [1, vocab]
  int32_t pool_size;       // candidates per window
  int32_t window;          // frames per label window
  float eps_prob;          // fraction of epsilon arcs
  float weight_max;        // weights ~ U(0, weight_max)
} klu_synth_cfg;

}  // extern "C"

namespace {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  int32_t below(int32_t n) { return (int32_t)(next() % (uint64_t)std::max(n, 1)); }
};

inline uint64_t mix(uint64_t a, uint64_t b) {
  Rng r(a * 0x9E3779B97F4A7C15ULL + b + 0x632BE59BD9B4E019ULL);
  return r.next();
}

struct Out {
  int32_t *src, *dst, *label, *dur;
  float *g, *a, *fg, *fa;
  int32_t* fdur;
};

// Generates one lattice.  When out == nullptr only counts.
void gen_word(const klu_synth_cfg& c, uint64_t seed, uint64_t id, int32_t* ns_out, int64_t* na_out, const Out* out) {
  Rng r(mix(seed, id));
  const int32_t T = std::max(2, (int32_t)lrint(c.frames * (0.9 + 0.2 * r.uni())));
  // states per frame
  std::vector<int32_t> first(T + 2);
  first[0] = 0;
  first[1] = 1;  // frame 0: start state only
  const int32_t base = (int32_t)floor(c.states_per_frame);
  const double frac = c.states_per_frame - base;
  for (int32_t t = 1; t < T; ++t) {
    int32_t k = base + (r.uni() < frac ? 1 : 0) + (r.below(3) - 1);
    first[t + 1] = first[t] + std::max(1, k);
  }
  first[T + 1] = first[T] + 1;  // frame T: the final state
  const int32_t ns = first[T + 1];
  // guaranteed parent of every state v at frame t >= 1: a state of frame t-1
  std::vector<int32_t> parent(ns, -1);
  for (int32_t t = 1; t <= T; ++t)
    for (int32_t v = first[t]; v < first[t + 1]; ++v) parent[v] = first[t - 1] + r.below(first[t] - first[t - 1]);
  int64_t na = 0;
  const int32_t max_skip = std::max(1, c.max_skip);
  const uint64_t pool_seed = mix(seed, 0xC0FFEE);
  auto draw_label = [&](int32_t t) -> int32_t {
    if (r.uni() < c.eps_prob) return 0;
    // Zipf-like rank: log-uniform over [1, pool]
    const int32_t rank = (int32_t)floor(exp(r.uni() * log((double)c.pool_size)));
    const int32_t w = t / std::max(1, c.window);
    return 1 + (int32_t)(mix(pool_seed, (uint64_t)w * 1000003ULL + (uint64_t)rank) % (uint64_t)c.vocab);
  };
  auto put = [&](int32_t u, int32_t v, int32_t tu, int32_t tv, int32_t label) {
    if (out) {
      out->src[na] = u;
      out->dst[na] = v;
      out->label[na] = label;
      out->dur[na] = tv - tu;
      out->g[na] = (float)(r.uni() * c.weight_max);
      out->a[na] = (float)(r.uni() * c.weight_max);
    } else {
      r.uni();
      r.uni();
    }
    ++na;
  };
  for (int32_t t = 0; t < T; ++t) {
    for (int32_t u = first[t]; u < first[t + 1]; ++u) {
      for (int32_t v = first[t + 1]; v < first[t + 2]; ++v)
        if (parent[v] == u) put(u, v, t, t + 1, draw_label(t));
      const int32_t deg = std::max(1, (int32_t)lrint(c.arcs_per_state * (0.5 + r.uni())));
      for (int32_t k = 0; k < deg; ++k) {
        const int32_t tv = std::min(T, t + 1 + r.below(max_skip));
        const int32_t v = first[tv] + r.below(first[tv + 1] - first[tv]);
        put(u, v, t, tv, draw_label(t));
      }
    }
  }
  if (out) {
    for (int32_t s = 0; s < ns; ++s) {
      out->fg[s] = INFINITY;
      out->fa[s] = INFINITY;
      out->fdur[s] = 0;
    }
    out->fg[ns - 1] = (float)(r.uni() * c.weight_max);
    out->fa[ns - 1] = (float)(r.uni() * c.weight_max);
  }
  *ns_out = ns;
  *na_out = na;
}

void gen_char(const klu_synth_cfg& c, uint64_t seed, uint64_t id, int32_t* ns_out, int64_t* na_out, const Out* out) {
  Rng r(mix(seed, id));
  const int32_t T = std::max(2, (int32_t)lrint(c.frames * (0.9 + 0.2 * r.uni())));
  // slot t has alternatives; every state of slot t connects to a subset of slot t+1.
  std::vector<int32_t> first(T + 2), is_ws(T + 1, 0);
  first[0] = 0;
  first[1] = 1;
  int32_t next_ws = 1 + r.below(12);
  for (int32_t t = 1; t < T; ++t) {
    const bool ws = (t == next_ws);
    if (ws) next_ws = t + 2 + r.below(12);
    is_ws[t] = ws;
    const int32_t k = ws ? 1 : std::max(1, (int32_t)lrint(c.states_per_frame * (0.4 + 1.2 * r.uni())));
    first[t + 1] = first[t] + k;
  }
  first[T + 1] = first[T] + 1;
  const int32_t ns = first[T + 1];
  // the label is a property of the destination state (so each state has one
  // incoming label, like a character lattice built from a confusion network)
  std::vector<int32_t> lab(ns, 0);
  for (int32_t t = 1; t <= T; ++t)
    for (int32_t v = first[t]; v < first[t + 1]; ++v)
      lab[v] = is_ws[t] ? 1 : 2 + r.below(std::max(1, c.vocab - 1));
  int64_t na = 0;
  auto put = [&](int32_t u, int32_t v) {
    if (out) {
      out->src[na] = u;
      out->dst[na] = v;
      out->label[na] = lab[v];
      out->dur[na] = 1;
      out->g[na] = (float)(r.uni() * c.weight_max);
      out->a[na] = (float)(r.uni() * c.weight_max);
    } else {
      r.uni();
      r.uni();
    }
    ++na;
  };
  for (int32_t t = 0; t < T; ++t) {
    const int32_t n1 = first[t + 2] - first[t + 1];
    for (int32_t u = first[t]; u < first[t + 1]; ++u) {
      // connect to a random non-empty subset; state u's "own" successor guarantees
      // coverage of slot t+1 states: successor index (u - first[t]) mod n1, and all
      // slot t+1 states beyond the slot t width are attached to the first state.
      const int32_t own = (u - first[t]) % n1;
      for (int32_t k = 0; k < n1; ++k) {
        const bool must = (k == own) || (u == first[t] && k >= first[t + 1] - first[t]);
        if (must || r.uni() < c.arcs_per_state / std::max(1.0f, (float)n1)) put(u, first[t + 1] + k);
      }
    }
  }
  if (out) {
    for (int32_t s = 0; s < ns; ++s) {
      out->fg[s] = INFINITY;
      out->fa[s] = INFINITY;
      out->fdur[s] = 0;
    }
    out->fg[ns - 1] = 0.0f;
    out->fa[ns - 1] = 0.0f;
  }
  *ns_out = ns;
  *na_out = na;
}

void gen(const klu_synth_cfg& c, uint64_t seed, uint64_t id, int32_t* ns, int64_t* na, const Out* out) {
  if (c.kind == 1) gen_char(c, seed, id, ns, na, out);
  else gen_word(c, seed, id, ns, na, out);
}

}  // namespace

extern "C" {

// Fills state_off/arc_off (n+1 entries, starting at 0) for lattices
// first_id .. first_id+n-1.
int klu_synth_sizes(const klu_synth_cfg* cfg, uint64_t seed, uint64_t first_id, int32_t n, int64_t* state_off,
                    int64_t* arc_off, int nthreads) {
  std::vector<int32_t> ns(n);
  std::vector<int64_t> na(n);
  nthreads = std::max(1, std::min(nthreads, n));
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&, t]() {
      for (int32_t i = t; i < n; i += nthreads) gen(*cfg, seed, first_id + i, &ns[i], &na[i], nullptr);
    });
  for (auto& x : th) x.join();
  state_off[0] = 0;
  arc_off[0] = 0;
  for (int32_t i = 0; i < n; ++i) {
    state_off[i + 1] = state_off[i] + ns[i];
    arc_off[i + 1] = arc_off[i] + na[i];
  }
  return 0;
}

// Fills the SoA arrays (sized from klu_synth_sizes).
int klu_synth_fill(const klu_synth_cfg* cfg, uint64_t seed, uint64_t first_id, int32_t n, const int64_t* state_off,
                   const int64_t* arc_off, int32_t* src, int32_t* dst, int32_t* label, int32_t* dur, float* graph,
                   float* acoustic, float* fin_graph, float* fin_acoustic, int32_t* fin_dur, int nthreads) {
  nthreads = std::max(1, std::min(nthreads, n));
  std::vector<std::thread> th;
  std::vector<int> bad(nthreads, 0);
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&, t]() {
      for (int32_t i = t; i < n; i += nthreads) {
        Out o{src + arc_off[i], dst + arc_off[i], label + arc_off[i], dur + arc_off[i], graph + arc_off[i],
              acoustic + arc_off[i], fin_graph + state_off[i], fin_acoustic + state_off[i], fin_dur + state_off[i]};
        int32_t ns;
        int64_t na;
        gen(*cfg, seed, first_id + i, &ns, &na, &o);
        if (ns != state_off[i + 1] - state_off[i] || na != arc_off[i + 1] - arc_off[i]) bad[t] = 1;
      }
    });
  for (auto& x : th) x.join();
  for (int b : bad)
    if (b) return 1;
  return 0;
}

}  // extern "C"
