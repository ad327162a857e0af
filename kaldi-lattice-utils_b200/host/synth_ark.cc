// synth_ark.cc -- writes the synthetic lattices of bench.py / the tests as a Kaldi
// CompactLattice table (binary or text), so that the drop-in binaries can be timed
// end to end, ark parse and write included (SURVEY.md 8d timing (iii)).
//
//   klu-synth-lattices <frames> <states/frame> <arcs/state> <max-skip> <vocab> <pool>
//                      <window> <eps-prob> <weight-max> <kind> <seed> <n> <lat-wspecifier>
// (bench_cli.py fills the shape parameters from kaldi-lattice-utils_b200/lattice.py SHAPES)
#include <stdlib.h>

#include <iostream>
#include <string>
#include <vector>

#include "kaldi_io.h"

extern "C" {
typedef struct klu_synth_cfg {
  int32_t kind;
  int32_t frames;
  float states_per_frame;
  float arcs_per_state;
  int32_t max_skip;
  int32_t vocab;
  int32_t pool_size;
  int32_t window;
  float eps_prob;
  float weight_max;
} klu_synth_cfg;
int klu_synth_sizes(const klu_synth_cfg* cfg, uint64_t seed, uint64_t first_id, int32_t n, int64_t* state_off,
                    int64_t* arc_off, int nthreads);
int klu_synth_fill(const klu_synth_cfg* cfg, uint64_t seed, uint64_t first_id, int32_t n, const int64_t* state_off,
                   const int64_t* arc_off, int32_t* src, int32_t* dst, int32_t* label, int32_t* dur, float* graph,
                   float* acoustic, float* fin_graph, float* fin_acoustic, int32_t* fin_dur, int nthreads);
}

int main(int argc, char** argv) {
  try {
    if (argc != 14) {
      std::cerr << "usage: klu-synth-lattices frames states/frame arcs/state max-skip vocab pool window eps-prob "
                   "weight-max kind seed n lat-wspecifier\n";
      return 1;
    }
    klu_synth_cfg cfg;
    cfg.frames = atoi(argv[1]);
    cfg.states_per_frame = (float)atof(argv[2]);
    cfg.arcs_per_state = (float)atof(argv[3]);
    cfg.max_skip = atoi(argv[4]);
    cfg.vocab = atoi(argv[5]);
    cfg.pool_size = atoi(argv[6]);
    cfg.window = atoi(argv[7]);
    cfg.eps_prob = (float)atof(argv[8]);
    cfg.weight_max = (float)atof(argv[9]);
    cfg.kind = atoi(argv[10]);
    const uint64_t seed = strtoull(argv[11], nullptr, 0);
    const int32_t n = atoi(argv[12]);
    kio::TableWriter writer(argv[13]);
    const int32_t chunk = 64;
    for (int32_t first = 0; first < n; first += chunk) {
      const int32_t m = std::min(chunk, n - first);
      std::vector<int64_t> so(m + 1), ao(m + 1);
      klu_synth_sizes(&cfg, seed, (uint64_t)first, m, so.data(), ao.data(), 8);
      const size_t S = (size_t)so[m], E = (size_t)ao[m];
      std::vector<int32_t> src(E), dst(E), label(E), dur(E), fdur(S);
      std::vector<float> g(E), a(E), fg(S), fa(S);
      if (klu_synth_fill(&cfg, seed, (uint64_t)first, m, so.data(), ao.data(), src.data(), dst.data(), label.data(),
                         dur.data(), g.data(), a.data(), fg.data(), fa.data(), fdur.data(), 8) != 0) {
        std::cerr << "klu_synth_fill failed\n";
        return 1;
      }
      for (int32_t l = 0; l < m; ++l) {
        kio::CompactLat lat;
        char key[32];
        snprintf(key, sizeof(key), "utt%07d", first + l);
        lat.key = key;
        lat.nstates = (int32_t)(so[l + 1] - so[l]);
        const size_t e0 = (size_t)ao[l], e1 = (size_t)ao[l + 1], s0 = (size_t)so[l], s1 = (size_t)so[l + 1];
        lat.src.assign(src.begin() + e0, src.begin() + e1);
        lat.dst.assign(dst.begin() + e0, dst.begin() + e1);
        lat.label.assign(label.begin() + e0, label.begin() + e1);
        lat.dur.assign(dur.begin() + e0, dur.begin() + e1);
        lat.graph.assign(g.begin() + e0, g.begin() + e1);
        lat.acoustic.assign(a.begin() + e0, a.begin() + e1);
        lat.tids.resize(e1 - e0);
        for (size_t e = e0; e < e1; ++e) lat.tids[e - e0].assign((size_t)dur[e], 1);  // one transition-id per frame
        lat.fin_graph.assign(fg.begin() + s0, fg.begin() + s1);
        lat.fin_acoustic.assign(fa.begin() + s0, fa.begin() + s1);
        lat.fin_dur.assign(fdur.begin() + s0, fdur.begin() + s1);
        lat.fin_tids.resize(s1 - s0);
        for (size_t s = s0; s < s1; ++s) lat.fin_tids[s - s0].assign((size_t)fdur[s], 1);
        std::ostream& os = writer.Begin(lat.key);
        kio::WriteCompactLattice(os, writer.binary(), lat);
        writer.End();
      }
    }
    writer.Close();
    return 0;
  } catch (const std::exception& e) {
    std::cerr << e.what();
    return 1;
  }
}
