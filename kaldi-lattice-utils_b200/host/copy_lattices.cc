// klu-copy-lattices: copies a table of CompactLattices (what Kaldi's lattice-copy does for
// compact lattices) through this package's table I/O layer -- text <-> binary conversion,
// and a GPU-free way to exercise the readers and writers.
//
//   klu-copy-lattices [--sequential] [--no-tids] <lattice-rspecifier> <lattice-wspecifier>
//
// --no-tids reads the way every tool but lattice-prune-dyn-beam does (transition-id strings
// dropped, only their lengths kept) and writes each string back as that many ids "1".
//
// Archives held in memory are read in blocks parsed on several threads
// (SequentialCompactLatticeReader::ReadBlock); --sequential forces one entry at a time.
#include <string.h>

#include "kaldi_io.h"

using namespace kio;

int main(int argc, char** argv) {
  try {
    if (argc == 2 && strcmp(argv[1], "--format-selftest") == 0) {  // tests/test_io.py
      const std::string diff = FormatSelfTest(2000000);
      if (!diff.empty()) KIO_ERR("text formatting differs: " << diff);
      KIO_LOG("Text formatting identical to iostream / printf on 6000000 values.");
      return 0;
    }
    bool sequential = false;
    int a = 1;
    if (a < argc && strcmp(argv[a], "--sequential") == 0) {
      sequential = true;
      ++a;
    }
    bool keep_tids = true;
    if (a < argc && strcmp(argv[a], "--no-tids") == 0) {
      keep_tids = false;
      ++a;
    }
    if (argc - a != 2) {
      std::cerr << "Usage: klu-copy-lattices [--sequential] [--no-tids] <lattice-rspecifier> <lattice-wspecifier>\n";
      return 1;
    }
    SequentialCompactLatticeReader reader(argv[a], keep_tids);
    TableWriter writer(argv[a + 1]);
    size_t n = 0;
    auto put = [&](CompactLat& lat) {
      if (!keep_tids) {
        lat.tids.assign(lat.src.size(), TidString());
        lat.fin_tids.assign((size_t)lat.nstates, TidString());
        for (size_t e = 0; e < lat.src.size(); ++e) lat.tids[e].assign((size_t)lat.dur[e], 1);
        for (int32_t s = 0; s < lat.nstates; ++s) lat.fin_tids[s].assign((size_t)lat.fin_dur[s], 1);
      }
      std::ostream& os = writer.Begin(lat.key);
      WriteCompactLattice(os, writer.binary(), lat);
      writer.End();
      ++n;
    };
    std::vector<CompactLat> block;
    while (!reader.Done()) {
      if (!sequential && reader.ReadBlock((int64_t)4 << 20, &block)) {
        for (CompactLat& lat : block) put(lat);
      } else {
        put(reader.Value());
        reader.Next();
      }
    }
    writer.Close();
    KIO_LOG("Copied " << n << " lattices.");
    return 0;
  } catch (const std::exception& e) {
    std::cerr << e.what();
    return 1;
  }
}
