// klu-copy-lattices: copies a table of CompactLattices (what Kaldi's lattice-copy does for
// compact lattices) through this package's table I/O layer -- text <-> binary conversion,
// and a GPU-free way to exercise the readers and writers.
//
//   klu-copy-lattices [--sequential] <lattice-rspecifier> <lattice-wspecifier>
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
    if (argc - a != 2) {
      std::cerr << "Usage: klu-copy-lattices [--sequential] <lattice-rspecifier> <lattice-wspecifier>\n";
      return 1;
    }
    SequentialCompactLatticeReader reader(argv[a]);
    TableWriter writer(argv[a + 1]);
    size_t n = 0;
    auto put = [&](const CompactLat& lat) {
      std::ostream& os = writer.Begin(lat.key);
      WriteCompactLattice(os, writer.binary(), lat);
      writer.End();
      ++n;
    };
    std::vector<CompactLat> block;
    while (!reader.Done()) {
      if (!sequential && reader.ReadBlock((int64_t)4 << 20, &block)) {
        for (const CompactLat& lat : block) put(lat);
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
