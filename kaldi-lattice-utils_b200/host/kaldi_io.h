// kaldi_io.h -- the slice of Kaldi's command-line and table I/O conventions the
// seven hot-path tools depend on, re-implemented without Kaldi/OpenFst
// (Kaldi/OpenFst behaviour marked [ext] is restated from their
// public formats):
//   * ParseOptions: --name=value, --config=file, --help, --print-args, --verbose,
//     positional arguments, usage text; exit codes as the reference's main()s
//   * rspecifier / wspecifier: ark:file, ark:-, ark:cmd|, ark:|cmd, ark,t:..,
//     scp:file (sequential reading, "key path[:offset]" lines)
//   * CompactLattice holder: text (4-column CompactLattice and 5-column Lattice
//     entries, kwsbin2/egs/lattice.ark.txt / lattice.char.ark.txt) and binary
//     (OpenFst VectorFst stream, arc types compactlattice44 / lattice4)
//   * writers: BasicTupleVectorHolder (util/basic-tuple-vector-holder.h:149-181),
//     Posterior, Int32Vector, CompactLattice
#ifndef KLU_KALDI_IO_H_
#define KLU_KALDI_IO_H_

#include <stdint.h>
#include <string.h>
#include <stdio.h>

#include <cmath>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <streambuf>
#include <string>
#include <tuple>
#include <vector>

namespace kio {

struct KaldiError : public std::runtime_error {
  explicit KaldiError(const std::string& m) : std::runtime_error(m) {}
};
#define KIO_ERR(msg)                                   \
  do {                                                 \
    std::ostringstream os__;                           \
    os__ << "ERROR (" << kio::g_program << "): " << msg << "\n"; \
    throw kio::KaldiError(os__.str());                 \
  } while (0)
#define KIO_WARN(msg) (std::cerr << "WARNING (" << kio::g_program << "): " << msg << std::endl)
#define KIO_LOG(msg) (std::cerr << "LOG (" << kio::g_program << "): " << msg << std::endl)
#define KIO_VLOG(v, msg) \
  do {                   \
    if (kio::g_verbose >= (v)) std::cerr << "VLOG[" << (v) << "] (" << kio::g_program << "): " << msg << std::endl; \
  } while (0)

extern std::string g_program;
extern int g_verbose;

// ------------------------------------------------------------ ParseOptions ---
class ParseOptions {
 public:
  explicit ParseOptions(const char* usage) : usage_(usage) {}
  void Register(const std::string& name, float* p, const std::string& doc) { Add(name, kFloat, p, doc, Str(*p)); }
  void Register(const std::string& name, int32_t* p, const std::string& doc) { Add(name, kInt, p, doc, Str(*p)); }
  void Register(const std::string& name, bool* p, const std::string& doc) {
    Add(name, kBool, p, doc, *p ? "true" : "false");
  }
  void Register(const std::string& name, std::string* p, const std::string& doc) { Add(name, kString, p, doc, *p); }
  // Returns 0, or exits like Kaldi does for --help / bad options.
  void Read(int argc, const char* const* argv);
  int NumArgs() const { return (int)args_.size(); }
  std::string GetArg(int i) const;  // 1-based
  std::string GetOptArg(int i) const { return i <= NumArgs() ? GetArg(i) : std::string(); }
  void PrintUsage(bool print_command_line = false) const;

 private:
  enum Type { kFloat, kInt, kBool, kString };
  struct Opt {
    Type type;
    void* ptr;
    std::string doc, def;
  };
  template <typename T>
  static std::string Str(T v) {
    std::ostringstream os;
    os << v;
    return os.str();
  }
  void Add(const std::string& name, Type t, void* p, const std::string& doc, const std::string& def);
  static std::string Normalize(const std::string& name);
  bool SetOption(const std::string& key, const std::string& value, bool has_value);
  void ReadConfigFile(const std::string& path);
  std::string usage_;
  std::map<std::string, Opt> opts_;
  std::vector<std::string> args_;
  std::vector<std::string> argv_;
};

bool SplitStringToIntegers(const std::string& full, const char* delim, bool omit_empty, std::vector<int32_t>* out);

// ------------------------------------------------------------------ lattice ---
// A transition-id string (the "string" half of a CompactLatticeWeight): one per arc and
// per final state, almost always a handful of ids.  Up to six live inside the object; a
// std::vector here meant one heap allocation per arc when lattices are read to be written
// back (four times the parse time of everything else together).
class TidString {
 public:
  TidString() {}
  TidString(const TidString& o) { CopyFrom(o); }
  TidString(TidString&& o) noexcept { MoveFrom(&o); }
  TidString(const std::vector<int32_t>& v) { Assign(v.data(), v.size()); }
  TidString& operator=(const TidString& o) {
    if (this != &o) {
      Release();
      CopyFrom(o);
    }
    return *this;
  }
  TidString& operator=(TidString&& o) noexcept {
    if (this != &o) {
      Release();
      MoveFrom(&o);
    }
    return *this;
  }
  ~TidString() { Release(); }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  int32_t* data() { return cap_ > kInline ? heap_ : inl_; }
  const int32_t* data() const { return cap_ > kInline ? heap_ : inl_; }
  int32_t& operator[](size_t i) { return data()[i]; }
  int32_t operator[](size_t i) const { return data()[i]; }
  const int32_t* begin() const { return data(); }
  const int32_t* end() const { return data() + n_; }
  void clear() { n_ = 0; }
  void resize(size_t n) {
    Reserve(n);
    if (n > n_) memset(data() + n_, 0, 4 * (n - n_));
    n_ = (uint32_t)n;
  }
  void assign(size_t n, int32_t v) {
    Reserve(n);
    n_ = (uint32_t)n;
    for (size_t i = 0; i < n; ++i) data()[i] = v;
  }
  void push_back(int32_t v) {
    if (n_ == cap_) Reserve(2 * (size_t)cap_);
    data()[n_++] = v;
  }
  void swap(TidString& o) {
    TidString t(std::move(o));
    o = std::move(*this);
    *this = std::move(t);
  }
  bool operator==(const TidString& o) const { return n_ == o.n_ && memcmp(data(), o.data(), 4 * (size_t)n_) == 0; }

 private:
  static constexpr uint32_t kInline = 6;
  void Reserve(size_t n) {
    if (n <= cap_) return;
    int32_t* nb = new int32_t[n];
    memcpy(nb, data(), 4 * (size_t)n_);
    if (cap_ > kInline) delete[] heap_;
    heap_ = nb;
    cap_ = (uint32_t)n;
  }
  void Release() {
    if (cap_ > kInline) delete[] heap_;
    cap_ = kInline;
    n_ = 0;
  }
  void Assign(const int32_t* p, size_t n) {
    Reserve(n);
    if (n) memcpy(data(), p, 4 * n);
    n_ = (uint32_t)n;
  }
  void CopyFrom(const TidString& o) { Assign(o.data(), o.n_); }
  void MoveFrom(TidString* o) {
    if (o->cap_ > kInline) {
      heap_ = o->heap_;
      cap_ = o->cap_;
      n_ = o->n_;
      o->cap_ = kInline;
      o->n_ = 0;
    } else {
      memcpy(inl_, o->inl_, sizeof(inl_));
      n_ = o->n_;
      o->n_ = 0;
    }
  }
  uint32_t n_ = 0, cap_ = kInline;
  union {
    int32_t inl_[kInline];
    int32_t* heap_;
  };
};

struct CompactLat {
  std::string key;
  int32_t nstates = 0;
  // arcs grouped by ascending src (stored order inside a state)
  std::vector<int32_t> src, dst, label, dur;
  std::vector<float> graph, acoustic;
  std::vector<TidString> tids;  // transition-id strings (kept for lattice output)
  std::vector<float> fin_graph, fin_acoustic;  // +inf = not final
  std::vector<int32_t> fin_dur;
  std::vector<TidString> fin_tids;
};

// Reads one table entry body (after "key ") from `is`: text or binary.
// keep_tids = false drops the transition-id strings (only their lengths, the arc
// durations, are needed unless the lattice is written back).
void ReadCompactLattice(std::istream& is, CompactLat* lat, bool keep_tids = true);
void WriteCompactLattice(std::ostream& os, bool binary, const CompactLat& lat);
// TopSortCompactLatticeIfNeeded [ext]; throws on cycles.
void TopSortIfNeeded(CompactLat* lat);

// -------------------------------------------------------------------- tables ---
struct Specifier {
  bool is_scp = false, text = false, is_ark = false;
  std::string ark, scp;  // file names / "-" / "cmd|" / "|cmd"
};
Specifier ParseSpecifier(const std::string& spec, bool writing);

// Read-only stream buffer over a block of memory (a memory-mapped archive, or the
// collected output of an input pipe).  Table entries can be parsed straight from
// cur() .. end() and the position moved with advance(); the std::istream layered on
// top sees the same position.
class MemBuf : public std::streambuf {
 public:
  MemBuf(const char* begin, size_t n) {
    char* b = const_cast<char*>(begin);
    setg(b, b, b + n);
  }
  const char* cur() const { return gptr(); }
  const char* end() const { return egptr(); }
  void advance(size_t n) { setg(eback(), gptr() + n, egptr()); }

 protected:
  pos_type seekoff(off_type off, std::ios_base::seekdir dir, std::ios_base::openmode) override {
    char* base = dir == std::ios_base::beg ? eback() : dir == std::ios_base::cur ? gptr() : egptr();
    char* p = base + off;
    if (p < eback() || p > egptr()) return pos_type(off_type(-1));
    setg(eback(), p, egptr());
    return pos_type(p - eback());
  }
  pos_type seekpos(pos_type pos, std::ios_base::openmode m) override { return seekoff(off_type(pos), std::ios_base::beg, m); }
};

// Stream buffer over a stdio FILE (the two ends of popen): pipes are streamed through a
// 1 MB buffer, never collected in memory.  A failed fwrite is remembered (failed()).
class StdioBuf : public std::streambuf {
 public:
  StdioBuf(FILE* f, bool writing);
  ~StdioBuf() override { sync(); }
  bool failed() const { return failed_; }

 protected:
  int_type underflow() override;
  int_type overflow(int_type c) override;
  std::streamsize xsputn(const char* s, std::streamsize n) override;
  int sync() override;

 private:
  bool FlushOut();
  FILE* f_;
  bool writing_, failed_ = false;
  std::vector<char> buf_;
};

class Input {  // file, stdin or pipe opened for reading (binary-safe)
 public:
  explicit Input(const std::string& name);
  ~Input();
  std::istream& Stream() { return *is_; }
  // Ends a pipe and raises a KaldiError when its command failed (Kaldi does the same when
  // it closes a pipe input); no-op for files.  The destructor closes quietly.
  void Close();

 private:
  std::string filebuf_;  // stream buffer of a file input (declared first: destroyed after owned_)
  std::unique_ptr<MemBuf> membuf_;
  void* map_ = nullptr;  // memory-mapped regular file
  size_t map_len_ = 0;
  std::unique_ptr<std::istream> owned_;
  std::istream* is_ = nullptr;
  FILE* pipe_ = nullptr;
  std::string pipe_cmd_;
  std::unique_ptr<StdioBuf> pipe_sb_;
};

class Output {
 public:
  explicit Output(const std::string& name);
  ~Output();
  std::ostream& Stream() { return *os_; }
  // Flushes; raises a KaldiError when a write failed (full disk, closed pipe) or the
  // consumer command of a pipe exited non-zero.  The destructor closes quietly.
  void Close();
  // True while every write so far has reached the stream.
  bool Good();

 private:
  std::unique_ptr<std::ostream> owned_;
  std::ostream* os_ = nullptr;
  FILE* pipe_ = nullptr;
  std::string name_;
  std::unique_ptr<StdioBuf> pipe_sb_;
  bool closed_ = false;
};

// fn(i) for every i in [0, n) on up to `threads` threads (0 = $KLU_IO_THREADS, else the
// hardware's, at most 16); the first exception thrown by any of them is rethrown.
void ParallelFor(size_t n, const std::function<void(size_t)>& fn, int threads = 0);

class SequentialCompactLatticeReader {
 public:
  explicit SequentialCompactLatticeReader(const std::string& rspecifier, bool keep_tids = true);
  bool Done() const { return done_; }
  void Next();
  const std::string& Key() const { return cur_.key; }
  CompactLat& Value() { return cur_; }
  // Block mode, for archives held in memory (memory-mapped files, collected pipes): moves
  // the current entry and the following ones into `out` until they hold `max_arcs` arcs
  // (the entry that crosses the mark included, as a sequential reader filling a batch
  // would) or number `max_lattices`, and leaves the reader on the entry after them.  The binary entries are located
  // by walking their length fields and parsed on several threads.  False, with nothing
  // consumed, when the archive is not a memory block (or Done()).
  bool ReadBlock(int64_t max_arcs, std::vector<CompactLat>* out, size_t max_lattices = (size_t)-1);

 private:
  void ReadOne();
  bool keep_tids_ = true;
  Specifier spec_;
  std::unique_ptr<Input> in_;       // ark stream, or the scp list
  std::unique_ptr<Input> scp_item_;
  CompactLat cur_;
  bool done_ = false;
};

class TableWriter {  // archive writer: "key " + payload
 public:
  explicit TableWriter(const std::string& wspecifier);
  bool binary() const { return !spec_.text; }
  std::ostream& Begin(const std::string& key);  // writes the key, returns the stream
  // End of one entry; a failed write raises here (the reference's TableWriter::Write
  // returns os.good() and Kaldi raises on false).
  void End();
  void Close() { out_->Close(); }
  bool IsOpen() const { return out_ != nullptr; }

 private:
  Specifier spec_;
  std::unique_ptr<Output> out_;
};

// Kaldi basic types
void WriteKaldiFloat(std::ostream& os, double v);          // text: precision-7 general format
std::string FormatSelfTest(size_t n);  // text formatters vs iostream / printf; "" = identical
void WriteBasicInt32(std::ostream& os, bool binary, int32_t v);
void WriteBasicFloat(std::ostream& os, bool binary, float v);
void WriteBasicDouble(std::ostream& os, bool binary, double v);
void WriteToken(std::ostream& os, bool binary, const std::string& tok);

}  // namespace kio

#endif  // KLU_KALDI_IO_H_
