// kaldi_io.cc -- see kaldi_io.h.
#include "kaldi_io.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string.h>

#include <algorithm>
#include <charconv>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <exception>
#include <mutex>
#include <thread>
#include <cstdlib>

namespace kio {

std::string g_program = "klu";
int g_verbose = 0;

// ------------------------------------------------------------ ParseOptions ---
std::string ParseOptions::Normalize(const std::string& name) {
  std::string out;
  for (char ch : name) out += (ch == '_') ? '-' : (char)tolower(ch);
  return out;
}

void ParseOptions::Add(const std::string& name, Type t, void* p, const std::string& doc, const std::string& def) {
  opts_[Normalize(name)] = Opt{t, p, doc, def};
}

static bool ToBool(const std::string& v, bool* out) {
  std::string s;
  for (char ch : v) s += (char)tolower(ch);
  if (s == "true" || s == "t" || s == "1" || s.empty()) {
    *out = true;
    return true;
  }
  if (s == "false" || s == "f" || s == "0") {
    *out = false;
    return true;
  }
  return false;
}

bool ParseOptions::SetOption(const std::string& key_in, const std::string& value, bool has_value) {
  const std::string key = Normalize(key_in);
  if (key == "help") {
    PrintUsage();
    exit(0);
  }
  if (key == "verbose") {
    g_verbose = atoi(value.c_str());
    return true;
  }
  if (key == "print-args") return true;
  if (key == "config") {
    ReadConfigFile(value);
    return true;
  }
  auto it = opts_.find(key);
  if (it == opts_.end()) return false;
  Opt& o = it->second;
  char* end = nullptr;
  switch (o.type) {
    case kFloat: {
      if (!has_value || value.empty()) KIO_ERR("Invalid floating-point option \"" << value << "\" for --" << key);
      const std::string low = Normalize(value);
      float f;
      if (low == "inf" || low == "+inf" || low == "infinity") f = std::numeric_limits<float>::infinity();
      else if (low == "-inf" || low == "-infinity") f = -std::numeric_limits<float>::infinity();
      else {
        f = strtof(value.c_str(), &end);
        if (end == value.c_str() || *end != '\0') KIO_ERR("Invalid floating-point option \"" << value << "\"");
      }
      *static_cast<float*>(o.ptr) = f;
      break;
    }
    case kInt: {
      if (!has_value || value.empty()) KIO_ERR("Invalid integer option \"" << value << "\" for --" << key);
      const long long v = strtoll(value.c_str(), &end, 10);
      if (end == value.c_str() || *end != '\0') KIO_ERR("Invalid integer option \"" << value << "\"");
      *static_cast<int32_t*>(o.ptr) = (int32_t)v;
      break;
    }
    case kBool: {
      bool b;
      if (!ToBool(has_value ? value : "", &b)) KIO_ERR("Invalid format for boolean argument [expected true or false]: " << value);
      *static_cast<bool*>(o.ptr) = b;
      break;
    }
    case kString:
      *static_cast<std::string*>(o.ptr) = value;
      break;
  }
  return true;
}

void ParseOptions::ReadConfigFile(const std::string& path) {
  std::ifstream f(path);
  if (!f) KIO_ERR("Cannot open config file: " << path);
  std::string line;
  while (std::getline(f, line)) {
    const size_t hash = line.find('#');
    if (hash != std::string::npos) line = line.substr(0, hash);
    size_t a = line.find_first_not_of(" \t\r"), b = line.find_last_not_of(" \t\r");
    if (a == std::string::npos) continue;
    line = line.substr(a, b - a + 1);
    if (line.compare(0, 2, "--") != 0) KIO_ERR("Reading config file " << path << ": line does not start with --: " << line);
    const size_t eq = line.find('=');
    const std::string key = line.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
    const std::string val = eq == std::string::npos ? "" : line.substr(eq + 1);
    if (!SetOption(key, val, eq != std::string::npos)) KIO_ERR("Invalid option " << line << " in config file " << path);
  }
}

void ParseOptions::Read(int argc, const char* const* argv) {
  if (argc > 0) {
    const char* slash = strrchr(argv[0], '/');
    g_program = slash ? slash + 1 : argv[0];
  }
  for (int i = 0; i < argc; ++i) argv_.push_back(argv[i]);
  bool print_args = true;
  int i = 1;
  for (; i < argc; ++i) {
    const std::string a = argv[i];
    if (a.compare(0, 2, "--") != 0) break;
    if (a == "--") {
      ++i;
      break;
    }
    const size_t eq = a.find('=');
    const std::string key = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
    const std::string val = eq == std::string::npos ? "" : a.substr(eq + 1);
    if (Normalize(key) == "print-args") {
      bool b = true;
      ToBool(val, &b);
      print_args = b;
      continue;
    }
    if (!SetOption(key, val, eq != std::string::npos)) {
      PrintUsage(true);
      KIO_ERR("Invalid option " << a);
    }
  }
  for (; i < argc; ++i) args_.push_back(argv[i]);
  if (print_args) {
    std::ostringstream os;
    for (int k = 0; k < argc; ++k) os << argv[k] << (k + 1 < argc ? " " : "");
    std::cerr << os.str() << std::endl;
  }
}

std::string ParseOptions::GetArg(int i) const {
  if (i < 1 || i > (int)args_.size()) KIO_ERR("ParseOptions::GetArg, invalid index " << i);
  return args_[i - 1];
}

void ParseOptions::PrintUsage(bool print_command_line) const {
  std::cerr << "\n" << usage_ << "\n";
  std::cerr << "Options:\n";
  for (const auto& kv : opts_)
    std::cerr << "  --" << kv.first << " : " << kv.second.doc << " (default = " << kv.second.def << ")\n";
  std::cerr << "\nStandard options:\n  --config, --help, --print-args, --verbose\n\n";
  if (print_command_line) {
    std::ostringstream os;
    for (const auto& a : argv_) os << a << " ";
    std::cerr << "Command line was: " << os.str() << "\n";
  }
}

bool SplitStringToIntegers(const std::string& full, const char* delim, bool omit_empty, std::vector<int32_t>* out) {
  out->clear();
  size_t start = 0;
  while (start <= full.size()) {
    size_t end = full.find_first_of(delim, start);
    if (end == std::string::npos) end = full.size();
    const std::string tok = full.substr(start, end - start);
    if (!tok.empty() || !omit_empty) {
      char* e = nullptr;
      const long long v = strtoll(tok.c_str(), &e, 10);
      if (tok.empty() || e == tok.c_str() || *e != '\0') return false;
      out->push_back((int32_t)v);
    }
    start = end + 1;
  }
  return true;
}

// ------------------------------------------------------------- basic types ---
// Text fields are formatted with std::to_chars into a stack buffer and handed to the stream
// buffer directly (no sentry / locale machinery per field): byte-for-byte what
// `os << int` and printf("%.7g") give -- FormatSelfTest() below checks exactly that -- at
// a quarter of the cost, which is what a text index of 1e8 entries is made of.
namespace {
inline void PutChars(std::ostream& os, const char* b, size_t n) {
  if ((size_t)os.rdbuf()->sputn(b, (std::streamsize)n) != n) os.setstate(std::ios::badbit);
}
}  // namespace

namespace {
// precision-7 general format into buf (>= 41 bytes); returns the length
inline size_t FormatKaldiFloat(char* buf, double v) {
  if (std::isinf(v)) {
    if (v > 0) {
      memcpy(buf, "inf", 3);
      return 3;
    }
    memcpy(buf, "-inf", 4);
    return 4;
  }
  if (std::isnan(v)) {
    memcpy(buf, "nan", 3);
    return 3;
  }
  const auto r = std::to_chars(buf, buf + 40, v, std::chars_format::general, 7);  // == "%.7g"
  return (size_t)(r.ptr - buf);
}
}  // namespace

void WriteKaldiFloat(std::ostream& os, double v) {
  char buf[48];
  PutChars(os, buf, FormatKaldiFloat(buf, v));
}

void WriteBasicInt32(std::ostream& os, bool binary, int32_t v) {
  if (binary) {  // straight into the stream buffer: no sentry per field
    char b[5];
    b[0] = 4;
    memcpy(b + 1, &v, 4);
    if (os.rdbuf()->sputn(b, 5) != 5) os.setstate(std::ios::badbit);
  } else {
    char buf[16];
    const auto r = std::to_chars(buf, buf + 15, v);
    *r.ptr = ' ';
    PutChars(os, buf, (size_t)(r.ptr - buf) + 1);
  }
}

// The text formatters against the iostream / printf formulations they replace, over n
// pseudo-random values of every magnitude plus the special ones.  Empty string = identical.
std::string FormatSelfTest(size_t n) {
  auto slow_float = [](double v) {
    std::ostringstream o;
    if (std::isinf(v)) o << (v > 0 ? "inf" : "-inf");
    else if (std::isnan(v)) o << "nan";
    else {
      char buf[64];
      snprintf(buf, sizeof(buf), "%.7g", v);
      o << buf;
    }
    return o.str();
  };
  auto slow_int = [](int32_t v) {
    std::ostringstream o;
    o << v << " ";
    return o.str();
  };
  auto fast_float = [](double v) {
    std::ostringstream o, o2;
    WriteKaldiFloat(o, v);
    WriteBasicDouble(o2, false, v);
    if (o2.str() != o.str() + " ") return std::string("<WriteBasicDouble differs>");
    return o.str();
  };
  auto fast_int = [](int32_t v) {
    std::ostringstream o;
    WriteBasicInt32(o, false, v);
    return o.str();
  };
  std::vector<double> dv = {0.0, -0.0, 1.0, -1.0, 0.5, 1e-5, 9.9999995e-5, 1e-4, 123456.7, 1234567.0, 12345678.0,
                            9999999.5, 0.1, 1.0 / 3.0, 2.5e-310, 1.7976931348623157e308, 5e-324,
                            std::numeric_limits<double>::infinity(), -std::numeric_limits<double>::infinity(),
                            std::numeric_limits<double>::quiet_NaN(), -36.04365338911715, (double)1.609438f};
  std::vector<int32_t> iv = {0, 1, -1, 9, 10, 99, 100, 2147483647, -2147483647 - 1, 65535, -65536};
  uint64_t x = 0x9E3779B97F4A7C15ULL;
  auto next = [&x] {
    x ^= x << 13;
    x ^= x >> 7;
    x ^= x << 17;
    return x;
  };
  for (size_t i = 0; i < n; ++i) {
    const uint64_t r = next();
    double d;
    memcpy(&d, &r, 8);  // any bit pattern: every exponent, subnormals, NaNs
    dv.push_back(d);
    dv.push_back(-(double)(r % 1000000007ULL) * 1e-6);         // typical log-posteriors
    dv.push_back((double)(float)(-(double)(r >> 40) * 1e-5));   // float-valued ones
    iv.push_back((int32_t)(r >> 32));
    iv.push_back((int32_t)(r % 70000));
  }
  for (double d : dv)
    if (slow_float(d) != fast_float(d)) {
      std::ostringstream m;
      m.precision(17);
      m << "float " << d << ": '" << slow_float(d) << "' vs '" << fast_float(d) << "'";
      return m.str();
    }
  for (int32_t v : iv)
    if (slow_int(v) != fast_int(v)) return "int " + std::to_string(v);
  return "";
}

void WriteBasicFloat(std::ostream& os, bool binary, float v) {
  if (binary) {
    char b[5];
    b[0] = 4;
    memcpy(b + 1, &v, 4);
    if (os.rdbuf()->sputn(b, 5) != 5) os.setstate(std::ios::badbit);
  } else {
    char buf[48];
    size_t n = FormatKaldiFloat(buf, v);
    buf[n++] = ' ';
    PutChars(os, buf, n);
  }
}

void WriteBasicDouble(std::ostream& os, bool binary, double v) {
  if (binary) {
    char b[9];
    b[0] = 8;
    memcpy(b + 1, &v, 8);
    if (os.rdbuf()->sputn(b, 9) != 9) os.setstate(std::ios::badbit);
  } else {
    char buf[48];
    size_t n = FormatKaldiFloat(buf, v);
    buf[n++] = ' ';
    PutChars(os, buf, n);
  }
}

void WriteToken(std::ostream& os, bool, const std::string& tok) { os << tok << " "; }

// --------------------------------------------------------------- specifiers ---
Specifier ParseSpecifier(const std::string& spec, bool writing) {
  Specifier s;
  const size_t colon = spec.find(':');
  if (colon == std::string::npos) KIO_ERR("Invalid " << (writing ? "wspecifier " : "rspecifier ") << spec);
  const std::string head = spec.substr(0, colon), rest = spec.substr(colon + 1);
  size_t start = 0;
  bool both = false;
  std::vector<std::string> kinds;
  while (start <= head.size()) {
    size_t end = head.find(',', start);
    if (end == std::string::npos) end = head.size();
    const std::string t = head.substr(start, end - start);
    if (t == "ark") { s.is_ark = true; kinds.push_back(t); }
    else if (t == "scp") { s.is_scp = true; kinds.push_back(t); }
    else if (t == "t") s.text = true;
    else if (t == "b") s.text = false;
    else if (t == "s" || t == "cs" || t == "o" || t == "p" || t == "f" || t == "ns" || t == "ncs" || t == "no" ||
             t == "np" || t == "nf" || t == "bg") { /* sorted/once/permissive/flush hints: no effect here */ }
    else KIO_ERR("Invalid " << (writing ? "wspecifier " : "rspecifier ") << spec);
    start = end + 1;
  }
  both = s.is_ark && s.is_scp;
  if (!s.is_ark && !s.is_scp) KIO_ERR("Invalid specifier " << spec);
  if (both) {
    if (!writing) KIO_ERR("Invalid rspecifier " << spec);
    const size_t comma = rest.find(',');
    if (comma == std::string::npos) KIO_ERR("Invalid wspecifier " << spec);
    if (kinds[0] == "ark") { s.ark = rest.substr(0, comma); s.scp = rest.substr(comma + 1); }
    else { s.scp = rest.substr(0, comma); s.ark = rest.substr(comma + 1); }
  } else if (s.is_ark) {
    s.ark = rest;
  } else {
    s.scp = rest;
  }
  return s;
}

Input::Input(const std::string& name_in) {
  std::string name = name_in;
  while (!name.empty() && isspace((unsigned char)name.back())) name.pop_back();
  if (name == "-" || name.empty()) {
    is_ = &std::cin;
  } else if (name.back() == '|') {
    pipe_cmd_ = name.substr(0, name.size() - 1);
    pipe_ = popen(pipe_cmd_.c_str(), "r");
    if (!pipe_) KIO_ERR("Failed opening pipe for reading, command is: " << name);
    pipe_sb_.reset(new StdioBuf(pipe_, false));  // streamed: block mode is off for pipes
    owned_.reset(new std::istream(pipe_sb_.get()));
    is_ = owned_.get();
  } else {
    // "path:offset" (scp entries)
    size_t off = 0;
    const size_t colon = name.rfind(':');
    if (colon != std::string::npos && colon + 1 < name.size() &&
        name.find_first_not_of("0123456789", colon + 1) == std::string::npos) {
      off = (size_t)strtoull(name.c_str() + colon + 1, nullptr, 10);
      name = name.substr(0, colon);
    }
    // regular files are memory-mapped: entries are parsed in place
    {
      const int fd = open(name.c_str(), O_RDONLY);
      struct stat st;
      if (fd >= 0 && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0 && (size_t)st.st_size >= off) {
        void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m != MAP_FAILED) {
          madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
          map_ = m;
          map_len_ = (size_t)st.st_size;
        }
      }
      if (fd >= 0) close(fd);
      if (map_) {
        membuf_.reset(new MemBuf(static_cast<const char*>(map_), map_len_));
        membuf_->advance(off);
        owned_.reset(new std::istream(membuf_.get()));
        is_ = owned_.get();
        return;
      }
    }
    auto* f = new std::ifstream();
    filebuf_.resize(4 << 20);  // large reads: table entries are parsed straight off the stream buffer
    f->rdbuf()->pubsetbuf(&filebuf_[0], (std::streamsize)filebuf_.size());
    f->open(name, std::ios::in | std::ios::binary);
    owned_.reset(f);
    if (!*f) KIO_ERR("Error opening input stream " << name);
    if (off) f->seekg((std::streamoff)off);
    is_ = f;
  }
}

void Input::Close() {
  if (!pipe_) return;
  FILE* f = pipe_;
  pipe_ = nullptr;
  const int status = pclose(f);
  if (status != 0) KIO_ERR("Pipe " << pipe_cmd_ << "| had nonzero return status " << status);
}

Input::~Input() {
  if (pipe_) pclose(pipe_);
  pipe_ = nullptr;
  owned_.reset();
  membuf_.reset();
  if (map_) munmap(map_, map_len_);
}

StdioBuf::StdioBuf(FILE* f, bool writing) : f_(f), writing_(writing), buf_(1 << 20) {
  if (writing_) setp(buf_.data(), buf_.data() + buf_.size());
  else setg(buf_.data(), buf_.data(), buf_.data());
}

StdioBuf::int_type StdioBuf::underflow() {
  if (writing_) return traits_type::eof();
  if (gptr() < egptr()) return traits_type::to_int_type(*gptr());
  const size_t n = fread(buf_.data(), 1, buf_.size(), f_);
  if (n == 0) return traits_type::eof();
  setg(buf_.data(), buf_.data(), buf_.data() + n);
  return traits_type::to_int_type(*gptr());
}

bool StdioBuf::FlushOut() {
  const size_t n = (size_t)(pptr() - pbase());
  if (n && fwrite(pbase(), 1, n, f_) != n) failed_ = true;
  setp(buf_.data(), buf_.data() + buf_.size());
  return !failed_;
}

StdioBuf::int_type StdioBuf::overflow(int_type c) {
  if (!writing_ || !FlushOut()) return traits_type::eof();
  if (!traits_type::eq_int_type(c, traits_type::eof())) {
    *pptr() = traits_type::to_char_type(c);
    pbump(1);
  }
  return traits_type::not_eof(c);
}

std::streamsize StdioBuf::xsputn(const char* s, std::streamsize n) {
  if (!writing_) return 0;
  if (n <= epptr() - pptr()) {
    memcpy(pptr(), s, (size_t)n);
    pbump((int)n);
    return n;
  }
  if (!FlushOut()) return 0;
  if ((size_t)n >= buf_.size()) {  // large blocks go straight through
    if (fwrite(s, 1, (size_t)n, f_) != (size_t)n) {
      failed_ = true;
      return 0;
    }
    return n;
  }
  memcpy(pptr(), s, (size_t)n);
  pbump((int)n);
  return n;
}

int StdioBuf::sync() {
  if (!writing_) return 0;
  if (!FlushOut()) return -1;
  if (fflush(f_) != 0) failed_ = true;
  return failed_ ? -1 : 0;
}

Output::Output(const std::string& name) : name_(name) {
  if (name == "-" || name.empty()) {
    os_ = &std::cout;
  } else if (name[0] == '|') {
    pipe_ = popen(name.substr(1).c_str(), "w");
    if (!pipe_) KIO_ERR("Failed opening pipe for writing, command is: " << name);
    pipe_sb_.reset(new StdioBuf(pipe_, true));  // streamed to the command as it is produced
    owned_.reset(new std::ostream(pipe_sb_.get()));
    os_ = owned_.get();
  } else {
    auto* f = new std::ofstream(name, std::ios::out | std::ios::binary);
    owned_.reset(f);
    if (!*f) KIO_ERR("Error opening output stream " << name);
    os_ = f;
  }
}

bool Output::Good() { return os_ && os_->good() && !(pipe_sb_ && pipe_sb_->failed()); }

void Output::Close() {
  if (closed_) return;
  closed_ = true;
  if (os_) os_->flush();
  bool ok = Good();
  int status = 0;
  if (pipe_) {
    FILE* f = pipe_;
    pipe_ = nullptr;
    status = pclose(f);
  }
  if (!ok) KIO_ERR("Error writing to " << (name_.empty() ? std::string("standard output") : name_));
  if (status != 0) KIO_ERR("Pipe " << name_ << " had nonzero return status " << status);
}

Output::~Output() {
  try {
    Close();
  } catch (const std::exception&) {  // a destructor cannot raise; explicit Close() calls do
  }
}

// ---------------------------------------------------------------- lattices ---
namespace {

struct RawArc {
  int32_t src, dst, ilabel, olabel;
  float g, a;
  TidString tids;
};

void ParseWeight(const std::string& tok, bool compact, float* g, float* a, TidString* tids) {
  // "g,a" or "g,a,t1_t2_..."; empty fields mean 0
  *g = 0.0f;
  *a = 0.0f;
  tids->clear();
  std::vector<std::string> parts;
  size_t start = 0;
  while (start <= tok.size()) {
    size_t end = tok.find(',', start);
    if (end == std::string::npos) end = tok.size();
    parts.push_back(tok.substr(start, end - start));
    start = end + 1;
  }
  auto tof = [&](const std::string& s) -> float {
    if (s.empty()) return 0.0f;
    std::string low;
    for (char ch : s) low += (char)tolower(ch);
    if (low == "infinity" || low == "inf") return std::numeric_limits<float>::infinity();
    if (low == "-infinity" || low == "-inf") return -std::numeric_limits<float>::infinity();
    char* e = nullptr;
    const float f = strtof(s.c_str(), &e);
    if (e == s.c_str() || *e != '\0') KIO_ERR("Bad lattice weight: " << tok);
    return f;
  };
  if (parts.size() < 1 || parts.size() > 3) KIO_ERR("Bad lattice weight: " << tok);
  *g = tof(parts[0]);
  if (parts.size() > 1) *a = tof(parts[1]);
  if (compact && parts.size() > 2 && !parts[2].empty()) {
    size_t s2 = 0;
    const std::string& str = parts[2];
    while (s2 <= str.size()) {
      size_t e2 = str.find('_', s2);
      if (e2 == std::string::npos) e2 = str.size();
      tids->push_back(atoi(str.substr(s2, e2 - s2).c_str()));
      s2 = e2 + 1;
    }
  }
}

struct RawLat {
  std::vector<RawArc> arcs;  // any order
  std::map<int32_t, std::tuple<float, float, TidString> > finals;
  int32_t nstates = 0;
  int32_t start = -1;
  bool compact = true;
};

// [ext] ConvertLattice(Lattice -> CompactLattice): kaldi's Factor() collapses
// linear chains (one arc in, one arc out, not initial/final, no olabel on the
// outgoing arc) into one arc whose string is the chain's ilabels.
void FactorLattice(RawLat* lat) {
  const int32_t n = lat->nstates;
  std::vector<int32_t> nin(n, 0), nout(n, 0), only_out(n, -1);
  for (size_t i = 0; i < lat->arcs.size(); ++i) {
    const RawArc& a = lat->arcs[i];
    nin[a.dst]++;
    nout[a.src]++;
    only_out[a.src] = (int32_t)i;
  }
  std::vector<char> remove(n, 0);
  for (int32_t s = 0; s < n; ++s)
    remove[s] = nin[s] == 1 && nout[s] == 1 && s != lat->start && !lat->finals.count(s) &&
                lat->arcs[only_out[s]].olabel == 0;
  std::vector<RawArc> out;
  for (const RawArc& a0 : lat->arcs) {
    if (remove[a0.src]) continue;
    RawArc a = a0;
    a.tids.clear();
    if (a0.ilabel != 0) a.tids.push_back(a0.ilabel);
    while (remove[a.dst]) {
      const RawArc& nx = lat->arcs[only_out[a.dst]];
      if (nx.ilabel != 0) a.tids.push_back(nx.ilabel);
      a.g = a.g + nx.g;  // Times() of LatticeWeight = component-wise float add
      a.a = a.a + nx.a;
      a.dst = nx.dst;
    }
    out.push_back(a);
  }
  // renumber the surviving states, keeping their relative order
  std::vector<int32_t> newid(n, -1);
  int32_t m = 0;
  for (int32_t s = 0; s < n; ++s)
    if (!remove[s]) newid[s] = m++;
  for (RawArc& a : out) {
    a.src = newid[a.src];
    a.dst = newid[a.dst];
  }
  std::map<int32_t, std::tuple<float, float, TidString> > fin;
  for (auto& kv : lat->finals) fin[newid[kv.first]] = kv.second;
  lat->finals.swap(fin);
  lat->arcs.swap(out);
  lat->start = lat->start >= 0 ? newid[lat->start] : -1;
  lat->nstates = m;
}

void Finish(RawLat* raw, CompactLat* lat) {
  if (!raw->compact) FactorLattice(raw);
  // OpenFst requires the start state to be 0 for Kaldi's lattice functions; text
  // lattices start at the source of the first line.
  int32_t n = raw->nstates;
  if (raw->start > 0) {  // swap ids so the start is 0
    const int32_t st = raw->start;
    auto sw = [&](int32_t s) { return s == st ? 0 : (s == 0 ? st : s); };
    for (RawArc& a : raw->arcs) {
      a.src = sw(a.src);
      a.dst = sw(a.dst);
    }
    std::map<int32_t, std::tuple<float, float, TidString> > fin;
    for (auto& kv : raw->finals) fin[sw(kv.first)] = kv.second;
    raw->finals.swap(fin);
  }
  std::stable_sort(raw->arcs.begin(), raw->arcs.end(), [](const RawArc& x, const RawArc& y) { return x.src < y.src; });
  lat->nstates = n;
  const size_t na = raw->arcs.size();
  lat->src.resize(na);
  lat->dst.resize(na);
  lat->label.resize(na);
  lat->dur.resize(na);
  lat->graph.resize(na);
  lat->acoustic.resize(na);
  lat->tids.resize(na);
  for (size_t i = 0; i < na; ++i) {
    RawArc& a = raw->arcs[i];
    lat->src[i] = a.src;
    lat->dst[i] = a.dst;
    lat->label[i] = a.olabel;
    lat->dur[i] = (int32_t)a.tids.size();
    lat->graph[i] = a.g;
    lat->acoustic[i] = a.a;
    lat->tids[i].swap(a.tids);
  }
  const float inf = std::numeric_limits<float>::infinity();
  lat->fin_graph.assign(n, inf);
  lat->fin_acoustic.assign(n, inf);
  lat->fin_dur.assign(n, 0);
  lat->fin_tids.assign(n, TidString());
  for (auto& kv : raw->finals) {
    lat->fin_graph[kv.first] = std::get<0>(kv.second);
    lat->fin_acoustic[kv.first] = std::get<1>(kv.second);
    lat->fin_dur[kv.first] = (int32_t)std::get<2>(kv.second).size();
    lat->fin_tids[kv.first] = std::get<2>(kv.second);
  }
  TopSortIfNeeded(lat);
}

void ReadText(std::istream& is, CompactLat* lat) {
  // The text form starts with '\n' after the key and ends with an empty line.
  RawLat raw;
  std::string line;
  std::getline(is, line);  // rest of the key line
  bool decided = false;
  int32_t maxs = -1;
  while (std::getline(is, line)) {
    size_t a = line.find_first_not_of(" \t\r");
    if (a == std::string::npos) break;  // blank line terminates the entry
    std::istringstream ls(line);
    std::vector<std::string> tok;
    std::string t;
    while (ls >> t) tok.push_back(t);
    if (tok.size() <= 2) {  // final state
      const int32_t s = atoi(tok[0].c_str());
      if (s < 0) KIO_ERR("Lattice " << lat->key << ": negative state id in line: " << line);
      float g = 0, w = 0;
      TidString tids;
      if (tok.size() == 2) ParseWeight(tok[1], true, &g, &w, &tids);
      raw.finals[s] = std::make_tuple(g, w, tids);
      maxs = std::max(maxs, s);
      if (raw.start < 0) raw.start = s;
      continue;
    }
    RawArc arc;
    arc.src = atoi(tok[0].c_str());
    arc.dst = atoi(tok[1].c_str());
    if (arc.src < 0 || arc.dst < 0) KIO_ERR("Lattice " << lat->key << ": negative state id in line: " << line);
    bool is_compact;
    if (tok.size() == 3) is_compact = true;
    else if (tok.size() == 5) is_compact = false;
    else is_compact = tok[3].find(',') != std::string::npos;  // 4 columns: weight or olabel
    if (!decided) {
      raw.compact = is_compact;
      decided = true;
    } else if (raw.compact != is_compact) {
      KIO_ERR("Lattice " << lat->key << " mixes CompactLattice and Lattice text lines");
    }
    if (is_compact) {
      arc.ilabel = arc.olabel = atoi(tok[2].c_str());
      arc.g = arc.a = 0.0f;
      if (tok.size() == 4) ParseWeight(tok[3], true, &arc.g, &arc.a, &arc.tids);
    } else {
      arc.ilabel = atoi(tok[2].c_str());
      arc.olabel = atoi(tok[3].c_str());
      arc.g = arc.a = 0.0f;
      TidString dummy;
      if (tok.size() == 5) ParseWeight(tok[4], false, &arc.g, &arc.a, &dummy);
    }
    if (raw.start < 0) raw.start = arc.src;
    maxs = std::max(maxs, std::max(arc.src, arc.dst));
    raw.arcs.push_back(arc);
  }
  raw.nstates = maxs + 1;
  Finish(&raw, lat);
}

template <typename T>
T ReadRaw(std::istream& is) {
  T v;
  is.read(reinterpret_cast<char*>(&v), sizeof(T));
  if (!is) KIO_ERR("Unexpected end of binary lattice stream");
  return v;
}

std::string ReadFstString(std::istream& is) {
  const int32_t n = ReadRaw<int32_t>(is);
  if (n < 0 || n > (1 << 20)) KIO_ERR("Corrupt FST header");
  std::string s((size_t)n, '\0');
  if (n) is.read(&s[0], n);
  return s;
}

// Bulk reads straight from the stream buffer (no sentry / formatting layer per field).
inline void GetBytes(std::streambuf* sb, void* p, std::streamsize n) {
  if (n > 0 && sb->sgetn(reinterpret_cast<char*>(p), n) != n) KIO_ERR("Unexpected end of binary lattice stream");
}

// Walks the length fields of a binary CompactLattice body (nstates states from p) without
// reading anything else: the number of arcs and the end of the body.  OpenFst's VectorFst
// writer leaves the header's arc count at zero, so this walk is how the arcs are counted.
// False if the body is truncated or a length is negative.
bool CountBinaryCompactArcs(const char* p, const char* end, int64_t nstates, int64_t* narcs, const char** body_end) {
  int64_t total = 0;
  for (int64_t s = 0; s < nstates; ++s) {
    if (end - p < 12) return false;
    int32_t sz;
    memcpy(&sz, p + 8, 4);
    if (sz < 0 || (end - p - 12) / 4 < sz) return false;
    p += 12 + 4 * (size_t)sz;
    if (end - p < 8) return false;
    int64_t na;
    memcpy(&na, p, 8);
    p += 8;
    if (na < 0 || na > (end - p) / 24) return false;
    for (int64_t k = 0; k < na; ++k) {
      if (end - p < 24) return false;
      int32_t asz;
      memcpy(&asz, p + 16, 4);
      if (asz < 0 || (end - p - 24) / 4 < asz) return false;
      p += 24 + 4 * (size_t)asz;
    }
    total += na;
  }
  *narcs = total;
  *body_end = p;
  return true;
}

// Fast path for the common case -- a binary CompactLattice whose start state is 0:
// the OpenFst body lists the arcs state by state, i.e. already grouped by source, so
// the SoA arrays are filled directly (one read for the fixed part of every arc, one
// for its transition ids, which are kept only for the tool that writes lattices back).
// The same over a memory block: no copies except into the SoA arrays.
// narcs < 0: not counted yet.
bool ReadBinaryCompactMem(MemBuf* mb, int64_t nstates, int64_t narcs, bool keep_tids, CompactLat* lat) {
  if (narcs < 0) {
    const char* body_end;
    if (!CountBinaryCompactArcs(mb->cur(), mb->end(), nstates, &narcs, &body_end))
      KIO_ERR("Unexpected end of binary lattice " << lat->key << " (or corrupt length field)");
  }
  if (narcs >= ((int64_t)1 << 31)) KIO_ERR("Lattice " << lat->key << " has too many arcs");
  const char* p = mb->cur();
  const char* const end = mb->end();
  const int32_t n = (int32_t)nstates;
  const size_t na_total = (size_t)narcs;
  lat->nstates = n;
  const float inf = std::numeric_limits<float>::infinity();
  lat->fin_graph.assign(n, inf);
  lat->fin_acoustic.assign(n, inf);
  lat->fin_dur.assign(n, 0);
  if (keep_tids) lat->fin_tids.assign(n, TidString());
  lat->src.resize(na_total);
  lat->dst.resize(na_total);
  lat->label.resize(na_total);
  lat->dur.resize(na_total);
  lat->graph.resize(na_total);
  lat->acoustic.resize(na_total);
  if (keep_tids) lat->tids.resize(na_total);
  size_t e = 0;
  auto need = [&](size_t bytes) {
    if ((size_t)(end - p) < bytes) KIO_ERR("Unexpected end of binary lattice " << lat->key);
  };
  for (int32_t s = 0; s < n; ++s) {
    need(12);
    float g, a;
    int32_t sz;
    memcpy(&g, p, 4);
    memcpy(&a, p + 4, 4);
    memcpy(&sz, p + 8, 4);
    p += 12;
    if (sz < 0) KIO_ERR("Corrupt binary lattice " << lat->key);
    need(4 * (size_t)sz + 8);
    if (!(std::isinf(g) && std::isinf(a))) {
      lat->fin_graph[s] = g;
      lat->fin_acoustic[s] = a;
      lat->fin_dur[s] = sz;
      if (keep_tids) {
        lat->fin_tids[s].resize((size_t)sz);
        if (sz) memcpy(lat->fin_tids[s].data(), p, 4 * (size_t)sz);
      }
    }
    p += 4 * (size_t)sz;
    int64_t na;
    memcpy(&na, p, 8);
    p += 8;
    if (na < 0 || e + (size_t)na > na_total) KIO_ERR("Corrupt binary lattice " << lat->key << ": arc counts disagree");
    for (int64_t k = 0; k < na; ++k, ++e) {
      need(24);
      int32_t olabel, asz, dst;
      memcpy(&olabel, p + 4, 4);
      memcpy(&lat->graph[e], p + 8, 4);
      memcpy(&lat->acoustic[e], p + 12, 4);
      memcpy(&asz, p + 16, 4);
      p += 20;
      if (asz < 0) KIO_ERR("Corrupt binary lattice " << lat->key);
      need(4 * (size_t)asz + 4);
      if (keep_tids) {
        lat->tids[e].resize((size_t)asz);
        if (asz) memcpy(lat->tids[e].data(), p, 4 * (size_t)asz);
      }
      p += 4 * (size_t)asz;
      memcpy(&dst, p, 4);
      p += 4;
      lat->src[e] = s;
      lat->dst[e] = dst;
      lat->label[e] = olabel;
      lat->dur[e] = asz;
    }
  }
  if (e != na_total) KIO_ERR("Corrupt binary lattice " << lat->key << ": arc counts disagree");
  mb->advance((size_t)(p - mb->cur()));
  TopSortIfNeeded(lat);
  return true;
}

void ReadBinaryCompactFast(std::istream& is, int64_t nstates, int64_t narcs_hint, bool keep_tids, CompactLat* lat) {
  if (MemBuf* mb = dynamic_cast<MemBuf*>(is.rdbuf()))
    if (ReadBinaryCompactMem(mb, nstates, -1, keep_tids, lat)) return;  // the header's count is only a hint
  std::streambuf* sb = is.rdbuf();
  struct FinalHead { float g, a; int32_t sz; };
  struct ArcHead { int32_t ilabel, olabel; float g, a; int32_t sz; };
  static_assert(sizeof(FinalHead) == 12 && sizeof(ArcHead) == 20, "packed layout");
  const int32_t n = (int32_t)nstates;
  lat->nstates = n;
  const float inf = std::numeric_limits<float>::infinity();
  lat->fin_graph.assign(n, inf);
  lat->fin_acoustic.assign(n, inf);
  lat->fin_dur.assign(n, 0);
  if (keep_tids) lat->fin_tids.assign(n, TidString());
  if (narcs_hint > 0 && narcs_hint < ((int64_t)1 << 31)) {
    const size_t r = (size_t)narcs_hint;
    lat->src.reserve(r);
    lat->dst.reserve(r);
    lat->label.reserve(r);
    lat->dur.reserve(r);
    lat->graph.reserve(r);
    lat->acoustic.reserve(r);
    if (keep_tids) lat->tids.reserve(r);
  }
  std::vector<int32_t> scratch;
  for (int32_t s = 0; s < n; ++s) {
    FinalHead fh;
    GetBytes(sb, &fh, sizeof(fh));
    if (fh.sz < 0) KIO_ERR("Corrupt binary lattice " << lat->key);
    scratch.resize((size_t)fh.sz);
    GetBytes(sb, scratch.data(), 4 * (std::streamsize)fh.sz);
    if (!(std::isinf(fh.g) && std::isinf(fh.a))) {
      lat->fin_graph[s] = fh.g;
      lat->fin_acoustic[s] = fh.a;
      lat->fin_dur[s] = fh.sz;
      if (keep_tids) lat->fin_tids[s] = scratch;
    }
    int64_t na;
    GetBytes(sb, &na, 8);
    if (na < 0) KIO_ERR("Corrupt binary lattice " << lat->key);
    for (int64_t k = 0; k < na; ++k) {
      ArcHead ah;
      GetBytes(sb, &ah, sizeof(ah));
      if (ah.sz < 0) KIO_ERR("Corrupt binary lattice " << lat->key);
      scratch.resize((size_t)ah.sz);
      GetBytes(sb, scratch.data(), 4 * (std::streamsize)ah.sz);
      int32_t dst;
      GetBytes(sb, &dst, 4);
      lat->src.push_back(s);
      lat->dst.push_back(dst);
      lat->label.push_back(ah.olabel);
      lat->dur.push_back(ah.sz);
      lat->graph.push_back(ah.g);
      lat->acoustic.push_back(ah.a);
      if (keep_tids) lat->tids.push_back(scratch);
    }
  }
  if (!keep_tids) {
    lat->tids.clear();
    lat->fin_tids.clear();
  }
  // the stream position moved underneath the istream: nothing is buffered above the streambuf
  TopSortIfNeeded(lat);
}

void ReadBinary(std::istream& is, CompactLat* lat, bool keep_tids) {
  // [ext] OpenFst FstHeader + VectorFst body
  const int32_t magic = ReadRaw<int32_t>(is);
  if (magic != 2125659606) KIO_ERR("Reading lattice " << lat->key << ": bad FST magic number");
  const std::string fsttype = ReadFstString(is), arctype = ReadFstString(is);
  /*version*/ ReadRaw<int32_t>(is);
  const int32_t flags = ReadRaw<int32_t>(is);
  /*props*/ ReadRaw<uint64_t>(is);
  const int64_t start = ReadRaw<int64_t>(is), nstates = ReadRaw<int64_t>(is);
  const int64_t narcs = ReadRaw<int64_t>(is);
  if (fsttype != "vector") KIO_ERR("Unsupported FST type " << fsttype);
  if (flags & 3) KIO_ERR("Lattices with symbol tables are not supported");
  if (arctype == "compactlattice44" && start == 0 && nstates > 0 && nstates < ((int64_t)1 << 31)) {
    ReadBinaryCompactFast(is, nstates, narcs, keep_tids, lat);
    return;
  }
  RawLat raw;
  if (arctype == "compactlattice44") raw.compact = true;
  else if (arctype == "lattice4") raw.compact = false;
  else KIO_ERR("Unsupported arc type " << arctype << " (expected compactlattice44 or lattice4)");
  if (nstates < 0 || nstates >= ((int64_t)1 << 31) || start < -1 || start >= std::max<int64_t>(nstates, 1))
    KIO_ERR("Corrupt FST header of lattice " << lat->key << " (states " << nstates << ", start " << start << ")");
  raw.nstates = (int32_t)nstates;
  raw.start = (int32_t)start;
  auto read_weight = [&](float* g, float* a, TidString* tids) {
    *g = ReadRaw<float>(is);
    *a = ReadRaw<float>(is);
    tids->clear();
    if (raw.compact) {
      const int32_t sz = ReadRaw<int32_t>(is);
      tids->resize(sz);
      for (int32_t i = 0; i < sz; ++i) (*tids)[i] = ReadRaw<int32_t>(is);
    }
  };
  for (int64_t s = 0; s < nstates; ++s) {
    float g, a;
    TidString tids;
    read_weight(&g, &a, &tids);
    if (!(std::isinf(g) && std::isinf(a))) raw.finals[(int32_t)s] = std::make_tuple(g, a, tids);
    const int64_t na = ReadRaw<int64_t>(is);
    for (int64_t k = 0; k < na; ++k) {
      RawArc arc;
      arc.src = (int32_t)s;
      arc.ilabel = ReadRaw<int32_t>(is);
      arc.olabel = ReadRaw<int32_t>(is);
      read_weight(&arc.g, &arc.a, &arc.tids);
      arc.dst = ReadRaw<int32_t>(is);
      if (arc.dst < 0 || arc.dst >= nstates)
        KIO_ERR("Lattice " << lat->key << ": arc refers to state " << arc.dst << " outside [0, " << nstates << ")");
      raw.arcs.push_back(arc);
    }
  }
  Finish(&raw, lat);
}

}  // namespace

void ReadCompactLattice(std::istream& is, CompactLat* lat, bool keep_tids) {
  int c = is.peek();
  if (c == '\0') {  // tolerate a Kaldi binary marker "\0B" in front of the FST
    is.get();
    if (is.peek() == 'B') is.get();
    c = is.peek();
  }
  if (c == EOF) KIO_ERR("End of stream detected reading CompactLattice " << lat->key);
  if (isspace(c)) ReadText(is, lat);
  else if (c == 214) ReadBinary(is, lat, keep_tids);
  else KIO_ERR("Reading compact lattice " << lat->key << ": does not appear to be an FST");
}

// State ids straight from a file index vectors below (and in every tool): refuse anything
// outside [0, nstates) here, with the reference's error path, before they are used.
void ValidateStateIds(const CompactLat& lat) {
  const int32_t n = lat.nstates;
  const size_t na = lat.src.size();
  for (size_t i = 0; i < na; ++i) {
    const int32_t u = lat.src[i], v = lat.dst[i];
    if (u < 0 || u >= n || v < 0 || v >= n)
      KIO_ERR("Lattice " << lat.key << ": arc " << i << " refers to state " << (u < 0 || u >= n ? u : v)
                         << " outside [0, " << n << ")");
  }
}

void TopSortIfNeeded(CompactLat* lat) {
  ValidateStateIds(*lat);
  const size_t na = lat->src.size();
  // transition-id strings are only there when the lattice is read to be written back
  const bool have_tids = lat->tids.size() == na && na > 0;
  const bool have_fin_tids = lat->fin_tids.size() == (size_t)lat->nstates && lat->nstates > 0;
  bool sorted = true;
  for (size_t i = 0; i < na && sorted; ++i) sorted = lat->src[i] < lat->dst[i];
  if (sorted) return;
  const int32_t n = lat->nstates;
  std::vector<int32_t> first(n + 1, 0);
  for (size_t i = 0; i < na; ++i) first[lat->src[i] + 1]++;
  for (int32_t s = 0; s < n; ++s) first[s + 1] += first[s];
  // [ext] fst::TopSort: DFS from the start, then from every unvisited state in id
  // order; new order = reverse finishing order
  std::vector<char> color(n, 0);
  std::vector<int32_t> finish;
  std::vector<std::pair<int32_t, int32_t> > stack;
  for (int32_t pass = -1; pass < n; ++pass) {
    const int32_t root = pass < 0 ? 0 : pass;
    if (n == 0 || color[root]) continue;
    color[root] = 1;
    stack.push_back(std::make_pair(root, first[root]));
    while (!stack.empty()) {
      auto& top = stack.back();
      if (top.second < first[top.first + 1]) {
        const int32_t d = lat->dst[top.second++];
        if (color[d] == 1) KIO_ERR("Topological sorting of lattice " << lat->key << " failed (cyclic lattice)");
        if (color[d] == 0) {
          color[d] = 1;
          stack.push_back(std::make_pair(d, first[d]));
        }
      } else {
        color[top.first] = 2;
        finish.push_back(top.first);
        stack.pop_back();
      }
    }
  }
  std::vector<int32_t> order(n);
  for (int32_t i = 0; i < n; ++i) order[finish[n - 1 - i]] = i;
  CompactLat out;
  out.key = lat->key;
  out.nstates = n;
  std::vector<size_t> perm(na);
  for (size_t i = 0; i < na; ++i) perm[i] = i;
  std::stable_sort(perm.begin(), perm.end(), [&](size_t x, size_t y) { return order[lat->src[x]] < order[lat->src[y]]; });
  for (size_t p : perm) {
    out.src.push_back(order[lat->src[p]]);
    out.dst.push_back(order[lat->dst[p]]);
    out.label.push_back(lat->label[p]);
    out.dur.push_back(lat->dur[p]);
    out.graph.push_back(lat->graph[p]);
    out.acoustic.push_back(lat->acoustic[p]);
    if (have_tids) out.tids.push_back(lat->tids[p]);
  }
  out.fin_graph.resize(n);
  out.fin_acoustic.resize(n);
  out.fin_dur.resize(n);
  if (have_fin_tids) out.fin_tids.resize(n);
  for (int32_t s = 0; s < n; ++s) {
    out.fin_graph[order[s]] = lat->fin_graph[s];
    out.fin_acoustic[order[s]] = lat->fin_acoustic[s];
    out.fin_dur[order[s]] = lat->fin_dur[s];
    if (have_fin_tids) out.fin_tids[order[s]] = lat->fin_tids[s];
  }
  *lat = out;
}

namespace {
void WriteTextWeight(std::ostream& os, float g, float a, const TidString& tids) {
  WriteKaldiFloat(os, g);
  os << ",";
  WriteKaldiFloat(os, a);
  os << ",";
  for (size_t i = 0; i < tids.size(); ++i) os << (i ? "_" : "") << tids[i];
}
template <typename T>
void WriteRaw(std::ostream& os, T v) {
  os.write(reinterpret_cast<const char*>(&v), sizeof(T));
}
}  // namespace

void WriteCompactLattice(std::ostream& os, bool binary, const CompactLat& lat) {
  const size_t na = lat.src.size();
  const float inf = std::numeric_limits<float>::infinity();
  if (!binary) {
    // [ext] WriteCompactLattice text: newline after the key, FstPrinter acceptor
    // format with tabs, final states after a state's arcs, blank line at the end
    os << '\n';
    size_t e = 0;
    for (int32_t s = 0; s < lat.nstates; ++s) {
      for (; e < na && lat.src[e] == s; ++e) {
        os << s << '\t' << lat.dst[e] << '\t' << lat.label[e] << '\t';
        WriteTextWeight(os, lat.graph[e], lat.acoustic[e], lat.tids[e]);
        os << '\n';
      }
      if (!(lat.fin_graph[s] == inf && lat.fin_acoustic[s] == inf)) {
        os << s << '\t';
        WriteTextWeight(os, lat.fin_graph[s], lat.fin_acoustic[s], lat.fin_tids[s]);
        os << '\n';
      }
    }
    os << '\n';
    return;
  }
  // The entry is laid out in memory and written with one call ("\0B" = Kaldi's binary-mode
  // marker, CompactLatticeHolder::Write -> InitKaldiOutputStream [ext]; then OpenFst's
  // FstHeader and VectorFst body).
  size_t bytes = 2 + 4 + (4 + 6) + (4 + 16) + 4 + 4 + 8 + 3 * 8 + (size_t)lat.nstates * 20 + na * 24;
  for (int32_t s = 0; s < lat.nstates; ++s) bytes += 4 * lat.fin_tids[s].size();
  for (size_t e = 0; e < na; ++e) bytes += 4 * lat.tids[e].size();
  std::string buf(bytes, '\0');
  char* q = &buf[0];
  auto put = [&q](const void* v, size_t n) {
    memcpy(q, v, n);
    q += n;
  };
  auto put32 = [&put](int32_t v) { put(&v, 4); };
  auto put64 = [&put](int64_t v) { put(&v, 8); };
  auto putf = [&put](float v) { put(&v, 4); };
  put("\0B", 2);
  put32(2125659606);
  put32(6), put("vector", 6);
  put32(16), put("compactlattice44", 16);
  put32(2);                      // file version
  put32(0);                      // flags: no symbol tables
  put64(0x3LL);                  // kExpanded | kMutable, everything else unknown
  put64(lat.nstates ? 0 : -1);   // start state
  put64(lat.nstates);
  put64(0);                      // arc count: left at zero by OpenFst's VectorFst writer
  size_t e = 0;
  for (int32_t s = 0; s < lat.nstates; ++s) {
    putf(lat.fin_graph[s]);
    putf(lat.fin_acoustic[s]);
    put32((int32_t)lat.fin_tids[s].size());
    put(lat.fin_tids[s].data(), 4 * lat.fin_tids[s].size());
    size_t e1 = e;
    while (e1 < na && lat.src[e1] == s) ++e1;
    put64((int64_t)(e1 - e));
    for (; e < e1; ++e) {
      put32(lat.label[e]);
      put32(lat.label[e]);
      putf(lat.graph[e]);
      putf(lat.acoustic[e]);
      put32((int32_t)lat.tids[e].size());
      put(lat.tids[e].data(), 4 * lat.tids[e].size());
      put32(lat.dst[e]);
    }
  }
  os.write(buf.data(), (std::streamsize)(q - buf.data()));
}

// ------------------------------------------------------------------- tables ---
SequentialCompactLatticeReader::SequentialCompactLatticeReader(const std::string& rspecifier, bool keep_tids)
    : keep_tids_(keep_tids) {
  spec_ = ParseSpecifier(rspecifier, false);
  in_.reset(new Input(spec_.is_scp ? spec_.scp : spec_.ark));
  ReadOne();
}

void SequentialCompactLatticeReader::Next() { ReadOne(); }

void SequentialCompactLatticeReader::ReadOne() {
  std::istream& is = in_->Stream();
  cur_ = CompactLat();
  if (spec_.is_scp) {
    std::string line;
    while (std::getline(is, line)) {
      const size_t a = line.find_first_not_of(" \t\r");
      if (a == std::string::npos) continue;
      const size_t sp = line.find_first_of(" \t", a);
      if (sp == std::string::npos) KIO_ERR("Invalid scp line: " << line);
      cur_.key = line.substr(a, sp - a);
      std::string path = line.substr(line.find_first_not_of(" \t", sp));
      scp_item_.reset(new Input(path));
      ReadCompactLattice(scp_item_->Stream(), &cur_, keep_tids_);
      scp_item_->Close();  // "cmd|" entries: a failing command is an error
      return;
    }
    done_ = true;
    in_->Close();
    return;
  }
  // archive: skip whitespace, read the key token, one space, then the object
  int c;
  while ((c = is.peek()) != EOF && isspace(c)) is.get();
  if (c == EOF) {
    done_ = true;
    in_->Close();  // pipe rspecifiers: a failing producer command is an error, not an empty table
    return;
  }
  std::string key;
  while ((c = is.get()) != EOF && !isspace(c)) key += (char)c;
  if (c == EOF) KIO_ERR("Invalid archive file format: expected space after key " << key);
  cur_.key = key;
  if (c == '\n') is.unget();  // text lattices: the newline belongs to the holder
  ReadCompactLattice(is, &cur_, keep_tids_);
}

namespace {
int IoThreads() {
  const char* e = getenv("KLU_IO_THREADS");
  return std::max(1, e ? atoi(e) : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency())));
}
}  // namespace

void ParallelFor(size_t n, const std::function<void(size_t)>& fn, int threads) {
  if (threads <= 0) threads = IoThreads();
  threads = (int)std::min<size_t>((size_t)std::max(threads, 1), n);
  if (threads <= 1) {
    for (size_t i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<size_t> next(0);
  std::mutex mu;
  std::exception_ptr failure;
  auto work = [&] {
    try {
      for (size_t i; (i = next.fetch_add(1)) < n;) fn(i);
    } catch (...) {
      std::lock_guard<std::mutex> lk(mu);
      if (!failure) failure = std::current_exception();
      next.store(n);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < threads; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  if (failure) std::rethrow_exception(failure);
}

namespace {
// One archive entry located but not parsed: "key \0B" + FstHeader + body, a binary
// CompactLattice with start state 0 (what ReadBinary's fast path takes).
struct EntrySpan {
  std::string key;
  const char* body = nullptr;
  size_t len = 0;
  int64_t nstates = 0, narcs = 0;
};

// Locates the entry at the buffer's position and moves past it.  False -- nothing
// consumed -- at the end of the archive and for anything the fast path does not take
// (text entries, other FST types, a truncated body): the sequential reader then deals
// with the entry, error messages included.
bool ScanEntry(MemBuf* mb, EntrySpan* sp) {
  const char* p = mb->cur();
  const char* const end = mb->end();
  while (p < end && isspace((unsigned char)*p)) ++p;
  const char* k0 = p;
  while (p < end && !isspace((unsigned char)*p)) ++p;
  if (p == k0 || p == end || *p != ' ') return false;
  sp->key.assign(k0, p);
  ++p;
  if (end - p >= 2 && p[0] == '\0' && p[1] == 'B') p += 2;  // Kaldi's binary-mode marker (tolerated when absent)
  auto rd32 = [&](int32_t* v) {
    if (end - p < 4) return false;
    memcpy(v, p, 4);
    p += 4;
    return true;
  };
  auto rd64 = [&](int64_t* v) {
    if (end - p < 8) return false;
    memcpy(v, p, 8);
    p += 8;
    return true;
  };
  auto rdstr = [&](std::string* str) {
    int32_t n;
    if (!rd32(&n) || n < 0 || end - p < n) return false;
    str->assign(p, (size_t)n);
    p += n;
    return true;
  };
  int32_t magic, version, flags;
  int64_t props, start, nstates, hdr_arcs;
  std::string fsttype, arctype;
  if (!rd32(&magic) || magic != 2125659606 || !rdstr(&fsttype) || !rdstr(&arctype) || !rd32(&version) || !rd32(&flags) ||
      !rd64(&props) || !rd64(&start) || !rd64(&nstates) || !rd64(&hdr_arcs))
    return false;
  if (fsttype != "vector" || (flags & 3) || arctype != "compactlattice44" || start != 0 || nstates <= 0 ||
      nstates >= ((int64_t)1 << 31))
    return false;
  const char* body_end;
  int64_t narcs;
  if (!CountBinaryCompactArcs(p, end, nstates, &narcs, &body_end) || narcs >= ((int64_t)1 << 31)) return false;
  sp->body = p;
  sp->len = (size_t)(body_end - p);
  sp->nstates = nstates;
  sp->narcs = narcs;
  mb->advance((size_t)(body_end - mb->cur()));
  return true;
}
}  // namespace

bool SequentialCompactLatticeReader::ReadBlock(int64_t max_arcs, std::vector<CompactLat>* out, size_t max_lattices) {
  out->clear();
  if (done_ || spec_.is_scp) return false;
  MemBuf* mb = dynamic_cast<MemBuf*>(in_->Stream().rdbuf());
  if (!mb) return false;
  int64_t arcs = (int64_t)cur_.src.size();
  out->push_back(std::move(cur_));
  // This thread walks the archive from entry to entry; the others parse what it has
  // found so far (deques: elements stay put while more are appended).
  std::deque<EntrySpan> spans;
  std::deque<CompactLat> lats;
  std::mutex mu;
  std::condition_variable cv;
  size_t next = 0;
  bool finished = false;
  std::exception_ptr failure;
  const bool keep = keep_tids_;
  auto parse = [&] {
    for (;;) {
      EntrySpan* sp;
      CompactLat* lat;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return next < spans.size() || finished; });
        if (next >= spans.size()) return;
        sp = &spans[next];
        lat = &lats[next];
        ++next;
      }
      try {
        lat->key = sp->key;
        MemBuf body(sp->body, sp->len);
        ReadBinaryCompactMem(&body, sp->nstates, sp->narcs, keep, lat);
      } catch (...) {
        std::lock_guard<std::mutex> lk(mu);
        if (!failure) failure = std::current_exception();
      }
    }
  };
  std::vector<std::thread> helpers;
  for (int t = 1, n = IoThreads(); t < n; ++t) helpers.emplace_back(parse);
  size_t found = 1;  // the current entry
  while (arcs < max_arcs && found < max_lattices) {
    EntrySpan sp;
    if (!ScanEntry(mb, &sp)) break;
    ++found;
    arcs += sp.narcs;
    {
      std::lock_guard<std::mutex> lk(mu);
      spans.push_back(std::move(sp));
      lats.emplace_back();
    }
    cv.notify_one();
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    finished = true;
  }
  cv.notify_all();
  parse();
  for (auto& t : helpers) t.join();
  if (failure) std::rethrow_exception(failure);
  out->reserve(1 + lats.size());
  for (CompactLat& l : lats) out->push_back(std::move(l));
  ReadOne();  // the entry after the block (or the end of the archive)
  return true;
}

void TableWriter::End() {
  if (spec_.ark == "-") out_->Stream().flush();
  if (!out_->Good()) KIO_ERR("Write failure to " << (spec_.ark.empty() ? std::string("table") : spec_.ark));
}

TableWriter::TableWriter(const std::string& wspecifier) {
  if (wspecifier.empty()) return;
  spec_ = ParseSpecifier(wspecifier, true);
  if (spec_.is_scp && !spec_.is_ark) KIO_ERR("scp-only wspecifiers are not supported: " << wspecifier);
  if (spec_.is_scp) KIO_WARN("ark,scp wspecifier: writing the archive only");
  out_.reset(new Output(spec_.ark));
}

std::ostream& TableWriter::Begin(const std::string& key) {
  std::ostream& os = out_->Stream();
  os << key << ' ';
  return os;
}

}  // namespace kio
