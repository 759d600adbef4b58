// scan_files.cpp -- the wire formats on either side of the path, host only:
//   uos point files    ScanIO_uos -> readASCII / handle_line (reference src/scanio/helper.cc:577-835): one point
//                      per line, three blank-separated values, '#' comments, up to 10 unparsable lines tolerated
//                      at the top of the file, \n or \r\n
//   .pose files        six values, position then Euler angles in DEGREES (src/scanio/helper.cc:228-232)
//   .frames files      BasicScan::saveFrames (src/slam6d/basicScan.cc:902-917): per frame 16 doubles written with
//                      the default ostream format, each followed by a blank, then the AlgoType integer
//                      (operator<< include/slam6d/globals.icc:123-132; enum include/slam6d/scan.h:126)
// plus the frame bookkeeping of Scan::transform (src/slam6d/scan.cc:941-1000) that decides which scans receive a
// frame of which type.  The reader parses the file in parallel chunks (std::from_chars), because text parsing is
// what remains on the host once the correspondence search is gone.
#include "../../include/b200icp.h"

#include <algorithm>
#include <cerrno>
#include <charconv>
#include <clocale>
#include <locale.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <fstream>
#include <sstream>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

extern "C" int b200icp_set_error_(int code, const char* msg);

namespace {

inline bool is_blank(char c) { return c == ' ' || c == '\t'; }

// strtod semantics on one whole token (reference: strtoval -> strtod with an end-of-token check)
bool parse_token(const char* b, const char* e, double* out) {
  if (b == e) return false;
  const char* p = b;
  if (*p == '+') ++p;
  auto r = std::from_chars(p, e, *out);
  if (r.ec == std::errc() && r.ptr == e) return true;
  // out of range (1e400, 1e-400): the reference's strtoval rejects the token when strtod sets ERANGE
  // (helper.cc:242-271) -- "unable to parse line", not inf / 0
  if (r.ec == std::errc::result_out_of_range) return false;
  std::string tmp(b, e);       // rare spellings from_chars does not take (hex floats, "infinity")
  char* end = nullptr;
  errno = 0;
  static const locale_t c_loc = newlocale(LC_ALL_MASK, "C", (locale_t)0);   // the reference parses in the C locale
  const double v = c_loc ? strtod_l(tmp.c_str(), &end, c_loc) : strtod(tmp.c_str(), &end);
  if (errno == ERANGE) return false;
  if (end != tmp.c_str() + tmp.size()) return false;
  *out = v;
  return true;
}

// Clinger's fast path for plain decimals: at most 15 digits and a decimal exponent within +-22 make both the integer
// mantissa and the power of ten exact doubles, so ONE multiplication or division is correctly rounded -- the same
// value strtod / from_chars return, at a quarter of their cost.  Everything else (longer mantissas, "inf", hex,
// malformed tokens) returns false with p untouched and goes through from_chars / strtod.
inline bool fast_decimal(const char*& p, const char* e, double& out) {
  static const double P10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  const char* s = p;
  bool neg = false;
  if (s < e && *s == '-') { neg = true; ++s; }
  unsigned long long m = 0;
  int nd = 0, exp10 = 0;
  while (s < e && (unsigned)(*s - '0') <= 9u) { m = m * 10 + (unsigned)(*s - '0'); ++nd; ++s; if (nd > 15) return false; }
  if (s < e && *s == '.') {
    ++s;
    while (s < e && (unsigned)(*s - '0') <= 9u) { m = m * 10 + (unsigned)(*s - '0'); ++nd; --exp10; ++s; if (nd > 15) return false; }
  }
  if (nd == 0) return false;
  if (s < e && (*s == 'e' || *s == 'E')) {
    const char* t = s + 1;
    bool eneg = false;
    if (t < e && (*t == '-' || *t == '+')) { eneg = *t == '-'; ++t; }
    if (!(t < e && (unsigned)(*t - '0') <= 9u)) return false;
    int ex = 0;
    while (t < e && (unsigned)(*t - '0') <= 9u) { ex = ex * 10 + (*t - '0'); ++t; if (ex > 400) return false; }
    exp10 += eneg ? -ex : ex;
    s = t;
  }
  if (exp10 < -22 || exp10 > 22) return false;
  double d = (double)m;
  d = exp10 < 0 ? d / P10[-exp10] : d * P10[exp10];
  out = neg ? -d : d;
  p = s;
  return true;
}

// one line of a uos file: 0 = no point (empty / comment), 1 = point stored in v, -1 = parse error.
// Fast path: fast_decimal finds the end of each number itself; anything it does not take whole (long mantissas, a
// leading '+', "inf", hex floats, trailing garbage) goes through parse_token on the blank-delimited token.
int parse_line(const char* b, const char* e, double v[3]) {
  if (e > b && e[-1] == '\r') --e;
  while (b < e && is_blank(*b)) ++b;
  if (b == e || *b == '#') return 0;
  int nv = 0;
  const char* p = b;
  while (p < e && *p != '#') {
    if (nv == 3) return -1;                       // "too many values in line"
    const char* fp = p;
    if (fast_decimal(fp, e, v[nv]) && (fp == e || is_blank(*fp) || *fp == '#')) {
      p = fp;
    } else {
      const char* t = p;
      while (t < e && !is_blank(*t) && *t != '#') ++t;
      if (!parse_token(p, t, &v[nv])) return -1;
      p = t;
    }
    ++nv;
    while (p < e && is_blank(*p)) ++p;
  }
  return nv == 3 ? 1 : -1;                         // "less values than in spec"
}

struct Frame { double m[16]; int type; };

}  // namespace

struct b200icp_frames {
  std::vector<std::vector<Frame>> scans;
};

extern "C" {

void b200icp_free(void* p) { free(p); }

int b200icp_read_uos(const char* path, double** xyz_out, size_t* n_out) {
  if (!path || !xyz_out || !n_out) return b200icp_set_error_(B200ICP_EINVAL, "read_uos: NULL argument");
  *xyz_out = nullptr;
  *n_out = 0;
  // the file is mapped, not copied: the parser threads fault its pages in as they go
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return b200icp_set_error_(B200ICP_EINVAL, (std::string("read_uos: cannot open ") + path).c_str());
  struct stat sb;
  if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) {
    close(fd);
    return b200icp_set_error_(B200ICP_EINVAL, (std::string("read_uos: not a regular file: ") + path).c_str());
  }
  const size_t fsize = (size_t)sb.st_size;
  struct Mapping {
    void* p = nullptr; size_t n = 0;
    ~Mapping() { if (p && p != MAP_FAILED) munmap(p, n); }
  } map;
  if (fsize > 0) {
    map.p = mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0);
    map.n = fsize;
  }
  close(fd);
  if (fsize > 0 && map.p == MAP_FAILED) {
    map.p = nullptr;
    return b200icp_set_error_(B200ICP_EINVAL, "read_uos: mmap failed");
  }
  const char* base = fsize ? (const char*)map.p : "";
  const char* end = base + fsize;
  // chunk boundaries on line starts
  unsigned nthr = std::thread::hardware_concurrency();
  if (nthr == 0) nthr = 1;
  if (nthr > 32) nthr = 32;
  if (fsize < (1u << 20)) nthr = 1;
  std::vector<const char*> cut(nthr + 1);
  cut[0] = base;
  cut[nthr] = end;
  for (unsigned t = 1; t < nthr; ++t) {
    const char* p = base + fsize / nthr * t;
    while (p < end && *p != '\n') ++p;
    cut[t] = p < end ? p + 1 : end;
    if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
  }
  struct Part { std::vector<double> xyz; std::vector<std::pair<size_t, signed char>> events; size_t lines = 0; };
  // events: (line index inside the chunk, status) for every line that is not a plain point -- errors (-1) and
  // point-free lines (0); the header rule needs their order relative to the points
  std::vector<Part> parts(nthr);
  auto work = [&](unsigned t) {
    Part& P = parts[t];
    const char* p = cut[t];
    const char* e = cut[t + 1];
    P.xyz.reserve((size_t)(e - p) / 24 * 3);
    size_t line = 0;
    while (p < e) {
      const char* q = (const char*)memchr(p, '\n', (size_t)(e - p));
      const char* le = q ? q : e;
      double v[3];
      const int s = parse_line(p, le, v);
      if (s == 1) {
        if (P.xyz.empty()) P.events.push_back({line, 1});   // remember where the chunk's first point is
        P.xyz.insert(P.xyz.end(), v, v + 3);
      } else {
        P.events.push_back({line, (signed char)s});
      }
      ++line;
      p = q ? q + 1 : e;
    }
    P.lines = line;
  };
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nthr; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  // header rule of readASCII (helper.cc:752-822): unparsable lines are skipped while no line has been read
  // successfully yet (at most 10 of them); a successfully handled line -- a point, an empty line or a comment --
  // ends the header, after which any unparsable line is fatal
  int header = 10;
  size_t line_base = 0;
  for (unsigned t = 0; t < nthr; ++t) {
    for (auto& ev : parts[t].events) {
      if (ev.second == -1) {
        header -= 1;
        if (header < 0) {
          std::ostringstream m;
          m << "read_uos: unable to parse line " << (line_base + ev.first + 1) << " of " << path;
          return b200icp_set_error_(B200ICP_EINVAL, m.str().c_str());
        }
      } else if (header >= 0) {
        header = -1;
      }
    }
    line_base += parts[t].lines;
  }
  size_t total = 0;
  for (auto& P : parts) total += P.xyz.size();
  double* out = (double*)malloc((total ? total : 1) * sizeof(double));
  if (!out) return b200icp_set_error_(B200ICP_ENOMEM, "read_uos: out of memory");
  size_t off = 0;
  for (auto& P : parts) {
    if (!P.xyz.empty()) memcpy(out + off, P.xyz.data(), P.xyz.size() * sizeof(double));
    off += P.xyz.size();
  }
  *xyz_out = out;
  *n_out = total / 3;
  return B200ICP_OK;
}

int b200icp_write_uos(const char* path, const double* xyz, size_t n, double scale, int format) {
  // write_uos(DataXYZ&, FILE*, scaleFac, hexfloat, high_precision) (src/scanio/writer.cc:146-178), the output of
  // bin/scan_red: "%lf %lf %lf" (format 0), "%.016e" x3 (1, round-trips a double) or "%.013a" x3 (2, hex floats).
  // Lines are formatted in parallel chunks and written in order.
  if (!path || (!xyz && n) || format < 0 || format > 2) return b200icp_set_error_(B200ICP_EINVAL, "write_uos: bad argument");
  FILE* f = fopen(path, "wb");
  if (!f) return b200icp_set_error_(B200ICP_EINVAL, (std::string("write_uos: cannot open ") + path).c_str());
  static const char* const fmts[3] = {"%lf %lf %lf\n", "%.016e %.016e %.016e\n", "%.013a %.013a %.013a\n"};
  const char* fmt = fmts[format];
  unsigned nthr = std::thread::hardware_concurrency();
  if (nthr == 0) nthr = 1;
  if (nthr > 32) nthr = 32;
  if (n < 50000) nthr = 1;
  const size_t block = 1u << 16;                      // points per formatted piece
  bool ok = true;
  for (size_t base = 0; base < n && ok; base += block * nthr) {
    std::vector<std::string> piece(nthr);
    auto work = [&](unsigned t) {
      const size_t lo = base + (size_t)t * block, hi = std::min(n, lo + block);
      if (lo >= hi) return;
      std::string& out = piece[t];
      out.reserve((hi - lo) * 48);
      char line[256];
      for (size_t j = lo; j < hi; ++j) {
        const int len = snprintf(line, sizeof line, fmt, scale * xyz[3 * j], scale * xyz[3 * j + 1], scale * xyz[3 * j + 2]);
        if (len > 0) out.append(line, (size_t)std::min<int>(len, (int)sizeof line - 1));
      }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nthr; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    for (unsigned t = 0; t < nthr && ok; ++t)
      if (!piece[t].empty()) ok = fwrite(piece[t].data(), 1, piece[t].size(), f) == piece[t].size();
  }
  if (fclose(f) != 0) ok = false;
  return ok ? B200ICP_OK : b200icp_set_error_(B200ICP_EINVAL, "write_uos: write failed");
}

int b200icp_read_pose(const char* path, double rPos[3], double rPosTheta[3]) {
  if (!path || !rPos || !rPosTheta) return b200icp_set_error_(B200ICP_EINVAL, "read_pose: NULL argument");
  std::ifstream f(path);
  if (!f) return b200icp_set_error_(B200ICP_EINVAL, (std::string("read_pose: cannot open ") + path).c_str());
  double v[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 6; ++i) f >> v[i];                 // "read 6 plain doubles"
  for (int i = 0; i < 3; ++i) {
    rPos[i] = v[i];
    rPosTheta[i] = v[i + 3] * M_PI / 180.0;              // rad(), globals.icc:172-175
  }
  return B200ICP_OK;
}

b200icp_frames* b200icp_frames_create(int n_scans) {
  if (n_scans < 0) return nullptr;
  b200icp_frames* f = new b200icp_frames();
  f->scans.resize((size_t)n_scans);
  return f;
}

void b200icp_frames_destroy(b200icp_frames* f) { delete f; }

int b200icp_frames_add(b200icp_frames* f, int scan, const double transMat[16], int type) {
  if (!f || !transMat || scan < 0 || scan >= (int)f->scans.size())
    return b200icp_set_error_(B200ICP_EINVAL, "frames_add: bad argument");
  Frame fr;
  memcpy(fr.m, transMat, sizeof fr.m);
  fr.type = type;
  f->scans[(size_t)scan].push_back(fr);
  return B200ICP_OK;
}

int b200icp_frames_transform(b200icp_frames* f, int scan, const double* transmats, int type, int islum) {
  // the frame part of Scan::transform (scan.cc:941-1000) for a BasicScan `scan`; transmats = current transMat of
  // EVERY scan (16 doubles each, allScans order)
  if (!f || !transmats || scan < 0 || scan >= (int)f->scans.size())
    return b200icp_set_error_(B200ICP_EINVAL, "frames_transform: bad argument");
  if (type == B200ICP_FRAME_INVALID) return B200ICP_OK;
  const int n = (int)f->scans.size();
  int found = 0;
  switch (islum) {
    case -1: break;
    case 0:
      for (int i = 0; i < n; ++i) {
        if (i == scan) { found = i; b200icp_frames_add(f, i, transmats + 16 * i, type); }
        else b200icp_frames_add(f, i, transmats + 16 * i, found == 0 ? B200ICP_FRAME_ICPINACTIVE : B200ICP_FRAME_INVALID);
      }
      break;
    case 1: b200icp_frames_add(f, scan, transmats + 16 * scan, type); break;
    case 2:
      for (int i = 0; i < n; ++i) {
        if (i == scan) {
          found = i;
          b200icp_frames_add(f, i, transmats + 16 * i, type);
          b200icp_frames_add(f, 0, transmats, type);
          continue;
        }
        if (found != 0) b200icp_frames_add(f, i, transmats + 16 * i, B200ICP_FRAME_INVALID);
      }
      break;
    default: return b200icp_set_error_(B200ICP_EINVAL, "invalid point transformation mode");
  }
  return B200ICP_OK;
}

int b200icp_frames_count(const b200icp_frames* f, int scan) {
  if (!f || scan < 0 || scan >= (int)f->scans.size()) return -1;
  return (int)f->scans[(size_t)scan].size();
}

int b200icp_frames_get(const b200icp_frames* f, int scan, int k, double transMat[16], int* type) {
  if (!f || scan < 0 || scan >= (int)f->scans.size() || k < 0 || k >= (int)f->scans[(size_t)scan].size())
    return b200icp_set_error_(B200ICP_EINVAL, "frames_get: bad index");
  const Frame& fr = f->scans[(size_t)scan][(size_t)k];
  if (transMat) memcpy(transMat, fr.m, sizeof fr.m);
  if (type) *type = fr.type;
  return B200ICP_OK;
}

int b200icp_frames_load(b200icp_frames* f, int scan, const char* path) {
  // BasicScan::readFrames (basicScan.cc:872-900): the list is cleared first; empty lines and lines starting with
  // '#' are skipped; every other line must hold 16 numbers and an unsigned type
  if (!f || !path || scan < 0 || scan >= (int)f->scans.size())
    return b200icp_set_error_(B200ICP_EINVAL, "frames_load: bad argument");
  std::ifstream file(path);
  if (!file) return b200icp_set_error_(B200ICP_EINVAL, (std::string("frames_load: cannot open ") + path).c_str());
  std::vector<Frame> out;
  std::string line;
  while (std::getline(file, line)) {
    if (line.length() == 0) continue;
    if (line[0] == '#') continue;
    std::istringstream ls(line);
    Frame fr;
    bool ok = true;
    for (int i = 0; i < 16 && ok; ++i) ok = (bool)(ls >> fr.m[i]);
    unsigned int type = 0;
    if (ok) ok = (bool)(ls >> type);
    if (!ok) return b200icp_set_error_(B200ICP_EINVAL, (std::string("Malformed line in ") + path + ": " + line).c_str());
    fr.type = (int)type;
    out.push_back(fr);
  }
  f->scans[(size_t)scan].swap(out);
  return B200ICP_OK;
}

int b200icp_graph_read_net(const char* path, int* links, int cap, int* n_links, int* n_scans) {
  // Graph::Graph(const std::string& netfile) (graph.cc:52-74): "<nrScans> <nrLinks>" then nrLinks pairs "from to";
  // like there, the scan count reported is what Graph::addLink (graph.cc:157-174) counts -- ids not seen before
  if (!path || !n_links) return b200icp_set_error_(B200ICP_EINVAL, "graph_read_net: NULL argument");
  std::ifstream file(path);
  if (!file) return b200icp_set_error_(B200ICP_EINVAL, (std::string("graph_read_net: cannot open ") + path).c_str());
  int file_scans = 0, file_links = 0;
  file >> file_scans >> file_links;
  std::vector<int> from, to;
  int nr_scans = 0;
  for (int j = 0; j < file_links; ++j) {
    if (!file.good()) return b200icp_set_error_(B200ICP_EINVAL, "Error while reading network structure");
    int a = 0, b = 0;
    file >> a >> b;
    int present = 0;
    for (size_t k = 0; k < from.size(); ++k) present += (from[k] == a) + (to[k] == a);
    if (present == 0) ++nr_scans;
    present = 0;
    for (size_t k = 0; k < from.size(); ++k) present += (from[k] == b) + (to[k] == b);
    if (present == 0) ++nr_scans;
    from.push_back(a);
    to.push_back(b);
  }
  *n_links = (int)from.size();
  if (n_scans) *n_scans = nr_scans;
  if (links) {
    if ((int)from.size() > cap) return b200icp_set_error_(B200ICP_EINVAL, "graph_read_net: links array too small");
    for (size_t k = 0; k < from.size(); ++k) { links[2 * k] = from[k]; links[2 * k + 1] = to[k]; }
  }
  return B200ICP_OK;
}

int b200icp_frames_save(const b200icp_frames* f, int scan, const char* path, int append) {
  if (!f || !path || scan < 0 || scan >= (int)f->scans.size())
    return b200icp_set_error_(B200ICP_EINVAL, "frames_save: bad argument");
  std::ofstream file(path, append ? std::ios_base::app : std::ios_base::out);
  if (!file) return b200icp_set_error_(B200ICP_EINVAL, (std::string("frames_save: cannot open ") + path).c_str());
  for (const Frame& fr : f->scans[(size_t)scan]) {
    for (int i = 0; i < 16; ++i) {
      if (std::isnan(fr.m[i])) return b200icp_set_error_(B200ICP_EINVAL, "will not write out NAN value");
      file << fr.m[i] << " ";
    }
    file << fr.type << '\n';
  }
  file << std::flush;
  return file.good() ? B200ICP_OK : b200icp_set_error_(B200ICP_EINVAL, "frames_save: write failed");
}

}  // extern "C"
