// pm_io.hpp — host-side readers of the reference's text inputs (no device involved).
//
//   vertex metadata (-v)   include/havoqgt/vertex_data_db.hpp:139-262: `base` = <directory>/<filename prefix>; every
//                          regular file of the directory whose name matches "<prefix>.*" is read (:139-165), one
//                          "vertex label" pair per line (:177-186).  The reference spreads the files over its ranks
//                          and applies the pairs through visitors in no particular order; here files are read in
//                          name order and a later pair for the same vertex wins.
//   edge metadata (-e)     include/havoqgt/edge_data_db.hpp (same file discovery; "source target data" per line).
//                          The pattern matching path never reads the values (src/run_pattern_matching_beta.cpp:906
//                          binds them to an unused reference): the files are parsed and validated only.
//   edge lists             include/havoqgt/parallel_edge_list_reader.hpp:236-262: "source target [weight]" per line;
//                          with `undirected` every edge also yields its reverse (:126-150), which is what
//                          src/ingest_edge_list.cpp -u 1 feeds the graph constructor.
// Paths are relative to /root/reference.
#pragma once
#include <dirent.h>
#include <stdint.h>
#include <sys/stat.h>

#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace pm {
namespace io {

// files of dirname(base) whose name starts with basename(base), in name order (vertex_data_db.hpp:139-165)
inline bool files_with_prefix(const std::string& base, std::vector<std::string>& out, std::string& err) {
  out.clear();
  std::string dir = ".", prefix = base;
  const size_t slash = base.find_last_of('/');
  if (slash != std::string::npos) {
    dir = slash == 0 ? "/" : base.substr(0, slash);
    prefix = base.substr(slash + 1);
  }
  DIR* d = opendir(dir.c_str());
  if (!d) { err = "Error: Invalid directory path."; return false; }
  while (dirent* e = readdir(d)) {
    const std::string name = e->d_name;
    if (name.compare(0, prefix.size(), prefix) != 0) continue;
    const std::string path = dir + "/" + name;
    struct stat st;
    if (stat(path.c_str(), &st) != 0 || !S_ISREG(st.st_mode)) continue;
    out.push_back(path);
  }
  closedir(d);
  std::sort(out.begin(), out.end());
  if (out.empty()) { err = "Error: Failed to read input files."; return false; }
  return true;
}

inline bool parse_u64(const std::string& tok, uint64_t& v) {
  if (tok.empty() || tok.size() > 19 || tok.find_first_not_of("0123456789") != std::string::npos) return false;
  v = std::strtoull(tok.c_str(), nullptr, 10);
  return true;
}

// labels[v] for every "v label" pair found; vertices without a pair keep what `labels` holds (the caller zeroes it:
// VertexData is value-initialised in the reference, beta.cpp:344-349).  n_pairs: pairs applied.
inline bool read_vertex_data(const std::string& base, uint64_t n_vertices, uint64_t* labels, uint64_t* n_pairs,
                             std::string& err) {
  std::vector<std::string> files;
  if (!files_with_prefix(base, files, err)) return false;
  uint64_t n = 0;
  for (const std::string& path : files) {
    std::ifstream f(path);
    if (!f) { err = "cannot open " + path; return false; }
    std::string line, a, b;
    uint64_t ln = 0;
    while (std::getline(f, line)) {
      ++ln;
      std::istringstream ss(line);
      if (!(ss >> a)) continue;  // blank line
      uint64_t v, l;
      if (!(ss >> b) || !parse_u64(a, v) || !parse_u64(b, l)) {
        err = path + ":" + std::to_string(ln) + ": expected '<vertex> <label>'";
        return false;
      }
      if (v >= n_vertices) {
        err = path + ":" + std::to_string(ln) + ": vertex " + a + " is not in the graph (" + std::to_string(n_vertices) + " vertices)";
        return false;
      }
      labels[v] = l;
      ++n;
    }
  }
  if (n_pairs) *n_pairs = n;
  return true;
}

// validates "source target data" files; returns the number of records
inline bool check_edge_data(const std::string& base, uint64_t n_vertices, uint64_t* n_records, std::string& err) {
  std::vector<std::string> files;
  if (!files_with_prefix(base, files, err)) return false;
  uint64_t n = 0;
  for (const std::string& path : files) {
    std::ifstream f(path);
    if (!f) { err = "cannot open " + path; return false; }
    std::string line, a, b, c;
    uint64_t ln = 0;
    while (std::getline(f, line)) {
      ++ln;
      std::istringstream ss(line);
      if (!(ss >> a)) continue;
      uint64_t s, t, w;
      if (!(ss >> b >> c) || !parse_u64(a, s) || !parse_u64(b, t) || !parse_u64(c, w) || s >= n_vertices || t >= n_vertices) {
        err = path + ":" + std::to_string(ln) + ": expected '<source> <target> <data>' with vertices of the graph";
        return false;
      }
      ++n;
    }
  }
  if (n_records) *n_records = n;
  return true;
}

// directed slots of the listed edge-list files in file order; undirected: (s, t) then (t, s) per line, the order of
// the reference's iterator (parallel_edge_list_reader.hpp:126-150).  n_vertices = largest id + 1.
inline bool read_edge_lists(const std::vector<std::string>& files, bool undirected, std::vector<uint32_t>& src,
                            std::vector<uint32_t>& dst, uint64_t& n_vertices, std::string& err) {
  src.clear();
  dst.clear();
  uint64_t maxv = 0;
  bool any = false;
  for (const std::string& path : files) {
    std::ifstream f(path);
    if (!f) { err = "cannot open " + path; return false; }
    std::string line, a, b;
    uint64_t ln = 0;
    while (std::getline(f, line)) {
      ++ln;
      std::istringstream ss(line);
      if (!(ss >> a)) continue;
      if (a[0] == '#' || a[0] == '%') continue;  // comment lines of the usual edge-list dumps
      uint64_t s, t;
      if (!(ss >> b) || !parse_u64(a, s) || !parse_u64(b, t)) {
        err = path + ":" + std::to_string(ln) + ": expected '<source> <target> [weight]'";
        return false;
      }
      if (s >= (1ull << 31) || t >= (1ull << 31)) { err = path + ":" + std::to_string(ln) + ": vertex id above 2^31"; return false; }
      src.push_back((uint32_t)s);
      dst.push_back((uint32_t)t);
      if (undirected) {
        src.push_back((uint32_t)t);
        dst.push_back((uint32_t)s);
      }
      maxv = std::max(maxv, std::max(s, t));
      any = true;
    }
  }
  n_vertices = any ? maxv + 1 : 0;
  return true;
}

}  // namespace io
}  // namespace pm
