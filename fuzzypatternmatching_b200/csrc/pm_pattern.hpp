// pm_pattern.hpp — host-side reader of the reference's pattern directory
// `<pattern_dir>/<ps>/pattern_*` (format contract of
// /root/reference/src/run_pattern_matching_beta.cpp:433-441,473-475).
//
//   pattern_edge                 "s t" per line, both directions, sorted by s
//                                (include/havoqgt/graph.hpp:195-207, 224-270)
//   pattern_vertex_data          "id label"; labels are taken in file order (graph.hpp:181-193)
//   pattern_stat                 "diameter : <n>", key case-insensitive (graph.hpp:337-358)
//   pattern_nlc                  "P.. : I.. : C : valid_cycle : interleave : selected_vertices"
//                                (include/havoqgt/pattern_util.hpp:172-210)
//   pattern_non_local_constraint "<ignored> : e.. : agg.." one line per pattern_nlc line
//                                (pattern_util.hpp:254-278)
// pattern_vertex and pattern_edge_data are read by the reference but never used
// on this path, so they are not required here.
//
// Approximate matching (run_pattern_matching_beta_2.cpp:459-476, include/havoqgt/approximate_pattern_matching/
// pattern_graph.hpp:282-337, 604-622): a pattern_edge line may carry a third column, "s t flag" with flag 1 = mandatory
// and 0 = optional edge (lines without the column are mandatory, the format above), and
//   pattern_vertex_local_constraints   "v : min_optional_edge_count" per template vertex (-1 / 0: no requirement)
// gives vertex_min_optional_edge_count.  Either one switches the local constraint to the approximate form of
// approximate_pattern_matching/local_constraint_checking.hpp:625-651, 1062-1113.
#pragma once

#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace pm {

struct Constraint {
  std::vector<uint64_t> P;     // vertex labels along the walk
  std::vector<uint32_t> I;     // template vertex ids along the walk
  uint64_t C = 0;              // pattern_cycle_length (largest itr_count)
  bool valid_cycle = false;    // the walk must end at its source
  bool interleave = false;     // run LCC after this constraint removed a source
  bool selected_vertices = false;
  std::vector<uint32_t> enum_idx;   // history rule per hop (TDS)
  std::vector<uint32_t> agg_steps;  // parsed like the reference, unused like the reference
};

struct Pattern {
  int n_vertices = 0;
  int n_edges = 0;   // lines of pattern_edge (= directed template edges)
  int diameter = 0;  // LCC supersteps per call
  std::vector<uint64_t> vertex_label;
  uint16_t N[16] = {0};      // template neighbours over MANDATORY edges (every edge, for an exact pattern)
  uint16_t No[16] = {0};     // template neighbours over OPTIONAL edges (approximate matching)
  int min_opt[16] = {0};     // vertex_min_optional_edge_count (values <= 0: no requirement)
  bool approximate = false;  // the pattern has an optional edge or a pattern_vertex_local_constraints file
  std::vector<Constraint> constraints;
};

namespace detail {
inline std::string strip(const std::string& s) {
  const char* ws = " \t\r\n";
  size_t b = s.find_first_not_of(ws);
  if (b == std::string::npos) return std::string();
  return s.substr(b, s.find_last_not_of(ws) - b + 1);
}
inline std::vector<std::string> fields(const std::string& line) {
  std::vector<std::string> out;
  std::stringstream ss(line);
  std::string f;
  while (std::getline(ss, f, ':')) out.push_back(strip(f));
  return out;
}
template <class T>
inline bool numbers(const std::string& s, std::vector<T>& out) {
  std::stringstream ss(s);
  std::string tok;
  while (ss >> tok) {
    if (tok.find_first_not_of("0123456789") != std::string::npos) return false;
    if (tok.size() > 19) return false;  // does not fit 64 bits: rejected like any other malformed token
    out.push_back((T)std::strtoull(tok.c_str(), nullptr, 10));
  }
  return true;
}
}  // namespace detail

// Returns an empty string on success, otherwise the reason the directory was rejected.
inline std::string load_pattern_dir(const std::string& dir, Pattern& pat) {
  using namespace detail;
  pat = Pattern();
  const std::string base = dir + "/pattern";
  std::string line;
  {
    std::ifstream f(base + "_edge");
    if (!f) return "cannot open " + base + "_edge";
    long long prev = -1;
    while (std::getline(f, line)) {
      std::vector<uint64_t> st;
      if (strip(line).empty()) continue;
      if (!numbers(line, st) || st.size() < 2) return "pattern_edge: bad line '" + line + "'";
      if (st[0] > 15 || st[1] > 15)
        return "pattern_edge: template vertex ids must be < 16 (std::bitset<16>, beta.cpp:270-271)";
      if ((long long)st[0] < prev) return "pattern_edge: lines must be sorted by source (graph.hpp:224-270)";
      prev = (long long)st[0];
      if (st.size() >= 3 && st[2] == 0) {  // optional edge (approximate_pattern_matching/pattern_graph.hpp:320-337, 609-616)
        pat.No[st[0]] |= (uint16_t)(1u << st[1]);
        pat.approximate = true;
      } else {
        pat.N[st[0]] |= (uint16_t)(1u << st[1]);
      }
      pat.n_edges++;
    }
    if (prev < 0) return "pattern_edge is empty";
    pat.n_vertices = (int)prev + 1;
  }
  {
    std::ifstream f(base + "_vertex_data");
    if (!f) return "cannot open " + base + "_vertex_data";
    while (std::getline(f, line)) {
      std::vector<uint64_t> kv;
      if (strip(line).empty()) continue;
      if (!numbers(line, kv) || kv.size() < 2) return "pattern_vertex_data: bad line '" + line + "'";
      pat.vertex_label.push_back(kv[1]);
    }
    if (pat.vertex_label.empty() || pat.vertex_label.size() > 16)
      return "pattern_vertex_data: need 1..16 template vertices";
    if ((int)pat.vertex_label.size() != pat.n_vertices)
      return "pattern_vertex_data lists " + std::to_string(pat.vertex_label.size()) + " template vertices, pattern_edge " +
             std::to_string(pat.n_vertices);
  }
  {
    // vertex_min_optional_edge_count (approximate_pattern_matching/pattern_graph.hpp:282-315), optional file
    std::ifstream f(base + "_vertex_local_constraints");
    while (f && std::getline(f, line)) {
      if (strip(line).empty()) continue;
      auto kv = fields(line);
      std::vector<uint64_t> v;
      if (kv.size() < 2 || !numbers(kv[0], v) || v.size() != 1 || v[0] > 15)
        return "pattern_vertex_local_constraints: expected '<template vertex> : <min optional edge count>'";
      const std::string cnt = strip(kv[1]);
      char* end = nullptr;
      const long k = std::strtol(cnt.c_str(), &end, 10);
      if (cnt.empty() || *end != 0) return "pattern_vertex_local_constraints: bad count '" + cnt + "'";
      pat.min_opt[v[0]] = (int)std::max<long>(k, 0);
      pat.approximate = true;
    }
  }
  {
    std::ifstream f(base + "_stat");
    if (!f) return "cannot open " + base + "_stat";
    while (std::getline(f, line)) {
      auto kv = fields(line);
      if (kv.size() < 2) continue;
      std::string k = kv[0];
      std::transform(k.begin(), k.end(), k.begin(), [](unsigned char ch) { return (char)std::tolower(ch); });
      if (k == "diameter") {
        std::vector<uint64_t> d;
        if (!numbers(kv[1], d) || d.size() != 1) return "pattern_stat: bad diameter";
        pat.diameter = (int)d[0];
      }
    }
    if (pat.diameter <= 0) return "pattern_stat: diameter missing or zero";
  }
  {
    std::ifstream f(base + "_nlc");
    while (f && std::getline(f, line)) {
      if (strip(line).empty()) continue;
      auto kv = fields(line);
      if (kv.size() < 6) return "pattern_nlc: expected 6 ':'-separated fields";
      Constraint c;
      std::vector<uint64_t> scal;
      if (!numbers(kv[0], c.P) || !numbers(kv[1], c.I)) return "pattern_nlc: bad walk";
      for (int i = 2; i < 6; ++i)
        if (!numbers(kv[i], scal) || (int)scal.size() != i - 1) return "pattern_nlc: bad scalar field";
      c.C = scal[0];
      c.valid_cycle = scal[1] != 0;
      c.interleave = scal[2] != 0;
      c.selected_vertices = scal[3] != 0;
      if (c.P.size() != c.I.size() || c.P.size() != c.C + 2)
        return "pattern_nlc: walk length must equal cycle_length + 2";
      if (c.P.size() > 16) return "pattern_nlc: walks longer than 16 vertices overflow visited_vertices (tds_batch_1.hpp:964)";
      for (uint32_t id : c.I)
        if (id > 15) return "pattern_nlc: template vertex ids must be < 16";
      if (c.selected_vertices) return "pattern_nlc: selected_vertices = 1 is not supported";
      pat.constraints.push_back(c);
    }
    std::ifstream g(base + "_non_local_constraint");
    size_t k = 0;
    while (g && std::getline(g, line)) {
      if (strip(line).empty()) continue;
      auto kv = fields(line);
      if (kv.size() < 3) return "pattern_non_local_constraint: expected 3 ':'-separated fields";
      if (k < pat.constraints.size()) {
        Constraint& c = pat.constraints[k];
        if (!numbers(kv[1], c.enum_idx) || !numbers(kv[2], c.agg_steps))
          return "pattern_non_local_constraint: bad indices";
      }
      ++k;
    }
  }
  return std::string();
}

}  // namespace pm
