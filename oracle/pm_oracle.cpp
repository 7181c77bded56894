// pm_oracle.cpp — CPU ORACLE for the LCC/NLCC pruning path.
//
// TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs.  Never part of the product.
//
// PARITY STATUS: pinned by the reference's own code — its driver and visitor
// headers compile over the single-rank runtime stand-in of oracle/ref_shim
// (oracle/_ref); tests/test_oracle_vs_reference.py and tests/golden/reference_runs
// hold this oracle to that binary's result files.  See pm_oracle.h.
//
// What is restated (paths relative to /root/reference):
//   R-MAT stream      src/generate_rmat.cpp:197-205,
//                     include/havoqgt/rmat_edge_generator.hpp:126-139,218-259,
//                     include/havoqgt/detail/hash.hpp:65-143
//                     (boost::mt19937 == std::mt19937; the legacy
//                      boost::uniform_01<engine> of Boost 1.57 returns
//                      double(x) * 1/(max-min+1) = x * 2^-32)
//   degrees / labels  include/havoqgt/impl/delegate_partitioned_graph.ipp:437-470,1768-1781,
//                     include/havoqgt/vertex_data_db_degree.hpp:109
//   pattern files     include/havoqgt/graph.hpp:73-110,181-270,337-358,
//                     include/havoqgt/pattern_util.hpp:89-115,172-210,254-278
//   LCC               include/havoqgt/label_propagation_pattern_matching_nonunique_ee.hpp
//                     :148-459 (receiver), :467-636 (sender), :646-816 (per message),
//                     :827-1027 (post step), :1029-1153 (superstep loop)
//   NLCC nem_1        include/havoqgt/token_passing_pattern_matching_nonunique_nem_1.hpp:98-303,311-861
//   NLCC TDS          include/havoqgt/token_passing_pattern_matching_nonunique_tds_batch_1.hpp:122-335,347-919,976-1324
//   outer loop, rows  src/run_pattern_matching_beta.cpp:481-1425
//
// The reference executes visitors asynchronously; all state a visitor reads
// during an LCC superstep is written only in the post step (or is written with
// a value that does not depend on arrival order), so a superstep can be
// evaluated in any message order.  NLCC is evaluated level by level (one hop
// per level).  Where the reference's outcome DOES depend on arrival order the
// oracle raises a hazard counter instead of guessing (orc_run_hazards).

#include "pm_oracle.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <unordered_set>
#include <vector>
#include <unordered_map>
#include <map>
#include <deque>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

double now_s() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------------------
// hash_nbits  (detail/hash.hpp:65-143)
// ---------------------------------------------------------------------------
inline uint32_t mix32(uint32_t a) {  // hash32, :65-74
  a = (a + 0x7ed55d16u) + (a << 12);
  a = (a ^ 0xc761c23cu) ^ (a >> 19);
  a = (a + 0x165667b1u) + (a << 5);
  a = (a + 0xd3a2646cu) ^ (a << 9);
  a = (a + 0xfd7046c5u) + (a << 3);
  a = (a ^ 0xb55a4f09u) ^ (a >> 16);
  return a;
}
// hash16, :76-85.  The reference computes in `int` after integer promotion and
// truncates to uint16_t on every assignment; masking after each line is the same.
inline uint32_t mix16(uint32_t a) {
  a &= 0xffffu;
  a = ((a + 0x5d16u) + (a << 6)) & 0xffffu;
  a = ((a ^ 0xc23cu) ^ (a >> 9)) & 0xffffu;
  a = ((a + 0x67b1u) + (a << 5)) & 0xffffu;
  a = ((a + 0x646cu) ^ (a << 7)) & 0xffffu;
  a = ((a + 0x46c5u) + (a << 3)) & 0xffffu;
  a = ((a ^ 0x4f09u) ^ (a >> 8)) & 0xffffu;
  return a;
}
inline uint64_t window32(uint64_t x, int n) {  // shifted_n_hash32, :87-99
  uint64_t h = mix32((uint32_t)((x >> n) & 0xffffffffull));
  uint64_t mask = 0xffffffffull << n;
  return (x & ~mask) | (h << n);
}
inline uint64_t window16(uint64_t x, int n) {  // shifted_n_hash16, :101-113
  uint64_t h = mix16((uint32_t)((x >> n) & 0xffffull));
  uint64_t mask = 0xffffull << n;
  return (x & ~mask) | (h << n);
}
uint64_t hash_nbits(uint64_t x, int n) {  // :115-143
  if (n == 32) {
    x = mix32((uint32_t)x);
  } else if (n > 32) {
    int k = n - 32;
    for (int i = 0; i <= k; ++i) x = window32(x, i);
    for (int i = k; i >= 0; --i) x = window32(x, i);
  } else {
    int k = n - 16;  // reference asserts n > 16
    for (int i = 0; i <= k; ++i) x = window16(x, i);
    for (int i = k; i >= 0; --i) x = window16(x, i);
  }
  return x;
}

// ---------------------------------------------------------------------------
// R-MAT edge (rmat_edge_generator.hpp:218-259).  Compiled with
// -ffp-contract=off: the reference's doubles are evaluated without FMA.
// ---------------------------------------------------------------------------
struct Uniform01 {  // boost::uniform_01<boost::mt19937>, Boost 1.57 legacy form
  std::mt19937 eng;
  explicit Uniform01(uint32_t seed) : eng(seed) {}
  double operator()() { return (double)eng() * (1.0 / 4294967296.0); }
};

inline void rmat_edge(Uniform01& gen, uint64_t scale, uint64_t& u_out, uint64_t& v_out) {
  double a = 0.57, b = 0.19, c = 0.19, d = 0.05;  // generate_rmat.cpp:204
  uint64_t u = 0, v = 0;
  uint64_t step = (uint64_t(1) << scale) / 2;
  for (uint64_t j = 0; j < scale; ++j) {
    double p = gen();
    if (p < a) {
    } else if (p >= a && p < a + b) {
      v += step;
    } else if (p >= a + b && p < a + b + c) {
      u += step;
    } else {
      u += step;
      v += step;
    }
    step /= 2;
    a *= 0.9 + 0.2 * gen();
    b *= 0.9 + 0.2 * gen();
    c *= 0.9 + 0.2 * gen();
    d *= 0.9 + 0.2 * gen();
    double S = a + b + c + d;
    a /= S;
    b /= S;
    c /= S;
    d = 1. - a - b - c;
  }
  u_out = hash_nbits(u, (int)scale);  // scramble = true, generate_rmat.cpp:205
  v_out = hash_nbits(v, (int)scale);
}

}  // namespace

// ---------------------------------------------------------------------------
// graph
// ---------------------------------------------------------------------------
struct orc_graph {
  uint64_t V = 0;
  uint64_t slots_multi = 0;
  std::vector<uint64_t> rowptr;  // distinct-neighbour CSR
  std::vector<uint32_t> col;
  std::vector<uint64_t> rev;     // mirror slot of (v,u) = slot of (u,v)
  std::vector<uint64_t> degree;  // multigraph out-degree (dups + self loops)
};

namespace {

// src/dst hold every directed slot (u32 ids).  Builds degree, dedup CSR, rev.
orc_graph* build_graph(uint64_t V, std::vector<uint32_t>& src, std::vector<uint32_t>& dst) {
  orc_graph* g = new orc_graph();
  g->V = V;
  const uint64_t M = src.size();
  g->slots_multi = M;
  g->degree.assign(V, 0);
  // out-degree with duplicates and self loops (ipp:437-470, degree() ipp:1768-1781)
  for (uint64_t i = 0; i < M; ++i) g->degree[src[i]]++;
  std::vector<uint64_t> mrow(V + 1, 0);
  for (uint64_t v = 0; v < V; ++v) mrow[v + 1] = mrow[v] + g->degree[v];
  std::vector<uint32_t> mcol(M);
  {
    std::vector<uint64_t> pos(mrow.begin(), mrow.end() - 1);
    for (uint64_t i = 0; i < M; ++i) mcol[pos[src[i]]++] = dst[i];
  }
  std::vector<uint32_t>().swap(src);
  std::vector<uint32_t>().swap(dst);
  std::vector<uint64_t> ddeg(V);
#pragma omp parallel for schedule(dynamic, 4096)
  for (uint64_t v = 0; v < V; ++v) {
    auto b = mcol.begin() + mrow[v], e = mcol.begin() + mrow[v + 1];
    std::sort(b, e);
    ddeg[v] = (uint64_t)(std::unique(b, e) - b);
  }
  g->rowptr.assign(V + 1, 0);
  for (uint64_t v = 0; v < V; ++v) g->rowptr[v + 1] = g->rowptr[v] + ddeg[v];
  g->col.resize(g->rowptr[V]);
#pragma omp parallel for schedule(dynamic, 4096)
  for (uint64_t v = 0; v < V; ++v)
    std::copy(mcol.begin() + mrow[v], mcol.begin() + mrow[v] + ddeg[v],
              g->col.begin() + g->rowptr[v]);
  std::vector<uint32_t>().swap(mcol);
  g->rev.resize(g->col.size());
#pragma omp parallel for schedule(dynamic, 4096)
  for (uint64_t v = 0; v < V; ++v) {
    for (uint64_t j = g->rowptr[v]; j < g->rowptr[v + 1]; ++j) {
      uint32_t u = g->col[j];
      auto b = g->col.begin() + g->rowptr[u], e = g->col.begin() + g->rowptr[u + 1];
      auto it = std::lower_bound(b, e, (uint32_t)v);
      // symmetric input is a precondition ("only handling undirected graphs",
      // ..._nonunique_ee.hpp:592); a missing mirror is marked and never used.
      g->rev[j] = (it != e && *it == (uint32_t)v) ? (uint64_t)(it - g->col.begin()) : UINT64_MAX;
    }
  }
  return g;
}

}  // namespace

extern "C" {

void orc_rmat_stream(uint64_t scale, uint64_t rank, uint64_t n_edges, uint64_t* out) {
  Uniform01 gen((uint32_t)(5489ull + 3ull * rank));  // generate_rmat.cpp:202
  for (uint64_t i = 0; i < n_edges; ++i) rmat_edge(gen, scale, out[2 * i], out[2 * i + 1]);
}

uint64_t orc_hash_nbits(uint64_t input, int n) { return hash_nbits(input, n); }

orc_graph* orc_graph_from_slots(uint64_t n_vertices, uint64_t n_slots, const uint64_t* src,
                                const uint64_t* dst) {
  std::vector<uint32_t> s(n_slots), d(n_slots);
  for (uint64_t i = 0; i < n_slots; ++i) {
    s[i] = (uint32_t)src[i];
    d[i] = (uint32_t)dst[i];
  }
  return build_graph(n_vertices, s, d);
}

orc_graph* orc_graph_rmat(uint64_t scale, uint64_t gen_ranks, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  const uint64_t V = uint64_t(1) << scale;
  const uint64_t per_rank = V * 16 / gen_ranks;  // generate_rmat.cpp:201
  const uint64_t n_gen = per_rank * gen_ranks;
  std::vector<uint32_t> src(2 * n_gen), dst(2 * n_gen);
  // Each generated edge consumes exactly 5*scale draws (uniform_01 never
  // rejects), so a stream can be cut into chunks by discarding draws.
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  uint64_t chunks_per_rank = std::max<uint64_t>(1, (uint64_t)(2 * nthreads) / gen_ranks);
  uint64_t chunk_len = (per_rank + chunks_per_rank - 1) / chunks_per_rank;
  int64_t n_chunks = (int64_t)(gen_ranks * chunks_per_rank);
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t ci = 0; ci < n_chunks; ++ci) {
    uint64_t r = (uint64_t)ci / chunks_per_rank, k = (uint64_t)ci % chunks_per_rank;
    uint64_t e0 = k * chunk_len, e1 = std::min(per_rank, e0 + chunk_len);
    if (e0 >= e1) continue;
    Uniform01 gen((uint32_t)(5489ull + 3ull * r));
    gen.eng.discard(e0 * 5ull * scale);
    for (uint64_t e = e0; e < e1; ++e) {
      uint64_t u, v;
      rmat_edge(gen, scale, u, v);
      // iterator yields (u,v) then the swapped pair (rmat_edge_generator.hpp:126-139)
      uint64_t o = 2 * (r * per_rank + e);
      src[o] = (uint32_t)u;
      dst[o] = (uint32_t)v;
      src[o + 1] = (uint32_t)v;
      dst[o + 1] = (uint32_t)u;
    }
  }
  return build_graph(V, src, dst);
}

void orc_graph_free(orc_graph* g) { delete g; }
uint64_t orc_graph_num_vertices(const orc_graph* g) { return g->V; }
uint64_t orc_graph_num_slots_multi(const orc_graph* g) { return g->slots_multi; }
uint64_t orc_graph_num_slots(const orc_graph* g) { return g->col.size(); }
const uint64_t* orc_graph_rowptr(const orc_graph* g) { return g->rowptr.data(); }
const uint32_t* orc_graph_col(const orc_graph* g) { return g->col.data(); }
const uint64_t* orc_graph_degree(const orc_graph* g) { return g->degree.data(); }

void orc_labels_degree_log2(const orc_graph* g, uint64_t* out) {
  // vertex_data_db_degree.hpp:109, evaluated in double exactly as written there
  for (uint64_t v = 0; v < g->V; ++v)
    out[v] = static_cast<uint64_t>(std::ceil(std::log2((double)(g->degree[v] + 1))));
}

}  // extern "C"

// ---------------------------------------------------------------------------
// pattern directory
// ---------------------------------------------------------------------------
struct orc_constraint {
  std::vector<uint64_t> P;   // labels along the walk
  std::vector<uint32_t> I;   // template vertex ids along the walk
  uint64_t C = 0;            // pattern_cycle_length = max itr_count
  bool valid_cycle = false, interleave = false, selected = false;
  std::vector<uint32_t> enumidx;  // pattern_non_local_constraint field 2
  std::vector<uint32_t> agg;      // field 3 (read, unused by the reference)
};

struct orc_pattern {
  int nv = 0, ne = 0, diameter = 0;
  std::vector<uint64_t> vlabel;
  uint16_t N[16] = {0};     // template neighbours over mandatory edges (every edge of an exact pattern)
  // approximate matching: approximate_pattern_matching/pattern_graph.hpp:282-337, 604-622
  uint16_t No[16] = {0};    // ... over optional edges ("s t 0" lines of pattern_edge)
  int min_opt[16] = {0};    // vertex_min_optional_edge_count (pattern_vertex_local_constraints, "v : count")
  bool approximate = false;
  std::vector<orc_constraint> cons;
  std::string err;
};

namespace {

std::string trim(const std::string& s) {
  size_t b = s.find_first_not_of(" \t\r\n"), e = s.find_last_not_of(" \t\r\n");
  return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
}
std::vector<std::string> split_colon(const std::string& line) {
  std::vector<std::string> out;
  std::string tok;
  std::istringstream iss(line);
  while (std::getline(iss, tok, ':')) out.push_back(trim(tok));
  return out;
}
template <class T>
std::vector<T> parse_uints(const std::string& s) {
  std::vector<T> out;
  std::istringstream iss(s);
  unsigned long long x;
  while (iss >> x) out.push_back((T)x);
  return out;
}

}  // namespace

extern "C" {

orc_pattern* orc_pattern_load(const char* dir) {
  orc_pattern* p = new orc_pattern();
  std::string base = std::string(dir) + "/pattern";
  std::string line;
  // pattern_edge: "s t", both directions, sorted by s (graph.hpp:195-207,224-270)
  {
    std::ifstream f(base + "_edge");
    if (!f) { p->err = "cannot open " + base + "_edge"; return p; }
    long long last_s = -1;
    while (std::getline(f, line)) {
      if (trim(line).empty()) continue;
      std::istringstream iss(line);
      unsigned long long s = 0, t = 0, flag = 1;
      iss >> s >> t;
      const bool has_flag = (bool)(iss >> flag);  // "s t flag": 1 mandatory, 0 optional (approximate .. pattern_graph.hpp:320-337)
      if (s >= 16 || t >= 16) { p->err = "template vertex id >= 16 (beta.cpp:270-271)"; return p; }
      if ((long long)s < last_s) { p->err = "pattern_edge not sorted by source (graph.hpp:224-270)"; return p; }
      last_s = (long long)s;
      if (has_flag && flag == 0) {
        p->No[s] |= (uint16_t)(1u << t);
        p->approximate = true;
      } else {
        p->N[s] |= (uint16_t)(1u << t);
      }
      p->ne++;
    }
    p->nv = (int)(last_s + 1);  // vertex count = last source + 1 (graph.hpp:226,262)
  }
  // pattern_vertex_data: "id label"; labels are taken in FILE ORDER (graph.hpp:181-193)
  {
    std::ifstream f(base + "_vertex_data");
    if (!f) { p->err = "cannot open " + base + "_vertex_data"; return p; }
    while (std::getline(f, line)) {
      if (trim(line).empty()) continue;
      std::istringstream iss(line);
      unsigned long long id = 0, lab = 0;
      iss >> id >> lab;
      p->vlabel.push_back(lab);
    }
    if (p->vlabel.size() > 16) { p->err = "more than 16 template vertices"; return p; }
  }
  // pattern_vertex_local_constraints: "v : min_optional_edge_count" (approximate .. pattern_graph.hpp:282-315)
  {
    std::ifstream f(base + "_vertex_local_constraints");
    while (f && std::getline(f, line)) {
      auto t = split_colon(line);
      if (t.size() < 2) continue;
      const unsigned long long v = std::stoull(t[0]);
      const long k = std::stol(t[1]);
      if (v < 16) p->min_opt[v] = (int)std::max<long>(k, 0);
      p->approximate = true;
    }
  }
  // pattern_stat: "diameter : <int>", key case-insensitive (graph.hpp:337-358)
  {
    std::ifstream f(base + "_stat");
    while (f && std::getline(f, line)) {
      auto t = split_colon(line);
      if (t.size() < 2) continue;
      std::string k = t[0];
      std::transform(k.begin(), k.end(), k.begin(), ::tolower);
      if (k == "diameter") p->diameter = (int)std::stoull(t[1]);
    }
  }
  // pattern_nlc (pattern_util.hpp:172-210) and pattern_non_local_constraint (:254-278)
  {
    std::ifstream f(base + "_nlc");
    while (f && std::getline(f, line)) {
      if (trim(line).empty()) continue;
      auto t = split_colon(line);
      if (t.size() < 6) { p->err = "pattern_nlc: expected 6 ':' separated fields"; return p; }
      orc_constraint c;
      c.P = parse_uints<uint64_t>(t[0]);
      c.I = parse_uints<uint32_t>(t[1]);
      c.C = std::stoull(t[2]);
      c.valid_cycle = std::stoull(t[3]) != 0;
      c.interleave = std::stoull(t[4]) != 0;
      c.selected = std::stoull(t[5]) != 0;
      if (c.P.size() != c.I.size() || c.P.size() != c.C + 2) {
        p->err = "pattern_nlc: walk length must be cycle_length + 2";
        return p;
      }
      for (auto i : c.I) if (i >= 16) { p->err = "pattern_nlc: template id >= 16"; return p; }
      p->cons.push_back(c);
    }
    std::ifstream f2(base + "_non_local_constraint");
    size_t k = 0;
    while (f2 && std::getline(f2, line)) {
      if (trim(line).empty()) continue;
      auto t = split_colon(line);
      if (t.size() < 3) { p->err = "pattern_non_local_constraint: expected 3 fields"; return p; }
      if (k < p->cons.size()) {
        p->cons[k].enumidx = parse_uints<uint32_t>(t[1]);
        p->cons[k].agg = parse_uints<uint32_t>(t[2]);
      }
      ++k;
    }
  }
  return p;
}
void orc_pattern_free(orc_pattern* p) { delete p; }
const char* orc_pattern_error(const orc_pattern* p) { return p->err.empty() ? nullptr : p->err.c_str(); }
int orc_pattern_num_vertices(const orc_pattern* p) { return p->nv; }
int orc_pattern_num_edges(const orc_pattern* p) { return p->ne; }
int orc_pattern_diameter(const orc_pattern* p) { return p->diameter; }
int orc_pattern_num_constraints(const orc_pattern* p) { return (int)p->cons.size(); }

}  // extern "C"

// ---------------------------------------------------------------------------
// run
// ---------------------------------------------------------------------------
struct orc_run {
  uint64_t V = 0;
  std::vector<uint8_t> active;    // vertex_active           (beta.cpp:320)
  std::vector<uint16_t> T_arr;    // template_vertices[v]    (beta.cpp:328)
  std::vector<uint8_t> inmap;     // v in vertex_state_map   (beta.cpp:307)
  std::vector<uint16_t> T_state;  // vertex_state.template_vertices
  std::vector<uint16_t> heard;    // vertex_state.template_neighbors
  std::vector<uint8_t> estate;    // per distinct slot: 0 absent, 1 in E_v flag 0, 2 in E_v flag 1 (heard this superstep), 3 flag 1 set by nem_1
  std::vector<orc_row> rows;
  uint64_t iterations = 0;
  double search_seconds = 0;
  std::vector<std::vector<uint32_t>> subgraphs;  // per constraint, flattened
  std::vector<int> subgraph_width;
  uint64_t path_count = 0;  // file-static, never reset (tds_batch_1.hpp:14,1243)
  uint64_t edges_processed = 0;
  uint64_t hazards[8] = {0};
  int n_ranks = 1;
  std::vector<uint32_t> hubs;  // sorted hub vertex ids (delegate ids), ipp:681
  std::string err;
  const orc_graph* g = nullptr;
  std::vector<uint64_t> rank_counts;                    // per row: R x {vertices, edges}
  std::vector<std::pair<uint64_t, double>> step_rows;   // result_step: (itr, LP seconds)
  std::vector<double> iter_seconds;                     // result_iteration
};

namespace {
int owner_rank(const orc_run* r, uint64_t v) {
  // non-delegates: v mod R (ipp:1682-1697); hubs: controller = delegate_id mod R
  // (delegate_partitioned_graph.hpp:231-233), delegate ids follow the sorted hub list (ipp:681)
  if (!r->hubs.empty()) {
    auto it = std::lower_bound(r->hubs.begin(), r->hubs.end(), (uint32_t)v);
    if (it != r->hubs.end() && *it == (uint32_t)v) return (int)((it - r->hubs.begin()) % r->n_ranks);
  }
  return (int)(v % (uint64_t)r->n_ranks);
}
}  // namespace

namespace {

struct Ctx {
  const orc_graph* g;
  const uint64_t* label;
  const orc_pattern* pat;
  orc_run* r;
  orc_options opt;
  std::vector<uint16_t> NB;  // NB[T] = OR of N(a), a in T  (valid-parent test, ee.hpp:673-722)
  bool keep_sub = true;

  uint16_t labelmask(uint64_t lab) const {  // ee.hpp:371-380, 523-537
    uint16_t m = 0;
    for (size_t i = 0; i < pat->vlabel.size(); ++i)
      if (pat->vlabel[i] == lab) m |= (uint16_t)(1u << i);
    return m;
  }
  // bits p of T whose template neighbourhood is non-empty and fully heard (ee.hpp:901-939)
  uint16_t cover(uint16_t T, uint16_t heard) const {
    uint16_t out = 0;
    for (int p = 0; p < 16; ++p)
      if ((T >> p) & 1) {
        uint16_t need = pat->N[p];
        if (!pat->approximate) {
          if (need != 0 && (need & heard) == need) out |= (uint16_t)(1u << p);
          continue;
        }
        // approximate local constraint, approximate_pattern_matching/local_constraint_checking.hpp:1062-1113
        bool mandatory_ok = need == 0 || (need & heard) == need;                               // :1080-1088
        bool optional_ok = true;                                                               // :1090-1099
        if (pat->min_opt[p] > 0) {
          uint16_t got = pat->No[p] & heard;
          optional_ok = got == pat->No[p] && __builtin_popcount(got) >= pat->min_opt[p];
        }
        if (mandatory_ok && optional_ok) out |= (uint16_t)(1u << p);
      }
    return out;
  }
  void count(uint64_t& nv, uint64_t& ne) const {  // ee.hpp:1112-1128, beta.cpp:1094-1110
    const uint64_t V = g->V;
    const int Rk = r->n_ranks;
    std::vector<uint64_t> acc((size_t)Rk * 2, 0);
#pragma omp parallel
    {
      std::vector<uint64_t> loc((size_t)Rk * 2, 0);
#pragma omp for schedule(static, 8192) nowait
      for (uint64_t v = 0; v < V; ++v)
        if (r->inmap[v]) {
          int k = owner_rank(r, v);
          loc[2 * k]++;
          uint64_t ce = 0;
          for (uint64_t j = g->rowptr[v]; j < g->rowptr[v + 1]; ++j) ce += r->estate[j] != 0;
          loc[2 * k + 1] += ce;
        }
#pragma omp critical
      for (size_t i = 0; i < acc.size(); ++i) acc[i] += loc[i];
    }
    nv = ne = 0;
    for (int k = 0; k < Rk; ++k) { nv += acc[2 * k]; ne += acc[2 * k + 1]; }
    r->rank_counts.insert(r->rank_counts.end(), acc.begin(), acc.end());
  }

  // ------------------------------------------------------------------ LCC
  // label_propagation_pattern_matching_bsp, ee.hpp:1029-1153
  void lcc(bool init, uint64_t itr, bool& not_finished) {
    const uint64_t V = g->V;
    auto& R = *r;
    double call_t0 = now_s();
    for (int k = 0; k < pat->diameter; ++k) {  // fixed superstep count, ee.hpp:1069
      const bool first = (k == 0 && init);
      double t0 = now_s();
      uint64_t msgs = 0, asym = 0;
      if (first) {
        // every vertex evaluates its own label once (sender side ee.hpp:519-546;
        // receivers recompute the same value ee.hpp:368-401)
#pragma omp parallel for schedule(static, 8192)
        for (uint64_t v = 0; v < V; ++v)
          if (R.active[v]) {
            uint16_t lm = labelmask(label[v]);
            if (lm == 0) R.active[v] = 0; else R.T_arr[v] = lm;
          }
      }
      // message phase: u sends T_arr(u) along CSR (first) or keys(E_u) (later)
#pragma omp parallel for reduction(+ : msgs, asym) schedule(dynamic, 2048)
      for (uint64_t u = 0; u < V; ++u) {
        if (!R.active[u]) continue;                 // ee.hpp:470
        if (!first && !R.inmap[u]) continue;        // ee.hpp:481-486
        const uint16_t m = R.T_arr[u];
        if (m == 0) continue;                       // ee.hpp:541-543, 575-577
        for (uint64_t j = g->rowptr[u]; j < g->rowptr[u + 1]; ++j) {
          if (!first && R.estate[j] == 0) continue; // ee.hpp:589 iterates E_u only
          const uint32_t v = g->col[j];
          const uint64_t jr = g->rev[j];
          ++msgs;
          // ---- receiver pre_visit, ee.hpp:148-459
          if (!R.active[v]) continue;               // :151
          if (!first && !R.inmap[v]) continue;      // :414-417
          const uint16_t Tv = R.T_arr[v];           // first: == labelmask(v) (:401)
          if (Tv == 0) continue;                    // :434-437
          // ---- verify_and_update_vertex_state, ee.hpp:646-816
          if ((NB[Tv] & m) == 0) continue;          // no valid parent, :719-722
          if (first && !R.inmap[v]) {               // :734-746 (same value from every sender)
            R.T_state[v] = Tv;
            R.inmap[v] = 1;
          }
          __atomic_fetch_or(&R.heard[v], m, __ATOMIC_RELAXED);  // :775
          if (first) {
            if (jr != UINT64_MAX) R.estate[jr] = 2; // insert / overwrite flag 1, :794-807
          } else if (jr != UINT64_MAX && R.estate[jr] != 0) {
            R.estate[jr] = 2;                       // :811
          } else {
            ++asym;                                 // "did not find the expected item", :801
          }
        }
      }
      R.edges_processed += msgs;
      R.hazards[2] += asym;
      // ---- post step, ee.hpp:827-1027
      uint64_t removed = 0, grew = 0, flagged = 0;
#pragma omp parallel for reduction(+ : removed, grew, flagged) schedule(dynamic, 2048)
      for (uint64_t v = 0; v < V; ++v) {
        if (first && R.active[v] && !R.inmap[v]) {  // :841-852
          R.active[v] = 0;
          for (uint64_t j = g->rowptr[v]; j < g->rowptr[v + 1]; ++j) R.estate[j] = 0;
          continue;
        }
        if (!R.inmap[v]) continue;
        uint16_t ts = cover(R.T_state[v], R.heard[v]);  // :901-939
        if (ts == 0) {                              // :941-946
          R.inmap[v] = 0;
          R.active[v] = 0;
          for (uint64_t j = g->rowptr[v]; j < g->rowptr[v + 1]; ++j) R.estate[j] = 0;
          ++removed;
        } else {                                    // :947-964
          if (ts & ~R.T_arr[v]) ++grew;
          R.T_state[v] = ts;
          R.T_arr[v] = ts;
          R.heard[v] = 0;
          for (uint64_t j = g->rowptr[v]; j < g->rowptr[v + 1]; ++j) {
            flagged += R.estate[j] == 3;  // survives this post step on the flag alone, :954-962
            R.estate[j] = (R.estate[j] >= 2) ? 1 : 0;
          }
        }
      }
      if (removed) not_finished = true;             // :968-970
      R.hazards[3] += grew;
      R.hazards[5] += flagged;
      double t1 = now_s();
      orc_row row;
      row.itr = itr; row.kind = 0; row.index = k; row.seconds = t1 - t0;
      count(row.n_vertices, row.n_edges);
      R.rows.push_back(row);
    }
    R.step_rows.push_back({itr, now_s() - call_t0});  // beta.cpp:593-596, 1194-1197
  }

  // ------------------------------------------------------------ NLCC nem_1
  static inline uint64_t key(uint32_t v, uint32_t s) { return ((uint64_t)v << 32) | s; }

  bool hop_ok(const orc_constraint& c, uint32_t v, uint64_t h) const {
    // nem_1.hpp:101 (active), :186 (label), :192-210 (template bit)
    return r->active[v] && label[v] == c.P[h] && ((r->T_arr[v] >> c.I[h]) & 1);
  }

  struct Tok { uint32_t v, s, parent; };

  void nem1(const orc_constraint& c, std::vector<uint32_t>& sources, std::vector<uint8_t>& ok) {
    const uint64_t V = g->V;
    auto& R = *r;
    const size_t n = c.P.size();
    std::vector<Tok> cur, nxt;
    // sources, nem_1.hpp:387-527
    for (uint64_t v = 0; v < V; ++v) {
      if (!R.active[v] || label[v] != c.P[0]) continue;           // :314, :428
      uint16_t T = R.T_arr[v];
      if (T == 0 || !((T >> c.I[0]) & 1)) continue;                // :440
      if (!c.valid_cycle && !((T >> c.I[n - 1]) & 1)) continue;    // :447-451
      sources.push_back((uint32_t)v);                              // token_source_map[v]=0, :469-477
      for (uint64_t j = g->rowptr[v]; j < g->rowptr[v + 1]; ++j)   // :486-525
        if (R.estate[j]) cur.push_back({g->col[j], (uint32_t)v, (uint32_t)v});
    }
    std::unordered_set<uint64_t> seen;  // vertex_token_source_set, keyed (vertex, source)
    for (uint64_t h = 1; h <= c.C + 1 && !cur.empty(); ++h) {
      const bool interior = c.C > h - 1;  // max_itr_count > itr_count
      R.edges_processed += cur.size();
      std::sort(cur.begin(), cur.end(), [](const Tok& a, const Tok& b) {
        if (a.v != b.v) return a.v < b.v;
        if (a.s != b.s) return a.s < b.s;
        return a.parent < b.parent;
      });
      nxt.clear();
      size_t i = 0;
      while (i < cur.size()) {
        size_t e = i;
        while (e < cur.size() && cur[e].v == cur[i].v && cur[e].s == cur[i].s) ++e;
        const uint32_t v = cur[i].v, s = cur[i].s;
        const bool st = hop_ok(c, v, h);
        if (interior) {
          if (seen.count(key(v, s))) {              // :131-139 (checked before anything else)
            if (st && v != s) R.hazards[0]++;
          } else if (v != s && st) {                // :174-177, :186-210
            seen.insert(key(v, s));                 // :270-285
            size_t np = 1;
            for (size_t t = i + 1; t < e; ++t) np += cur[t].parent != cur[t - 1].parent;
            // forward along E_v; the reference skips the parent of the FIRST token
            // to arrive (:836-838).  One distinct parent: deterministic.  Several:
            // forward to all and raise the hazard if a skipped parent could matter.
            for (uint64_t j = g->rowptr[v]; j < g->rowptr[v + 1]; ++j) {
              if (!R.estate[j]) continue;
              uint32_t nb = g->col[j];
              if (np == 1 && nb == cur[i].parent) continue;
              nxt.push_back({nb, s, v});
            }
            if (np > 1) {
              for (size_t t = i; t < e; ++t) {
                if (t > i && cur[t].parent == cur[t - 1].parent) continue;
                uint32_t pv = cur[t].parent;
                if (!hop_ok(c, pv, h + 1)) continue;
                bool nxt_interior = c.C > h;
                bool viable = nxt_interior ? (pv != s && !seen.count(key(pv, s)))
                                           : (c.valid_cycle ? pv == s : pv != s);
                if (viable) R.hazards[1]++;
              }
            }
          }
        } else if (st) {                            // final hop, :661-791
          for (size_t t = i; t < e; ++t) {
            if (!c.valid_cycle) {
              if (v != s) ok[s] = 1;                // ack_success -> :326-342
            } else if (v == s) {
              ok[s] = 1;                            // :749-758
              // mark the edge the successful token arrived on, :764-770
              auto b = g->col.begin() + g->rowptr[v], en = g->col.begin() + g->rowptr[v + 1];
              auto it = std::lower_bound(b, en, cur[t].parent);
              if (it != en && *it == cur[t].parent && R.estate[it - g->col.begin()])
                R.estate[it - g->col.begin()] = 3;  // flag 1 set OUTSIDE LCC
            }
          }
        }
        i = e;
      }
      cur.swap(nxt);
    }
  }

  // --------------------------------------------------------------- NLCC TDS
  // tds_batch_1.hpp: history rule :284-302 / :622-639 / :808-886
  static bool hist_rule(const orc_constraint& c, const uint32_t* hist, uint64_t hp, uint32_t x) {
    if (hp >= c.enumidx.size()) return false;
    uint64_t e = c.enumidx[hp];
    if (e == hp) {
      for (uint64_t i = 0; i < hp; ++i) if (hist[i] == x) return false;
      return true;
    } else if (e < hp) {
      return hist[e] == x;
    }
    return false;  // "invalid value" branches drop the token
  }

  void tds(const orc_constraint& c, int pl, std::vector<uint32_t>& sources, std::vector<uint8_t>& ok) {
    const uint64_t V = g->V;
    auto& R = *r;
    std::vector<uint32_t> cur, nxt;  // flattened histories, stride h+1
    for (uint64_t v = 0; v < V; ++v) {  // :1067-1100 and :425-512
      if (!R.active[v] || label[v] != c.P[0]) continue;
      uint16_t T = R.T_arr[v];
      if (T == 0 || !((T >> c.I[0]) & 1)) continue;
      sources.push_back((uint32_t)v);
      for (uint64_t j = g->rowptr[v]; j < g->rowptr[v + 1]; ++j)
        if (R.estate[j]) { cur.push_back((uint32_t)v); cur.push_back(g->col[j]); }
    }
    auto& out = R.subgraphs[pl];
    out.clear();  // the file is truncated on every outer iteration, beta.cpp:713-717
    R.subgraph_width[pl] = (int)(c.C + 2);
    for (uint64_t h = 1; h <= c.C + 1 && !cur.empty(); ++h) {
      const uint64_t stride = h + 1;
      const uint64_t ntok = cur.size() / stride;
      const bool interior = c.C > h - 1;
      R.edges_processed += ntok;
      nxt.clear();
      for (uint64_t t = 0; t < ntok; ++t) {
        const uint32_t* hist = &cur[t * stride];
        const uint32_t v = hist[h], s = hist[0];
        if (!hop_ok(c, v, h)) continue;                          // :125, :207-233
        if (interior) {
          if (!hist_rule(c, hist, h, v)) continue;               // :284-302
          for (uint64_t j = g->rowptr[v]; j < g->rowptr[v + 1]; ++j) {  // :793-909
            if (!R.estate[j]) continue;
            uint32_t nb = g->col[j];
            if (c.C == h) {                                      // penultimate hop, :808-845
              if (c.valid_cycle) { if (nb != s) continue; }
              else { if (nb == s) continue; if (!hist_rule(c, hist, h + 1, nb)) continue; }
            } else if (!hist_rule(c, hist, h + 1, nb)) {         // :846-886
              continue;
            }
            nxt.insert(nxt.end(), hist, hist + stride);
            nxt.push_back(nb);
          }
        } else {                                                 // final hop, :641-754
          bool success = c.valid_cycle ? (v == s) : (v != s);
          if (!success) continue;
          ok[s] = 1;
          R.path_count++;
          if (keep_sub) out.insert(out.end(), hist, hist + stride);
        }
      }
      cur.swap(nxt);
    }
    if (!keep_sub) out.clear();
  }
};

}  // namespace

extern "C" {

orc_run* orc_run_pattern(const orc_graph* g, const uint64_t* labels, const orc_pattern* pat,
                         const orc_options* opt_in) {
  orc_run* r = new orc_run();
  orc_options opt = *opt_in;
#ifdef _OPENMP
  if (opt.threads > 0) omp_set_num_threads(opt.threads);
#endif
  const uint64_t V = g->V;
  r->V = V;
  r->g = g;
  r->n_ranks = opt.n_ranks > 0 ? opt.n_ranks : 1;
  if (!pat->err.empty()) { r->err = pat->err; return r; }
  for (auto& c : pat->cons)
    if (c.selected) { r->err = "selected_vertices=1 constraints are not supported"; return r; }
  // beta.cpp:484-492
  r->active.assign(V, 1);
  r->T_arr.assign(V, 0);
  r->inmap.assign(V, 0);
  r->T_state.assign(V, 0);
  r->heard.assign(V, 0);
  r->estate.assign(g->col.size(), 0);
  r->subgraphs.resize(pat->cons.size());
  r->subgraph_width.assign(pat->cons.size(), 0);
  if (opt.delegate_threshold)
    for (uint64_t v = 0; v < V; ++v)
      if (g->degree[v] >= opt.delegate_threshold) r->hubs.push_back((uint32_t)v);

  Ctx cx{g, labels, pat, r, opt, {}, opt.keep_subgraphs != 0};
  cx.NB.assign(65536, 0);
  for (uint32_t T = 1; T < 65536; ++T) {
    uint32_t low = T & (~T + 1);
    int b = __builtin_ctz(low);
    cx.NB[T] = (uint16_t)(cx.NB[T ^ low] | pat->N[b] | pat->No[b]);  // valid parents come over mandatory and optional edges (approximate .. local_constraint_checking.hpp:641-651)
  }

  const int max_it = opt.max_iterations > 0 ? opt.max_iterations : 1000;
  bool init = true, nf = false;
  uint64_t itr = 0;
  double t_begin = now_s();
  do {  // beta.cpp:544-1351
    double it0 = now_s();
    nf = false;
    cx.lcc(init, itr, nf);            // :577-583
    init = false;                     // :602-604
    if (opt.lcc_only) {
      // OUR extension for BASELINE config 2 ("LCC-only pruning to fixed point"):
      // repeat LCC calls while a call removed a vertex from the map (the
      // reference's own continue signal, ee.hpp:968-970); no NLCC at all.
    } else {
      if (itr == 0) nf = true;        // forced token passing, :686-688
      if (nf) {                       // :695
        nf = false;                   // :697
        for (size_t pl = 0; pl < pat->cons.size(); ++pl) {  // :710
          const orc_constraint& c = pat->cons[pl];
          double t0 = now_s();
          std::vector<uint32_t> sources;        // token_source_map keys (cleared, :791-793)
          std::vector<uint8_t> ok(V, 0);
          bool do_tds = opt.tds_from_pl >= 0 && (int)pl >= opt.tds_from_pl;  // :762-767
          if (do_tds) cx.tds(c, (int)pl, sources, ok); else cx.nem1(c, sources, ok);
          bool deleted = false;
          for (uint32_t s : sources) {          // :964-1005
            if (ok[s]) continue;
            uint16_t T = r->T_arr[s];
            if (T == 0) continue;
            T &= (uint16_t)~(1u << c.I[0]);
            r->T_arr[s] = T;
            if (T == 0) r->active[s] = 0;
            nf = true;
            deleted = true;
          }
          for (uint32_t s : sources)            // :1043-1062
            if (!r->active[s] && r->inmap[s]) r->inmap[s] = 0;
          double t1 = now_s();
          orc_row row;
          row.itr = itr; row.kind = 1; row.index = (int)pl; row.seconds = t1 - t0;
          cx.count(row.n_vertices, row.n_edges);  // :1094-1120
          r->rows.push_back(row);
          if (deleted && c.interleave) cx.lcc(false, itr, nf);  // :1163-1184
        }
      }
    }
    r->iter_seconds.push_back(now_s() - it0);  // :1337-1338
    ++itr;                            // :1341
    if ((int)itr >= max_it && nf) { r->hazards[4]++; break; }
  } while (nf);                       // :1351
  r->iterations = itr;
  r->search_seconds = now_s() - t_begin;
  return r;
}

void orc_run_free(orc_run* r) { delete r; }
uint64_t orc_run_num_rows(const orc_run* r) { return r->rows.size(); }
const orc_row* orc_run_rows(const orc_run* r) { return r->rows.data(); }
uint64_t orc_run_iterations(const orc_run* r) { return r->iterations; }
double orc_run_search_seconds(const orc_run* r) { return r->search_seconds; }
const uint16_t* orc_run_template_vertices(const orc_run* r) { return r->T_arr.data(); }
const uint8_t* orc_run_in_map(const orc_run* r) { return r->inmap.data(); }
const char* orc_run_error(const orc_run* r) { return r->err.empty() ? nullptr : r->err.c_str(); }
uint64_t orc_run_cumulative_path_count(const orc_run* r) { return r->path_count; }
uint64_t orc_run_edges_processed(const orc_run* r) { return r->edges_processed; }
const uint64_t* orc_run_hazards(const orc_run* r) { return r->hazards; }
uint64_t orc_run_num_subgraphs(const orc_run* r, int pl) {
  if (pl < 0 || (size_t)pl >= r->subgraphs.size() || r->subgraph_width[pl] == 0) return 0;
  return r->subgraphs[pl].size() / r->subgraph_width[pl];
}
int orc_run_subgraph_width(const orc_run* r, int pl) { return r->subgraph_width[pl]; }
const uint32_t* orc_run_subgraphs(const orc_run* r, int pl) { return r->subgraphs[pl].data(); }

}  // extern "C"

extern "C" uint64_t orc_run_num_active_edges(const orc_run* r) {
  uint64_t n = 0;
  const orc_graph* g = r->g;
  for (uint64_t v = 0; v < g->V; ++v) {
    if (!r->inmap[v]) continue;  // beta.cpp:1386 iterates vertex_state_map
    for (uint64_t j = g->rowptr[v]; j < g->rowptr[v + 1]; ++j) n += r->estate[j] != 0;
  }
  return n;
}
extern "C" void orc_run_active_edges(const orc_run* r, uint64_t* pairs_out) {
  uint64_t n = 0;
  const orc_graph* g = r->g;
  for (uint64_t v = 0; v < g->V; ++v) {
    if (!r->inmap[v]) continue;
    for (uint64_t j = g->rowptr[v]; j < g->rowptr[v + 1]; ++j)
      if (r->estate[j]) { pairs_out[2 * n] = v; pairs_out[2 * n + 1] = g->col[j]; ++n; }
  }
}

namespace {
std::string bits16(uint16_t x) {  // std::bitset<16> stream output, MSB first
  std::string s(16, '0');
  for (int i = 0; i < 16; ++i) if ((x >> i) & 1) s[15 - i] = '1';
  return s;
}
}  // namespace

extern "C" int orc_run_write_results(const orc_run* r, const orc_graph* g, const uint64_t* labels,
                                     const orc_pattern* pat, const char* outdir) {
  const int R = r->n_ranks;
  std::string base = std::string(outdir), ps = base + "/0";
  auto open = [](const std::string& p, std::ofstream& f) { f.open(p, std::ofstream::out); return (bool)f; };
  std::ofstream f_set, f_itr, f_step, f_ss;
  if (!open(base + "/result_pattern_set", f_set)) return -1;   // beta.cpp:413-414
  if (!open(ps + "/result_iteration", f_itr)) return -1;       // :504-511
  if (!open(ps + "/result_step", f_step)) return -1;
  if (!open(ps + "/result_superstep", f_ss)) return -1;
  // rank 0 files: times
  for (size_t i = 0; i < r->iter_seconds.size(); ++i) f_itr << i << ", " << r->iter_seconds[i] << "\n";  // beta.cpp:1337
  for (auto& sr : r->step_rows) f_step << sr.first << ", LP, " << sr.second << "\n";                   // :594, :1195
  for (const orc_row& w : r->rows)   // ee.hpp:1106-1108, beta.cpp:1088-1090
    f_ss << w.itr << (w.kind == 0 ? ", LP, " : ", TP, ") << w.index << ", " << w.seconds << "\n";
  f_set << 0 << ", " << R << ", " << r->iterations << ", " << r->search_seconds << ", "
        << pat->ne << ", " << pat->nv << ", " << pat->cons.size() << "\n";  // beta.cpp:1375-1381
  for (int k = 0; k < R; ++k) {
    std::ofstream fvc, fec, fv, fe, fm;
    if (!open(ps + "/all_ranks_active_vertices_count/active_vertices_" + std::to_string(k), fvc)) return -1;
    if (!open(ps + "/all_ranks_active_edges_count/active_edges_" + std::to_string(k), fec)) return -1;
    if (!open(ps + "/all_ranks_active_vertices/active_vertices_" + std::to_string(k), fv)) return -1;
    if (!open(ps + "/all_ranks_active_edges/active_edges_" + std::to_string(k), fe)) return -1;
    if (!open(ps + "/all_ranks_messages/messages_" + std::to_string(k), fm)) return -1;
    for (size_t i = 0; i < r->rows.size(); ++i) {
      const orc_row& w = r->rows[i];
      const char* kind = w.kind == 0 ? ", LP, " : ", TP, ";
      fvc << w.itr << kind << w.index << ", " << r->rank_counts[(i * R + k) * 2] << "\n";      // ee.hpp:1131-1133
      fec << w.itr << kind << w.index << ", " << r->rank_counts[(i * R + k) * 2 + 1] << "\n";  // ee.hpp:1136-1138
      fm << w.itr << kind << w.index << ", 0\n";  // message counts are transport specific
    }
    for (uint64_t v = 0; v < g->V; ++v) {
      if (!r->inmap[v] || owner_rank(r, v) != k) continue;
      fv << k << ", " << v << ", 0, " << labels[v] << ", " << bits16(r->T_arr[v]) << "\n";  // beta.cpp:1390-1394
      for (uint64_t j = g->rowptr[v]; j < g->rowptr[v + 1]; ++j)
        if (r->estate[j]) fe << k << ", " << v << ", " << g->col[j] << "\n";                // :1398-1403
    }
    for (size_t pl = 0; pl < pat->cons.size(); ++pl) {
      std::ofstream fs;
      if (!open(ps + "/all_ranks_subgraphs/subgraphs_" + std::to_string(pl) + "_" + std::to_string(k), fs)) return -1;
      int w = r->subgraph_width[pl];
      if (w == 0) continue;
      const auto& sg = r->subgraphs[pl];
      for (size_t i = 0; i + w <= sg.size(); i += w) {
        uint32_t last = sg[i + w - 1];
        if (owner_rank(r, last) != k) continue;   // written by the rank that owns the final vertex
        fs << "[" << k << "], ";                  // tds_batch_1.hpp:685-689
        for (int x = 0; x < w; ++x) fs << sg[i + x] << ", ";
        fs << "[" << last << "]\n";
      }
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------
// run_fuzzy path: unique-label LCC + cycle token passing over the UNPRUNED adjacency.
//   LCC    include/havoqgt/label_propagation_pattern_matching_bsp.hpp:66-310 (visitor), :317-520
//          (per-message state), :527-593 (post step), :598-699 (superstep loop)
//   NLCC   include/havoqgt/token_passing_pattern_matching.hpp:51-508
//   loop   src/run_pattern_matching.cpp:340-722 (the compiling twin of the stale
//          src/run_fuzzy_pattern_matching.cpp:287-557; TP_ORIG is defined at :31)
// Restated literally, message by message, with the reference's containers:
//   vertex_state {is_active, vertex_pattern_index, pattern_vertex_itr_count_map}  (bsp.hpp:9-26)
//   vertex_active, vertex_state_map, token_source_map, vertex_token_source_set.
// ---------------------------------------------------------------------------
namespace {

struct FzState {
  bool is_active = false;
  uint32_t vpi = 0;
  std::map<uint32_t, uint8_t> count;  // pattern_vertex_itr_count_map
};

struct FzCtx {
  const orc_graph* g;
  const uint64_t* label;
  const orc_pattern* pat;
  std::vector<uint8_t> active;                 // vertex_active
  std::unordered_map<uint32_t, FzState> map;   // vertex_state_map
};

// lppm_visitor::verify_and_update_vertex_state_map (bsp.hpp:317-520) for itr_count == 1
int fz_verify_update(FzCtx& c, uint32_t v, uint32_t q, uint32_t parent_idx) {
  const orc_pattern& P = *c.pat;
  if (!((P.N[q] >> parent_idx) & 1u)) return 0;                        // :330-339
  auto it = c.map.find(v);
  if (it == c.map.end()) {                                             // :360-370
    it = c.map.insert({v, FzState()}).first;
    it->second.vpi = q;
  }
  FzState& st = it->second;
  if (st.is_active) return 1;                                          // :379-381
  if (st.count.empty())                                                // :385-414
    for (uint32_t b = 0; b < 16; ++b)
      if ((P.N[q] >> b) & 1u) st.count.insert({b, 0});
  if (st.count.empty()) return 0;                                      // :416-418
  auto f = st.count.find(parent_idx);
  if (f == st.count.end()) return 0;                                   // :420-430 ("did not find the expected item")
  if (f->second < 1) f->second = 1;                                    // :433-435
  bool all = true;                                                     // :438-466
  for (auto& kv : st.count)
    if (kv.second == 0) { all = false; break; }
  st.is_active = all;
  if (all)                                                             // :513-517
    for (auto& kv : st.count) kv.second = 0;
  return 1;
}

// one message of superstep `superstep` from u (as template index p) to v: pre_visit (:91-155) + visit (:163-310)
void fz_deliver(FzCtx& c, uint32_t v, uint32_t p, bool map_required) {
  const orc_pattern& P = *c.pat;
  if (!c.active[v]) return;                                            // :93-95
  // pre_visit: the FIRST template vertex carrying v's label decides (:113-141)
  bool match = false;
  for (uint32_t q = 0; q < (uint32_t)P.nv; ++q) {
    if (P.vlabel[q] != c.label[v]) continue;
    match = true;
    if (!((P.N[q] >> p) & 1u)) return;                                 // :133-135
    break;                                                             // :139-141
  }
  if (!match) return;
  // visit (:163-...)
  if (map_required && c.map.find(v) == c.map.end()) return;            // :173-178
  for (uint32_t q = 0; q < (uint32_t)P.nv; ++q)
    if (P.vlabel[q] == c.label[v]) fz_verify_update(c, v, q, p);       // :253-262
}

}  // namespace

extern "C" orc_run* orc_run_fuzzy(const orc_graph* g, const uint64_t* labels, const orc_pattern* pat,
                                   const orc_options* opt_in) {
  orc_run* r = new orc_run();
  r->g = g;
  r->V = g->V;
  orc_options opt{1, -1, 0, 0, 0, 0, 0};
  if (opt_in) opt = *opt_in;
  r->n_ranks = opt.n_ranks > 0 ? opt.n_ranks : 1;
  const uint64_t V = g->V;
  const orc_pattern& P = *pat;
  FzCtx c{g, labels, pat, std::vector<uint8_t>(V, 1), {}};
  // walks whose interior hops repeat a template vertex make the (source-keyed) aggregation set of
  // token_passing_pattern_matching.hpp:104-109 depend on message order: not restated
  for (auto& k : P.cons) {
    for (size_t a = 1; a <= k.C && a < k.I.size(); ++a)
      for (size_t b = a + 1; b <= k.C && b < k.I.size(); ++b)
        if (k.I[a] == k.I[b]) { r->err = "token walk repeats a template vertex at interior hops (order dependent)"; return r; }
  }
  const int max_it = opt.max_iterations > 0 ? opt.max_iterations : 1000;
  bool initstep = true, nf = false;
  const double t_begin = now_s();
  uint64_t itr = 0;
  do {
    nf = false;
    // ---- label_propagation_pattern_matching_bsp (bsp.hpp:598-699) ----
    for (int ss = 0; ss < P.diameter; ++ss) {
      const double t0 = now_s();
      const bool map_required = ss > 0 || !initstep;                   // :173
      // senders are fixed at the start of the superstep: membership only changes by insertion
      // during the very first superstep (where it is not consulted) and in the post step
      std::vector<uint32_t> senders;
      for (uint64_t u = 0; u < V; ++u) {
        if (!c.active[u]) continue;                                    // :165-167
        if (map_required && c.map.find((uint32_t)u) == c.map.end()) continue;
        bool match = false;
        for (int q = 0; q < P.nv; ++q) match = match || P.vlabel[q] == labels[u];
        if (!match) { c.active[u] = 0; continue; }                     // :199-203
        senders.push_back((uint32_t)u);
      }
      for (uint32_t u : senders)                                       // :207-225: all neighbours, all matching indices
        for (uint64_t e = g->rowptr[u]; e < g->rowptr[u + 1]; ++e) {
          r->edges_processed++;
          for (int p = 0; p < P.nv; ++p)
            if (P.vlabel[p] == labels[u]) fz_deliver(c, g->col[e], (uint32_t)p, map_required);
        }
      // post step (:527-593)
      std::vector<uint32_t> gone;
      for (auto& kv : c.map) {
        if (!kv.second.is_active) { gone.push_back(kv.first); c.active[kv.first] = 0; }
        else kv.second.is_active = false;
      }
      if (!gone.empty()) nf = true;
      for (uint32_t v : gone) c.map.erase(v);
      orc_row row{itr, 0, ss, c.map.size(), 0, now_s() - t0};
      r->rows.push_back(row);
    }
    initstep = false;
    // ---- token passing (run_pattern_matching.cpp:511-640), only if something was removed ----
    if (nf) {
      nf = false;
      for (size_t pl = 0; pl < P.cons.size(); ++pl) {
        const orc_constraint& k = P.cons[pl];
        std::unordered_map<uint32_t, bool> token_source_map;           // cleared per constraint (:523)
        std::unordered_map<uint32_t, std::unordered_set<uint32_t>> forwarded;  // vertex_token_source_set
        struct Tok { uint32_t v, target, itr, parent_idx; };
        std::deque<Tok> q;
        for (auto& kv : c.map)                                          // init visit (tp.hpp:208-228)
          if (kv.second.vpi == k.I[0] && labels[kv.first] == k.P[0]) {
            token_source_map.insert({kv.first, false});
            for (uint64_t e = g->rowptr[kv.first]; e < g->rowptr[kv.first + 1]; ++e)
              q.push_back({g->col[e], kv.first, 0, k.I[0]});
          }
        while (!q.empty()) {
          const Tok t = q.front();
          q.pop_front();
          r->edges_processed++;
          const bool interior = k.C > t.itr;
          if (interior && forwarded[t.v].count(t.target)) continue;     // pre_visit :104-109
          auto f = c.map.find(t.v);
          if (f == c.map.end()) continue;                               // :111-114
          const uint32_t ni = t.itr + 1;
          if (ni >= k.P.size()) continue;
          if (!(labels[t.v] == k.P[ni] && f->second.vpi == k.I[ni] && t.parent_idx == k.I[ni - 1])) continue;  // :127-131, 157-164
          if (interior) {
            forwarded[t.v].insert(t.target);                            // :137-149
            for (uint64_t e = g->rowptr[t.v]; e < g->rowptr[t.v + 1]; ++e)   // visit :289-295
              q.push_back({g->col[e], t.target, ni, f->second.vpi});
          } else if (t.v == t.target && k.valid_cycle) {                // :250-262
            token_source_map[t.v] = true;
          }
        }
        for (auto& kv : token_source_map)                               // TP_ORIG post-processing (:583-629)
          if (!kv.second) { c.active[kv.first] = 0; c.map.erase(kv.first); nf = true; }
      }
      orc_row row{itr, 1, 0, c.map.size(), 0, 0.0};                     // "itr, TP, 0, count" (:664-666)
      r->rows.push_back(row);
    }
    ++itr;
    if ((int)itr >= max_it && nf) { r->hazards[4]++; break; }
  } while (nf);
  r->iterations = itr;
  r->search_seconds = now_s() - t_begin;
  r->inmap.assign(V, 0);
  r->T_arr.assign(V, 0);
  for (auto& kv : c.map) { r->inmap[kv.first] = 1; r->T_arr[kv.first] = (uint16_t)(1u << kv.second.vpi); }
  r->active = c.active;
  r->estate.assign(g->col.size(), 0);  // this path keeps no edge maps
  r->subgraphs.resize(P.cons.size());
  r->subgraph_width.assign(P.cons.size(), 0);
  return r;
}
