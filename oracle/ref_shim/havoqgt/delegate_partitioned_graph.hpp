// havoqgt/delegate_partitioned_graph.hpp — an in-memory graph with the interface the pattern matching path uses,
// for ONE rank and without delegates (see README.md in oracle/ref_shim).  Written from scratch: the real class
// (include/havoqgt/delegate_partitioned_graph.hpp + impl/*.hpp of the reference) lives in a Boost.Interprocess segment and
// is partitioned over MPI ranks; this one is a CSR of the directed slots of a text file.
//   * vertex v is owned by rank 0 with local id v: label_to_locator / locator_to_label are the identity
//   * edges_begin(v) .. edges_end(v) walk the slots of v in file order (a multigraph keeps its duplicates)
//   * degree(v) = number of slots of v
//   * vertex_data<T> / edge_data<T>: one value per vertex / slot (impl/vertex_data.hpp:63-127 of the reference for the
//     member functions the path calls: operator[], reset, clear, all_min_reduce, all_max_reduce — reductions over one rank)
#pragma once
#include <havoqgt/environment.hpp>
#include <havoqgt/mpi.hpp>

#include <boost/interprocess/managed_heap_memory.hpp>

#include <cassert>
#include <cstdint>
#include <deque>
#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <string>
#include <type_traits>
#include <vector>

namespace havoqgt {
namespace mpi {

template <typename SegmentManager>
class delegate_partitioned_graph {
 public:
  class vertex_locator {
   public:
    vertex_locator() : m_id(~0ull) {}
    explicit vertex_locator(uint64_t id) : m_id(id) {}
    bool is_valid() const { return m_id != ~0ull; }
    bool is_delegate() const { return false; }
    bool is_delegate_master() const { return false; }
    uint32_t get_bcast() const { return 0; }
    void set_bcast(uint32_t) {}
    bool is_intercept() const { return false; }
    void set_intercept(uint32_t) {}
    uint32_t owner() const { return 0; }
    uint64_t local_id() const { return m_id; }
    size_t hash() const { return std::hash<uint64_t>()(m_id); }
    bool is_equal(const vertex_locator x) const { return m_id == x.m_id; }
    friend bool operator==(const vertex_locator& a, const vertex_locator& b) { return a.m_id == b.m_id; }
    friend bool operator!=(const vertex_locator& a, const vertex_locator& b) { return a.m_id != b.m_id; }
    friend bool operator<(const vertex_locator& a, const vertex_locator& b) { return a.m_id < b.m_id; }
    friend bool operator>(const vertex_locator& a, const vertex_locator& b) { return a.m_id > b.m_id; }

   private:
    uint64_t m_id;
  };

  class vertex_iterator {
   public:
    vertex_iterator() : m_at(0) {}
    explicit vertex_iterator(uint64_t at) : m_at(at) {}
    vertex_locator operator*() const { return vertex_locator(m_at); }
    vertex_iterator& operator++() { ++m_at; return *this; }
    vertex_iterator operator++(int) { vertex_iterator t = *this; ++m_at; return t; }
    friend bool operator==(const vertex_iterator& a, const vertex_iterator& b) { return a.m_at == b.m_at; }
    friend bool operator!=(const vertex_iterator& a, const vertex_iterator& b) { return a.m_at != b.m_at; }

   private:
    uint64_t m_at;
  };
  typedef vertex_iterator controller_iterator;  // no delegates: the controller and delegate ranges are empty

  class edge_iterator {
   public:
    edge_iterator() : m_graph(nullptr), m_source(0), m_at(0) {}
    edge_iterator(const delegate_partitioned_graph* g, uint64_t source, uint64_t at) : m_graph(g), m_source(source), m_at(at) {}
    vertex_locator source() const { return vertex_locator(m_source); }
    vertex_locator target() const { return vertex_locator(m_graph->m_targets[m_at]); }
    uint64_t slot() const { return m_at; }
    edge_iterator& operator++() { ++m_at; return *this; }
    edge_iterator operator++(int) { edge_iterator t = *this; ++m_at; return t; }
    friend bool operator==(const edge_iterator& a, const edge_iterator& b) { return a.m_at == b.m_at; }
    friend bool operator!=(const edge_iterator& a, const edge_iterator& b) { return a.m_at != b.m_at; }

   private:
    const delegate_partitioned_graph* m_graph;
    uint64_t m_source, m_at;
  };

  template <typename T, typename Allocator = std::allocator<T>>
  class vertex_data {
   public:
    typedef T value_type;
    vertex_data() {}
    explicit vertex_data(const delegate_partitioned_graph& g, Allocator = Allocator()) : m_data(g.num_local_vertices()) {}
    T& operator[](const vertex_locator& v) { assert(v.local_id() < m_data.size()); return m_data[v.local_id()]; }
    const T& operator[](const vertex_locator& v) const { assert(v.local_id() < m_data.size()); return m_data[v.local_id()]; }
    void reset(const T& r) { for (auto& x : m_data) x = r; }
    void clear() { for (auto& x : m_data) x.clear(); }
    // reductions over the copies of the delegates on the ranks: none here
    void all_reduce() {}
    void all_max_reduce() {}
    void all_min_reduce() {}
    size_t size() const { return m_data.size(); }

   private:
    // std::deque<bool> hands out real references (std::vector<bool> would not)
    typename std::conditional<std::is_same<T, bool>::value, std::deque<bool>, std::vector<T>>::type m_data;
  };

  template <typename T, typename Allocator = std::allocator<T>>
  class edge_data {
   public:
    typedef T value_type;
    edge_data() {}
    explicit edge_data(const delegate_partitioned_graph& g, Allocator = Allocator()) : m_data(g.m_targets.size()) {}
    T& operator[](const edge_iterator& e) { return m_data[e.slot()]; }
    const T& operator[](const edge_iterator& e) const { return m_data[e.slot()]; }
    void reset(const T& r) { for (auto& x : m_data) x = r; }
    size_t size() const { return m_data.size(); }

   private:
    std::vector<T> m_data;
  };

  delegate_partitioned_graph() {}

  // text file: first line the number of vertices, then one directed slot "source target" per line (both directions of an
  // undirected edge are listed, as the reference's ingest with -u 1 stores them)
  bool load_slots(const std::string& path) {
    std::ifstream f(path);
    if (!f) return false;
    uint64_t n = 0, s = 0, t = 0;
    if (!(f >> n)) return false;
    std::vector<uint64_t> src, dst;
    while (f >> s >> t) {
      if (s >= n || t >= n) return false;
      src.push_back(s);
      dst.push_back(t);
    }
    m_offsets.assign(n + 1, 0);
    for (uint64_t x : src) m_offsets[x + 1]++;
    for (uint64_t v = 0; v < n; ++v) m_offsets[v + 1] += m_offsets[v];
    m_targets.resize(src.size());
    std::vector<uint64_t> at(m_offsets.begin(), m_offsets.end() - 1);
    for (size_t i = 0; i < src.size(); ++i) m_targets[at[src[i]]++] = dst[i];
    return true;
  }

  vertex_iterator vertices_begin() const { return vertex_iterator(0); }
  vertex_iterator vertices_end() const { return vertex_iterator(num_local_vertices()); }
  vertex_iterator delegate_vertices_begin() const { return vertex_iterator(0); }
  vertex_iterator delegate_vertices_end() const { return vertex_iterator(0); }
  controller_iterator controller_begin() const { return controller_iterator(0); }
  controller_iterator controller_end() const { return controller_iterator(0); }

  edge_iterator edges_begin(vertex_locator v) const { return edge_iterator(this, v.local_id(), m_offsets[v.local_id()]); }
  edge_iterator edges_end(vertex_locator v) const { return edge_iterator(this, v.local_id(), m_offsets[v.local_id() + 1]); }
  uint64_t degree(vertex_locator v) const { return m_offsets[v.local_id() + 1] - m_offsets[v.local_id()]; }
  uint64_t local_degree(vertex_locator v) const { return degree(v); }

  vertex_locator label_to_locator(uint64_t label) const { return vertex_locator(label); }
  uint64_t locator_to_label(vertex_locator v) const { return v.local_id(); }
  uint32_t master(const vertex_locator&) const { return 0; }

  uint64_t num_local_vertices() const { return m_offsets.empty() ? 0 : m_offsets.size() - 1; }
  uint64_t max_global_vertex_id() const { return num_local_vertices() ? num_local_vertices() - 1 : 0; }
  uint64_t max_local_vertex_id() const { return max_global_vertex_id(); }
  size_t num_delegates() const { return 0; }
  void print_graph_statistics() const {
    std::cout << "vertices " << num_local_vertices() << " directed slots " << m_targets.size() << std::endl;
  }

 private:
  std::vector<uint64_t> m_offsets, m_targets;
};

}  // namespace mpi
}  // namespace havoqgt
