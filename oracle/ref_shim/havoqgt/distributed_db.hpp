// havoqgt/distributed_db.hpp — "opens" a graph: the file named by -i is a text slot list (see
// delegate_partitioned_graph.hpp in this directory) instead of a Boost.Interprocess image.
#pragma once
#include <havoqgt/delegate_partitioned_graph.hpp>

#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <type_traits>
#include <utility>

namespace havoqgt {

struct db_open {};
struct db_create {};

class slot_file_segment_manager;
typedef mpi::delegate_partitioned_graph<slot_file_segment_manager> slot_file_graph;

class slot_file_segment_manager {
 public:
  // find<T>("graph_obj") -> the graph; find<T>("graph_edge_data_obj") -> one zero value per slot (an image ingested
  // without weights; the pattern matching path never reads the values, run_pattern_matching.cpp only asserts that the
  // object exists); any other name -> nothing stored
  template <typename T>
  std::pair<T*, size_t> find(const char* name) {
    if constexpr (std::is_same<T, slot_file_graph>::value) {
      if (std::strcmp(name, "graph_obj") == 0) return std::make_pair(m_graph, size_t(1));
    } else {
      if (std::strcmp(name, "graph_edge_data_obj") == 0) return std::make_pair(edge_data_object<T>(), size_t(1));
    }
    return std::make_pair(static_cast<T*>(nullptr), size_t(0));
  }
  slot_file_graph* m_graph = nullptr;

 private:
  template <typename T>
  T* edge_data_object() {
    static T* obj = nullptr;  // one graph per process
    if (!obj) {
      obj = new T(*m_graph);
      obj->reset(typename T::value_type());
    }
    return obj;
  }

 public:
};

class distributed_db {
 public:
  typedef slot_file_segment_manager segment_manager_type;

  distributed_db(db_open, const char* path) {
    if (!m_graph.load_slots(path)) {
      std::cerr << "cannot read the slot list " << path << std::endl;
      std::exit(-1);
    }
    m_manager.m_graph = &m_graph;
  }
  segment_manager_type* get_segment_manager() { return &m_manager; }
  static void transfer(const char*, const char*) {}

 private:
  slot_file_graph m_graph;
  segment_manager_type m_manager;
};

}  // namespace havoqgt
