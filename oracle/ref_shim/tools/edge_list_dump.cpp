// edge_list_dump — prints what the reference's OWN edge list reader (include/havoqgt/parallel_edge_list_reader.hpp, compiled
// from /root/reference over the single-rank runtime stand-in) hands the graph constructor of src/ingest_edge_list.cpp:
//   line 1: "<max vertex id> <has edge data> <edge count>", then one "<source> <target>" line per iterated edge, in
//   iteration order (with -u 1: every edge followed by its reverse, :126-150).
// Test infrastructure (tests/test_oracle_vs_reference.py pins csrc/pm_io.hpp::read_edge_lists with it).
//   usage: edge_list_dump <undirected 0|1> file...
#include <cassert>
#include <cstdint>
#include <iostream>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include <havoqgt/environment.hpp>
#ifndef HAVOQGT_ERROR_MSG
#define HAVOQGT_ERROR_MSG(msg) do { std::cerr << "ERROR: " << msg << std::endl; } while (0)
#endif
#include <havoqgt/parallel_edge_list_reader.hpp>

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  const bool undirected = std::string(argv[1]) == "1";
  std::vector<std::string> files(argv + 2, argv + argc);
  typedef uint8_t edge_data_type;  // src/ingest_edge_list.cpp
  havoqgt::parallel_edge_list_reader<edge_data_type> reader(files, undirected);
  uint64_t n = 0;
  std::ostringstream body;
  for (auto it = reader.begin(); it != reader.end(); ++it) {
    body << std::get<0>(*it) << " " << std::get<1>(*it) << "\n";
    ++n;
  }
  std::cout << reader.max_vertex_id() << " " << (reader.has_edge_data() ? 1 : 0) << " " << n << "\n" << body.str();
  return 0;
}
