// rmat_edge_dump — prints the directed pairs the reference's OWN R-MAT generator (include/havoqgt/rmat_edge_generator.hpp and
// include/havoqgt/detail/hash.hpp, compiled from /root/reference; Boost.Random through ref_shim/boost/random.hpp) yields for one
// generating rank, constructed exactly as src/generate_rmat.cpp:202-205 constructs it: seed 5489 + 3 * rank,
// 16 * 2^scale / ranks edges, a b c d = .57 .19 .19 .05, scramble on, undirected on.
// Test infrastructure (tests/test_oracle_vs_reference.py pins the oracle's R-MAT stream and hash_nbits with it).
//   usage: rmat_edge_dump <scale> <rank> <ranks> [max generated edges]        (-ffp-contract=off, like the oracle)
//          rmat_edge_dump hash <n bits> <x> ...      prints detail::hash_nbits(x, n) of every x
#include <cassert>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <iterator>
#include <sstream>
#include <string>

#include <havoqgt/rmat_edge_generator.hpp>

int main(int argc, char** argv) {
  if (argc >= 3 && std::string(argv[1]) == "hash") {
    const int n = std::atoi(argv[2]);
    for (int i = 3; i < argc; ++i) std::cout << havoqgt::detail::hash_nbits(std::strtoull(argv[i], nullptr, 10), n) << "\n";
    return 0;
  }
  if (argc < 4) return 2;
  const uint64_t scale = std::strtoull(argv[1], nullptr, 10), rank = std::strtoull(argv[2], nullptr, 10),
                 ranks = std::strtoull(argv[3], nullptr, 10);
  const uint64_t num_vertices = uint64_t(1) << scale;
  uint64_t num_edges_per_rank = num_vertices * 16 / ranks;  // generate_rmat.cpp:201
  if (argc > 4) num_edges_per_rank = std::min<uint64_t>(num_edges_per_rank, std::strtoull(argv[4], nullptr, 10));
  havoqgt::rmat_edge_generator rmat(uint64_t(5489) + rank * 3ULL, scale, num_edges_per_rank, 0.57, 0.19, 0.19, 0.05, true, true);
  std::ostringstream out;
  uint64_t n = 0;
  for (auto it = rmat.begin(); it != rmat.end(); ++it, ++n) out << (*it).first << " " << (*it).second << "\n";
  std::cout << rmat.max_vertex_id() << " " << n << "\n" << out.str();
  return 0;
}
