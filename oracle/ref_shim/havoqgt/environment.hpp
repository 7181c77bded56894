// havoqgt/environment.hpp — the process environment of a single rank.
#pragma once
#include <havoqgt/mpi.hpp>

#include <unistd.h>  // getopt: the real header reaches it through its own includes

#include <cmath>
#include <iostream>

namespace havoqgt {

class single_rank_comm {
 public:
  int rank() const { return 0; }
  int size() const { return 1; }
  MPI_Comm comm() const { return MPI_COMM_WORLD; }
  void barrier() const {}
};

class environment {
 public:
  const single_rank_comm& world_comm() const { return m_comm; }
  const single_rank_comm& node_local_comm() const { return m_comm; }
  const single_rank_comm& node_offset_comm() const { return m_comm; }
  void print() const { std::cout << "single-rank runtime (oracle/ref_shim)" << std::endl; }

 private:
  single_rank_comm m_comm;
};

inline environment& get_environment() {
  static environment e;
  return e;
}
inline environment* havoqgt_env() { return &get_environment(); }
inline void havoqgt_init(int*, char***) {}
inline void havoqgt_finalize() {}

}  // namespace havoqgt
