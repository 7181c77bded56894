// havoqgt/mpi.hpp — collectives over a world of one rank: every reduction returns its input.
#pragma once
#include <mpi.h>

#include <cstdlib>
#include <iostream>
#include <vector>

#define CHK_MPI(a)                                                                  \
  {                                                                                 \
    if ((a) != MPI_SUCCESS) {                                                       \
      std::cerr << "MPI call failed: " #a << std::endl;                             \
      std::exit(-1);                                                                \
    }                                                                               \
  }

namespace havoqgt {
namespace mpi {

template <typename T, typename Op>
T mpi_all_reduce(T in, Op, MPI_Comm) { return in; }

template <typename T, typename Op>
void mpi_all_reduce(std::vector<T>& in, std::vector<T>& out, Op, MPI_Comm) { out = in; }

template <typename Vec, typename Op>
void mpi_all_reduce_inplace(Vec&, Op, MPI_Comm) {}

template <typename T>
T mpi_bcast(T in, int, MPI_Comm) { return in; }

inline int mpi_comm_rank() { return 0; }
inline int mpi_comm_size() { return 1; }

}  // namespace mpi
}  // namespace havoqgt
