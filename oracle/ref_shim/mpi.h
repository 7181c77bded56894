// mpi.h — a world of ONE rank (see README.md in this directory).  Only what the pattern matching path calls.
#pragma once
#include <chrono>

typedef int MPI_Comm;
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0

inline int MPI_Init(int*, char***) { return MPI_SUCCESS; }
inline int MPI_Finalize() { return MPI_SUCCESS; }
inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int* s) { *s = 1; return MPI_SUCCESS; }
inline int MPI_Abort(MPI_Comm, int code) { std::exit(code); return MPI_SUCCESS; }
inline double MPI_Wtime() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
