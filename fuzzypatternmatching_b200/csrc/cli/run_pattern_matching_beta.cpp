// run_pattern_matching_beta — command-line twin of the reference driver
// (/root/reference/src/run_pattern_matching_beta.cpp) on top of the C ABI of libpmgpu.so.
//
// Same flags and usage text (beta.cpp:67-142): -i <graph base> -b <backup base>
// [-v vertex metadata] [-e edge metadata] -p <pattern base dir> -o <output base dir>
// [-x batch] -h.  `-i/-b` name a PMGRAPH container written by this repo's generate_rmat
// (or `rmat:<scale>:<gen_ranks>` to generate in place); `-b` restores the container to the
// `-i` location first, like distributed_db::transfer (beta.cpp:209-211) — without the
// reference's accidental truncation when only -i is given.  The loop
// below is the reference's do/while (beta.cpp:544-1351) spelled out over pm_lcc / pm_nlcc
// so that every step of the driver maps to one ABI call; it prints the same progress lines.
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <vector>

#include "../../../include/pmgpu.h"
#include "pm_container.hpp"

static void usage() {
  std::cerr << "Usage: -i <string> -p <string> -o <string>\n"
            << " -i <string>   - input graph base filename (required)\n"
            << " -b <string>   - backup graph base filename. If set, \"input\" graph will be deleted if it exists\n"
            << " -v <string>   - vertex metadata base filename (optional, Default is degree based metadata)\n"
            << " -e <string>   - edge metadata base filename (optional)\n"
            << " -p <string>   - pattern base directory (required)\n"
            << " -o <string>   - output base directory (required)\n"
            << " -x <int>      - Token Passing batch size (optional, Default/Max batch size is 1 , Min batch size is 1)\n"
            << " -t <int>      - first constraint index searched with Template Driven Search (Default is 4)\n"
            << " -n <int>      - ranks = GPUs (the reference's mpirun -np; Default is 1)\n"
            << " -h            - print help and exit\n\n";
}

static double now_s() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

#define CHECK(call)                                                     \
  do {                                                                  \
    if ((call) != 0) {                                                  \
      std::cerr << "Error: " << pm_last_error(ctx) << std::endl;        \
      return 1;                                                         \
    }                                                                   \
  } while (0)

struct Options {
  std::string graph_input, backup_graph_input, vertex_metadata_input, edge_metadata_input, pattern_input, result_output;
  int tds_from = 4;
  int n_ranks = 1;
};

static bool file_exists(const std::string& p) { return access(p.c_str(), R_OK) == 0; }

// One rank of the run (the reference's main is collective over the MPI ranks, beta.cpp:144-191): device `rank`,
// partition owner(v) = v mod n_ranks, per-rank result files *_<rank> (beta.cpp:504-535).  Progress lines come from
// rank 0 only, like the reference's `if (mpi_rank == 0)` guards.
static int run_rank(const Options& o, int rank, int n_ranks, const char* comm_id) {
  struct Quiet : std::streambuf { int overflow(int c) override { return c; } } quiet;
  std::ostream null_out(&quiet);
  std::ostream& out = rank == 0 ? std::cout : null_out;
  pm_ctx* ctx = nullptr;
  if (pm_create(&ctx, rank) != 0) { std::cerr << "Error: no CUDA device " << rank << " (this engine has no CPU path)." << std::endl; return 1; }
  if (n_ranks > 1) CHECK(pm_comm_init(ctx, rank, n_ranks, comm_id));
  out << "MPI Initialized With " << n_ranks << " Ranks." << std::endl;

  // ---- load graph (beta.cpp:200-244)
  out << "Loading Graph ... " << std::endl;
  uint64_t delegate_threshold = 1048576;
  if (o.graph_input.rfind("rmat:", 0) == 0) {
    unsigned long long scale = 0, gen = 0;
    char colon;
    std::istringstream ss(o.graph_input.substr(5));
    ss >> scale >> colon >> gen;
    CHECK(pm_graph_rmat(ctx, scale, gen ? gen : 4));
  } else {
    std::string err;
    if (rank == 0 && !o.backup_graph_input.empty() && o.backup_graph_input != o.graph_input &&
        !pmcli::copy_file(pmcli::container_path(o.backup_graph_input), pmcli::container_path(o.graph_input), err)) {
      std::cerr << "Error: " << err << std::endl;
      return 1;
    }
    pmcli::Container c;
    const std::string path = pmcli::container_path(!o.backup_graph_input.empty() && rank != 0 ? o.backup_graph_input : o.graph_input);
    if (!pmcli::read_container(path, c, err)) { std::cerr << "Error: " << err << std::endl; return 1; }
    delegate_threshold = c.delegate_threshold;
    if (n_ranks == 1) {
      CHECK(pm_graph_from_csr(ctx, c.n_vertices, c.rowptr.data(), c.col.data(), c.degree_multi.data()));
    } else {
      // the rows this rank owns: local row i = vertex i * n_ranks + rank (impl/delegate_partitioned_graph.ipp:1682-1697)
      std::vector<uint64_t> rowptr(1, 0), degm;
      std::vector<uint32_t> col;
      for (uint64_t v = rank; v < c.n_vertices; v += n_ranks) {
        col.insert(col.end(), c.col.begin() + c.rowptr[v], c.col.begin() + c.rowptr[v + 1]);
        rowptr.push_back(col.size());
        degm.push_back(c.degree_multi[v]);
      }
      if (degm.empty()) degm.push_back(0);
      CHECK(pm_graph_from_csr(ctx, c.n_vertices, rowptr.data(), col.data(), degm.data()));
    }
  }
  out << "Done Loading Graph." << std::endl;
  // several ranks: hubs are attributed to their controller ranks in the per-rank files, like a reference run with delegates
  if (n_ranks > 1) CHECK(pm_graph_set_delegate_threshold(ctx, delegate_threshold));
  uint64_t n_delegates = 0;
  pm_graph_num_delegates(ctx, &n_delegates);
  out << "Delegate threshold : " << delegate_threshold << ", delegates : " << n_delegates << std::endl;

  // ---- vertex data (beta.cpp:358-377), edge data (beta.cpp:379-400: loaded, never read by the search, :906)
  out << "Fuzzy Pattern Matching ... " << std::endl;
  double t0 = now_s();
  if (!o.vertex_metadata_input.empty()) {
    out << "Building distributed vertex data db ... " << std::endl;
    CHECK(pm_labels_from_files(ctx, o.vertex_metadata_input.c_str()));
    out << "Done building vertex data db." << std::endl;
  } else {
    CHECK(pm_labels_degree_log2(ctx));
  }
  out << "Fuzzy Pattern Matching Time | Vertex Data DB : " << now_s() - t0 << std::endl;
  if (!o.edge_metadata_input.empty() && rank == 0) {
    pm_graph_info_t gi;
    CHECK(pm_graph_info(ctx, &gi));
    char err[512];
    uint64_t n = 0;
    if (pm_io_check_edge_data(o.edge_metadata_input.c_str(), gi.n_vertices, &n, err, sizeof(err)) != 0) {
      std::cerr << "Error: " << err << std::endl;
      return 1;
    }
    out << "Edge Data DB : " << n << " records (not used by the search)" << std::endl;
  }

  // ---- the pattern set (beta.cpp:424): <pattern_dir>/0, /1, ... while present; the reference stops after 0 (a TODO)
  for (int ps = 0; ps == 0 || file_exists(o.pattern_input + "/" + std::to_string(ps) + "/pattern_edge"); ++ps) {
    out << "Setting up Pattern [" << ps << "] ... " << std::endl;
    CHECK(pm_pattern_load_dir(ctx, (o.pattern_input + "/" + std::to_string(ps)).c_str()));
    pm_pattern_info_t pi;
    CHECK(pm_pattern_info(ctx, &pi));
    out << "Fuzzy Pattern Matching | Searching Pattern [" << ps << "] : \ndiameter : " << pi.diameter << std::endl;

    // ---- the loop (beta.cpp:481-1351)
    CHECK(pm_state_reset(ctx));
    std::vector<pm_counts_t> counts(pi.diameter);
    bool global_init_step = true;
    int global_not_finished = 0;
    uint64_t global_itr_count = 0;
    const double pattern_time_start = now_s();
    do {
      global_not_finished = 0;
      const double itr_time_start = now_s();
      out << "Label Propagation ... " << std::endl;
      double lp0 = now_s();
      CHECK(pm_lcc(ctx, global_init_step, &global_not_finished, counts.data()));
      for (int k = 0; k < pi.diameter; ++k)
        out << "Label Propagation | Superstep #" << k << " | Synchronizing ... | Time : " << counts[k].seconds << std::endl;
      out << "Fuzzy Pattern Matching Time | Label Propagation : " << now_s() - lp0 << std::endl;
      global_init_step = false;
      out << "Fuzzy Pattern Matching | Global Finished Status : " << (global_not_finished ? "Continue" : "Stop") << std::endl;
      if (global_itr_count == 0) global_not_finished = 1;  // forced token passing (beta.cpp:686-688)
      if (global_not_finished) {
        global_not_finished = 0;
        for (int pl = 0; pl < pi.n_constraints; ++pl) {
          const bool do_tds_tp = o.tds_from >= 0 && pl >= o.tds_from;  // beta.cpp:762-767
          if (do_tds_tp) out << "Token Passing [" << pl << "] | Template Driven Search " << std::endl;
          int found = 0, deleted = 0;
          pm_counts_t tp;
          CHECK(pm_nlcc(ctx, pl, do_tds_tp ? PM_NLCC_TDS : PM_NLCC_NEM1, &found, &deleted, &tp));
          out << "Fuzzy Pattern Matching Time | Token Passing [" << pl << "] : " << tp.seconds << std::endl;
          out << "Token Passing [" << pl << "] | Found Pattern : " << (found ? "True" : "False") << std::endl;
          out << "Token Passing [" << pl << "] | Token Source Deleted Status : " << (deleted ? "Deleted" : "Not Deleted") << std::endl;
          if (deleted) global_not_finished = 1;
          // interleave token passing with label propagation (beta.cpp:1163-1197)
          pm_constraint_info_t ci;
          CHECK(pm_pattern_constraint_info(ctx, pl, &ci));
          if (deleted && ci.interleave_lcc) {
            double l0 = now_s();
            CHECK(pm_lcc(ctx, 0, &global_not_finished, counts.data()));
            out << "Fuzzy Pattern Matching Time | Label Propagation (Interleaved) : " << now_s() - l0 << std::endl;
          } else {
            out << "Fuzzy Pattern Matching | Skipping Label Propagation (Interleaved)." << std::endl;
          }
        }
      } else {
        out << "Fuzzy Pattern Matching | Skipping Token Passing." << std::endl;
      }
      out << "Fuzzy Pattern Matching | Global Finished Status : " << (global_not_finished ? "Continue" : "Stop") << std::endl;
      out << "Fuzzy Pattern Matching Time | Pattern [" << ps << "] | Iteration [" << global_itr_count
          << "] : " << now_s() - itr_time_start << std::endl;
      CHECK(pm_end_iteration(ctx, now_s() - itr_time_start));
      global_itr_count++;
    } while (global_not_finished);
    out << "Fuzzy Pattern Matching Time | Pattern [" << ps << "] : " << now_s() - pattern_time_start << std::endl;
    out << "Fuzzy Pattern Matching | Pattern [" << ps << "] | # Iterations : " << global_itr_count << std::endl;

    // ---- results (beta.cpp:1370-1425); like the reference, directories are never created
    CHECK(pm_write_results_ps(ctx, o.result_output.c_str(), ps));
  }
  pm_destroy(ctx);
  return 0;
}

int main(int argc, char** argv) {
  Options o;
  bool help = false;
  int required = 0;
  std::cout << "CMD Line :";
  for (int i = 0; i < argc; ++i) std::cout << " " << argv[i];
  std::cout << std::endl;
  int ch;
  while ((ch = getopt(argc, argv, "i:b:v:e:p:o:x:t:n:h")) != -1) {
    switch (ch) {
      case 'h': help = true; break;
      case 'i': o.graph_input = optarg; required |= 1; break;
      case 'b': o.backup_graph_input = optarg; break;
      case 'v': o.vertex_metadata_input = optarg; break;
      case 'e': o.edge_metadata_input = optarg; break;
      case 'p': o.pattern_input = optarg; required |= 2; break;
      case 'o': o.result_output = optarg; required |= 4; break;
      case 'x': break;  // parsed and unused by the reference as well (beta.cpp:125-130,187)
      case 't': o.tds_from = std::atoi(optarg); break;
      case 'n': o.n_ranks = std::atoi(optarg); break;
      default:
        std::cerr << "Unrecognized Option : " << (char)ch << ", Ignore." << std::endl;
        help = true;
        break;
    }
  }
  if (help || required != 7 || o.n_ranks < 1 || o.n_ranks > 8) { usage(); return -1; }
  if (o.n_ranks == 1) return run_rank(o, 0, 1, nullptr);

  // Several ranks: where the reference is started under mpirun, this driver starts one process per GPU itself
  // (fork before any CUDA call) and hands rank 0's communicator id to the others through pipes.
  std::cout.flush();
  std::vector<int> rd(o.n_ranks, -1), wr(o.n_ranks, -1);
  for (int r = 1; r < o.n_ranks; ++r) {
    int fd[2];
    if (pipe(fd) != 0) { std::cerr << "Error: pipe failed." << std::endl; return 1; }
    rd[r] = fd[0];
    wr[r] = fd[1];
  }
  std::vector<pid_t> kids;
  for (int r = 0; r < o.n_ranks; ++r) {
    const pid_t pid = fork();
    if (pid < 0) { std::cerr << "Error: fork failed." << std::endl; return 1; }
    if (pid == 0) {
      char id[PM_COMM_ID_BYTES];
      if (r == 0) {
        if (pm_comm_unique_id(id) != 0) { std::cerr << "Error: NCCL is not available." << std::endl; _exit(1); }
        for (int q = 1; q < o.n_ranks; ++q)
          if (write(wr[q], id, sizeof(id)) != (ssize_t)sizeof(id)) _exit(1);
      } else {
        size_t got = 0;
        while (got < sizeof(id)) {
          const ssize_t n = read(rd[r], id + got, sizeof(id) - got);
          if (n <= 0) _exit(1);
          got += (size_t)n;
        }
      }
      const int rc = run_rank(o, r, o.n_ranks, id);
      std::cout.flush();
      _exit(rc);
    }
    kids.push_back(pid);
  }
  int worst = 0;
  for (pid_t k : kids) {
    int st = 0;
    waitpid(k, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) worst = 1;
  }
  return worst;
}
