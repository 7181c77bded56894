// run_pattern_matching_beta — command-line twin of the reference driver
// (/root/reference/src/run_pattern_matching_beta.cpp) on top of the C ABI of libpmgpu.so.
//
// Same flags and usage text (beta.cpp:67-142): -i <graph base> -b <backup base>
// [-v vertex metadata] [-e edge metadata] -p <pattern base dir> -o <output base dir>
// [-x batch] -h.  `-i/-b` name a PMGRAPH1 container written by this repo's generate_rmat
// (or `rmat:<scale>:<gen_ranks>` to generate in place); `-b` restores the container to the
// `-i` location first, like distributed_db::transfer (beta.cpp:209-211) — without the
// reference's accidental truncation when only -i is given.  The loop
// below is the reference's do/while (beta.cpp:544-1351) spelled out over pm_lcc / pm_nlcc
// so that every step of the driver maps to one ABI call; it prints the same progress lines.
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

int main(int argc, char** argv) {
  std::string graph_input, backup_graph_input, vertex_metadata_input, edge_metadata_input, pattern_input, result_output;
  int tds_from = 4;
  bool help = false;
  int required = 0;
  std::cout << "CMD Line :";
  for (int i = 0; i < argc; ++i) std::cout << " " << argv[i];
  std::cout << std::endl;
  int ch;
  while ((ch = getopt(argc, argv, "i:b:v:e:p:o:x:t:h")) != -1) {
    switch (ch) {
      case 'h': help = true; break;
      case 'i': graph_input = optarg; required |= 1; break;
      case 'b': backup_graph_input = optarg; break;
      case 'v': vertex_metadata_input = optarg; break;
      case 'e': edge_metadata_input = optarg; break;
      case 'p': pattern_input = optarg; required |= 2; break;
      case 'o': result_output = optarg; required |= 4; break;
      case 'x': break;  // parsed and unused by the reference as well (beta.cpp:125-130,187)
      case 't': tds_from = std::atoi(optarg); break;
      default:
        std::cerr << "Unrecognized Option : " << (char)ch << ", Ignore." << std::endl;
        help = true;
        break;
    }
  }
  if (help || required != 7) { usage(); return -1; }

  pm_ctx* ctx = nullptr;
  if (pm_create(&ctx, 0) != 0) { std::cerr << "Error: no CUDA device (this engine has no CPU path)." << std::endl; return 1; }
  std::cout << "MPI Initialized With 1 Ranks." << std::endl;

  // ---- load graph (beta.cpp:200-244)
  std::cout << "Loading Graph ... " << std::endl;
  if (graph_input.rfind("rmat:", 0) == 0) {
    unsigned long long scale = 0, gen = 0;
    char colon;
    std::istringstream ss(graph_input.substr(5));
    ss >> scale >> colon >> gen;
    CHECK(pm_graph_rmat(ctx, scale, gen ? gen : 4));
  } else {
    std::string err;
    if (!backup_graph_input.empty() && backup_graph_input != graph_input &&
        !pmcli::copy_file(pmcli::container_path(backup_graph_input), pmcli::container_path(graph_input), err)) {
      std::cerr << "Error: " << err << std::endl;
      return 1;
    }
    pmcli::Container c;
    if (!pmcli::read_container(pmcli::container_path(graph_input), c, err)) { std::cerr << "Error: " << err << std::endl; return 1; }
    CHECK(pm_graph_from_csr(ctx, c.n_vertices, c.rowptr.data(), c.col.data(), c.degree_multi.data()));
  }
  std::cout << "Done Loading Graph." << std::endl;

  // ---- vertex data (beta.cpp:358-377)
  std::cout << "Fuzzy Pattern Matching ... " << std::endl;
  double t0 = now_s();
  if (!vertex_metadata_input.empty()) {
    // "vertex label" lines (include/havoqgt/vertex_data_db.hpp:169-194)
    pm_graph_info_t gi;
    CHECK(pm_graph_info(ctx, &gi));
    std::vector<uint64_t> labels(gi.n_vertices, 0);
    std::ifstream f(vertex_metadata_input);
    if (!f) { std::cerr << "Error: cannot open " << vertex_metadata_input << std::endl; return 1; }
    unsigned long long v, l;
    while (f >> v >> l) if (v < gi.n_vertices) labels[v] = l;
    CHECK(pm_labels_set(ctx, labels.data()));
  } else {
    CHECK(pm_labels_degree_log2(ctx));
  }
  std::cout << "Fuzzy Pattern Matching Time | Vertex Data DB : " << now_s() - t0 << std::endl;

  // ---- pattern (beta.cpp:424-479): only <pattern_dir>/0 is read
  const int ps = 0;
  std::cout << "Setting up Pattern [" << ps << "] ... " << std::endl;
  CHECK(pm_pattern_load_dir(ctx, (pattern_input + "/" + std::to_string(ps)).c_str()));
  pm_pattern_info_t pi;
  CHECK(pm_pattern_info(ctx, &pi));
  std::cout << "Fuzzy Pattern Matching | Searching Pattern [" << ps << "] : \ndiameter : " << pi.diameter << std::endl;

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
    std::cout << "Label Propagation ... " << std::endl;
    double lp0 = now_s();
    CHECK(pm_lcc(ctx, global_init_step, &global_not_finished, counts.data()));
    for (int k = 0; k < pi.diameter; ++k)
      std::cout << "Label Propagation | Superstep #" << k << " | Synchronizing ... | Time : " << counts[k].seconds << std::endl;
    std::cout << "Fuzzy Pattern Matching Time | Label Propagation : " << now_s() - lp0 << std::endl;
    global_init_step = false;
    std::cout << "Fuzzy Pattern Matching | Global Finished Status : " << (global_not_finished ? "Continue" : "Stop") << std::endl;
    if (global_itr_count == 0) global_not_finished = 1;  // forced token passing (beta.cpp:686-688)
    if (global_not_finished) {
      global_not_finished = 0;
      for (int pl = 0; pl < pi.n_constraints; ++pl) {
        const bool do_tds_tp = tds_from >= 0 && pl >= tds_from;  // beta.cpp:762-767
        if (do_tds_tp) std::cout << "Token Passing [" << pl << "] | Template Driven Search " << std::endl;
        int found = 0, deleted = 0;
        pm_counts_t tp;
        CHECK(pm_nlcc(ctx, pl, do_tds_tp ? PM_NLCC_TDS : PM_NLCC_NEM1, &found, &deleted, &tp));
        std::cout << "Fuzzy Pattern Matching Time | Token Passing [" << pl << "] : " << tp.seconds << std::endl;
        std::cout << "Token Passing [" << pl << "] | Found Pattern : " << (found ? "True" : "False") << std::endl;
        std::cout << "Token Passing [" << pl << "] | Token Source Deleted Status : " << (deleted ? "Deleted" : "Not Deleted") << std::endl;
        if (deleted) global_not_finished = 1;
        // interleave token passing with label propagation (beta.cpp:1163-1197)
        pm_constraint_info_t ci;
        CHECK(pm_pattern_constraint_info(ctx, pl, &ci));
        if (deleted && ci.interleave_lcc) {
          double l0 = now_s();
          CHECK(pm_lcc(ctx, 0, &global_not_finished, counts.data()));
          std::cout << "Fuzzy Pattern Matching Time | Label Propagation (Interleaved) : " << now_s() - l0 << std::endl;
        } else {
          std::cout << "Fuzzy Pattern Matching | Skipping Label Propagation (Interleaved)." << std::endl;
        }
      }
    } else {
      std::cout << "Fuzzy Pattern Matching | Skipping Token Passing." << std::endl;
    }
    std::cout << "Fuzzy Pattern Matching | Global Finished Status : " << (global_not_finished ? "Continue" : "Stop") << std::endl;
    std::cout << "Fuzzy Pattern Matching Time | Pattern [" << ps << "] | Iteration [" << global_itr_count
              << "] : " << now_s() - itr_time_start << std::endl;
    CHECK(pm_end_iteration(ctx, now_s() - itr_time_start));
    global_itr_count++;
  } while (global_not_finished);
  std::cout << "Fuzzy Pattern Matching Time | Pattern [" << ps << "] : " << now_s() - pattern_time_start << std::endl;
  std::cout << "Fuzzy Pattern Matching | Pattern [" << ps << "] | # Iterations : " << global_itr_count << std::endl;

  // ---- results (beta.cpp:1370-1425); like the reference, directories are never created
  CHECK(pm_write_results(ctx, result_output.c_str()));
  pm_destroy(ctx);
  return 0;
}
