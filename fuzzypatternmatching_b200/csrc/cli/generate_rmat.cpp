// generate_rmat — command-line twin of /root/reference/src/generate_rmat.cpp for this engine.
// Same flags (-s -d -o -b -p -f -c, generate_rmat.cpp:78-150); the number of generating
// ranks, which the reference takes from the MPI world size and which is part of the
// graph's identity (generate_rmat.cpp:201-205), is the extra flag -r (default 4, the
// README's `--ntasks-per-node=4`).  The graph is generated and built on the GPU through
// the C ABI and written as a PMGRAPH1 container (see pm_container.hpp).
#include <unistd.h>

#include <cstdlib>
#include <iostream>

#include "../../../include/pmgpu.h"
#include "pm_container.hpp"

static void usage() {
  std::cerr << "Usage: -s <int> -d <int> -o <string>\n"
            << " -s <int>      - RMAT graph Scale (default 17)\n"
            << " -d <int>      - delegate threshold (Default is 1048576)\n"
            << " -o <string>   - output graph base filename\n"
            << " -b <string>   - backup graph base filename \n"
            << " -p <int>      - number of Low & High partition passes (Default is 1)\n"
            << " -f <float>    - Gigabytes reserved per rank (Default is 0.25)\n"
            << " -c <int>      - Edge partitioning chunk size (Defulat is 8192)\n"
            << " -r <int>      - generating ranks (the reference's MPI world size; Default is 4)\n"
            << " -h            - print help and exit\n\n";
}

int main(int argc, char** argv) {
  uint64_t scale = 17, threshold = 1048576, gen_ranks = 4;
  std::string out, backup;
  bool have_out = false, help = false;
  int ch;
  while ((ch = getopt(argc, argv, "s:d:o:b:p:f:c:r:h")) != -1) {
    switch (ch) {
      case 's': scale = std::strtoull(optarg, nullptr, 10); break;
      case 'd': threshold = std::strtoull(optarg, nullptr, 10); break;
      case 'o': out = optarg; have_out = true; break;
      case 'b': backup = optarg; break;
      case 'r': gen_ranks = std::strtoull(optarg, nullptr, 10); break;
      case 'p': case 'f': case 'c': break;  // knobs of the reference's mmap store and partitioning passes: nothing to tune here
      case 'h': help = true; break;
      default: help = true; break;
    }
  }
  if (help || !have_out) { usage(); return -1; }
  std::cout << "Building Graph500\nBuilding graph Scale: " << scale << "\nGenerating ranks = " << gen_ranks
            << "\nFile name = " << out << std::endl;
  pm_ctx* ctx = nullptr;
  if (pm_create(&ctx, 0) != 0) { std::cerr << "Error: no CUDA device (this engine has no CPU path)." << std::endl; return 1; }
  if (pm_graph_rmat(ctx, scale, gen_ranks) != 0) { std::cerr << "Error: " << pm_last_error(ctx) << std::endl; return 1; }
  pm_graph_info_t gi;
  pm_graph_info(ctx, &gi);
  pmcli::Container c;
  c.n_vertices = gi.n_vertices; c.n_slots = gi.n_slots; c.n_slots_multi = gi.n_slots_multi;
  c.scale = scale; c.gen_ranks = gen_ranks;
  c.delegate_threshold = threshold;  // kept with the graph: run_pattern_matching_beta attributes hubs to their controllers
  c.rowptr.resize(gi.n_vertices + 1);
  c.degree_multi.resize(gi.n_vertices);
  c.col.resize(gi.n_slots);
  if (pm_graph_get_csr(ctx, c.rowptr.data(), c.col.data()) != 0 || pm_graph_get_degree(ctx, c.degree_multi.data()) != 0) {
    std::cerr << "Error: " << pm_last_error(ctx) << std::endl;
    return 1;
  }
  std::cout << "Graph Ready, Calculating Stats. " << std::endl;
  std::cout << "Max Degree = " << gi.max_degree << std::endl;
  {
    uint64_t hubs = 0;
    for (uint64_t d : c.degree_multi) hubs += d >= threshold;
    std::cout << "Delegate threshold = " << threshold << ", delegates = " << hubs << std::endl;
  }
  std::string err;
  if (!pmcli::write_container(pmcli::container_path(out), c, err)) { std::cerr << "Error: " << err << std::endl; return 1; }
  if (!backup.empty() && !pmcli::copy_file(pmcli::container_path(out), pmcli::container_path(backup), err)) {
    std::cerr << "Error: " << err << std::endl;
    return 1;
  }
  pm_destroy(ctx);
  return 0;
}
