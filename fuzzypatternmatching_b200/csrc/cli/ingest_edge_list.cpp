// ingest_edge_list — command-line twin of /root/reference/src/ingest_edge_list.cpp for this engine.
// Same flags and usage text (ingest_edge_list.cpp:82-160): -o <output graph base> [-b backup] [-d delegate
// threshold] [-p] [-f] [-c] [-u treat as undirected] file ...; the files hold "source target [weight]" lines
// (include/havoqgt/parallel_edge_list_reader.hpp:236-262).  The graph is built on the GPU through the C ABI
// (pm_graph_from_slots) and written as a PMGRAPH container (pm_container.hpp) that run_pattern_matching_beta opens.
// The pattern matching path needs both directions of every edge: ingest with -u 1 unless the files list them.
#include <unistd.h>

#include <cstdlib>
#include <iostream>
#include <vector>

#include "../../../include/pmgpu.h"
#include "pm_container.hpp"

static void usage() {
  std::cerr << "Usage: -o <string> -d <int> [file ...]\n"
            << " -o <string>   - output graph base filename (required)\n"
            << " -b <string>   - backup graph base filename \n"
            << " -d <int>      - delegate threshold (Default is 1048576)\n"
            << " -h            - print help and exit\n"
            << " -p <int>      - number of Low & High partition passes (Default is 1)\n"
            << " -f <float>    - Gigabytes reserved per rank (Default is 0.25)\n"
            << " -c <int>      - Edge partitioning chunk size (Defulat is 8192)\n"
            << " -u <bool>     - Treat edgelist as undirected (Default is 0)\n"
            << "[file ...] - list of edge list files to ingest\n\n";
}

int main(int argc, char** argv) {
  uint64_t threshold = 1048576;
  std::string out, backup;
  bool have_out = false, help = false, undirected = false;
  std::cout << "CMD line:";
  for (int i = 0; i < argc; ++i) std::cout << " " << argv[i];
  std::cout << std::endl;
  int ch;
  while ((ch = getopt(argc, argv, "o:d:p:f:c:b:u:h")) != -1) {
    switch (ch) {
      case 'h': help = true; break;
      case 'd': threshold = std::strtoull(optarg, nullptr, 10); break;
      case 'o': out = optarg; have_out = true; break;
      case 'b': backup = optarg; break;
      case 'p': case 'f': case 'c': break;  // knobs of the reference's mmap store and partitioning passes: nothing to tune here
      case 'u': undirected = std::atoi(optarg) != 0; break;
      default:
        std::cerr << "Unrecognized option: " << (char)ch << ", ignore." << std::endl;
        help = true;
        break;
    }
  }
  if (help || !have_out) { usage(); return -1; }
  std::vector<const char*> files;
  for (int i = optind; i < argc; ++i) files.push_back(argv[i]);
  std::cout << "Ingesting graph from " << files.size() << " files." << std::endl;

  char err[512];
  uint64_t n_vertices = 0, n_slots = 0;
  if (pm_io_read_edge_lists(files.data(), (int)files.size(), undirected, &n_vertices, &n_slots, nullptr, nullptr, err, sizeof(err)) != 0) {
    std::cerr << "Error: " << err << std::endl;
    return 1;
  }
  std::vector<uint32_t> src(n_slots ? n_slots : 1), dst(n_slots ? n_slots : 1);
  if (pm_io_read_edge_lists(files.data(), (int)files.size(), undirected, &n_vertices, &n_slots, src.data(), dst.data(), err, sizeof(err)) != 0) {
    std::cerr << "Error: " << err << std::endl;
    return 1;
  }
  if (n_vertices == 0) { std::cerr << "Error: no edges read." << std::endl; return 1; }

  pm_ctx* ctx = nullptr;
  if (pm_create(&ctx, 0) != 0) { std::cerr << "Error: no CUDA device (this engine has no CPU path)." << std::endl; return 1; }
  if (pm_graph_from_slots(ctx, n_vertices, n_slots, src.data(), dst.data()) != 0) {
    std::cerr << "Error: " << pm_last_error(ctx) << std::endl;
    return 1;
  }
  pm_graph_info_t gi;
  pm_graph_info(ctx, &gi);
  pmcli::Container c;
  c.n_vertices = gi.n_vertices; c.n_slots = gi.n_slots; c.n_slots_multi = gi.n_slots_multi;
  c.scale = 0; c.gen_ranks = 0; c.delegate_threshold = threshold;
  c.rowptr.resize(gi.n_vertices + 1);
  c.degree_multi.resize(gi.n_vertices);
  c.col.resize(gi.n_slots);
  if (pm_graph_get_csr(ctx, c.rowptr.data(), c.col.data()) != 0 || pm_graph_get_degree(ctx, c.degree_multi.data()) != 0) {
    std::cerr << "Error: " << pm_last_error(ctx) << std::endl;
    return 1;
  }
  std::cout << "Graph Ready, Calculating Stats. " << std::endl;
  std::cout << "Vertices = " << gi.n_vertices << ", directed edges = " << gi.n_slots_multi << ", distinct = " << gi.n_slots
            << "\nMax Degree = " << gi.max_degree << std::endl;
  std::string e2;
  if (!pmcli::write_container(pmcli::container_path(out), c, e2)) { std::cerr << "Error: " << e2 << std::endl; return 1; }
  if (!backup.empty() && !pmcli::copy_file(pmcli::container_path(out), pmcli::container_path(backup), e2)) {
    std::cerr << "Error: " << e2 << std::endl;
    return 1;
  }
  pm_destroy(ctx);
  return 0;
}
