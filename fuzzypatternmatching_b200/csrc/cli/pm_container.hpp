// pm_container.hpp — on-disk graph container of this engine.
//
// The reference keeps its graph as a Boost.Interprocess managed_mapped_file image
// (`<base>_<rank>_of_<size>`, include/havoqgt/distributed_db.hpp:353-359) whose layout
// depends on the Boost version and cannot be read without it, so `-i/-o` name a plain
// little-endian container instead:
//   u64 magic "PMGRAPH2", u64 n_vertices, u64 n_slots, u64 n_slots_multi, u64 scale, u64 gen_ranks,
//   u64 delegate_threshold (the -d of generate_rmat / ingest_edge_list; "PMGRAPH1" files lack this word: 1048576)
//   u64 rowptr[n_vertices + 1], u64 degree_multi[n_vertices], u32 col[n_slots]
#pragma once
#include <stdint.h>

#include <cstdio>
#include <string>
#include <vector>

namespace pmcli {

static const uint64_t kMagic1 = 0x3148504152474d50ull;  // "PMGRAPH1"
static const uint64_t kMagic = 0x3248504152474d50ull;   // "PMGRAPH2"

struct Container {
  uint64_t n_vertices = 0, n_slots = 0, n_slots_multi = 0, scale = 0, gen_ranks = 0, delegate_threshold = 1048576;
  std::vector<uint64_t> rowptr, degree_multi;
  std::vector<uint32_t> col;
};

inline std::string container_path(const std::string& base) { return base + "_0_of_1.pmg"; }

inline bool write_container(const std::string& path, const Container& c, std::string& err) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) { err = "cannot create " + path; return false; }
  uint64_t hdr[7] = {kMagic, c.n_vertices, c.n_slots, c.n_slots_multi, c.scale, c.gen_ranks, c.delegate_threshold};
  bool ok = std::fwrite(hdr, 8, 7, f) == 7 &&
            std::fwrite(c.rowptr.data(), 8, c.rowptr.size(), f) == c.rowptr.size() &&
            std::fwrite(c.degree_multi.data(), 8, c.degree_multi.size(), f) == c.degree_multi.size() &&
            std::fwrite(c.col.data(), 4, c.col.size(), f) == c.col.size();
  std::fclose(f);
  if (!ok) err = "short write to " + path;
  return ok;
}

inline bool read_container(const std::string& path, Container& c, std::string& err) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) { err = "cannot open " + path; return false; }
  uint64_t hdr[6];
  if (std::fread(hdr, 8, 6, f) != 6 || (hdr[0] != kMagic && hdr[0] != kMagic1)) {
    std::fclose(f);
    err = path + " is not a PMGRAPH container";
    return false;
  }
  c.n_vertices = hdr[1]; c.n_slots = hdr[2]; c.n_slots_multi = hdr[3]; c.scale = hdr[4]; c.gen_ranks = hdr[5];
  c.delegate_threshold = 1048576;
  if (hdr[0] == kMagic && std::fread(&c.delegate_threshold, 8, 1, f) != 1) { std::fclose(f); err = path + " is truncated"; return false; }
  c.rowptr.resize(c.n_vertices + 1);
  c.degree_multi.resize(c.n_vertices);
  c.col.resize(c.n_slots);
  bool ok = std::fread(c.rowptr.data(), 8, c.rowptr.size(), f) == c.rowptr.size() &&
            std::fread(c.degree_multi.data(), 8, c.degree_multi.size(), f) == c.degree_multi.size() &&
            std::fread(c.col.data(), 4, c.col.size(), f) == c.col.size();
  std::fclose(f);
  if (!ok) err = path + " is truncated";
  return ok;
}

inline bool copy_file(const std::string& from, const std::string& to, std::string& err) {
  FILE* a = std::fopen(from.c_str(), "rb");
  if (!a) { err = "cannot open " + from; return false; }
  FILE* b = std::fopen(to.c_str(), "wb");
  if (!b) { std::fclose(a); err = "cannot create " + to; return false; }
  std::vector<char> buf(1 << 22);
  size_t n;
  bool ok = true;
  while ((n = std::fread(buf.data(), 1, buf.size(), a)) > 0) ok = ok && std::fwrite(buf.data(), 1, n, b) == n;
  std::fclose(a);
  std::fclose(b);
  if (!ok) err = "short write to " + to;
  return ok;
}

}  // namespace pmcli
