/*
 * pm_oracle.h — C API of the CPU ORACLE (test infrastructure, NOT the product).
 *
 * The oracle is a CPU restatement of the reference's pattern-matching pruning
 * path (HavoqGT run_pattern_matching_beta: LCC label propagation + NLCC token
 * passing).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product path (libpmgpu.so) never
 * links, loads or calls anything declared here.
 *
 * PARITY STATUS: pinned by the reference's own code.  The reference ships no
 * golden vectors / tests for this path and cannot be built
 * as it ships (cmake, MPI, Boost), but its driver and visitor headers compile
 * from /root/reference over a single-rank runtime stand-in (oracle/ref_shim ->
 * oracle/_ref/run_pattern_matching_beta).  The oracle is held to
 *   (0) that binary's result files: tests/test_oracle_vs_reference.py (live) and
 *       the JSON files of tests/golden/reference_runs (committed outputs),
 *   (1) a literal, dict/set based Python transliteration of the reference
 *       visitors with randomised message delivery (oracle/ref_literal.py),
 *   (2) a brute-force subgraph-isomorphism property check (networkx),
 *   (3) hand-derived known-answer tests (tests/golden/).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#ifndef PM_ORACLE_H
#define PM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_graph orc_graph;
typedef struct orc_pattern orc_pattern;
typedef struct orc_run orc_run;

/* ---- R-MAT stream (src/generate_rmat.cpp:197-205,
 *      include/havoqgt/rmat_edge_generator.hpp:126-139,218-259,
 *      include/havoqgt/detail/hash.hpp:65-143) ---- */
/* Writes the first n_edges GENERATED edges of generating rank `rank`
 * (seed 5489+3*rank) as (u,v) pairs: out[2*i], out[2*i+1].  */
void orc_rmat_stream(uint64_t scale, uint64_t rank, uint64_t n_edges, uint64_t* out);
uint64_t orc_hash_nbits(uint64_t input, int n);

/* ---- graph ---- */
/* Directed slot list exactly as the reference's edge iterator yields it (both
 * directions present, duplicates and self loops kept).  */
orc_graph* orc_graph_from_slots(uint64_t n_vertices, uint64_t n_slots,
                                const uint64_t* src, const uint64_t* dst);
/* generate_rmat -s scale on gen_ranks ranks; threads = host threads to use. */
orc_graph* orc_graph_rmat(uint64_t scale, uint64_t gen_ranks, int threads);
void orc_graph_free(orc_graph*);
uint64_t orc_graph_num_vertices(const orc_graph*);
uint64_t orc_graph_num_slots_multi(const orc_graph*);  /* with duplicates      */
uint64_t orc_graph_num_slots(const orc_graph*);        /* distinct (v,u) pairs */
const uint64_t* orc_graph_rowptr(const orc_graph*);    /* V+1, distinct CSR    */
const uint32_t* orc_graph_col(const orc_graph*);       /* sorted per row       */
const uint64_t* orc_graph_degree(const orc_graph*);    /* multigraph out-degree */
/* label = (uint64)ceil(log2(degree+1)) (vertex_data_db_degree.hpp:109) */
void orc_labels_degree_log2(const orc_graph*, uint64_t* labels_out);

/* ---- pattern directory "<p>/<ps>" (graph.hpp:73-110,181-270,337-358;
 *      pattern_util.hpp:89-115,172-210,254-278) ---- */
orc_pattern* orc_pattern_load(const char* dir);
void orc_pattern_free(orc_pattern*);
const char* orc_pattern_error(const orc_pattern*);
int orc_pattern_num_vertices(const orc_pattern*);
int orc_pattern_num_edges(const orc_pattern*);
int orc_pattern_diameter(const orc_pattern*);
int orc_pattern_num_constraints(const orc_pattern*);

/* ---- run (src/run_pattern_matching_beta.cpp:481-1425) ---- */
typedef struct {
  int n_ranks;         /* partitions for the per-rank result layout (v mod R) */
  int tds_from_pl;     /* constraints with index >= this use TDS (beta.cpp:762); <0: never */
  int max_iterations;  /* safety cap on the outer do/while (A.6 #4); 0 = 1000  */
  int lcc_only;        /* 1: skip NLCC entirely (config 2: LCC to fixed point) */
  int threads;         /* OpenMP threads; 0 = all                              */
  int keep_subgraphs;  /* 1: keep enumerated subgraph rows in memory           */
  uint64_t delegate_threshold; /* hubs: multigraph out-degree >= this; 0 = none */
} orc_options;

typedef struct {
  uint64_t itr;
  int32_t kind;       /* 0 = LP (LCC superstep), 1 = TP (NLCC constraint) */
  int32_t index;      /* superstep k or constraint pl                     */
  uint64_t n_vertices;/* |vertex_state_map|                               */
  uint64_t n_edges;   /* sum |E_v| over the map                           */
  double seconds;
} orc_row;

orc_run* orc_run_pattern(const orc_graph*, const uint64_t* labels,
                         const orc_pattern*, const orc_options*);
void orc_run_free(orc_run*);
uint64_t orc_run_num_rows(const orc_run*);
const orc_row* orc_run_rows(const orc_run*);
uint64_t orc_run_iterations(const orc_run*);
double orc_run_search_seconds(const orc_run*);
/* final state */
const uint16_t* orc_run_template_vertices(const orc_run*); /* T_arr, V entries  */
const uint8_t* orc_run_in_map(const orc_run*);             /* V entries         */
uint64_t orc_run_num_active_edges(const orc_run*);
/* fills (v,u) pairs of the final active edges sorted by (v,u) */
void orc_run_active_edges(const orc_run*, uint64_t* pairs_out);
/* enumerated subgraphs of constraint pl, final outer iteration only (A.6 #5) */
uint64_t orc_run_num_subgraphs(const orc_run*, int pl);
int orc_run_subgraph_width(const orc_run*, int pl);
const uint32_t* orc_run_subgraphs(const orc_run*, int pl); /* n*width ids, unsorted */
uint64_t orc_run_cumulative_path_count(const orc_run*);    /* never reset (A.6 #5) */
/* edges scanned: sum of LCC messages (slots walked) + NLCC token fan-out */
uint64_t orc_run_edges_processed(const orc_run*);
/* hazards: situations where the reference's result depends on message order,
 * or where a level-synchronous restatement is not provably identical.
 *  [0] nem_1 token reached a (vertex,source) already claimed at another hop
 *  [1] nem_1 >=2 distinct parents and an excluded parent was a viable target
 *  [2] LCC delivery along an edge the receiver no longer holds (A.6 #11)
 *  [3] T_arr grew between supersteps (bit resurrected, A.6 #4)
 *  [4] outer loop hit max_iterations
 *  [5] (not a hazard, a coverage counter) edges that survived an LCC post step only because nem_1 had
 *      set their flag outside LCC (A.6 #11)                                  */
const uint64_t* orc_run_hazards(const orc_run*);
/* writes the reference's result tree under outdir (must pre-exist like the
 * reference requires, beta.cpp:413-414,504-535); returns 0 on success */
int orc_run_write_results(const orc_run*, const orc_graph*, const uint64_t* labels,
                          const orc_pattern*, const char* outdir);

/* ---- run_fuzzy path: unique-label LCC (label_propagation_pattern_matching_bsp.hpp) + cycle
 * token passing over the unpruned adjacency (token_passing_pattern_matching.hpp), loop of
 * src/run_pattern_matching.cpp:340-722.  Rows: (itr, LP, k, |map|, 0) per superstep and (itr, TP, 0, |map|, 0)
 * once per iteration that ran token passing; template_vertices[v] = 1 << vertex_pattern_index.  ---- */
orc_run* orc_run_fuzzy(const orc_graph*, const uint64_t* labels, const orc_pattern*, const orc_options*);

#ifdef __cplusplus
}
#endif
#endif
