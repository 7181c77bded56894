/*
 * pmgpu.h — C ABI of libpmgpu.so, the B200-native (sm_100a) replacement for the
 * LCC/NLCC pruning path of HavoqGT's run_pattern_matching_beta.
 *
 * The reference's plugin API for this path is the C++ *visitor concept*
 * (doc/developer_guide.dox:40-57, include/havoqgt/visitor_queue.hpp:69-95,
 * 221-251, 395-411): templates over graph/queue types that cannot cross a
 * device boundary.  The drop-in seam is therefore the function level directly
 * above it — the free functions the driver calls once per superstep loop /
 * constraint — plus graph load, label build and the result writers.  Each entry
 * point below names the reference interface it replaces (paths relative to the
 * reference root).
 *
 * Conventions: every function returns 0 on success and a negative pm_status on
 * failure; pm_last_error(ctx) gives a message.  Handles are opaque.  All
 * pointers are HOST pointers to caller-owned buffers unless a name ends in
 * `_dev`; the engine owns all device memory.  One host thread per context, one
 * context per GPU (one process per GPU when several GPUs cooperate).
 * There is NO CPU fallback: without a CUDA device pm_create fails.
 */
#ifndef PMGPU_H
#define PMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pm_ctx pm_ctx;

enum pm_status {
  PM_OK = 0,
  PM_ERR_CUDA = -1,      /* a CUDA call failed                               */
  PM_ERR_ARG = -2,       /* bad argument / call order                        */
  PM_ERR_IO = -3,        /* pattern or result file problem                   */
  PM_ERR_PATTERN = -4,   /* malformed pattern directory                      */
  PM_ERR_CAPACITY = -5,  /* token pool exhausted even after growing          */
  PM_ERR_COMM = -6,      /* NCCL failure                                     */
  PM_ERR_UNSUPPORTED = -7
};

/* NLCC walker selection — the reference picks by constraint index
 * (`if (pl >= 4) do_tds_tp = true`, src/run_pattern_matching_beta.cpp:762-767). */
enum pm_nlcc_mode {
  PM_NLCC_NEM1 = 0, /* token_passing_pattern_matching_nonunique_nem_1.hpp:908-922   */
  PM_NLCC_TDS = 1   /* token_passing_pattern_matching_nonunique_tds_batch_1.hpp:976-984 */
};

/* ---- context ------------------------------------------------------------ */
/* replaces havoqgt_init + per-rank state of main (beta.cpp:144-162) */
int pm_create(pm_ctx** out, int device);
void pm_destroy(pm_ctx* ctx);
const char* pm_last_error(const pm_ctx* ctx);
/* number of kernels this context has launched so far (bench.py: gpu_launches) */
uint64_t pm_kernel_launches(const pm_ctx* ctx);

/* ---- multi-GPU: 1-D vertex partition owner(v) = v mod n_ranks ------------
 * replaces the MPI mailbox / visitor queue exchange
 * (include/havoqgt/new_mailbox.hpp:289-428, visitor_queue.hpp:395-434) and
 * vertex_data::all_{min,max}_reduce (impl/vertex_data.hpp:114-127).           */
#define PM_COMM_ID_BYTES 128
int pm_comm_unique_id(char id_out[PM_COMM_ID_BYTES]);
int pm_comm_init(pm_ctx* ctx, int rank, int n_ranks, const char id[PM_COMM_ID_BYTES]);

/* ---- graph store ---------------------------------------------------------
 * replaces delegate_partitioned_graph construction / distributed_db open
 * (src/generate_rmat.cpp:211-213, beta.cpp:209-223,
 *  include/havoqgt/impl/delegate_partitioned_graph.ipp:112-165).              */
/* Directed slots exactly as the reference edge iterator yields them (both
 * directions present, duplicates and self loops kept).  With n_ranks > 1 every
 * rank passes the slots whose SOURCE it owns (or all slots; foreign sources
 * are dropped).                                                               */
int pm_graph_from_slots(pm_ctx* ctx, uint64_t n_vertices, uint64_t n_slots,
                        const uint32_t* src, const uint32_t* dst);
/* generate_rmat -s scale on gen_ranks generating ranks, built on the GPU
 * (src/generate_rmat.cpp:197-213, include/havoqgt/rmat_edge_generator.hpp:218-259) */
int pm_graph_rmat(pm_ctx* ctx, uint64_t scale, uint64_t gen_ranks);

/* Distinct-neighbour CSR held in HOST memory (rows ascending, no duplicates)
 * plus the multigraph out-degrees the labels derive from: the analogue of
 * opening the reference's memory-mapped graph image (beta.cpp:213-223).  This is
 * the entry point whose host->device copies an end-to-end measurement includes. */
int pm_graph_from_csr(pm_ctx* ctx, uint64_t n_vertices, const uint64_t* rowptr /* n_vertices+1 */,
                      const uint32_t* col, const uint64_t* degree_multi /* n_vertices */);

/* Delegates (generate_rmat / ingest_edge_list -d; include/havoqgt/impl/delegate_partitioned_graph.ipp:501-512): the vertices
 * whose multigraph out-degree reaches `threshold` are hubs, numbered in ascending vertex order; a hub's controller is
 * delegate_id % n_ranks (delegate_partitioned_graph.hpp:231-233).  With several ranks the count files and the vertex / edge /
 * subgraph rows of a hub are attributed to its controller, like a reference run with delegates writes them; the hub's
 * adjacency itself stays with rank v mod n_ranks (deviations).  0 switches delegates off.  Collective. */
int pm_graph_set_delegate_threshold(pm_ctx* ctx, uint64_t threshold);
int pm_graph_num_delegates(const pm_ctx* ctx, uint64_t* n_out);

typedef struct {
  uint64_t n_vertices;      /* global                                  */
  uint64_t n_local;         /* vertices owned by this rank             */
  uint64_t n_slots_multi;   /* local directed slots with duplicates    */
  uint64_t n_slots;         /* local distinct (v,u) pairs              */
  uint64_t n_slots_padded;  /* local slots after 32-byte row padding   */
  uint64_t max_degree;      /* largest local multigraph out-degree     */
  uint64_t device_bytes;    /* device memory held by the graph store   */
} pm_graph_info_t;
int pm_graph_info(const pm_ctx* ctx, pm_graph_info_t* out);
/* copies of the distinct-neighbour CSR of the LOCAL vertices (tests) */
int pm_graph_get_degree(const pm_ctx* ctx, uint64_t* degree_multi_out /* n_local */);
int pm_graph_get_csr(const pm_ctx* ctx, uint64_t* rowptr_out /* n_local+1 */, uint32_t* col_out /* n_slots */);

/* ---- vertex labels -------------------------------------------------------
 * pm_labels_degree_log2 replaces vertex_data_db_degree
 * (include/havoqgt/vertex_data_db_degree.hpp:109: ceil(log2(degree+1)));
 * pm_labels_set replaces the -v loader (vertex_data_db.hpp:169-194).          */
int pm_labels_degree_log2(pm_ctx* ctx);
int pm_labels_set(pm_ctx* ctx, const uint64_t* labels /* n_vertices, global ids */);
int pm_labels_get(const pm_ctx* ctx, uint64_t* labels_out /* n_vertices */);

/* -v <base>: every file of dirname(base) whose name starts with basename(base) holds "vertex label" lines
 * (include/havoqgt/vertex_data_db.hpp:139-262); vertices without a line get label 0.  Collective over the ranks
 * (every rank reads the same files).                                                                          */
int pm_labels_from_files(pm_ctx* ctx, const char* base);

/* ---- text inputs of the reference tools, host only (no context, no device) -------------------------------
 * pm_io_read_vertex_data   the -v reader above into a caller buffer (labels_inout keeps the entries of
 *                          vertices that have no line)
 * pm_io_check_edge_data    the -e files (include/havoqgt/edge_data_db.hpp: "source target data" lines).  The
 *                          path never reads the values (beta.cpp:906, an unused reference): parsed and validated only
 * pm_io_read_edge_lists    "source target [weight]" lines (include/havoqgt/parallel_edge_list_reader.hpp:236-262)
 *                          of src/ingest_edge_list.cpp; undirected != 0 adds the reverse of every edge (-u 1).
 *                          Call with src_out = NULL for the sizes, then with buffers of n_slots entries.        */
int pm_io_read_vertex_data(const char* base, uint64_t n_vertices, uint64_t* labels_inout, uint64_t* n_pairs_out,
                           char* err_out, size_t err_cap);
int pm_io_check_edge_data(const char* base, uint64_t n_vertices, uint64_t* n_records_out, char* err_out, size_t err_cap);
int pm_io_read_edge_lists(const char* const* files, int n_files, int undirected, uint64_t* n_vertices_out,
                          uint64_t* n_slots_out, uint32_t* src_out, uint32_t* dst_out, char* err_out, size_t err_cap);

/* ---- pattern -------------------------------------------------------------
 * replaces ::graph 5-file ctor (include/havoqgt/graph.hpp:73-110) and
 * pattern_util (include/havoqgt/pattern_util.hpp:89-115); dir = "<p>/<ps>".   */
int pm_pattern_load_dir(pm_ctx* ctx, const char* dir);
typedef struct {
  int n_vertices, n_edges, diameter, n_constraints;
} pm_pattern_info_t;
int pm_pattern_info(const pm_ctx* ctx, pm_pattern_info_t* out);
/* one line of pattern_nlc (pattern_util.hpp:172-210) */
typedef struct {
  int walk_length;         /* vertices on the walk = cycle_length + 2                     */
  int valid_cycle;         /* 1: the walk must close at its source                        */
  int interleave_lcc;      /* 1: the driver runs LCC after this constraint removed a source */
  int order_independent;   /* 1: nem_1's result does not depend on message order (A.6 #7)  */
} pm_constraint_info_t;
int pm_pattern_constraint_info(const pm_ctx* ctx, int pl, pm_constraint_info_t* out);
/* Host-only check of a pattern directory with the reader pm_pattern_load_dir uses: no context, no device.
 * Returns PM_OK and fills *info_out (and up to constraints_cap entries of constraints_out, which may be null),
 * or PM_ERR_PATTERN with the reason in err_out (the message the reference's ::graph / pattern_util would
 * have failed on later, graph.hpp:195-270, pattern_util.hpp:172-278).                                      */
int pm_pattern_check_dir(const char* dir, pm_pattern_info_t* info_out, pm_constraint_info_t* constraints_out,
                         int constraints_cap, char* err_out, size_t err_cap);

/* ---- per-pattern state ---------------------------------------------------
 * replaces the container reset of beta.cpp:484-492                            */
int pm_state_reset(pm_ctx* ctx);

typedef struct {
  uint64_t n_vertices; /* |vertex_state_map| on this rank        */
  uint64_t n_edges;    /* sum of |vertex_active_edges_map[v]|    */
  double seconds;      /* device time of the superstep           */
} pm_counts_t;

/* ---- LCC -----------------------------------------------------------------
 * replaces label_propagation_pattern_matching_bsp
 * (include/havoqgt/label_propagation_pattern_matching_nonunique_ee.hpp:1029-1040):
 * runs `diameter` supersteps; *not_finished is OR-ed with "a vertex left the
 * vertex_state_map" (ee.hpp:968-970); counts_out receives one entry per
 * superstep (the rows the reference appends to its count files, :1131-1138).
 * With several ranks the flag and nothing else is reduced here; counts are
 * per rank like the reference's per-rank files.                               */
int pm_lcc(pm_ctx* ctx, int global_init_step, int* not_finished, pm_counts_t* counts_out);

/* CUDA-event timing of the LCC scan kernels on the context's stream, by kernel class, accumulated since
 * pm_create.  bin 0: first-superstep scan of the main row list (walks the pristine adjacency + label stream);
 * bin 1: later scans (active edge maps + mask gathers); bin 2: CTA-per-row scans (rows above 4096 slots);
 * bin 3: the per-pattern initialisation + signature filter; bin 4: the renaming scan (second superstep of the
 * first call: slots -> compact ids).                                                                       */
typedef struct {
  uint64_t launches;
  double ms;         /* sum of launch durations                    */
  uint64_t slots;    /* adjacency slots walked                     */
  uint64_t vertices; /* rows walked                                */
} pm_kernel_stats_t;
int pm_get_kernel_stats(const pm_ctx* ctx, int bin, pm_kernel_stats_t* out);

/* ---- NLCC ----------------------------------------------------------------
 * replaces token_passing_pattern_matching (both overloads) AND the driver's
 * post-processing of token_source_map (beta.cpp:956-1062): failed sources lose
 * bit pattern_indices[0]; vertices left without a bit are deactivated.
 * *pattern_found: a walk completed (beta.cpp:1136); *token_source_deleted: a
 * source failed (beta.cpp:1149) — both already reduced over ranks.            */
int pm_nlcc(pm_ctx* ctx, int pl, int mode, int* pattern_found, int* token_source_deleted,
            pm_counts_t* counts_out);

/* ---- the driver loop -----------------------------------------------------
 * replaces the do/while of beta.cpp:544-1351                                  */
typedef struct {
  int tds_from_pl;       /* constraints with index >= this use PM_NLCC_TDS (reference: 4); <0 never */
  int max_iterations;    /* safety cap, 0 = 1000                                                    */
  int lcc_only;          /* 1: repeat LCC while it removes vertices, never run NLCC                  */
  int keep_subgraphs;    /* 1: materialise enumerated subgraphs for pm_get_subgraphs / the writer    */
} pm_run_options_t;

typedef struct {
  uint64_t iterations;
  double search_seconds;       /* host wall clock of the loop                      */
  double device_seconds;       /* sum of per-row device times                      */
  uint64_t n_rows;
  uint64_t n_active_vertices;  /* final, this rank                                 */
  uint64_t n_active_edges;     /* final, this rank                                 */
  uint64_t path_count;         /* cumulative enumerated walks (never reset, A.6#5) */
  uint64_t edges_processed;    /* adjacency slots scanned + token fan-out          */
  uint64_t algorithmic_bytes;  /* SURVEY §8(d) byte model over the run             */
} pm_run_summary_t;

typedef struct {
  uint64_t itr;
  int32_t kind;  /* 0 = LP, 1 = TP */
  int32_t index; /* superstep or constraint */
  uint64_t n_vertices, n_edges;
  double seconds;
} pm_row_t;

int pm_run(pm_ctx* ctx, const pm_run_options_t* opt, pm_run_summary_t* out);
/* The run_fuzzy_pattern_matching path (src/run_fuzzy_pattern_matching.cpp:287-557; the driver is stale at this
 * revision, its compiling twin is src/run_pattern_matching.cpp:340-722): unique-label LCC
 * (include/havoqgt/label_propagation_pattern_matching_bsp.hpp:598-699) — every vertex stands for the FIRST template
 * vertex carrying its label and must hear all of that vertex's template neighbours among ALL its graph neighbours
 * each superstep, no edge elimination — followed, when LCC removed something, by cycle token passing over the
 * unpruned adjacency (include/havoqgt/token_passing_pattern_matching.hpp:514-530) whose failed sources leave the
 * vertex_state_map.  Rows: (itr, LP, k, |map|, 0) and one (itr, TP, 0, |map|, 0) per iteration that passed
 * tokens; pm_get_active_vertices returns (vertex, 1 << vertex_pattern_index).  Only opt->max_iterations is used.
 * Labels < 64.  Collective over the ranks of pm_comm_init (1-D partition; the mask array is replicated by slot,
 * removals travel as deltas, tokens through the owners' inboxes). */
int pm_run_fuzzy(pm_ctx* ctx, const pm_run_options_t* opt, pm_run_summary_t* out);
/* For a host driver that spells the loop out over pm_lcc / pm_nlcc itself (as the
 * reference main does): closes outer iteration `global_itr_count` (beta.cpp:1327-1341)
 * so that later rows carry the next iteration number and result_iteration gets its row. */
int pm_end_iteration(pm_ctx* ctx, double iteration_seconds);
int pm_get_rows(const pm_ctx* ctx, pm_row_t* rows_out /* n_rows */);

/* ---- results -------------------------------------------------------------
 * final vertex_state_map / vertex_active_edges_map of this rank
 * (beta.cpp:1386-1425); vertices ascending, edges sorted by (v, u).           */
int pm_get_active_vertices(const pm_ctx* ctx, uint64_t* vertices_out, uint16_t* template_bits_out);
int pm_get_active_edges(const pm_ctx* ctx, uint64_t* pairs_out /* 2 * n_active_edges */);
/* enumerated subgraphs of constraint pl from the LAST outer iteration that ran it */
int pm_get_subgraph_count(const pm_ctx* ctx, int pl, uint64_t* count_out, int* width_out);
int pm_get_subgraphs(const pm_ctx* ctx, int pl, uint32_t* rows_out /* count * width */);
/* writes the reference's result tree (names and row grammar of beta.cpp:504-535,
 * 1375-1425); like the reference it never creates directories.               */
int pm_write_results(const pm_ctx* ctx, const char* outdir);
/* the same for element `ps` of a pattern set: files under <outdir>/<ps>/, one more row of <outdir>/result_pattern_set
 * (appended for ps > 0).  The reference loops `for ps < 1` with a TODO (beta.cpp:424); the set loop is the drivers'. */
int pm_write_results_ps(const pm_ctx* ctx, const char* outdir, int ps);

#ifdef __cplusplus
}
#endif
#endif
