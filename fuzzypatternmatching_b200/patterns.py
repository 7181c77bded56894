"""Pattern-directory writer for the reference's on-disk template format.

The engine reads `<pattern_dir>/<ps>/pattern_{edge,vertex,vertex_data,edge_data,
stat,nlc,non_local_constraint}` exactly like the reference driver
(/root/reference/src/run_pattern_matching_beta.cpp:433-441,473-475; grammar in
include/havoqgt/graph.hpp:181-270,337-358 and include/havoqgt/pattern_util.hpp:172-210,
254-278).  This module only WRITES such directories from a small Python
description, so tests and bench.py can create their templates without shipping
data files.

A spec is a dict:
  labels      : list, label of template vertex i
  edges       : list of undirected (a, b)
  diameter    : number of LCC supersteps per call (pattern_stat "diameter")
  constraints : list of dicts {walk: [template ids], cycle: bool,
                               interleave: bool (default True), tds: bool}
  optional_edges : (approximate matching, run_pattern_matching_beta_2.cpp:459-476) undirected (a, b) that are
                   OPTIONAL: their pattern_edge lines carry the flag column "s t 0", every other line "s t 1"
                   (approximate_pattern_matching/pattern_graph.hpp:320-337); must be a subset of `edges`
  min_optional   : {template vertex: vertex_min_optional_edge_count} -> pattern_vertex_local_constraints
                   ("v : count" per template vertex, -1 where none is given; pattern_graph.hpp:282-315)
Constraints flagged tds must come last: the reference switches to template
driven search by constraint INDEX (`pl >= 4`, beta.cpp:762), which the engine
exposes as the `tds_from_pl` option; `tds_from_pl(spec)` returns that index.
"""
import os


def _enum_indices(walk):
    # pattern_non_local_constraint field 2: position of the first occurrence of
    # walk[h] ("new vertex" when it equals h, "must equal visited[e]" when smaller;
    # token_passing_pattern_matching_nonunique_tds_batch_1.hpp:284-302)
    return [walk.index(w) for w in walk]


def tds_from_pl(spec):
    for i, c in enumerate(spec.get("constraints", [])):
        if c.get("tds"):
            return i
    return -1


def write_pattern_dir(base, spec, ps=0):
    """Writes `<base>/<ps>/pattern_*`; returns `<base>/<ps>`."""
    d = os.path.join(base, str(ps))
    os.makedirs(d, exist_ok=True)
    labels = spec["labels"]
    both = sorted(set((a, b) for a, b in spec["edges"]) | set((b, a) for a, b in spec["edges"]))
    eid = {}
    for a, b in spec["edges"]:
        eid[(a, b)] = eid[(b, a)] = len(eid) // 2
    optional = set()
    for a, b in spec.get("optional_edges", []):
        optional |= {(a, b), (b, a)}
    approximate = bool(optional) or "min_optional" in spec
    with open(os.path.join(d, "pattern_edge"), "w") as f:
        for a, b in both:
            if approximate:
                f.write("%d %d %d\n" % (a, b, 0 if (a, b) in optional else 1))
            else:
                f.write("%d %d\n" % (a, b))
    lc = os.path.join(d, "pattern_vertex_local_constraints")
    if approximate:
        with open(lc, "w") as f:
            for i in range(len(labels)):
                f.write("%d : %d\n" % (i, spec.get("min_optional", {}).get(i, -1)))
    elif os.path.exists(lc):
        os.remove(lc)
    with open(os.path.join(d, "pattern_edge_data"), "w") as f:
        for a, b in both:
            f.write("%d %d %d 55\n" % (a, b, eid[(a, b)]))
    with open(os.path.join(d, "pattern_vertex"), "w") as f:
        f.write("\n")
    with open(os.path.join(d, "pattern_vertex_data"), "w") as f:
        for i, l in enumerate(labels):
            f.write("%d %d\n" % (i, l))
    with open(os.path.join(d, "pattern_stat"), "w") as f:
        f.write("diameter : %d\n" % spec["diameter"])
    seen_tds = False
    with open(os.path.join(d, "pattern_nlc"), "w") as f, \
            open(os.path.join(d, "pattern_non_local_constraint"), "w") as g:
        for c in spec.get("constraints", []):
            if seen_tds and not c.get("tds"):
                raise ValueError("tds constraints must come last (beta.cpp:762)")
            seen_tds = seen_tds or bool(c.get("tds"))
            walk = c["walk"]
            f.write("%s : %s : %d : %d : %d : 0\n" % (
                " ".join(str(labels[w]) for w in walk), " ".join(str(w) for w in walk),
                len(walk) - 2, int(bool(c.get("cycle"))), int(c.get("interleave", True))))
            agg = [0] * len(walk) if not c.get("tds") else [0] + [1] * (len(walk) - 1)
            g.write("%s : %s : %s\n" % (" ".join(str(w) for w in walk),
                                        " ".join(str(e) for e in _enum_indices(walk)),
                                        " ".join(str(a) for a in agg)))
    return d


# The README's example template (examples/rmat_log2_tree_pattern/0): a 7-vertex
# tree with degree-log2 labels 3,4,7,2,3,5,7; four path constraints between the
# repeated labels (3..3 and 7..7, both directions) and one full-template walk.
RMAT_LOG2_TREE = {
    "labels": [3, 4, 7, 2, 3, 5, 7],
    "edges": [(0, 1), (1, 2), (1, 3), (3, 5), (4, 5), (5, 6)],
    "diameter": 8,
    "constraints": [
        {"walk": [4, 5, 3, 1, 0]},
        {"walk": [0, 1, 3, 5, 4]},
        {"walk": [2, 1, 3, 5, 6]},
        {"walk": [6, 5, 3, 1, 2]},
        {"walk": [0, 1, 2, 1, 3, 5, 4, 5, 6], "tds": True},
    ],
}


def triangle(la, lb, lc):
    return {
        "labels": [la, lb, lc],
        "edges": [(0, 1), (1, 2), (0, 2)],
        "diameter": 2,
        "constraints": [{"walk": [0, 1, 2, 0], "cycle": True},
                        {"walk": [0, 1, 2, 0], "cycle": True, "tds": True}],
    }


def cycle4(la, lb, lc, ld):
    return {
        "labels": [la, lb, lc, ld],
        "edges": [(0, 1), (1, 2), (2, 3), (0, 3)],
        "diameter": 3,
        "constraints": [{"walk": [0, 1, 2, 3, 0], "cycle": True},
                        {"walk": [0, 1, 2, 3, 0], "cycle": True, "tds": True}],
    }


def cycle6_chords(labels):
    """6-cycle 0-1-2-3-4-5-0 with chords (0,3) and (1,4); `labels` has 6 entries."""
    assert len(labels) == 6
    return {
        "labels": list(labels),
        "edges": [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (0, 5), (0, 3), (1, 4)],
        "diameter": 3,
        "constraints": [{"walk": [0, 1, 2, 3, 4, 5, 0], "cycle": True},
                        {"walk": [0, 1, 2, 3, 0], "cycle": True},
                        {"walk": [1, 2, 3, 4, 1], "cycle": True},
                        # one walk that covers every template edge
                        {"walk": [0, 1, 2, 3, 4, 5, 0, 3, 4, 1], "tds": True}],
    }


# ---------------------------------------------------------------------------
# BASELINE configs[3]: "edit-distance-k prototype set (k = 1, 2)".  The reference holds no
# prototype logic (grep -i "edit.?dist|prototype" over src/ and include/ finds nothing; its drivers
# only ever read pattern directory 0, run_fuzzy_pattern_matching.cpp:208), so the SET is ours: the
# connected templates obtained from a base template by deleting up to k edges.  Each prototype is an
# ordinary pattern directory <base>/<i>/ and is searched independently.
# ---------------------------------------------------------------------------
def _connected(n, edges):
    adj = {i: set() for i in range(n)}
    for a, b in edges:
        adj[a].add(b)
        adj[b].add(a)
    seen, todo = {0}, [0]
    while todo:
        x = todo.pop()
        for y in adj[x]:
            if y not in seen:
                seen.add(y)
                todo.append(y)
    return len(seen) == n


def _diameter(n, edges):
    adj = {i: set() for i in range(n)}
    for a, b in edges:
        adj[a].add(b)
        adj[b].add(a)
    best = 0
    for s in range(n):
        dist, todo = {s: 0}, [s]
        for x in todo:
            for y in adj[x]:
                if y not in dist:
                    dist[y] = dist[x] + 1
                    todo.append(y)
        best = max(best, max(dist.values()))
    return best


def edit_distance_prototypes(spec, k):
    """[(deleted_edges, prototype_spec)] for every connected template within edit distance k of `spec`
    (edge deletions only; distance 0 = the base template itself comes first).  A constraint is kept iff
    every edge of its walk is still present; the superstep count follows the prototype's diameter."""
    import itertools
    n = len(spec["labels"])
    base = [tuple(sorted(e)) for e in spec["edges"]]
    out = []
    for dist in range(0, k + 1):
        for gone in itertools.combinations(base, dist):
            edges = [e for e in base if e not in gone]
            if not edges or not _connected(n, edges):
                continue
            es = set(edges)
            cons = [c for c in spec.get("constraints", [])
                    if all(tuple(sorted((a, b))) in es for a, b in zip(c["walk"], c["walk"][1:]))]
            out.append((list(gone), {"labels": list(spec["labels"]), "edges": edges,
                                     "diameter": max(spec.get("diameter", 1) if dist == 0 else 1, _diameter(n, edges)),
                                     "constraints": cons}))
    return out


def write_prototype_set(base_dir, spec, k):
    """Writes <base_dir>/<i>/pattern_* for every prototype; returns [(i, deleted_edges, dir)]."""
    out = []
    for i, (gone, proto) in enumerate(edit_distance_prototypes(spec, k)):
        out.append((i, gone, write_pattern_dir(base_dir, proto, ps=i)))
    return out
