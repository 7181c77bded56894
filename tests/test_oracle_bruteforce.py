"""Soundness of the restated pruning against a brute-force enumerator (networkx):
every vertex and edge that takes part in at least one exact match survives, with the
template bit of its role still set, and template driven search enumerates exactly
the label-preserving monomorphisms."""
import networkx as nx
import numpy as np
import pytest
from networkx.algorithms import isomorphism as iso

from fuzzypatternmatching_b200 import patterns as PT
from tests import cases


def _matches(n, edges, labels, spec):
    G = nx.Graph()
    G.add_nodes_from((i, {"l": int(labels[i])}) for i in range(n))
    G.add_edges_from((a, b) for a, b in edges if a != b)
    P = nx.Graph()
    P.add_nodes_from((i, {"l": l}) for i, l in enumerate(spec["labels"]))
    P.add_edges_from(spec["edges"])
    gm = iso.GraphMatcher(G, P, node_match=lambda a, b: a["l"] == b["l"])
    return [{p: g for g, p in m.items()} for m in gm.subgraph_monomorphisms_iter()]


@pytest.mark.parametrize("name,spec,labelset,tds_from", cases.SPECS[:3], ids=[s[0] for s in cases.SPECS[:3]])
def test_pruning_is_sound_and_enumeration_is_exact(oracle, name, spec, labelset, tds_from):
    found_any = 0
    for seed in range(8):
        n, m = 40, 130 + 20 * (seed % 3)
        edges = cases.random_multigraph(seed + 100, n, m)
        labels = cases.random_labels(seed + 100, n, labelset)
        ms = _matches(n, edges, labels, spec)
        g = oracle.Graph.from_undirected(n, edges)
        r = oracle.Run(g, labels, oracle.Pattern(cases.pattern_dir(spec)), tds_from_pl=tds_from, max_iterations=50)
        T = r.template_vertices
        alive = r.in_map
        E = set(map(tuple, r.active_edges.tolist()))
        for mp in ms:
            for p, v in mp.items():
                assert alive[v] and (T[v] >> p) & 1, "a matched vertex lost its role"
            for a, b in spec["edges"]:
                assert (mp[a], mp[b]) in E and (mp[b], mp[a]) in E, "a matched edge was pruned"
        # the last constraint of every spec is the full-template walk
        walk = spec["constraints"][-1]["walk"]
        want = sorted(tuple(mp[w] for w in walk) for mp in ms)
        got = sorted(map(tuple, r.subgraphs[len(spec["constraints"]) - 1].tolist()))
        assert got == want
        found_any += len(ms) > 0
    assert found_any >= 2
