"""The oracle's multi-rank attribution with delegates (what tests/multi_gpu_check.py holds the engine to at 2/4/8 GPUs), against
the reference's rules restated independently here, straight from the edge list:
  * hubs = vertices whose out-degree — directed slots, parallel edges and self loops counted — reaches the threshold
    (impl/delegate_partitioned_graph.ipp:501-512); delegate id = position in the ascending list of hubs (:681-683);
    controller = delegate id mod ranks (delegate_partitioned_graph.hpp:231-233);
  * a vertex of the final vertex_state_map is written, with its edges, by its controller if it is a hub and by rank
    `v mod ranks` otherwise (beta.cpp:1380-1410), and counted in that rank's count files (ee.hpp:1112-1125);
  * nothing else changes: the union over the ranks is the single-rank result and the per-rank rows add up to its rows."""
import pytest

from fuzzypatternmatching_b200 import patterns as PT
from tests import cases




@pytest.mark.parametrize("ranks,threshold", [(2, 14), (3, 15), (4, 16), (8, 17)])
def test_hub_rows_belong_to_their_controller(oracle, tmp_path, ranks, threshold):
    spec = PT.triangle(1, 2, 3)
    n = 160
    edges = cases.random_multigraph(ranks * 31 + threshold, n, 900, dup=0.2, loops=0.1)
    labels = cases.random_labels(ranks * 31 + threshold, n, [1, 2, 3])
    st = cases.check_multi_rank_attribution(oracle, n, edges, labels, spec, 1, ranks, threshold, str(tmp_path / "tree"))
    assert 5 <= st["hubs"] < (2 * n) // 3
    assert st["final_vertices"] > 10 and st["hubs_moved"] > 0, "the input must end with hubs whose controller is not their modulo owner"


def test_attribution_over_templates_and_rank_counts(oracle, tmp_path):
    """the same check over the tree template (path constraints + enumeration), the 4-cycle and the twin template, 2..8 ranks,
    with and without hubs"""
    import random
    rng = random.Random(99)
    moved = 0
    for i in range(24):
        name, spec, labelset, tds_from = cases.SPECS[i % 3]
        n = rng.choice([60, 150, 300])
        m = int(n * rng.choice([3.0, 5.0]))
        if i % 2:
            edges, labels = cases.planted(i, n, m, spec, labelset)
        else:
            edges, labels = cases.random_multigraph(i, n, m, dup=0.2, loops=0.1), cases.random_labels(i, n, labelset)
        ranks = rng.choice([2, 3, 4, 8])
        threshold = rng.choice([0, int(2 * m / n) + 2, int(2 * m / n) + 5])
        st = cases.check_multi_rank_attribution(oracle, n, edges, labels, spec, tds_from, ranks, threshold, str(tmp_path / ("t%d" % i)))
        moved += st["hubs_moved"]
    assert moved > 0
