"""Literal transliteration of the run_fuzzy path with RANDOMISED message delivery.

TEST INFRASTRUCTURE (like everything under oracle/): pins oracle/pm_oracle.cpp:orc_run_fuzzy.
Follows, container by container,
  include/havoqgt/label_propagation_pattern_matching_bsp.hpp:66-310 (lppm_visitor), :317-520
  (verify_and_update_vertex_state_map), :527-593 (post step), :598-699 (superstep loop),
  include/havoqgt/token_passing_pattern_matching.hpp:98-335 (tppm_visitor) and the loop of
  src/run_pattern_matching.cpp:340-722 (TP_ORIG defined).  The mailbox delivers the messages of a
  traversal in a random order; results must not depend on it."""
import random


class VertexState:  # bsp.hpp:9-26
    def __init__(self):
        self.is_active = False
        self.vertex_pattern_index = 0
        self.pattern_vertex_itr_count_map = {}


def run(adj, labels, pat_labels, pat_adj, diameter, constraints, seed=0, max_iterations=50):
    """adj: list of neighbour lists (with multiplicity), labels: list, pat_adj: list of sets,
    constraints: list of (P, I, C, valid_cycle).  Returns (rows, {vertex: vertex_pattern_index})."""
    rng = random.Random(seed)
    V = len(adj)
    vertex_active = [True] * V
    vertex_state_map = {}
    rows = []

    def verify_and_update(v, q, parent_idx):  # bsp.hpp:317-520, itr_count == 1
        if parent_idx not in pat_adj[q]:
            return 0
        st = vertex_state_map.get(v)
        if st is None:
            st = vertex_state_map[v] = VertexState()
            st.vertex_pattern_index = q
        if st.is_active:
            return 1
        if not st.pattern_vertex_itr_count_map:
            for b in pat_adj[q]:
                st.pattern_vertex_itr_count_map.setdefault(b, 0)
        if not st.pattern_vertex_itr_count_map:
            return 0
        if parent_idx not in st.pattern_vertex_itr_count_map:
            return 0
        if st.pattern_vertex_itr_count_map[parent_idx] < 1:
            st.pattern_vertex_itr_count_map[parent_idx] = 1
        st.is_active = all(c != 0 for c in st.pattern_vertex_itr_count_map.values())
        if st.is_active:
            for b in st.pattern_vertex_itr_count_map:
                st.pattern_vertex_itr_count_map[b] = 0
        return 1

    def lcc(initstep):
        removed_any = False
        for superstep in range(diameter):
            need_map = superstep > 0 or not initstep
            msgs = []
            for u in range(V):  # init visit of every vertex (bsp.hpp:163-231)
                if not vertex_active[u]:
                    continue
                if need_map and u not in vertex_state_map:
                    continue
                idx = [q for q, l in enumerate(pat_labels) if l == labels[u]]
                if not idx:
                    vertex_active[u] = False
                    continue
                for w in adj[u]:
                    for p in idx:
                        msgs.append((w, p))
            rng.shuffle(msgs)
            for v, p in msgs:  # pre_visit :91-155, visit :235-290
                if not vertex_active[v]:
                    continue
                first = next((q for q, l in enumerate(pat_labels) if l == labels[v]), None)
                if first is None or p not in pat_adj[first]:
                    continue
                if need_map and v not in vertex_state_map:
                    continue
                for q, l in enumerate(pat_labels):
                    if l == labels[v]:
                        verify_and_update(v, q, p)
            gone = [v for v, st in vertex_state_map.items() if not st.is_active]  # :527-593
            for st in vertex_state_map.values():
                st.is_active = False
            for v in gone:
                vertex_active[v] = False
                del vertex_state_map[v]
            removed_any = removed_any or bool(gone)
            rows.append((itr, "LP", superstep, len(vertex_state_map), 0))
        return removed_any

    def token_passing(P, I, C, valid_cycle):  # token_passing_pattern_matching.hpp
        token_source_map, forwarded = {}, {}
        queue = []
        for v, st in vertex_state_map.items():
            if st.vertex_pattern_index == I[0] and labels[v] == P[0]:
                token_source_map[v] = False
                for w in adj[v]:
                    queue.append((w, v, 0, I[0]))
        while queue:
            v, target, c, parent_idx = queue.pop(rng.randrange(len(queue)))
            interior = C > c
            if interior and target in forwarded.get(v, ()):
                continue
            st = vertex_state_map.get(v)
            if st is None:
                continue
            ni = c + 1
            if not (labels[v] == P[ni] and st.vertex_pattern_index == I[ni] and parent_idx == I[ni - 1]):
                continue
            if interior:
                forwarded.setdefault(v, set()).add(target)
                for w in adj[v]:
                    queue.append((w, target, ni, st.vertex_pattern_index))
            elif v == target and valid_cycle:
                token_source_map[v] = True
        return token_source_map

    itr, initstep = 0, True
    while True:
        nf = lcc(initstep)
        initstep = False
        if nf:
            nf = False
            for P, I, C, valid_cycle in constraints:
                for v, ok in token_passing(P, I, C, valid_cycle).items():
                    if not ok:
                        vertex_active[v] = False
                        del vertex_state_map[v]
                        nf = True
            rows.append((itr, "TP", 0, len(vertex_state_map), 0))
        itr += 1
        if not nf or itr >= max_iterations:
            break
    return rows, {v: st.vertex_pattern_index for v, st in vertex_state_map.items()}, itr
