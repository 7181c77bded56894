"""Literal, dict/set based transliteration of the reference visitors with a
RANDOMISED asynchronous scheduler — TEST INFRASTRUCTURE ONLY (small graphs).

Purpose: pin oracle/pm_oracle.cpp.  The C++ oracle evaluates LCC supersteps in
a fixed order and NLCC level by level; this module executes the visitors the
way HavoqGT does (init visitors, in-flight messages, local queues) but picks
the next event at random, so any dependence of the result on message order
shows up as a run-to-run difference, and equality with the C++ oracle over many
seeds is evidence that the restatement is order-faithful.

Every function names the reference lines it mirrors (relative to
/root/reference).  Single logical rank, no delegates (hubs behave as ordinary
vertices logically).
"""
import random


class PatternFiles:
    """graph.hpp:73-110,181-270,337-358 and pattern_util.hpp:172-210,254-278"""

    def __init__(self, directory):
        base = directory + "/pattern"
        self.N = {}
        last = -1
        self.n_edges = 0
        for line in open(base + "_edge"):
            if not line.strip():
                continue
            s, t = [int(x) for x in line.split()[:2]]
            self.N.setdefault(s, []).append(t)
            last = s
            self.n_edges += 1
        self.nv = last + 1
        self.vertex_data = [int(l.split()[1]) for l in open(base + "_vertex_data") if l.strip()]
        self.diameter = 0
        for line in open(base + "_stat"):
            k = line.split(":")
            if len(k) > 1 and k[0].strip().lower() == "diameter":
                self.diameter = int(k[1])
        self.cons = []
        try:
            for line in open(base + "_nlc"):
                if not line.strip():
                    continue
                f = [x.strip() for x in line.split(":")]
                self.cons.append(dict(P=[int(x) for x in f[0].split()], I=[int(x) for x in f[1].split()],
                                      C=int(f[2]), valid_cycle=bool(int(f[3])), interleave=bool(int(f[4])),
                                      selected=bool(int(f[5])), enum=None))
            k = 0
            for line in open(base + "_non_local_constraint"):
                if not line.strip():
                    continue
                f = [x.strip() for x in line.split(":")]
                if k < len(self.cons):
                    self.cons[k]["enum"] = [int(x) for x in f[1].split()]
                k += 1
        except FileNotFoundError:
            pass


class LiteralRun:
    def __init__(self, n_vertices, slots, labels, pattern, seed=0, tds_from_pl=4, max_iterations=50):
        """slots: list of directed (u, v) exactly as the reference edge iterator yields them."""
        self.rng = random.Random(seed)
        self.V = n_vertices
        self.adj = [[] for _ in range(n_vertices)]  # CSR with multiplicity (ee.hpp:555-560)
        for u, v in slots:
            self.adj[u].append(v)
        self.label = list(labels)
        self.pg = pattern
        self.tds_from_pl = tds_from_pl
        self.max_iterations = max_iterations
        # beta.cpp:484-492
        self.state_map = {}
        self.active = [True] * n_vertices
        self.template_vertices = [0] * n_vertices
        self.E = [dict() for _ in range(n_vertices)]  # vertex_active_edges_map
        self.token_source_set = [set() for _ in range(n_vertices)]
        self.rows = []
        self.subgraphs = [[] for _ in pattern.cons]
        self.errors = []
        self.run()

    # ------------------------------------------------------------ scheduler
    def traverse(self, init_visitors, pre_visit, visit):
        """visitor_queue.hpp:221-251 (init_visitor_traversal_new), :395-411 (queue_visitor).
        Events: start an init visitor / deliver an in-flight visitor (pre_visit,
        then push) / pop a queued visitor (visit).  Chosen uniformly at random."""
        init = list(init_visitors)
        self.rng.shuffle(init)
        inflight, queued = [], []
        self._send = inflight.append
        while init or inflight or queued:
            pools = [p for p in (init, inflight, queued) if p]
            pool = self.rng.choice(pools)
            i = self.rng.randrange(len(pool))
            pool[i], pool[-1] = pool[-1], pool[i]
            vis = pool.pop()
            if pool is init:
                visit(vis)  # do_init_visit -> init_visit == visit, no pre_visit
            elif pool is inflight:
                if pre_visit(vis):
                    queued.append(vis)
            else:
                visit(vis)

    # ------------------------------------------------------------------ LCC
    def nbr_bits(self, p):
        m = 0
        for t in self.pg.N.get(p, []):
            m |= 1 << t
        return m

    def labelmask(self, lab):
        m = 0
        for i, l in enumerate(self.pg.vertex_data):
            if l == lab:
                m |= 1 << i
        return m

    def lp_verify(self, v, parent, pbits, tv):
        """ee.hpp:646-816"""
        match_found = valid_parent = False
        for a in range(16):
            if (tv >> a) & 1:
                match_found = True
                for i in range(16):
                    if (pbits >> i) & 1 and i in self.pg.N.get(a, []):
                        valid_parent = True
                        break
            if valid_parent:
                break
        if not match_found or not valid_parent:
            return 0
        if v not in self.state_map:
            self.state_map[v] = dict(tv=tv, tn=0)  # :734-746
        self.state_map[v]["tn"] |= pbits  # :775
        first = self.superstep == 0 and self.init_step
        if parent not in self.E[v]:  # :791-813
            if first:
                self.E[v][parent] = 1
            else:
                self.errors.append(("lp_edge_missing", v, parent))
                return 0
        else:
            self.E[v][parent] = 1
        return 1

    def lp_pre_visit(self, vis):
        """ee.hpp:148-459, non-delegate path"""
        v, parent, pbits, msg_type = vis
        if not self.active[v]:
            return False
        if msg_type == 0:
            return True
        first = self.superstep == 0 and self.init_step
        if first:  # :368-406
            tv = self.labelmask(self.label[v])
            if tv == 0:
                self.active[v] = False
                return False
            if pbits == 0:
                return False
            self.template_vertices[v] = tv
            self.lp_verify(v, parent, pbits, tv)
            return False
        if v not in self.state_map:  # :414-417
            return False
        if pbits == 0:
            return False
        if self.state_map[v]["tv"] == 0:
            self.errors.append(("lp_no_bit", v))
            return False
        tv = self.template_vertices[v]
        if tv == 0:
            self.errors.append(("lp_no_bit_arr", v))
            return False
        self.lp_verify(v, parent, pbits, tv)
        return False

    def lp_visit(self, vis):
        """ee.hpp:467-636"""
        v, parent, pbits, msg_type = vis
        if not self.active[v]:
            return
        first = self.superstep == 0 and self.init_step
        if not first and v not in self.state_map:  # :481-486
            return
        if first:  # :519-569
            tv = self.labelmask(self.label[v])
            if tv == 0:
                self.active[v] = False
                return
            self.template_vertices[v] = tv
            if msg_type == 0:
                for n in self.adj[v]:
                    self._send((n, v, tv, 1))
            return
        tv = self.template_vertices[v]  # :573-624
        if tv == 0:
            return
        if msg_type == 0:
            for n in list(self.E[v].keys()):
                self._send((n, v, tv, 1))

    def lp_post(self):
        """ee.hpp:827-1027"""
        first = self.superstep == 0 and self.init_step
        if first:  # :841-852
            for v in range(self.V):
                if self.active[v] and v not in self.state_map:
                    self.active[v] = False
                    self.E[v].clear()
        remove = []
        for v, st in self.state_map.items():  # :886-966
            for p in range(16):
                if (st["tv"] >> p) & 1:
                    need = self.nbr_bits(p)
                    got = need & st["tn"]
                    if not (need == got and got != 0):
                        st["tv"] &= ~(1 << p)
            if st["tv"] == 0:
                remove.append(v)
                self.active[v] = False
                self.E[v].clear()
            else:
                self.template_vertices[v] = st["tv"]
                st["tn"] = 0
                for n in list(self.E[v].keys()):
                    if not self.E[v][n]:
                        del self.E[v][n]
                    else:
                        self.E[v][n] = 0
        if remove:
            self.not_finished = True
        for v in remove:
            del self.state_map[v]

    def counts(self):
        return len(self.state_map), sum(len(self.E[v]) for v in self.state_map)

    def lcc(self):
        """ee.hpp:1029-1153"""
        for k in range(self.pg.diameter):
            self.superstep = k
            self.traverse([(v, None, 0, 0) for v in range(self.V)], self.lp_pre_visit, self.lp_visit)
            self.lp_post()
            nv, ne = self.counts()
            self.rows.append((self.itr, "LP", k, nv, ne))

    # ---------------------------------------------------------------- nem_1
    def tp_static(self, v, h):
        c = self.c
        if self.label[v] != c["P"][h]:
            return False
        tv = self.template_vertices[v]
        return tv != 0 and (tv >> c["I"][h]) & 1 == 1

    def nem_pre_visit(self, vis):
        """nem_1.hpp:98-303"""
        c = self.c
        v = vis["vertex"]
        if not self.active[v]:
            return False
        if vis["ack"]:
            return True
        interior = c["C"] > vis["itr"]
        if interior and vis["target"] in self.token_source_set[v]:  # :131-139
            return False
        if interior and v == vis["target"]:  # :174-177
            return False
        h = vis["itr"] + 1
        if not self.tp_static(v, h):  # :186-210
            return False
        if vis["ppi"] != c["I"][h - 1]:  # :228-231
            return False
        if interior:  # :270-285
            if vis["target"] in self.token_source_set[v]:
                self.errors.append(("tp_set_dup", v))
                return False
            self.token_source_set[v].add(vis["target"])
        return True

    def nem_visit(self, vis):
        """nem_1.hpp:311-861"""
        c = self.c
        v = vis["vertex"]
        if not self.active[v]:
            return
        if vis["ack"]:  # :326-342
            if v not in self.token_source_map:
                self.errors.append(("tp_ack_missing", v))
                return
            self.token_source_map[v] = 1
            return
        if vis["init"]:  # :387-527
            if self.label[v] != c["P"][0]:
                return
            tv = self.template_vertices[v]
            if tv == 0 or not (tv >> c["I"][0]) & 1:
                return
            if not c["valid_cycle"] and not ((tv >> c["I"][0]) & 1 and (tv >> c["I"][-1]) & 1):
                return
            self.token_source_map.setdefault(v, 0)
            for n in list(self.E[v].keys()):
                self._send(dict(vertex=n, parent=v, target=v, itr=0, ppi=c["I"][0], init=False, ack=False))
            return
        interior = c["C"] > vis["itr"]
        if interior and v == vis["target"]:  # :543-546
            return
        h = vis["itr"] + 1
        if not self.tp_static(v, h):
            return
        if interior:  # :608-659
            if vis["ppi"] != c["I"][h - 1]:
                return
            for n in list(self.E[v].keys()):  # :832-851
                if n == vis["parent"]:
                    continue
                self._send(dict(vertex=n, parent=v, target=vis["target"], itr=h, ppi=c["I"][h], init=False, ack=False))
            return
        # final hop, :661-791
        if vis["ppi"] != c["I"][h - 1]:
            return
        if not c["valid_cycle"]:
            if v == vis["target"]:
                return
            self._send(dict(vertex=vis["target"], parent=v, target=vis["target"], itr=vis["itr"], ppi=0,
                            init=False, ack=True))
        elif v == vis["target"]:
            if v not in self.token_source_map:
                self.errors.append(("tp_cycle_missing", v))
                return
            self.token_source_map[v] = 1
            if vis["parent"] not in self.E[v]:  # :764-770
                self.errors.append(("tp_edge_missing", v, vis["parent"]))
            else:
                self.E[v][vis["parent"]] = 1

    # ------------------------------------------------------------------ TDS
    def hist_rule(self, hist, hp, x):
        e = self.c["enum"]
        if hp >= len(e):
            return False
        if e[hp] == hp:
            return x not in hist[:hp]
        if e[hp] < hp:
            return hist[e[hp]] == x
        return False

    def tds_pre_visit(self, vis):
        """tds_batch_1.hpp:122-335"""
        c = self.c
        v = vis["vertex"]
        if not self.active[v]:
            return False
        if vis["ack"]:
            return True
        h = vis["itr"] + 1
        if not self.tp_static(v, h) or vis["ppi"] != c["I"][h - 1]:  # :207-254
            return False
        if c["C"] > vis["itr"]:  # :260-304
            if not self.hist_rule(vis["hist"], h, v):
                return False
        return True

    def tds_visit(self, vis):
        """tds_batch_1.hpp:347-919"""
        c = self.c
        v = vis["vertex"]
        if not self.active[v]:
            return
        if vis["ack"]:
            self.token_source_map[v] = 1
            return
        if vis["init"]:  # :425-512
            if v not in self.token_source_map or self.label[v] != c["P"][0]:
                return
            tv = self.template_vertices[v]
            if tv == 0 or not (tv >> c["I"][0]) & 1:
                return
            for n in list(self.E[v].keys()):
                self._send(dict(vertex=n, parent=v, target=v, itr=0, ppi=c["I"][0], init=False, ack=False,
                                hist=[v, n]))
            return
        h = vis["itr"] + 1
        hist = vis["hist"]
        if not self.tp_static(v, h):
            return
        if c["C"] > vis["itr"]:  # interior, :588-639 then :793-909
            if vis["ppi"] != c["I"][h - 1]:
                return
            if not self.hist_rule(hist, h, v):
                return
            for n in list(self.E[v].keys()):
                if c["C"] == h:
                    if c["valid_cycle"]:
                        if n != vis["target"]:
                            continue
                    else:
                        if n == vis["target"]:
                            continue
                        if not self.hist_rule(hist, h + 1, n):
                            continue
                elif not self.hist_rule(hist, h + 1, n):
                    continue
                self._send(dict(vertex=n, parent=v, target=vis["target"], itr=h, ppi=c["I"][h], init=False,
                                ack=False, hist=hist[: h + 1] + [n]))
            return
        # final hop, :641-754
        if vis["ppi"] != c["I"][h - 1]:
            return
        if not c["valid_cycle"]:
            if v == vis["target"]:
                return
            self._send(dict(vertex=vis["target"], parent=v, target=vis["target"], itr=vis["itr"], ppi=0,
                            init=False, ack=True))
            self.subgraphs[self.pl].append(tuple(hist[: h + 1]))
            self.path_count += 1
        elif v == vis["target"] and v == hist[0]:
            self.token_source_map[v] = 1
            self.subgraphs[self.pl].append(tuple(hist[: h + 1]))
            self.path_count += 1
        else:
            self.errors.append(("tds_wrong_branch", v))

    # ----------------------------------------------------------- outer loop
    def run(self):
        """beta.cpp:544-1351"""
        self.init_step = True
        self.itr = 0
        self.path_count = 0
        while True:
            self.not_finished = False
            self.lcc()
            self.init_step = False
            if self.itr == 0:
                self.not_finished = True
            if self.not_finished:
                self.not_finished = False
                for pl, c in enumerate(self.pg.cons):
                    self.c, self.pl = c, pl
                    self.token_source_map = {}
                    for s in self.token_source_set:
                        s.clear()
                    self.subgraphs[pl] = []  # file truncated, beta.cpp:713-717
                    init = [dict(vertex=v, init=True, ack=False, itr=0) for v in range(self.V)]
                    if pl >= self.tds_from_pl >= 0:
                        for v in range(self.V):  # tds_batch_1.hpp:1067-1100
                            tv = self.template_vertices[v]
                            if self.active[v] and self.label[v] == c["P"][0] and tv and (tv >> c["I"][0]) & 1:
                                self.token_source_map[v] = 0
                        self.traverse(init, self.tds_pre_visit, self.tds_visit)
                    else:
                        self.traverse(init, self.nem_pre_visit, self.nem_visit)
                    deleted = False
                    for s, okf in self.token_source_map.items():  # beta.cpp:964-1005
                        if okf:
                            continue
                        tv = self.template_vertices[s]
                        if tv == 0:
                            continue
                        if (tv >> c["I"][0]) & 1:
                            tv &= ~(1 << c["I"][0])
                            self.template_vertices[s] = tv
                        if tv == 0:
                            self.active[s] = False
                        self.not_finished = True
                        deleted = True
                    for s in self.token_source_map:  # :1043-1062
                        if not self.active[s] and s in self.state_map:
                            del self.state_map[s]
                    nv, ne = self.counts()
                    self.rows.append((self.itr, "TP", pl, nv, ne))
                    if deleted and c["interleave"]:
                        self.lcc()
            self.itr += 1
            if not self.not_finished or self.itr >= self.max_iterations:
                break
        self.iterations = self.itr

    # -------------------------------------------------------------- results
    def final_vertices(self):
        return sorted((v, self.template_vertices[v]) for v in self.state_map)

    def final_edges(self):
        return sorted((v, n) for v in self.state_map for n in self.E[v])
