#!/usr/bin/env python
"""bench.py — LCC+NLCC search throughput on synthetic R-MAT graphs (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA engine
  python bench.py --impl reference --gpus N ...          # reference CPU path (oracle port) on host cores

One "step" = one full search (per-pattern state reset, LCC supersteps, NLCC token
walks, until the reference's loop terminates) of every template of the workload
over one resident R-MAT graph.  `value` = directed edge slots of the graph
(16 * 2^scale generated edges, both directions) * templates per step / step time:
input edges searched per second.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from fuzzypatternmatching_b200 import patterns as PT  # noqa: E402

# BASELINE.json configs[2]: R-MAT with cyclic templates (triangle, 4-cycle, 6-vertex
# cycle + chords) exercising NLCC token walks; labels = degree classes ceil(log2(d+1)).
# Interior hop labels are pairwise distinct, so the reference result is independent of
# message order.  The README's tree template rides along as template 0.
WORKLOADS = {
    "cyclic": [("tree", PT.RMAT_LOG2_TREE), ("triangle_678", PT.triangle(6, 7, 8)),
               ("cycle4_5678", PT.cycle4(5, 6, 7, 8)), ("cycle6_chords_456789", PT.cycle6_chords([4, 5, 6, 7, 8, 9]))],
    "tree": [("tree", PT.RMAT_LOG2_TREE)],
    # hub classes (degree labels 12, 14, 16: rows of 2^11 .. 2^16 slots): the CTA-per-row kernels carry the scans;
    # cycle checking only (the enumeration walk over hub neighbourhoods is combinatorial)
    "hubs": [("triangle_hubs_12_14_16", dict(PT.triangle(12, 14, 16), constraints=PT.triangle(12, 14, 16)["constraints"][:1]))],
}


def write_patterns(workload):
    base = tempfile.mkdtemp(prefix="pm_bench_")
    out = []
    for name, spec in WORKLOADS[workload]:
        d = PT.write_pattern_dir(os.path.join(base, name), spec)
        out.append((name, d, PT.tds_from_pl(spec)))
    return out


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled every 20 ms from a
    thread (nvidia-smi -lms needs most of a second to start, longer than a short timed region); nvidia-smi is the
    fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []
        self.sm, self.mask, self.stop_flag, self.t, self.max_mhz, self.nvml = [], 0, False, None, None, None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.idx])
            except (ValueError, IndexError):
                pass
        return self.idx

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml = (pynvml, h)
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        pynvml, h = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                self.mask |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                pass
            time.sleep(0.02)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.t.join(timeout=1.0)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(n for b, n in self.REASONS.items() if self.mask & b), "samples": len(sm),
                    "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference(scale, gen_ranks, pats, steps, warmup, budget_s=150.0):
    """The reference's CPU path for the same templates: the oracle port (the reference
    binary needs MPI + Boost and cannot be built on this image), all host threads."""
    from oracle import oracle as O
    threads = host_threads()
    g = O.Graph.rmat(scale, gen_ranks, threads)
    labels = g.labels_degree_log2()
    ps = [(O.Pattern(d), tds) for _, d, tds in pats]

    def one_step():
        t = 0.0
        for p, tds in ps:
            r = O.Run(g, labels, p, tds_from_pl=tds, threads=threads, keep_subgraphs=False)
            t += r.search_seconds
        return t

    t_begin = time.time()
    for _ in range(warmup):
        one_step()
        if time.time() - t_begin > budget_s / 3:
            break
    times = []
    for _ in range(steps):
        times.append(one_step())
        if time.time() - t_begin > budget_s:
            break
    edges = g.n_slots_multi * len(ps)
    return {"value": edges * len(times) / sum(times), "seconds_per_step": sum(times) / len(times),
            "steps_done": len(times), "cores": threads, "scale": scale, "edges_per_step": edges}


def cpu_reference_binary(scale, gen_ranks, workload, steps, warmup, budget_s=150.0):
    """The reference's OWN code for the same templates: oracle/_ref/run_pattern_matching_beta — the reference driver and
    visitor headers compiled from /root/reference over the single-rank runtime stand-in of oracle/ref_shim (the binary
    travels with the repository).  One rank is all that runtime gives a process, so the host cores are used the way an
    MPI job would use them, minus the communication: one independent instance per host thread, each searching its own
    copy of the sample graph; the value is the aggregate.  Returns None where the binary is absent."""
    from oracle import oracle as O
    from oracle import reference_run as R
    if not R.available():
        return None
    import shutil
    import numpy as np
    copies = host_threads()
    work = tempfile.mkdtemp(prefix="pm_bench_ref_")
    try:
        per = (16 << scale) // gen_ranks
        edges = np.concatenate([O.rmat_stream(scale, r, per) for r in range(gen_ranks)])
        both = np.empty((2 * len(edges), 2), dtype=np.uint64)  # "(s, t) then (t, s)" per generated edge, like ingest -u 1
        both[0::2] = edges
        both[1::2] = edges[:, ::-1]
        graph = os.path.join(work, "graph.slots")
        with open(graph, "w") as f:
            f.write("%d\n" % (1 << scale))
            np.savetxt(f, both, fmt="%d")
        pdirs = []
        for name, spec in WORKLOADS[workload]:
            spec4, _ = R.tds_at_constraint_4(spec)
            pdirs.append(os.path.dirname(PT.write_pattern_dir(os.path.join(work, "pattern_" + name), spec4)))
        run_no = [0]

        def one_step():
            total = 0.0
            for pd in pdirs:
                procs = []
                for c in range(copies):
                    run_no[0] += 1
                    procs.append(R.launch(graph, pd, os.path.join(work, "out_%d" % (run_no[0] % (2 * copies)))))
                slowest = 0.0
                for p in procs:
                    out, err = p.communicate()
                    if p.returncode != 0:
                        raise RuntimeError("reference driver failed: " + err[-500:])
                    slowest = max(slowest, R.search_seconds(out))
                total += slowest
            return total

        t_begin = time.time()
        for _ in range(warmup):
            one_step()
            if time.time() - t_begin > budget_s / 3:
                break
        times = []
        for _ in range(steps):
            times.append(one_step())
            if time.time() - t_begin > budget_s:
                break
        edges_per_step = 2 * len(edges) * len(pdirs)
        return {"value": copies * edges_per_step * len(times) / sum(times), "seconds_per_step": sum(times) / len(times),
                "steps_done": len(times), "cores": copies, "scale": scale, "edges_per_step": edges_per_step, "copies": copies}
    finally:
        shutil.rmtree(work, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=int, default=int(os.environ.get("PM_BENCH_SCALE", "26")))
    ap.add_argument("--gen-ranks", type=int, default=1024, help="generating ranks of the R-MAT stream (part of the graph's identity)")
    ap.add_argument("--workload", default="cyclic", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-scale", type=int, default=int(os.environ.get("PM_BENCH_CPU_SCALE", "20")))
    ap.add_argument("--ref-scale", type=int, default=int(os.environ.get("PM_BENCH_REF_SCALE", "16")),
                    help="R-MAT scale of the sample the reference's own binary (oracle/_ref) is timed on")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    pats = write_patterns(args.workload)
    metric, unit = "lcc_nlcc_search_edges_per_second", "edges/s"
    # Weak scaling: the per-GPU share of the graph stays that of R-MAT scale `--scale` on one GPU, so N GPUs
    # search ONE R-MAT graph of scale + log2(N) partitioned 1-D (owner(v) = v mod N); N = 4 is BASELINE
    # configs[4] (scale 28).  PM_BENCH_STRONG=1 keeps the scale fixed instead.
    strong = os.environ.get("PM_BENCH_STRONG", "0") == "1"
    scale = args.scale if (strong or world == 1) else args.scale + max(0, (world - 1).bit_length())
    config = {"workload": "rmat_s%d_%s" % (scale, args.workload), "scale": scale,
              "gen_ranks": args.gen_ranks, "templates": [n for n, _, _ in pats],
              "labels": "ceil(log2(degree+1))", "parallelism": "1d_vertex_partition_x%d" % args.gpus}

    # ---------------------------------------------------------------- reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        # the reference's own code where its binary is here (oracle/_ref, built in the container that holds the reference
        # tree and shipped with the repository); PM_BENCH_REF_PORT=1 or a missing / failing binary: the oracle port
        r, kind = None, "reference"
        if os.environ.get("PM_BENCH_REF_PORT", "0") != "1":
            try:
                cpu_scale, cpu_gen = min(scale, args.ref_scale), 4
                r = cpu_reference_binary(cpu_scale, cpu_gen, args.workload, args.steps, max(args.warmup, 1))
            except Exception as e:  # noqa: BLE001 — the arm must still print its line
                print("reference binary leg failed, falling back to the oracle port: %s" % e, file=sys.stderr)
                r = None
        if r is not None:
            sample = ("the reference's own driver + visitor headers (oracle/_ref/run_pattern_matching_beta, single-rank runtime "
                      "stand-in oracle/ref_shim), %d independent instances (one per host thread), each on R-MAT scale %d with %d "
                      "generating ranks (bounded sample of the scale-%d workload), same templates; its own clock around each "
                      "template's search loop, slowest instance per template" % (r["copies"], cpu_scale, cpu_gen, scale))
            par = "independent_single_rank_instances_x%d" % r["copies"]
        else:
            kind = "port"
            cpu_scale = min(scale, args.cpu_scale)
            cpu_gen = min(args.gen_ranks, 4) if cpu_scale != scale else args.gen_ranks
            r = cpu_reference(cpu_scale, cpu_gen, pats, args.steps, max(args.warmup, 1))
            sample = ("oracle port of the reference CPU path, R-MAT scale %d with %d generating ranks (bounded sample of the "
                      "scale-%d workload), same templates, %d host threads" % (cpu_scale, cpu_gen, scale, r["cores"]))
            par = "openmp_x%d" % r["cores"]
        # the config states what RAN; the workload it samples stays named beside it
        config.update({"scale": cpu_scale, "gen_ranks": cpu_gen, "directed_edge_slots": r["edges_per_step"] // len(pats),
                       "sample_of": {"workload": config["workload"], "scale": scale, "gen_ranks": args.gen_ranks},
                       "parallelism": par})
        line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": args.gpus,
                "steps": r["steps_done"], "warmup": args.warmup, "ms_per_step": r["seconds_per_step"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": kind, "sample": sample},
                "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------------- our arm
    import numpy as np
    import torch
    import torch.distributed as dist
    from fuzzypatternmatching_b200.engine import Engine
    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gloo = dist.new_group(backend="gloo")  # object collectives of the parity check
    torch.cuda.set_device(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = Engine(local_rank)
    if world > 1:
        # one engine context per GPU; the NCCL id of the engine's own communicator travels through torch.distributed
        ids = [Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0, device=torch.device("cuda", local_rank))
        eng.comm_init(rank, world, ids[0])
    t0 = time.time()
    # every rank generates its share of the generating ranks' streams; the slots are shuffled to their owners
    eng.graph_rmat(scale, args.gen_ranks)
    torch.cuda.synchronize()
    t_lab = time.perf_counter()
    eng.labels_degree_log2()   # degree labels + the per-graph index the search reads (byte labels, label stream, signatures)
    torch.cuda.synchronize()
    index_ms = (time.perf_counter() - t_lab) * 1e3
    gi = eng.graph_info()
    build_s = time.time() - t0
    n_local_edges = gi["n_slots_multi"]
    if world > 1:
        t = torch.tensor([n_local_edges], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        n_global_edges = int(t.item())
    else:
        n_global_edges = n_local_edges
    n_edges = n_global_edges * len(pats)      # whole-job edges searched per step

    step_acc = {"edges_scanned": 0, "algorithmic_bytes": 0, "device_seconds": 0.0}

    def one_step(fetch=False):
        got = 0
        step_acc.update(edges_scanned=0, algorithmic_bytes=0, device_seconds=0.0)
        for _, d, tds in pats:
            eng.pattern_load_dir(d)
            s = eng.run(tds_from_pl=tds, keep_subgraphs=False)
            step_acc["edges_scanned"] += int(s["edges_processed"])
            step_acc["algorithmic_bytes"] += int(s["algorithmic_bytes"])
            step_acc["device_seconds"] += float(s["device_seconds"])
            if fetch:  # device -> host read of the step's result
                v, b = eng.active_vertices()
                e = eng.active_edges()
                got += v.nbytes + b.nbytes + e.nbytes
        return got

    for _ in range(args.warmup):
        one_step()
    ks0 = [eng.kernel_stats(b) for b in range(5)]
    l0 = eng.kernel_launches()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    barrier()
    elapsed = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = eng.kernel_launches() - l0
    ks1 = [eng.kernel_stats(b) for b in range(5)]
    if world > 1:
        t = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    value = n_edges * args.steps / elapsed
    # what the kernels themselves counted over one step (pm_run_summary_t): adjacency slots walked by the LCC scans +
    # token fan-out of the NLCC walks, and the SURVEY section 8(d) byte model over them; several ranks: summed
    acc = [step_acc["edges_scanned"], step_acc["algorithmic_bytes"]]
    if world > 1:
        t = torch.tensor(acc, dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        acc = [int(x) for x in t.tolist()]
    ms_step = elapsed / args.steps * 1e3

    # roofline of the dominant kernel: the LCC scan class (first superstep / later supersteps / CTA-per-row)
    # with the largest share of the timed region, timed with CUDA events on the engine's stream
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dk = [dict((k, ks1[b][k] - ks0[b][k]) for k in ks1[b]) for b in range(5)]
    scan_classes = [0, 1, 2, 4]                       # 3 is the init filter (no adjacency walked)
    alg = lambda b: dk[b]["slots"] * 6.25 + dk[b]["vertices"] * 12.25  # noqa: E731  SURVEY section 8(d)
    longest = max(dk[b]["ms"] for b in scan_classes)
    # the dominant class: largest CUDA-event time; classes within 3 % of the largest are tied (run-to-run noise is of
    # that order) and the one that moves more algorithmic bytes is named.  Every class is printed below it.
    top = max((b for b in scan_classes if dk[b]["ms"] >= 0.97 * longest), key=alg)
    names = {0: "k_lcc_scan<FIRST=true> (first superstep: pristine adjacency + label stream)",
             1: "k_lcc_scan<FIRST=false> (later supersteps: active edge maps + mask gathers)",
             2: "k_lcc_scan_big (CTA per row)",
             4: "k_lcc_scan<XLATE=true> (second superstep: mask gathers + renaming slots to compact ids)"}
    traffic = None
    try:  # DRAM bytes per launch of the same kernel class from the committed ncu --set full capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        # only a capture of THIS workload on this many GPUs describes this run
        tr = tr.get("%s_n%d" % (config["workload"], world), {})
        traffic = tr.get({0: "scan_first", 1: "scan_later", 2: "scan_big", 4: "scan_xlate"}[top], {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roof = None
    if dk[top]["launches"] and dk[top]["ms"] > 0:
        alg_bytes = dk[top]["slots"] * 6.25 + dk[top]["vertices"] * 12.25
        achieved = alg_bytes / dk[top]["launches"] / (dk[top]["ms"] / dk[top]["launches"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": names[top], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback",
                "algorithmic_bytes_per_launch": alg_bytes / dk[top]["launches"],
                "avg_launch_ms": dk[top]["ms"] / dk[top]["launches"], "launches": dk[top]["launches"],
                "share_of_step": dk[top]["ms"] * 1e-3 / elapsed,
                "other_classes": {names[b].split(" ")[0]: {"ms": dk[b]["ms"], "launches": dk[b]["launches"],
                                                              "frac": (alg(b) / (dk[b]["ms"] * 1e-3) / 1e9 / peak) if dk[b]["ms"] > 0 else None}
                                  for b in scan_classes if b != top},
                "dominant_rule": "largest CUDA-event time over the timed region; classes within 3 % of it are tied and the "
                                 "one with more algorithmic bytes is named",
                "model": "6.25 B per scanned slot + 12.25 B per scanned vertex"}

    # ---- several GPUs: correctness of THIS build on THIS many ranks, outside the timed region.  R-MAT scale 18 is
    # searched by the partitioned engine and by the CPU oracle with n_ranks = N (the checker, tests/multi_gpu_check.py);
    # per-rank rows, vertex / edge lists and enumerated subgraphs must be equal.
    parity = None
    strong_rec = None
    if world > 1:
        from tests.multi_gpu_check import partition_parity

        class _Gloo:
            @staticmethod
            def broadcast_object_list(box, src=0):
                dist.broadcast_object_list(box, src=src, group=gloo)

            @staticmethod
            def gather_object(obj, lst, dst=0):
                dist.gather_object(obj, lst, dst=dst, group=gloo)

        from oracle import oracle as O  # the checker, rank 0 only uses it
        oks, names = [], []
        for nm, spec in WORKLOADS[args.workload]:
            tds = PT.tds_from_pl(spec)
            ok = partition_parity(eng, _Gloo, rank, world, "rmat18/" + nm, lambda: eng.graph_rmat(18, 4),
                                  lambda: O.Graph.rmat(18, 4), None, spec, tds, log=lambda m: sys.stderr.write(m + "\n"))
            oks.append(ok)
            names.append(nm)
        if rank == 0:
            parity = {"ranks": world, "ok": bool(all(oks)), "graph": "rmat scale 18, 4 generating ranks",
                      "templates": names, "checker": "CPU oracle with n_ranks = %d" % world}
        # ---- strong scaling beside the weak-scaling value: the N=1 graph (scale `--scale`) over N GPUs
        if not strong:
            eng.graph_rmat(args.scale, args.gen_ranks)
            eng.labels_degree_log2()
            for _ in range(args.warmup):
                one_step()
            barrier()
            t0s = time.perf_counter()
            for _ in range(args.steps):
                one_step()
            barrier()
            el = time.perf_counter() - t0s
            t = torch.tensor([el], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
            ne = 2 ** (args.scale + 5) * len(pats)
            strong_rec = {"scale": args.scale, "n_gpus": world, "ms_per_step": el / args.steps * 1e3,
                          "value": ne * args.steps / el, "unit": unit,
                          "what": "the N=1 graph (scale %d) partitioned over %d GPUs" % (args.scale, world)}
            # back to the weak-scaling graph for the end-to-end leg
            eng.graph_rmat(scale, args.gen_ranks)
            eng.labels_degree_log2()

    # end to end through the C ABI with HOST buffers: upload the host CSR, search, read results back
    e2e = None
    if world > 1:
        args.e2e_steps = min(args.e2e_steps, 1)  # reading the rows back for the upload is a host loop per rank: keep the run short
    if (rank == 0 or world > 1) and args.e2e_steps > 0:
        rowptr, col = eng.graph_csr()
        degm = eng.graph_degree()
        pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
        rowptr, col, degm = pin(rowptr), pin(col), pin(degm)
        h2d = rowptr.nbytes + col.nbytes + degm.nbytes
        if world > 1:
            t = torch.tensor([h2d], dtype=torch.int64, device="cuda")
            dist.all_reduce(t)
            h2d = int(t.item())
        d2h = 0
        # one untimed end-to-end step first (like the warm-up of the resident arm: the first upload sizes the
        # device block cache and the NLCC scratch)
        eng.graph_from_csr(rowptr, col, degm, n_vertices=gi["n_vertices"])
        eng.labels_degree_log2()
        one_step(fetch=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            eng.graph_from_csr(rowptr, col, degm, n_vertices=gi["n_vertices"])
            eng.labels_degree_log2()
            d2h = one_step(fetch=True)
        barrier()
        e2e_elapsed = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_elapsed], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_elapsed = float(t.item())
        e2e = {"value": n_edges * args.e2e_steps / e2e_elapsed, "unit": unit,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_elapsed / args.e2e_steps * 1e3,
               "steps": args.e2e_steps,
               "what": "pm_graph_from_csr (pinned host CSR -> device store) + degree labels + search of every template + "
                       "device->host read of the final active vertex and edge lists"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = None
        if os.environ.get("PM_BENCH_REF_PORT", "0") != "1":
            try:
                cs = min(scale, args.ref_scale)
                r = cpu_reference_binary(cs, 4, args.workload, 2, 1, budget_s=45.0)
            except Exception as e:  # noqa: BLE001 — the bench line must still be printed
                print("reference binary leg failed, falling back to the oracle port: %s" % e, file=sys.stderr)
                r = None
        if r is not None:
            cpu = {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "reference",
                   "sample": "the reference's own driver + visitor headers (oracle/_ref, single-rank runtime stand-in), %d "
                             "independent instances (one per host thread) on R-MAT scale %d, same templates, %d steps"
                             % (r["copies"], cs, r["steps_done"])}
        else:
            cs = min(scale, args.cpu_scale)
            r = cpu_reference(cs, 4 if cs != scale else args.gen_ranks, pats, 2, 1, budget_s=60.0)
            cpu = {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": "port",
                   "sample": "oracle port on R-MAT scale %d, same templates, %d steps, %d host threads"
                             % (cs, r["steps_done"], r["cores"])}

    if rank == 0:
        free_b, total_b = torch.cuda.mem_get_info()
        config.update({"inputs_exceed_l2": gi["n_slots_padded"] * 4 > 126e6, "graph_build_seconds": build_s,
                       "device_bytes_in_use_rank0": int(total_b - free_b), "graph_store_bytes_rank0": gi["device_bytes"],
                       "directed_edge_slots": n_global_edges, "distinct_slots_rank0": gi["n_slots"],
                       "multi_gpu": ("one graph, 1-D partition owner(v) = v mod %d, per-GPU share = scale %d" % (world, args.scale))
                       if world > 1 else "single gpu"})
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
                "scaling": "strong" if (strong and world > 1) else "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic", "config": config,
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
                "search_ms_per_template": elapsed / args.steps / len(pats) * 1e3,
                # the kernels' own counters for one step, and the whole step against the HBM roofline
                "edges_scanned_per_step": acc[0], "algorithmic_bytes_per_step": acc[1],
                "edges_scanned_per_second": acc[0] / (ms_step * 1e-3),
                "step_roofline_frac": acc[1] / (ms_step * 1e-3) / 1e9 / (peak * world),
                "launches_per_step": launches / args.steps,
                # the per-graph index (byte labels, label stream, neighbour-label signatures) is built with the labels,
                # outside the timed search like the reference's label pass; `value_incl_index` charges one build per step
                "index_build_ms": index_ms,
                "value_incl_index": n_edges / ((ms_step + index_ms) * 1e-3),
                "value_definition": "directed edge slots of the graph x templates / step time (input edges searched per second)"}
        if parity is not None:
            line["parity_check"] = parity
        if strong_rec is not None:
            line["strong_s%d" % args.scale] = strong_rec
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
