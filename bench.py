#!/usr/bin/env python3
"""bench.py -- the driver's measurement contract for the two ProGraphMSA hot paths on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched under torch.distributed.run, one rank per GPU)

Workload (BASELINE.json configs[1]): the all-pairs NW + ML distance matrix of a synthetic
500-sequence x 400-aa WAG family (124,750 pairs, 2.0e10 DP cells).  One step = one full matrix.
  value   NW cell updates per second (GCUPS), all stages included (forward DP with packed traceback,
          traceback walk, ML Newton), sequences already resident in HBM, results left in HBM
  e2e     the same through the C-ABI call with HOST buffers (H2D of the sequences and D2H of the
          distance/variance entries inside the timed region)
N > 1 shards the pair range over the ranks (no data-path collective while computing) and gathers the
(distance, variance) entries with one NCCL all-gather -- strong scaling of a fixed matrix.
At N = 1 the line also carries `graph_dp`: the graph-DP GCUPS half of BASELINE.json's metric -- `exact` / `tensor`: the 1020 jobs
of a whole progressive pass in one batch (latency of its few long sweeps included); `leaf_level`: the 2500 leaf x leaf jobs at the
bottom of BASELINE configs[4] (the fill kernels' throughput) -- and `other_kernels`: the ML kernel against the FP64 rate measured in
the run, the Prealigned path on the bench family and on rows shaped like configs[4]'s alignment.

--impl reference times the reference's own CPU binary (oracle/_ref/ProGraphMSA_64, single-threaded by
design; one independent instance per host core) on bounded samples of the same family.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

N_SEQ, SEQ_LEN, SEED = 500, 400, 1
GRAPH_N_SEQ = 256
GRAPH_REPLICAS = 4
LEAF_LEVEL_JOBS = 2500      # bottom level of BASELINE configs[4] (5000 sequences x 500 aa)
LEAF_LEVEL_LEN = 500
WORKLOAD = "all-pairs NW+ML distance matrix, %d synthetic protein seqs x %d aa, WAG, -a -m" % (N_SEQ, SEQ_LEN)
METRIC = "all-pairs NW+ML distance-matrix throughput (graph-DP GCUPS in graph_dp)"
UNIT = "GCUPS"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback (B200_PROFILING.md)")


def workload_config(seqs, names):
    """`config` of the JSON line: the same dict in both arms (ours and --impl reference), so the driver compares like with like"""
    L = np.array([len(seqs[k]) for k in names], dtype=np.int64)
    cells = 0
    for i in range(len(L)):
        cells += int(((L[i] + 1) * (L[i + 1:] + 1)).sum())
    return dict(workload=WORKLOAD, n_seq=len(L), pairs=len(L) * (len(L) - 1) // 2, nw_cells=cells,
                l2="GPU arm: 256 MiB flush write between timed steps, and the per-step working set (7 GB of packed traceback) >> 126 MB L2; "
                   "CPU arm: not applicable")


def make_family():
    from prographmsa_b200 import model, synth
    from prographmsa_b200.options import Options
    opts = Options(mldist_flag=True, nwdist_flag=True)
    mf = model.getDefault(opts)
    seqs, root = synth.simulate_family(N_SEQ, SEQ_LEN, mf, seed=SEED)
    names = sorted(seqs)
    return opts, mf, seqs, names, root


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own binary on the host cores
# ------------------------------------------------------------------------------------------------------
def _fasta(seqs, names):
    from prographmsa_b200.alphabet import AA_LETTERS
    return "".join(">%s\n%s\n" % (k, "".join(AA_LETTERS[c] for c in seqs[k])) for k in names)


def _run_ref_instances(seqs, names, groups):
    """run one binary instance per group (each group = list of sequence names), in parallel; -> (wall s, cells, pairs)"""
    from oracle import pyoracle as po
    binary = po.ref_binary()
    cells = pairs = 0
    with tempfile.TemporaryDirectory() as d:
        procs = []
        for gi, g in enumerate(groups):
            path = os.path.join(d, "g%d.fa" % gi)
            with open(path, "w") as f:
                f.write(_fasta(seqs, g))
            L = [len(seqs[k]) for k in g]
            for a in range(len(g)):
                for b in range(a + 1, len(g)):
                    cells += (L[a] + 1) * (L[b] + 1); pairs += 1
        t0 = time.perf_counter()
        for gi in range(len(groups)):
            procs.append(subprocess.Popen([binary, "-a", "-m", "-i", "0", "-T", "g%d.fa" % gi], cwd=d,
                                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("reference binary failed")
        wall = time.perf_counter() - t0
    return wall, cells, pairs


def _port_sample(opts, mf, seqs, names, k):
    """fallback when the binary is absent: the oracle port, 1 thread"""
    from oracle import pyoracle as po
    from prographmsa_b200 import alphabet, distance
    tab, go, ge = distance.nw_table(alphabet.AA)
    sub = [alphabet.AA.nw_codes(seqs[x]) for x in names[:k]]
    t0 = time.perf_counter()
    po.nw_ml_all_pairs(sub, 20, tab, go, ge, mf.Q, mf.freqs, opts)
    wall = time.perf_counter() - t0
    L = [len(s) for s in sub]
    cells = sum((L[a] + 1) * (L[b] + 1) for a in range(k) for b in range(a + 1, k))
    return wall, cells, k * (k - 1) // 2


def cpu_baseline(opts, mf, seqs, names):
    from oracle import pyoracle as po
    if po.ref_binary():
        k = 144                                      # 10,296 pairs ~ 1.7e9 cells: ~10 s on one core of the GPU box
        wall, cells, pairs = _run_ref_instances(seqs, names, [names[:k]])
        return dict(value=cells / wall / 1e9, unit=UNIT, cores=1, kind="reference",
                    sample="oracle/_ref/ProGraphMSA_64 -a -m -i 0 -T on the first %d of the %d sequences (%d pairs, %.3g cells, %.1f s wall, "
                           "incl. its BioNJ of %d taxa)" % (k, N_SEQ, pairs, cells, wall, k))
    k = 24
    wall, cells, pairs = _port_sample(opts, mf, seqs, names, k)
    return dict(value=cells / wall / 1e9, unit=UNIT, cores=1, kind="port",
                sample="oracle port on the first %d sequences (%d pairs, %.1f s)" % (k, pairs, wall))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    opts, mf, seqs, names, _ = make_family()
    cores = os.cpu_count() or 1
    inst = max(1, min(cores, 64))
    per = 24                                         # 276 pairs ~ 4.4e7 cells ~ 1 s per instance per step
    per = min(per, max(2, N_SEQ // inst))
    rng = np.random.default_rng(0)

    def groups():
        perm = rng.permutation(N_SEQ)
        return [[names[i] for i in perm[g * per:(g + 1) * per]] for g in range(inst)]

    if not po.ref_binary():
        wall, cells, pairs = _port_sample(opts, mf, seqs, names, 24)
        val = cells / wall / 1e9
        line = dict(metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=1, warmup=0, ms_per_step=wall * 1e3,
                    higher_is_better=True, scaling="strong", vs_baseline=None, dtype="int32+f64", data="synthetic",
                    config=workload_config(seqs, names), impl="reference",
                    cpu_baseline=dict(value=val, unit=UNIT, cores=1, kind="port", sample="oracle port, first 24 sequences"),
                    e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        emit(line); return
    for _ in range(args.warmup):
        _run_ref_instances(seqs, names, groups())
    tot_wall = tot_cells = 0.0
    tot_pairs = 0
    for _ in range(args.steps):
        wall, cells, pairs = _run_ref_instances(seqs, names, groups())
        tot_wall += wall; tot_cells += cells; tot_pairs += pairs
    val = tot_cells / tot_wall / 1e9
    sample = ("%d parallel single-threaded instances of oracle/_ref/ProGraphMSA_64 (-a -m -i 0 -T), each on a random %d-sequence "
              "subsample of the family (%d pairs per step in total)" % (inst, per, tot_pairs // max(1, args.steps)))
    line = dict(metric=METRIC, value=val, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=tot_wall / args.steps * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype="int32+f64", data="synthetic", config=workload_config(seqs, names), impl="reference",
                cpu_baseline=dict(value=val, unit=UNIT, cores=inst, kind="reference", sample=sample),
                e2e=dict(value=val, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def graph_dp_bench(ctx, torch, steps, warmup, pk):
    """graph-DP GCUPS on 1 GPU: every internal-node alignGraphs job of a progressive pass over a
    GRAPH_N_SEQ x 400 aa family (leaf x leaf up to profile x profile graphs), staged once and run as
    one batch (value), and through alignGraphsBatch with host graphs (e2e)."""
    from prographmsa_b200 import model, synth, progressive, graphalign
    from prographmsa_b200.options import Options
    opts = Options(mldist_flag=True)
    mf = model.getDefault(opts)
    seqs, root = synth.simulate_family(GRAPH_N_SEQ, SEQ_LEN, mf, seed=SEED + 1)
    tree = progressive.from_synth(root)
    jobs = []

    def batch(js, o):
        jobs.extend(js)
        return graphalign.alignGraphsBatch(js, o, ctx=ctx)
    progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch)
    base_jobs = len(jobs)
    jobs = jobs * GRAPH_REPLICAS                      # the batch a 1000-sequence level-by-level pass would offer
    out = {}
    for mode_name, mode in (("exact", graphalign.EMISSION_EXACT), ("tensor", graphalign.EMISSION_TENSOR)):
        try:
            gb = graphalign.GraphBatch(jobs, opts, mode, ctx)
            gb.run(); ctx.sync()
        except Exception as e:      # tensor flavour may be unavailable in an early round
            out[mode_name] = dict(unavailable=str(e)); continue
        try:
            flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
            ctx.enable_timing(True)
            for _ in range(warmup):
                gb.run(); ctx.sync()
            def timed_runs(acc):
                ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
                u = cl = 0
                for a, b in ev:
                    flush.fill_(1); torch.cuda.synchronize()
                    a.record(); gb.run(); ctx.sync(); b.record(); torch.cuda.synchronize()
                    t = ctx.timing()
                    for k in ("emission_ms", "graph_fill_ms", "graph_trace_ms"):
                        acc[k] += t[k]
                    acc["launches"] += t["emission_launches"] + t["graph_fill_launches"] + t["graph_trace_launches"]
                    u, cl = t["graph_cell_updates"], t["graph_cells"]
                return ev, u, cl
            # per-kernel breakdown: the three phases one after the other over the whole batch
            os.environ["PGM_GRAPH_PIPELINE"] = "0"
            gb.run(); ctx.sync()
            tm = dict(emission_ms=0.0, graph_fill_ms=0.0, graph_trace_ms=0.0, launches=0)
            ev0, upd, cells = timed_runs(tm)
            ms_phased = sum(a.elapsed_time(b) for a, b in ev0) / steps
            # the product's schedule (and the reported value): every size class its own emission -> fill -> traceback pipeline
            os.environ.pop("PGM_GRAPH_PIPELINE")
            gb.run(); ctx.sync()
            tp = dict(emission_ms=0.0, graph_fill_ms=0.0, graph_trace_ms=0.0, launches=0)
            ev, upd, cells = timed_runs(tp)
            ms = sum(a.elapsed_time(b) for a, b in ev) / steps
            res = gb.fetch()
            # e2e: host graphs in, mappings out, per step (one untimed call first: buffer pools, side streams).  Two figures: through the
            # Python mirror of alignGraphs (alignGraphsBatch: marshals the graphs, builds result objects), and the C-ABI call alone
            # (pgm_align_graphs_batch on prebuilt pgm_graph / pgm_alignment arrays in host memory: pack, H2D, kernels, D2H)
            graphalign.alignGraphsBatch(jobs, opts, mode, ctx)
            t0 = time.perf_counter()
            for _ in range(max(1, steps // 2)):
                graphalign.alignGraphsBatch(jobs, opts, mode, ctx)
            e2e_py_ms = (time.perf_counter() - t0) / max(1, steps // 2) * 1e3
            import ctypes as C
            from prographmsa_b200 import _lib as L_
            keep = []
            nj = len(jobs)
            G1 = (L_.Graph * nj)(); G2 = (L_.Graph * nj)(); MM = (L_.DPModel * nj)()
            for k, (g1, g2, mdl) in enumerate(jobs):
                G1[k] = graphalign.c_graph(g1, keep); G2[k] = graphalign.c_graph(g2, keep); MM[k] = graphalign.c_model(mdl, keep)
            caps = np.array([2 * (g1.N + g2.N) + 8 for g1, g2, _ in jobs], dtype=np.int64)
            offs = np.concatenate([[0], np.cumsum(caps)])
            maps = np.empty((2, int(offs[-1])), dtype=np.int32)
            res_arr = np.zeros(nj, dtype=graphalign._ALIGNMENT)
            res_arr["mapping1"] = maps[0].ctypes.data + 4 * offs[:-1]; res_arr["mapping2"] = maps[1].ctypes.data + 4 * offs[:-1]
            res_arr["capacity"] = caps
            o_ = graphalign.c_opts(opts, mode)
            def cabi():
                L_.check(L_.lib().pgm_align_graphs_batch(ctx.handle, jobs[0][0].n, nj, G1, G2, MM, C.byref(o_), C.c_void_p(res_arr.ctypes.data)))
            cabi()
            t0 = time.perf_counter()
            for _ in range(max(1, steps // 2)):
                cabi()
            e2e_ms = (time.perf_counter() - t0) / max(1, steps // 2) * 1e3
            assert abs(float(res_arr["score"].astype(np.float64).sum()) - float(sum(r["score"] for r in res))) <= 1e-3 * abs(float(sum(r["score"] for r in res)))
            fill_ms = tm["graph_fill_ms"] / steps
            # algorithmic HBM bytes of the fill kernels: 4 B S read + 1 B packed traceback per cell.
            # What the kernels move: 4 B S + 2 B direction word per cell, plus the few kept rows / columns of the value store
            alg_bytes = 5.0 * cells
            out[mode_name] = dict(
                value=upd / ms / 1e6, unit="GCUPS", ms_per_step=ms, jobs=len(jobs), cell_updates=upd, cells=cells,
                schedule="per size class: emission -> fill -> traceback on its own stream (pgm_graph_batch_run); ms_per_step_phased = "
                         "the three phases one after the other over the whole batch (PGM_GRAPH_PIPELINE=0), which is what kernels_ms splits",
                ms_per_step_phased=ms_phased,
                e2e=dict(value=upd / e2e_ms / 1e6, unit="GCUPS", ms_per_step=e2e_ms, through="pgm_align_graphs_batch (C ABI, host graphs in, mappings out)",
                         python_mirror_ms_per_step=e2e_py_ms),
                kernels_ms=dict(emission=tm["emission_ms"] / steps, fill=fill_ms, trace=tm["graph_trace_ms"] / steps),
                gpu_launches=tp["launches"],
                roofline=dict(bound="hbm", kernel="graph_fill_strip_kernel + graph_fill_strip_gen_kernel + graph_fill_wave_kernel (concurrent launches, one fill section)",
                              achieved=alg_bytes / fill_ms / 1e6, peak=pk["hbm_gbs"], unit="GB/s",
                              frac=alg_bytes / fill_ms / 1e6 / pk["hbm_gbs"], traffic=None, peak_source=pk["source"],
                              moved_bytes_per_cell=6.0,
                              note="algorithmic bytes = 5 B/cell (SURVEY 8d: 4 B S read + 1 B packed traceback); the strip kernels move 4.5 B/cell (4 direction "
                                   "bits + S, class-coded S for leaf graphs), the anti-diagonal (wave) kernel 20 B/cell on the few branchy graphs it sweeps. "
                                   "Not the binding roof: the fill is bound by dependent-issue latency (roofline_issue; profiles/r2_*)"),
                roofline_issue=dict(bound="sm fp32 issue", kernel="fill section (all fill kernels)", achieved=17.0 * upd / fill_ms / 1e9,
                                    peak=148 * 128 * pk.get("sm_max_mhz", 1965.0) * 1e6 / 1e12, unit="Tlane-op/s",
                                    frac=17.0 * upd / fill_ms / 1e9 / (148 * 128 * pk.get("sm_max_mhz", 1965.0) * 1e6 / 1e12),
                                    note="algorithmic work = 17 fp32 add/max per predecessor-pair cell update (SURVEY 8d, GraphAlign.h:245-256); "
                                         "peak = 148 SM x 128 lanes x sm_max_mhz"),
                score_checksum=float(sum(r["score"] for r in res)))
        finally:
            gb.close()
    # Second workload: the bottom level of BASELINE configs[4] -- 2500 leaf x leaf jobs of a 5000-sequence x 500-aa family (half of a
    # progressive pass's cells sit in its two bottom levels).  Enough independent sweeps to fill the GPU, so this is the fill kernels'
    # THROUGHPUT, where the mixed batch above measures the latency of its few long sweeps.
    try:
        from prographmsa_b200.graph import sequence_graph
        seqs5, _ = synth.simulate_family(LEAF_LEVEL_JOBS * 2, LEAF_LEVEL_LEN, mf, seed=SEED + 2)
        names5 = sorted(seqs5)
        mdl = mf.getModel(0.3, 0.3)
        ljobs = [(sequence_graph(seqs5[names5[2 * k]], 20), sequence_graph(seqs5[names5[2 * k + 1]], 20), mdl) for k in range(LEAF_LEVEL_JOBS)]
        gb = graphalign.GraphBatch(ljobs, opts, graphalign.EMISSION_EXACT, ctx)
        try:
            flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
            ctx.enable_timing(True)
            os.environ["PGM_GRAPH_PIPELINE"] = "0"
            for _ in range(warmup):
                gb.run(); ctx.sync()
            tl = dict(emission_ms=0.0, graph_fill_ms=0.0, graph_trace_ms=0.0)
            tot = 0.0
            for _ in range(steps):
                flush.fill_(1); torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); gb.run(); ctx.sync(); b.record(); torch.cuda.synchronize()
                tot += a.elapsed_time(b)
                t = ctx.timing()
                for k in tl: tl[k] += t[k]
                upd5, cells5 = t["graph_cell_updates"], t["graph_cells"]
            os.environ.pop("PGM_GRAPH_PIPELINE")
            fill5 = tl["graph_fill_ms"] / steps
            lanes_peak = 148 * 128 * pk.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
            out["leaf_level"] = dict(
                workload="%d leaf x leaf alignGraphs jobs, %d aa (bottom level of a %d-sequence progressive pass, BASELINE configs[4])" % (LEAF_LEVEL_JOBS, LEAF_LEVEL_LEN, 2 * LEAF_LEVEL_JOBS),
                value=upd5 / (tot / steps) / 1e6, unit="GCUPS", ms_per_step=tot / steps, jobs=LEAF_LEVEL_JOBS, cell_updates=upd5, cells=cells5,
                kernels_ms=dict(emission=tl["emission_ms"] / steps, fill=fill5, trace=tl["graph_trace_ms"] / steps),
                fill_gcups=upd5 / fill5 / 1e6,
                roofline=dict(bound="hbm", kernel="graph_fill_strip_kernel", achieved=5.0 * cells5 / fill5 / 1e6, peak=pk["hbm_gbs"], unit="GB/s",
                              frac=5.0 * cells5 / fill5 / 1e6 / pk["hbm_gbs"], traffic=379235584.0, peak_source=pk["source"],
                              note="algorithmic bytes = 5 B/cell (SURVEY 8d: 4 B S read + 1 B packed traceback).  The kernel itself moves ~0.6 B/cell: "
                                   "leaf graphs are class-coded (S is a 21 x 21 table per job) and the traceback is 4 bits per cell, so this is the "
                                   "section-8d figure, not DRAM traffic (traffic: 55 MB read + 324 MB written per launch, ncu capture profiles/r2b_strip_leaf_level.txt); "
                                   "the binding roof is instruction issue (roofline_issue: ncu shows 64 % of the issue slots busy)"),
                roofline_issue=dict(bound="sm fp32 issue", kernel="graph_fill_strip_kernel", achieved=17.0 * upd5 / fill5 / 1e9, peak=lanes_peak, unit="Tlane-op/s",
                                    frac=17.0 * upd5 / fill5 / 1e9 / lanes_peak,
                                    note="algorithmic work = 17 fp32 add/max per cell update; the kernel executes 1.39 warp instructions per cell "
                                         "(25 for the recurrences and the 4 direction bits, the rest per-step overhead), profiles/r2_strip_chain.txt"))
        finally:
            gb.close()
    except Exception as e:
        out["leaf_level"] = dict(unavailable=str(e))
    out["workload"] = ("%d alignGraphs jobs in one batch: the %d internal-node jobs (leaf x leaf up to profile x profile graphs) of a "
                       "progressive pass over %d synthetic seqs x %d aa (WAG), replicated x%d" % (len(jobs), base_jobs, GRAPH_N_SEQ, SEQ_LEN, GRAPH_REPLICAS))
    return out


def other_kernels_bench(ctx, torch, lib, prm, acc, steps, pk, seqs, names):
    """SURVEY.md section 8d's remaining rooflines (N = 1 only): the ML kernel against the FP64 FMA rate measured in this
    run, and the Prealigned front-end (MSA column scan) against HBM."""
    from prographmsa_b200 import _lib
    out = {}
    tf = C.c_double(0.0)
    _lib.check(lib.pgm_probe_fp64(ctx.handle, C.byref(tf)))
    n = 20
    evals = acc["ml_evaluations"] / steps
    ml_ms = acc["ml_ms"] / steps
    alg_flop = evals * (6.0 * n ** 3 + 10.0 * n ** 2)                   # dense-algorithm flops per likelihood evaluation (section 8d)
    out["ml_kernel"] = dict(bound="fp64", kernel="ml_kernel", evaluations_per_step=evals, launch_ms=ml_ms,
                            achieved=alg_flop / ml_ms / 1e9, peak=tf.value, unit="TFLOP/s (FP64)", frac=alg_flop / ml_ms / 1e9 / tf.value,
                            peak_source="pgm_probe_fp64: 8 independent DFMA chains per thread, measured in this run",
                            note="algorithmic flops = (6 n^3 + 10 n^2) per Newton evaluation; the kernel itself runs nnz * 3n FMA per evaluation "
                                 "(sparse counts) and is shared-memory-bound (DESIGN.md section 4)")
    # Prealigned: a pseudo-alignment of the bench family (gap columns inserted at random), all pairs, device-resident
    rng = np.random.default_rng(3)
    L_aln = int(max(len(seqs[k]) for k in names) * 1.6) & ~15
    from prographmsa_b200 import alphabet as _alpha
    rows = np.full((len(names), L_aln), _alpha.GAP, dtype=np.uint8)
    for r, k in enumerate(names):
        pos = np.sort(rng.choice(L_aln, size=len(seqs[k]), replace=False))
        rows[r, pos] = seqs[k]
    N = len(names); P = N * (N - 1) // 2
    d_rows = torch.from_numpy(rows).cuda()
    d_out = torch.zeros(2, P, dtype=torch.float64, device="cuda")
    vp = lambda t: C.c_void_p(t.data_ptr())
    def step():
        _lib.check(lib.pgm_prealigned_ml_pair_range_dev(ctx.handle, C.byref(prm), vp(d_rows), N, L_aln, C.c_int64(0), C.c_int64(P),
                                                        vp(d_out[0]), vp(d_out[1]), None, None))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    alg = 2.0 * L_aln * P                                                # bytes if both rows of every pair are read once
    out["prealigned"] = dict(bound="hbm", kernel="prealigned_lists_kernel + prealigned_count_kernel (list merge) + ml_kernel<FRONT_COUNTS16> (Newton)", workload="%d rows x %d columns, %d pairs" % (N, L_aln, P),
                             ms_per_matrix=ms, pairs_per_s=P / ms * 1e3, achieved=alg / ms / 1e6, peak=pk["hbm_gbs"], unit="GB/s",
                             frac=alg / ms / 1e6 / pk["hbm_gbs"],
                             note="algorithmic bytes = 2 L_aln per pair (section 8d, rows re-read per pair); the rows (%.1f MB) stay in L2, so this "
                                  "is a fraction of HBM peak reached out of L2, and the time includes the Newton iterations" % (rows.nbytes / 1e6))
    # the same path on rows shaped like BASELINE configs[4]'s alignment: 11,000 columns, 500 residues per row (the rest gaps)
    try:
        N5, L5, R5 = 2000, 11008, 500
        rows5 = np.full((N5, L5), _alpha.GAP, dtype=np.uint8)
        shared = np.sort(rng.choice(L5, R5, replace=False))
        for r in range(N5):
            cols = shared.copy()
            k = rng.integers(0, R5, 60); cols[k] = rng.integers(0, L5, 60)      # ~10 % of a row's residues sit in columns of its own
            cols = np.unique(cols)
            rows5[r, cols] = rng.integers(0, 20, len(cols))
        P5 = N5 * (N5 - 1) // 2
        d_rows5 = torch.from_numpy(rows5).cuda()
        d_out5 = torch.zeros(2, P5, dtype=torch.float64, device="cuda")
        def step5():
            _lib.check(lib.pgm_prealigned_ml_pair_range_dev(ctx.handle, C.byref(prm), vp(d_rows5), N5, L5, C.c_int64(0), C.c_int64(P5),
                                                            vp(d_out5[0]), vp(d_out5[1]), None, None))
        step5(); torch.cuda.synchronize()
        a.record()
        for _ in range(3):
            step5()
        b.record(); torch.cuda.synchronize()
        ms5 = a.elapsed_time(b) / 3
        alg5 = 2.0 * L5 * P5
        out["prealigned_config5_shape"] = dict(bound="hbm", kernel=out["prealigned"]["kernel"], workload="%d rows x %d columns (%d residues each), %d pairs" % (N5, L5, R5, P5),
                                               ms_per_matrix=ms5, pairs_per_s=P5 / ms5 * 1e3, achieved=alg5 / ms5 / 1e6, peak=pk["hbm_gbs"], unit="GB/s",
                                               frac=alg5 / ms5 / 1e6 / pk["hbm_gbs"],
                                               note="algorithmic bytes = 2 L_aln per pair (section 8d).  The list merge reads 4 B per residue instead "
                                                    "(~8 KB per pair at 500 residues per row, from L2), so this is the section-8d figure, not traffic")
    except Exception as e:
        out["prealigned_config5_shape"] = dict(unavailable=str(e))
    return out


def run_ours(args):
    import torch
    from prographmsa_b200 import _lib, alphabet, distance, sharding
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ["NCCL_DEBUG"] = "WARN"                  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pk = peaks()
    opts, mf, seqs, names, _ = make_family()
    AA = alphabet.AA
    code_list = [AA.nw_codes(seqs[k]) for k in names]
    codes, offsets = distance.pack_sequences(code_list)
    tab, go, ge = distance.nw_table(AA)
    keep = []
    prm = distance.make_ml_params(mf, opts, keep)
    P = N_SEQ * (N_SEQ - 1) // 2
    per = sharding.shard_size(N_SEQ, world)                # equal-sized shards (the last one is padded)
    pb, pe = sharding.pair_shard(N_SEQ, rank, world)
    cells_total = 0
    L = np.diff(offsets.astype(np.int64))
    for i in range(N_SEQ):
        cells_total += int(((L[i] + 1) * (L[i + 1:] + 1)).sum())

    ctx = _lib.Context(local)
    ctx.enable_timing(True)
    lib = _lib.lib()
    h_codes = torch.from_numpy(codes).pin_memory()
    d_codes = h_codes.cuda()
    d_out = torch.zeros(2, per, dtype=torch.float64, device="cuda")          # [dist; var] of my shard
    d_all = torch.zeros(world, 2, per, dtype=torch.float64, device="cuda") if world > 1 else None
    h_out = torch.zeros(world, 2, per, dtype=torch.float64).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def vp(t):
        return C.c_void_p(t.data_ptr())

    def step_resident():
        _lib.check(lib.pgm_nw_ml_pair_range_dev(ctx.handle, C.byref(prm), vp(d_codes), offsets.ctypes.data_as(C.c_void_p), N_SEQ,
                                                tab.ctypes.data_as(C.c_void_p), go, ge, C.c_int64(pb), C.c_int64(pe),
                                                vp(d_out[0]), vp(d_out[1]), None, None, None))
        if world > 1:
            dist.all_gather_into_tensor(d_all.view(-1), d_out.view(-1))

    def step_e2e():
        d_codes.copy_(h_codes, non_blocking=True)                            # H2D of this step's input
        torch.cuda.current_stream().synchronize()
        step_resident()
        src = d_all if world > 1 else d_out.unsqueeze(0)
        if rank == 0:
            h_out.copy_(src, non_blocking=True)                               # D2H of the step's result
        torch.cuda.current_stream().synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, acc=None):
        tot = 0.0
        for _ in range(steps):
            flush.fill_(1)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tot += float(t.item())
            if acc is not None:
                tm = ctx.timing()
                for k in acc:
                    acc[k] += tm[k]
        return tot

    for _ in range(max(3, args.warmup)):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    acc = dict(nw_forward_ms=0.0, nw_walk_ms=0.0, ml_ms=0.0, nw_forward_launches=0, nw_walk_launches=0, ml_launches=0,
               tb_bytes=0, ml_evaluations=0, nw_cells=0)
    total_ms = timed(step_resident, args.steps, acc)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)
    # C-ABI host-buffer call (what the reference-side shim calls: plain host pointers in, host pointers out), rank-local
    # range; one untimed call first (buffer growth), then the mean of `steps` calls with the wall clock
    dd = np.zeros(max(1, pe - pb)); vv = np.zeros(max(1, pe - pb))

    def cabi_call():
        _lib.check(lib.pgm_nw_ml_pair_range(ctx.handle, C.byref(prm), codes.ctypes.data_as(C.c_void_p), offsets.ctypes.data_as(C.c_void_p), N_SEQ,
                                            tab.ctypes.data_as(C.c_void_p), go, ge, C.c_int64(pb), C.c_int64(pe),
                                            dd.ctypes.data_as(C.c_void_p), vv.ctypes.data_as(C.c_void_p), None, None, None))
    cabi_call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cabi_call()
    cabi_host_ms = (time.perf_counter() - t0) * 1e3 / args.steps

    if rank == 0:
        ms = total_ms / args.steps
        e2e = e2e_ms / args.steps
        fwd_launches = max(1, acc["nw_forward_launches"])
        fwd_ms = acc["nw_forward_ms"] / fwd_launches                          # average launch duration (this rank)
        cells_rank = acc["nw_cells"] / args.steps
        tb_rank = acc["tb_bytes"] / args.steps
        launches_per_step = (acc["nw_forward_launches"] + acc["nw_walk_launches"] + acc["ml_launches"]) / args.steps
        # algorithmic HBM bytes of the forward kernel per launch: the packed traceback it must write
        # (2 bits/cell rounded up to the lane layout) + both sequences of every pair read once
        seq_bytes = float(sum(int(L[i]) * (N_SEQ - 1 - i) + int(L[i + 1:].sum()) for i in range(N_SEQ))) / world
        alg_bytes_launch = (tb_rank + seq_bytes) / (fwd_launches / args.steps)
        cells_launch = cells_rank / (fwd_launches / args.steps)
        lane_ops = 148 * 128 * pk["sm_max_mhz"] * 1e6                        # int32 lane-ops/s at max clock
        line = dict(
            metric=METRIC, value=cells_total / ms / 1e6, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
            ms_per_step=ms, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="int32+f64", data="synthetic",
            config=workload_config(seqs, names),
            sharding="pair range / rank, NCCL all-gather of (dist,var)" if world > 1 else "single GPU",
            traceback_gb_per_step=tb_rank * world / 1e9,
            seconds_per_matrix=ms / 1e3,
            e2e=dict(value=cells_total / e2e / 1e6, unit=UNIT, ms_per_step=e2e, h2d_bytes_per_step=int(codes.nbytes),
                     d2h_bytes_per_step=int(world * 2 * per * 8), cabi_host_call_ms_rank0=cabi_host_ms),
            gpu_launches=int(launches_per_step * args.steps),
            kernels_ms_per_step=dict(nw_forward=acc["nw_forward_ms"] / args.steps, nw_walk=acc["nw_walk_ms"] / args.steps, ml=acc["ml_ms"] / args.steps),
            ml_evaluations_per_step=acc["ml_evaluations"] / args.steps,
            clocks=clocks,
            roofline=dict(bound="hbm", kernel="nw_forward_kernel", achieved=alg_bytes_launch / fwd_ms / 1e6, peak=pk["hbm_gbs"], unit="GB/s",
                          frac=alg_bytes_launch / fwd_ms / 1e6 / pk["hbm_gbs"],
                          # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload at N=1 (6.33 MB + 6.971 GB),
                          # from the ncu --set full capture summarised in profiles/r1_final_hp2.txt; not re-measured per run
                          traffic=(6.978e9 if world == 1 else None), peak_source=pk["source"],
                          launch_ms=fwd_ms, algorithmic_bytes_per_launch=alg_bytes_launch),
            roofline_issue=dict(bound="sm integer issue", kernel="nw_forward_kernel", achieved=9.0 * cells_launch / fwd_ms / 1e9,
                                peak=lane_ops / 1e12, unit="Tlane-op/s", frac=9.0 * cells_launch / fwd_ms / 1e9 / (lane_ops / 1e12),
                                note="algorithmic work = 9 int32 ops/cell (SURVEY.md §8d); peak = 148 SM x 128 lanes x sm_max_mhz. The 16-bit twin "
                                     "flavour retires two cells per lane-op (9 instructions per cell PAIR), so this is a fraction of the 32-bit-op bound, "
                                     "not of issue slots: ncu shows 74 % of issue slots busy (profiles/r1_final_hp2.txt)",
                                gcups_kernel=cells_launch / fwd_ms / 1e6),
        )
        if world == 1:
            line["cpu_baseline"] = cpu_baseline(opts, mf, seqs, names)
            try:
                line["other_kernels"] = other_kernels_bench(ctx, torch, lib, prm, acc, args.steps, pk, seqs, names)
            except Exception as e:
                line["other_kernels"] = dict(error=repr(e))
            try:
                line["graph_dp"] = graph_dp_bench(ctx, torch, max(3, args.steps // 2), 3, pk)
            except Exception as e:                                            # never lose the headline line
                line["graph_dp"] = dict(error=repr(e))
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line, on the real stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries (NCCL's version banner, torch.distributed warnings) write to file descriptor 1; the driver wants exactly one
    # JSON line there.  Everything else is sent to stderr for the lifetime of the process.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
