#!/usr/bin/env python3
"""bench.py -- throughput of the barcode-clustering hot path on N B200s (driver contract).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the CPU arm (oracle port on the host cores)

A step = one pass of the whole hot path (bucketing -> edit-distance pairs -> genotype-aware merge +
consensus) over one synthetic library.  `value` is reads/s with the inputs resident in HBM; `e2e` is the
same call with pinned HOST buffers (H2D of the reads, D2H of the cluster table arrays inside the timed
region); `files_e2e` is wall clock from the gzip files on disk to clusters.csv.gz written (SURVEY 8(d)
metric 1; N = 1, rank 0).  Between timed steps L2 is flushed by writing a 512 MiB buffer.  Timing: CUDA
events on the library's stream, max over ranks.  Outside the timed region every run checks its result
(`parity`): at N = 1 against the CPU oracle on a 10^5-read library of the same configuration, at N > 1 the
reassembled shards against the single-GPU result of the same library -- a mismatch exits non-zero.
`--impl reference` times the oracle port on the host cores at the FULL bench configuration (as many passes
as fit the time budget, at least one; `steps` reports the passes actually timed).
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from pacybara_b200 import synth  # noqa: E402

METRIC = "barcode reads clustered/sec"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=os.environ.get("PCB_BENCH_CONFIG", "cfg2"))
    ap.add_argument("--scale", type=float, default=float(os.environ.get("PCB_BENCH_SCALE", "1.0")))
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-sample-reads", type=int, default=600_000,
                    help="reads of the bounded sample the cpu_baseline leg of the default run clusters on the host cores")
    ap.add_argument("--cpu-budget-s", type=float, default=float(os.environ.get("PCB_CPU_BUDGET_S", "240")),
                    help="--impl reference: wall-clock budget for the timed full-size passes (at least one pass is always timed)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity check outside the timed region")
    ap.add_argument("--no-files-e2e", action="store_true", help="skip the gz files -> clusters.csv.gz measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--full-table", action="store_true",
                    help="materialise every pair up to 2*maxDiff (the calcEdits table) instead of on-demand distances")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, cmax = float(f[1]), float(f[2])
            except ValueError:
                continue
            if t0 - 0.05 <= t <= t1 + 0.05:
                sm.append(clk)
                mx.append(cmax)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            allclk = [float(l.split(",")[1]) for _, l in self.rows if len(l.split(",")) >= 9 and l.split(",")[1].strip().replace(".", "").isdigit()]
            return dict(sm_mhz=statistics.median(allclk) if allclk else None, sm_max_mhz=None, reasons=sorted(reasons),
                        samples=0)
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------- CPU arm
def cpu_pass(lib, params):
    """The oracle port on the host cores: the whole path once.  Returns seconds and counts."""
    from oracle import orc
    from pacybara_b200 import hostdata, pipeline
    t0 = time.perf_counter()
    r = orc.precluster(lib.seq, lib.qual, lib.length)
    U = r["n_buckets"]
    order = np.argsort(r["bucket_of_read"], kind="stable").astype(np.int32)
    ptr = np.zeros(U + 1, dtype=np.int32)
    np.cumsum(np.bincount(r["bucket_of_read"], minlength=U), out=ptr[1:])
    bseq, blen = lib.seq[r["first_read"]], lib.length[r["first_read"]]
    t1 = time.perf_counter()
    ei, ej, ed = orc.edit_pairs(bseq, blen, 2 * params["max_diff"], method="seeded")
    t2 = time.perf_counter()
    q, rr, d = hostdata.canonical_rows(ei, ej, ed)
    prob = hostdata.MergeProblem(n_reads=lib.n_reads, n_buckets=U, bucket_ptr=ptr, bucket_reads=order,
                                 lexrank=np.arange(lib.n_reads, dtype=np.int32), bc_id=r["bucket_of_read"], up_id=None,
                                 geno_ptr=lib.geno_ptr.astype(np.int32), geno_mut=lib.geno_mut, geno_q=lib.geno_q,
                                 mut_rank=np.arange(len(lib.mut_codes), dtype=np.int32), **params)
    rows = orc.cluster(prob, q, rr, d)
    t3 = time.perf_counter()
    return dict(seconds=t3 - t0, bucket_s=t1 - t0, pairs_s=t2 - t1, merge_s=t3 - t2, n_buckets=U, n_edges=len(ei),
                n_rows=rows.n_rows)


def bench_config(args, cfg, world=1):
    """The workload, named identically by both arms (`config` of the JSON line)."""
    g = cfg["gen"]
    return dict(workload=f"{args.config}: {g['n_clones']} clones / {g['n_reads']} reads" + (f" x{args.scale}" if args.scale != 1.0 else ""),
                n_reads=int(g["n_reads"] * args.scale), n_clones=int(g["n_clones"] * args.scale),
                barcode_nt=g["bc_len"] * (2 if g.get("combo") else 1), max_diff=cfg["max_diff"], min_qual=cfg["min_qual"],
                min_jaccard=0.2, min_matches=1, seed=args.seed, output="cluster table (clusters.csv.gz rows) incl. consensus",
                l2="GPU arm: flushed between steps (512 MiB write); CPU arm: not applicable")


def cpu_baseline(args, params, sample_reads=None, steps=1, warmup=0, budget_s=None):
    """Oracle port on every host core.  sample_reads=None: the full configuration."""
    from oracle import orc
    orc.set_threads(os.cpu_count() or 1)      # torchrun pins OMP_NUM_THREADS=1; the CPU arm uses every host core
    base = synth.CONFIGS[args.config]["gen"]
    scale = args.scale if sample_reads is None else min(args.scale, sample_reads / base["n_reads"])
    lib = synth.make_config(args.config, seed=args.seed, scale=scale)
    for _ in range(warmup):
        cpu_pass(lib, params)
    times, t_start = [], time.perf_counter()
    for _ in range(max(1, steps)):
        times.append(cpu_pass(lib, params))
        if budget_s is not None and time.perf_counter() - t_start + times[-1]["seconds"] > budget_s:
            break                                # the next pass would overrun the budget
    sec = statistics.mean(t["seconds"] for t in times)
    whole = scale == args.scale
    return dict(value=lib.n_reads / sec, unit="reads/s", cores=orc.num_threads(), kind="port",
                sample=(f"the whole workload ({lib.n_reads} reads)" if whole else
                        f"{args.config} scaled to {lib.n_reads} reads / {lib.params['n_clones']} clones") +
                       f", {len(times)} pass(es); oracle/pcb_oracle.c: bucketing incl. Phred sums + seeded banded-DP pairs up to 2*maxDiff "
                       f"(OpenMP) + literal merge restatement (1 thread, like the reference); the reference's own bowtie2 + R "
                       f"distance stage cannot run here",
                seconds_per_pass=sec, passes=len(times),
                stages={k: statistics.mean(t[k] for t in times) for k in ("bucket_s", "pairs_s", "merge_s")},
                host_cores=os.cpu_count()), lib.n_reads, sec, len(times)


def rows_digest(rows) -> str:
    """sha256 over the cluster rows as arrays: row order, member order, consensus barcode ids, calls."""
    h = hashlib.sha256()
    for f in ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut"):
        h.update(np.ascontiguousarray(getattr(rows, f), dtype=np.int64).tobytes())
    return h.hexdigest()


def finite(x):
    """JSON has no NaN/Infinity: non-finite floats become null, recursively."""
    if isinstance(x, float):
        return x if math.isfinite(x) else None
    if isinstance(x, dict):
        return {k: finite(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [finite(v) for v in x]
    if isinstance(x, np.generic):
        return finite(x.item())
    return x


def emit(line):
    print(json.dumps(finite(line), allow_nan=False))


def replay_traffic(config: str, world: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the replay launches of one step, from the committed ncu capture
    (profiles/replay_traffic.json, written by profiles/summarize.py from an `ncu --set full` run of this very command)."""
    try:
        with open(os.path.join(ROOT, "profiles", "replay_traffic.json")) as fh:
            t = json.load(fh)
        if t.get("config") == config and int(t.get("n_gpus", 1)) == world:
            return float(t["dram_bytes_per_step"]), t.get("source")
    except Exception:
        pass
    return None, None


# ----------------------------------------------------------------------------- GPU arm
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = synth.CONFIGS[args.config]
    params = dict(max_diff=cfg["max_diff"], min_qual=cfg["min_qual"], min_jaccard=0.2, min_matches=1)
    config = bench_config(args, cfg)

    if args.impl == "reference":
        if rank != 0:
            return
        # the full bench configuration (N > 1: the per-GPU share of the N-times library, i.e. the same 1x workload --
        # the host cannot cluster an 8x library within minutes; said in `sample`)
        base, n_reads, sec, passes = cpu_baseline(args, params, sample_reads=None, steps=args.steps, warmup=0,
                                                  budget_s=args.cpu_budget_s)
        if world > 1 or args.gpus > 1:
            base["sample"] += f"; --gpus {args.gpus}: the 1x configuration, not the {args.gpus}x library the GPU arm shards"
        line = dict(metric=METRIC, value=base["value"], unit="reads/s", n_gpus=args.gpus, steps=passes, warmup=0,
                    steps_requested=args.steps, warmup_requested=args.warmup,
                    ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="int32", data="synthetic", impl="reference", config=config, cpu_baseline=base,
                    e2e=dict(value=base["value"], unit="reads/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        emit(line)
        return

    import torch
    from pacybara_b200 import hostdata, multigpu, pipeline
    from pacybara_b200.api import Context

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    ctx = Context(local_rank, pinned_outputs=True)
    # N > 1: ONE library N times the size (reads per GPU fixed -> weak scaling), sharded as multigpu.py describes
    scale = args.scale * world
    lib = synth.make_config(args.config, seed=args.seed, scale=scale)
    ra, names = pipeline.arrays_from_library(lib, ctx, with_names=False)
    R = ra.n_reads
    dev = f"cuda:{local_rank}"
    host = {k: v for k, v in vars(ra).items() if v is not None}
    pinned = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in pinned.items()}
    # the quality matrix is not an input of the clustering call (only of pcb_bucket's per-bucket Phred sums, which the
    # cluster table does not contain)
    h2d_bytes = sum(v.numel() * v.element_size() for k, v in pinned.items() if k != "qual")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    ext = torch.cuda.ExternalStream(ctx.stream_handle(), device=dev)

    def run(arrs, on_device):
        if world > 1:
            # results stay sharded by component (north star: only the merge-edge lists are exchanged)
            return multigpu.cluster_reads_sharded(ctx, arrs, params, rank, world, on_device=on_device,
                                                  full_table=args.full_table, combine=False)
        a = arrs if on_device else {k: v.numpy() for k, v in arrs.items()}
        # host buffers: ask for one entry per ROW of the cluster table, so only the rows cross the bus
        return ctx.cluster_reads(a["seq"], a["qual"], a["length"], a["lexrank"], a.get("up_id"), a["geno_ptr"],
                                 a["geno_mut"], a["geno_q"], a["mut_rank"], full_table=args.full_table,
                                 compact_rows=not on_device, **params)

    LIB_STAGES = (("bucket", "t_bucket_ns"), ("index", "t_index_ns"), ("probe", "probe_ns"), ("hits", "t_hits_ns"),
                  ("prep", "t_prep_ns"), ("replay", "t_replay_ns"), ("scatter", "t_scatter_ns"))
    stage_acc = {}

    def timed(arrs, on_device, steps):
        total_ms, probe_ns, out = 0.0, [], None
        stage_acc.clear()
        for _ in range(steps):
            flush.fill_(1)
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            out = run(arrs, on_device)
            e1.record(ext)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if dist is not None:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            total_ms += ms
            if world > 1:
                probe_ns.append(out["local_stats"]["probe_ns"])
                for k, v in out["stage_ms"].items():
                    stage_acc.setdefault(k, []).append(v)
            else:
                probe_ns.append(ctx.stat("probe_ns"))
                for k, st in LIB_STAGES:
                    stage_acc.setdefault(k, []).append(ctx.stat(st) * 1e-6)
        return total_ms, probe_ns, out

    timed(resident, True, args.warmup)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = ctx.stat("kernel_launches")
    t_wall0 = time.time()
    total_ms, probe_ns, out = timed(resident, True, args.steps)
    t_wall1 = time.time()
    launches = ctx.stat("kernel_launches") - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    stats = ctx.stats()
    if world > 1:
        stats.update(out["local_stats"])
        stats["buckets"], stats["edges"] = out["n_buckets"], out["n_edges"]
    stage_ms = {k: statistics.mean(v) for k, v in stage_acc.items()}
    # e2e: same call, pinned host inputs, host outputs
    timed(pinned, False, 1)
    e2e_ms, _, out_h = timed(pinned, False, args.steps)
    d2h_bytes = sum(np.asarray(v).nbytes for k, v in out_h.items() if hasattr(v, "nbytes"))
    n_rows_local = int((np.asarray(out_h["row_size"]) > 0).sum())
    stage_limit = None
    if dist is not None:
        # whole job: every rank uploads its chunk of the reads (+ the small mutation table) and ships its own rows
        t = torch.tensor([d2h_bytes, n_rows_local, stats["pairs_scored"], stats["cells"], launches], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        d2h_bytes, n_rows, pairs_all, cells_all, launches = (int(x) for x in t.tolist())
        # how evenly the (query x target) blocks of the pair search and the component shards of the merge load the ranks
        mine = torch.tensor([stats["pairs_scored"], n_rows_local, int(out["local_stats"].get("components", 0))], dtype=torch.int64, device=dev)
        every = torch.empty(world * 3, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(every, mine)
        every = every.view(world, 3).tolist()
        shard_balance = dict(pairs_scored=[e[0] for e in every], rows=[e[1] for e in every], components=[e[2] for e in every],
                             pair_grid="%d query x %d target groups" % multigpu.pair_grid(world),
                             note="per rank; pairs: blocks of the pair grid, rows / components: shards by component label mod N")
        h2d_bytes += (world - 1) * pinned["mut_rank"].numel() * pinned["mut_rank"].element_size()
        # per stage: the slowest rank (what the step waits for), and which stage limits the step
        keys = sorted(stage_ms)
        t = torch.tensor([stage_ms[k] for k in keys], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        stage_ms = dict(zip(keys, t.tolist()))
    else:
        n_rows, pairs_all, cells_all = n_rows_local, stats["pairs_scored"], stats["cells"]

    # ---- parity, outside the timed region
    parity = None
    if not args.no_parity:
        if world == 1:
            from tests import util
            small = synth.make_config(args.config, seed=args.seed + 100, scale=args.scale * min(1.0, 100_000 / cfg["gen"]["n_reads"]))
            sra, snames = pipeline.arrays_from_library(small, ctx, with_names=False)
            snames["ids"] = np.arange(small.n_reads).astype(str).tolist()           # padded ids sort like their index
            sdev = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in vars(sra).items() if v is not None}
            sout = run(sdev, True)
            sout = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in sout.items()}
            got = hostdata.rows_from_merge_outputs(sout)
            from oracle import orc
            orc.set_threads(os.cpu_count() or 1)
            b = util.oracle_buckets(hostdata.FastqReads(snames["ids"], small.seq, small.qual, small.length))
            ei, ej, ed = orc.edit_pairs(b["bucket_seq"], b["bucket_len"], 2 * params["max_diff"], method="seeded")
            q, r, d = hostdata.canonical_rows(ei, ej, ed)
            prob = hostdata.MergeProblem(n_reads=small.n_reads, n_buckets=b["n_buckets"], bucket_ptr=b["bucket_ptr"],
                                         bucket_reads=b["bucket_reads"], lexrank=sra.lexrank, bc_id=b["bucket_of_read"], up_id=sra.up_id,
                                         geno_ptr=sra.geno_ptr, geno_mut=sra.geno_mut, geno_q=sra.geno_q, mut_rank=sra.mut_rank, **params)
            want = orc.cluster(prob, q, r, d)
            parity = dict(check="GPU rows vs CPU oracle (own bucketing, pair search, merge)", reads=small.n_reads, rows=want.n_rows,
                          digest=rows_digest(got), expected=rows_digest(want))
        else:
            mine = {k: (np.array(v) if hasattr(v, "shape") else v) for k, v in out_h.items() if k not in ("stage_ms", "local_stats")}
            parts = [None] * world
            dist.all_gather_object(parts, mine)
            if rank == 0:
                whole = hostdata.merge_outputs_from_shards(parts, R)
                got = hostdata.rows_from_merge_outputs(whole)
                a = resident
                single = ctx.cluster_reads(a["seq"], a["qual"], a["length"], a["lexrank"], a.get("up_id"), a["geno_ptr"], a["geno_mut"],
                                           a["geno_q"], a["mut_rank"], full_table=args.full_table, **params)
                single = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in single.items()}
                want = hostdata.rows_from_merge_outputs(single)
                same_buckets = bool(np.array_equal(whole["bucket_of_read"], single["bucket_of_read"]))
                parity = dict(check=f"rows reassembled from {world} shards vs the single-GPU run of the same library", reads=R,
                              rows=want.n_rows, digest=rows_digest(got), expected=rows_digest(want), same_buckets=same_buckets)
        if parity is not None:
            parity["digest_of"] = "sha256(row_ptr, row_members, row_bc, row_up, call_ptr, call_mut)"
            parity["equal"] = parity["digest"] == parity["expected"] and parity.get("same_buckets", True)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / args.steps
    value = R / (ms_per_step * 1e-3)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_source = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    probe_s = statistics.mean(probe_ns) * 1e-9 if probe_ns and statistics.mean(probe_ns) > 0 else float("nan")
    pairs, cells, U, E = stats["pairs_scored"], stats["cells"], stats["buckets"], stats["edges"]
    bc_len = config["barcode_nt"]
    words = 1 if bc_len + 6 <= 32 else 2
    # algorithmic bytes per launch: 16 B planes + 4 B length per query probe and per scored
    # candidate, 4 B per walked bin entry (>= scored), 12 B per hit; per rank (rank 0's slice at N > 1)
    kk = (2 if args.full_table else 1) * params["max_diff"]
    per_query = (kk + 1) * (2 * kk + 1)
    algo_bytes = (U // world) * per_query * 20 + pairs * 24 + (E // world) * 12
    roof_probe = dict(bound="hbm", achieved=algo_bytes / probe_s / 1e9, peak=hbm_peak, unit="GB/s",
                      frac=algo_bytes / probe_s / 1e9 / hbm_peak, traffic=None, kernel="probe_kernel_lanes (probe.cuh)",
                      note="integer-ALU bound kernel: the HBM fraction is small by construction, see roofline_alu")
    # the replay (merge.cu).  Algorithmic bytes per launch: per read 16 B of inputs (lexrank, barcode id,
    # bucket, genotype offset) + 12 B of per-read outputs; 8 B per genotype entry; 32 B per output row; 4 B per call.
    # N > 1: rank 0's components, i.e. 1/N of the library.
    replay_s = stage_ms.get("replay", 0.0) * 1e-3
    if replay_s <= 0:
        replay_s = float("nan")
    nnz = int(ra.geno_mut.shape[0])
    replay_bytes = (R * (16 + 12) + nnz * 8 + n_rows * 32 + nnz * 4) / world
    traffic, traffic_src = replay_traffic(args.config, world)
    roof = dict(bound="hbm", achieved=replay_bytes / replay_s / 1e9, peak=hbm_peak, unit="GB/s",
                frac=replay_bytes / replay_s / 1e9 / hbm_peak, traffic=traffic, traffic_source=traffic_src,
                algorithmic_bytes=replay_bytes, kernel="component replay (merge.cu)", ms=replay_s * 1e3,
                share_of_step=replay_s * 1e3 / ms_per_step, peak_source=peak_source,
                note="replay of the reference's greedy merge per connected component; achieved = algorithmic bytes of the "
                     "components this rank owns / stage time from CUDA events on the library's stream")
    alu = None
    try:
        al = C.CDLL(os.path.join(ROOT, "bench_aux", "_alu_peak.so"))
        vals = {}
        for mode, name in ((0, "lop3"), (1, "lop3+imad")):
            v, ms = C.c_double(0), C.c_double(0)
            if al.alu_peak_measure(local_rank, mode, C.byref(v), C.byref(ms)) == 0:
                vals[name] = v.value
        # textbook Myers/Hyyro op count: 21 ops per text column per machine word
        ops = 21.0 * cells / bc_len * words if bc_len else 0.0
        alu = dict(bound="int-alu", achieved=ops / probe_s / 1e12, peak=vals.get("lop3", float("nan")) / 1e12,
                   unit="Tlane-op/s", frac=ops / probe_s / vals["lop3"] if "lop3" in vals else None,
                   peak_dual_issue=vals.get("lop3+imad", float("nan")) / 1e12,
                   pairs_per_s=pairs / probe_s, gcups=cells / probe_s / 1e9,
                   note="achieved = 21 textbook ops x text columns x words per scored pair / kernel time; "
                        "peak = measured LOP3 issue rate (bench_aux/alu_peak.cu) at the clocks of this run")
    except Exception as e:  # pragma: no cover
        alu = dict(error=str(e))

    if world > 1:
        top = max(stage_ms, key=stage_ms.get)
        stage_limit = dict(stage=top, ms=stage_ms[top], note="slowest stage (max over ranks) of the resident step")
    line = dict(metric=METRIC, value=value, unit="reads/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int32",
                data="synthetic", config=config,
                run=dict(n_reads=R, n_buckets=U, n_distance_pairs=E, n_rows=n_rows, pair_search_k=kk,
                         distance_table="full (<= 2*maxDiff)" if args.full_table else "<= maxDiff + on-demand in merge",
                         seed_len=stats["seed_len"], phred_sums="not computed (the cluster table does not contain them; pcb_bucket does)",
                         parallelism=(f"x{world}: one library {world}x the configuration; pairs sharded by query bucket, edge "
                                      f"all-gather, merge and results sharded by component" if world > 1 else "1 GPU")),
                e2e=dict(value=R / (e2e_ms / args.steps * 1e-3), unit="reads/s", h2d_bytes_per_step=h2d_bytes,
                         d2h_bytes_per_step=d2h_bytes, ms_per_step=e2e_ms / args.steps),
                gpu_launches=launches, clocks=clocks, roofline=roof, roofline_probe=roof_probe, roofline_alu=alu,
                edit_distance_gcups=(cells_all / probe_s / 1e9), kernel_ms=dict(probe=probe_s * 1e3), stage_ms=stage_ms,
                parity=parity)
    if stage_limit:
        line["stage_limit"] = stage_limit
    if world > 1:
        line["shard_balance"] = shard_balance
    if world == 1 and not args.no_files_e2e:
        try:
            line["files_e2e"] = files_e2e(args, ctx, lib, params)
        except Exception as e:  # pragma: no cover - never lose the line to the side measurement
            line["files_e2e"] = dict(error=f"{type(e).__name__}: {e}")
    if world == 1 and not args.no_cpu_baseline:
        base, _, _, _ = cpu_baseline(args, params, sample_reads=args.cpu_sample_reads)
        line["cpu_baseline"] = base
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["equal"]:
        print("bench.py: PARITY MISMATCH " + json.dumps(parity), file=sys.stderr)
        sys.exit(3)


def files_e2e(args, ctx, lib, params):
    """SURVEY 8(d) metric (1): reads / wall time from the gzip input files on disk to clusters.csv.gz written, through
    the one-call program's own steps (pacybara_b200.fused_cli: gunzip -> device tokenizers -> pcb_cluster_reads -> table
    -> gzip).  One warm-up pass, one timed pass.  A tied >= 3-read barcode vote would shell out to `muscle`
    (pacybara_cluster.py:484-511), which is not on the GPU box: the first member's barcode is taken here."""
    import shutil
    import tempfile
    from pacybara_b200 import hostdata, ingest, pipeline
    d = tempfile.mkdtemp(prefix="pcb_bench_files_")
    try:
        paths = lib.write_files(d)
        same = paths["uptag"] == paths["bc"]
        laps = {}
        for rep in range(2):
            t0 = time.perf_counter()
            raw = ingest.read_all([paths["bc"], paths["geno"], None if same else paths["uptag"]])
            if same:
                raw[2] = raw[0]
            t1 = time.perf_counter()
            ra, names = ingest.arrays_from_bytes(ctx, raw[0], raw[1], raw[2])
            t2 = time.perf_counter()
            out = pipeline.cluster_reads(ctx, ra, full_table=args.full_table, **params)
            t3 = time.perf_counter()
            text = pipeline.table_text_from_outputs(ctx, out, ra, names, consensus=lambda bcs: bcs[0])      # rows formatted on the device (pcb_concat_tokens)
            table = None if text is not None else pipeline.table_from_outputs(out, ra, names, consensus=lambda bcs: bcs[0])
            t4 = time.perf_counter()
            if text is not None:
                hostdata.write_clusters_text(os.path.join(d, "clusters.csv.gz"), text)
            else:
                hostdata.write_clusters_csv(os.path.join(d, "clusters.csv.gz"), table)
            t5 = time.perf_counter()
            laps = dict(gunzip=t1 - t0, tokenize=t2 - t1, cluster=t3 - t2, table=t4 - t3, write=t5 - t4)
            n_table_rows = text.count(b"\n") - 1 if text is not None else len(table)
        sec = sum(laps.values())
        return dict(value=ra.n_reads / sec, unit="reads/s", seconds=sec, stages_s={k: round(v, 4) for k, v in laps.items()},
                    rows=n_table_rows, table="device (pcb_concat_tokens)" if text is not None else "host loop", input_bytes_gz=sum(os.path.getsize(p) for p in set(paths.values())),
                    note="wall clock, files on disk (page cache) -> clusters.csv.gz written; N = 1")
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
