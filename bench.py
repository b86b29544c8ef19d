#!/usr/bin/env python3
"""bench.py -- throughput of the barcode-clustering hot path on N B200s (driver contract).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the CPU arm (oracle port on the host cores)

A step = one pass of the whole hot path (bucketing -> edit-distance pairs at k = 2*maxDiff ->
genotype-aware merge + consensus) over one synthetic library.  `value` is reads/s with the
inputs resident in HBM; `e2e` is the same call with pinned HOST buffers (H2D of the reads,
D2H of the cluster table arrays inside the timed region).  Between timed steps L2 is flushed
by writing a 512 MiB buffer.  Timing: CUDA events on the library's stream, max over ranks.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
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
    ap.add_argument("--cpu-sample-reads", type=int, default=100_000)
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


def cpu_baseline(args, params, steps=1, warmup=0):
    from oracle import orc
    orc.set_threads(os.cpu_count() or 1)      # torchrun pins OMP_NUM_THREADS=1; the CPU arm uses every host core
    base = synth.CONFIGS[args.config]["gen"]
    scale = min(1.0, args.cpu_sample_reads / base["n_reads"])
    lib = synth.make_config(args.config, seed=args.seed, scale=scale)
    for _ in range(warmup):
        cpu_pass(lib, params)
    times = [cpu_pass(lib, params) for _ in range(max(1, steps))]
    sec = statistics.mean(t["seconds"] for t in times)
    return dict(value=lib.n_reads / sec, unit="reads/s", cores=orc.num_threads(), kind="port",
                sample=f"{args.config} scaled to {lib.n_reads} reads / {lib.params['n_clones']} clones "
                       f"(oracle/pcb_oracle.c: bucketing + seeded banded-DP pairs (OpenMP) + literal merge restatement; "
                       f"the reference's own bowtie2 + R distance stage cannot run here)",
                seconds_per_pass=sec, stages={k: statistics.mean(t[k] for t in times) for k in ("bucket_s", "pairs_s", "merge_s")},
                host_cores=os.cpu_count()), lib.n_reads, sec


# ----------------------------------------------------------------------------- GPU arm
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = synth.CONFIGS[args.config]
    params = dict(max_diff=cfg["max_diff"], min_qual=cfg["min_qual"], min_jaccard=0.2, min_matches=1)
    workload = f"{args.config}: {cfg['gen']['n_clones']} clones / {cfg['gen']['n_reads']} reads"
    if args.scale != 1.0:
        workload += f" x{args.scale}"

    if args.impl == "reference":
        if rank != 0:
            return
        base, n_reads, sec = cpu_baseline(args, params, steps=args.steps, warmup=min(args.warmup, 1))
        line = dict(metric=METRIC, value=base["value"], unit="reads/s", n_gpus=args.gpus, steps=args.steps,
                    warmup=args.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="int32", data="synthetic", impl="reference",
                    config=dict(workload=workload, sample=base["sample"], max_diff=params["max_diff"]),
                    cpu_baseline=base,
                    e2e=dict(value=base["value"], unit="reads/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    import torch
    from pacybara_b200 import multigpu, pipeline
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
    if world > 1:
        workload += f"; x{world} for {world} GPUs = {lib.params['n_clones']} clones / {lib.n_reads} reads, one library"
    ra, names = pipeline.arrays_from_library(lib, ctx, with_names=False)
    R = ra.n_reads
    dev = f"cuda:{local_rank}"
    host = {k: v for k, v in vars(ra).items() if v is not None}
    pinned = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in pinned.items()}
    # the quality matrix is not an input of the clustering call (only of pcb_bucket's per-bucket sums)
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

    STAGES = ("t_bucket_ns", "t_index_ns", "probe_ns", "t_hits_ns", "t_prep_ns", "t_replay_ns", "t_scatter_ns")
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
            probe_ns.append(ctx.stat("probe_ns"))
            for k in STAGES:
                stage_acc.setdefault(k, []).append(ctx.stat(k) * 1e-6)
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
    stage_ms = {(k[2:] if k.startswith("t_") else k)[:-3]: statistics.mean(v) for k, v in stage_acc.items()}
    # e2e: same call, pinned host inputs, host outputs
    timed(pinned, False, 1)
    e2e_ms, _, out_h = timed(pinned, False, args.steps)
    d2h_bytes = sum(np.asarray(v).nbytes for k, v in out_h.items() if hasattr(v, "nbytes"))
    if dist is not None:
        # whole job: every rank uploads its chunk of the reads (+ the small mutation table) and ships its own rows
        t = torch.tensor([d2h_bytes], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        d2h_bytes = int(t.item())
        h2d_bytes += (world - 1) * pinned["mut_rank"].numel() * pinned["mut_rank"].element_size()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / args.steps
    value = R / (ms_per_step * 1e-3)
    # ---- roofline of the dominant kernel (the Myers probe kernel)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    probe_s = statistics.mean(probe_ns) * 1e-9 if probe_ns else float("nan")
    pairs, cells, U, E = stats["pairs_scored"], stats["cells"], stats["buckets"], stats["edges"]
    bc_len = lib.params["bc_len"] * (2 if lib.params.get("combo") else 1)
    words = 1 if bc_len + 6 <= 32 else 2
    # algorithmic bytes per launch: 16 B planes + 4 B length per query probe and per scored
    # candidate, 4 B per walked bin entry (>= scored), 12 B per hit
    kk = (2 if args.full_table else 1) * params["max_diff"]
    per_query = (kk + 1) * (2 * kk + 1)
    algo_bytes = U * per_query * 20 + pairs * 24 + E * 12
    roof_probe = dict(bound="hbm", achieved=algo_bytes / probe_s / 1e9, peak=hbm_peak, unit="GB/s",
                      frac=algo_bytes / probe_s / 1e9 / hbm_peak, traffic=None, kernel="probe_kernel<false, true>",
                      note="integer-ALU bound kernel: the HBM fraction is small by construction, see roofline_alu")
    # the kernel that takes the largest share of the step: the per-component replay (merge.cu).  Algorithmic bytes
    # (DESIGN.md 4.4): per read 16 B of inputs (lexrank, barcode id, bucket, genotype offset) + 28 B of cluster state
    # written and read back + 12 B of per-read outputs; 8 B per genotype entry; 32 B per output row; 4 B per call.
    replay_s = stage_ms.get("replay", 0.0) * 1e-3
    if replay_s <= 0:                       # sharded runs call the stages one by one: no fused stage clock
        replay_s = float("nan")
    n_rows = int((np.asarray(out["row_size"].cpu() if hasattr(out["row_size"], "cpu") else out["row_size"]) > 0).sum())
    nnz = int(ra.geno_mut.shape[0])
    replay_bytes = R * (16 + 2 * 28 + 12) + nnz * 8 + n_rows * 32 + nnz * 4
    roof = dict(bound="hbm", achieved=replay_bytes / replay_s / 1e9, peak=hbm_peak, unit="GB/s",
                frac=replay_bytes / replay_s / 1e9 / hbm_peak, traffic=91.5e6 if args.config == "cfg2" and world == 1 else None,
                kernel="component_kernel_thread" if R / max(U, 1) < 7 else "component_kernel_warp", ms=replay_s * 1e3,
                share_of_step=replay_s * 1e3 / ms_per_step,
                peak_source="measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback",
                note="latency-bound sequential replay of the reference's greedy merge, one thread (or warp) per connected "
                     "component: DRAM traffic ~ algorithmic bytes (ncu r01e: 72 MB read + 19 MB written), the time goes to "
                     "dependent loads (long_scoreboard / wait stalls), not to bandwidth; traffic = ncu dram bytes at cfg 2")
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

    line = dict(metric=METRIC, value=value, unit="reads/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int32",
                data="synthetic",
                config=dict(workload=workload, n_reads=R, n_buckets=U, n_distance_pairs=E, barcode_nt=bc_len,
                            max_diff=params["max_diff"], pair_search_k=(2 if args.full_table else 1) * params["max_diff"],
                            distance_table="full (<= 2*maxDiff)" if args.full_table else "<= maxDiff + on-demand in merge",
                            seed_len=stats["seed_len"],
                            l2="flushed between steps (512 MiB write)", parallelism=(f"x{world}: pairs sharded by query bucket, edge all-gather, merge and results sharded by component"
                                         if world > 1 else "1 GPU")),
                e2e=dict(value=R / (e2e_ms / args.steps * 1e-3), unit="reads/s", h2d_bytes_per_step=h2d_bytes,
                         d2h_bytes_per_step=d2h_bytes, ms_per_step=e2e_ms / args.steps),
                gpu_launches=launches, clocks=clocks, roofline=roof, roofline_probe=roof_probe, roofline_alu=alu,
                edit_distance_gcups=(cells / probe_s / 1e9) if probe_s == probe_s else None,
                kernel_ms=dict(probe=probe_s * 1e3), stage_ms=stage_ms)
    if world == 1 and not args.no_cpu_baseline:
        base, _, _ = cpu_baseline(args, params)
        line["cpu_baseline"] = base
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
