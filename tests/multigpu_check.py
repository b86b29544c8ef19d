#!/usr/bin/env python3
"""Run under torchrun on N GPUs: the sharded hot path (pacybara_b200.multigpu) must give the very
same arrays as the single-GPU call, and the oracle's table.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multigpu_check.py [cfg] [scale]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pacybara_b200 import canon, multigpu, pipeline, synth  # noqa: E402
from pacybara_b200.api import Context  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ctx = Context(local)
    lib = synth.make_config(cfg, seed=2, scale=scale)
    ra, names = pipeline.arrays_from_library(lib, ctx)
    params = dict(max_diff=lib.params["max_diff"], min_qual=lib.params["min_qual"], min_jaccard=0.2, min_matches=1)
    dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in vars(ra).items() if v is not None}
    full = os.environ.get("PCB_FULL_TABLE", "0") == "1"
    out = multigpu.cluster_reads_sharded(ctx, dev, params, rank, world, on_device=True, full_table=full, combine=True)
    single = ctx.cluster_reads(dev["seq"], dev["qual"], dev["length"], dev["lexrank"], dev.get("up_id"), dev["geno_ptr"],
                               dev["geno_mut"], dev["geno_q"], dev["mut_rank"], full_table=full, **params)
    bad = [k for k in ("read_root", "read_pos", "row_size", "row_key", "row_bc", "row_up", "row_call_n", "bucket_of_read")
           if not torch.equal(out[k], single[k])]
    assert not bad, f"rank {rank}: sharded != single-GPU in {bad}"
    # where a row's calls sit in call_mut depends on the run (regions are handed out as components come by): compare rows
    from pacybara_b200 import hostdata as _hd
    ra_, rb_ = (_hd.rows_from_merge_outputs({k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in o.items()}) for o in (out, single))
    for f in ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut"):
        assert np.array_equal(getattr(ra_, f), getattr(rb_, f)), f"rank {rank}: sharded != single-GPU in {f}"
    assert out["n_edges"] == single["n_edges"]
    host = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}
    digest = canon.table_digest(canon.canon_table(pipeline.table_from_outputs(host, ra, names)))
    if rank == 0:
        from tests import util
        want = canon.canon_table(util.oracle_table_for_library(lib, ra, names, **params))
        assert digest == canon.table_digest(want), "sharded table differs from the oracle"
        print(f"multigpu_check ok: {world} ranks, {lib.n_reads} reads, {out['n_buckets']} buckets, {out['n_edges']} edges, "
              f"rank-0 share of scored pairs {out['local_stats']['pairs_scored']}")
    # results left sharded on the devices: the element-wise max over the ranks is the single-GPU result
    names = ("read_root", "read_pos", "row_size", "row_key", "row_bc", "row_up", "row_call_off", "row_call_n", "call_mut")
    part = multigpu.cluster_reads_sharded(ctx, dev, params, rank, world, on_device=True, full_table=full, combine=False)
    owned = int((part["read_root"] != -(1 << 31)).sum())
    for k in names:
        if k in ("row_call_off", "call_mut"):
            continue                       # run-dependent layout without PCB_MERGE_SHARD_LAYOUT; the compacted shards below carry the calls
        t = part[k].clone()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert torch.equal(t, single[k]), f"rank {rank}: max over shards != single-GPU in {k}"
    # host buffers: every rank uploads its chunk of the reads and gets its own reads and rows back, compacted
    from pacybara_b200 import hostdata
    pinned = {k: v.cpu().pin_memory() for k, v in dev.items()}
    shard = multigpu.cluster_reads_sharded(ctx, pinned, params, rank, world, on_device=False, full_table=full, combine=False)
    assert shard["own_read"].shape[0] == owned
    mine = {k: (np.array(v) if hasattr(v, "shape") else v) for k, v in shard.items()}
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    whole = hostdata.merge_outputs_from_shards(parts, lib.n_reads)
    ref = {k: single[k].cpu().numpy() for k in names + ("bucket_of_read",)}
    assert np.array_equal(whole["bucket_of_read"], ref["bucket_of_read"])
    a, b = hostdata.rows_from_merge_outputs(ref), hostdata.rows_from_merge_outputs(whole)
    for f in ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f"rank {rank}: host shards != single-GPU in {f}"
    if rank == 0:
        print(f"multigpu_check sharded results ok: rank 0 owns {owned} of {lib.n_reads} reads, {shard['row_rep'].shape[0]} rows")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
