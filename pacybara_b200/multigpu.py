"""The hot path over N GPUs of one box: one process per GPU, torch.distributed (NCCL) for the plumbing.

What shards and what is exchanged:
  1. bucketing            replicated -- every rank buckets the whole read set (HBM-bound, a few ms;
                          it yields the identical bucket numbering everywhere without an exchange);
  2. distance pairs       SHARDED over a (query groups x target groups) grid of the ranks: a rank indexes the targets
                          of its slice of [0, U) and scores the pairs whose query side lies in its query slice
                          (pcb_edit_pairs q_begin/q_end, t_begin/t_end; pair_grid) -- the dominant kernel, and
                          the index build divides with the target groups;
  3. edge exchange        the one real collective: all-gather of the per-rank edge lists (counts, then
                          the padded (i, j, d) payload) over NVLink, then pcb_sort_edges restores the
                          canonical row order;
  4. merge                SHARDED by connected component (round-robin, pcb_merge shard_rank/count):
                          components are independent units of the reference's sequential algorithm;
  5. result               stays SHARDED by default (combine=False): every rank holds the rows of its own
                          components (entries of reads it does not own are PCB_FILL = INT32_MIN) and, with host
                          buffers, ships only those (pcb_compact_shard).  combine=True replicates the whole result
                          on every rank with an element-wise MAX all-reduce of the per-read outputs.
With host buffers (on_device=False) every rank uploads 1/N of the reads and the ranks all-gather them over NVLink
(upload_sharded), instead of N full copies crossing PCIe.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from .api import Context
from .hostdata import MERGE_OUTPUTS, MergeProblem, PairList


def shard_range(n: int, rank: int, world: int):
    return (n * rank) // world, (n * (rank + 1)) // world


def pair_grid(world: int):
    """Ranks as a (query groups x target groups) grid for the pair search.  A rank indexes the targets of its target
    group and probes with the queries of its query group, so the index build costs U / Gt and the per-probe work
    (decode, query planes, bin bounds) U / Gq, while the bin entries walked are 1 / world either way.  Measured at
    8 * 10^6 reads (profiles/r02j_shard_profile_n8.txt): index 1.4 ms and probe overhead 0.93 ms for all of U, so one
    split of the queries pays from 4 ranks on: 1 x 2, 2 x 2, 2 x 4."""
    gq = 2 if world >= 4 and world % 2 == 0 else 1
    return gq, world // gq


def pair_ranges(U: int, rank: int, world: int):
    gq, gt = pair_grid(world)
    return shard_range(U, rank % gq, gq), shard_range(U, rank // gq, gt)


def gather_edges(ei, ej, ed, world: int):
    """All-gather variable-length edge lists.  Returns concatenated device tensors (rank order)."""
    import torch
    import torch.distributed as dist
    dev = ei.device
    n = torch.tensor([ei.numel()], dtype=torch.int64, device=dev)
    counts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, n)
    counts_h = counts.tolist()
    m = max(1, max(counts_h))
    mine = torch.zeros(3, m, dtype=torch.int32, device=dev)
    k = ei.numel()
    mine[0, :k], mine[1, :k], mine[2, :k] = ei, ej, ed
    allbuf = torch.empty(world * 3, m, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(allbuf, mine)
    allbuf = allbuf.view(world, 3, m)
    parts = [allbuf[r, :, : counts_h[r]] for r in range(world)]
    cat = torch.cat(parts, dim=1).contiguous()
    return cat[0].contiguous(), cat[1].contiguous(), cat[2].contiguous()


def read_chunk(n: int, rank: int, world: int):
    """Equal chunks (the last ones may be short or empty): rank r uploads reads [lo, hi)."""
    chunk = (n + world - 1) // world
    return chunk, min(n, rank * chunk), min(n, (rank + 1) * chunk)


def upload_sharded(arrs: Dict, rank: int, world: int, dev) -> Dict:
    """Host arrays (the full read set, identical on every rank) -> full device arrays on every rank, with every rank
    copying only its chunk of the reads over PCIe and the chunks all-gathered between the GPUs.  The genotype CSR
    travels as it is: geno_ptr holds global offsets, so the concatenation of the chunks IS the pointer array."""
    import torch
    import torch.distributed as dist
    R = int(arrs["length"].shape[0])
    chunk, lo, hi = read_chunk(R, rank, world)
    out = {}

    def gather_rows(src, extra_tail=0):
        shape = (chunk,) + tuple(src.shape[1:])
        mine = torch.zeros(shape, dtype=src.dtype, device=dev)
        if hi > lo:
            mine[: hi - lo].copy_(src[lo:hi], non_blocking=True)
        full = torch.empty((world * chunk + extra_tail,) + tuple(src.shape[1:]), dtype=src.dtype, device=dev)
        dist.all_gather_into_tensor(full[: world * chunk], mine)
        return full

    for k in ("seq", "length", "lexrank", "up_id"):
        if arrs.get(k) is not None:
            out[k] = gather_rows(arrs[k])[:R]
    gptr = arrs["geno_ptr"]
    nnz = int(gptr[R])
    full = gather_rows(gptr[:R], extra_tail=1)
    full[R] = nnz
    out["geno_ptr"] = full[: R + 1]
    # genotype entries of my reads: [ptr[lo], ptr[hi]); the counts of all ranks follow from the (replicated) host pointer
    bounds = [int(gptr[min(R, r * chunk)]) for r in range(world + 1)]
    m = max(1, max(bounds[r + 1] - bounds[r] for r in range(world)))
    for k in ("geno_mut", "geno_q"):
        mine = torch.zeros(m, dtype=arrs[k].dtype, device=dev)
        n = bounds[rank + 1] - bounds[rank]
        if n:
            mine[:n].copy_(arrs[k][bounds[rank]: bounds[rank + 1]], non_blocking=True)
        allbuf = torch.empty(world * m, dtype=arrs[k].dtype, device=dev)
        dist.all_gather_into_tensor(allbuf, mine)
        parts = [allbuf[r * m: r * m + bounds[r + 1] - bounds[r]] for r in range(world)]
        out[k] = torch.cat(parts) if world > 1 else parts[0]
        assert out[k].numel() == nnz
    out["mut_rank"] = arrs["mut_rank"].to(dev, non_blocking=True)
    return out


class ShardError(RuntimeError):
    """A local stage failed on some rank; every rank raises this at the same point instead of hanging in the next collective."""


def _all_ranks_ok(stage: str, local_error, dev) -> None:
    """Collective: ranks agree on whether `stage` went through everywhere (an error raised on one rank only -- a missing
    uptag in a component it owns, a full hit buffer, no memory -- would leave the others blocked in the next all-gather)."""
    import torch
    import torch.distributed as dist
    flag = torch.tensor([0 if local_error is None else 1], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if int(flag.item()):
        if local_error is not None:
            raise local_error
        raise ShardError(f"another rank failed in stage '{stage}'")


def cluster_reads_sharded(ctx: Context, arrs: Dict, params: Dict, rank: int, world: int, on_device: bool = True,
                          full_table: bool = False, combine: bool = False) -> Dict:
    """`arrs`: dict of torch tensors (seq, qual, length, lexrank, geno_ptr, geno_mut, geno_q, mut_rank[, up_id]),
    CUDA tensors when on_device else pinned host tensors holding the full read set on every rank.
    Result: on_device -> per-read device arrays (combine=False: entries of reads other ranks own are PCB_FILL);
    host buffers -> combine=False: this rank's reads and rows, compacted (Context.compact_shard; put the ranks' parts
    together with hostdata.merge_outputs_from_shards); combine=True: the full per-read arrays on every rank."""
    import torch
    import torch.distributed as dist
    dev = f"cuda:{ctx.device}"
    stage_ms: Dict[str, float] = {}

    def lib_stages(*names):
        for n in names:                                   # device time of the library's own stages (CUDA events on its stream)
            stage_ms[n] = stage_ms.get(n, 0.0) + ctx.stat(("t_" if n != "probe" else "") + n + "_ns") * 1e-6

    def torch_stage(name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = fn()
        e1.record()
        e1.synchronize()
        stage_ms[name] = stage_ms.get(name, 0.0) + e0.elapsed_time(e1)
        return res

    a = arrs if on_device else torch_stage("upload_allgather", lambda: upload_sharded(arrs, rank, world, dev))
    max_diff = int(params["max_diff"])
    R = int(a["length"].shape[0])
    # 1 + 2. buckets (replicated) and my block of the (query, target) grid of the pair search -- one library call; the
    # bucket arrays and the packed barcodes stay in the context for the merge
    gq, gt = pair_grid(world)
    err, bor, U, n_mine = None, None, 0, 0
    try:
        bor, U, n_mine = ctx.shard_pairs(a["seq"], a["length"], (2 if full_table else 1) * max_diff, rank % gq, gq, rank // gq, gt)
    except Exception as e:          # noqa: BLE001 - re-raised on every rank below
        err = e
    lib_stages("bucket", "index", "probe", "hits")
    local_stats = dict(pairs_scored=ctx.stat("pairs_scored"), cells=ctx.stat("cells"), probe_ns=ctx.stat("probe_ns"))

    # 3. exchange edges: the counts (a negative count says "my search failed": every rank then raises at the same point
    # instead of hanging in the next collective), then the padded (i, j, d) rows
    def exchange():
        n = torch.tensor([n_mine if err is None else -1], dtype=torch.int64, device=dev)
        counts = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, n)
        counts_h = counts.tolist()
        if min(counts_h) < 0:
            if err is not None:
                raise err
            raise ShardError("another rank failed in stage 'pair search'")
        m = max(1, max(counts_h))
        mine = torch.empty(3, m, dtype=torch.int32, device=dev)
        if n_mine:
            ctx.shard_edges_into(mine)          # (every library call returns with its stream drained)
        allbuf = torch.empty(world, 3, m, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(allbuf.view(world * 3, m), mine)
        return allbuf, counts_h
    allbuf, counts_h = torch_stage("edge_allgather", exchange)
    # 4. my components (the edge sort is part of the library call and of its `prep` time)
    err = None
    try:
        out, n_all = ctx.shard_merge(allbuf, counts_h, bor, a["lexrank"], a.get("up_id"), a["geno_ptr"], a["geno_mut"], a["geno_q"],
                                     a["mut_rank"], max_diff, int(params["min_qual"]), float(params["min_jaccard"]),
                                     int(params["min_matches"]), on_demand=not full_table, shard=(rank, world), shard_layout=combine)
    except Exception as e:          # noqa: BLE001 - e.g. PCB_E_UPTAG, raised only where the component lives
        err = e
    _all_ranks_ok("merge", err, dev)
    lib_stages("prep", "replay", "scatter")
    local_stats.update(links=ctx.stat("links"), tests=ctx.stat("tests"), components=ctx.stat("components"))
    b = dict(bucket_of_read=bor, n_buckets=U)
    # 5. result
    if combine:
        for name in list(out.keys()):
            dist.all_reduce(out[name], op=dist.ReduceOp.MAX)
        torch.cuda.current_stream(ctx.device).synchronize()
    extra = dict(bucket_of_read=b["bucket_of_read"], n_buckets=U, n_edges=n_all, local_stats=local_stats, stage_ms=stage_ms)
    if not on_device and not combine:
        # bucket ids travel for my reads only (row_bc names buckets; the barcode of a bucket is that of any of its reads)
        shard = ctx.compact_shard(dict(out, bucket_of_read=b["bucket_of_read"]), to_host=True)
        shard.update(n_buckets=U, n_edges=extra["n_edges"], local_stats=local_stats, stage_ms=stage_ms)
        return shard
    out.update(extra)
    if not on_device:
        # results to page-locked host buffers (reused by the next call, like Context(pinned_outputs=True))
        host = {}
        for k, v in out.items():
            if hasattr(v, "is_cuda") and v.is_cuda:
                key = ("mg_" + k, tuple(v.shape), str(v.dtype))
                buf = ctx._pinned.get(key)
                if buf is None:
                    buf = ctx._pinned[key] = torch.empty(v.shape, dtype=v.dtype).pin_memory()
                buf.copy_(v, non_blocking=True)
                host[k] = buf
            else:
                host[k] = v
        torch.cuda.current_stream(ctx.device).synchronize()
        out = {k: (v.numpy() if hasattr(v, "numpy") and not isinstance(v, (int, dict)) else v) for k, v in host.items()}
    return out


def cluster_reads_virtual(ctx: Context, arrs: Dict, params: Dict, world: int, full_table: bool = False) -> Dict:
    """The sharded algorithm with `world` VIRTUAL ranks on ONE GPU, one after the other: what every rank computes in
    cluster_reads_sharded (its slice of the pair search, its components of the merge, its compacted shard), with the
    collectives replaced by concatenation.  No NCCL involved: this exists so that the decomposition itself -- target
    slices that partition the pairs, components dealt to shards, shards reassembled -- is checked wherever one GPU is
    available.  `arrs`: CUDA tensors as for cluster_reads_sharded(on_device=True).  Returns per-read host arrays."""
    import torch
    from . import hostdata
    max_diff = int(params["max_diff"])
    R = int(arrs["length"].shape[0])
    b = ctx.bucket(arrs["seq"], None, arrs["length"], want_qsum=False)
    U = b["n_buckets"]
    bseq, blen = b["bucket_seq"].contiguous(), b["bucket_len"].contiguous()
    parts = []
    for r in range(world):
        q_range, t_range = pair_ranges(U, r, world)
        parts.append(ctx.edit_pairs(bseq, blen, (2 if full_table else 1) * max_diff, q_range=q_range, t_range=t_range))
    ei, ej, ed = (torch.cat([p[t] for p in parts]).contiguous() for t in range(3))        # the all-gather
    ctx.sort_edges(ei, ej, ed, U)
    ped = torch.where((ed >= 1) & (ed <= max_diff), ed, torch.zeros_like(ed))
    prob = MergeProblem(n_reads=R, n_buckets=U, bucket_ptr=b["bucket_ptr"].contiguous(), bucket_reads=b["bucket_reads"].contiguous(),
                        lexrank=arrs["lexrank"], bc_id=b["bucket_of_read"].contiguous(), up_id=arrs.get("up_id"), geno_ptr=arrs["geno_ptr"],
                        geno_mut=arrs["geno_mut"], geno_q=arrs["geno_q"], mut_rank=arrs["mut_rank"], max_diff=max_diff,
                        min_qual=int(params["min_qual"]), min_jaccard=float(params["min_jaccard"]), min_matches=int(params["min_matches"]))
    shards = []
    for r in range(world):
        kw = {} if full_table else dict(bucket_seq=bseq, bucket_len=blen, on_demand_k=2 * max_diff)
        out = ctx.merge(prob, PairList(ei, ej, ed, ped), shard=(r, world), pairs_sorted=True, shard_layout=False, **kw)
        shard = ctx.compact_shard(dict(out, bucket_of_read=b["bucket_of_read"]), to_host=True)
        shards.append({k: (np.array(v) if hasattr(v, "shape") else v) for k, v in shard.items()})
    whole = hostdata.merge_outputs_from_shards(shards, R)
    whole.update(n_buckets=U, n_edges=int(ei.numel()), shard_rows=[int(s["row_rep"].shape[0]) for s in shards])
    return whole
