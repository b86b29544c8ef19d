"""The hot path over N GPUs of one box: one process per GPU, torch.distributed (NCCL) for the plumbing.

What shards and what is exchanged:
  1. bucketing            replicated -- every rank buckets the whole read set (HBM-bound, a few ms;
                          it yields the identical bucket numbering everywhere without an exchange);
  2. distance pairs       SHARDED by query bucket: rank r scores the pairs whose query side lies in
                          its slice of [0, U) (pcb_edit_pairs q_begin/q_end) -- the dominant kernel;
  3. edge exchange        the one real collective: all-gather of the per-rank edge lists (counts, then
                          the padded (i, j, d) payload) over NVLink, then pcb_sort_edges restores the
                          canonical row order;
  4. merge                SHARDED by connected component (round-robin, pcb_merge shard_rank/count):
                          components are independent units of the reference's sequential algorithm;
  5. table exchange       element-wise MAX all-reduce of the per-read outputs (entries a rank does not
                          own are PCB_FILL = INT32_MIN).
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from .api import Context
from .hostdata import MERGE_OUTPUTS, MergeProblem, PairList


def shard_range(n: int, rank: int, world: int):
    return (n * rank) // world, (n * (rank + 1)) // world


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


def cluster_reads_sharded(ctx: Context, arrs: Dict, params: Dict, rank: int, world: int, on_device: bool = True,
                          full_table: bool = False) -> Dict:
    """`arrs`: dict of torch tensors (seq, qual, length, lexrank, geno_ptr, geno_mut, geno_q, mut_rank[, up_id]),
    CUDA tensors when on_device else pinned host tensors (copied in here; results copied back)."""
    import torch
    import torch.distributed as dist
    dev = f"cuda:{ctx.device}"
    a = arrs if on_device else {k: v.to(dev, non_blocking=True) for k, v in arrs.items() if k != "qual"}
    max_diff = int(params["max_diff"])
    R = int(a["length"].shape[0])
    # 1. buckets (replicated)
    b = ctx.bucket(a["seq"], None, a["length"], want_qsum=False)      # qualities are not needed for clustering
    U = b["n_buckets"]
    # 2. my slice of the query buckets
    lo, hi = shard_range(U, rank, world)
    bseq, blen = b["bucket_seq"].contiguous(), b["bucket_len"].contiguous()
    ei, ej, ed = ctx.edit_pairs(bseq, blen, (2 if full_table else 1) * max_diff, q_range=(lo, hi))
    local_stats = dict(pairs_scored=ctx.stat("pairs_scored"), cells=ctx.stat("cells"), probe_ns=ctx.stat("probe_ns"))
    # 3. exchange edges
    ei, ej, ed = gather_edges(ei, ej, ed, world)
    ctx.sort_edges(ei, ej, ed, U)
    ped = torch.where((ed >= 1) & (ed <= max_diff), ed, torch.zeros_like(ed))
    # 4. my components
    prob = MergeProblem(n_reads=R, n_buckets=U, bucket_ptr=b["bucket_ptr"].contiguous(), bucket_reads=b["bucket_reads"].contiguous(),
                        lexrank=a["lexrank"], bc_id=b["bucket_of_read"].contiguous(), up_id=a.get("up_id"), geno_ptr=a["geno_ptr"],
                        geno_mut=a["geno_mut"], geno_q=a["geno_q"], mut_rank=a["mut_rank"], max_diff=max_diff,
                        min_qual=int(params["min_qual"]), min_jaccard=float(params["min_jaccard"]),
                        min_matches=int(params["min_matches"]))
    if full_table:
        out = ctx.merge(prob, PairList(ei, ej, ed, ped), shard=(rank, world))
    else:
        out = ctx.merge(prob, PairList(ei, ej, ed, ped), shard=(rank, world), bucket_seq=bseq, bucket_len=blen,
                        on_demand_k=2 * max_diff)
    # 5. combine
    for name in list(out.keys()):
        dist.all_reduce(out[name], op=dist.ReduceOp.MAX)
    torch.cuda.current_stream(ctx.device).synchronize()
    out["bucket_of_read"] = b["bucket_of_read"]
    out["n_buckets"] = U
    out["n_edges"] = int(ei.numel())
    out["local_stats"] = local_stats
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
