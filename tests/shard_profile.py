#!/usr/bin/env python3
"""What one rank of an N-GPU run spends where, measured on ONE GPU (no NCCL): python tests/shard_profile.py N [reads_per_gpu_scale]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pacybara_b200 import multigpu, pipeline, synth  # noqa: E402
from pacybara_b200.api import Context  # noqa: E402
from pacybara_b200.hostdata import MergeProblem, PairList  # noqa: E402


def main():
    N = int(sys.argv[1])
    lib = synth.make_config("cfg2", seed=1, scale=float(N))
    ctx = Context(0)
    ra, _ = pipeline.arrays_from_library(lib, ctx, with_names=False)
    a = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in vars(ra).items() if v is not None}
    R = ra.n_reads

    def timed(f):
        torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); return r, (time.perf_counter() - t) * 1e3

    for rep in range(3):
        b, t_bucket = timed(lambda: ctx.bucket(a["seq"], None, a["length"], want_qsum=False))
        U = b["n_buckets"]
        bseq, blen = b["bucket_seq"].contiguous(), b["bucket_len"].contiguous()
        lo, hi = multigpu.shard_range(U, 0, N)
        (ei, ej, ed), t_pairs_shard = timed(lambda: ctx.edit_pairs(bseq, blen, 2, t_range=(lo, hi)))
        st_pairs = {k: round(ctx.stat(k) * 1e-6, 3) for k in ("t_index_ns", "probe_ns", "t_hits_ns")}
        (fi, fj, fd), t_pairs_all = timed(lambda: ctx.edit_pairs(bseq, blen, 2))
        fi, fj, fd = fi.clone(), fj.clone(), fd.clone()
        _, t_sort = timed(lambda: ctx.sort_edges(fi, fj, fd, U))
        ped = torch.where((fd >= 1) & (fd <= 2), fd, torch.zeros_like(fd))
        prob = MergeProblem(n_reads=R, n_buckets=U, bucket_ptr=b["bucket_ptr"].contiguous(), bucket_reads=b["bucket_reads"].contiguous(),
                            lexrank=a["lexrank"], bc_id=b["bucket_of_read"].contiguous(), up_id=None, geno_ptr=a["geno_ptr"],
                            geno_mut=a["geno_mut"], geno_q=a["geno_q"], mut_rank=a["mut_rank"], max_diff=2, min_qual=32)
        out, t_merge = timed(lambda: ctx.merge(prob, PairList(fi, fj, fd, ped), shard=(0, N), bucket_seq=bseq, bucket_len=blen, on_demand_k=4))
        st_merge = {k: round(ctx.stat(k) * 1e-6, 3) for k in ("t_prep_ns", "t_replay_ns", "t_scatter_ns")}
        _, t_compact = timed(lambda: ctx.compact_shard(dict(out, bucket_of_read=b["bucket_of_read"]), to_host=True))
        nbytes = sum(v.numel() * v.element_size() for v in out.values())
    print(f"N={N} reads={R} U={U} edges={fi.numel()}: bucket {t_bucket:.2f} ms, pairs(shard 1/{N}) {t_pairs_shard:.2f} ms "
          f"(all: {t_pairs_all:.2f}), sort_edges {t_sort:.2f}, merge(shard) {t_merge:.2f} ms, compact+D2H {t_compact:.2f} ms, all-reduce payload {nbytes / 1e6:.0f} MB; "
          f"pair stages {st_pairs}, merge stages {st_merge}")


if __name__ == "__main__":
    main()
