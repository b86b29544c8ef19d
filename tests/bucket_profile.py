"""Bucketing alone at several sizes, device-resident inputs (run on the GPU box): wall clock and the library's stage clock."""
import json
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pacybara_b200 import synth
from pacybara_b200.api import Context

ctx = Context(0)
rng = np.random.default_rng(1)
for R in [int(x) for x in (sys.argv[1:] or ["5000000", "10000000", "16000000", "17000000", "20000000"])]:
    n_clones = R // 10
    codes = rng.integers(0, 4, size=(n_clones, 25), dtype=np.int8)
    clone = rng.integers(0, n_clones, size=R)
    seq = np.zeros((R, 31), dtype=np.uint8)
    seq[:, :25] = synth.BASES[codes[clone]]
    err = rng.random(R) < 0.1
    seq[err, rng.integers(0, 25)] = ord("A")
    length = np.full(R, 25, dtype=np.int32)
    d_seq, d_len = torch.from_numpy(seq).cuda(), torch.from_numpy(length).cuda()
    out = []
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        b = ctx.bucket(d_seq, None, d_len, want_qsum=False)
        torch.cuda.synchronize()
        out.append((round((time.perf_counter() - t0) * 1e3, 2), round(ctx.stat("t_bucket_ns") * 1e-6, 2)))
    print(json.dumps(dict(R=R, U=b["n_buckets"], wall_ms_and_stage_ms=out, torch_reserved_gb=round(torch.cuda.memory_reserved() / 2**30, 2))), flush=True)
