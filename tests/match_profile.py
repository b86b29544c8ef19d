"""f3 measurement: queries/s of pcb_match_barcodes (device-resident arrays) and the exhaustive oracle on a sample.

    python tests/match_profile.py [library_rows] [queries] [max_err]   -> one JSON line
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np   # noqa: E402
import torch         # noqa: E402

from pacybara_b200.api import Context   # noqa: E402


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    Q = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    n = 25
    rng = np.random.default_rng(7)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    lib = acgt[rng.integers(0, 4, (L, n))]
    # queries: a library barcode with 0..3 substitutions (85 %), or random (15 %)
    src = rng.integers(0, L, Q)
    qs = lib[src].copy()
    n_sub = rng.integers(0, 4, Q)
    for e in range(3):
        rows = np.nonzero(n_sub > e)[0]
        qs[rows, rng.integers(0, n, rows.shape[0])] = acgt[rng.integers(0, 4, rows.shape[0])]
    rnd = rng.random(Q) < 0.15
    qs[rnd] = acgt[rng.integers(0, 4, (int(rnd.sum()), n))]
    ll, ql = np.full(L, n, np.int32), np.full(Q, n, np.int32)
    ctx = Context(0)
    d = [torch.from_numpy(a).cuda() for a in (lib, ll, qs, ql)]
    for _ in range(3):
        out = ctx.match_barcodes(*d, k)
    ext = torch.cuda.ExternalStream(ctx.stream_handle(), device="cuda:0")
    ms = []
    for _ in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        out = ctx.match_barcodes(*d, k)
        e1.record(ext)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    st = ctx.stats()
    t = sorted(ms)[len(ms) // 2]
    best = out["best_d"].cpu().numpy()
    line = dict(workload=f"{L} library barcodes x {Q} distinct queries, {n} nt, maxErr {k}", ms=t, queries_per_s=Q / (t * 1e-3),
                matched=int((best >= 0).sum()), ambiguous=int((out["n_best"].cpu().numpy() > 1).sum()), pairs_within_k=out["n_pairs"],
                pairs_scored=st["pairs_scored"], probe_ms=st["probe_ns"] * 1e-6, gcups=st["cells"] / (st["probe_ns"] * 1e-9) / 1e9,
                seed_len=st["seed_len"])
    from oracle import orc
    orc.lib().orc_set_threads(os.cpu_count())
    m = 2000
    t0 = time.perf_counter()
    ref = orc.match_barcodes(lib, ll, qs[:m], ql[:m], k)
    dt = time.perf_counter() - t0
    assert np.array_equal(ref["best_d"], best[:m]) and np.array_equal(ref["n_best"], out["n_best"].cpu().numpy()[:m])
    line["cpu_oracle"] = dict(queries_per_s=m / dt, cores=int(orc.lib().orc_num_threads()), sample=f"first {m} queries, exhaustive DP against all {L} rows")
    print(json.dumps(line))


if __name__ == "__main__":
    main()
