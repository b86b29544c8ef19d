#!/usr/bin/env python3
"""Large-configuration checks on a B200 (run by hand under gpurun, not by pytest: minutes of CPU time).

    python tests/scale_check.py <cfg> <scale> [oracle]

Runs the fused path on the named BASELINE.json config, checks size-independent invariants (every read in
exactly one row; planted clones come back; collisions stay apart; on-demand == full table when asked) and,
with `oracle`, the exact table against the CPU oracle.  Prints one JSON line with counts and stage times.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from pacybara_b200 import canon, hostdata, pipeline, synth  # noqa: E402
from pacybara_b200.api import Context  # noqa: E402


def main():
    cfg, scale = sys.argv[1], float(sys.argv[2])
    want_oracle = "oracle" in sys.argv[3:]
    want_full = "full" in sys.argv[3:]
    t0 = time.time()
    lib = synth.make_config(cfg, seed=1, scale=scale)
    t_gen = time.time() - t0
    ctx = Context(0)
    p = dict(max_diff=lib.params["max_diff"], min_qual=lib.params["min_qual"], min_jaccard=0.2, min_matches=1)
    # array inputs without building 20M id strings: numeric id order == index order (padded ids)
    up_id = None
    if lib.uptag_len is not None:
        b = ctx.bucket(lib.uptag_seq(), None, lib.uptag_len, want_qsum=False)
        up_id = b["bucket_of_read"].astype(np.int32)
    R = lib.n_reads
    args = (lib.seq, None, lib.length.astype(np.int32), np.arange(R, dtype=np.int32), up_id, lib.geno_ptr.astype(np.int32),
            lib.geno_mut, lib.geno_q, np.arange(len(lib.mut_codes), dtype=np.int32))
    ctx.cluster_reads(*args, **p)                                  # warm-up (pool growth)
    t0 = time.time()
    out = ctx.cluster_reads(*args, **p)
    t_run = time.time() - t0
    st = ctx.stats()
    rows = hostdata.rows_from_merge_outputs(out)                  # raises unless the rows partition the reads
    assert np.array_equal(np.sort(rows.row_members), np.arange(R))
    sizes = np.diff(rows.row_ptr)
    row_of_read = np.empty(R, dtype=np.int64)
    row_of_read[rows.row_members] = np.repeat(np.arange(rows.n_rows), sizes)
    # purity against the planted truth: a row is pure when all its reads come from one clone
    clone = lib.clone_of_read
    first_clone = clone[rows.row_members[rows.row_ptr[:-1]]]
    pure = np.ones(rows.n_rows, dtype=bool)
    np.logical_and.at(pure, row_of_read, clone == first_clone[row_of_read])
    big = sizes >= 2
    n_clones_seen = len(np.unique(clone))
    res = dict(cfg=cfg, scale=scale, reads=R, clones=int(lib.params["n_clones"]), buckets=out["n_buckets"], edges=out["n_edges"],
               rows=rows.n_rows, clusters=rows.n_clusters, pure_cluster_frac=float(pure[big].mean()) if big.any() else None,
               clusters_per_seen_clone=rows.n_clusters / n_clones_seen, max_cluster=int(sizes.max()),
               components=st["components"], links=st["links"], tests=st["tests"], pairs_scored=st["pairs_scored"],
               seed_len=st["seed_len"], huge_components=st["huge_components"], huge_rounds=st["huge_rounds"], gen_s=round(t_gen, 1), wall_ms=round(t_run * 1e3, 1),
               stage_ms={k[2:-3]: round(st[k] * 1e-6, 3) for k in st if k.startswith("t_")}, probe_ms=round(st["probe_ns"] * 1e-6, 3))
    if want_full:
        out2 = ctx.cluster_reads(*args, full_table=True, **p)
        for k in ("read_root", "read_pos", "row_size", "row_key", "row_bc", "row_up", "row_call_n"):
            assert np.array_equal(out[k], out2[k]), f"on-demand and full-table runs differ in {k}"
        res["full_table_edges"] = out2["n_edges"]
        res["full_table_probe_ms"] = round(ctx.stat("probe_ns") * 1e-6, 3)
    if want_oracle:
        from tests import util
        ra = pipeline.ReadArrays(seq=lib.seq, qual=lib.qual, length=args[2], lexrank=args[3], geno_ptr=args[5], geno_mut=args[6],
                                 geno_q=args[7], mut_rank=args[8], up_id=up_id)
        ids = np.arange(R).astype(str).tolist()
        muts = np.arange(len(lib.mut_codes)).astype(str).tolist()
        names = dict(ids=ids, mut_names=muts, up_names=None if up_id is None else np.arange(int(up_id.max()) + 1).astype(str).tolist())
        t0 = time.time()
        want = canon.canon_table(util.oracle_table_for_library(lib, ra, names, **p))
        res["oracle_s"] = round(time.time() - t0, 1)
        got = canon.canon_table(pipeline.table_from_outputs(out, ra, names))
        assert got == want, canon.diff_tables(got, want)
        res["oracle_equal"] = True
    if "oracle_merge" in sys.argv[3:]:
        # Full size, where the oracle's own pair search is out of reach: the GPU's full table (every pair <= 2*maxDiff)
        # is (i) spot-checked pair by pair against the DP, (ii) fed to the oracle's literal merge restatement, whose
        # rows must equal the GPU's -- rows, member order, consensus, everything.
        import ctypes as C
        from oracle import orc
        out2 = ctx.cluster_reads(*args, full_table=True, **p)
        E = out2["n_edges"]
        ei, ej, ed = (np.empty(E, dtype=np.int32) for _ in range(3))
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        assert ctx.lib.pcb_edit_pairs_fetch(ctx.h, 0, ptr(ei), ptr(ej), ptr(ed), E) == 0
        bor = out2["bucket_of_read"]
        U = out2["n_buckets"]
        first = np.full(U, R, dtype=np.int64)
        np.minimum.at(first, bor, np.arange(R))
        bseq, blen = lib.seq[first], lib.length[first]
        rng = np.random.default_rng(0)
        pick = rng.choice(E, size=min(E, 200_000), replace=False)
        for k in pick[:50_000]:
            a, b = ei[k], ej[k]
            assert orc.lev(bseq[a, :blen[a]].tobytes(), bseq[b, :blen[b]].tobytes()) == ed[k]
        # random non-edges must be farther than 2*maxDiff apart
        have = set(zip(ei[pick].tolist(), ej[pick].tolist()))
        ra_, rb_ = rng.integers(0, U, 50_000), rng.integers(0, U, 50_000)
        key = np.minimum(ra_, rb_).astype(np.int64) * U + np.maximum(ra_, rb_)
        edge_key = np.sort(ei.astype(np.int64) * U + ej)
        is_edge = edge_key[np.clip(np.searchsorted(edge_key, key), 0, E - 1)] == key
        for a, b, e_ in zip(ra_, rb_, is_edge):
            if a != b and not e_:
                assert orc.lev(bseq[a, :blen[a]].tobytes(), bseq[b, :blen[b]].tobytes()) > 2 * p["max_diff"]
        order = np.argsort(bor, kind="stable").astype(np.int32)
        bptr = np.zeros(U + 1, dtype=np.int32)
        np.cumsum(np.bincount(bor, minlength=U), out=bptr[1:])
        q, r, d = hostdata.canonical_rows(ei, ej, ed)
        prob = hostdata.MergeProblem(n_reads=R, n_buckets=U, bucket_ptr=bptr, bucket_reads=order, lexrank=args[3], bc_id=bor,
                                     up_id=up_id, geno_ptr=args[5], geno_mut=args[6], geno_q=args[7], mut_rank=args[8], **p)
        t0 = time.time()
        ref = orc.cluster(prob, q, r, d)
        res["oracle_merge_s"] = round(time.time() - t0, 1)
        for name, rws in (("full", hostdata.rows_from_merge_outputs(out2)), ("on_demand", rows)):
            for f in ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut"):
                assert np.array_equal(getattr(rws, f), getattr(ref, f)), (name, f)
        res["oracle_merge_equal"] = True
        res["full_table_edges"] = int(E)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
