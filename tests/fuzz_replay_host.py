#!/usr/bin/env python3
"""CPU fuzz: the replay the GPU threads execute (csrc/merge_core.h, compiled for the host by tests/host_emul) against the
literal restatement of pacybara_cluster.py (oracle/pcb_oracle.c:orc_cluster) on many small random libraries with harsh
parameters.  Not collected by pytest (it runs as long as you let it):

    python tests/fuzz_replay_host.py [seconds] [first_seed]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import orc  # noqa: E402
from pacybara_b200 import hostdata, synth  # noqa: E402
from tests import util  # noqa: E402


STATS = dict(small=0, components=0)


def one(seed: int) -> str:
    rng = np.random.default_rng(seed)
    combo = bool(rng.random() < 0.2)
    kw = dict(n_clones=int(rng.integers(3, 60)), n_reads=int(rng.integers(20, 500)), bc_len=int(rng.choice([8, 12, 25])),
              p_err=float(rng.choice([0.0, 0.02, 0.06])), p_near=float(rng.choice([0.0, 0.3, 0.7])),
              p_collide=float(rng.choice([0.0, 0.2, 0.5])), seed=seed, combo=combo,
              id_style=str(rng.choice(["padded", "plain"])), skew=float(rng.choice([0.0, 1.0, 2.0])))
    if rng.random() < 0.5:
        kw.update(p_private=float(rng.choice([0.1, 0.5])), p_private_hq=float(rng.choice([0.0, 0.3])), p_dropout=float(rng.choice([0.0, 0.3])))
    params = dict(max_diff=int(rng.integers(0, 4)), min_qual=int(rng.choice([1, 20, 32, 60])),
                  min_jaccard=float(rng.choice([0.05, 0.2, 0.5, 1.0])), min_matches=int(rng.choice([1, 1, 2, 3])))
    lib = synth.make_library(**kw)
    ids = lib.read_ids()
    reads = hostdata.FastqReads(ids, lib.seq, lib.qual, lib.length)
    b = util.oracle_buckets(reads)
    up_id = None
    if combo:
        ub = orc.precluster(lib.uptag_seq(), lib.qual, lib.uptag_len)
        up_id = ub["bucket_of_read"].astype(np.int32)
        if rng.random() < 0.3:
            up_id[rng.integers(0, lib.n_reads, 2)] = -1              # reads without uptag record: KeyError if clustered
    mut_names = lib.mut_strings()
    from pacybara_b200.pipeline import mut_rank_of
    prob = hostdata.MergeProblem(n_reads=lib.n_reads, n_buckets=b["n_buckets"], bucket_ptr=b["bucket_ptr"], bucket_reads=b["bucket_reads"],
                                 lexrank=hostdata.lexrank_of(ids), bc_id=b["bucket_of_read"], up_id=up_id,
                                 geno_ptr=lib.geno_ptr.astype(np.int32), geno_mut=lib.geno_mut.astype(np.int32),
                                 geno_q=lib.geno_q.astype(np.int32), mut_rank=mut_rank_of(mut_names), **params)
    ei, ej, ed = orc.edit_pairs(b["bucket_seq"], b["bucket_len"], 2 * params["max_diff"], method="brute")
    q, r, d = hostdata.canonical_rows(ei, ej, ed)
    if rng.random() < 0.3 and len(q):                                  # arbitrary row order / duplicates, like a foreign EDFILE
        perm = rng.permutation(len(q))
        q, r, d = q[perm], r[perm], d[perm]
    try:
        want = orc.cluster(prob, q, r, d)
        want_err = None
    except Exception as e:        # the reference's KeyError (read without uptag in a cluster)
        want, want_err = None, type(e).__name__
    pairs = hostdata.dedupe_ed_rows(q, r, d, prob.n_buckets, prob.max_diff)
    out = util.emul_merge(prob, pairs)
    if int(out["status"][0]) == 3 or want_err is not None:
        return "" if (int(out["status"][0]) == 3) == (want_err is not None) else f"error behaviour differs: emul status {out['status'][:2]}, oracle {want_err}"
    got = hostdata.rows_from_merge_outputs(out)
    for f in ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut"):
        if not np.array_equal(getattr(got, f), getattr(want, f)):
            return f"{f} differs ({kw}, {params})"
    # the warp kernel for small components (csrc/merge_small.cuh) on an emulated warp, lanes interleaved at random
    small = util.emul_merge_small(prob, pairs, seed=seed)
    STATS["small"] += small["n_small"]; STATS["components"] += small["n_components"]
    if int(small["status"][0]) == 3:
        return "small-component kernel reports a missing uptag, the oracle does not"
    got = hostdata.rows_from_merge_outputs(small)
    for f in ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut"):
        if not np.array_equal(getattr(got, f), getattr(want, f)):
            return f"small-component kernel: {f} differs ({kw}, {params})"
    # the generic warp replay (csrc/merge_warp.cuh) and the giant tier (csrc/merge_huge.cuh: rounds of deterministic
    # reservations) on emulated warps, with scratch tables and windows drawn small enough that rows are carried from round
    # to round and units wait for the big scratch -- the caller's row order (shuffled or not) is what they must replay
    for label, kw_emul in (("warp replay", dict(mode=2)),
                           ("giant tier", dict(mode=3, huge_cap=int(rng.choice([16, 32, 64, 1024])), huge_window=int(rng.choice([1, 3, 8, 40])),
                                               use_small=seed))):
        alt = util.emul_merge(prob, pairs, **kw_emul)
        STATS[label] = STATS.get(label, 0) + 1
        if int(alt["status"][0]) == 3:
            return f"{label} reports a missing uptag, the oracle does not"
        got = hostdata.rows_from_merge_outputs(alt)
        for f in ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut"):
            if not np.array_equal(getattr(got, f), getattr(want, f)):
                return f"{label}: {f} differs ({kw}, {params}, {kw_emul})"
        if int(alt["status"][2]) != int(out["status"][2]):
            return f"{label}: {int(alt['status'][2])} links, the thread replay made {int(out['status'][2])}"
    # on-demand distances: only the pairs a pass can visit are listed, the transitive rule scores the rest on the spot,
    # and two shards put together are the whole (canonical row order only: that is what the fused call produces)
    keep = ed <= params["max_diff"]
    q2, r2, d2 = hostdata.canonical_rows(ei[keep], ej[keep], ed[keep])
    qf, rf, df = hostdata.canonical_rows(ei, ej, ed)
    full_out = util.emul_merge(prob, hostdata.dedupe_ed_rows(qf, rf, df, prob.n_buckets, prob.max_diff))
    planes = util.pack_planes(b["bucket_seq"], b["bucket_len"])
    pairs2 = hostdata.dedupe_ed_rows(q2, r2, d2, prob.n_buckets, prob.max_diff)
    small2 = util.emul_merge_small(prob, pairs2, planes=planes, blen=b["bucket_len"], compute_k=2 * params["max_diff"], seed=seed + 7)
    parts = [util.emul_merge(prob, pairs2, shard=(s, 2), planes=planes, blen=b["bucket_len"], compute_k=2 * params["max_diff"]) for s in range(2)]
    # (the rows above may have been shuffled, which is a different problem: the canonical order decides here)
    if (int(full_out["status"][0]) == 3) != any(int(o["status"][0]) == 3 for o in parts):
        return "uptag error: on-demand shards and the full table disagree"
    if int(full_out["status"][0]) == 3:
        return ""
    full = hostdata.rows_from_merge_outputs(full_out)
    merged = {k: np.maximum(parts[0][k], parts[1][k]) for k in [n for n, _ in hostdata.MERGE_OUTPUTS] + ["call_mut"]}
    lazy = hostdata.rows_from_merge_outputs(merged)
    for f in ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut"):
        if not np.array_equal(getattr(lazy, f), getattr(full, f)):
            return f"on-demand + 2 shards: {f} differs ({kw}, {params})"
    lazy_small = hostdata.rows_from_merge_outputs(small2)
    for f in ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut"):
        if not np.array_equal(getattr(lazy_small, f), getattr(full, f)):
            return f"small-component kernel, on-demand distances: {f} differs ({kw}, {params})"
    return ""


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    t0, n, bad = time.time(), 0, []
    while time.time() - t0 < budget:
        msg = one(seed)
        if msg:
            bad.append((seed, msg))
            print(f"seed {seed}: {msg}", flush=True)
        seed += 1
        n += 1
    print(f"{n} libraries, {len(bad)} discrepancies, seeds {seed - n}..{seed - 1}; "
          f"{STATS['small']} of {STATS['components']} components went through the small-component warp kernel; "
          f"{STATS.get('warp replay', 0)} libraries through the generic warp replay, {STATS.get('giant tier', 0)} through the giant tier")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
