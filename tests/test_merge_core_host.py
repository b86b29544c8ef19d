"""The per-component replay (csrc/merge_core.h), compiled for the host, against the reference's
golden tables and against the oracle.  This exercises the exact function the GPU threads run,
without a GPU; the CUDA launch path itself is covered by the -m gpu tests."""
import numpy as np
import pytest

from oracle import orc
from pacybara_b200 import canon, hostdata, synth
from tests import util

CASES = [n for n in util.golden_names() if n != "kat10"]


def _run(job, shard_count=1):
    p = job.problem
    pairs = hostdata.dedupe_ed_rows(job.ed_q, job.ed_r, job.ed_d, p.n_buckets, p.max_diff)
    outs = [util.emul_merge(p, pairs, (s, shard_count)) for s in range(shard_count)]
    merged = {k: np.maximum.reduce([o[k] for o in outs]) for k, _ in hostdata.MERGE_OUTPUTS}
    merged["call_mut"] = np.maximum.reduce([o["call_mut"] for o in outs])
    return hostdata.rows_from_merge_outputs(merged), outs


@pytest.mark.parametrize("name", CASES)
def test_replay_matches_reference(name):
    case = util.load_golden(name)
    job = util.job_from_golden(case)
    rows, _ = _run(job)
    if case["exit_code"] != 0:
        assert (rows.row_bc == hostdata.NEEDS_ALIGNMENT).any() or (rows.row_up == hostdata.NEEDS_ALIGNMENT).any()
        return
    got = util.table_of(job, rows)
    want = util.golden_table(case)
    assert got == want, canon.diff_tables(got, want)


@pytest.mark.parametrize("name", ["rand_dense_d2", "rand_collide", "rand_d3_combo"])
def test_sharded_replay_is_identical(name):
    """Components dealt to 3 shards and combined with an element-wise max (the multi-GPU merge)."""
    case = util.load_golden(name)
    job = util.job_from_golden(case)
    rows, _ = _run(job, shard_count=3)
    got = util.table_of(job, rows)
    assert got == util.golden_table(case)


@pytest.mark.parametrize("name", ["rand_dense_d2", "rand_lexids", "rand_private_heavy"])
def test_row_and_member_order_match_oracle(name):
    """Beyond the canonical table: cluster rows come out in the reference's creation order and
    members in its list order (the oracle keeps both)."""
    case = util.load_golden(name)
    job = util.job_from_golden(case)
    rows, _ = _run(job)
    ref = orc.cluster(job.problem, job.ed_q, job.ed_r, job.ed_d)
    assert rows.n_clusters == ref.n_clusters
    assert np.array_equal(rows.row_ptr, ref.row_ptr)
    assert np.array_equal(rows.row_members, ref.row_members)
    assert np.array_equal(rows.row_bc, ref.row_bc)
    assert np.array_equal(rows.row_up, ref.row_up)
    assert np.array_equal(rows.call_ptr, ref.call_ptr)
    assert np.array_equal(rows.call_mut, ref.call_mut)


@pytest.mark.parametrize("name", ["rand_dense_d2", "rand_collide", "rand_d3_combo", "rand_d1", "rand_shortbc"])
def test_on_demand_distances_give_the_same_clusters(name):
    """The fused path's default: the pair list only holds dist <= maxDiff, the distances in
    (maxDiff, 2*maxDiff] that the transitive rule needs are scored on demand from the bit planes."""
    case = util.load_golden(name)
    job = util.job_from_golden(case)
    p = job.problem
    keep = job.ed_d <= p.max_diff
    pairs = hostdata.dedupe_ed_rows(job.ed_q[keep], job.ed_r[keep], job.ed_d[keep], p.n_buckets, p.max_diff)
    seqs = [s.encode() for s in job.bc_names]
    # bucket b's barcode: bc_id of its first read
    first = p.bucket_reads[p.bucket_ptr[:-1]]
    mat, lens = hostdata.strings_to_matrix([seqs[i] for i in p.bc_id[first]], 64)
    out = util.emul_merge(p, pairs, planes=util.pack_planes(mat, lens), blen=lens, compute_k=2 * p.max_diff)
    got = util.table_of(job, hostdata.rows_from_merge_outputs(out))
    assert got == util.golden_table(case)
    # and without them the truncated table is NOT enough (the test would be vacuous otherwise)
    if name in ("rand_dense_d2", "rand_d3_combo"):
        out2 = util.emul_merge(p, pairs)
        assert util.table_of(job, hostdata.rows_from_merge_outputs(out2)) != util.golden_table(case)


@pytest.mark.parametrize("seed,kw", [
    (101, dict(n_clones=400, n_reads=6000, p_err=0.03, p_near=0.5, p_collide=0.2, skew=1.2)),
    (102, dict(n_clones=50, n_reads=4000, p_err=0.04, p_near=0.6, p_private=0.4, p_private_hq=0.1, p_dropout=0.2)),
    (103, dict(n_clones=300, n_reads=3000, bc_len=10, p_err=0.03, p_near=0.5)),
])
def test_replay_vs_oracle_on_fresh_random_libraries(seed, kw):
    """Beyond the committed goldens: dense, skewed libraries (big buckets, many satellites per cluster,
    collisions) through the oracle and through the replay, full table and on-demand distances."""
    lib = synth.make_library(seed=seed, **kw)
    ids = lib.read_ids()
    reads = hostdata.FastqReads(ids, lib.seq, lib.qual, lib.length)
    b = util.oracle_buckets(reads)
    md = 2
    ei, ej, ed = orc.edit_pairs(b["bucket_seq"], b["bucket_len"], 2 * md, method="seeded")
    q, r, d = hostdata.canonical_rows(ei, ej, ed)
    from pacybara_b200 import pipeline
    prob = hostdata.MergeProblem(n_reads=lib.n_reads, n_buckets=b["n_buckets"], bucket_ptr=b["bucket_ptr"],
                                 bucket_reads=b["bucket_reads"], lexrank=hostdata.lexrank_of(ids), bc_id=b["bucket_of_read"],
                                 up_id=None, geno_ptr=lib.geno_ptr.astype(np.int32), geno_mut=lib.geno_mut, geno_q=lib.geno_q,
                                 mut_rank=pipeline.mut_rank_of(lib.mut_strings()), max_diff=md, min_qual=32)
    ref = orc.cluster(prob, q, r, d)
    full = hostdata.rows_from_merge_outputs(util.emul_merge(prob, hostdata.pairs_from_edges(ei, ej, ed, md)))
    keep = ed <= md
    fast = hostdata.rows_from_merge_outputs(util.emul_merge(
        prob, hostdata.pairs_from_edges(ei[keep], ej[keep], ed[keep], md),
        planes=util.pack_planes(b["bucket_seq"], b["bucket_len"]), blen=b["bucket_len"], compute_k=2 * md))
    for rows in (full, fast):
        for f in ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut"):
            assert np.array_equal(getattr(rows, f), getattr(ref, f)), f


def test_many_small_random_libraries_host():
    """The same 120 extreme tiny libraries the GPU test uses, through the host-compiled replay."""
    from pacybara_b200 import pipeline
    rng = np.random.default_rng(2024)
    for case in range(120):
        kw = dict(n_clones=int(rng.integers(1, 40)), n_reads=int(rng.integers(1, 400)), bc_len=int(rng.choice([8, 12, 25])),
                  p_err=float(rng.choice([0.0, 0.01, 0.05])), p_near=float(rng.choice([0.0, 0.3, 0.8])),
                  p_collide=float(rng.choice([0.0, 0.3])), p_private=float(rng.choice([0.0, 0.15, 0.6])),
                  p_private_hq=float(rng.choice([0.0, 0.2])), p_dropout=float(rng.choice([0.0, 0.3])),
                  skew=float(rng.choice([0.0, 1.5])), seed=1000 + case, id_style=str(rng.choice(["padded", "plain"])))
        params = dict(max_diff=int(rng.integers(0, 4)), min_qual=int(rng.choice([5, 32, 60])),
                      min_jaccard=float(rng.choice([0.1, 0.2, 0.5, 0.9])), min_matches=int(rng.integers(1, 4)))
        lib = synth.make_library(**kw)
        ids = lib.read_ids()
        b = util.oracle_buckets(hostdata.FastqReads(ids, lib.seq, lib.qual, lib.length))
        md = params["max_diff"]
        ei, ej, ed = orc.edit_pairs(b["bucket_seq"], b["bucket_len"], 2 * md, method="brute")
        q, r, d = hostdata.canonical_rows(ei, ej, ed)
        prob = hostdata.MergeProblem(n_reads=lib.n_reads, n_buckets=b["n_buckets"], bucket_ptr=b["bucket_ptr"],
                                     bucket_reads=b["bucket_reads"], lexrank=hostdata.lexrank_of(ids), bc_id=b["bucket_of_read"],
                                     up_id=None, geno_ptr=lib.geno_ptr.astype(np.int32), geno_mut=lib.geno_mut, geno_q=lib.geno_q,
                                     mut_rank=pipeline.mut_rank_of(lib.mut_strings()), **params)
        ref = orc.cluster(prob, q, r, d)
        keep = ed <= md
        runs = [util.emul_merge(prob, hostdata.pairs_from_edges(ei, ej, ed, md)),
                util.emul_merge(prob, hostdata.pairs_from_edges(ei[keep], ej[keep], ed[keep], md),
                                planes=util.pack_planes(b["bucket_seq"], b["bucket_len"]), blen=b["bucket_len"], compute_k=2 * md)]
        for out in runs:
            rows = hostdata.rows_from_merge_outputs(out)
            for f in ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut"):
                assert np.array_equal(getattr(rows, f), getattr(ref, f)), (case, f, kw, params)
