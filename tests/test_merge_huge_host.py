"""The generic warp replay (csrc/merge_warp.cuh) and the giant tier (csrc/merge_huge.cuh: seed step per bucket, main loop
in rounds of deterministic reservations, consensus per cluster) WITHOUT a GPU: tests/host_emul runs the very functions the
kernels call on emulated warps, the units of every "launch" in shuffled order, against the reference's golden tables and
the oracle.  Small scratch tables and windows force the paths a real library rarely takes: rows that lose their
reservation and are carried over, units that wait for the big scratch, the big scratch growing between batches."""
import numpy as np
import pytest

from oracle import orc
from pacybara_b200 import canon, hostdata, pipeline, synth
from tests import util

FILL = -(1 << 31)

CASES = [n for n in util.golden_names() if n not in ("kat10", "cfg1_full", "x_dense_sw")]
FIELDS = ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut")
SHAPES = [dict(mode=2), dict(mode=3, huge_cap=16, huge_window=4, use_small=1), dict(mode=3, huge_cap=1024, huge_window=12, use_small=2)]


@pytest.mark.parametrize("shape", SHAPES, ids=["warp", "huge-tight", "huge-roomy"])
@pytest.mark.parametrize("name", CASES)
def test_warp_and_giant_replay_match_the_reference(name, shape):
    case = util.load_golden(name)
    job = util.job_from_golden(case)
    p = job.problem
    pairs = hostdata.dedupe_ed_rows(job.ed_q, job.ed_r, job.ed_d, p.n_buckets, p.max_diff)
    out = util.emul_merge(p, pairs, **shape)
    rows = hostdata.rows_from_merge_outputs(out)
    if case["exit_code"] != 0:
        assert (rows.row_bc == hostdata.NEEDS_ALIGNMENT).any() or (rows.row_up == hostdata.NEEDS_ALIGNMENT).any()
        return
    got = util.table_of(job, rows)
    want = util.golden_table(case)
    assert got == want, canon.diff_tables(got, want)
    ref = orc.cluster(p, job.ed_q, job.ed_r, job.ed_d)          # beyond the canonical table: row order, member order
    for f in FIELDS:
        assert np.array_equal(getattr(rows, f), getattr(ref, f)), f
    assert int(out["status"][2]) == ref.n_links if hasattr(ref, "n_links") else True


@pytest.mark.parametrize("seed,kw,md", [
    (301, dict(n_clones=300, n_reads=2500, p_err=0.03, p_near=0.5, p_collide=0.2), 2),
    (302, dict(n_clones=200, n_reads=2000, p_err=0.04, p_near=0.6, p_private=0.5, p_private_hq=0.2, p_dropout=0.2), 2),
    (303, dict(n_clones=250, n_reads=2000, bc_len=10, p_err=0.03, p_near=0.5), 3),       # one giant component plus debris
    (304, dict(n_clones=40, n_reads=600, combo=True, p_err=0.01, p_near=0.4, skew=1.0), 3),
    (305, dict(n_clones=400, n_reads=2500, bc_len=9, p_err=0.02, p_near=0.3), 2),
])
def test_giant_replay_vs_oracle_on_random_libraries(seed, kw, md):
    """Full table and on-demand distances, split over two shards; scratch so small that most units need the big one."""
    lib = synth.make_library(seed=seed, **kw)
    ids = lib.read_ids()
    b = util.oracle_buckets(hostdata.FastqReads(ids, lib.seq, lib.qual, lib.length))
    ei, ej, ed = orc.edit_pairs(b["bucket_seq"], b["bucket_len"], 2 * md, method="seeded")
    q, r, d = hostdata.canonical_rows(ei, ej, ed)
    prob = hostdata.MergeProblem(n_reads=lib.n_reads, n_buckets=b["n_buckets"], bucket_ptr=b["bucket_ptr"],
                                 bucket_reads=b["bucket_reads"], lexrank=hostdata.lexrank_of(ids), bc_id=b["bucket_of_read"],
                                 up_id=None, geno_ptr=lib.geno_ptr.astype(np.int32), geno_mut=lib.geno_mut, geno_q=lib.geno_q,
                                 mut_rank=pipeline.mut_rank_of(lib.mut_strings()), max_diff=md, min_qual=32)
    ref = orc.cluster(prob, q, r, d)
    full = util.emul_merge(prob, hostdata.pairs_from_edges(ei, ej, ed, md), mode=3, huge_cap=32, huge_window=7, use_small=seed)
    keep = ed <= md
    planes = util.pack_planes(b["bucket_seq"], b["bucket_len"])
    lazy = [util.emul_merge(prob, hostdata.pairs_from_edges(ei[keep], ej[keep], ed[keep], md), shard=(s, 2), planes=planes,
                            blen=b["bucket_len"], compute_k=2 * md, mode=3, huge_cap=256, huge_window=50, use_small=seed + s)
            for s in range(2)]
    merged = {k: np.where(lazy[0][k] != FILL, lazy[0][k], lazy[1][k]) for k, _ in hostdata.MERGE_OUTPUTS if k != "row_call_off"}
    rows = hostdata.rows_from_merge_outputs(full)
    for f in FIELDS:
        assert np.array_equal(getattr(rows, f), getattr(ref, f)), f
    # the sharded run: same rows (call regions are per shard, so compare everything but their offsets)
    for k in ("read_root", "read_pos", "row_size", "row_bc", "row_up", "row_call_n"):
        assert np.array_equal(merged[k], full[k]), k


def test_dense_golden_through_the_giant_tier():
    """tests/golden/x_dense_sw: 3148 reads of a dense SWSW barcode space clustered by the unmodified reference -- 839
    buckets, 65 602 distance rows, nearly all in one component."""
    case = util.load_golden("x_dense_sw")
    job = util.job_from_golden(case)
    p = job.problem
    pairs = hostdata.dedupe_ed_rows(job.ed_q, job.ed_r, job.ed_d, p.n_buckets, p.max_diff)
    out = util.emul_merge(p, pairs, mode=3, huge_cap=256, huge_window=512, use_small=5)
    got = util.table_of(job, hostdata.rows_from_merge_outputs(out))
    want = util.golden_table(case)
    assert got == want, canon.diff_tables(got, want)
