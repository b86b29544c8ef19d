"""The warp kernel for small components (csrc/merge_small.cuh) WITHOUT a GPU: tests/host_emul runs the very source the
device executes on an emulated warp (32 coroutines that meet at every collective of csrc/simt.h, lane order shuffled at
every scheduling sweep) against the reference's golden tables and against the oracle.  Unit tests of the kernel source;
the CUDA launch path is covered by the -m gpu tests."""
import numpy as np
import pytest

from oracle import orc
from pacybara_b200 import canon, hostdata, pipeline, synth
from tests import util

CASES = [n for n in util.golden_names() if n != "kat10"]
FIELDS = ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut")


@pytest.mark.parametrize("name", CASES)
def test_small_component_kernel_matches_reference(name):
    case = util.load_golden(name)
    job = util.job_from_golden(case)
    p = job.problem
    pairs = hostdata.dedupe_ed_rows(job.ed_q, job.ed_r, job.ed_d, p.n_buckets, p.max_diff)
    out = util.emul_merge_small(p, pairs, seed=3)
    rows = hostdata.rows_from_merge_outputs(out)
    if case["exit_code"] != 0:
        assert (rows.row_bc == hostdata.NEEDS_ALIGNMENT).any() or (rows.row_up == hostdata.NEEDS_ALIGNMENT).any()
        return
    assert out["n_small"] > 0
    got = util.table_of(job, rows)
    want = util.golden_table(case)
    assert got == want, canon.diff_tables(got, want)
    ref = orc.cluster(p, job.ed_q, job.ed_r, job.ed_d)          # beyond the canonical table: row order, member order
    for f in FIELDS:
        assert np.array_equal(getattr(rows, f), getattr(ref, f)), f


def test_missing_uptag_is_reported():
    case = util.load_golden("kat2b")
    job = util.job_from_golden(case)
    job.problem.up_id = np.full_like(job.problem.up_id, -1)
    pairs = hostdata.dedupe_ed_rows(job.ed_q, job.ed_r, job.ed_d, job.problem.n_buckets, 2)
    assert int(util.emul_merge_small(job.problem, pairs)["status"][0]) == 3


@pytest.mark.parametrize("seed,kw,md", [
    (201, dict(n_clones=300, n_reads=2500, p_err=0.03, p_near=0.5, p_collide=0.2), 2),
    (202, dict(n_clones=200, n_reads=2000, p_err=0.04, p_near=0.6, p_private=0.5, p_private_hq=0.2, p_dropout=0.2), 2),
    (203, dict(n_clones=250, n_reads=2000, bc_len=10, p_err=0.03, p_near=0.5), 3),
    (204, dict(n_clones=60, n_reads=1500, combo=True, p_err=0.01, p_near=0.4, skew=1.0), 3),
    (205, dict(n_clones=400, n_reads=2500, p_err=0.02, p_near=0.3), 1),
])
def test_small_component_kernel_vs_oracle_on_random_libraries(seed, kw, md):
    """Full table and on-demand distances; components the kernel declines (> 32 reads or mutations) take the generic path."""
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
    full = util.emul_merge_small(prob, hostdata.pairs_from_edges(ei, ej, ed, md), seed=seed)
    keep = ed <= md
    lazy = util.emul_merge_small(prob, hostdata.pairs_from_edges(ei[keep], ej[keep], ed[keep], md),
                                 planes=util.pack_planes(b["bucket_seq"], b["bucket_len"]), blen=b["bucket_len"], compute_k=2 * md,
                                 seed=seed + 1)
    assert full["n_small"] >= 1          # (10-nt barcodes at maxDiff 3 are one giant component plus debris)
    for out in (full, lazy):
        rows = hostdata.rows_from_merge_outputs(out)
        for f in FIELDS:
            assert np.array_equal(getattr(rows, f), getattr(ref, f)), f


def test_the_emulator_catches_a_missing_sync():
    """The harness itself: a lane that reads another lane's shared-memory write without a warp sync in between must
    produce wrong answers under the shuffled lane order (so a kernel that forgets one cannot pass by luck)."""
    import ctypes as C
    lib = util.emul_lib()
    assert lib.emul_selftest_missing_sync() == 1        # racy variant: detected
    assert lib.emul_selftest_with_sync() == 0           # synced variant: clean
