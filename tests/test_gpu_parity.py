"""Parity of the CUDA path with the oracle / the reference's golden tables, through the C-ABI.

Bit-exact is the bar everywhere (integer and index work; the one float compare is the same
IEEE fp64 divide the reference does)."""
import io
import os

import numpy as np
import pytest

from pacybara_b200 import canon, hostdata, synth
from tests import util

pytestmark = pytest.mark.gpu

CASES = [n for n in util.golden_names() if n != "kat10"]


@pytest.fixture(scope="module")
def ctx():
    from pacybara_b200.api import Context
    c = Context(0)
    yield c
    c.close()


# ----------------------------------------------------------------------------- a1 bucketing
@pytest.mark.parametrize("name", util.golden_names())
def test_bucket_golden(ctx, name):
    case = util.load_golden(name)
    reads = hostdata.parse_fastq_for_precluster(io.StringIO(case["fastq"]))
    res = ctx.bucket(reads.seq, reads.qual, reads.length)
    assert util.preclust_text(reads, res) == case["preclust"]


@pytest.mark.parametrize("cfg,scale", [("cfg1", 1.0), ("cfg3", 0.02), ("cfg2", 1.0)])
def test_bucket_vs_oracle(ctx, cfg, scale):
    from oracle import orc
    lib = synth.make_config(cfg, seed=3, scale=scale)
    res = ctx.bucket(lib.seq, lib.qual, lib.length)
    ref = orc.precluster(lib.seq, lib.qual, lib.length)
    assert res["n_buckets"] == ref["n_buckets"]
    assert np.array_equal(res["bucket_of_read"], ref["bucket_of_read"])
    assert np.array_equal(res["first_read"], ref["first_read"])
    assert np.array_equal(res["bucket_qsum"], ref["qsum"])
    # arrival order inside every bucket == stable sort of the reads by bucket
    assert np.array_equal(res["bucket_reads"], np.argsort(ref["bucket_of_read"], kind="stable"))
    assert np.array_equal(np.diff(res["bucket_ptr"]), np.bincount(ref["bucket_of_read"], minlength=ref["n_buckets"]))


def test_bucket_edge_shapes(ctx):
    # a single read; all reads identical; every read different; 64-nt barcodes
    one = hostdata.strings_to_matrix([b"ACGT"])
    r = ctx.bucket(one[0], np.full_like(one[0], 73), one[1])
    assert r["n_buckets"] == 1 and r["bucket_qsum"][0].tolist() == [40, 40, 40, 40]
    same = hostdata.strings_to_matrix([b"ACGTACGT"] * 1000)
    r = ctx.bucket(same[0], np.full_like(same[0], 35), same[1])
    assert r["n_buckets"] == 1 and int(r["bucket_ptr"][1]) == 1000
    assert r["bucket_qsum"][0].tolist() == [93] * 8                  # 1000 * 2 capped at 93
    assert np.array_equal(r["bucket_reads"], np.arange(1000))
    rng = np.random.default_rng(0)
    long_ = rng.integers(0, 4, size=(500, 64))
    mat = synth.BASES[long_].astype(np.uint8)
    r = ctx.bucket(mat, np.full_like(mat, 40), np.full(500, 64, dtype=np.int32))
    assert r["n_buckets"] == len({row.tobytes() for row in mat})
    # same bases, different lengths are different barcodes
    pre = hostdata.strings_to_matrix([b"AAAA", b"AAA", b"AAAA", b"AAAAA", b"AAA"])
    r = ctx.bucket(pre[0], np.full_like(pre[0], 50), pre[1])
    assert r["bucket_of_read"].tolist() == [0, 1, 0, 2, 1]


def test_barcodes_that_do_not_fit_the_planes_are_kept_as_opaque_strings(ctx):
    """The reference treats barcodes as opaque strings: an N, a 70-nt insertion call or an empty barcode must not abort
    the library.  They bucket with their exact copies, in first-seen order like everything else, and the pair search
    leaves them alone."""
    from pacybara_b200._lib import PcbError, PCB_E_INPUT
    long_ = b"ACGT" * 17 + b"AC"                                   # 70 nt
    seqs = [b"ACGTACGTAC", b"ACNTACGTAC", long_, b"ACGTACGTAC", b"ACNTACGTAC", b"", long_, b"ACGTACGTAA", b"", long_[:-1] + b"G"]
    mat, lens = hostdata.strings_to_matrix(seqs)
    r = ctx.bucket(mat, np.full_like(mat, 50), lens)
    assert r["bucket_of_read"].tolist() == [0, 1, 2, 0, 1, 3, 2, 4, 3, 5]
    assert r["n_buckets"] == 6 and np.diff(r["bucket_ptr"]).tolist() == [2, 2, 2, 2, 1, 1]
    assert r["bucket_reads"].tolist() == [0, 3, 1, 4, 2, 6, 5, 8, 7, 9]
    assert r["bucket_qsum"][2, :70].tolist() == [34] * 70 and r["bucket_qsum"][3].tolist() == [0] * mat.shape[1]
    # pairs: only among the barcodes that fit the planes (ACGTACGTAC ~ ACGTACGTAA); the N barcode is one edit from
    # bucket 0 as a string, but it is opaque here
    ei, ej, ed = ctx.edit_pairs(r["bucket_seq"], r["bucket_len"], 2)
    assert list(zip(ei.tolist(), ej.tolist(), ed.tolist())) == [(0, 4, 1)]
    # the whole path: rows for everybody
    R = len(seqs)
    out = ctx.cluster_reads(mat, None, lens, np.arange(R, dtype=np.int32), None, np.zeros(R + 1, dtype=np.int32), np.zeros(0, dtype=np.int32),
                            np.zeros(0, dtype=np.int32), np.zeros(1, dtype=np.int32), max_diff=2)
    rows = hostdata.rows_from_merge_outputs(out)
    members = sorted(sorted(rows.row_members[rows.row_ptr[i]:rows.row_ptr[i + 1]].tolist()) for i in range(rows.n_rows))
    assert members == [[0, 3, 7], [1, 4], [2, 6], [5, 8], [9]]          # (7: one read against two, one edit away, no calls: it joins)
    # what is still refused: a length the matrix cannot hold
    with pytest.raises(PcbError) as e:
        ctx.bucket(mat, np.full_like(mat, 50), np.where(np.arange(R) == 4, mat.shape[1] + 1, lens).astype(np.int32))
    assert e.value.code == PCB_E_INPUT
    only_opaque = hostdata.strings_to_matrix([b"NNNN", b"NNNA", b"NNNN"])
    assert ctx.bucket(only_opaque[0], np.full_like(only_opaque[0], 40), only_opaque[1])["bucket_of_read"].tolist() == [0, 1, 0]
    assert len(ctx.edit_pairs(only_opaque[0][:2], only_opaque[1][:2], 2)[0]) == 0


# ----------------------------------------------------------------------------- a2+a3 distances
def _edges_equal(got, want):
    for g, w in zip(got, want):
        assert np.array_equal(np.asarray(g), np.asarray(w))


@pytest.mark.parametrize("k", [0, 1, 2, 4, 6])
def test_edit_pairs_vs_exhaustive_dp(ctx, k):
    """Every pair goes through the full DP in the oracle (no candidate filter)."""
    from oracle import orc
    lib = synth.make_library(n_clones=120, n_reads=2500, p_err=0.03, p_near=0.4, seed=21 + k)
    b = util.oracle_buckets(hostdata.FastqReads([], lib.seq, lib.qual, lib.length))
    got = ctx.edit_pairs(b["bucket_seq"], b["bucket_len"], k)
    want = orc.edit_pairs(b["bucket_seq"], b["bucket_len"], k, method="brute")
    assert len(want[0]) > 0 or k == 0
    _edges_equal(got, want)


@pytest.mark.parametrize("bc_len,combo,k", [(25, False, 4), (25, True, 6), (12, False, 2), (6, False, 4), (3, False, 3)])
def test_edit_pairs_lengths(ctx, bc_len, combo, k):
    """One-word and two-word barcodes, and barcodes shorter than k+1 (q = 0: one bin, all pairs)."""
    from oracle import orc
    lib = synth.make_library(n_clones=150, n_reads=1500, bc_len=bc_len, combo=combo, p_err=0.03, p_near=0.3, seed=5)
    b = util.oracle_buckets(hostdata.FastqReads([], lib.seq, lib.qual, lib.length))
    got = ctx.edit_pairs(b["bucket_seq"], b["bucket_len"], k)
    want = orc.edit_pairs(b["bucket_seq"], b["bucket_len"], k, method="brute")
    _edges_equal(got, want)


def test_edit_pairs_cfg1_full_and_sharded(ctx):
    from oracle import orc
    lib = synth.make_config("cfg1", seed=1)
    b = util.oracle_buckets(hostdata.FastqReads([], lib.seq, lib.qual, lib.length))
    U = b["n_buckets"]
    want = orc.edit_pairs(b["bucket_seq"], b["bucket_len"], 4, method="seeded")
    got = ctx.edit_pairs(b["bucket_seq"], b["bucket_len"], 4)
    _edges_equal(got, want)
    assert ctx.stat("edges") == len(want[0]) and ctx.stat("pairs_scored") >= len(want[0])
    # query shards: disjoint, union = everything
    cuts = [0, U // 3, U // 2, U]
    parts = [ctx.edit_pairs(b["bucket_seq"], b["bucket_len"], 4, q_range=(cuts[i], cuts[i + 1])) for i in range(3)]
    assert sum(len(p[0]) for p in parts) == len(want[0])
    ei, ej, ed = (np.concatenate([p[t] for p in parts]) for t in range(3))
    order = np.lexsort((ej, ei))
    _edges_equal((ei[order], ej[order], ed[order]), want)


# ----------------------------------------------------------------------------- a4..a13 merge
def _merge_table(ctx, job, shards=1):
    p = job.problem
    pairs = hostdata.dedupe_ed_rows(job.ed_q, job.ed_r, job.ed_d, p.n_buckets, p.max_diff)
    outs = [ctx.merge(p, pairs, (s, shards)) for s in range(shards)]
    merged = {k: np.maximum.reduce([o[k] for o in outs]) for k in outs[0]}
    return hostdata.rows_from_merge_outputs(merged)


@pytest.mark.parametrize("name", CASES)
def test_merge_golden(ctx, name):
    case = util.load_golden(name)
    job = util.job_from_golden(case)
    rows = _merge_table(ctx, job)
    if case["exit_code"] != 0:
        assert (rows.row_bc == hostdata.NEEDS_ALIGNMENT).any() or (rows.row_up == hostdata.NEEDS_ALIGNMENT).any()
        return
    got = util.table_of(job, rows)
    want = util.golden_table(case)
    assert got == want, canon.diff_tables(got, want)


@pytest.mark.parametrize("name", ["rand_dense_d2", "rand_collide", "rand_d3_combo"])
def test_merge_sharded(ctx, name):
    case = util.load_golden(name)
    job = util.job_from_golden(case)
    got = util.table_of(job, _merge_table(ctx, job, shards=4))
    assert got == util.golden_table(case)


def test_merge_order_matches_oracle(ctx):
    from oracle import orc
    case = util.load_golden("rand_lexids")
    job = util.job_from_golden(case)
    rows = _merge_table(ctx, job)
    ref = orc.cluster(job.problem, job.ed_q, job.ed_r, job.ed_d)
    for f in ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut"):
        assert np.array_equal(getattr(rows, f), getattr(ref, f)), f


def test_missing_uptag_is_an_error(ctx):
    from pacybara_b200._lib import PcbError, PCB_E_UPTAG
    case = util.load_golden("kat2b")
    job = util.job_from_golden(case)
    job.problem.up_id = job.problem.up_id.copy()
    job.problem.up_id[:] = -1
    pairs = hostdata.dedupe_ed_rows(job.ed_q, job.ed_r, job.ed_d, job.problem.n_buckets, 2)
    with pytest.raises(PcbError) as e:
        ctx.merge(job.problem, pairs)
    assert e.value.code == PCB_E_UPTAG


# ----------------------------------------------------------------------------- the fused path
def _fused_vs_oracle(ctx, lib, device=False, full_table=False, **params):
    from pacybara_b200 import pipeline
    ra, names = pipeline.arrays_from_library(lib, ctx)
    p = dict(max_diff=lib.params.get("max_diff", 2), min_qual=lib.params.get("min_qual", 32), min_jaccard=0.2, min_matches=1)
    p.update(params)
    want = canon.canon_table(util.oracle_table_for_library(lib, ra, names, **p))
    p["full_table"] = full_table
    if device:
        import torch
        dev = {k: (None if v is None else torch.from_numpy(np.ascontiguousarray(v)).cuda())
               for k, v in vars(ra).items()}
        out = ctx.cluster_reads(dev["seq"], dev["qual"], dev["length"], dev["lexrank"], dev["up_id"], dev["geno_ptr"],
                                dev["geno_mut"], dev["geno_q"], dev["mut_rank"], **p)
        out = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}
    else:
        out = pipeline.cluster_reads(ctx, ra, **p)
    got = canon.canon_table(pipeline.table_from_outputs(out, ra, names))
    assert len(got) == len(want)
    assert got == want, canon.diff_tables(got, want)
    return out


@pytest.mark.parametrize("seed,full", [(1, False), (2, False), (1, True)])
def test_fused_cfg1(ctx, seed, full):
    """full=False: pairs up to maxDiff + on-demand transitive distances; full=True: the whole table."""
    out = _fused_vs_oracle(ctx, synth.make_config("cfg1", seed=seed), full_table=full)
    assert ctx.stat("links") > 0


def test_fused_cfg1_device_memory(ctx):
    _fused_vs_oracle(ctx, synth.make_config("cfg1", seed=4), device=True)


@pytest.mark.parametrize("device", [False, True])
@pytest.mark.parametrize("cfg,scale", [("cfg1", 1.0), ("cfg3", 0.01), ("cfg2", 0.1)])
def test_fused_compact_rows(ctx, cfg, scale, device):
    """PCB_CLUSTER_COMPACT_ROWS: one entry per row instead of one per read -- the same rows, member order and calls."""
    from pacybara_b200 import pipeline
    lib = synth.make_config(cfg, seed=6, scale=scale)
    ra, names = pipeline.arrays_from_library(lib, ctx)
    if device:
        import torch
        ra = pipeline.ReadArrays(**{k: (torch.from_numpy(np.ascontiguousarray(v)).cuda() if v is not None else None)
                                    for k, v in vars(ra).items()})
    host = lambda d: {k: (v.cpu().numpy() if hasattr(v, "data_ptr") else v) for k, v in d.items()}
    per_read = host(pipeline.cluster_reads(ctx, ra))
    compact = host(pipeline.cluster_reads(ctx, ra, compact_rows=True))
    n_rows = int((per_read["row_size"] > 0).sum())
    assert ctx.stat("rows") == n_rows == compact["row_rep"].shape[0] == compact["row_size"].shape[0]
    assert np.array_equal(compact["row_rep"], np.nonzero(per_read["row_size"] > 0)[0])
    assert compact["call_mut"].shape[0] == ctx.stat("calls") == int(per_read["row_call_n"][compact["row_rep"]].clip(0).sum())
    for k in ("read_root", "read_pos", "bucket_of_read"):
        assert np.array_equal(compact[k], per_read[k]), k
    a, b = hostdata.rows_from_merge_outputs(per_read), hostdata.rows_from_merge_outputs(compact)
    for f in ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert a.n_clusters == b.n_clusters


def test_fused_cfg1_golden_digest(ctx):
    """The committed digest of the reference's own cfg-1 table (tests/golden/cfg1_seed1)."""
    import hashlib
    from oracle.make_golden import lib_texts
    from pacybara_b200 import pipeline
    case = util.load_golden("cfg1_seed1")
    lib = synth.make_config("cfg1", seed=1)
    fastq, geno, uptag = lib_texts(lib)
    if hashlib.sha256((fastq + geno).encode()).hexdigest() != case["input_digest"]:
        pytest.skip("numpy random streams differ from the build container's")
    # through the text parsers: the reference drops reads whose quality line starts with '@'
    ra, names = pipeline.arrays_from_text(io.StringIO(fastq), io.StringIO(geno), io.StringIO(uptag))
    out = pipeline.cluster_reads(ctx, ra, full_table=True, **case["params"])
    got = canon.canon_table(pipeline.table_from_outputs(out, ra, names))
    assert 2 * out["n_edges"] == case["n_ed_rows"]
    assert canon.table_digest(got) == case["digest"]
    out = pipeline.cluster_reads(ctx, ra, **case["params"])          # on-demand distances: same table
    assert out["n_edges"] < case["n_ed_rows"] // 2
    assert canon.table_digest(canon.canon_table(pipeline.table_from_outputs(out, ra, names))) == case["digest"]


@pytest.mark.parametrize("kw", [
    dict(n_clones=300, n_reads=4000, p_err=0.03, p_near=0.4, p_collide=0.2, seed=31),
    dict(n_clones=300, n_reads=4000, combo=True, p_err=0.01, p_near=0.3, seed=32, max_diff=3),
    dict(n_clones=200, n_reads=5000, p_err=0.02, p_near=0.4, skew=1.5, seed=33, id_style="plain"),
    dict(n_clones=400, n_reads=3000, p_err=0.02, p_near=0.3, seed=34, max_diff=1),
])
def test_fused_random(ctx, kw):
    kw = dict(kw)
    md = kw.pop("max_diff", 2)
    lib = synth.make_library(**kw)
    for full in (False, True):
        try:
            _fused_vs_oracle(ctx, lib, max_diff=md, full_table=full)
        except RuntimeError as e:      # a tied >=3-read consensus: both sides must flag it
            assert "muscle" in str(e)


def test_fused_cfg2_scaled_vs_oracle(ctx):
    _fused_vs_oracle(ctx, synth.make_config("cfg2", seed=1, scale=0.2))


def test_fused_cfg2_full_size_properties(ctx):
    """cfg 2 at full size (1M reads): invariants that need no oracle run."""
    from pacybara_b200 import pipeline
    lib = synth.make_config("cfg2", seed=1)
    ra, names = pipeline.arrays_from_library(lib, ctx)
    out = pipeline.cluster_reads(ctx, ra)
    R = lib.n_reads
    rows = hostdata.rows_from_merge_outputs(out)             # raises unless every read is in exactly one row
    assert int(rows.row_ptr[-1]) == R
    assert np.array_equal(np.sort(rows.row_members), np.arange(R))
    bor = out["bucket_of_read"]
    # reads of one row are within 2*maxDiff buckets of each other only through edges: same bucket => same component,
    # and a cluster never mixes reads whose clones have different planted barcodes AND disjoint genotypes
    sizes = np.diff(rows.row_ptr)
    assert sizes.max() < 2000 and (sizes >= 1).all()
    # idempotence of bucketing: bucketing the unique barcodes again gives the identity
    first = np.full(out["n_buckets"], R, dtype=np.int64)
    np.minimum.at(first, bor, np.arange(R))
    again = ctx.bucket(lib.seq[first], lib.qual[first], lib.length[first], want_qsum=False)
    assert again["n_buckets"] == out["n_buckets"]
    assert np.array_equal(again["bucket_of_read"], np.arange(out["n_buckets"]))
    # most clones come back as one dominant cluster
    top = rows.row_members[rows.row_ptr[:-1]]
    assert rows.n_clusters > 0.9 * lib.params["n_clones"]


# ----------------------------------------------------------------------------- the multi-GPU decomposition on one GPU
@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("kw", [dict(cfg="cfg1"), dict(n_clones=300, n_reads=5000, p_err=0.03, p_near=0.4, p_collide=0.2, seed=41),
                                dict(n_clones=200, n_reads=3000, combo=True, p_err=0.01, p_near=0.3, seed=42, max_diff=3)])
def test_sharded_algorithm_with_virtual_ranks(ctx, kw, world):
    """What N ranks compute in pacybara_b200.multigpu -- target slices of the pair search, components dealt to shards,
    compacted shards put back together -- executed rank after rank on this GPU (no NCCL): the rows must be those of the
    single call, which the other tests pin to the oracle.  (The collectives themselves run in bench.py --gpus N, whose
    line carries the same comparison, and in tests/multigpu_check.py.)"""
    import torch
    from pacybara_b200 import multigpu, pipeline
    kw = dict(kw)
    md = kw.pop("max_diff", 2)
    lib = synth.make_config(kw["cfg"], seed=5) if "cfg" in kw else synth.make_library(**kw)
    ra, names = pipeline.arrays_from_library(lib, ctx)
    params = dict(max_diff=md, min_qual=32, min_jaccard=0.2, min_matches=1)
    dev = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in vars(ra).items() if v is not None}
    single = ctx.cluster_reads(dev["seq"], dev["qual"], dev["length"], dev["lexrank"], dev.get("up_id"), dev["geno_ptr"], dev["geno_mut"],
                               dev["geno_q"], dev["mut_rank"], **params)
    single = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in single.items()}
    want = hostdata.rows_from_merge_outputs(single)
    for full in (False, True):
        whole = multigpu.cluster_reads_virtual(ctx, dev, params, world, full_table=full)
        assert np.array_equal(whole["bucket_of_read"], single["bucket_of_read"])
        _rows_equal(hostdata.rows_from_merge_outputs(whole), want)
        assert sum(whole["shard_rows"]) == want.n_rows and min(whole["shard_rows"]) > 0


# ----------------------------------------------------------------------------- many small random libraries
def _rows_equal(a, b):
    for f in ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


@pytest.mark.parametrize("shape", ["thread", "warp", "auto", "huge"])
def test_many_small_random_libraries_both_replay_shapes(ctx, shape, monkeypatch):
    """120 tiny libraries with extreme settings (dense neighbours, collisions, short barcodes, d = 1..3,
    minMatches 1..3, minJaccard 0.1..0.9, low/high minQual, skewed clone sizes): the fused path must give the
    oracle's rows exactly -- row order, member order, consensus -- with either replay kernel."""
    from oracle import orc
    from pacybara_b200 import pipeline
    monkeypatch.setenv("PCB_REPLAY", shape)
    if shape == "huge":                      # giant tier with scratch and windows so small that rows wait and are carried over
        monkeypatch.setenv("PCB_HUGE_CAP", "32")
        monkeypatch.setenv("PCB_HUGE_WINDOW", "5")
    rng = np.random.default_rng(2024)
    n_alignment = 0
    for case in range(120):
        kw = dict(n_clones=int(rng.integers(1, 40)), n_reads=int(rng.integers(1, 400)), bc_len=int(rng.choice([8, 12, 25])),
                  p_err=float(rng.choice([0.0, 0.01, 0.05])), p_near=float(rng.choice([0.0, 0.3, 0.8])),
                  p_collide=float(rng.choice([0.0, 0.3])), p_private=float(rng.choice([0.0, 0.15, 0.6])),
                  p_private_hq=float(rng.choice([0.0, 0.2])), p_dropout=float(rng.choice([0.0, 0.3])),
                  skew=float(rng.choice([0.0, 1.5])), seed=1000 + case, id_style=str(rng.choice(["padded", "plain"])))
        params = dict(max_diff=int(rng.integers(0, 4)), min_qual=int(rng.choice([5, 32, 60])),
                      min_jaccard=float(rng.choice([0.1, 0.2, 0.5, 0.9])), min_matches=int(rng.integers(1, 4)))
        lib = synth.make_library(**kw)
        ra, names = pipeline.arrays_from_library(lib, ctx)
        reads = hostdata.FastqReads(names["ids"], lib.seq, lib.qual, lib.length)
        b = util.oracle_buckets(reads)
        ei, ej, ed = orc.edit_pairs(b["bucket_seq"], b["bucket_len"], 2 * params["max_diff"], method="brute")
        q, r, d = hostdata.canonical_rows(ei, ej, ed)
        prob = hostdata.MergeProblem(n_reads=lib.n_reads, n_buckets=b["n_buckets"], bucket_ptr=b["bucket_ptr"],
                                     bucket_reads=b["bucket_reads"], lexrank=ra.lexrank, bc_id=b["bucket_of_read"], up_id=None,
                                     geno_ptr=ra.geno_ptr, geno_mut=ra.geno_mut, geno_q=ra.geno_q, mut_rank=ra.mut_rank, **params)
        ref = orc.cluster(prob, q, r, d)
        n_alignment += int((ref.row_bc == hostdata.NEEDS_ALIGNMENT).any())
        for full in (False, True):
            out = pipeline.cluster_reads(ctx, ra, full_table=full, **params)
            try:
                _rows_equal(hostdata.rows_from_merge_outputs(out), ref)
            except AssertionError as e:
                raise AssertionError(f"case {case} full={full} {kw} {params}: {e}") from None
    assert n_alignment < 60


def test_big_buckets_with_collisions(ctx):
    """A few clones with a hundred or more reads each, half of them sharing barcodes with disjoint genotypes: buckets
    of several hundred reads that do NOT collapse into one seed cluster (the quadratic case of the seed step).
    (Sized for the oracle: its literal restatement is quadratic in the product of two adjacent bucket sizes.)"""
    lib = synth.make_library(n_clones=24, n_reads=3000, p_err=0.004, p_near=0.2, p_collide=0.5, skew=0.5, seed=77)
    for shape in ("thread", "warp", "auto", "huge"):
        os.environ["PCB_REPLAY"] = shape
        try:
            _fused_vs_oracle(ctx, lib, max_diff=2)
        finally:
            os.environ.pop("PCB_REPLAY", None)
    assert lib.n_reads / ctx.stat("buckets") > 7


def test_mixed_light_and_heavy_components(ctx):
    """Mean bucket size below the warp-kernel threshold, but a few clones with hundreds of reads: the heavy
    components (>= 64 reads, numbered first) go to the warp kernel, the rest to the thread kernels."""
    lib = synth.make_library(n_clones=1500, n_reads=12000, p_err=0.01, p_near=0.1, skew=2.2, seed=5)
    _fused_vs_oracle(ctx, lib, max_diff=2)
    assert lib.n_reads / ctx.stat("buckets") < 7


# ----------------------------------------------------------------------------- f1: collision tags + soft filter
@pytest.mark.parametrize("name", ["rand_collide", "rand_dense_d2", "rand_d3_combo", "kat1", "kat5a"])
def test_tag_rows_on_cluster_tables(ctx, name):
    from oracle import f1
    from pacybara_b200 import tagging
    case = util.load_golden(name)
    table = [(r[1], r[2], r[0], r[3], r[4]) for r in util.golden_table(case)]     # canonical rows: members first
    for ratio in (0.5, 0.666, 0.9):
        assert tagging.tag_table(ctx, table, ratio) == f1.tag_table(table, ratio)
    if name in ("rand_collide", "kat1", "kat5a"):
        assert any(r[5] == "collision" for r in tagging.tag_table(ctx, table))     # the collision rule is exercised


def test_tag_rows_random(ctx):
    from oracle import f1
    from pacybara_b200 import tagging
    rng = np.random.default_rng(3)
    table = [(f"B{rng.integers(0, 3000)}", f"U{rng.integers(0, 1500)}", f"r{i}", int(rng.integers(1, 40)), "=") for i in range(20000)]
    assert tagging.tag_table(ctx, table, 0.666) == f1.tag_table(table, 0.666)


# ----------------------------------------------------------------------------- degenerate shapes through the fused call
def _tiny(ctx, seqs, genos, **params):
    mat, lens = hostdata.strings_to_matrix([s.encode() for s in seqs]) if seqs else (np.zeros((0, 8), np.uint8), np.zeros(0, np.int32))
    R = len(seqs)
    ptr = np.zeros(R + 1, dtype=np.int32)
    mut, q = [], []
    for i, g in enumerate(genos):
        for m, qq in g:
            mut.append(m); q.append(qq)
        ptr[i + 1] = len(mut)
    out = ctx.cluster_reads(mat, None, lens, np.arange(R, dtype=np.int32), None, ptr, np.asarray(mut, dtype=np.int32),
                            np.asarray(q, dtype=np.int32), np.arange(8, dtype=np.int32), **params)
    return out


def test_degenerate_inputs(ctx):
    out = _tiny(ctx, [], [])                                   # no reads at all
    assert out["n_buckets"] == 0 and out["n_edges"] == 0 and len(out["read_root"]) == 0
    out = _tiny(ctx, ["ACGTACGT"], [[(0, 50)]])                 # one read: a singleton row with its high-quality call
    assert out["n_buckets"] == 1 and out["read_root"].tolist() == [0] and out["row_size"].tolist() == [1]
    assert out["row_call_n"].tolist() == [1] and out["call_mut"][out["row_call_off"][0]] == 0
    out = _tiny(ctx, ["ACGTACGT"] * 3, [[], [], []])            # one bucket, wild type: one cluster [r1, r0, r2]
    rows = hostdata.rows_from_merge_outputs(out)
    assert rows.n_clusters == 1 and rows.row_members.tolist() == [1, 0, 2]
    out = _tiny(ctx, ["ACGTACGT", "ACGTACGA"], [[], []])        # two singletons one edit apart never merge (1 v 1)
    assert out["n_edges"] == 1 and sorted(out["row_size"].tolist()) == [1, 1]
    out = _tiny(ctx, ["ACGTACGT", "ACGTACGT", "ACGTACGA"], [[], [], []], max_diff=0)   # maxDiff 0: no pair search at all
    assert out["n_edges"] == 0 and hostdata.rows_from_merge_outputs(out).n_clusters == 1
