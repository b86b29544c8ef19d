"""f3: pcb_match_barcodes (barcode -> library matching with maxErr) against the exhaustive-DP oracle."""
import numpy as np
import pytest

from pacybara_b200 import hostdata

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from pacybara_b200.api import Context
    c = Context(0)
    yield c
    c.close()


def rand_bc(rng, n):
    return "".join(rng.choice(list("ACGT"), n))


def mutate(rng, s, n_edits):
    s = list(s)
    for _ in range(n_edits):
        op = rng.integers(3)
        p = int(rng.integers(len(s))) if s else 0
        if op == 0 and s:
            s[p] = "ACGT"[rng.integers(4)]
        elif op == 1 and len(s) > 1:
            del s[p]
        else:
            s.insert(p, "ACGT"[rng.integers(4)])
    return "".join(s)


def make_case(seed, n_lib, n_q, bc_len, len_jitter=0):
    rng = np.random.default_rng(seed)
    lib = [rand_bc(rng, bc_len + int(rng.integers(-len_jitter, len_jitter + 1))) for _ in range(n_lib)]
    for i in range(0, n_lib, 17):                      # near-duplicates and exact duplicates inside the library
        lib[i] = mutate(rng, lib[(i * 7) % n_lib], int(rng.integers(0, 3)))
    qs = []
    for _ in range(n_q):
        r = rng.random()
        if r < 0.15:
            qs.append(rand_bc(rng, bc_len))
        else:
            qs.append(mutate(rng, lib[int(rng.integers(n_lib))], int(rng.integers(0, 5))))
    qs = [q[:64] for q in qs]
    return lib, qs


@pytest.mark.parametrize("bc_len,jitter,k", [(25, 0, 0), (25, 0, 1), (25, 0, 2), (25, 2, 2), (25, 2, 3), (18, 3, 2), (50, 2, 3), (40, 8, 2),
                                              (6, 2, 2)])
def test_match_vs_oracle(ctx, bc_len, jitter, k):
    from oracle import orc
    lib, qs = make_case(100 + bc_len + k, 1500, 4000, bc_len, jitter)
    lm, ll = hostdata.strings_to_matrix([s.encode() for s in lib])
    qm, ql = hostdata.strings_to_matrix([s.encode() for s in qs])
    got = ctx.match_barcodes(lm, ll, qm, ql, k, want_pairs=True)
    ref = orc.match_barcodes(lm, ll, qm, ql, k)
    for f in ("best_d", "n_best", "first_hit"):
        assert np.array_equal(got[f], ref[f]), f
    assert (got["best_d"] >= 0).sum() > 500
    # every pair within k, each once, with the DP's distance
    key = got["pair_q"].astype(np.int64) * len(lib) + got["pair_row"]
    assert (np.diff(key) > 0).all()
    rng = np.random.default_rng(1)
    for x in rng.integers(0, got["n_pairs"], 300):
        a, b = qs[got["pair_q"][x]], lib[got["pair_row"][x]]
        assert got["pair_d"][x] == orc.lev(a.encode(), b.encode())
    n_ref = sum(int(orc.match_barcodes(lm, ll, qm[i:i + 1], ql[i:i + 1], k)["n_best"][0]) for i in range(50))
    assert n_ref == int(got["n_best"][:50].sum())


def test_match_device_memory_and_degenerate(ctx):
    import torch
    from oracle import orc
    lib, qs = make_case(9, 300, 500, 25, 1)
    lm, ll = hostdata.strings_to_matrix([s.encode() for s in lib])
    qm, ql = hostdata.strings_to_matrix([s.encode() for s in qs])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    got = ctx.match_barcodes(t(lm), t(ll), t(qm), t(ql), 2)
    ref = orc.match_barcodes(lm, ll, qm, ql, 2)
    for f in ("best_d", "n_best", "first_hit"):
        assert np.array_equal(got[f].cpu().numpy(), ref[f]), f
    none = ctx.match_barcodes(lm[:0], ll[:0], qm, ql, 2)
    assert (none["best_d"] == -1).all() and (none["n_best"] == 0).all() and none["n_pairs"] == 0
    empty = ctx.match_barcodes(lm, ll, qm[:0], ql[:0], 2)
    assert empty["best_d"].shape[0] == 0
    one = ctx.match_barcodes(lm[:1], ll[:1], lm[:1], ll[:1], 2)
    assert one["best_d"].tolist() == [0] and one["n_best"].tolist() == [1] and one["first_hit"].tolist() == [0]


def test_bc_matcher_mirrors_the_call_site(ctx):
    from pacybara_b200.barseq_match import BcMatcher
    lib = ["ACGTACGTACGTACGTACGTACGTA", "ACGTACGTACGTACGTACGTACGTC", None, "TTTTTTTTTTTTTTTTTTTTTTTTT", "ACGTACGTACGTACGTACGTACGTA"]
    m = BcMatcher(lib, err_cutoff=2, ctx=ctx)
    reads = ["ACGTACGTACGTACGTACGTACGTA",       # exact: rows 1 and 5
             "ACGTACGTACGTACGTACGTACGTG",       # one substitution away from rows 1, 2, 5
             None,                              # failed extraction
             "ACGTNCGTACGTACGTACGTACGTA",       # N costs one edit: rows 1 and 5 at distance 1 (row 2 would be 2)
             "GGGGGGGGGGGGGGGGGGGGGGGGG",       # nothing within 2
             "TTTTTTTTTTTTTTTTTTTTTTTT",        # one deletion away from row 4
             "ACGTACGTACGTACGTACGTACGTA"]       # duplicate read
    r = m.find_matches(reads)
    assert r["hits"] == [[1, 5], [1, 2, 5], [], [1, 5], [], [4], [1, 5]]
    assert r["diffs"] == [0, 1, None, 1, None, 1, 0]
    assert r["nhits"] == [2, 3, 0, 2, 0, 1, 2]
    # paired reads: the closer read decides, equal distances pool their rows
    r2 = m.find_matches(reads, ["TTTTTTTTTTTTTTTTTTTTTTTTT", None, "ACGTACGTACGTACGTACGTACGTC", None, None, "TTTTTTTTTTTTTTTTTTTTTTTTT", None])
    assert r2["hits"] == [[1, 4, 5], [1, 2, 5], [2], [1, 5], [], [4], [1, 5]]
    assert r2["diffs"] == [0, 1, 0, 1, None, 0, 0]


@pytest.mark.parametrize("bc_len,k", [(25, 2), (50, 3), (12, 1)])
def test_match_queries_with_n(ctx, bc_len, k):
    """Query characters outside ACGT equal nothing: same answer as the DP over the raw strings."""
    from oracle import orc
    lib, qs = make_case(300 + bc_len, 800, 3000, bc_len, 1)
    rng = np.random.default_rng(2)
    qs = ["".join(("N" if rng.random() < 0.04 else c) for c in q) for q in qs]
    qs[0] = "N" * bc_len
    assert sum("N" in q for q in qs) > 1000
    lm, ll = hostdata.strings_to_matrix([s.encode() for s in lib])
    qm, ql = hostdata.strings_to_matrix([s.encode() for s in qs])
    got = ctx.match_barcodes(lm, ll, qm, ql, k)
    ref = orc.match_barcodes(lm, ll, qm, ql, k)
    for f in ("best_d", "n_best", "first_hit"):
        assert np.array_equal(got[f], ref[f]), f
    # a library row with a non-base character is an opaque string: it stays in the library but matches nothing
    self_match = ctx.match_barcodes(qm, ql, qm, ql, k)
    for i, s in enumerate(qs):
        if "N" not in s:
            assert self_match["best_d"][i] == 0 and "N" not in qs[int(self_match["first_hit"][i])]
