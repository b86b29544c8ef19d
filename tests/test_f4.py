"""Row f4: library merge (src/pacybara_merge.R:48-116).  The oracle restatement on hand-made cases (CPU), the CUDA path
against it (GPU).  Parity unpinned against R (not available here): both follow the R source line by line."""
import numpy as np
import pytest

from oracle import f4


def row(vbc, geno, size, reads, up=None):
    return dict(virtualBarcode=vbc, upBarcode=up or vbc[:4], reads=reads, size=size, hgvsc=geno)


LIB1 = [row("AAAA", "c.1A>G", 5, "a1|a2"), row("CCCC", "c.2C>T", 3, "c1"), row("CCCC", "c.9G>A", 9, "c2", up="CCCX"),
        row("GGGG", "=", 4, "g1"), row("GGGG", "=", 4, "g2"), row("GGGG", "=", 7, "g3")]
LIB2 = [row("AAAA", "c.1A>G", 2, "x1"),        # merge into row 0
        row("TTTT", "c.5T>C", 6, "x2"),        # new
        row("CCCC", "c.7A>T", 1, "x3"),        # conflict
        row("GGGG", "=", 5, "x4"),             # three candidates: the largest (row 5, 7 -> 12)
        row("GGGG", "=", 1, "x5"),             # again row 5 (12 -> 13)
        row("CCCC", "c.9G>A", 2, "x6"),        # merge into row 2
        row("AAAA", "c.1A>G", 1, "x7")]        # row 0 again: read lists keep the order of the second library


def test_oracle_on_the_hand_made_case():
    out = f4.merge_tables(LIB1, LIB2)
    assert [(r["virtualBarcode"], r["size"], r["reads"]) for r in out] == [
        ("AAAA", 8, "a1|a2|x1|x7"), ("CCCC", 3, "c1"), ("CCCC", 11, "c2|x6"), ("GGGG", 4, "g1"), ("GGGG", 4, "g2"),
        ("GGGG", 13, "g3|x4|x5"), ("TTTT", 6, "x2"), ("CCCC", 1, "x3")]
    assert [r["collision"] for r in out] == ["", "collision", "collision", "collision", "collision", "collision", "", "collision"]
    # equal sizes: which.max takes the first
    out = f4.merge_tables([row("GGGG", "=", 4, "g1"), row("GGGG", "=", 4, "g2")], [row("GGGG", "=", 1, "y1"), row("GGGG", "=", 1, "y2")])
    assert [(r["size"], r["reads"]) for r in out] == [(6, "g1|y1|y2"), (4, "g2")]


def random_tables(seed, n1, n2, n_bc, n_geno):
    rng = np.random.default_rng(seed)
    def lib(n, tag):
        return [row(f"B{rng.integers(0, n_bc):05d}", f"g{rng.integers(0, n_geno)}", int(rng.integers(1, 30)), f"{tag}{i}",
                    up=f"U{rng.integers(0, n_bc // 2 + 1)}") for i in range(n)]
    return lib(n1, "p"), lib(n2, "q")


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n1,n2,n_bc,n_geno", [(1, 0, 5, 3, 2), (2, 5, 0, 3, 2), (3, 400, 300, 150, 3), (4, 20000, 15000, 9000, 4),
                                                    (5, 3000, 3000, 40, 2)])
def test_cuda_library_merge_against_the_oracle(seed, n1, n2, n_bc, n_geno):
    from pacybara_b200 import libmerge
    from pacybara_b200.api import Context
    ctx = Context(0)
    try:
        for a, b in ((LIB1, LIB2),) + (random_tables(seed, n1, n2, n_bc, n_geno),):
            assert libmerge.merge_tables(ctx, a, b) == f4.merge_tables(a, b)
    finally:
        ctx.close()
