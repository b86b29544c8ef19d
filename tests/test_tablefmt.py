"""The writer without a Python loop over the reads (pacybara_b200/tablefmt.py): the items it lists, concatenated, must be the
text hostdata.rows_to_table + write_clusters_csv produce -- on the golden tables of the reference and on random rows."""
import io

import numpy as np
import pytest

from pacybara_b200 import hostdata, tablefmt
from tests import util


def _slices_of(strings):
    raw = "\x00".join(strings).encode("ascii")           # (a separator byte between the strings: slices must not run over)
    lens = np.array([len(s) for s in strings], dtype=np.int32)
    off = np.zeros(len(strings), dtype=np.int64)
    if len(strings) > 1:
        np.cumsum(lens[:-1].astype(np.int64) + 1, out=off[1:])
    return tablefmt.Slices(np.frombuffer(raw, dtype=np.uint8), off, lens)


def _text_of(table):
    buf = io.StringIO()
    for vbc, ubc, reads, size, geno in table:
        buf.write(f"{vbc},{ubc},{reads},{size},{geno}\n")
    return buf.getvalue().encode("ascii")


@pytest.mark.parametrize("name", ["kat1", "kat7a", "rand_collide", "rand_dense_d2", "x_duprows", "x_opaque", "x_textquirks", "x_dense_sw"])
def test_items_give_the_reference_rows(name):
    from oracle import orc
    case = util.load_golden(name)
    job = util.job_from_golden(case)
    rows = orc.cluster(job.problem, job.ed_q, job.ed_r, job.ed_d)
    if (rows.row_bc == hostdata.NEEDS_ALIGNMENT).any() or (rows.row_up == hostdata.NEEDS_ALIGNMENT).any():
        pytest.skip("needs the alignment fallback")
    want = _text_of(hostdata.rows_to_table(rows, job.ids, job.bc_names, job.up_names, job.mut_names))
    items = tablefmt.table_items(rows, _slices_of(job.ids), _slices_of(job.bc_names),
                                 None if job.up_names is None else _slices_of(job.up_names), _slices_of(job.mut_names))
    assert tablefmt.concat_host(*items) == want


def test_random_rows_with_missing_uptags_and_no_calls():
    rng = np.random.default_rng(5)
    R, U, M = 500, 60, 30
    ids = [f"m{rng.integers(0, 10**6)}/{i}/ccs" for i in range(R)]
    bcs = ["".join(rng.choice(list("ACGT"), size=rng.integers(3, 30))) for _ in range(U)]
    ups = ["".join(rng.choice(list("ACGT"), size=12)) for _ in range(40)]
    muts = [f"c.{rng.integers(1, 900)}A>T" for _ in range(M)]
    members = rng.permutation(R).astype(np.int32)
    cuts = np.sort(rng.choice(np.arange(1, R), size=119, replace=False))
    row_ptr = np.concatenate([[0], cuts, [R]]).astype(np.int32)
    n = len(row_ptr) - 1
    ncall = rng.integers(0, 4, size=n)
    call_ptr = np.concatenate([[0], np.cumsum(ncall)]).astype(np.int32)
    rows = hostdata.ClusterRows(row_ptr, members, rng.integers(0, U, size=n).astype(np.int32),
                                rng.integers(-1, 40, size=n).astype(np.int32), call_ptr,
                                rng.integers(0, M, size=int(call_ptr[-1])).astype(np.int32), n_clusters=n)
    want = _text_of(hostdata.rows_to_table(rows, ids, bcs, ups, muts))
    got = tablefmt.concat_host(*tablefmt.table_items(rows, _slices_of(ids), _slices_of(bcs), _slices_of(ups), _slices_of(muts)))
    assert got == want
    # no uptag table: the uptag column names barcodes
    rows.row_up[:] = rows.row_bc
    want = _text_of(hostdata.rows_to_table(rows, ids, bcs, None, muts))
    assert tablefmt.concat_host(*tablefmt.table_items(rows, _slices_of(ids), _slices_of(bcs), None, _slices_of(muts))) == want


def test_rows_resolved_by_an_alignment_take_their_strings():
    """A tied barcode vote (pacybara_cluster.py:484-511): the caller resolves the row and hands the string over."""
    ids, bcs, muts = ["a", "b", "c", "d"], ["ACGT", "AGGT"], ["c.1A>T"]
    rows = hostdata.ClusterRows(np.array([0, 3, 4], dtype=np.int32), np.array([0, 1, 2, 3], dtype=np.int32),
                                np.array([hostdata.NEEDS_ALIGNMENT, 1], dtype=np.int32), np.array([hostdata.NEEDS_ALIGNMENT, 0], dtype=np.int32),
                                np.array([0, 1, 1], dtype=np.int32), np.array([0], dtype=np.int32), n_clusters=1)
    got = tablefmt.concat_host(*tablefmt.table_items(rows, _slices_of(ids), _slices_of(bcs), None, _slices_of(muts),
                                                     bc_text={0: "ACNT"}, up_text={0: "AC-T"}))
    assert got == b"ACNT,AC-T,a|b|c,3,c.1A>T\nAGGT,ACGT,d,1,=\n"
