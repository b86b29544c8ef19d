"""Host-side helpers around the tokenizers and the writer (no GPU): gzip members, file reading, string slicing,
table building from rows."""
import gzip
import io
import os

import numpy as np

from pacybara_b200 import hostdata, ingest


def test_gzip_members_concatenate_to_one_file(tmp_path):
    data = b"".join(b"row %d,ACGT\n" % i for i in range(200000))
    members = hostdata.gzip_members(data, chunk=1 << 18)
    assert len(members) > 4
    blob = b"".join(members)
    assert gzip.decompress(blob) == data                       # RFC 1952: members back to back are one file
    p = tmp_path / "x.csv.gz"
    p.write_bytes(blob)
    with gzip.open(p, "rt") as fh:
        assert sum(1 for _ in fh) == 200000
    assert gzip.decompress(b"".join(hostdata.gzip_members(b""))) == b""


def test_write_clusters_csv_format(tmp_path):
    table = [("ACGT", "ACGT", "r1|r2", 2, "5A>C;7del"), ("TTTT", "None", "r3", 1, "=")]
    p = str(tmp_path / "clusters.csv.gz")
    hostdata.write_clusters_csv(p, table)
    with gzip.open(p, "rt") as fh:
        assert fh.read() == "virtualBarcode,upBarcode,reads,size,geno\nACGT,ACGT,r1|r2,2,5A>C;7del\nTTTT,None,r3,1,=\n"


def test_read_all_plain_and_gzip(tmp_path):
    a, b = tmp_path / "a.fastq.gz", tmp_path / "b.csv"
    with gzip.open(a, "wb") as fh:
        fh.write(b"@r\nA\n+\nI\n")
    b.write_bytes(b'"r","="\n')
    got = ingest.read_all([str(a), str(b), None])
    assert got == [b"@r\nA\n+\nI\n", b'"r","="\n', None]


def test_slices_and_decode_cache():
    raw = b"@read/1 x\nACGT\n+\nIIII\n@read/22\nAC\n+\nII\n"
    off, ln = np.array([1, 23], dtype=np.int64), np.array([6, 7], dtype=np.int32)
    assert ingest._slices(raw, off, ln) == ["read/1", "read/22"]
    assert ingest._text(raw) is ingest._text(raw)               # decoded once per buffer
    ingest._decoded.clear()


def test_rows_to_table_matches_a_plain_loop():
    rng = np.random.default_rng(4)
    R, n_rows = 500, 120
    cuts = np.sort(rng.choice(np.arange(1, R), n_rows - 1, replace=False))
    row_ptr = np.concatenate([[0], cuts, [R]]).astype(np.int32)
    ncall = rng.integers(0, 4, n_rows)
    call_ptr = np.concatenate([[0], np.cumsum(ncall)]).astype(np.int32)
    rows = hostdata.ClusterRows(row_ptr, rng.permutation(R).astype(np.int32), rng.integers(0, 30, n_rows).astype(np.int32),
                                rng.integers(-1, 30, n_rows).astype(np.int32), call_ptr,
                                rng.integers(0, 50, int(call_ptr[-1])).astype(np.int32), n_clusters=n_rows)
    ids = [f"m/{i}/ccs" for i in range(R)]
    bc = [f"BC{i}" for i in range(30)]
    up = [f"UP{i}" for i in range(30)]
    mn = [f"{i}A>C" for i in range(50)]
    got = hostdata.rows_to_table(rows, ids, bc, up, mn)
    for i, (vbc, ubc, reads, size, geno) in enumerate(got):
        mem = rows.row_members[row_ptr[i]:row_ptr[i + 1]]
        calls = rows.call_mut[call_ptr[i]:call_ptr[i + 1]]
        assert vbc == bc[rows.row_bc[i]]
        assert ubc == ("None" if rows.row_up[i] < 0 else up[rows.row_up[i]])
        assert reads == "|".join(ids[r] for r in mem) and size == len(mem)
        assert geno == (";".join(mn[m] for m in calls) if len(calls) else "=")


def test_lazy_ids_find_a_byte_without_building_the_strings():
    """ingest.LazyIds.any_contains: the '-' check of the one-call program on the raw FASTQ bytes (quality strings are full of
    '-', only the id slices count)."""
    import numpy as np
    from pacybara_b200 import ingest
    raw = b"@r1 x\nACGT\n+\n--I-\n@q-2 y\nAC\n+\n-I\n@r3\nA\n+\n-\n"
    off = np.array([1, 19, 34], dtype=np.int64)
    length = np.array([2, 3, 2], dtype=np.int32)
    ids = ingest.LazyIds(raw, off, length)
    assert list(ids) == ["r1", "q-2", "r3"] and len(ids) == 3 and ids == ["r1", "q-2", "r3"]
    assert ids.any_contains(b"-") == "q-2"
    assert ingest.LazyIds(raw, off[[0, 2]], length[[0, 2]]).any_contains(b"-") is None
    assert ingest.LazyIds(b"", off[:0], length[:0]).any_contains(b"-") is None
