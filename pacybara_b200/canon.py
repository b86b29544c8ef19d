"""Canonical form of a cluster table (what "bit-exact" is defined on).

The reference's own output is not byte-reproducible: singleton rows come out in
``set`` order (src/pacybara_cluster.py:547,560) and same-position mutations are
ordered by ``set`` iteration (:338-347,475), both PYTHONHASHSEED dependent.  Two
tables are equal when, after this canonicalisation, they are the same set of rows:

    key   = the set of read ids of the row
    value = (virtualBarcode, upBarcode, size, mutations sorted by (digit key, string))
"""
from __future__ import annotations

import gzip
import hashlib
import re
from typing import Iterable, List, Sequence, Tuple

_DIGITS = re.compile(r"\D+")


def mut_sort_key(m: str) -> Tuple[int, str]:
    """(int of all digits, string): the reference's sort key (:475) plus a tie-break."""
    d = _DIGITS.sub("", m)
    return (int(d) if d else -1, m)


def canon_geno(geno: str) -> str:
    if geno == "=" or geno == "":
        return "="
    return ";".join(sorted(geno.split(";"), key=mut_sort_key))


def canon_row(vbc: str, ubc: str, reads: str, size, geno: str) -> Tuple[str, str, str, int, str]:
    members = "|".join(sorted(reads.split("|")))
    return (members, vbc, ubc, int(size), canon_geno(geno))


def canon_table(rows: Iterable[Sequence]) -> List[Tuple[str, str, str, int, str]]:
    return sorted(canon_row(*r) for r in rows)


def read_clusters_csv(path: str) -> List[List[str]]:
    """Rows of a clusters.csv.gz (header virtualBarcode,upBarcode,reads,size,geno; :551)."""
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as fh:
        header = fh.readline().rstrip("\n")
        if header != "virtualBarcode,upBarcode,reads,size,geno":
            raise ValueError(f"unexpected clusters header: {header!r}")
        return [line.rstrip("\n").split(",") for line in fh if line.strip()]


def table_digest(table: List[Tuple]) -> str:
    h = hashlib.sha256()
    for row in table:
        h.update(("\t".join(str(x) for x in row) + "\n").encode())
    return h.hexdigest()


def diff_tables(a: List[Tuple], b: List[Tuple], limit: int = 10) -> str:
    sa, sb = set(a), set(b)
    only_a = sorted(sa - sb)[:limit]
    only_b = sorted(sb - sa)[:limit]
    return (f"rows {len(a)} vs {len(b)}; only-left {len(sa - sb)}, only-right {len(sb - sa)}\n"
            + "\n".join(f"  < {r}" for r in only_a) + "\n" + "\n".join(f"  > {r}" for r in only_b))
