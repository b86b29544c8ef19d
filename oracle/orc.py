"""ctypes front end of oracle/pcb_oracle.c (test infrastructure, never a product path)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional, Tuple

import numpy as np

from pacybara_b200.hostdata import ClusterRows, MergeProblem

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpcb_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pcb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_edit_pairs_brute.restype = C.c_int64
        _lib.orc_edit_pairs_seeded.restype = C.c_int64
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_threads(n: int) -> None:
    lib().orc_set_threads(C.c_int32(int(n)))


def precluster(seq: np.ndarray, qual: np.ndarray, length: np.ndarray):
    seq, qual, length = _c(seq, np.uint8), _c(qual, np.uint8), _c(length, np.int32)
    R, stride = seq.shape
    bucket = np.empty(R, dtype=np.int32)
    first = np.empty(max(R, 1), dtype=np.int32)
    qsum = np.zeros((max(R, 1), stride), dtype=np.uint8)
    U = lib().orc_precluster(_p(seq), _p(qual), _p(length), C.c_int32(R), C.c_int32(stride),
                             _p(bucket), _p(first), _p(qsum))
    return dict(n_buckets=int(U), bucket_of_read=bucket, first_read=first[:U].copy(), qsum=qsum[:U].copy())


def lev(a: bytes, b: bytes) -> int:
    return int(lib().orc_lev(a, C.c_int32(len(a)), b, C.c_int32(len(b))))


def lev_banded(a: bytes, b: bytes, k: int) -> int:
    return int(lib().orc_lev_banded(a, C.c_int32(len(a)), b, C.c_int32(len(b)), C.c_int32(k)))


def edit_pairs(seq: np.ndarray, length: np.ndarray, k: int, method: str = "seeded"
               ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """All i<j with Levenshtein <= k, sorted by (i, j)."""
    seq, length = _c(seq, np.uint8), _c(length, np.int32)
    U, stride = seq.shape
    fn = lib().orc_edit_pairs_brute if method == "brute" else lib().orc_edit_pairs_seeded
    cap = max(1024, 8 * U)
    while True:
        oi, oj, od = (np.empty(cap, dtype=np.int32) for _ in range(3))
        n = int(fn(_p(seq), _p(length), C.c_int32(U), C.c_int32(stride), C.c_int32(k),
                   _p(oi), _p(oj), _p(od), C.c_int64(cap)))
        if n <= cap:
            return oi[:n].copy(), oj[:n].copy(), od[:n].copy()
        cap = n


def match_barcodes(lib_seq: np.ndarray, lib_len: np.ndarray, q_seq: np.ndarray, q_len: np.ndarray, k: int) -> Dict[str, np.ndarray]:
    """f3: per query the smallest Levenshtein distance <= k to a library row (-1: none), the number of rows at that
    distance and the first of them; exhaustive DP."""
    lib_seq, lib_len, q_seq, q_len = _c(lib_seq, np.uint8), _c(lib_len, np.int32), _c(q_seq, np.uint8), _c(q_len, np.int32)
    L, Q = lib_len.shape[0], q_len.shape[0]
    best, nb, first = (np.empty(max(Q, 1), dtype=np.int32) for _ in range(3))
    lib().orc_match_brute(_p(lib_seq), _p(lib_len), C.c_int32(L), C.c_int32(lib_seq.shape[1] if L else 1),
                          _p(q_seq), _p(q_len), C.c_int32(Q), C.c_int32(q_seq.shape[1] if Q else 1), C.c_int32(k),
                          _p(best), _p(nb), _p(first))
    return dict(best_d=best[:Q], n_best=nb[:Q], first_hit=first[:Q])


def cluster(p: MergeProblem, ed_q: np.ndarray, ed_r: np.ndarray, ed_d: np.ndarray) -> ClusterRows:
    """Raw ED rows (file order, both directions) -> cluster table, by the literal restatement."""
    R = p.n_reads
    nnz = int(p.geno_ptr[-1])
    row_ptr = np.zeros(R + 2, dtype=np.int32)
    row_members = np.zeros(max(R, 1), dtype=np.int32)
    row_bc = np.zeros(max(R, 1), dtype=np.int32)
    row_up = np.zeros(max(R, 1), dtype=np.int32)
    call_ptr = np.zeros(R + 2, dtype=np.int32)
    call_mut = np.zeros(max(nnz, 1), dtype=np.int32)
    stats = np.zeros(4, dtype=np.int64)
    ed_q, ed_r, ed_d = _c(ed_q, np.int32), _c(ed_r, np.int32), _c(ed_d, np.int32)
    a = dict(bucket_ptr=_c(p.bucket_ptr, np.int32), bucket_reads=_c(p.bucket_reads, np.int32),
             lexrank=_c(p.lexrank, np.int32), bc_id=_c(p.bc_id, np.int32),
             up_id=None if p.up_id is None else _c(p.up_id, np.int32),
             geno_ptr=_c(p.geno_ptr, np.int32), geno_mut=_c(p.geno_mut, np.int32),
             geno_q=_c(p.geno_q, np.int32), mut_rank=_c(p.mut_rank, np.int32))
    n = lib().orc_cluster(C.c_int32(R), C.c_int32(p.n_buckets), _p(a["bucket_ptr"]), _p(a["bucket_reads"]),
                          _p(a["lexrank"]), _p(a["bc_id"]), _p(a["up_id"]),
                          _p(a["geno_ptr"]), _p(a["geno_mut"]), _p(a["geno_q"]), _p(a["mut_rank"]),
                          C.c_int64(ed_q.shape[0]), _p(ed_q), _p(ed_r), _p(ed_d),
                          C.c_int32(p.max_diff), C.c_int32(p.min_qual), C.c_double(p.min_jaccard),
                          C.c_int32(p.min_matches),
                          _p(row_ptr), _p(row_members), _p(row_bc), _p(row_up), _p(call_ptr), _p(call_mut),
                          _p(stats))
    if n < 0:
        raise KeyError(f"oracle cluster failed with code {n} (a clustered read has no uptag record)")
    return ClusterRows(row_ptr[: n + 1].copy(), row_members[: int(row_ptr[n])].copy(), row_bc[:n].copy(),
                       row_up[:n].copy(), call_ptr[: n + 1].copy(), call_mut[: int(call_ptr[n])].copy(),
                       n_clusters=int(stats[0]), stats=dict(links=int(stats[1]), visits=int(stats[2])))
