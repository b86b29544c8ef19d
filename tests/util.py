"""Shared helpers for the parity tests (golden fixtures -> jobs -> canonical tables)."""
from __future__ import annotations

import glob
import gzip
import io
import json
import os
from typing import Dict, List

import numpy as np

from pacybara_b200 import canon, hostdata

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefixes=("kat", "x_", "q1", "rand_")) -> List[str]:
    names = sorted(os.path.basename(p)[:-8] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.json.gz")))
    return [n for n in names if n.startswith(tuple(prefixes))]


def load_golden(name: str) -> Dict:
    with gzip.open(os.path.join(GOLDEN_DIR, name + ".json.gz"), "rt") as fh:
        return json.load(fh)


def ed_text(case: Dict) -> str:
    return hostdata.ED_HEADER + "\n" + "".join(f"\"{a}\",\"{b}\",NA,NA,{d}\n" for a, b, d in case["ed_rows"])


def job_from_golden(case: Dict) -> hostdata.ClusterJob:
    p = case["params"]
    up = None if case["uptag"] is None else io.StringIO(case["uptag"])
    return hostdata.load_cluster_job(io.StringIO(ed_text(case)), io.StringIO(case["geno"]),
                                     io.StringIO(case["preclust"]), up, max_diff=p["max_diff"],
                                     min_qual=p["min_qual"], min_jaccard=p["min_jaccard"],
                                     min_matches=p["min_matches"])


def golden_table(case: Dict):
    return [tuple(r) for r in case["table"]]


def table_of(job: hostdata.ClusterJob, rows: hostdata.ClusterRows):
    return canon.canon_table(hostdata.rows_to_table(rows, job.ids, job.bc_names, job.up_names, job.mut_names))


def preclust_text(reads: hostdata.FastqReads, res: Dict) -> str:
    """Render a bucketing result (dict with bucket_ptr/bucket_reads/bucket_seq/bucket_len/bucket_qsum)."""
    out = io.StringIO()
    hostdata.write_preclust_fastq(out, reads.ids, res["bucket_ptr"], res["bucket_reads"], res["bucket_seq"],
                                  res["bucket_len"], res["bucket_qsum"])
    return out.getvalue()


def oracle_buckets(reads: hostdata.FastqReads) -> Dict:
    """Bucketing through the C oracle, in the layout the CUDA path returns."""
    from oracle import orc
    r = orc.precluster(reads.seq, reads.qual, reads.length)
    U = r["n_buckets"]
    order = np.argsort(r["bucket_of_read"], kind="stable").astype(np.int32)
    counts = np.bincount(r["bucket_of_read"], minlength=U)
    ptr = np.zeros(U + 1, dtype=np.int32)
    np.cumsum(counts, out=ptr[1:])
    return dict(n_buckets=U, bucket_of_read=r["bucket_of_read"], bucket_ptr=ptr, bucket_reads=order,
                bucket_seq=reads.seq[r["first_read"]], bucket_len=reads.length[r["first_read"]],
                bucket_qsum=r["qsum"])


# ---------------------------------------------------------------------------------------------
# host emulation of the per-component replay (tests/host_emul/merge_emul.cpp): unit-test only

_EMUL = None


def emul_lib():
    global _EMUL
    if _EMUL is None:
        import ctypes as C
        import subprocess
        here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul")
        so = os.path.join(here, "libmerge_emul.so")
        src = os.path.join(here, "merge_emul.cpp")
        hdr = os.path.join(os.path.dirname(here), "..", "pacybara_b200", "csrc", "merge_core.h")
        csrc = os.path.dirname(hdr)
        deps = [src, hdr, os.path.join(here, "simt_emul.cpp"), os.path.join(here, "simt_emul.h")] + \
               [os.path.join(csrc, f) for f in ("merge_small.cuh", "merge_warp.cuh", "merge_huge.cuh", "simt.h", "myers.cuh")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.run([cxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", csrc, "-I", here, "-o", so, src,
                            os.path.join(here, "simt_emul.cpp")], check=True)
        _EMUL = C.CDLL(so)
    return _EMUL


def pack_planes(seq: np.ndarray, length: np.ndarray) -> np.ndarray:
    """uint64 [U, 2] bit planes (L, H) of ASCII barcodes -- the layout csrc/myers.cuh works on."""
    code = np.zeros(seq.shape, dtype=np.uint64)
    for ch, v in ((ord("C"), 1), (ord("G"), 2), (ord("T"), 3)):
        code[seq == ch] = v
    cols = np.arange(seq.shape[1], dtype=np.uint64)[None, :]
    valid = cols < length[:, None].astype(np.uint64)
    L = np.where(valid, (code & np.uint64(1)) << cols, np.uint64(0)).sum(axis=1, dtype=np.uint64)
    H = np.where(valid, ((code >> np.uint64(1)) & np.uint64(1)) << cols, np.uint64(0)).sum(axis=1, dtype=np.uint64)
    return np.ascontiguousarray(np.stack([L, H], axis=1))


def emul_merge_small(p: hostdata.MergeProblem, pairs: hostdata.PairList, planes=None, blen=None, compute_k=-1, seed=1, shard=(0, 1)):
    """The warp kernel for small components (csrc/merge_small.cuh) on an emulated warp, generic replay for the rest.
    out['n_small'] = components it replayed."""
    out = emul_merge(p, pairs, shard=shard, planes=planes, blen=blen, compute_k=compute_k, use_small=seed)
    out["n_small"] = int(out["status"][1]) if out["status"][0] == 0 else 0
    return out


def emul_merge(p: hostdata.MergeProblem, pairs: hostdata.PairList, shard=(0, 1), planes=None, blen=None,
               compute_k=-1, use_small=0, mode=0, huge_cap=64, huge_window=16) -> Dict[str, np.ndarray]:
    """mode 0: process_component() per component (with use_small: the small-component warp kernel first); 2: the generic
    warp replay (merge_warp.cuh) on an emulated warp; 3: the giant tier (merge_huge.cuh), use_small seeds the shuffles."""
    import ctypes as C
    R = p.n_reads
    nnz = int(p.geno_ptr[-1])
    out = {k: np.zeros(max(R, 1), dtype=dt) for k, dt in hostdata.MERGE_OUTPUTS}
    out["call_mut"] = np.zeros(max(nnz, 1), dtype=np.int32)
    status = np.zeros(4, dtype=np.int32)
    ptr = lambda a: None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
    keep = [np.ascontiguousarray(x, dtype=np.int32) for x in
            (p.bucket_ptr, p.bucket_reads, p.lexrank, p.bc_id, p.geno_ptr, p.geno_mut, p.geno_q, p.mut_rank,
             pairs.q, pairs.r, pairs.d, pairs.ed)]
    up = None if p.up_id is None else np.ascontiguousarray(p.up_id, dtype=np.int32)
    n = emul_lib().emul_merge(C.c_int32(R), C.c_int32(p.n_buckets), ptr(keep[0]), ptr(keep[1]), ptr(keep[2]),
                              ptr(keep[3]), ptr(up), ptr(keep[4]), ptr(keep[5]), ptr(keep[6]), ptr(keep[7]),
                              C.c_int64(len(pairs.q)), ptr(keep[8]), ptr(keep[9]), ptr(keep[10]), ptr(keep[11]),
                              C.c_int32(p.max_diff), C.c_int32(p.min_qual), C.c_double(p.min_jaccard),
                              C.c_int32(p.min_matches), C.c_int32(shard[0]), C.c_int32(shard[1]),
                              *[ptr(out[k]) for k, _ in hostdata.MERGE_OUTPUTS], ptr(out["call_mut"]), ptr(status),
                              ptr(planes), ptr(None if blen is None else np.ascontiguousarray(blen, dtype=np.int32)),
                              C.c_int32(compute_k), C.c_int32(0 if blen is None or int(np.max(blen)) <= 32 else 1),
                              C.c_int32(use_small), C.c_int32(mode), C.c_int32(huge_cap), C.c_int32(huge_window))
    if n < 0:
        raise RuntimeError(f"emul_merge: {n} (the small-component path needs component-order numbering / uniform lanes)")
    out = {k: v[:R] if k != "call_mut" else v[:nnz] for k, v in out.items()}
    out["status"] = status
    out["n_components"] = n
    return out


def oracle_table_for_library(lib, ra, names, max_diff=2, min_qual=32, min_jaccard=0.2, min_matches=1):
    """The whole hot path through the CPU oracle: bucketing, seeded DP distances at k = 2*maxDiff,
    literal cluster restatement.  Rows as (virtualBarcode, upBarcode, reads, size, geno)."""
    from oracle import orc
    reads = hostdata.FastqReads(names["ids"], lib.seq, lib.qual, lib.length)
    b = oracle_buckets(reads)
    ei, ej, ed = orc.edit_pairs(b["bucket_seq"], b["bucket_len"], 2 * max_diff, method="seeded")
    q, r, d = hostdata.canonical_rows(ei, ej, ed)
    prob = hostdata.MergeProblem(n_reads=lib.n_reads, n_buckets=b["n_buckets"], bucket_ptr=b["bucket_ptr"],
                                 bucket_reads=b["bucket_reads"], lexrank=ra.lexrank, bc_id=b["bucket_of_read"],
                                 up_id=ra.up_id, geno_ptr=ra.geno_ptr, geno_mut=ra.geno_mut, geno_q=ra.geno_q,
                                 mut_rank=ra.mut_rank, max_diff=max_diff, min_qual=min_qual,
                                 min_jaccard=min_jaccard, min_matches=min_matches)
    rows = orc.cluster(prob, q, r, d)
    bc_names = hostdata.matrix_to_strings(b["bucket_seq"], b["bucket_len"])
    return hostdata.rows_to_table(rows, names["ids"], bc_names, names["up_names"], names["mut_names"])


# ---------------------------------------------------------------------------------------------
# parity at sizes where the oracle's own pair search is out of reach (BASELINE cfg 2 / 4 / 5 at full size)

def library_arrays(lib, ctx):
    """Array inputs of Context.cluster_reads for a synth.Library without building R id strings
    (padded ids sort like their index)."""
    up_id = None
    if lib.uptag_len is not None:
        b = ctx.bucket(lib.uptag_seq(), None, lib.uptag_len, want_qsum=False)
        up_id = b["bucket_of_read"].astype(np.int32)
    R = lib.n_reads
    return (lib.seq, None, lib.length.astype(np.int32), np.arange(R, dtype=np.int32), up_id, lib.geno_ptr.astype(np.int32),
            lib.geno_mut, lib.geno_q, np.arange(len(lib.mut_codes), dtype=np.int32))


def rows_equal(a, b, what=""):
    for f in ("row_ptr", "row_members", "row_bc", "row_up", "call_ptr", "call_mut"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), (what, f)


def check_full_size_against_oracle_merge(ctx, lib, n_spot=20_000, seed=0):
    """The GPU's full distance table (every pair <= 2*maxDiff) is (i) spot-checked pair by pair against the textbook DP,
    edges and random non-edges alike, (ii) fed to the oracle's literal restatement of pacybara_cluster.py
    (:385-455, :464-477, :515-537, :546-567), whose rows must equal the GPU's rows -- row order, member order, consensus
    barcode, calls -- for BOTH the full-table run and the on-demand-distance run.  Returns a few counts."""
    import ctypes as C
    from oracle import orc
    p = dict(max_diff=lib.params["max_diff"], min_qual=lib.params["min_qual"], min_jaccard=0.2, min_matches=1)
    args = library_arrays(lib, ctx)
    R = lib.n_reads
    out_full = ctx.cluster_reads(*args, full_table=True, **p)
    E = out_full["n_edges"]
    ei, ej, ed = (np.empty(max(E, 1), dtype=np.int32) for _ in range(3))
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    assert ctx.lib.pcb_edit_pairs_fetch(ctx.h, 0, ptr(ei), ptr(ej), ptr(ed), E) == 0
    ei, ej, ed = ei[:E], ej[:E], ed[:E]
    bor = out_full["bucket_of_read"]
    U = out_full["n_buckets"]
    first = np.full(U, R, dtype=np.int64)
    np.minimum.at(first, bor, np.arange(R))
    bseq, blen = lib.seq[first], lib.length[first]
    rng = np.random.default_rng(seed)
    lev = lambda a, b: orc.lev(bseq[a, :blen[a]].tobytes(), bseq[b, :blen[b]].tobytes())
    for k in rng.choice(E, size=min(E, n_spot), replace=False):
        assert lev(ei[k], ej[k]) == ed[k], ("distance", int(ei[k]), int(ej[k]))
    assert (ei < ej).all() and (ed >= 0).all() and (ed <= 2 * p["max_diff"]).all()
    key = ei.astype(np.int64) * U + ej
    assert (np.diff(key) > 0).all()                      # sorted by (i, j), no duplicate
    ra_, rb_ = rng.integers(0, U, n_spot), rng.integers(0, U, n_spot)
    probe = np.minimum(ra_, rb_).astype(np.int64) * U + np.maximum(ra_, rb_)
    is_edge = key[np.clip(np.searchsorted(key, probe), 0, max(E - 1, 0))] == probe if E else np.zeros(n_spot, dtype=bool)
    for a, b, e_ in zip(ra_, rb_, is_edge):
        if a != b and not e_:
            assert lev(a, b) > 2 * p["max_diff"], ("missing pair", int(a), int(b))
    order = np.argsort(bor, kind="stable").astype(np.int32)
    bptr = np.zeros(U + 1, dtype=np.int32)
    np.cumsum(np.bincount(bor, minlength=U), out=bptr[1:])
    q, r, d = hostdata.canonical_rows(ei, ej, ed)
    prob = hostdata.MergeProblem(n_reads=R, n_buckets=U, bucket_ptr=bptr, bucket_reads=order, lexrank=args[3], bc_id=bor,
                                 up_id=args[4], geno_ptr=args[5], geno_mut=args[6], geno_q=args[7], mut_rank=args[8], **p)
    ref = orc.cluster(prob, q, r, d)
    rows_equal(hostdata.rows_from_merge_outputs(out_full), ref, "full table")
    out = ctx.cluster_reads(*args, **p)
    rows_equal(hostdata.rows_from_merge_outputs(out), ref, "on-demand distances")
    assert np.array_equal(np.sort(ref.row_members), np.arange(R))
    return dict(reads=R, buckets=U, edges_full=int(E), edges_on_demand=int(out["n_edges"]), rows=ref.n_rows,
                clusters=ref.n_clusters, links=ctx.stat("links"))
