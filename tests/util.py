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
