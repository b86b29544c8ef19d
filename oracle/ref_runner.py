"""Runs the UNMODIFIED reference scripts as subprocesses (golden-vector generation only).

/root/reference exists only in the build container, never on the GPU box, so nothing in
tests/, smoke() or bench.py calls this at run time: oracle/make_golden.py uses it once to
produce tests/golden/*.json.gz, which are committed.
"""
from __future__ import annotations

import gzip
import os
import subprocess
import sys
import tempfile
from typing import Dict, List, Optional, Sequence, Tuple

REF_SRC = os.environ.get("PACYBARA_REFERENCE", "/root/reference/src")


def available() -> bool:
    return os.path.exists(os.path.join(REF_SRC, "pacybara_cluster.py"))


def _env(hashseed: int) -> Dict[str, str]:
    env = dict(os.environ)
    env["PYTHONHASHSEED"] = str(hashseed)
    return env


def run_precluster(fastq_text: str, hashseed: int = 0) -> str:
    """stdin FASTQ -> stdout FASTQ (src/pacybara_precluster.py)."""
    p = subprocess.run([sys.executable, os.path.join(REF_SRC, "pacybara_precluster.py")],
                       input=fastq_text, capture_output=True, text=True, env=_env(hashseed), check=True)
    return p.stdout


def run_cluster(ed_text: str, geno_text: str, preclust_text: str, uptag_text: Optional[str],
                max_diff: int = 2, min_qual: int = 32, min_jaccard: float = 0.2, min_matches: int = 1,
                hashseed: int = 0) -> Tuple[int, Optional[List[List[str]]], str]:
    """Returns (exit code, rows of clusters.csv.gz or None, stderr tail)."""
    with tempfile.TemporaryDirectory() as td:
        def put(name, text):
            path = os.path.join(td, name)
            with gzip.open(path, "wt") as fh:
                fh.write(text)
            return path
        ed = put("editDistance.csv.gz", ed_text)
        ge = put("genoExtract.csv.gz", geno_text)
        pc = put("bcPreclust.fastq.gz", preclust_text)
        out = os.path.join(td, "clusters.csv.gz")
        cmd = [sys.executable, "-u", os.path.join(REF_SRC, "pacybara_cluster.py"), "--out", out,
               "--minJaccard", repr(min_jaccard), "--minMatches", str(min_matches),
               "--maxDiff", str(max_diff), "--minQual", str(min_qual)]
        if uptag_text is not None:
            cmd += ["--uptagBarcodeFile", put("uptag.fastq.gz", uptag_text)]
        cmd += [ed, ge, pc]
        p = subprocess.run(cmd, capture_output=True, text=True, env=_env(hashseed))
        rows = None
        if p.returncode == 0 and os.path.exists(out):
            with gzip.open(out, "rt") as fh:
                header = fh.readline().rstrip("\n")
                assert header == "virtualBarcode,upBarcode,reads,size,geno", header
                rows = [line.rstrip("\n").split(",") for line in fh if line.strip()]
        return p.returncode, rows, p.stderr[-2000:]
