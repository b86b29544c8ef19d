"""The drop-in command-line boundary: same flags, files, exit codes."""
import gzip
import io
import os
import subprocess
import sys

import pytest

from pacybara_b200 import canon, cluster
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")


def _run_main(argv, capsys):
    rc = cluster.main(argv)
    cap = capsys.readouterr()
    return rc, cap.out, cap.err


# ---------------------------------------------------------------- argument handling (no GPU needed)
def test_missing_positionals_exit_1_with_usage(capsys):
    rc, out, err = _run_main([], capsys)
    assert rc == 1 and "Usage: pacybara_cluster.py" in out and "Missing required arguments" in err


def test_help_exits_0(capsys):
    rc, out, _ = _run_main(["--help"], capsys)
    assert rc == 0 and "Usage:" in out


@pytest.mark.parametrize("extra,msg", [
    (["-j", "0"], "minJaccard must be between 0 and 1"), (["-j", "1.5"], "minJaccard must be between 0 and 1"),
    (["-j", "x"], "minJaccard must be a number"), (["-m", "0"], "minMatches must be at least 1"),
    (["-d", "4"], "maxDiff must be between 0 and 3"), (["-d", "-1"], "maxDiff must be between 0 and 3"),
    (["-q", "0"], "minQual must be a positive integer"), (["-q", "x"], "minQual must be an integer number"),
    (["--bogus"], "not recognized"),
])
def test_option_validation(tmp_path, capsys, extra, msg):
    files = []
    for n in ("ed.csv.gz", "geno.csv.gz", "pre.fastq.gz"):
        p = tmp_path / n
        with gzip.open(p, "wt") as fh:
            fh.write("")
        files.append(str(p))
    rc, out, err = _run_main(extra + files, capsys)
    assert rc == 1 and msg in err and "Usage:" in out


def test_missing_file_and_extra_args(tmp_path, capsys):
    rc, _, err = _run_main(["a", "b", "c"], capsys)
    assert rc == 1 and "does not exist" in err
    rc, _, err = _run_main(["a", "b", "c", "d"], capsys)
    assert rc == 1 and "Unrecognized additional arguments" in err


def test_dash_in_read_id_is_refused(tmp_path, capsys):
    """KAT-10: the reference dies with ValueError at :418; the drop-in refuses before any work."""
    case = util.load_golden("kat10")
    paths = {}
    for name, text in (("ed", util.ed_text(case)), ("geno", case["geno"]), ("pre", case["preclust"])):
        paths[name] = str(tmp_path / f"{name}.gz")
        with gzip.open(paths[name], "wt") as fh:
            fh.write(text)
    rc, _, err = _run_main([paths["ed"], paths["geno"], paths["pre"]], capsys)
    assert rc == 1 and "contains '-'" in err and case["exit_code"] == 1


def test_bin_scripts_exist_and_are_executable():
    for n in ("pacybara_precluster.py", "pacybara_cluster.py", "pacybara_calcEdits.R",
              "shims/seqret", "shims/bowtie2-build", "shims/bowtie2"):
        p = os.path.join(BIN, n)
        assert os.path.exists(p) and os.access(p, os.X_OK), p


# ---------------------------------------------------------------- the commands themselves (GPU)
def _write_extract_dir(case, d):
    os.makedirs(d, exist_ok=True)
    with gzip.open(os.path.join(d, "bcExtract_1.fastq.gz"), "wt") as fh:
        fh.write(case["uptag"] if case["uptag"] is not None else case["fastq"])
    with gzip.open(os.path.join(d, "bcExtract_combo.fastq.gz"), "wt") as fh:
        fh.write(case["fastq"])
    with gzip.open(os.path.join(d, "genoExtract.csv.gz"), "wt") as fh:
        fh.write(case["geno"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["kat1", "kat3a", "q1_at_quality", "rand_dense_d2", "rand_collide"])
def test_precluster_cli(name):
    case = util.load_golden(name)
    p = subprocess.run([sys.executable, os.path.join(BIN, "pacybara_precluster.py")], input=case["fastq"],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert p.stdout == case["preclust"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["kat1", "kat3a", "kat3b", "kat8a", "x_duprows", "rand_dense_d2", "rand_d3_combo"])
def test_cluster_cli_with_the_golden_ed_file(tmp_path, name):
    """pacybara_cluster.py honours the EDFILE it is given (row order, duplicates, directions)."""
    case = util.load_golden(name)
    files = {}
    for key, text in (("ed", util.ed_text(case)), ("geno", case["geno"]), ("pre", case["preclust"]), ("up", case["uptag"])):
        files[key] = str(tmp_path / f"{key}.gz")
        with gzip.open(files[key], "wt") as fh:
            fh.write(text)
    out = str(tmp_path / "clusters.csv.gz")
    p = case["params"]
    cmd = [sys.executable, os.path.join(BIN, "pacybara_cluster.py"), "-u", files["up"], "-o", out, "-j", repr(p["min_jaccard"]),
           "-m", str(p["min_matches"]), "-d", str(p["max_diff"]), "-q", str(p["min_qual"]), files["ed"], files["geno"], files["pre"]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Done!" in r.stdout
    got = canon.canon_table(canon.read_clusters_csv(out))
    want = util.golden_table(case)
    assert got == want, canon.diff_tables(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("name,mode", [("rand_dense_d2", "uptag"), ("rand_collide", "uptag"), ("rand_d3_combo", "virtual")])
def test_clustering_section_of_the_driver(tmp_path, name, mode):
    """The reference driver's call sequence (tests/clustering_section.sh, after src/pacybara.sh:425-489)
    with bin/shims and bin first on PATH: precluster -> calcEdits -> cluster, files in the same places."""
    case = util.load_golden(name)
    ext, clu = str(tmp_path / "extract"), str(tmp_path / "clustering")
    _write_extract_dir(case, ext)
    p = case["params"]
    env = dict(os.environ)
    env["PATH"] = os.path.join(BIN, "shims") + os.pathsep + BIN + os.pathsep + env["PATH"]
    r = subprocess.run(["bash", os.path.join(ROOT, "tests", "clustering_section.sh"), ext, clu, str(p["max_diff"]),
                        str(p["min_qual"]), repr(p["min_jaccard"]), str(p["min_matches"]), mode],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    with gzip.open(os.path.join(clu, "bcPreclust.fastq.gz"), "rt") as fh:
        assert fh.read() == case["preclust"]
    with gzip.open(os.path.join(clu, "editDistance.csv.gz"), "rt") as fh:
        lines = fh.read().splitlines()
    assert lines[0] == '"query","ref","mismatch","indel","dist"'
    rows = [[a.strip('"'), b.strip('"'), int(d)] for a, b, _, _, d in (l.split(",") for l in lines[1:])]
    assert rows == case["ed_rows"]                      # the canonical table the golden run was fed
    got = canon.canon_table(canon.read_clusters_csv(os.path.join(clu, "clusters.csv.gz")))
    assert got == util.golden_table(case)
    for qc in ("bcPreclust_distr.txt", "edDistr.txt", "clusters_distr.txt"):
        assert os.path.getsize(os.path.join(clu, "qc", qc)) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("name,extra", [("rand_dense_d2", []), ("rand_collide", ["--full-table"]), ("rand_d3_combo", [])])
def test_fused_cli(tmp_path, name, extra):
    """bin/pacybara_b200_cluster: reads + genotypes in, the reference's cluster table out, one GPU call."""
    case = util.load_golden(name)
    ext = str(tmp_path / "extract")
    _write_extract_dir(case, ext)
    p = case["params"]
    out = str(tmp_path / "clusters.csv.gz")
    bc = "bcExtract_combo.fastq.gz" if name == "rand_d3_combo" else "bcExtract_1.fastq.gz"
    cmd = [sys.executable, os.path.join(BIN, "pacybara_b200_cluster"), "-u", os.path.join(ext, "bcExtract_1.fastq.gz"), "-o", out,
           "-j", repr(p["min_jaccard"]), "-m", str(p["min_matches"]), "-d", str(p["max_diff"]), "-q", str(p["min_qual"])] + extra + \
          [os.path.join(ext, bc), os.path.join(ext, "genoExtract.csv.gz")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    got = canon.canon_table(canon.read_clusters_csv(out))
    assert got == util.golden_table(case)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rand_dense_d2", "rand_d3_combo"])
def test_fused_cli_two_gpus(tmp_path, name):
    """The same command under torchrun on 2 GPUs: sharded pair search and merge, rank 0 writes the one table."""
    import socket
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    case = util.load_golden(name)
    ext = str(tmp_path / "extract")
    _write_extract_dir(case, ext)
    p = case["params"]
    out = str(tmp_path / "clusters.csv.gz")
    bc = "bcExtract_combo.fastq.gz" if name == "rand_d3_combo" else "bcExtract_1.fastq.gz"
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(BIN, "pacybara_b200_cluster"), "-u", os.path.join(ext, "bcExtract_1.fastq.gz"),
           "-o", out, "-j", repr(p["min_jaccard"]), "-m", str(p["min_matches"]), "-d", str(p["max_diff"]), "-q", str(p["min_qual"]),
           os.path.join(ext, bc), os.path.join(ext, "genoExtract.csv.gz")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    got = canon.canon_table(canon.read_clusters_csv(out))
    assert got == util.golden_table(case)
