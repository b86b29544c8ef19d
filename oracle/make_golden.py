#!/usr/bin/env python3
"""Generates tests/golden/*.json.gz by running the UNMODIFIED reference scripts here.

    python -m oracle.make_golden            (build container only; needs /root/reference)

Each fixture holds the text inputs the reference was given and what it produced:

    fastq        the extracted-barcode FASTQ fed to src/pacybara_precluster.py (stdin)
    preclust     that script's stdout, verbatim
    ed_rows      [[query title, ref title, dist], ...] in file order
    geno         the genoExtract text, verbatim
    uptag        the uptag FASTQ text (or null)
    params       -d -q -j -m
    exit_code    of src/pacybara_cluster.py (KAT-10 is a deliberate crash)
    table        the canonicalised cluster table (pacybara_b200.canon), PYTHONHASHSEED=0

The known-answer cases are KAT-1..10 / Q1 of SURVEY.md section 4.3; the random ones come from
pacybara_b200.synth with high error / near-neighbour / collision rates so that ordering,
transitivity and genotype rules all fire.  Where `ed_rows` is "canonical" the table was
written from exhaustive Levenshtein distances (oracle DP) in the canonical (q, r) row order.
"""
from __future__ import annotations

import gzip
import io
import json
import os
import sys

import numpy as np

from oracle import orc, ref_runner
from pacybara_b200 import canon, hostdata, synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

A = "ACGTACGTACGTACGTACGTACGTA"
A1 = "ACGTACGTACGTTCGTACGTACGTA"      # pos13 G>T
A1p = "ACGTACGTACGTACGTACGTACGTT"     # last A>T
A2 = "ACGTACGTACGTTCGTACGAACGTA"      # A1 + pos20 T>A
A3 = "ACGTACCTACGTTCGTACGAACGTA"      # A2 + pos7 G>C
B = "TTGCATTGCATTGCATTGCATTGCA"
B1 = "TTGCATTGCATTGCATTGCATTGCC"
Cc = "GGGGGCCCCCAAAAATTTTTGGGGG"


def _fq(reads, qual_char="I"):
    return "".join(f"@{rid} pos=17 len={len(bc)}\n{bc}\n+\n{qual_char * len(bc)}\n" for rid, bc, _ in reads)


def _geno(reads):
    return "".join(f"\"{rid}\",\"{g}\"\n" for rid, _, g in reads)


def kat_cases():
    W = "="
    def reads(*items):
        return [tuple(x.split(":", 2)) for x in items]
    def named(name, rds, ed, **params):
        return dict(name=name, reads=rds, ed=ed, params=params)
    a4 = [f"a{i}:{A}:=" for i in range(1, 5)]
    a8 = [f"a{i}:{A}:=" for i in range(1, 9)]
    cases = [
        named("kat1", [("r1", A, "10A>G:50;20C>T:40"), ("r2", A, "10A>G:45"), ("r3", A, "10A>G:60;33del:10"),
                       ("r4", A1p, "10A>G:50"), ("r5", Cc, W), ("r6", A, "99G>C:70;120insACG:66")],
              [("r4", "r1", 1), ("r1", "r4", 1)]),
        named("kat2a", reads(f"a1:{A}:=", f"b1:{A1}:="), [("a1", "b1", 1), ("b1", "a1", 1)]),
        named("kat2b", reads(f"a1:{A}:=", f"a2:{A}:=", f"b1:{A1}:="), [("a1", "b1", 1), ("b1", "a1", 1)]),
        named("kat2c", reads(*a4[:3], f"c1:{A2}:="), [("a1", "c1", 2), ("c1", "a1", 2)]),
        named("kat2d", reads(*a4, f"c1:{A2}:="), [("a1", "c1", 2), ("c1", "a1", 2)]),
        named("kat3a", reads(f"a1:{A}:=", f"a2:{A}:=", f"s:{A1}:=", f"c1:{A2}:=", f"c2:{A2}:="),
              [("a1", "s", 1), ("s", "c1", 1), ("a1", "c1", 2)]),
        named("kat3b", reads(f"a1:{A}:=", f"a2:{A}:=", f"s:{A1}:=", f"c1:{A2}:=", f"c2:{A2}:="),
              [("s", "c1", 1), ("a1", "s", 1), ("a1", "c1", 2)]),
        named("kat4a", reads(*a4, f"s:{A1}:=", f"t:{A2}:="), [("a1", "s", 1), ("s", "t", 1), ("a1", "t", 2)]),
        named("kat4b", reads(*a4, f"s:{A1}:=", f"t:{A2}:="), [("a1", "s", 1), ("s", "t", 1)]),
        named("kat4c", reads(*a8, f"s:{A1}:=", f"t:{A2}:=", f"u:{A3}:="),
              [("a1", "s", 1), ("s", "t", 1), ("t", "u", 1), ("a1", "t", 2), ("s", "u", 2), ("a1", "u", 3)]),
        named("kat5a", [("w1", A, W), ("w2", A, W), ("m1", A, "100A>G:80"), ("m2", A, "100A>G:75")], []),
        named("kat5b", [("p1", A, "1A>C:80;2A>C:80;3A>C:80"), ("p2", A, "1A>C:80;4A>C:80;5A>C:80"),
                        ("q1", B, "1A>C:80;2A>C:80;3A>C:80"), ("q2", B, "1A>C:80;2A>C:80;4A>C:80;5A>C:80")], []),
        named("kat5c", [("p1", A, "1A>C:80;2A>C:80"), ("p2", A, "1A>C:80")], [], m=2),
        named("kat6a", [("p1", A, "7C>T:32"), ("p2", A, W)], []),
        named("kat6b", [("p1", A, "7C>T:33"), ("p2", A, W)], []),
        named("kat6c", [("p1", A, "7C>T:10"), ("p2", A, W), ("p3", A, "7C>T:10")], []),
        named("kat7a", [("a1", A, "9G>A:65"), ("a2", A, "9G>A:1"), ("a3", A, "9G>A:1;50T>A:65")], []),
        named("kat7b", [("s1", A, "50T>A:32;9G>A:31;7C>T:99")], []),
        named("kat8a", [("m1", A, "100A>G:80"), ("m2", A, "100A>G:80"), ("s", A1, "100A>G:50"),
                        ("b1", B, "5A>T:90"), ("b2", B, "5A>T:90"), ("w", B1, W)],
              [("m1", "s", 1), ("b1", "w", 1)]),
        named("kat8c", [("m1", A, "100A>G:80;200C>T:80"), ("m2", A, "100A>G:80;200C>T:80"), ("s", A1, "100A>G:50")],
              [("m1", "s", 1)], m=2),
        named("kat9", reads(f"a1:{A}:=", f"a2:{A}:=", f"s:{A1}:="), [("a1", "s", 1)]),
        named("kat10", reads(f"x-1:{A}:=", f"x-2:{A}:=", f"y-1:{A1}:="), [("x-1", "y-1", 1)]),
        # extra hand-made cases: duplicate rows at different distances, reverse-only row,
        # rows beyond 2*maxDiff, zero-quality calls, duplicate mutation keys, d=0/3 limits
        named("x_duprows", reads(*a4, f"s:{A1}:=", f"t:{A2}:="),
              [("a1", "t", 3), ("s", "a1", 1), ("a1", "t", 2), ("t", "s", 1), ("a1", "s", 2), ("a1", "t", 9)]),
        named("x_zeroq", [("z1", A, "5A>C:0;6A>C:40"), ("z2", A, "5A>C:0"), ("z3", A, "6A>C:0;6A>C:41;9G>T:3")], []),
        named("x_d0", reads(f"a1:{A}:=", f"a2:{A}:=", f"b1:{A1}:="), [("a1", "b1", 1)], d=0),
        named("x_d3", reads(*a8, f"s:{A1}:=", f"t:{A2}:=", f"u:{A3}:="),
              [("a1", "s", 1), ("s", "t", 1), ("t", "u", 1), ("a1", "t", 2), ("s", "u", 2), ("a1", "u", 3)], d=3),
        named("x_lexorder", [("r10", A, W), ("r9", A, W), ("r100", A1, W), ("r2", A2, W), ("r11", A2, W)],
              [("r10", "r100", 1), ("r100", "r2", 1), ("r10", "r2", 2)]),
    ]
    return cases


def _params(p):
    return dict(max_diff=p.get("d", 2), min_qual=p.get("q", 32), min_jaccard=p.get("j", 0.2),
                min_matches=p.get("m", 1))


def run_case(name, fastq, geno, uptag, ed_spec, params, note=""):
    """ed_spec: list of (read-or-title, read-or-title, d) | "canonical"."""
    preclust = ref_runner.run_precluster(fastq)
    pc = hostdata.parse_preclust(io.StringIO(preclust))
    read2title = {}
    for b, t in enumerate(pc.titles):
        for rid in t.split("|"):
            read2title[rid] = t
    if ed_spec == "canonical":
        mat, lens = hostdata.strings_to_matrix([s.encode() for s in pc.seqs])
        ei, ej, ed = orc.edit_pairs(mat, lens, 2 * params["max_diff"], method="brute")
        q, r, d = hostdata.canonical_rows(ei, ej, ed)
        rows = [[pc.titles[a], pc.titles[b], int(c)] for a, b, c in zip(q, r, d)]
    else:
        rows = [[read2title.get(a, a), read2title.get(b, b), int(d)] for a, b, d in ed_spec]
    ed_text = hostdata.ED_HEADER + "\n" + "".join(f"\"{a}\",\"{b}\",NA,NA,{d}\n" for a, b, d in rows)
    code, out, err = ref_runner.run_cluster(ed_text, geno, preclust, uptag, **params)
    table = canon.canon_table(out) if out is not None else None
    return dict(name=name, note=note, fastq=fastq, preclust=preclust, ed_rows=rows, geno=geno, uptag=uptag,
                params=params, exit_code=code, stderr_tail=None if code == 0 else err[-300:],
                table=table, n_rows=None if table is None else len(table),
                digest=None if table is None else canon.table_digest(table))


def lib_texts(lib: synth.Library):
    ids = lib.read_ids()
    muts = lib.mut_strings()
    def fq(seq, lens, pos):
        return "".join(f"@{ids[i]} pos={pos} len={int(lens[i])}\n{seq[i, :lens[i]].tobytes().decode()}\n+\n"
                       f"{lib.qual[i, :lens[i]].tobytes().decode()}\n" for i in range(lib.n_reads))
    fastq = fq(lib.seq, lib.length, "NA" if lib.uptag_len is not None else 17)
    uptag = fq(lib.seq, lib.uptag_len, 17) if lib.uptag_len is not None else fastq
    geno = "".join(f"\"{ids[i]}\",\"{lib.geno_str(i, muts)}\"\n" for i in range(lib.n_reads))
    return fastq, geno, uptag


RANDOM_CASES = [
    # name, generator kwargs, cluster params
    ("rand_small_d2", dict(n_clones=60, n_reads=500, p_err=0.02, p_near=0.3, seed=11), dict()),
    ("rand_dense_d2", dict(n_clones=150, n_reads=2000, p_err=0.03, p_near=0.4, seed=12, skew=1.0), dict()),
    ("rand_collide", dict(n_clones=150, n_reads=1500, p_err=0.01, p_near=0.2, p_collide=0.3, seed=13,
                          q_clone=(20, 50), q_private=(3, 15)), dict(q=27)),
    ("rand_d1", dict(n_clones=200, n_reads=1500, p_err=0.02, p_near=0.3, seed=14), dict(d=1)),
    ("rand_d3_combo", dict(n_clones=120, n_reads=1200, combo=True, p_err=0.015, p_near=0.3, seed=15), dict(d=3)),
    ("rand_m2_j05", dict(n_clones=150, n_reads=1500, p_err=0.02, p_near=0.3, seed=16), dict(m=2, j=0.5)),
    ("rand_lexids", dict(n_clones=100, n_reads=1200, p_err=0.02, p_near=0.4, seed=17, id_style="plain",
                         skew=1.5), dict()),
    ("rand_shortbc", dict(n_clones=80, n_reads=800, bc_len=12, p_err=0.02, p_near=0.4, seed=18), dict(d=1)),
    ("rand_private_heavy", dict(n_clones=100, n_reads=1500, p_err=0.02, p_near=0.3, p_private=0.5,
                                p_private_hq=0.2, p_dropout=0.3, seed=19), dict(q=40)),
]


def textquirks_texts():
    """A small library whose files are formatted the way real tools sometimes leave them: CRLF records, trailing
    blanks, a quality string starting with '@', no newline at the end, unquoted and CRLF genotype lines, a third
    column, records of reads that were never extracted, a read recorded twice (the later record counts) -- everything
    the reference's line loops accept (src/pacybara_precluster.py:47-63, src/pacybara_cluster.py:302-331)."""
    lib = synth.make_library(name="x_textquirks", n_clones=40, n_reads=320, p_err=0.03, p_near=0.4, seed=23)
    ids, muts = lib.read_ids(), lib.mut_strings()
    fq, ge, up = [], [], []
    for i in range(lib.n_reads):
        n = int(lib.length[i])
        seq, qual = lib.seq[i, :n].tobytes().decode(), lib.qual[i, :n].tobytes().decode()
        if i % 37 == 5:
            qual = "@" + qual[1:]                                  # Phred 31 first: the reference drops this read
        eol = "\r\n" if i % 5 == 0 else "\n"
        pad = "  " if i % 7 == 0 else ""
        fq.append(f"@{ids[i]} pos=17 len={n}{eol}{seq}{pad}{eol}+{eol}{qual}{eol}")
        g = lib.geno_str(i, muts)
        if i % 11 == 3:
            ge.append(f"\"{ids[i]}\",\"999A>C:55\"\n")                  # an earlier record of the same read: overwritten below
        quote = "" if i % 4 == 1 else "\""
        tail = ",extra" if i % 9 == 2 else ""
        ge.append(f"{quote}{ids[i]}{quote},{quote}{g}{quote}{tail}" + ("\r\n" if i % 6 == 0 else "\n"))
        if i % 50 == 0:
            ge.append(f"\"ghost/{i}/ccs\",\"5A>T:40\"\n")                  # a read that is not in the FASTQ
        if i % 13 == 4:
            up.append(f"@{ids[i]} pos=1 len={n}\n{'T' * n}\n+\n{'I' * n}\n")   # an earlier uptag record of this read
        up.append(f"@{ids[i]} pos=17 len={n}{eol}{seq}{pad}{eol}+{eol}{'I' * n}{eol}")
    fastq, uptag = "".join(fq), "".join(up)
    return fastq[:-1] if fastq.endswith("\n") and not fastq.endswith("\r\n") else fastq, "".join(ge), uptag


def opaque_texts():
    """A small library in which some barcodes do not fit 2-bit planes: a 70-nt insertion call (three reads), a barcode
    with an N (two reads), one read with an N all of its own.  The reference treats them like any other string."""
    long_ = "ACGTTGCA" * 8 + "ACGTAC"                      # 70 nt
    n_bc = "TTGCANTGCATTGCATTGCATTGCA"
    reads = [("r01", A, "10A>G:50"), ("r02", long_, "7C>T:60"), ("r03", A, "10A>G:55"), ("r04", n_bc, "="), ("r05", A1, "10A>G:45"),
             ("r06", long_, "7C>T:58"), ("r07", B, "5A>T:90"), ("r08", n_bc, "="), ("r09", long_, "7C>T:61;99G>C:12"),
             ("r10", B, "5A>T:88"), ("r11", "GGGGGCCCCCAAAAANTTTTGGGGG", "3A>C:70"), ("r12", B1, "=")]
    return _fq(reads), _geno(reads), _fq(reads)


def add_opaque():
    """Adds tests/golden/x_opaque.json.gz (and its INDEX entry) without touching the other fixtures."""
    fastq, geno, uptag = opaque_texts()
    case = run_case("x_opaque", fastq, geno, uptag, "canonical", _params({}), note="oracle.make_golden.opaque_texts()")
    assert case["exit_code"] == 0, case["stderr_tail"]
    with gzip.GzipFile(os.path.join(GOLDEN_DIR, "x_opaque.json.gz"), "wb", mtime=0) as fh:
        fh.write(json.dumps(case, indent=0).encode())
    ipath = os.path.join(GOLDEN_DIR, "INDEX.json")
    with open(ipath) as fh:
        index = [e for e in json.load(fh) if e["name"] != "x_opaque"]
    index.append(dict(name="x_opaque", exit_code=case["exit_code"], n_rows=case["n_rows"], digest=case["digest"]))
    with open(ipath, "w") as fh:
        json.dump(index, fh, indent=1)
    print(f"x_opaque exit={case['exit_code']} rows={case['n_rows']}")
    for row in case["table"]:
        print("  ", row)


DENSE_GEN = dict(n_clones=450, n_reads=3200, bc_len=13, pattern="SW", p_err=0.01, p_near=0.0, seed=61)


def add_dense():
    """Adds tests/golden/x_dense_sw.json.gz: the reference template's SWSW... barcode design at a density where the link
    graph percolates (450 clones in 2^13 codes: nearly all reads hang together in ONE component)."""
    lib = synth.make_library(name="x_dense_sw", **DENSE_GEN)
    fastq, geno, uptag = lib_texts(lib)
    case = run_case("x_dense_sw", fastq, geno, uptag, "canonical", _params({}), note=f"synth.make_library({DENSE_GEN})")
    assert case["exit_code"] == 0, case["stderr_tail"]
    with gzip.GzipFile(os.path.join(GOLDEN_DIR, "x_dense_sw.json.gz"), "wb", mtime=0) as fh:
        fh.write(json.dumps(case, indent=0).encode())
    ipath = os.path.join(GOLDEN_DIR, "INDEX.json")
    with open(ipath) as fh:
        index = [e for e in json.load(fh) if e["name"] != "x_dense_sw"]
    index.append(dict(name="x_dense_sw", exit_code=case["exit_code"], n_rows=case["n_rows"], digest=case["digest"]))
    with open(ipath, "w") as fh:
        json.dump(index, fh, indent=1)
    print(f"x_dense_sw exit={case['exit_code']} rows={case['n_rows']} ed_rows={len(case['ed_rows'])}")


def add_textquirks():
    """Adds tests/golden/x_textquirks.json.gz (and its INDEX entry) without touching the other fixtures."""
    fastq, geno, uptag = textquirks_texts()
    case = run_case("x_textquirks", fastq, geno, uptag, "canonical", _params({}), note="oracle.make_golden.textquirks_texts()")
    assert case["exit_code"] == 0, case["stderr_tail"]
    with gzip.GzipFile(os.path.join(GOLDEN_DIR, "x_textquirks.json.gz"), "wb", mtime=0) as fh:
        fh.write(json.dumps(case, indent=0).encode())
    ipath = os.path.join(GOLDEN_DIR, "INDEX.json")
    with open(ipath) as fh:
        index = [e for e in json.load(fh) if e["name"] != "x_textquirks"]
    index.append(dict(name="x_textquirks", exit_code=case["exit_code"], n_rows=case["n_rows"], digest=case["digest"]))
    with open(ipath, "w") as fh:
        json.dump(index, fh, indent=1)
    print(f"x_textquirks exit={case['exit_code']} rows={case['n_rows']}")


def main():
    if "--add-dense" in sys.argv:
        if not ref_runner.available():
            sys.exit("reference scripts not found")
        return add_dense()
    if "--add-opaque" in sys.argv:
        if not ref_runner.available():
            sys.exit("reference scripts not found")
        return add_opaque()
    if "--add-textquirks" in sys.argv:
        if not ref_runner.available():
            sys.exit("reference scripts not found")
        return add_textquirks()
    if not ref_runner.available():
        sys.exit("reference scripts not found; golden vectors can only be generated in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    index = []

    def save(case):
        path = os.path.join(GOLDEN_DIR, case["name"] + ".json.gz")
        with gzip.GzipFile(path, "wb", mtime=0) as fh:
            fh.write(json.dumps(case, indent=0).encode())
        index.append(dict(name=case["name"], exit_code=case["exit_code"], n_rows=case["n_rows"], digest=case["digest"]))
        print(f"{case['name']:22s} exit={case['exit_code']} rows={case['n_rows']}")

    for c in kat_cases():
        fastq, geno = _fq(c["reads"]), _geno(c["reads"])
        save(run_case(c["name"], fastq, geno, fastq, c["ed"], _params(c["params"])))

    # Q1: quality line starting with '@' (Phred 31) makes the reference drop the read
    q1 = "@a pos=1 len=4\nACGT\n+\nIIII\n@b pos=1 len=4\nACGT\n+\n@III\n@c pos=1 len=4\nACGT\n+\nIIII\n@d pos=1 len=4\nTTTT\n+\n!~5I\n"
    g1 = "".join(f"\"{r}\",\"=\"\n" for r in "abcd")
    save(run_case("q1_at_quality", q1, g1, None, [], _params({})))

    for name, gen, par in RANDOM_CASES:
        lib = synth.make_library(name=name, **gen)
        fastq, geno, uptag = lib_texts(lib)
        case = run_case(name, fastq, geno, uptag, "canonical", _params(par), note=f"synth.make_library({gen})")
        if case["exit_code"] != 0:
            print(f"!! {name}: reference exited {case['exit_code']}: {case['stderr_tail']}")
        save(case)

    # cfg 1 at full size: inputs are regenerated from the seed at test time; only digests are stored
    lib = synth.make_config("cfg1", seed=1)
    fastq, geno, uptag = lib_texts(lib)
    case = run_case("cfg1_seed1", fastq, geno, uptag, "canonical", _params({}), note="synth.make_config('cfg1', seed=1)")
    import hashlib
    slim = dict(name=case["name"], note=case["note"], params=case["params"], exit_code=case["exit_code"],
                n_rows=case["n_rows"], digest=case["digest"], n_ed_rows=len(case["ed_rows"]),
                input_digest=hashlib.sha256((fastq + geno).encode()).hexdigest(),
                preclust_digest=hashlib.sha256(case["preclust"].encode()).hexdigest(),
                table_head=case["table"][:20])
    with gzip.GzipFile(os.path.join(GOLDEN_DIR, "cfg1_seed1.json.gz"), "wb", mtime=0) as fh:
        fh.write(json.dumps(slim, indent=0).encode())
    index.append(dict(name="cfg1_seed1", exit_code=case["exit_code"], n_rows=case["n_rows"], digest=case["digest"]))
    print(f"cfg1_seed1 rows={case['n_rows']} ed_rows={len(case['ed_rows'])}")

    with open(os.path.join(GOLDEN_DIR, "INDEX.json"), "w") as fh:
        json.dump(index, fh, indent=1)


if __name__ == "__main__":
    main()
