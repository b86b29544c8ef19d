#!/usr/bin/env python3
"""Build-container only: random, oddly formatted FASTQ text through the UNMODIFIED src/pacybara_precluster.py and through
the host parser + oracle bucketing of this repo; the two outputs must be the same text (or both must fail).

    python -m oracle.fuzz_against_reference [seconds] [first_seed]
    python -m oracle.fuzz_against_reference --cluster [seconds] [first_seed]     (whole libraries: precluster + cluster vs the oracle)

This pins pacybara_b200.hostdata.parse_fastq_for_precluster (the line parser the device tokenizers are tested against
and hand odd files to) and oracle bucketing to the reference beyond the committed fixtures.
"""
import io
import subprocess
import sys
import time

import numpy as np

from oracle import ref_runner
from pacybara_b200 import hostdata
from tests import util


def random_fastq(rng) -> str:
    n = int(rng.integers(1, 40))
    pool = ["".join(rng.choice(list("ACGT"), int(rng.integers(1, 12)))) for _ in range(int(rng.integers(1, 6)))]
    out = []
    for i in range(n):
        seq = pool[int(rng.integers(len(pool)))]
        qual = "".join(chr(int(c)) for c in rng.integers(33, 127, len(seq)))
        if rng.random() < 0.1:
            qual = "@" + qual[1:]
        eol = "\r\n" if rng.random() < 0.2 else "\n"
        pad = rng.choice(["", "", " ", "\t", "  "])
        head = f"@r{int(rng.integers(1000))}/{i}" + rng.choice(["", " pos=17", " pos=NA len=3", "\tx"])
        rec = f"{head}{pad}{eol}{seq}{pad}{eol}+{rng.choice(['', 'r'])}{eol}{qual}{pad if pad != ' ' or True else ''}{eol}"
        if rng.random() < 0.05:
            rec += eol                                   # blank line between records
        if rng.random() < 0.03:
            rec = rec.replace("+", "", 1)                # a record with a line missing
        out.append(rec)
    txt = "".join(out)
    if rng.random() < 0.3:
        txt = txt.rstrip("\r\n")
    return txt


def ours(text: str):
    try:
        reads = hostdata.parse_fastq_for_precluster(io.StringIO(text, newline=None))
    except Exception as e:                                # noqa: BLE001
        return None, type(e).__name__
    if len(reads.ids) == 0:
        return "", None
    try:
        return util.preclust_text(reads, util.oracle_buckets(reads)), None
    except Exception as e:                                # noqa: BLE001
        return None, type(e).__name__


def mangle(rng, geno: str, uptag: str):
    """Formatting the reference's loaders accept (src/pacybara_cluster.py:302-331): quotes or none, CRLF, extra columns,
    records of unknown reads, a read recorded twice (the later one counts), trailing blanks; uptag records likewise."""
    out = []
    for line in geno.splitlines():
        rid, g = [f.strip('"') for f in line.split(",")[:2]]
        if rng.random() < 0.1:
            out.append(f'"{rid}","77A>C:{int(rng.integers(1, 90))}"\n')            # overwritten by the real record below
        q = '"' if rng.random() < 0.7 else ""
        tail = ",x,y" if rng.random() < 0.1 else ""
        pad = rng.choice(["", "", " ", "\t"])
        out.append(f"{q}{rid}{q},{q}{g}{q}{tail}{pad}" + ("\r\n" if rng.random() < 0.2 else "\n"))
        if rng.random() < 0.05:
            out.append(f"ghost{int(rng.integers(99))},5A>T:40\n")
    up = []
    lines = uptag.split("\n")
    for k in range(0, len(lines) - 3, 4):                       # lib_texts writes plain 4-line records
        head, seq, _, qual = lines[k: k + 4]
        if rng.random() < 0.1:
            up.append(f"{head}\n{'A' * len(seq)}\n+\n{qual}\n")                    # an earlier record of the same read
        eol = "\r\n" if rng.random() < 0.2 else "\n"
        up.append(f"{head}{eol}{seq}{rng.choice(['', ' '])}{eol}+{eol}{qual}{eol}")
    return "".join(out), "".join(up)


def cluster_fuzz(budget: float, seed: int) -> int:
    """Random small libraries through the reference's precluster + cluster scripts and through the oracle."""
    from oracle import make_golden, orc
    from pacybara_b200 import canon, synth
    t0, n, bad, died = time.time(), 0, 0, 0
    while time.time() - t0 < budget:
        rng = np.random.default_rng(seed)
        kw = dict(n_clones=int(rng.integers(3, 40)), n_reads=int(rng.integers(20, 300)), bc_len=int(rng.choice([10, 25])),
                  p_err=float(rng.choice([0.0, 0.03, 0.06])), p_near=float(rng.choice([0.0, 0.4, 0.7])),
                  p_collide=float(rng.choice([0.0, 0.3])), seed=seed, combo=bool(rng.random() < 0.2),
                  id_style=str(rng.choice(["padded", "plain"])), skew=float(rng.choice([0.0, 1.5])))
        if rng.random() < 0.5:
            kw.update(p_private=0.5, p_private_hq=float(rng.choice([0.0, 0.3])), p_dropout=float(rng.choice([0.0, 0.3])))
        par = dict(d=int(rng.integers(0, 4)), q=int(rng.choice([1, 20, 32, 60])), j=float(rng.choice([0.05, 0.2, 0.5, 1.0])),
                   m=int(rng.choice([1, 1, 2])))
        lib = synth.make_library(name=f"fuzz{seed}", **kw)
        fastq, geno, uptag = make_golden.lib_texts(lib)
        if rng.random() < 0.6:
            geno, uptag = mangle(rng, geno, uptag)
        case = make_golden.run_case(f"fuzz{seed}", fastq, geno, uptag, "canonical", make_golden._params(par))
        job = util.job_from_golden(case)
        rows = orc.cluster(job.problem, job.ed_q, job.ed_r, job.ed_d)
        if case["exit_code"] != 0:
            died += 1                                        # muscle is not installed: the reference dies in its tie fallback
            if not ((rows.row_bc == hostdata.NEEDS_ALIGNMENT).any() or (rows.row_up == hostdata.NEEDS_ALIGNMENT).any()):
                bad += 1
                print(f"seed {seed}: reference exited {case['exit_code']} ({case['stderr_tail'][-120:]!r}) but the oracle flags no tie", flush=True)
        else:
            got, want = util.table_of(job, rows), util.golden_table(case)
            if got != want:
                bad += 1
                print(f"seed {seed}: tables differ ({kw}, {par})\n{canon.diff_tables(got, want)[:600]}", flush=True)
        seed += 1
        n += 1
    print(f"{n} libraries through the reference, {bad} discrepancies, {died} where the reference needed muscle")
    return 1 if bad else 0


def main():
    if not ref_runner.available():
        sys.exit("reference scripts not found (build container only)")
    if "--cluster" in sys.argv:
        args = [a for a in sys.argv[1:] if a != "--cluster"]
        return cluster_fuzz(float(args[0]) if args else 60.0, int(args[1]) if len(args) > 1 else 1)
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    t0, n, bad, both_fail = time.time(), 0, 0, 0
    while time.time() - t0 < budget:
        rng = np.random.default_rng(seed)
        text = random_fastq(rng)
        try:
            want, want_err = ref_runner.run_precluster(text), None
        except subprocess.CalledProcessError as e:
            want, want_err = None, (e.stderr or "").strip().splitlines()[-1:]
        got, got_err = ours(text)
        if (want is None) != (got is None):
            # the one deliberate difference: a quality line whose length differs from its sequence is refused here,
            # where the reference would write a record whose lines disagree (or die on the index, depending on order)
            if got_err == "FormatError":
                both_fail += 1
            else:
                bad += 1
                print(f"seed {seed}: reference {'failed ' + str(want_err) if want is None else 'ok'}, this repo {'failed ' + str(got_err) if got is None else 'ok'}", flush=True)
        elif want is None:
            both_fail += 1
        elif want != got:
            bad += 1
            print(f"seed {seed}: outputs differ\n--- reference\n{want[:400]}\n--- this repo\n{got[:400]}", flush=True)
        seed += 1
        n += 1
    print(f"{n} texts, {bad} discrepancies, {both_fail} refused by both (or refused here for unequal lengths)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
