#!/usr/bin/env python3
"""Wall time of the clustering section of pacybara.sh (:425-489) with the drop-ins on PATH, on a synthetic library:
    python scripts/time_dropins.py [cfg] [scale]
Prints one JSON line: seconds per command and in total, rows of the result."""
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pacybara_b200 import synth  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    lib = synth.make_config(cfg, seed=1, scale=scale)
    d = tempfile.mkdtemp(prefix="pcb_dropins_")
    try:
        ex, cl = os.path.join(d, "extract"), os.path.join(d, "cluster")
        os.makedirs(cl)
        lib.write_files(ex)
        # `muscle` (the reference's tie-break for a barcode vote, :484-511) is not on the GPU box: a stand-in that hands the
        # sequences back unaligned, for the few tied rows of a synthetic library (timing only, never a parity claim)
        fake = os.path.join(d, "fakebin")
        os.makedirs(fake)
        with open(os.path.join(fake, "muscle"), "w") as fh:
            fh.write("#!/usr/bin/env python3\n# usage: muscle -align IN -output OUT -- pads the sequences to one length\nimport sys\n"
                     "rec = open(sys.argv[2]).read().split()\nn = max(len(s) for s in rec[1::2])\n"
                     "open(sys.argv[4], 'w').write(''.join(f'{h}\\n{s.ljust(n, chr(45))}\\n' for h, s in zip(rec[0::2], rec[1::2])))\n")
        os.chmod(os.path.join(fake, "muscle"), 0o755)
        have_muscle = shutil.which("muscle") is not None
        env = dict(os.environ, PATH=os.pathsep.join([os.path.join(ROOT, "bin", "shims"), os.path.join(ROOT, "bin")] +
                                                    ([] if have_muscle else [fake]) + [os.environ["PATH"]]))
        p = lib.params
        laps = {}

        def run(name, cmd):
            t = time.perf_counter()
            r = subprocess.run(cmd, shell=True, check=False, env=env, executable="/bin/bash", stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
            laps[name] = round(time.perf_counter() - t, 3)
            if r.returncode != 0:
                raise SystemExit(f"{name} failed ({r.returncode}):\n{r.stderr[-3000:]}")

        run("precluster", f"zcat {ex}/bcExtract_1.fastq.gz | pacybara_precluster.py | gzip -1 -c > {cl}/bcPreclust.fastq.gz")
        run("calcEdits", f"pacybara_calcEdits.R {cl}/bcMatches.sam.gz {cl}/bcPreclust.fastq.gz --maxErr {2 * p['max_diff']} --output {cl}/editDistance.csv.gz")
        run("cluster", f"python3 -u $(command -v pacybara_cluster.py) --uptagBarcodeFile {ex}/bcExtract_1.fastq.gz --out {cl}/clusters.csv.gz "
                       f"--minJaccard 0.2 --minMatches 1 --maxDiff {p['max_diff']} --minQual {p['min_qual']} "
                       f"{cl}/editDistance.csv.gz {ex}/genoExtract.csv.gz {cl}/bcPreclust.fastq.gz")
        with gzip.open(os.path.join(cl, "clusters.csv.gz"), "rt") as fh:
            rows = sum(1 for _ in fh) - 1
        print(json.dumps(dict(cfg=cfg, scale=scale, reads=lib.n_reads, seconds=laps, total_s=round(sum(laps.values()), 3), rows=rows,
                              note="each command is a fresh process (interpreter + CUDA context start-up included)")))
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
