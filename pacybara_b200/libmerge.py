"""Row f4: merging two Pacybara libraries -- the drop-in for src/pacybara_merge.R.

    python -m pacybara_b200.libmerge LIB1.csv.gz LIB2.csv.gz [--out merged.csv.gz]

Both inputs are translated cluster tables (columns virtualBarcode, upBarcode, reads, size, hgvsc, ...).  Rows of the
second library whose virtual barcode and genotype are already in the first are merged into that row (sizes add up, read
lists are joined with "|"); the others are appended; the collision columns are computed afresh
(src/pacybara_merge.R:48-116).  The row classification and the size bookkeeping run on the GPU (pcb_merge_libraries,
pcb_tag_rows); this module interns the strings and assembles the table.
"""
from __future__ import annotations

import argparse
import csv
import gzip
import sys
from typing import Dict, List, Sequence

import numpy as np

from . import hostdata


def merge_tables(ctx, lib1: Sequence[Dict], lib2: Sequence[Dict]) -> List[Dict]:
    """Rows as dicts (virtualBarcode, upBarcode, reads, size, hgvsc, ...) -> the merged table with fresh collision tags."""
    n1 = len(lib1)
    vbc, vbc_names = hostdata.intern_strings([r["virtualBarcode"] for r in lib1] + [r["virtualBarcode"] for r in lib2])
    geno, _ = hostdata.intern_strings([r["hgvsc"] for r in lib1] + [r["hgvsc"] for r in lib2])
    size = np.asarray([int(r["size"]) for r in lib1] + [int(r["size"]) for r in lib2], dtype=np.int32)
    res = ctx.merge_libraries(vbc[:n1], geno[:n1], size[:n1], vbc[n1:], geno[n1:], size[n1:], len(vbc_names))
    out = [dict(r) for r in lib1]
    for i, r in enumerate(out):
        r["size"] = int(res["size1"][i])
    appended = []
    for row, (k, t) in enumerate(zip(res["klass2"].tolist(), res["target2"].tolist())):
        if k == 1:
            out[t]["reads"] = out[t]["reads"] + "|" + lib2[row]["reads"]        # in the order of the second library's rows
        else:
            appended.append(dict(lib2[row]))
    out += appended
    bc_id, bc_names = hostdata.intern_strings([r["virtualBarcode"] for r in out])
    up_id, up_names = hostdata.intern_strings([r["upBarcode"] for r in out])
    tags = ctx.tag_rows(bc_id, up_id, np.asarray([int(r["size"]) for r in out], dtype=np.int32), len(bc_names), len(up_names))
    for i, r in enumerate(out):
        r["collision"] = "collision" if tags["collision"][i] else ""
        r["upTagCollision"] = "collision" if tags["up_collision"][i] else ""
    return out


def read_table(path: str) -> List[Dict]:
    with hostdata.open_text(path) as fh:
        rows = list(csv.DictReader(fh))
    for r in rows:
        r["size"] = int(r["size"])
    return rows


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="pacybara_b200.libmerge", description="Merge Pacybara libraries")
    ap.add_argument("lib1", help="first translated clusters csv.gz file")
    ap.add_argument("lib2", help="second translated clusters csv.gz file")
    ap.add_argument("--out", default="merged.csv.gz", help="output file name")
    args = ap.parse_args(argv)
    from .api import open_context_or_exit
    lib1, lib2 = read_table(args.lib1), read_table(args.lib2)
    ctx = open_context_or_exit(0)
    out = merge_tables(ctx, lib1, lib2)
    ctx.close()
    cols = list(lib1[0].keys()) if lib1 else (list(lib2[0].keys()) if lib2 else [])
    for c in ("collision", "upTagCollision"):
        if c not in cols:
            cols.append(c)
    with gzip.open(args.out, "wt", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=cols, quoting=csv.QUOTE_NONNUMERIC, extrasaction="ignore")      # write.csv quotes strings
        w.writeheader()
        w.writerows(out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
