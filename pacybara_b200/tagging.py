"""Row f1: collision tags and the soft filter on a cluster table.

    python -m pacybara_b200.tagging clusters.csv.gz [--minRatio 0.666] [--out clusters_tagged.csv.gz]

Adds to every row of a clusters.csv.gz the two columns src/pacybara_translate.R:58-68 computes (`collision`,
`upTagCollision`: "collision" when the virtual / uptag barcode occurs in more than one row, else empty) and a
`usable` column holding the verdict of src/pacybara_softfilter.R:46-63 (the cluster holds more than minRatio of the
reads of its uptag and has more than one read).  The HGVS translation of pacybara_translate.R is out of scope.
The counting runs on the GPU (pcb_tag_rows); this module is the host glue.
"""
from __future__ import annotations

import argparse
import gzip
import sys
from typing import List, Sequence, Tuple

import numpy as np

from . import canon, hostdata


def tag_table(ctx, table: Sequence[Sequence], min_ratio: float = 0.666) -> List[Tuple]:
    """table rows (virtualBarcode, upBarcode, reads, size, geno) -> rows + (collision, upTagCollision, usable)."""
    bc_id, bc_names = hostdata.intern_strings([r[0] for r in table])
    up_id, up_names = hostdata.intern_strings([r[1] for r in table])
    size = np.asarray([int(r[3]) for r in table], dtype=np.int32)
    res = ctx.tag_rows(bc_id, up_id, size, len(bc_names), len(up_names), min_ratio)
    out = []
    for i, r in enumerate(table):
        out.append(tuple(r) + ("collision" if res["collision"][i] else "", "collision" if res["up_collision"][i] else "",
                               bool(res["usable"][i])))
    return out


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="pacybara_b200.tagging", description=__doc__.split("\n\n")[0])
    ap.add_argument("infile", help="clusters csv.gz file")
    ap.add_argument("--minRatio", type=float, default=0.666, help="minimum ccs ratio for cluster dominance")
    ap.add_argument("--out", default="clusters_tagged.csv.gz", help="output file name")
    args = ap.parse_args(argv)
    if not (0.5 <= args.minRatio <= 1):
        print("minRatio must be between 0.5 and 1", file=sys.stderr)
        return 1
    from .api import open_context_or_exit
    rows = canon.read_clusters_csv(args.infile)
    ctx = open_context_or_exit(0)
    tagged = tag_table(ctx, rows, args.minRatio)
    with gzip.open(args.out, "wt") as fh:
        fh.write("virtualBarcode,upBarcode,reads,size,geno,collision,upTagCollision,usable\n")
        for r in tagged:
            fh.write(",".join(str(x) for x in r[:7]) + "," + ("TRUE" if r[7] else "FALSE") + "\n")
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
