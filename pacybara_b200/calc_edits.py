"""Drop-in for src/pacybara_calcEdits.R (and, with it, for the bowtie2 all-vs-all step).

    pacybara_calcEdits.R <samFile> <preclustFASTQ> [--maxErr N] [--output F] [--chunkSize N]

Same positional arguments and options as the reference (src/pacybara_calcEdits.R:79-89) so that
src/pacybara.sh:466-468 can call it unchanged.  The SAM file is NOT read: the distance of every
pair of bucket barcodes is computed directly (unit-cost Levenshtein, pcb_edit_pairs) and every pair
with dist <= maxErr is written, in both directions, as `"query","ref",NA,NA,dist` rows sorted by
(query bucket, ref bucket).  Only fields 1, 2 and 5 are consumed downstream
(src/pacybara_cluster.py:257-260, src/pacybara.sh:470-471).
"""
from __future__ import annotations

import argparse
import gzip
import sys

from . import hostdata


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="calcEdits.R", description="calculate edit distance between barcode clusters")
    ap.add_argument("samFile", help="sam.gz file with cluster alignments (ignored: distances are computed directly)")
    ap.add_argument("preclustFASTQ", help="fastq.gz file of cluster consensus seqs")
    ap.add_argument("--output", default="editDistance.csv.gz", help="output file")
    ap.add_argument("--maxErr", type=int, default=3, help="Maximum allowed number of errors in barcode")
    ap.add_argument("--chunkSize", type=int, default=1000, help="accepted for compatibility; unused")
    args = ap.parse_args(argv)
    from .api import open_context_or_exit
    with hostdata.open_text(args.preclustFASTQ) as fh:
        pc = hostdata.parse_preclust(fh)
    mat, lens = hostdata.strings_to_matrix([s.encode("ascii") for s in pc.seqs])
    ctx = open_context_or_exit(0)
    ei, ej, ed = ctx.edit_pairs(mat, lens, args.maxErr) if pc.n_buckets > 1 else ([], [], [])
    import numpy as np
    ei, ej, ed = (np.asarray(x, dtype=np.int32) for x in (ei, ej, ed))
    opener = gzip.open if args.output.endswith(".gz") else open
    with opener(args.output, "wt") as out:
        n = hostdata.write_ed(out, pc.titles, ei, ej, ed)
    print(f"{pc.n_buckets} barcodes, {n} rows with dist <= {args.maxErr} "
          f"({ctx.stat('pairs_scored')} candidate pairs scored on the GPU)")
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
