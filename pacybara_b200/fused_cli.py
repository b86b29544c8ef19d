"""One-call variant of the clustering section: extracted barcodes + genotypes -> clusters.csv.gz.

    pacybara_b200_cluster [-u UPTAG.fastq.gz] [-j f] [-m i] [-d i] [-q i] [-o clusters.csv.gz] [--full-table]
                          BARCODES.fastq.gz GENOEXTRACT.csv.gz

Does what precluster -> (bowtie2 + calcEdits) -> cluster do together (src/pacybara.sh:425-489), with the same
options as pacybara_cluster.py and the same output file, through a single pcb_cluster_reads call: nothing but the
reads goes to the GPU and nothing but the cluster table comes back.  Intermediate files are not written; use the
three drop-ins in bin/ when bcPreclust.fastq.gz / editDistance.csv.gz are wanted.
"""
from __future__ import annotations

import sys

from . import cluster as cluster_cli
from . import hostdata, pipeline


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    full = "--full-table" in argv
    argv = [a for a in argv if a != "--full-table"]
    # same option parsing / validation as the drop-in; the first positional slot is unused here
    try:
        o = cluster_cli.parse_args(argv[:-2] + [argv[-2], argv[-1], argv[-2]] if len(argv) >= 2 else argv)
    except cluster_cli.Die:
        return 1
    if o["help"]:
        print(__doc__)
        return 0
    bc_file, geno_file = o["ed"], o["geno"]
    from .api import Context
    ctx = Context(0)
    from . import ingest
    ra, names = ingest.arrays_from_files(ctx, bc_file, geno_file, o["uptag"] or None, log=cluster_cli.log)
    bad = [r for r in names["ids"] if "-" in r]
    if bad:
        cluster_cli.log(f"ERROR: read id {bad[0]!r} contains '-'", file=sys.stderr)
        return 1
    cluster_cli.log(f"Clustering {ra.n_reads} reads on the GPU...")
    out = pipeline.cluster_reads(ctx, ra, max_diff=o["max_diff"], min_qual=o["min_qual"], min_jaccard=o["min_jaccard"],
                                 min_matches=o["min_matches"], full_table=full)
    table = pipeline.table_from_outputs(out, ra, names)
    cluster_cli.log(f" - {out['n_buckets']} barcode groups, {out['n_edges']} barcode pairs, {len(table)} rows")
    hostdata.write_clusters_csv(o["out"], table)
    cluster_cli.log("Done!")
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
