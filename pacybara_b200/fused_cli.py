"""One-call variant of the clustering section: extracted barcodes + genotypes -> clusters.csv.gz.

    pacybara_b200_cluster [-u UPTAG.fastq.gz] [-j f] [-m i] [-d i] [-q i] [-o clusters.csv.gz] [--full-table]
                          BARCODES.fastq.gz GENOEXTRACT.csv.gz

Does what precluster -> (bowtie2 + calcEdits) -> cluster do together (src/pacybara.sh:425-489), with the same
options as pacybara_cluster.py and the same output file, through a single pcb_cluster_reads call: nothing but the
reads goes to the GPU and nothing but the cluster table comes back.  Intermediate files are not written; use the
three drop-ins in bin/ when bcPreclust.fastq.gz / editDistance.csv.gz are wanted.

Several GPUs of one box: launch it under torchrun, one process per GPU --

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bin/pacybara_b200_cluster ...

every rank tokenizes the files on its own GPU, the pair search is sharded by query bucket, the merge by connected
component (pacybara_b200/multigpu.py), and rank 0 collects the ranks' rows and writes the one table.
"""
from __future__ import annotations

import sys

from . import cluster as cluster_cli
from . import hostdata, pipeline

USAGE = ("Usage: pacybara_b200_cluster [-u|--uptagBarcodeFile <FASTQFILE>] [-j|--minJaccard <FLOAT>] [-m|--minMatches <INT>]\n"
         "  [-d|--maxDiff <INT>] [-q|--minQual <INT>] [-o|--out <FILE>] [--full-table] [--help] <BARCODES.fastq.gz> <GENOEXTRACT.csv.gz>")


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if "--help" in argv or "-h" in argv:
        print(__doc__)
        return 0
    full = "--full-table" in argv
    argv = [a for a in argv if a != "--full-table"]
    # same option parsing / validation as the drop-in (its messages name the option at fault); the first
    # positional slot is unused here
    try:
        o = cluster_cli.parse_args(argv[:-2] + [argv[-2], argv[-1], argv[-2]] if len(argv) >= 2 else argv, usage=USAGE)
    except cluster_cli.Die:
        return 1
    bc_file, geno_file = o["ed"], o["geno"]
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    log = cluster_cli.log if rank == 0 else (lambda *a, **k: None)
    from .api import open_context_or_exit
    from . import ingest
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ctx = open_context_or_exit(local)
    try:
        ra, names = ingest.arrays_from_files(ctx, bc_file, geno_file, o["uptag"] or None, log=log)
        ids = names["ids"]
        bad = [ids.any_contains(b"-")] if hasattr(ids, "any_contains") else [r for r in ids if "-" in r][:1]
        if bad and bad[0] is not None:    # (the same files on every rank: every rank leaves here)
            log(f"ERROR: read id {bad[0]!r} contains '-'", file=sys.stderr)
            ctx.close()
            if world > 1:
                dist.destroy_process_group()
            return 1
        params = dict(max_diff=o["max_diff"], min_qual=o["min_qual"], min_jaccard=o["min_jaccard"], min_matches=o["min_matches"])
        if world == 1:
            log(f"Clustering {ra.n_reads} reads on the GPU...")
            out = pipeline.cluster_reads(ctx, ra, full_table=full, **params)
        else:
            log(f"Clustering {ra.n_reads} reads on {world} GPUs...")
            out = _cluster_sharded(ctx, ra, params, rank, world, full)
        if rank == 0:
            text = pipeline.table_text_from_outputs(ctx, out, ra, names)      # formatted on the device when the slices are there
            if text is not None:
                log(f" - {out['n_buckets']} barcode groups, {out['n_edges']} barcode pairs, {text.count(b"\n") - 1} rows")
                hostdata.write_clusters_text(o["out"], text)
            else:
                table = pipeline.table_from_outputs(out, ra, names)
                log(f" - {out['n_buckets']} barcode groups, {out['n_edges']} barcode pairs, {len(table)} rows")
                hostdata.write_clusters_csv(o["out"], table)
            log("Done!")
    except Exception:
        ctx.close()
        if world > 1:                     # no barrier on the way out of an error: the other ranks may not get there
            dist.destroy_process_group()
        raise
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _cluster_sharded(ctx, ra, params, rank: int, world: int, full: bool):
    """The sharded path; rank 0 gets the whole result as per-read host arrays, the others None."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from . import multigpu
    dev = f"cuda:{ctx.device}"
    arrs = {k: (v if hasattr(v, "data_ptr") else torch.from_numpy(np.ascontiguousarray(v)).to(dev))
            for k, v in vars(ra).items() if v is not None}            # line-parser route: host arrays
    part = multigpu.cluster_reads_sharded(ctx, arrs, params, rank, world, on_device=True, full_table=full, combine=False)
    shard = ctx.compact_shard(part, to_host=True)                      # my reads and rows only
    shard = {k: (np.array(v) if hasattr(v, "shape") else v) for k, v in shard.items()}
    parts = [None] * world
    dist.all_gather_object(parts, shard)
    if rank != 0:
        return None
    out = hostdata.merge_outputs_from_shards(parts, ra.n_reads)
    out.update(n_buckets=part["n_buckets"], n_edges=part["n_edges"])
    return out


if __name__ == "__main__":
    sys.exit(main())
