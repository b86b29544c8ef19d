"""Drop-in for src/pacybara_cluster.py: same options, positionals, validation, exit codes and
output file (src/pacybara_cluster.py:34-55, 68-143, 546-567); the clustering itself (seed step,
greedy merge, consensus) runs on the GPU through pcb_merge.

Differences, all deliberate:
  * read ids containing '-' make the reference die with a ValueError half way (:418); they are
    refused up front with exit code 1;
  * singleton rows come out in read order (the reference's order is a PYTHONHASHSEED-dependent
    set iteration, :547,560);
  * the muscle fallback (:484-511) is only run when a >=3-read cluster has tied barcode counts,
    exactly as in the reference, and needs `muscle` on PATH in that case.
"""
from __future__ import annotations

import collections
import getopt
import os
import subprocess
import sys
import tempfile
from datetime import datetime

from . import hostdata

USAGE = """
pacybara_cluster.py (pacybara_b200 drop-in)

Barcode Clustering for Pacybara

Usage: pacybara_cluster.py [-u|--uptagBarcodeFile <FASTQFILE>]
  [-j|--minJaccard <FLOAT>] [-m|--minMatches <INT>] [-d|--maxDiff <INT>] [-q|--minQual <INT>]
  [-o|--out <FILE>] [--help] <EDFILE> <GENOFILE> <PRECLUSTFILE>

-u|--uptagBarcodeFile : optional uptag barcode FASTQ file
-j|--minJaccard       : minimum jaccard coefficient to accept cluster
-m|--minMatches       : minimum number of shared variants in the seed clustering
-d|--maxDiff          : maximum edit distance to consider
-q|--minQual          : minimum PHRED quality to accept variant call
-o|--out              : output file. defaults to clusters.csv.gz
--help                : show this help text
EDFILE                : The edit distance file
GENOFILE              : The bcGenos file
PRECLUSTFILE          : The bcPreClust file
"""


def log(msg, file=sys.stdout):
    print(datetime.now().strftime("%Y/%m/%d %H:%M:%S") + " : " + str(msg), file=file)


class Die(Exception):
    pass


_usage_text = None        # a caller with other positionals (fused_cli) shows its own usage


def usage_and_die(err):
    log(err, file=sys.stderr)
    print(USAGE if _usage_text is None else _usage_text)
    raise Die()


def parse_args(argv, usage=None):
    global _usage_text
    _usage_text = usage
    try:
        opts, args = getopt.getopt(argv, "u:j:m:d:q:o:",
                                   ["uptagBarcodeFile=", "minJaccard=", "minMatches=", "maxDiff=", "minQual=", "out=", "help"])
    except getopt.GetoptError as err:
        usage_and_die(err)
    o = dict(uptag=None, min_jaccard=0.2, min_matches=1, max_diff=2, min_qual=32, out="clusters.csv.gz", help=False)
    for k, a in opts:
        if k == "--help":
            o["help"] = True
        elif k in ("-u", "--uptagBarcodeFile"):
            o["uptag"] = a
        elif k in ("-j", "--minJaccard"):
            try:
                o["min_jaccard"] = float(a)
            except ValueError:
                usage_and_die("ERROR: minJaccard must be a number")
        elif k in ("-m", "--minMatches"):
            try:
                o["min_matches"] = int(a)
            except ValueError:
                usage_and_die("ERROR: minMatches must be an integer number")
        elif k in ("-d", "--maxDiff"):
            try:
                o["max_diff"] = int(a)
            except ValueError:
                usage_and_die("ERROR: maxDiff must be an integer number")
        elif k in ("-q", "--minQual"):
            try:
                o["min_qual"] = int(a)
            except ValueError:
                usage_and_die("ERROR: minQual must be an integer number")
        elif k in ("-o", "--out"):
            o["out"] = a
    if o["help"]:
        return o
    if len(args) < 3:
        usage_and_die("ERROR: Missing required arguments!")
    elif len(args) > 3:
        usage_and_die("Unrecognized additional arguments!")
    o["ed"], o["geno"], o["preclust"] = args
    for key in ("ed", "geno", "preclust", "uptag"):
        if o[key] is not None and not os.path.exists(o[key]):
            usage_and_die("ERROR: The file " + o[key] + " does not exist")
    if o["min_jaccard"] > 1 or o["min_jaccard"] <= 0:
        usage_and_die("ERROR: minJaccard must be between 0 and 1")
    if o["min_matches"] < 1:
        usage_and_die("ERROR: minMatches must be at least 1")
    if o["max_diff"] < 0 or o["max_diff"] > 3:
        usage_and_die("ERROR: maxDiff must be between 0 and 3")
    if o["min_qual"] < 1:
        usage_and_die("ERROR: minQual must be a positive integer")
    return o


def alignment_consensus(bcs):
    """src/pacybara_cluster.py:484-511, verbatim behaviour: muscle, column majority, gaps dropped."""
    with tempfile.NamedTemporaryFile(mode="w", suffix=".fa") as fasta, \
            tempfile.NamedTemporaryFile(mode="r", suffix=".aln") as aln, \
            tempfile.NamedTemporaryFile(mode="r", suffix=".log") as logf:
        for i, bc in enumerate(bcs):
            fasta.write(f">{i}\n{bc}\n")
        fasta.flush()
        ret = subprocess.call(f"muscle -align {fasta.name} -output {aln.name} >{logf.name} 2>&1", shell=True)
        if ret != 0:
            log("Muscle aligment failed. Error message below:")
            log("".join(logf.readlines()))
            raise Exception("Alignment failed!")
        alignment = [line.strip() for line in aln if not line.startswith(">")]
    cons = []
    for i in range(len(alignment[0])):
        counts = collections.Counter(a[i] for a in alignment)
        top = max(counts.values())
        cons.append([k for k, v in counts.items() if v == top][0])
    return "".join(b for b in cons if b != "-")


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    try:
        o = parse_args(argv)
    except Die:
        return 1
    if o["help"]:
        print(USAGE)
        return 0
    from .api import open_context_or_exit
    log("Loading data...")
    params = dict(max_diff=o["max_diff"], min_qual=o["min_qual"], min_jaccard=o["min_jaccard"], min_matches=o["min_matches"])
    with hostdata.open_text(o["preclust"]) as fh:
        pc = hostdata.parse_preclust(fh)
    bad = [r for r in pc.ids if "-" in r]
    if bad:                               # before any work (the reference dies at :418 once such a pair comes up)
        log(f"ERROR: read id {bad[0]!r} contains '-': the pair-id scheme of pacybara_cluster.py (:152-157,418) "
            "cannot represent it", file=sys.stderr)
        return 1
    ctx = open_context_or_exit(0)
    job, slices = None, None
    try:
        # genotypes and uptags through the device tokenizers (the line loaders cost a Python iteration per read)
        from . import ingest
        try:
            job, slices = ingest.cluster_job_from_files(ctx, o["ed"], o["geno"], o["preclust"], o["uptag"], pc=pc, **params)
        except ingest.NeedsLineParser as e:
            log(f" - tokenizer declined ({e}); using the line parser")
    except BaseException:
        ctx.close()
        raise
    if job is None:
        with hostdata.open_text(o["ed"]) as ed, hostdata.open_text(o["geno"]) as ge, hostdata.open_text(o["preclust"]) as pc:
            up = hostdata.open_text(o["uptag"]) if o["uptag"] is not None else None
            try:
                job = hostdata.load_cluster_job(ed, ge, pc, up, **params)
            except BaseException:
                ctx.close()
                raise
            finally:
                if up is not None:
                    up.close()
    p = job.problem
    pairs = hostdata.dedupe_ed_rows(job.ed_q, job.ed_r, job.ed_d, p.n_buckets, p.max_diff)
    log(f"Clustering {p.n_reads} reads in {p.n_buckets} barcode groups, {len(pairs.q)} barcode pairs on the GPU...")
    out = ctx.merge(p, pairs)
    rows = hostdata.rows_from_merge_outputs(out)
    log(f" - {ctx.stat('components')} components, {ctx.stat('tests')} cluster tests, {ctx.stat('links')} links")

    def on_alignment(i, members, b, u):
        if b == hostdata.NEEDS_ALIGNMENT:
            b = alignment_consensus([job.bc_names[job.problem.bc_id[r]] for r in members])
        if u == hostdata.NEEDS_ALIGNMENT:
            names = job.up_names if job.up_names is not None else job.bc_names
            ids = job.problem.up_id if job.problem.up_id is not None else job.problem.bc_id
            u = alignment_consensus([names[ids[r]] for r in members])
        return b, u

    log(f"Writing output to file {o['out']}")
    if slices is not None:
        # the rows as a ragged concatenation of slices of the input texts, on the device (tablefmt / pcb_concat_tokens)
        import numpy as np
        from . import tablefmt
        mk = lambda t: tablefmt.Slices(np.frombuffer(t[0], dtype=np.uint8), np.asarray(t[1]), np.asarray(t[2]))
        bc_text, up_text = {}, {}
        for i in np.flatnonzero((rows.row_bc == hostdata.NEEDS_ALIGNMENT) | (rows.row_up == hostdata.NEEDS_ALIGNMENT)).tolist():
            b, u = on_alignment(i, rows.row_members[rows.row_ptr[i]: rows.row_ptr[i + 1]], int(rows.row_bc[i]), int(rows.row_up[i]))
            if isinstance(b, str):
                bc_text[i] = b
            if isinstance(u, str):
                up_text[i] = u
        text = tablefmt.table_bytes(ctx, rows, mk(slices["ids"]), mk(slices["bc"]), mk(slices["up"]) if slices["up"] is not None else None,
                                    mk(slices["muts"]), bc_text, up_text)
        hostdata.write_clusters_text(o["out"], text)
    else:
        table = hostdata.rows_to_table(rows, job.ids, job.bc_names, job.up_names, job.mut_names, on_alignment=on_alignment)
        hostdata.write_clusters_csv(o["out"], table)
    ctx.close()
    log("Done!")
    return 0


if __name__ == "__main__":
    sys.exit(main())
