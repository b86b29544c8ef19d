"""Drop-in for src/pacybara_precluster.py: FASTQ on stdin -> FASTQ of unique barcodes on stdout.

Same contract as the reference (no arguments; header parse `@([^ ]+)`; output record
`@<rid|rid|...> size=<n>`, buckets in first-seen order, reads in arrival order, quality =
chr(min(93, sum of Phred) + 33); a read whose quality line starts with '@' is dropped, as the
reference's line state machine does, src/pacybara_precluster.py:47-63).  The bucketing itself
runs on the GPU through pcb_bucket; without a CUDA device the command fails.
"""
from __future__ import annotations

import sys

from . import hostdata


def main(argv=None) -> int:
    import io
    from . import ingest
    from .api import open_context_or_exit
    raw = sys.stdin.buffer.read()
    out = sys.stdout
    ctx = open_context_or_exit(0)
    try:                                   # device tokenizer (csrc/ingest.cu) ...
        reads = ingest.fastq_from_bytes(ctx, raw)
    except ingest.NeedsLineParser:         # ... or, for anything it only counts, the reference's line loop
        reads = hostdata.parse_fastq_for_precluster(io.TextIOWrapper(io.BytesIO(raw)))
    if len(reads.ids) == 0:
        ctx.close()
        return 0
    res = ctx.bucket(reads.seq, reads.qual, reads.length)
    res = {k: (v.cpu().numpy() if hasattr(v, "data_ptr") else v) for k, v in res.items()}
    hostdata.write_preclust_fastq(out, reads.ids, res["bucket_ptr"], res["bucket_reads"], res["bucket_seq"],
                                  res["bucket_len"], res["bucket_qsum"])
    out.flush()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
