"""The fused hot path on arrays: what bench.py times and what the drop-in CLIs call.

reads (barcode, quality, genotype) -> buckets -> distance pairs -> clusters, one C-ABI call
(`pcb_cluster_reads`) on one GPU, or the sharded variant over torch.distributed (multigpu.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import hostdata
from .api import Context
from .canon import mut_sort_key


@dataclass
class ReadArrays:
    """Array form of (bcExtract_*.fastq.gz, genoExtract.csv.gz [, uptag fastq]) for R reads."""
    seq: np.ndarray          # uint8 [R, stride]
    qual: np.ndarray         # uint8 [R, stride]
    length: np.ndarray       # int32 [R]
    lexrank: np.ndarray      # int32 [R]
    geno_ptr: np.ndarray     # int32 [R+1]
    geno_mut: np.ndarray     # int32 [nnz]
    geno_q: np.ndarray       # int32 [nnz]
    mut_rank: np.ndarray     # int32 [M]
    up_id: Optional[np.ndarray] = None   # int32 [R]

    @property
    def n_reads(self) -> int:
        return int(self.length.shape[0])


def mut_rank_of(mut_strings: List[str]) -> np.ndarray:
    order = sorted(range(len(mut_strings)), key=lambda i: mut_sort_key(mut_strings[i]))
    rank = np.empty(len(order), dtype=np.int32)
    rank[np.asarray(order, dtype=np.int64)] = np.arange(len(order), dtype=np.int32)
    return rank


def arrays_from_library(lib, ctx: Optional[Context] = None, with_names: bool = True) -> Tuple[ReadArrays, Dict]:
    """ReadArrays of a synth.Library plus the string tables the writer needs.  For combo
    libraries the uptag ids come from bucketing the uptag strings (needs `ctx`).
    with_names=False skips building the R read-id strings (zero-padded ids sort like their index)."""
    if not with_names and lib.id_style == "padded":
        ids = None
        lexrank = np.arange(lib.n_reads, dtype=np.int32)
    else:
        ids = lib.read_ids()
        lexrank = hostdata.lexrank_of(ids)
    names = dict(ids=ids, mut_names=lib.mut_strings(), up_names=None)
    up_id = None
    if lib.uptag_len is not None:
        if ctx is None:
            raise ValueError("a combo library needs a Context to bucket its uptags")
        useq = lib.uptag_seq()
        b = ctx.bucket(useq, lib.qual, lib.uptag_len, want_qsum=False)
        up_id = b["bucket_of_read"].astype(np.int32)
        names["up_names"] = hostdata.matrix_to_strings(b["bucket_seq"], b["bucket_len"])
    ra = ReadArrays(seq=lib.seq, qual=lib.qual, length=lib.length.astype(np.int32),
                    lexrank=lexrank, geno_ptr=lib.geno_ptr.astype(np.int32),
                    geno_mut=lib.geno_mut.astype(np.int32), geno_q=lib.geno_q.astype(np.int32),
                    mut_rank=mut_rank_of(names["mut_names"]), up_id=up_id)
    return ra, names


def cluster_reads(ctx: Context, ra: ReadArrays, max_diff: int = 2, min_qual: int = 32, min_jaccard: float = 0.2,
                  min_matches: int = 1, full_table: bool = False, compact_rows: bool = False) -> Dict:
    return ctx.cluster_reads(ra.seq, ra.qual, ra.length, ra.lexrank, ra.up_id, ra.geno_ptr, ra.geno_mut, ra.geno_q,
                             ra.mut_rank, max_diff=max_diff, min_qual=min_qual, min_jaccard=min_jaccard,
                             min_matches=min_matches, full_table=full_table, compact_rows=compact_rows)


def table_from_outputs(out: Dict, ra: ReadArrays, names: Dict, consensus=None):
    """Rows (virtualBarcode, upBarcode, reads, size, geno) of a fused run (device results are brought to the host).
    consensus(list of barcode strings) -> string resolves the rows whose barcode vote is tied among >= 3 reads; the
    default is the reference's own `muscle` alignment (pacybara_cluster.py:484-511)."""
    host = lambda a: a.cpu().numpy() if hasattr(a, "data_ptr") else a
    out = {k: host(v) for k, v in out.items()}
    ra = ReadArrays(**{k: host(getattr(ra, k)) for k in ("seq", "qual", "length", "lexrank", "geno_ptr", "geno_mut", "geno_q",
                                                        "mut_rank", "up_id")})
    rows = hostdata.rows_from_merge_outputs(out)
    bor = np.asarray(out["bucket_of_read"])
    U = int(out["n_buckets"])
    first = np.full(U, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(first, bor, np.arange(bor.shape[0]))
    bc_names = hostdata.matrix_to_strings(np.asarray(ra.seq)[first], np.asarray(ra.length)[first])
    if consensus is None:
        from .cluster import alignment_consensus as consensus
    up_names = names["up_names"] if names["up_names"] is not None else bc_names
    up_of = ra.up_id if ra.up_id is not None else bor

    def on_alignment(i, members, b, u):
        if b == hostdata.NEEDS_ALIGNMENT:
            b = consensus([bc_names[bor[r]] for r in members])
        if u == hostdata.NEEDS_ALIGNMENT:
            u = consensus([up_names[up_of[r]] for r in members])
        return b, u

    return hostdata.rows_to_table(rows, list(names["ids"]), bc_names, names["up_names"], names["mut_names"], on_alignment=on_alignment)


def table_text_from_outputs(ctx: Context, out: Dict, ra: ReadArrays, names: Dict, consensus=None) -> Optional[bytes]:
    """The text of clusters.csv (header + rows) with the rows formatted on the device (tablefmt / pcb_concat_tokens), or None
    when that route is not available: no slice offsets (the line parsers loaded the files) -- table_from_outputs handles
    that.  consensus: as for table_from_outputs (the few rows with a tied barcode vote are resolved on the host and their
    strings handed to the formatter)."""
    sl = names.get("slices")
    if sl is None:
        return None
    from . import tablefmt
    host = lambda a: a.cpu().numpy() if hasattr(a, "data_ptr") else np.asarray(a)
    keys = ("read_root", "read_pos", "row_size", "row_key", "row_bc", "row_up", "row_call_off", "row_call_n", "call_mut")
    rows = hostdata.rows_from_merge_outputs({k: host(out[k]) for k in keys})
    as_slices = lambda t: tablefmt.Slices(np.frombuffer(t[0], dtype=np.uint8), np.asarray(t[1]), np.asarray(t[2]))
    bor = host(out["bucket_of_read"])
    bc = tablefmt.barcode_slices(host(ra.seq), host(ra.length), bor, int(out["n_buckets"]))
    ids = as_slices(sl["ids"])
    up = None
    if sl.get("up") is not None:
        up = as_slices(sl["up"])
        if sl["up"][0] is sl["ids"][0]:                       # the uptag file is the barcode file: one copy of its bytes
            up.text = ids.text
    # rows whose barcode / uptag vote is tied among >= 3 reads: the reference aligns the members' strings (:484-511)
    bc_text, up_text = {}, {}
    need = np.flatnonzero((rows.row_bc == hostdata.NEEDS_ALIGNMENT) | (rows.row_up == hostdata.NEEDS_ALIGNMENT))
    if need.size:
        if consensus is None:
            from .cluster import alignment_consensus as consensus
        text_of = lambda s, i: s.text[int(s.off[i]): int(s.off[i]) + int(s.len[i])].tobytes().decode("ascii")
        up_of = host(ra.up_id) if ra.up_id is not None else bor
        up_src = up if up is not None else bc
        for i in need.tolist():
            members = rows.row_members[rows.row_ptr[i]: rows.row_ptr[i + 1]]
            if rows.row_bc[i] == hostdata.NEEDS_ALIGNMENT:
                bc_text[i] = consensus([text_of(bc, bor[r]) for r in members])
            if rows.row_up[i] == hostdata.NEEDS_ALIGNMENT:
                up_text[i] = consensus([text_of(up_src, up_of[r]) for r in members])
    return tablefmt.table_bytes(ctx, rows, ids, bc, up, as_slices(sl["muts"]), bc_text, up_text)


def arrays_from_text(fastq_stream, geno_stream, uptag_stream=None, ctx: Optional[Context] = None
                     ) -> Tuple[ReadArrays, Dict]:
    """ReadArrays from the text inputs of the path (extracted-barcode FASTQ, genoExtract CSV and
    optionally the uptag FASTQ), parsed with the reference's own quirks (hostdata)."""
    reads = hostdata.parse_fastq_for_precluster(fastq_stream)
    index = {rid: i for i, rid in enumerate(reads.ids)}
    if len(index) != len(reads.ids):
        raise hostdata.FormatError("duplicate read id in the barcode FASTQ")
    geno = hostdata.parse_geno(geno_stream, index, len(reads.ids))
    up_id, up_names = None, None
    if uptag_stream is not None:
        tags = hostdata.parse_uptags(uptag_stream)
        if len(tags) > 0:
            up_id, up_names = hostdata.intern_strings([tags.get(rid) for rid in reads.ids])
    ra = ReadArrays(seq=reads.seq, qual=reads.qual, length=reads.length, lexrank=hostdata.lexrank_of(reads.ids),
                    geno_ptr=geno.geno_ptr, geno_mut=geno.geno_mut, geno_q=geno.geno_q, mut_rank=geno.mut_rank(),
                    up_id=up_id)
    return ra, dict(ids=reads.ids, mut_names=geno.mut_strings, up_names=up_names)
