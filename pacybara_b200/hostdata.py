"""Host-side data model: text formats of the hot path <-> the flat arrays the C-ABI takes.

Every parser mirrors the reference's own loader line by line (same splitting, same
"last one wins" dictionary semantics), because those quirks are part of the contract:

* FASTQ for bucketing        src/pacybara_precluster.py:47-63
* bcPreclust.fastq.gz        src/pacybara_cluster.py:268-298
* genoExtract.csv.gz         src/pacybara_cluster.py:300-313
* uptag FASTQ                src/pacybara_cluster.py:315-331
* editDistance.csv.gz        src/pacybara_cluster.py:250-266 (reader),
                             src/pacybara_calcEdits.R:123,164-171 (writer)
* clusters.csv.gz            src/pacybara_cluster.py:546-567

This module is pure host logic (numpy); it never computes distances or clusters.
"""
from __future__ import annotations

import gzip
import io
import re
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .canon import mut_sort_key

MAX_BARCODE = 64          # 2-bit planes in at most two 64-bit words; longer barcodes are kept as opaque strings
_RID = re.compile(r"@([^ ]+)")


class FormatError(ValueError):
    pass


def open_text(path: str, mode: str = "rt"):
    if path == "-":
        import sys
        return sys.stdin if "r" in mode else sys.stdout
    return gzip.open(path, mode) if path.endswith(".gz") else open(path, mode)


# --------------------------------------------------------------------------- strings <-> matrices

def strings_to_matrix(items: Sequence[bytes], stride: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
    """List of byte strings -> (uint8 [n, stride] zero padded, int32 lengths)."""
    n = len(items)
    lens = np.fromiter((len(s) for s in items), dtype=np.int32, count=n)
    lmax = int(lens.max()) if n else 0
    if stride is None:
        stride = max(1, lmax)
    if lmax > stride:
        raise FormatError(f"string of length {lmax} does not fit stride {stride}")
    mat = np.zeros((n, stride), dtype=np.uint8)
    if n:
        flat = np.frombuffer(b"".join(items), dtype=np.uint8)
        starts = np.zeros(n, dtype=np.int64)
        np.cumsum(lens[:-1], out=starts[1:])
        row = np.repeat(np.arange(n), lens)
        col = np.arange(flat.shape[0]) - np.repeat(starts, lens)
        mat[row, col] = flat
    return mat, lens


def matrix_to_strings(mat: np.ndarray, lens: np.ndarray) -> List[str]:
    raw = mat.tobytes()
    stride = mat.shape[1]
    return [raw[i * stride: i * stride + int(lens[i])].decode("ascii") for i in range(mat.shape[0])]


# --------------------------------------------------------------------------- a1 input: FASTQ for bucketing

@dataclass
class FastqReads:
    ids: List[str]
    seq: np.ndarray      # uint8 [R, stride]
    qual: np.ndarray     # uint8 [R, stride]
    length: np.ndarray   # int32 [R]


def parse_fastq_for_precluster(stream: Iterable[str]) -> FastqReads:
    """The reference's line state machine (pacybara_precluster.py:47-63), including its
    quirk that a quality line starting with '@' is taken for a header, which drops that read."""
    ids: List[str] = []
    seqs: List[bytes] = []
    quals: List[bytes] = []
    i = 0
    rid = seq = ""
    for line in stream:
        line = line.rstrip()
        if line.startswith("@"):
            i = 0
            m = _RID.match(line)
            if not m:
                raise Exception("Unable to extract read id!")
            rid = m.group(1)
        i += 1
        if i == 2:
            seq = line
        elif i == 4:
            if len(line) != len(seq):
                raise FormatError(f"read {rid}: quality length {len(line)} != sequence length {len(seq)}")
            ids.append(rid)
            seqs.append(seq.encode("ascii"))
            quals.append(line.encode("ascii"))
    smat, lens = strings_to_matrix(seqs)
    qmat, _ = strings_to_matrix(quals, smat.shape[1])
    return FastqReads(ids, smat, qmat, lens)


def write_preclust_fastq(out, ids: Sequence[str], bucket_ptr: np.ndarray, bucket_reads: np.ndarray,
                         bucket_seq: np.ndarray, bucket_len: np.ndarray, bucket_qsum: np.ndarray) -> None:
    """pacybara_precluster.py:68-76: title = '|'-joined ids, size tag, chr(sum+33)."""
    seqs = matrix_to_strings(bucket_seq, bucket_len)
    quals = matrix_to_strings((bucket_qsum + 33).astype(np.uint8), bucket_len)
    w = out.write
    for b in range(len(seqs)):
        lo, hi = int(bucket_ptr[b]), int(bucket_ptr[b + 1])
        title = "|".join(ids[r] for r in bucket_reads[lo:hi])
        w(f"@{title} size={hi - lo}\n{seqs[b]}\n+\n{quals[b]}\n")


# --------------------------------------------------------------------------- cluster-stage inputs

@dataclass
class Preclust:
    ids: List[str]                 # read ids, bucket by bucket (index = read number)
    bucket_ptr: np.ndarray         # int32 [U+1]
    bucket_reads: np.ndarray       # int32 [R]  (identity here: reads are numbered in file order)
    titles: List[str]
    seqs: List[str]                # one per bucket

    @property
    def n_reads(self) -> int:
        return len(self.ids)

    @property
    def n_buckets(self) -> int:
        return len(self.titles)


def parse_preclust(stream: Iterable[str]) -> Preclust:
    """pacybara_cluster.py:270-298."""
    ids: List[str] = []
    ptr = [0]
    titles: List[str] = []
    seqs: List[str] = []
    lcount = 0
    have = False
    for line in stream:
        line = line.rstrip()
        lcount += 1
        if line.startswith("@"):
            lcount = 1
            title = line[1:].split(" ")[0]
            rids = title.split("|")
            ids.extend(rids)
            ptr.append(len(ids))
            titles.append(title)
            seqs.append("")
            have = True
        elif lcount == 2 and have:
            seqs[-1] = line
    R = len(ids)
    return Preclust(ids, np.asarray(ptr, dtype=np.int32), np.arange(R, dtype=np.int32), titles, seqs)


@dataclass
class Genotypes:
    geno_ptr: np.ndarray     # int32 [R+1]
    geno_mut: np.ndarray     # int32 [nnz]
    geno_q: np.ndarray       # int32 [nnz]
    mut_strings: List[str]

    def mut_rank(self) -> np.ndarray:
        """Rank of every mutation in the (digit key, string) order (pacybara_cluster.py:475)."""
        order = sorted(range(len(self.mut_strings)), key=lambda i: mut_sort_key(self.mut_strings[i]))
        rank = np.empty(len(order), dtype=np.int32)
        rank[np.asarray(order, dtype=np.int64)] = np.arange(len(order), dtype=np.int32)
        return rank


def parse_geno(stream: Iterable[str], read_index: Dict[str, int], n_reads: int) -> Genotypes:
    """pacybara_cluster.py:302-313.  Reads absent from `read_index` are ignored (they are
    never looked up); a bucketed read absent from the file is the reference's KeyError."""
    per_read: List[Optional[Dict[int, int]]] = [None] * n_reads
    mut_id: Dict[str, int] = {}
    for line in stream:
        line = line.rstrip()
        fields = line.split(",")
        rid = fields[0].strip('"')
        idx = read_index.get(rid)
        geno = fields[1].strip('"')
        gd: Dict[int, int] = {}
        if geno != "=":
            for mut in geno.split(";"):
                m, q = mut.split(":")
                k = mut_id.get(m)
                if k is None:
                    k = mut_id[m] = len(mut_id)
                gd[k] = int(q)
        if idx is not None:
            per_read[idx] = gd
    missing = [i for i, g in enumerate(per_read) if g is None]
    if missing:
        raise KeyError(f"{len(missing)} bucketed read(s) have no genotype record (first index {missing[0]})")
    counts = np.fromiter((len(g) for g in per_read), dtype=np.int64, count=n_reads)
    ptr = np.zeros(n_reads + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    nnz = int(ptr[-1])
    muts = np.fromiter((k for g in per_read for k in g.keys()), dtype=np.int32, count=nnz)
    qs = np.fromiter((v for g in per_read for v in g.values()), dtype=np.int32, count=nnz)
    names = [None] * len(mut_id)
    for s, k in mut_id.items():
        names[k] = s
    return Genotypes(ptr.astype(np.int32), muts, qs, names)


def parse_uptags(stream: Iterable[str]) -> Dict[str, str]:
    """pacybara_cluster.py:318-331."""
    out: Dict[str, str] = {}
    lcount = 0
    rid = ""
    for line in stream:
        line = line.rstrip()
        lcount += 1
        if line.startswith("@"):
            lcount = 1
            rid = line[1:].split(" ")[0]
        elif lcount == 2 and rid != "":
            out[rid] = line
    return out


def intern_strings(strings: Sequence[Optional[str]]) -> Tuple[np.ndarray, List[str]]:
    """ids per item (-1 for None) and the table of distinct strings, first-seen order."""
    table: Dict[str, int] = {}
    ids = np.empty(len(strings), dtype=np.int32)
    for i, s in enumerate(strings):
        if s is None:
            ids[i] = -1
            continue
        k = table.get(s)
        if k is None:
            k = table[s] = len(table)
        ids[i] = k
    names = [None] * len(table)
    for s, k in table.items():
        names[k] = s
    return ids, names


def lexrank_of(ids: Sequence[str]) -> np.ndarray:
    """Rank of each read id in Python string order (makePID compares the strings, :152-157)."""
    arr = np.asarray(ids, dtype=object)
    order = np.argsort(arr, kind="stable")
    rank = np.empty(len(ids), dtype=np.int32)
    rank[order] = np.arange(len(ids), dtype=np.int32)
    return rank


# --------------------------------------------------------------------------- edit-distance table

ED_HEADER = '"query","ref","mismatch","indel","dist"'


def parse_ed(stream: Iterable[str], title_index: Dict[str, int]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """pacybara_cluster.py:252-266: header skipped, fields 0,1,4 used.  Titles must be
    bucket titles of the accompanying bcPreclust file."""
    q: List[int] = []
    r: List[int] = []
    d: List[int] = []
    first = True
    for line in stream:
        if first:
            first = False
            continue
        line = line.rstrip()
        if not line:
            continue
        f = line.split(",")
        try:
            q.append(title_index[f[0].strip('"')])
            r.append(title_index[f[1].strip('"')])
        except KeyError as e:
            raise FormatError(f"edit-distance row names a bucket that is not in the preclust file: {e}") from None
        d.append(int(f[4]))
    return (np.asarray(q, dtype=np.int32), np.asarray(r, dtype=np.int32), np.asarray(d, dtype=np.int32))


def write_ed(out, titles: Sequence[str], ei: np.ndarray, ej: np.ndarray, ed: np.ndarray) -> int:
    """Canonical table: every unordered pair (i<j, d) becomes the two rows (i,j) and (j,i);
    rows sorted by (query bucket, ref bucket).  Columns as pacybara_calcEdits.R:123,164-171;
    the mismatch/indel split is not defined for a Levenshtein distance and is written NA
    (only fields 0, 1 and 4 are read downstream, pacybara_cluster.py:257-260)."""
    q = np.concatenate([ei, ej]).astype(np.int64)
    r = np.concatenate([ej, ei]).astype(np.int64)
    d = np.concatenate([ed, ed])
    order = np.lexsort((r, q))
    out.write(ED_HEADER + "\n")
    w = out.write
    for k in order:
        w(f"\"{titles[q[k]]}\",\"{titles[r[k]]}\",NA,NA,{int(d[k])}\n")
    return int(order.shape[0])


def canonical_rows(ei: np.ndarray, ej: np.ndarray, ed: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """The (q, r, d) rows write_ed() emits, as arrays."""
    q = np.concatenate([ei, ej]).astype(np.int32)
    r = np.concatenate([ej, ei]).astype(np.int32)
    d = np.concatenate([ed, ed]).astype(np.int32)
    order = np.lexsort((r, q))
    return q[order], r[order], d[order]


@dataclass
class PairList:
    """Unique unordered bucket pairs in the order the reference's main loop meets them.

    For each pair: (q, r) = direction of the row that first put it into the pass it is
    processed in, d = the distance the lookup table ends up holding (last row wins,
    pacybara_cluster.py:266), ed = the pass that processes it (0: lookup only)."""
    q: np.ndarray
    r: np.ndarray
    d: np.ndarray
    ed: np.ndarray


def dedupe_ed_rows(q: np.ndarray, r: np.ndarray, d: np.ndarray, n_buckets: int, max_diff: int) -> PairList:
    """Collapse raw rows (both directions, duplicates, any order) into a PairList.

    Mirrors the dictionary semantics of pacybara_cluster.py:261-266 and :413-419: rows with
    dist > 2*maxDiff are ignored; a pair is visited in the first pass ed in 1..maxDiff for which
    it has a row, at the position of its first such row (dict insertion order), and is `known`
    afterwards; edistLookup keeps the distance of the pair's last row."""
    q = np.asarray(q, dtype=np.int64)
    r = np.asarray(r, dtype=np.int64)
    d = np.asarray(d, dtype=np.int64)
    keep = (d >= 0) & (d <= 2 * max_diff) & (q != r)
    rows = np.nonzero(keep)[0]
    q, r, d = q[rows], r[rows], d[rows]
    if q.size == 0:
        z = np.zeros(0, dtype=np.int32)
        return PairList(z, z.copy(), z.copy(), z.copy())
    key = np.minimum(q, r) * np.int64(n_buckets) + np.maximum(q, r)
    _, inv = np.unique(key, return_inverse=True)
    n = int(inv.max()) + 1
    big = np.iinfo(np.int64).max
    look = np.zeros(n, dtype=np.int64)
    look[inv] = d                                    # repeated index: the last assignment stays
    ped = np.full(n, big, dtype=np.int64)
    m = (d >= 1) & (d <= max_diff)
    np.minimum.at(ped, inv[m], d[m])
    first = np.full(n, big, dtype=np.int64)
    sel = m & (d == ped[inv])
    np.minimum.at(first, inv[sel], np.nonzero(sel)[0])
    anyrow = np.full(n, big, dtype=np.int64)
    np.minimum.at(anyrow, inv, np.arange(q.size))
    has = ped != big
    rep = np.where(has, first, anyrow)               # representative row of each pair
    ped = np.where(has, ped, 0)
    sort_ed = np.where(has, ped, max_diff + 1)        # lookup-only pairs go last
    order = np.lexsort((rep, sort_ed))
    return PairList(q[rep][order].astype(np.int32), r[rep][order].astype(np.int32),
                    look[order].astype(np.int32), ped[order].astype(np.int32))


def pairs_from_edges(ei: np.ndarray, ej: np.ndarray, ed: np.ndarray, max_diff: int) -> PairList:
    """PairList of the canonical table (rows sorted by (q, r), both directions): the first row
    of pair {i<j} is (q=i, r=j), so the visiting order is (i, j) order within each pass."""
    ei = np.asarray(ei, dtype=np.int64)
    ej = np.asarray(ej, dtype=np.int64)
    ed = np.asarray(ed, dtype=np.int64)
    keep = ed <= 2 * max_diff
    ei, ej, ed = ei[keep], ej[keep], ed[keep]
    ped = np.where((ed >= 1) & (ed <= max_diff), ed, 0)
    order = np.lexsort((ej, ei, np.where(ped > 0, ped, max_diff + 1)))
    return PairList(ei[order].astype(np.int32), ej[order].astype(np.int32),
                    ed[order].astype(np.int32), ped[order].astype(np.int32))


# --------------------------------------------------------------------------- the merge problem / result

@dataclass
class MergeProblem:
    n_reads: int
    n_buckets: int
    bucket_ptr: np.ndarray      # int32 [U+1]
    bucket_reads: np.ndarray    # int32 [R], arrival order inside each bucket
    lexrank: np.ndarray         # int32 [R]
    bc_id: np.ndarray           # int32 [R] id of the read's barcode string
    up_id: Optional[np.ndarray]  # int32 [R] id of the read's uptag string, -1 missing; None: no uptag file
    geno_ptr: np.ndarray        # int32 [R+1]
    geno_mut: np.ndarray        # int32 [nnz]
    geno_q: np.ndarray          # int32 [nnz]
    mut_rank: np.ndarray        # int32 [M]
    max_diff: int = 2
    min_qual: int = 32
    min_jaccard: float = 0.2
    min_matches: int = 1


@dataclass
class ClusterRows:
    """Cluster table as arrays: clusters first, singletons after (pacybara_cluster.py:552-567)."""
    row_ptr: np.ndarray         # int32 [n_rows+1]
    row_members: np.ndarray     # int32 [R] read numbers in list order
    row_bc: np.ndarray          # int32 [n_rows]; -2: the reference would run muscle (:530)
    row_up: np.ndarray          # int32 [n_rows]
    call_ptr: np.ndarray        # int32 [n_rows+1]
    call_mut: np.ndarray        # int32 [ncalls]
    n_clusters: int = 0
    stats: Dict = field(default_factory=dict)

    @property
    def n_rows(self) -> int:
        return int(self.row_ptr.shape[0]) - 1


NEEDS_ALIGNMENT = -2


def rows_to_table(rows: ClusterRows, ids: Sequence[str], bc_names: Sequence[str],
                  up_names: Optional[Sequence[str]], mut_names: Sequence[str],
                  on_alignment=None) -> List[Tuple[str, str, str, int, str]]:
    """(virtualBarcode, upBarcode, reads, size, geno) per row, as the reference writes them."""
    up_names = bc_names if up_names is None else up_names
    # gather once, slice per row: the id and mutation strings in member / call order
    member_ids = np.asarray(ids, dtype=object)[rows.row_members].tolist() if len(ids) else []
    call_names = np.asarray(mut_names, dtype=object)[rows.call_mut].tolist() if len(mut_names) else []
    row_ptr, call_ptr = rows.row_ptr.tolist(), rows.call_ptr.tolist()
    row_bc, row_up = rows.row_bc.tolist(), rows.row_up.tolist()
    out = []
    import gc
    gc_was_on = gc.isenabled()
    gc.disable()                 # millions of live strings: the cyclic collector would rescan them for every batch of rows
    try:
        _rows_to_table_loop(rows, out, row_ptr, call_ptr, row_bc, row_up, member_ids, call_names, bc_names, up_names, on_alignment)
    finally:
        if gc_was_on:
            gc.enable()
    return out


def _rows_to_table_loop(rows, out, row_ptr, call_ptr, row_bc, row_up, member_ids, call_names, bc_names, up_names, on_alignment):
    for i in range(rows.n_rows):
        lo, hi = row_ptr[i], row_ptr[i + 1]
        b, u = row_bc[i], row_up[i]
        if b == NEEDS_ALIGNMENT or u == NEEDS_ALIGNMENT:
            if on_alignment is None:
                raise RuntimeError("a cluster of >= 3 reads has tied barcode counts: the reference resolves "
                                   "this with an external `muscle` alignment (pacybara_cluster.py:484-511,530)")
            b, u = on_alignment(i, rows.row_members[lo:hi], b, u)
        vbc = b if isinstance(b, str) else bc_names[b]
        ubc = u if isinstance(u, str) else ("None" if u < 0 else up_names[u])
        clo, chi = call_ptr[i], call_ptr[i + 1]
        geno = ";".join(call_names[clo:chi]) if chi > clo else "="
        out.append((vbc, ubc, "|".join(member_ids[lo:hi]), hi - lo, geno))


def gzip_members(data: bytes, chunk: Optional[int] = None, level: int = 6) -> List[bytes]:
    """`data` as a sequence of gzip members compressed concurrently (zlib releases the GIL); their
    concatenation is one valid gzip file (RFC 1952 2.2), which gzip.open, zcat and R's gzfile read as a whole.
    chunk: bytes per member; default: two members per host core, at least 1 MiB each."""
    import os
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    if chunk is None:
        chunk = max(1 << 20, -(-len(data) // (2 * (os.cpu_count() or 1))))

    def one(i):
        c = zlib.compressobj(level, zlib.DEFLATED, 31)
        return c.compress(data[i:i + chunk]) + c.flush()

    starts = list(range(0, max(len(data), 1), chunk))
    if len(starts) == 1:
        return [one(0)]
    import os
    with ThreadPoolExecutor(max_workers=min(len(starts), os.cpu_count() or 1)) as ex:
        return list(ex.map(one, starts))


def write_clusters_csv(path: str, table: Sequence[Tuple[str, str, str, int, str]]) -> None:
    """pacybara_cluster.py:549-567: gzip, unquoted, fixed header."""
    buf = io.StringIO()
    buf.write("virtualBarcode,upBarcode,reads,size,geno\n")
    for vbc, ubc, reads, size, geno in table:
        buf.write(f"{vbc},{ubc},{reads},{size},{geno}\n")
    with open(path, "wb") as fh:
        for member in gzip_members(buf.getvalue().encode("utf-8")):
            fh.write(member)


def write_clusters_text(path: str, text: bytes) -> None:
    """The same file from the table's text (pacybara_b200/tablefmt.py)."""
    with open(path, "wb") as fh:
        for member in gzip_members(text):
            fh.write(member)


# --------------------------------------------------------------------------- loading a cluster-stage job

@dataclass
class ClusterJob:
    """Everything pacybara_cluster.py reads, as arrays plus the string tables for the writer."""
    problem: MergeProblem
    ed_q: np.ndarray
    ed_r: np.ndarray
    ed_d: np.ndarray
    ids: List[str]
    bc_names: List[str]
    up_names: Optional[List[str]]
    mut_names: List[str]


def load_cluster_job(ed_stream: Iterable[str], geno_stream: Iterable[str], preclust_stream: Iterable[str],
                     uptag_stream: Optional[Iterable[str]], max_diff: int = 2, min_qual: int = 32,
                     min_jaccard: float = 0.2, min_matches: int = 1) -> ClusterJob:
    pc = parse_preclust(preclust_stream)
    read_index = {rid: i for i, rid in enumerate(pc.ids)}
    if len(read_index) != len(pc.ids):
        raise FormatError("a read id occurs in more than one bucket title")
    title_index = {t: i for i, t in enumerate(pc.titles)}
    ed_q, ed_r, ed_d = parse_ed(ed_stream, title_index)
    geno = parse_geno(geno_stream, read_index, pc.n_reads)
    n_per = np.diff(pc.bucket_ptr)
    seq_id, bc_names = intern_strings(pc.seqs)
    bc_id = np.repeat(seq_id, n_per).astype(np.int32)
    up_id = None
    up_names = None
    if uptag_stream is not None:
        tags = parse_uptags(uptag_stream)
        if len(tags) > 0:                                   # pacybara_cluster.py:537
            up_id, up_names = intern_strings([tags.get(rid) for rid in pc.ids])
    prob = MergeProblem(n_reads=pc.n_reads, n_buckets=pc.n_buckets, bucket_ptr=pc.bucket_ptr,
                        bucket_reads=pc.bucket_reads, lexrank=lexrank_of(pc.ids), bc_id=bc_id, up_id=up_id,
                        geno_ptr=geno.geno_ptr, geno_mut=geno.geno_mut, geno_q=geno.geno_q,
                        mut_rank=geno.mut_rank(), max_diff=max_diff, min_qual=min_qual,
                        min_jaccard=min_jaccard, min_matches=min_matches)
    return ClusterJob(prob, ed_q, ed_r, ed_d, pc.ids, bc_names, up_names, geno.mut_strings)


# --------------------------------------------------------------------------- a14: rows from the merge outputs

MERGE_OUTPUTS = (("read_root", np.int32), ("read_pos", np.int32), ("row_size", np.int32), ("row_key", np.int64),
                 ("row_bc", np.int32), ("row_up", np.int32), ("row_call_off", np.int64), ("row_call_n", np.int32))


def merge_outputs_from_shards(shards: Sequence[Dict], n_reads: int) -> Dict[str, np.ndarray]:
    """Per-read merge outputs from the compacted results of all shards (Context.compact_shard): every read is owned
    by exactly one shard, rows are named by their representative read."""
    out = {name: np.full(n_reads, 0 if name == "row_size" else -1, dtype=dt) for name, dt in MERGE_OUTPUTS}
    if all("own_bucket" in s for s in shards):
        out["bucket_of_read"] = np.full(n_reads, -1, dtype=np.int32)
    calls, base = [], 0
    for s in shards:
        own = np.asarray(s["own_read"]).astype(np.int64)
        out["read_root"][own] = s["own_root"]
        out["read_pos"][own] = s["own_pos"]
        if "bucket_of_read" in out:
            out["bucket_of_read"][own] = s["own_bucket"]
        rep = np.asarray(s["row_rep"]).astype(np.int64)
        for name in ("row_size", "row_key", "row_bc", "row_up", "row_call_n"):
            out[name][rep] = s[name]
        out["row_call_off"][rep] = np.asarray(s["row_call_off"]) + base
        calls.append(np.asarray(s["call_mut"]))
        base += int(calls[-1].shape[0])
    if (out["read_root"] < 0).any():
        raise RuntimeError("some reads belong to no shard")
    out["call_mut"] = np.concatenate(calls) if calls else np.zeros(0, dtype=np.int32)
    return out


def rows_from_merge_outputs(out: Dict[str, np.ndarray]) -> ClusterRows:
    """Per-read / per-representative arrays of pcb_merge -> the table layout of
    pacybara_cluster.py:546-567: clusters in creation order of the surviving cluster id, then
    the unclustered reads (the reference emits those in `set` order; read order here)."""
    if "row_rep" in out:            # PCB_CLUSTER_COMPACT_ROWS: one entry per row already; same code on the expanded view
        rep0 = np.asarray(out["row_rep"]).astype(np.int64)
        R = out["read_root"].shape[0]
        full = {}
        for name, dt in MERGE_OUTPUTS[2:]:
            a = np.full(R, -1 if name != "row_size" else 0, dtype=dt)
            a[rep0] = out[name]
            full[name] = a
        out = dict(out, **full)
        del out["row_rep"]
    row_size = out["row_size"]
    rep = np.nonzero(row_size > 0)[0]
    key = out["row_key"][rep]
    single = key < 0
    order = np.lexsort((rep, key, single))
    rep = rep[order]
    n_rows = rep.shape[0]
    R = row_size.shape[0]
    row_of_rep = np.full(R, -1, dtype=np.int64)
    row_of_rep[rep] = np.arange(n_rows)
    row_of_read = row_of_rep[out["read_root"]]
    sizes = row_size[rep].astype(np.int64)
    row_ptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(sizes, out=row_ptr[1:])
    if int(row_ptr[-1]) != R or (R and int(row_of_read.min()) < 0):
        raise RuntimeError(f"merge outputs are inconsistent: rows hold {int(row_ptr[-1])} reads, expected {R}")
    # a read's place in the table: start of its row + its position in the member list (no sort over the reads)
    dest = row_ptr[row_of_read] + np.asarray(out["read_pos"]).astype(np.int64)
    members = np.full(R, -1, dtype=np.int32)
    if R and (int(dest.min()) < 0 or int(dest.max()) >= R):
        raise RuntimeError("merge outputs are inconsistent: a read's position lies outside its row")
    members[dest] = np.arange(R, dtype=np.int32)
    if R and int(members.min()) < 0:
        raise RuntimeError("merge outputs are inconsistent: two reads share a place in a row")
    ncall = out["row_call_n"][rep].astype(np.int64)
    call_ptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(ncall, out=call_ptr[1:])
    tot = int(call_ptr[-1])
    src = np.repeat(out["row_call_off"][rep] - call_ptr[:-1], ncall) + np.arange(tot)
    call_mut = out["call_mut"][src] if tot else np.zeros(0, dtype=np.int32)
    return ClusterRows(row_ptr.astype(np.int32), members, out["row_bc"][rep].astype(np.int32),
                       out["row_up"][rep].astype(np.int32), call_ptr.astype(np.int32),
                       call_mut.astype(np.int32), n_clusters=int((~single).sum()))
