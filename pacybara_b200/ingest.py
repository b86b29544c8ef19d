"""f2: from the gzip files of the clustering section to device-resident ReadArrays without a Python line loop.

    bcExtract_*.fastq.gz, genoExtract.csv.gz [, uptag FASTQ]  --gunzip-->  bytes  --H2D-->  pcb_text tokenizers

The tokenizers (csrc/ingest.cu) restate the reference's loaders (pacybara_precluster.py:47-63,
pacybara_cluster.py:302-331) as data-parallel passes and COUNT everything they do not model exactly; when any
such counter is non-zero this module hands the same files to the line parsers of hostdata.py (which mirror the
reference line by line and raise its errors).  What stays on the host: gunzip (zlib, one thread per file), the
names of the M distinct mutations and their (digit key, string) order, and the id strings the writer prints.
"""
from __future__ import annotations

import gzip
import threading
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import hostdata, pipeline
from .api import Context, Text


class NeedsLineParser(Exception):
    """The device tokenizers met something only the line parser models (message says what)."""


def read_bytes(path: str) -> bytes:
    if path == "-":
        import sys
        return sys.stdin.buffer.read()
    with open(path, "rb") as fh:
        head = fh.read(2)
    if head == b"\x1f\x8b":
        with gzip.open(path, "rb") as fh:
            return fh.read()
    with open(path, "rb") as fh:
        return fh.read()


def read_all(paths: List[Optional[str]]) -> List[Optional[bytes]]:
    """Decompress the files concurrently (zlib releases the GIL)."""
    out: List[Optional[bytes]] = [None] * len(paths)
    errs: List[BaseException] = []

    def work(i):
        try:
            out[i] = read_bytes(paths[i])
        except BaseException as e:       # noqa: BLE001 - re-raised below
            errs.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i, p in enumerate(paths) if p is not None]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errs:
        raise errs[0]
    return out


_decoded: Dict[int, tuple] = {}


def _text(raw: bytes) -> str:
    """ASCII text of a file's bytes, decoded once per buffer (the uptag file is often the barcode file itself)."""
    hit = _decoded.get(id(raw))
    if hit is None or hit[0] is not raw:
        _decoded.clear()
        hit = _decoded[id(raw)] = (raw, raw.decode("ascii"))
    return hit[1]


def _slices(raw: bytes, off: np.ndarray, length: np.ndarray) -> List[str]:
    import gc
    txt = _text(raw)
    on = gc.isenabled()
    gc.disable()                 # a million fresh strings: keep the cyclic collector from rescanning them as the list grows
    try:
        return [txt[o:e] for o, e in zip(off.tolist(), (off + length).tolist())]
    finally:
        if on:
            gc.enable()


def fastq_from_bytes(ctx: Context, fastq: bytes) -> hostdata.FastqReads:
    """The reads pacybara_precluster.py:47-63 would keep, as device matrices (torch CUDA tensors) + id strings."""
    with Text(ctx, fastq) as fq:
        if fq.n_exotic:
            raise NeedsLineParser(f"FASTQ has bytes outside plain ASCII text: {fq.n_exotic}")
        rec = fq.fastq_records()
        if rec["n_bad"]:
            raise NeedsLineParser(f"FASTQ records the tokenizer does not model: {rec['n_bad']}")
        f = fq.fastq_fetch(max(4, (rec["max_len"] + 3) // 4 * 4))
    ids = _slices(fastq, f["id_off"].cpu().numpy(), f["id_len"].cpu().numpy())
    _decoded.clear()
    return hostdata.FastqReads(ids, f["seq"], f["qual"], f["length"])


def arrays_from_bytes(ctx: Context, fastq: bytes, geno: bytes, uptag: Optional[bytes] = None, laps: Optional[Dict] = None
                      ) -> Tuple[pipeline.ReadArrays, Dict]:
    """Device-resident ReadArrays (torch CUDA tensors) + the string tables of the writer.
    Raises NeedsLineParser when the tokenizers decline.  laps: filled with seconds per step."""
    import time
    import torch
    dev = f"cuda:{ctx.device}"
    t_last = [time.perf_counter()]

    def lap(name):
        if laps is not None:
            now = time.perf_counter()
            laps[name] = laps.get(name, 0.0) + round(now - t_last[0], 4)
            t_last[0] = now

    def decline(what: str, n):
        raise NeedsLineParser(f"{what}: {n}")

    with Text(ctx, fastq) as fq:
        if fq.n_exotic:
            decline("barcode FASTQ has bytes outside plain ASCII text", fq.n_exotic)
        lap("open_fastq")
        rec = fq.fastq_records()
        if rec["n_bad"]:
            decline("barcode FASTQ records the tokenizer does not model", rec["n_bad"])
        if rec["max_len"] > 4 * hostdata.MAX_BARCODE:
            decline("a barcode too long for the tokenizer's matrices", rec["max_len"])
        stride = max(4, (rec["max_len"] + 3) // 4 * 4)
        f = fq.fastq_fetch(stride)
        R = rec["n_records"]
        lap("fastq_records")
        lexrank = fq.rank_strings(f["id_off"], f["id_len"], rec["max_id_len"])
        lap("rank_ids")
        with Text(ctx, geno) as ge:
            lap("open_geno")
            if ge.n_exotic:
                decline("genoExtract has bytes outside plain ASCII text", ge.n_exotic)
            j = ge.geno_join(fq, f["id_off"], f["id_len"])
            if j["collisions"]:
                decline("duplicate read ids (or a hash collision)", j["collisions"])
            if j["malformed"]:
                decline("genoExtract lines the tokenizer does not model", j["malformed"])
            if j["repeated"]:
                decline("reads naming a mutation twice", j["repeated"])
            if j["missing"]:
                raise KeyError(f"{j['missing']} bucketed read(s) have no genotype record")
            g = ge.geno_fetch()
            lap("geno_join")
            mut_off, mut_len = g["mut_off"].cpu().numpy(), g["mut_len"].cpu().numpy()
            mut_names = _slices(geno, mut_off, mut_len)
        lap("mut_names")
        up_id, up_names, up_slices = None, None, None
        if uptag is not None:
            up = fq if uptag is fastq else Text(ctx, uptag)                # the same file: one copy, one index
            try:
                if up.n_exotic:
                    decline("uptag FASTQ has bytes outside plain ASCII text", up.n_exotic)
                uj = up.uptag_join(fq, f["id_off"], f["id_len"])
                if uj["collisions"] or uj["too_long"]:
                    decline("uptag records the tokenizer does not model", uj["collisions"] + uj["too_long"])
                if uj["n_records"] > 0:                      # `if len(tags) > 0`, pacybara_cluster.py:376
                    u = up.uptag_fetch()
                    up_id = u["up_id"]
                    tag_off, tag_len = u["tag_off"].cpu().numpy(), u["tag_len"].cpu().numpy()
                    up_names = _slices(uptag, tag_off, tag_len)
                    up_slices = (uptag, tag_off, tag_len)
            finally:
                if up is not fq:
                    up.close()
    lap("uptags")
    # the read ids stay slices of the FASTQ bytes: the writer formats the table from them on the device
    # (pacybara_b200/tablefmt.py); the million Python strings are only made when somebody asks (LazyIds)
    id_off, id_len = f["id_off"].cpu().numpy(), f["id_len"].cpu().numpy()
    ids = LazyIds(fastq, id_off, id_len)
    lap("id_strings")
    mut_rank = torch.from_numpy(pipeline.mut_rank_of(mut_names)).to(dev)
    ra = pipeline.ReadArrays(seq=f["seq"], qual=f["qual"], length=f["length"], lexrank=lexrank, geno_ptr=g["geno_ptr"],
                             geno_mut=g["geno_mut"], geno_q=g["geno_q"], mut_rank=mut_rank, up_id=up_id)
    assert ra.n_reads == R
    _decoded.clear()
    lap("mut_rank")
    slices = dict(ids=(fastq, id_off, id_len), muts=(geno, mut_off, mut_len), up=up_slices)
    return ra, dict(ids=ids, mut_names=mut_names, up_names=up_names, slices=slices)


def cluster_job_from_files(ctx: Context, ed_file: str, geno_file: str, preclust_file: str, uptag_file: Optional[str],
                           max_diff: int = 2, min_qual: int = 32, min_jaccard: float = 0.2, min_matches: int = 1, pc=None):
    """What hostdata.load_cluster_job builds for the pacybara_cluster.py drop-in, with the two big loaders -- genotypes and
    uptags, one Python iteration per read in the line parsers (3.5 + 1.7 s at 10^6 reads) -- done by the device tokenizers:
    the read ids of the bucket titles are laid out as a text of their own, which the joins take like the id slices of a
    FASTQ.  Returns (ClusterJob, slices for tablefmt).  Raises NeedsLineParser when a tokenizer declines."""
    if pc is None:
        with hostdata.open_text(preclust_file) as fh:
            pc = hostdata.parse_preclust(fh)
    if len(set(pc.ids)) != len(pc.ids):
        raise hostdata.FormatError("a read id occurs in more than one bucket title")
    title_index = {t: i for i, t in enumerate(pc.titles)}
    with hostdata.open_text(ed_file) as fh:
        ed_q, ed_r, ed_d = hostdata.parse_ed(fh, title_index)
    id_len = np.fromiter((len(r) for r in pc.ids), dtype=np.int32, count=len(pc.ids))
    id_off = np.zeros(len(pc.ids), dtype=np.int64)
    if len(pc.ids) > 1:
        np.cumsum(id_len[:-1].astype(np.int64) + 1, out=id_off[1:])
    try:
        ids_text = ("\n".join(pc.ids) + "\n").encode("ascii")
    except UnicodeEncodeError:
        raise NeedsLineParser("read ids outside plain ASCII") from None
    raw = read_all([geno_file, uptag_file])
    geno, uptag = raw[0], raw[1]

    def decline(what: str, n):
        raise NeedsLineParser(f"{what}: {n}")

    up_id, up_names, up_slices = None, None, None
    with Text(ctx, ids_text) as idt:
        with Text(ctx, geno) as ge:
            if ge.n_exotic:
                decline("genoExtract has bytes outside plain ASCII text", ge.n_exotic)
            j = ge.geno_join(idt, id_off, id_len)
            if j["collisions"]:
                decline("duplicate read ids (or a hash collision)", j["collisions"])
            if j["malformed"]:
                decline("genoExtract lines the tokenizer does not model", j["malformed"])
            if j["repeated"]:
                decline("reads naming a mutation twice", j["repeated"])
            if j["missing"]:
                raise KeyError(f"{j['missing']} bucketed read(s) have no genotype record")
            g = ge.geno_fetch(device=False)
            mut_off, mut_len = np.asarray(g["mut_off"]), np.asarray(g["mut_len"])
            mut_names = _slices(geno, mut_off, mut_len)
        if uptag is not None:
            with Text(ctx, uptag) as up:
                if up.n_exotic:
                    decline("uptag FASTQ has bytes outside plain ASCII text", up.n_exotic)
                uj = up.uptag_join(idt, id_off, id_len)
                if uj["collisions"] or uj["too_long"]:
                    decline("uptag records the tokenizer does not model", uj["collisions"] + uj["too_long"])
                if uj["n_records"] > 0:                      # `if len(tags) > 0`, pacybara_cluster.py:376
                    u = up.uptag_fetch(device=False)
                    up_id = np.asarray(u["up_id"]).astype(np.int32)
                    tag_off, tag_len = np.asarray(u["tag_off"]), np.asarray(u["tag_len"])
                    up_names = _slices(uptag, tag_off, tag_len)
                    up_slices = (uptag, tag_off, tag_len)
    _decoded.clear()
    n_per = np.diff(pc.bucket_ptr)
    seq_id, bc_names = hostdata.intern_strings(pc.seqs)
    bc_id = np.repeat(seq_id, n_per).astype(np.int32)
    prob = hostdata.MergeProblem(n_reads=pc.n_reads, n_buckets=pc.n_buckets, bucket_ptr=pc.bucket_ptr, bucket_reads=pc.bucket_reads,
                                 lexrank=hostdata.lexrank_of(pc.ids), bc_id=bc_id, up_id=up_id,
                                 geno_ptr=np.asarray(g["geno_ptr"]).astype(np.int32), geno_mut=np.asarray(g["geno_mut"]).astype(np.int32),
                                 geno_q=np.asarray(g["geno_q"]).astype(np.int32), mut_rank=pipeline.mut_rank_of(mut_names),
                                 max_diff=max_diff, min_qual=min_qual, min_jaccard=min_jaccard, min_matches=min_matches)
    job = hostdata.ClusterJob(prob, ed_q, ed_r, ed_d, pc.ids, bc_names, up_names, mut_names)
    bc_len = np.fromiter((len(b) for b in bc_names), dtype=np.int32, count=len(bc_names))
    bc_off = np.zeros(len(bc_names), dtype=np.int64)
    if len(bc_names) > 1:
        np.cumsum(bc_len[:-1], out=bc_off[1:])
    slices = dict(ids=(ids_text, id_off, id_len), muts=(geno, mut_off, mut_len), up=up_slices,
                  bc=("".join(bc_names).encode("ascii"), bc_off, bc_len))
    return job, slices


class LazyIds:
    """The read ids as a sequence of str that is only built (10^6 slices, 0.15 s) when it is indexed or iterated."""

    def __init__(self, raw: bytes, off: np.ndarray, length: np.ndarray):
        self.raw, self.off, self.length = raw, off, length
        self._list: Optional[List[str]] = None

    def _get(self) -> List[str]:
        if self._list is None:
            self._list = _slices(self.raw, self.off, self.length)
            _decoded.clear()
        return self._list

    def __len__(self):
        return int(self.off.shape[0])

    def __getitem__(self, i):
        return self._get()[i]

    def __iter__(self):
        return iter(self._get())

    def __eq__(self, other):
        return self._get() == (other._get() if isinstance(other, LazyIds) else other)

    __hash__ = None

    def any_contains(self, ch: bytes) -> Optional[str]:
        """The first id that contains the byte `ch`, without building the strings (None: no id does)."""
        buf = np.frombuffer(self.raw, dtype=np.uint8)
        hits = np.flatnonzero(buf == ch[0])
        if hits.size == 0 or len(self) == 0:
            return None
        order = np.argsort(self.off, kind="stable")
        off, length = self.off[order].astype(np.int64), self.length[order].astype(np.int64)
        k = np.searchsorted(off, hits, side="right") - 1
        ok = (k >= 0) & (hits < off[np.maximum(k, 0)] + length[np.maximum(k, 0)])
        if not ok.any():
            return None
        i = int(order[k[ok].min()])
        return self.raw[int(self.off[i]): int(self.off[i]) + int(self.length[i])].decode("ascii")


def arrays_from_files(ctx: Context, bc_file: str, geno_file: str, uptag_file: Optional[str] = None, log=None
                      ) -> Tuple[pipeline.ReadArrays, Dict]:
    """The tokenizer path with the line parsers behind it."""
    same = uptag_file is not None and uptag_file == bc_file
    raw = read_all([bc_file, geno_file, None if same else uptag_file])
    if same:
        raw[2] = raw[0]
    try:
        return arrays_from_bytes(ctx, raw[0], raw[1], raw[2])
    except NeedsLineParser as e:
        if log:
            log(f" - tokenizer declined ({e}); using the line parser")
    import io
    streams = [io.TextIOWrapper(io.BytesIO(b)) if b is not None else None for b in raw]     # as hostdata.open_text reads them
    return pipeline.arrays_from_text(streams[0], streams[1], streams[2], ctx)
