"""Row a14, the writer (src/pacybara_cluster.py:546-567), without a Python loop over the reads.

A row of clusters.csv.gz is `virtualBarcode,upBarcode,read|read|...,size,mut;mut;...` (`=` for no call): slices of the
texts the fused path already holds -- read ids in the barcode FASTQ, mutation names in genoExtract, barcodes in the
sequence matrix, uptags in the uptag FASTQ -- glued with one-byte separators.  `table_items` lists, with numpy, which
slice every output item is and which separator follows it; `pcb_concat_tokens` (csrc/format.cu) sizes, scans and copies
on the device.  The rows, their order and their text are those of hostdata.rows_to_table + write_clusters_csv, which
stays the path for tables that need the reference's `muscle` fallback (a tied barcode vote, :484-511) and for inputs
the device tokenizers declined (no slice offsets then).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import hostdata

HEADER = b"virtualBarcode,upBarcode,reads,size,geno\n"
COMMA, BAR, SEMI, NL = (ord(c) for c in ",|;\n")


@dataclass
class Slices:
    """Slices of one text: text[off[i] : off[i] + len[i]]."""
    text: np.ndarray          # uint8
    off: np.ndarray           # int64
    len: np.ndarray           # int32


def _bytes(b) -> np.ndarray:
    return np.frombuffer(b, dtype=np.uint8) if isinstance(b, (bytes, bytearray, memoryview)) else np.ascontiguousarray(b, dtype=np.uint8).reshape(-1)


def table_items(rows: hostdata.ClusterRows, ids: Slices, bc: Slices, up: Optional[Slices], muts: Slices,
                bc_text: Optional[Dict[int, str]] = None, up_text: Optional[Dict[int, str]] = None
                ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """(src, tok_off, tok_len, item_tok, item_sep) of the table's rows (without the header line).
    ids: per read; bc: per barcode id (row_bc indexes it); up: per uptag id (row_up indexes it; None: the barcodes);
    muts: per mutation id.  bc_text / up_text: row -> string for the rows whose barcode / uptag is not an id (the
    alignment consensus of a tied vote, :484-511, which the caller has resolved)."""
    up = bc if up is None else up
    extra = list((bc_text or {}).items()), list((up_text or {}).items())
    extra_strings = [v for _, v in extra[0]] + [v for _, v in extra[1]]
    n_rows = rows.n_rows
    sizes = np.diff(rows.row_ptr).astype(np.int64)
    ncall = np.diff(rows.call_ptr).astype(np.int64)
    # ---- one source buffer, one token table: ids, barcodes, uptags, mutations, decimal sizes, "=", "None"
    max_size = int(sizes.max()) if n_rows else 0
    numbers = "".join(str(v) for v in range(max_size + 1)).encode("ascii")
    num_len = np.array([len(str(v)) for v in range(max_size + 1)], dtype=np.int32)
    num_off = np.zeros(max_size + 1, dtype=np.int64)
    np.cumsum(num_len[:-1], out=num_off[1:])
    consts = b"=None"
    ex_len = np.array([len(v) for v in extra_strings], dtype=np.int32)
    ex_off = np.zeros(len(extra_strings), dtype=np.int64)
    if len(extra_strings) > 1:
        np.cumsum(ex_len[:-1], out=ex_off[1:])
    parts = [ids, bc, up, muts, Slices(_bytes(numbers), num_off, num_len),
             Slices(_bytes(consts), np.array([0, 1], dtype=np.int64), np.array([1, 4], dtype=np.int32)),
             Slices(_bytes("".join(extra_strings).encode("ascii")), ex_off, ex_len)]
    texts, base, seen = [], [], {}
    at = 0
    for p in parts:                                           # (the uptag file is often the barcode file: one copy)
        key = id(p.text)
        if key not in seen:
            seen[key] = at
            texts.append(p.text)
            at += int(p.text.shape[0])
        base.append(seen[key])
    src = np.concatenate(texts) if texts else np.zeros(0, dtype=np.uint8)
    tok_off = np.concatenate([p.off.astype(np.int64) + b for p, b in zip(parts, base)])
    tok_len = np.concatenate([p.len.astype(np.int32) for p in parts])
    first = np.zeros(len(parts) + 1, dtype=np.int64)
    np.cumsum([p.off.shape[0] for p in parts], out=first[1:])
    T_ID, T_BC, T_UP, T_MUT, T_NUM, T_CONST, T_EXTRA = (int(x) for x in first[:-1])
    # ---- items: vbc , ubc , id | id ... , size , mut ; mut ... \n
    k_row = 3 + sizes + np.maximum(ncall, 1)
    item_ptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(k_row, out=item_ptr[1:])
    K = int(item_ptr[-1])
    item_tok = np.empty(K, dtype=np.int32)
    item_sep = np.empty(K, dtype=np.uint8)
    p0 = item_ptr[:-1]
    item_tok[p0] = T_BC + np.maximum(rows.row_bc.astype(np.int64), 0)
    item_sep[p0] = COMMA
    rup = rows.row_up.astype(np.int64)
    item_tok[p0 + 1] = np.where(rup >= 0, T_UP + rup, T_CONST + 1)                     # "None" (:558)
    item_sep[p0 + 1] = COMMA
    for k, (row, _) in enumerate(extra[0]):
        item_tok[p0[row]] = T_EXTRA + k
    for k, (row, _) in enumerate(extra[1]):
        item_tok[p0[row] + 1] = T_EXTRA + len(extra[0]) + k
    n_mem = int(sizes.sum())
    within = np.arange(n_mem, dtype=np.int64) - np.repeat(rows.row_ptr[:-1].astype(np.int64), sizes)
    pos = np.repeat(p0 + 2, sizes) + within
    item_tok[pos] = T_ID + rows.row_members.astype(np.int64)
    item_sep[pos] = BAR
    item_sep[p0 + 1 + sizes] = COMMA                                                      # the last member of every row
    item_tok[p0 + 2 + sizes] = T_NUM + sizes
    item_sep[p0 + 2 + sizes] = COMMA
    c0 = p0 + 3 + sizes
    n_call = int(ncall.sum())
    if n_call:
        cwithin = np.arange(n_call, dtype=np.int64) - np.repeat(rows.call_ptr[:-1].astype(np.int64), ncall)
        cpos = np.repeat(c0, ncall) + cwithin
        item_tok[cpos] = T_MUT + rows.call_mut.astype(np.int64)
        item_sep[cpos] = SEMI
    none = ncall == 0
    item_tok[c0[none]] = T_CONST                                                          # "=" (:553)
    item_sep[c0 + np.maximum(ncall, 1) - 1] = NL                                          # the last call (or "=") ends the row
    return src, tok_off, tok_len, item_tok, item_sep


def concat_host(src, tok_off, tok_len, item_tok, item_sep) -> bytes:
    """What pcb_concat_tokens computes, with numpy -- for the CPU tests of table_items (never the product path)."""
    out = bytearray()
    raw = src.tobytes()
    for t, s in zip(item_tok.tolist(), item_sep.tolist()):
        out += raw[int(tok_off[t]): int(tok_off[t]) + int(tok_len[t])]
        if s:
            out.append(s)
    return bytes(out)


def table_bytes(ctx, rows: hostdata.ClusterRows, ids: Slices, bc: Slices, up: Optional[Slices], muts: Slices,
                bc_text: Optional[Dict[int, str]] = None, up_text: Optional[Dict[int, str]] = None) -> bytes:
    """The whole text of clusters.csv (header + rows), the rows formatted on the device."""
    src, tok_off, tok_len, item_tok, item_sep = table_items(rows, ids, bc, up, muts, bc_text, up_text)
    return HEADER + ctx.concat_tokens(src, tok_off, tok_len, item_tok, item_sep)


def barcode_slices(seq: np.ndarray, length: np.ndarray, bucket_of_read: np.ndarray, n_buckets: int) -> Slices:
    """Barcode of every bucket id as a slice of the [R, stride] sequence matrix (the bucket's first read)."""
    R = int(bucket_of_read.shape[0])
    first = np.full(n_buckets, R, dtype=np.int64)
    np.minimum.at(first, bucket_of_read, np.arange(R, dtype=np.int64))
    stride = int(seq.shape[1])
    return Slices(np.ascontiguousarray(seq, dtype=np.uint8).reshape(-1), first * stride, np.asarray(length)[first].astype(np.int32))
