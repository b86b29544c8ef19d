"""f3: barcode -> library matching with maxErr, the matcher barseq_caller.R builds at :57 and calls at :242.

    matcher <- new.bc.matcher(bcRef$barcode, errCutoff=args$maxErr)         # src/barseq_caller.R:57
    matches <- matcher$findMatches(bcReads, bcReads2)                       # :242
    # "hits: list of library row numbers with matching barcodes; diffs: the edit distance between barcode and
    #  match; nhits: the number of equidistant hits"  (:238-241)

`yogiseq` (where new.bc.matcher lives) is not part of the reference tree, so the matcher is restated from this call
site and its comment: per read the library rows at the smallest distance <= errCutoff, that distance, their number.
The distance is unit-cost Levenshtein, computed by the same bit-parallel kernel and seed index as the clustering
path (pcb_match_barcodes).  Parity is pinned to an exhaustive-DP checker in the test tree (orc_match_brute), not to
yogiseq: "parity unpinned" in the sense of DESIGN.md section 2.

Reads are matched once per distinct barcode (exact duplicates are bucketed on the GPU first).  A read barcode may hold
characters outside ACGT (N): such a character equals nothing and costs one edit, as in a DP over the strings.  A
barcode that is missing (None / NA: failed extraction, :139-195), empty or longer than 64 is reported unmatched;
library rows that are not 1..64 ACGT bases can never be hit.  With a second read (paired-end) the read with the
smaller distance decides; on equal distance the rows of both count.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import hostdata
from .api import Context

_OK = np.zeros(256, dtype=bool)
_OK[[ord(c) for c in "ACGT"]] = True


def _scorable(bc: Optional[str]) -> bool:
    return isinstance(bc, str) and 1 <= len(bc) <= hostdata.MAX_BARCODE


def _valid(bc: Optional[str]) -> bool:
    """1..64 bases over ACGT (what a library row must be, and what the exact-duplicate bucketing takes)."""
    return _scorable(bc) and bool(_OK[np.frombuffer(bc.encode("latin-1", "replace"), dtype=np.uint8)].all())


class BcMatcher:
    """new.bc.matcher(barcodes, errCutoff): rows are numbered from 1, as R does."""

    def __init__(self, barcodes: Sequence[Optional[str]], err_cutoff: int = 2, ctx: Optional[Context] = None):
        if err_cutoff < 0 or err_cutoff > 8:
            raise ValueError("errCutoff must be between 0 and 8")
        self.err_cutoff = int(err_cutoff)
        self.ctx = ctx if ctx is not None else Context(0)
        self.n_rows = len(barcodes)
        self.rows = np.array([i for i, b in enumerate(barcodes) if _valid(b)], dtype=np.int64)   # rows that can match
        self.lib_seq, self.lib_len = hostdata.strings_to_matrix([barcodes[i].encode() for i in self.rows])

    def match_arrays(self, barcodes: Sequence[Optional[str]]) -> Dict:
        """Per read: best distance (-1: none), number of rows at it, and the CSR list of those rows (0-based)."""
        n = len(barcodes)
        ok = np.array([_scorable(b) for b in barcodes], dtype=bool)
        best = np.full(n, -1, dtype=np.int32)
        nhits = np.zeros(n, dtype=np.int32)
        ptr = np.zeros(n + 1, dtype=np.int64)
        hits = np.zeros(0, dtype=np.int64)
        idx = np.nonzero(ok)[0]
        if idx.shape[0] and self.rows.shape[0]:
            enc = [barcodes[i].encode("latin-1", "replace") for i in idx]
            seq, length = hostdata.strings_to_matrix(enc)
            # one query per distinct barcode: plain ACGT barcodes are bucketed on the GPU, the few others go as they are
            clean = _OK[seq].all(axis=1, where=np.arange(seq.shape[1])[None, :] < length[:, None])
            ci, oi = np.nonzero(clean)[0], np.nonzero(~clean)[0]
            bor = np.empty(idx.shape[0], dtype=np.int64)
            n_u = 0
            useq, ulen = seq[:0], length[:0]
            if ci.shape[0]:
                b = self.ctx.bucket(seq[ci], None, length[ci], want_qsum=False)
                n_u, useq, ulen = b["n_buckets"], b["bucket_seq"], b["bucket_len"]
                bor[ci] = b["bucket_of_read"]
            bor[oi] = n_u + np.arange(oi.shape[0])
            useq, ulen = np.concatenate([useq, seq[oi]]), np.concatenate([ulen, length[oi]])
            n_u += oi.shape[0]
            u = self.ctx.match_barcodes(self.lib_seq, self.lib_len, useq, ulen, self.err_cutoff, want_pairs=True)
            keep = u["pair_d"] == u["best_d"][u["pair_q"]]                     # the equidistant best rows
            pq, prow = u["pair_q"][keep], self.rows[u["pair_row"][keep]]
            uptr = np.zeros(n_u + 1, dtype=np.int64)
            np.cumsum(np.bincount(pq, minlength=n_u), out=uptr[1:])
            best[idx] = u["best_d"][bor]
            nhits[idx] = u["n_best"][bor]
            np.cumsum(nhits, out=ptr[1:])
            src = np.repeat(uptr[bor] - ptr[idx], nhits[idx]) + np.arange(int(ptr[-1]))
            hits = prow[src]
        return dict(best_d=best, nhits=nhits, hit_ptr=ptr, hits=hits)

    def find_matches(self, bc_reads: Sequence[Optional[str]], bc_reads2: Optional[Sequence[Optional[str]]] = None) -> Dict[str, List]:
        """matcher$findMatches: {'hits': [[row, ...] 1-based], 'diffs': [d or None], 'nhits': [n]} per read."""
        a = self.match_arrays(bc_reads)
        rows = [a["hits"][a["hit_ptr"][i]: a["hit_ptr"][i + 1]] for i in range(len(bc_reads))]
        best = a["best_d"].copy()
        if bc_reads2 is not None:
            if len(bc_reads2) != len(bc_reads):
                raise ValueError("R1 and R2 barcodes differ in number")
            b = self.match_arrays(bc_reads2)
            for i in range(len(bc_reads)):
                r2 = b["hits"][b["hit_ptr"][i]: b["hit_ptr"][i + 1]]
                d1, d2 = int(best[i]), int(b["best_d"][i])
                if d2 >= 0 and (d1 < 0 or d2 < d1):
                    rows[i], best[i] = r2, d2
                elif d2 >= 0 and d2 == d1:
                    rows[i] = np.union1d(rows[i], r2)
        return dict(hits=[(r + 1).tolist() for r in rows], diffs=[int(d) if d >= 0 else None for d in best],
                    nhits=[int(len(r)) for r in rows])
