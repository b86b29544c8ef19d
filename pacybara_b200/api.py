"""Python host API over the C-ABI (include/pacybara_b200.h).

One `Context` per GPU.  Every method takes either numpy arrays (host memory; the library copies
in and out) or torch CUDA tensors (device memory; nothing is copied) -- all arrays of one call
must be of the same kind -- and returns arrays of that kind.  torch is only used to own device
memory; every computation happens in the hand-written kernels behind the ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .hostdata import MERGE_OUTPUTS, MergeProblem, PairList

_NP2TORCH = None


def _torch():
    import torch
    global _NP2TORCH
    if _NP2TORCH is None:
        _NP2TORCH = {np.dtype(np.uint8): torch.uint8, np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64}
    return torch


def _is_dev(x) -> bool:
    return x is not None and hasattr(x, "data_ptr")


class Context:
    def __init__(self, device: int = 0, pinned_outputs: bool = False):
        """pinned_outputs: host-memory results are written into page-locked buffers that are REUSED by
        the next call of the same method (copy what you keep) -- device-to-host copies then run at
        full PCIe rate instead of being staged through pageable memory."""
        self.lib = _lib.load()
        self.device = int(device)
        self.pinned_outputs = bool(pinned_outputs)
        self._pinned: Dict = {}
        h = C.c_void_p()
        rc = self.lib.pcb_ctx_create(self.device, C.byref(h))
        if rc != 0:
            raise _lib.PcbError(rc, (self.lib.pcb_last_error(None) or b"").decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.pcb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _check(self, rc: int):
        if rc != 0:
            raise _lib.PcbError(rc, (self.lib.pcb_last_error(self.h) or b"").decode())

    def stat(self, name: str) -> int:
        return int(self.lib.pcb_stat(self.h, _lib.STAT[name]))

    def stats(self) -> Dict[str, int]:
        return {k: self.stat(k) for k in _lib.STAT}

    def sync(self):
        self._check(self.lib.pcb_sync(self.h))

    def set_replay_shape(self, shape: Optional[str] = None):
        """'auto' (default), 'thread' or 'warp' (pcb_set_option).  None: what the environment variable PCB_REPLAY
        says -- read here on the host, per call, so that tests and measurements can switch shapes on a live context."""
        import os
        shape = shape or os.environ.get("PCB_REPLAY") or "auto"
        if shape not in _lib.REPLAY_SHAPES:
            raise ValueError(f"unknown replay shape {shape!r}")
        self._check(self.lib.pcb_set_option(self.h, _lib.OPT_REPLAY_SHAPE, _lib.REPLAY_SHAPES[shape]))
        if os.environ.get("PCB_TRACE"):
            self._check(self.lib.pcb_set_option(self.h, 1, 1))
        staged = 1 if os.environ.get("PCB_PROBE_STAGING") else 0
        if getattr(self, "_staging_set", 0) != staged:
            self._check(self.lib.pcb_set_option(self.h, _lib.OPT_PROBE_STAGING, staged))
            self._staging_set = staged
        classes = 0 if os.environ.get("PCB_PROBE_CLASSES") == "0" else 1
        if getattr(self, "_classes_set", 1) != classes:
            self._check(self.lib.pcb_set_option(self.h, _lib.OPT_PROBE_CLASSES, classes))
            self._classes_set = classes
        sw = int(os.environ.get("PCB_SMALL_WARPS") or getattr(self, "_small_warps_default", 1))
        if getattr(self, "_sw_set", 1) != sw:
            self._check(self.lib.pcb_set_option(self.h, _lib.OPT_SMALL_WARPS, sw))
            self._sw_set = sw
        if os.environ.get("PCB_HUGE_WARPS"):
            self._check(self.lib.pcb_set_option(self.h, 101, int(os.environ["PCB_HUGE_WARPS"])))
        if os.environ.get("PCB_SMALL_PAD"):
            self._check(self.lib.pcb_set_option(self.h, 100, int(os.environ["PCB_SMALL_PAD"])))
        hb = int(os.environ.get("PCB_HUGE_BUCKETS") or getattr(self, "_huge_buckets_default", 16))
        if getattr(self, "_hb_set", 16) != hb:
            self._check(self.lib.pcb_set_option(self.h, _lib.OPT_HUGE_BUCKETS, hb))
            self._hb_set = hb
        bins = int(os.environ.get("PCB_SEED_BINS_LOG2") or 26)
        if getattr(self, "_bins_set", 26) != bins:
            self._check(self.lib.pcb_set_option(self.h, _lib.OPT_SEED_BINS_LOG2, bins))
            self._bins_set = bins
        # giant-tier knobs: environment > set_huge() > the library's defaults, applied per call like the shape
        for env, opt, default in (("PCB_HUGE_READS", _lib.OPT_HUGE_READS, 256), ("PCB_HUGE_CAP", _lib.OPT_HUGE_CAP, 1024),
                                  ("PCB_HUGE_WINDOW", _lib.OPT_HUGE_WINDOW, 32768)):
            v = int(os.environ[env]) if os.environ.get(env) else getattr(self, "_huge", {}).get(opt, default)
            if getattr(self, "_huge_set", {}).get(opt) != v:
                self._check(self.lib.pcb_set_option(self.h, opt, v))
                self._huge_set = dict(getattr(self, "_huge_set", {}))
                self._huge_set[opt] = v

    def set_huge(self, reads: Optional[int] = None, cap: Optional[int] = None, window: Optional[int] = None):
        """Knobs of the giant tier (pcb_set_option): component size above which it is used, slots of a warp's scratch
        table, rows per round.  Results do not depend on them."""
        self._huge = dict(getattr(self, "_huge", {}))
        for v, opt in ((reads, _lib.OPT_HUGE_READS), (cap, _lib.OPT_HUGE_CAP), (window, _lib.OPT_HUGE_WINDOW)):
            if v is not None:
                self._huge[opt] = int(v)

    def stream_handle(self) -> int:
        return int(self.lib.pcb_stream(self.h) or 0)

    def _kind(self, *arrays) -> int:
        kinds = {_is_dev(a) for a in arrays if a is not None}
        if len(kinds) > 1:
            raise TypeError("all arrays of a call must be numpy (host) or all torch CUDA tensors (device)")
        dev = kinds.pop() if kinds else False
        if dev:
            for a in arrays:
                if a is not None and (not a.is_cuda or a.device.index != self.device or not a.is_contiguous()):
                    raise TypeError(f"device arrays must be contiguous tensors on cuda:{self.device}")
            # the library works on its own stream: whatever torch still has queued for these tensors
            # must have finished (the library itself returns only after its stream has drained)
            _torch().cuda.current_stream(self.device).synchronize()
        return _lib.MEM_DEVICE if dev else _lib.MEM_HOST

    def _in(self, a, dtype, keep: list):
        """pointer of an input array (numpy arrays are made contiguous / cast; kept alive in `keep`)."""
        if a is None:
            return None
        if _is_dev(a):
            if a.dtype != _NP2TORCH[np.dtype(dtype)]:
                raise TypeError(f"device array has dtype {a.dtype}, expected {np.dtype(dtype)}")
            return C.c_void_p(a.data_ptr())
        a = np.ascontiguousarray(a, dtype=dtype)
        keep.append(a)
        return a.ctypes.data_as(C.c_void_p)

    def _out(self, mem: int, shape, dtype, tag: str = ""):
        if mem == _lib.MEM_DEVICE:
            torch = _torch()
            t = torch.empty(shape, dtype=_NP2TORCH[np.dtype(dtype)], device=f"cuda:{self.device}")
            return t, C.c_void_p(t.data_ptr())
        if self.pinned_outputs and tag:
            torch = _torch()
            key = (tag, tuple(shape), np.dtype(dtype).str)
            t = self._pinned.get(key)
            if t is None:
                t = self._pinned[key] = torch.empty(shape, dtype=_NP2TORCH[np.dtype(dtype)]).pin_memory()
            a = t.numpy()
            return a, a.ctypes.data_as(C.c_void_p)
        a = np.empty(shape, dtype=dtype)
        return a, a.ctypes.data_as(C.c_void_p)

    # ------------------------------------------------------------------ a1
    def bucket(self, seq, qual, length, want_qsum: bool = True) -> Dict:
        """Exact-barcode bucketing (pcb_bucket).  seq/qual: [R, stride] uint8, length: [R] int32."""
        if _is_dev(seq):
            _torch()
        mem = self._kind(seq, qual, length)
        R, stride = int(seq.shape[0]), int(seq.shape[1])
        keep: list = []
        bor, p_bor = self._out(mem, (max(R, 1),), np.int32)
        bptr, p_bptr = self._out(mem, (R + 1,), np.int32)
        breads, p_breads = self._out(mem, (max(R, 1),), np.int32)
        if want_qsum:
            qsum, p_qsum = self._out(mem, (max(R, 1), stride), np.uint8)
        else:
            qsum, p_qsum = None, None
        nb = C.c_int32(0)
        self._check(self.lib.pcb_bucket(self.h, mem, self._in(seq, np.uint8, keep), self._in(qual, np.uint8, keep),
                                        self._in(length, np.int32, keep), R, stride, p_bor, C.byref(nb), p_bptr,
                                        p_breads, p_qsum))
        U = int(nb.value)
        first = breads[bptr[:U].long()] if mem == _lib.MEM_DEVICE else breads[bptr[:U]]
        return dict(n_buckets=U, bucket_of_read=bor[:R], bucket_ptr=bptr[: U + 1], bucket_reads=breads[:R],
                    bucket_qsum=None if qsum is None else qsum[:U], first_read=first,
                    bucket_seq=seq[first.long()] if mem == _lib.MEM_DEVICE else np.asarray(seq)[first],
                    bucket_len=length[first.long()] if mem == _lib.MEM_DEVICE else np.asarray(length)[first])

    # ------------------------------------------------------------------ a2+a3
    def edit_pairs(self, seq, length, k: int, q_range: Optional[Tuple[int, int]] = None,
                   t_range: Optional[Tuple[int, int]] = None):
        """All pairs i<j of the U barcodes with Levenshtein <= k, sorted by (i, j) (pcb_edit_pairs).  Every unordered pair
        has one query side and one target side: q_range / t_range restrict the call to the pairs whose query / target
        lies in the range (disjoint ranges that cover [0, U) give disjoint edge sets whose union is everything)."""
        if _is_dev(seq):
            _torch()
        mem = self._kind(seq, length)
        self.set_replay_shape()
        U, stride = int(seq.shape[0]), int(seq.shape[1])
        lo, hi = (0, U) if q_range is None else q_range
        tlo, thi = (0, U) if t_range is None else t_range
        keep: list = []
        n = C.c_int64(0)
        self._check(self.lib.pcb_edit_pairs(self.h, mem, self._in(seq, np.uint8, keep), self._in(length, np.int32, keep),
                                            U, stride, int(k), int(lo), int(hi), int(tlo), int(thi), C.byref(n)))
        E = int(n.value)
        ei, p_i = self._out(mem, (max(E, 1),), np.int32)
        ej, p_j = self._out(mem, (max(E, 1),), np.int32)
        ed, p_d = self._out(mem, (max(E, 1),), np.int32)
        self._check(self.lib.pcb_edit_pairs_fetch(self.h, mem, p_i, p_j, p_d, max(E, 1)))
        return ei[:E], ej[:E], ed[:E]

    def sort_edges(self, ei, ej, ed, n_buckets: int):
        """Sort edge arrays in place by (ei, ej) (pcb_sort_edges)."""
        if _is_dev(ei):
            _torch()
        mem = self._kind(ei, ej, ed)
        n = int(ei.shape[0])
        if mem == _lib.MEM_HOST:
            for a in (ei, ej, ed):
                if a.dtype != np.int32 or not a.flags.c_contiguous:
                    raise TypeError("host edge arrays must be contiguous int32 (sorted in place)")
        keep: list = []
        self._check(self.lib.pcb_sort_edges(self.h, mem, self._in(ei, np.int32, keep), self._in(ej, np.int32, keep),
                                            self._in(ed, np.int32, keep), n, int(n_buckets)))

    # ------------------------------------------------------------------ a4..a13
    def merge(self, p: MergeProblem, pairs: PairList, shard: Tuple[int, int] = (0, 1), bucket_seq=None,
              bucket_len=None, on_demand_k: int = -1, pairs_sorted: bool = False, shard_layout: Optional[bool] = None) -> Dict:
        """Seed clustering + greedy merge + consensus (pcb_merge).  Arrays inside `p` and `pairs`
        may be numpy or torch CUDA tensors.  bucket_seq/bucket_len/on_demand_k: score distances the
        pair list lacks on demand (see the header); never used for an EDFILE-driven run.
        pairs_sorted: q < r everywhere and ascending by (q, r) (pcb_edit_pairs / pcb_sort_edges order).
        shard_layout: call regions laid out identically on every shard, so that the shards' outputs can be combined
        with an element-wise max (default: only when shard[1] > 1)."""
        arrays = [p.bucket_ptr, p.bucket_reads, p.lexrank, p.bc_id, p.up_id, p.geno_ptr, p.geno_mut, p.geno_q,
                  p.mut_rank, pairs.q, pairs.r, pairs.d, pairs.ed]
        extra = [bucket_seq, bucket_len]
        if any(_is_dev(a) for a in arrays):
            _torch()
        mem = self._kind(*arrays, *extra)
        R = int(p.n_reads)
        nnz = int(p.geno_mut.shape[0])
        keep: list = []
        ins = [self._in(a, np.int32, keep) for a in arrays]
        p_bseq, p_blen = self._in(bucket_seq, np.uint8, keep), self._in(bucket_len, np.int32, keep)
        stride = int(bucket_seq.shape[1]) if bucket_seq is not None else 1
        outs, ptrs = {}, []
        for name, dt in MERGE_OUTPUTS:
            outs[name], ptr = self._out(mem, (max(R, 1),), dt, "m_" + name)
            ptrs.append(ptr)
        outs["call_mut"], p_call = self._out(mem, (max(nnz, 1),), np.int32, "m_call_mut")
        self.set_replay_shape()
        if shard_layout is None:
            shard_layout = int(shard[1]) > 1
        flags = (_lib.MERGE_PAIRS_SORTED if pairs_sorted else 0) | (_lib.MERGE_SHARD_LAYOUT if shard_layout else 0)
        self._check(self.lib.pcb_merge(self.h, mem, R, int(p.n_buckets), *ins[:8], nnz, ins[8], int(p.mut_rank.shape[0]),
                                       int(pairs.q.shape[0]), *ins[9:], int(p.max_diff), int(p.min_qual),
                                       float(p.min_jaccard), int(p.min_matches), p_bseq, p_blen, stride,
                                       int(on_demand_k) if bucket_seq is not None else -1, int(shard[0]), int(shard[1]), flags,
                                       *ptrs, p_call))
        out = {k: v[:R] for k, v in outs.items() if k != "call_mut"}
        out["call_mut"] = outs["call_mut"][:nnz]
        return out

    # ------------------------------------------------------------------ f1
    def tag_rows(self, row_bc, row_up, row_size, n_bc_ids: int, n_up_ids: int, min_ratio: float = 0.666) -> Dict:
        """Collision tags + soft-filter dominance test on a cluster table (pcb_tag_rows)."""
        if _is_dev(row_bc):
            _torch()
        mem = self._kind(row_bc, row_up, row_size)
        n = int(row_bc.shape[0])
        keep: list = []
        outs = {}
        ptrs = []
        for name in ("collision", "up_collision", "usable"):
            outs[name], ptr = self._out(mem, (max(n, 1),), np.uint8)
            ptrs.append(ptr)
        self._check(self.lib.pcb_tag_rows(self.h, mem, n, self._in(row_bc, np.int32, keep), self._in(row_up, np.int32, keep),
                                          self._in(row_size, np.int32, keep), int(n_bc_ids), int(n_up_ids), float(min_ratio), *ptrs))
        return {k: v[:n] for k, v in outs.items()}

    # ------------------------------------------------------------------ (e) one rank's share of a multi-GPU step
    def shard_pairs(self, seq, length, k: int, q_group: int, n_q_groups: int, t_group: int, n_t_groups: int):
        """Bucketing of the whole read set + this rank's block of the pair grid (pcb_shard_pairs), device tensors only.
        Returns (bucket_of_read, n_buckets, n_edges); the edges wait for shard_edges_into, the buckets for shard_merge."""
        torch = _torch()
        mem = self._kind(seq, length)
        self.set_replay_shape()
        R, stride = int(seq.shape[0]), int(seq.shape[1])
        keep: list = []
        bor, p_bor = self._out(mem, (max(R, 1),), np.int32)
        nb, ne = C.c_int32(0), C.c_int64(0)
        self._check(self.lib.pcb_shard_pairs(self.h, mem, self._in(seq, np.uint8, keep), self._in(length, np.int32, keep), R, stride, int(k),
                                             int(q_group), int(n_q_groups), int(t_group), int(n_t_groups), p_bor, C.byref(nb), C.byref(ne)))
        return bor[:R], int(nb.value), int(ne.value)

    def shard_edges_into(self, rows):
        """This rank's edges into the three rows of a [3, m] int32 device tensor (pcb_edit_pairs_fetch), m >= n_edges."""
        m = int(rows.shape[1])
        ptr = lambda r: C.c_void_p(rows[r].data_ptr())
        self._check(self.lib.pcb_edit_pairs_fetch(self.h, _lib.MEM_DEVICE, ptr(0), ptr(1), ptr(2), m))

    def shard_merge(self, gathered, counts, bucket_of_read, lexrank, up_id, geno_ptr, geno_mut, geno_q, mut_rank, max_diff: int,
                    min_qual: int, min_jaccard: float, min_matches: int, on_demand: bool, shard, shard_layout: bool = False):
        """The gathered edge lists ([world, 3, m] int32 device tensor, counts per rank) -> this shard's rows (pcb_shard_merge).
        Returns (outputs as merge(), total number of edges)."""
        torch = _torch()
        mem = self._kind(gathered, bucket_of_read, lexrank, up_id, geno_ptr, geno_mut, geno_q, mut_rank)
        world, m = int(gathered.shape[0]), int(gathered.shape[2])
        R, nnz = int(bucket_of_read.shape[0]), int(geno_mut.shape[0])
        keep: list = []
        cnt = np.ascontiguousarray(counts, dtype=np.int64)
        outs, ptrs = {}, []
        for name, dt in MERGE_OUTPUTS:
            outs[name], ptr = self._out(mem, (max(R, 1),), dt)
            ptrs.append(ptr)
        outs["call_mut"], p_call = self._out(mem, (max(nnz, 1),), np.int32)
        ne = C.c_int64(0)
        self._check(self.lib.pcb_shard_merge(self.h, mem, C.c_void_p(gathered.data_ptr()), world, m, cnt.ctypes.data_as(C.c_void_p),
                                             self._in(bucket_of_read, np.int32, keep), self._in(lexrank, np.int32, keep),
                                             self._in(up_id, np.int32, keep), self._in(geno_ptr, np.int32, keep),
                                             self._in(geno_mut, np.int32, keep), self._in(geno_q, np.int32, keep), nnz,
                                             self._in(mut_rank, np.int32, keep), int(mut_rank.shape[0]), int(max_diff), int(min_qual),
                                             float(min_jaccard), int(min_matches), 1 if on_demand else 0, int(shard[0]), int(shard[1]),
                                             _lib.MERGE_SHARD_LAYOUT if shard_layout else 0, C.byref(ne), *ptrs, p_call))
        out = {k: v[:R] for k, v in outs.items() if k != "call_mut"}
        out["call_mut"] = outs["call_mut"][:nnz]
        return out, int(ne.value)

    # ------------------------------------------------------------------ a14
    def concat_tokens(self, src: np.ndarray, tok_off: np.ndarray, tok_len: np.ndarray, item_tok: np.ndarray, item_sep: np.ndarray,
                      size_hint: Optional[int] = None) -> bytes:
        """Ragged concatenation on the device (pcb_concat_tokens): item k is src[tok_off[t] : tok_off[t] + tok_len[t]] for
        t = item_tok[k], followed by the byte item_sep[k] unless that is 0.  Host arrays in, bytes out."""
        keep: list = []
        n_src, n_tok, n_items = int(src.shape[0]), int(tok_off.shape[0]), int(item_tok.shape[0])
        cap = int(size_hint) if size_hint is not None else int(np.asarray(tok_len, dtype=np.int64)[np.asarray(item_tok, dtype=np.int64)].sum()
                                                                 + np.count_nonzero(item_sep))
        args = (self._in(src, np.uint8, keep), n_src, self._in(tok_off, np.int64, keep), self._in(tok_len, np.int32, keep), n_tok,
                self._in(item_tok, np.int32, keep), self._in(item_sep, np.uint8, keep), n_items)
        for _ in range(2):
            out = np.empty(max(cap, 1), dtype=np.uint8)
            n = C.c_int64(0)
            rc = self.lib.pcb_concat_tokens(self.h, _lib.MEM_HOST, *args, out.ctypes.data_as(C.c_void_p), cap, C.byref(n))
            if rc == _lib.PCB_E_CAPACITY and int(n.value) > cap:
                cap = int(n.value)
                continue
            self._check(rc)
            return out[:int(n.value)].tobytes()
        self._check(rc)

    # ------------------------------------------------------------------ f4
    def merge_libraries(self, vbc1, geno1, size1, vbc2, geno2, size2, n_ids: int) -> Dict:
        """Rows of a second cluster library against a first one (pcb_merge_libraries): per row of the second library its
        class (0 new, 1 merge, 2 conflict) and the row of the first it merges into; the first library's sizes afterwards."""
        arrays = [vbc1, geno1, size1, vbc2, geno2, size2]
        if any(_is_dev(a) for a in arrays):
            _torch()
        mem = self._kind(*arrays)
        n1, n2 = int(vbc1.shape[0]), int(vbc2.shape[0])
        keep: list = []
        ins = [self._in(a, np.int32, keep) for a in arrays]
        size_out, p_s = self._out(mem, (max(n1, 1),), np.int32)
        target, p_t = self._out(mem, (max(n2, 1),), np.int32)
        klass, p_k = self._out(mem, (max(n2, 1),), np.int32)
        self._check(self.lib.pcb_merge_libraries(self.h, mem, n1, ins[0], ins[1], ins[2], n2, ins[3], ins[4], ins[5], int(n_ids),
                                                 p_s, p_t, p_k))
        return dict(size1=size_out[:n1], target2=target[:n2], klass2=klass[:n2])

    # ------------------------------------------------------------------ the whole path
    def cluster_reads(self, seq, qual, length, lexrank, up_id, geno_ptr, geno_mut, geno_q, mut_rank,
                      max_diff: int = 2, min_qual: int = 32, min_jaccard: float = 0.2, min_matches: int = 1,
                      full_table: bool = False, compact_rows: bool = False) -> Dict:
        """Bucketing -> distances -> merge on the device (pcb_cluster_reads).  full_table=True
        materialises every pair up to 2*max_diff (the calcEdits table) instead of scoring the
        transitive-rule distances on demand; the clusters are the same.  compact_rows=True returns the
        row_* arrays with one entry per row (plus `row_rep`, the representative read of each row, and a
        call_mut holding only the rows' calls) instead of one per read: with host arrays only the rows
        cross the bus."""
        arrays = [seq, qual, length, lexrank, up_id, geno_ptr, geno_mut, geno_q, mut_rank]
        if any(_is_dev(a) for a in arrays):
            _torch()
        mem = self._kind(*arrays)
        R, stride = int(seq.shape[0]), int(seq.shape[1])
        nnz = int(geno_mut.shape[0])
        keep: list = []
        dts = [np.uint8, np.uint8] + [np.int32] * 7
        ins = [self._in(a, dt, keep) for a, dt in zip(arrays, dts)]
        bor, p_bor = self._out(mem, (max(R, 1),), np.int32, "c_bor")
        outs, ptrs = {}, []
        for name, dt in MERGE_OUTPUTS:
            outs[name], ptr = self._out(mem, (max(R, 1),), dt, "c_" + name)
            ptrs.append(ptr)
        outs["call_mut"], p_call = self._out(mem, (max(nnz, 1),), np.int32, "c_call_mut")
        nb, ne = C.c_int32(0), C.c_int64(0)
        rep, p_rep = self._out(mem, (max(R, 1),), np.int32, "c_row_rep") if compact_rows else (None, None)
        flags = (_lib.CLUSTER_FULL_TABLE if full_table else 0) | (_lib.CLUSTER_COMPACT_ROWS if compact_rows else 0)
        self.set_replay_shape()
        self._check(self.lib.pcb_cluster_reads(self.h, mem, ins[0], ins[1], ins[2], R, stride, *ins[3:8], nnz, ins[8],
                                               int(mut_rank.shape[0]), int(max_diff), int(min_qual),
                                               float(min_jaccard), int(min_matches), flags, p_bor, C.byref(nb), C.byref(ne),
                                               *ptrs, p_call, p_rep))
        if compact_rows:
            n_rows, n_calls = self.stat("rows"), self.stat("calls")
            out = {k: v[:(R if k in ("read_root", "read_pos") else n_rows)] for k, v in outs.items() if k != "call_mut"}
            out["call_mut"] = outs["call_mut"][:n_calls]
            out["row_rep"] = rep[:n_rows]
        else:
            out = {k: v[:R] for k, v in outs.items() if k != "call_mut"}
            out["call_mut"] = outs["call_mut"][:nnz]
        out["bucket_of_read"] = bor[:R]
        out["n_buckets"] = int(nb.value)
        out["n_edges"] = int(ne.value)
        return out

    # ------------------------------------------------------------------ f3
    def match_barcodes(self, lib_seq, lib_len, q_seq, q_len, max_err: int, want_pairs: bool = False) -> Dict:
        """Nearest library barcodes of every query within max_err (pcb_match_barcodes).  want_pairs=True also returns
        every (query, row, d) within max_err, sorted by (query, row)."""
        arrays = [lib_seq, lib_len, q_seq, q_len]
        if any(_is_dev(a) for a in arrays):
            _torch()
        mem = self._kind(*arrays)
        L, Q = int(lib_len.shape[0]), int(q_len.shape[0])
        ls = int(lib_seq.shape[1]) if L else 1
        qs = int(q_seq.shape[1]) if Q else 1
        keep: list = []
        best, p_best = self._out(mem, (max(Q, 1),), np.int32)
        nb, p_nb = self._out(mem, (max(Q, 1),), np.int32)
        first, p_first = self._out(mem, (max(Q, 1),), np.int32)
        n_pairs = C.c_int64(0)
        self._check(self.lib.pcb_match_barcodes(self.h, mem, self._in(lib_seq, np.uint8, keep), self._in(lib_len, np.int32, keep), L, ls,
                                                self._in(q_seq, np.uint8, keep), self._in(q_len, np.int32, keep), Q, qs, int(max_err),
                                                p_best, p_nb, p_first, C.byref(n_pairs)))
        out = dict(best_d=best[:Q], n_best=nb[:Q], first_hit=first[:Q], n_pairs=int(n_pairs.value))
        if want_pairs:
            n = out["n_pairs"]
            bufs = [self._out(mem, (max(n, 1),), np.int32) for _ in range(3)]
            self._check(self.lib.pcb_edit_pairs_fetch(self.h, mem, bufs[0][1], bufs[1][1], bufs[2][1], n))
            out["pair_q"], out["pair_row"], out["pair_d"] = (b[0][:n] for b in bufs)
        return out

    def compact_shard(self, out: Dict, to_host: bool = True) -> Dict:
        """The reads this shard owns and their rows, compacted (pcb_compact_shard).  `out`: per-read outputs of
        merge() / cluster_reads() (numpy or torch CUDA).  to_host=True returns numpy arrays (page-locked and reused when
        the context has pinned_outputs): from device inputs only the compacted entries cross the bus."""
        names = [n for n, _ in MERGE_OUTPUTS] + ["call_mut"]
        arrays = [out[n] for n in names]
        mem_in = self._kind(*arrays)
        mem_out = _lib.MEM_HOST if to_host else _lib.MEM_DEVICE
        if not to_host:
            _torch()
        R, nnz = int(out["read_root"].shape[0]), int(out["call_mut"].shape[0])
        keep: list = []
        dts = dict(MERGE_OUTPUTS)
        ins = [self._in(out[n], dts.get(n, np.int32), keep) for n in names]
        bor = out.get("bucket_of_read")
        if bor is not None and self._kind(bor) != mem_in:
            raise TypeError("bucket_of_read must live where the other arrays do")
        ins.append(self._in(bor, np.int32, keep))
        res, ptrs = {}, {}
        for n, dt, size in (("own_read", np.int32, R), ("own_root", np.int32, R), ("own_pos", np.int32, R), ("own_bucket", np.int32, R),
                            ("row_rep", np.int32, R),
                            ("row_size", np.int32, R), ("row_key", np.int64, R), ("row_bc", np.int32, R), ("row_up", np.int32, R),
                            ("row_call_off", np.int64, R), ("row_call_n", np.int32, R), ("call_mut", np.int32, nnz)):
            res[n], ptrs[n] = self._out(mem_out, (max(size, 1),), dt, "s_" + n)
        n_own, n_rows, n_calls = C.c_int32(0), C.c_int32(0), C.c_int64(0)
        self._check(self.lib.pcb_compact_shard(self.h, mem_in, mem_out, R, nnz, *ins, C.byref(n_own), ptrs["own_read"], ptrs["own_root"],
                                               ptrs["own_pos"], ptrs["own_bucket"] if bor is not None else None, C.byref(n_rows), C.byref(n_calls), ptrs["row_rep"], ptrs["row_size"],
                                               ptrs["row_key"], ptrs["row_bc"], ptrs["row_up"], ptrs["row_call_off"], ptrs["row_call_n"],
                                               ptrs["call_mut"]))
        cut = dict(own_read=n_own.value, own_root=n_own.value, own_pos=n_own.value, own_bucket=n_own.value, call_mut=n_calls.value)
        res = {k: v[:cut.get(k, n_rows.value)] for k, v in res.items() if not (k == "own_bucket" and bor is None)}
        res["n_reads"] = R
        return res


def open_context_or_exit(device: int = 0, **kw) -> Context:
    """For the command-line programs: a Context, or one line on stderr and exit status 1 (there is no CPU path)."""
    import sys
    try:
        return Context(device, **kw)
    except _lib.PcbError as e:
        print(f"ERROR: {e}", file=sys.stderr)
        sys.exit(1)


class Text:
    """The bytes of one text file on the device plus its newline index (pcb_text; f2 tokenizers).

    `data`: bytes / bytearray / numpy uint8 (copied to the device) or a torch CUDA uint8 tensor (used in place and kept
    alive).  `device=True` in the fetch methods returns torch CUDA tensors, else numpy arrays."""

    def __init__(self, ctx: Context, data):
        self.ctx = ctx
        self.h = None
        self._keep = data
        if _is_dev(data):
            _torch()
            mem = ctx._kind(data)
            if data.dtype != _NP2TORCH[np.dtype(np.uint8)]:
                raise TypeError("device text must be a uint8 tensor")
            n, ptr = int(data.numel()), C.c_void_p(data.data_ptr())
        else:
            mem = _lib.MEM_HOST
            arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
            self._keep = arr
            n, ptr = int(arr.shape[0]), arr.ctypes.data_as(C.c_void_p)
        h, nl, ex = C.c_void_p(), C.c_int64(0), C.c_int64(0)
        ctx._check(ctx.lib.pcb_text_open(ctx.h, mem, ptr, n, C.byref(h), C.byref(nl), C.byref(ex)))
        self.h = h
        self.n_bytes, self.n_lines, self.n_exotic = n, int(nl.value), int(ex.value)
        if mem == _lib.MEM_HOST:
            self._keep = None          # the library holds its own copy

    def close(self):
        if self.h is not None and getattr(self.ctx, "h", None):
            self.ctx.lib.pcb_text_close(self.ctx.h, self.h)
        self.h = None
        self._keep = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _mem(self, device: bool) -> int:
        if device:
            _torch()
        return _lib.MEM_DEVICE if device else _lib.MEM_HOST

    def fastq_records(self) -> Dict:
        n, ml, mi, bad = C.c_int32(0), C.c_int32(0), C.c_int32(0), C.c_int64(0)
        self.ctx._check(self.ctx.lib.pcb_fastq_records(self.ctx.h, self.h, C.byref(n), C.byref(ml), C.byref(mi), C.byref(bad)))
        self.n_records = int(n.value)
        return dict(n_records=int(n.value), max_len=int(ml.value), max_id_len=int(mi.value), n_bad=int(bad.value))

    def fastq_fetch(self, stride: int, device: bool = True) -> Dict:
        mem, R, o = self._mem(device), self.n_records, self.ctx._out
        seq, p_seq = o(mem, (R, stride), np.uint8)
        qual, p_qual = o(mem, (R, stride), np.uint8)
        length, p_len = o(mem, (R,), np.int32)
        id_off, p_off = o(mem, (R,), np.int64)
        id_len, p_idl = o(mem, (R,), np.int32)
        self.ctx._check(self.ctx.lib.pcb_fastq_fetch(self.ctx.h, self.h, mem, int(stride), p_seq, p_qual, p_len, p_off, p_idl))
        return dict(seq=seq, qual=qual, length=length, id_off=id_off, id_len=id_len)

    def rank_strings(self, off, length, max_len: int):
        mem = self.ctx._kind(off, length)
        keep: list = []
        n = int(off.shape[0])
        rank, p_rank = self.ctx._out(mem, (n,), np.int32)
        self.ctx._check(self.ctx.lib.pcb_rank_strings(self.ctx.h, self.h, mem, self.ctx._in(off, np.int64, keep),
                                                      self.ctx._in(length, np.int32, keep), n, int(max_len), p_rank))
        return rank

    def _join(self, fn, fastq: "Text", id_off, id_len):
        mem = self.ctx._kind(id_off, id_len)
        keep: list = []
        R = int(id_off.shape[0])
        a, b = C.c_int64(0), C.c_int32(0)
        counters = (C.c_int64 * 4)()
        self.ctx._check(fn(self.ctx.h, self.h, fastq.h, mem, self.ctx._in(id_off, np.int64, keep), self.ctx._in(id_len, np.int32, keep),
                           R, C.byref(a), C.byref(b), counters))
        self.n_reads = R
        return int(a.value), int(b.value), [int(c) for c in counters]

    def geno_join(self, fastq: "Text", id_off, id_len) -> Dict:
        self.nnz, self.n_distinct, c = self._join(self.ctx.lib.pcb_geno_join, fastq, id_off, id_len)
        return dict(nnz=self.nnz, n_mut=self.n_distinct, malformed=c[0], missing=c[1], repeated=c[2], collisions=c[3])

    def geno_fetch(self, device: bool = True) -> Dict:
        mem, o = self._mem(device), self.ctx._out
        gp, p_gp = o(mem, (self.n_reads + 1,), np.int32)
        gm, p_gm = o(mem, (self.nnz,), np.int32)
        gq, p_gq = o(mem, (self.nnz,), np.int32)
        mo, p_mo = o(mem, (self.n_distinct,), np.int64)
        ml, p_ml = o(mem, (self.n_distinct,), np.int32)
        self.ctx._check(self.ctx.lib.pcb_geno_fetch(self.ctx.h, self.h, mem, p_gp, p_gm, p_gq, p_mo, p_ml))
        return dict(geno_ptr=gp, geno_mut=gm, geno_q=gq, mut_off=mo, mut_len=ml)

    def uptag_join(self, fastq: "Text", id_off, id_len) -> Dict:
        n_rec, self.n_distinct, c = self._join(self.ctx.lib.pcb_uptag_join, fastq, id_off, id_len)
        return dict(n_records=n_rec, n_tags=self.n_distinct, too_long=c[0], collisions=c[3])

    def uptag_fetch(self, device: bool = True) -> Dict:
        mem, o = self._mem(device), self.ctx._out
        up, p_up = o(mem, (self.n_reads,), np.int32)
        to, p_to = o(mem, (self.n_distinct,), np.int64)
        tl, p_tl = o(mem, (self.n_distinct,), np.int32)
        self.ctx._check(self.ctx.lib.pcb_uptag_fetch(self.ctx.h, self.h, mem, p_up, p_to, p_tl))
        return dict(up_id=up, tag_off=to, tag_len=tl)
