"""The probe of the pair search (csrc/probe.cuh: seed-index candidates, composition filter, Myers recurrence with the
early exit, lanes that each carry their own pair) WITHOUT a GPU: tests/host_emul runs the source the device executes on
emulated warps, and every pair within k must come out with the distance of the exhaustive DP -- no more, no less.
Unit tests of the kernel source; the CUDA launch path is covered by the -m gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import orc
from pacybara_b200 import hostdata, synth
from tests import util

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul")
CSRC = os.path.join(os.path.dirname(HERE), "..", "pacybara_b200", "csrc")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "libprobe_emul.so")
        deps = [os.path.join(HERE, f) for f in ("probe_emul.cpp", "simt_emul.cpp", "simt_emul.h")] + \
               [os.path.join(CSRC, f) for f in ("probe.cuh", "simt.h", "myers.cuh")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.run([cxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I", CSRC, "-I", HERE, "-o", so,
                            os.path.join(HERE, "probe_emul.cpp"), os.path.join(HERE, "simt_emul.cpp")], check=True)
        _LIB = C.CDLL(so)
        _LIB.emul_probe.restype = C.c_int64
    return _LIB


def probe(bseq, blen, k, q_range=None, n_warps=3, seed=1):
    U = int(blen.shape[0])
    planes = util.pack_planes(bseq, blen)
    cap = U * U + 4096
    ei, ej, ed = (np.zeros(cap, dtype=np.int32) for _ in range(3))
    stats = np.zeros(6, dtype=np.uint64)
    lo, hi = q_range or (0, U)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    n = lib().emul_probe(ptr(planes), ptr(np.ascontiguousarray(blen, dtype=np.int32)), C.c_int32(U), C.c_int32(k), C.c_int32(lo),
                         C.c_int32(hi), C.c_int32(n_warps), C.c_uint32(seed), ptr(ei), ptr(ej), ptr(ed), C.c_int64(cap), ptr(stats))
    assert n >= 0, n
    return (ei[:n], ej[:n], ed[:n]), dict(zip(("hits", "started", "cells", "blocks", "walked", "sided"), stats.tolist()))


@pytest.mark.parametrize("k,bc_len,p_err,p_near", [(0, 25, 0.02, 0.3), (1, 25, 0.03, 0.4), (2, 25, 0.03, 0.4), (4, 25, 0.03, 0.4),
                                                    (2, 12, 0.05, 0.5), (3, 8, 0.05, 0.5), (4, 3, 0.1, 0.3), (6, 25, 0.04, 0.5)])
def test_probe_finds_exactly_the_pairs_within_k(k, bc_len, p_err, p_near):
    lib_ = synth.make_library(n_clones=80, n_reads=900, bc_len=bc_len, p_err=p_err, p_near=p_near, seed=40 + k)
    b = util.oracle_buckets(hostdata.FastqReads([], lib_.seq, lib_.qual, lib_.length))
    got, st = probe(b["bucket_seq"], b["bucket_len"], k, seed=k + 1)
    want = orc.edit_pairs(b["bucket_seq"], b["bucket_len"], k, method="brute")
    for g, w in zip(got, want):
        assert np.array_equal(g, np.asarray(w))
    assert st["started"] <= st["sided"] <= st["walked"]


def test_query_shards_are_disjoint_and_complete():
    lib_ = synth.make_library(n_clones=100, n_reads=1200, p_err=0.03, p_near=0.4, seed=77)
    b = util.oracle_buckets(hostdata.FastqReads([], lib_.seq, lib_.qual, lib_.length))
    U = b["n_buckets"]
    want = orc.edit_pairs(b["bucket_seq"], b["bucket_len"], 2, method="brute")
    parts = [probe(b["bucket_seq"], b["bucket_len"], 2, q_range=r, seed=5)[0] for r in ((0, U // 3), (U // 3, U // 2), (U // 2, U))]
    assert sum(len(p[0]) for p in parts) == len(want[0])
    ei, ej, ed = (np.concatenate([p[t] for p in parts]) for t in range(3))
    order = np.lexsort((ej, ei))
    for g, w in zip((ei[order], ej[order], ed[order]), want):
        assert np.array_equal(g, np.asarray(w))


def test_the_filters_and_the_early_exit_do_work():
    """On random 25-mers most candidates die in the composition filter or at the first checkpoint."""
    rng = np.random.default_rng(3)
    mat = synth.BASES[rng.integers(0, 4, size=(3000, 25))].astype(np.uint8)
    got, st = probe(mat, np.full(3000, 25, dtype=np.int32), 4, n_warps=2)
    want = orc.edit_pairs(mat, np.full(3000, 25, dtype=np.int32), 4, method="brute")
    for g, w in zip(got, want):
        assert np.array_equal(g, np.asarray(w))
    assert st["started"] < 0.7 * st["sided"]                     # composition + length filter
    assert st["cells"] < 0.7 * st["started"] * 25 * 25           # early exit


@pytest.mark.parametrize("k,bc_len", [(2, 25), (4, 25), (3, 8), (4, 3)])
def test_staged_walk_of_long_bins_finds_the_same_pairs(k, bc_len):
    """probe_warp<STAGED>: bins walked through the two staging buffers (cp.async.bulk + mbarrier on the device, memcpy in the
    emulator) -- short seeds make every bin long, so the chunking (odd bin starts, chunks of 256 entries, bins that end
    inside a chunk, the refill order) is what is exercised."""
    lib_ = synth.make_library(n_clones=100, n_reads=1200, bc_len=bc_len, p_err=0.04, p_near=0.4, seed=60 + k)
    b = util.oracle_buckets(hostdata.FastqReads([], lib_.seq, lib_.qual, lib_.length))
    got, st = probe(b["bucket_seq"], b["bucket_len"], k, seed=0x80000000 | (k + 1))
    plain, st0 = probe(b["bucket_seq"], b["bucket_len"], k, seed=k + 1)
    want = orc.edit_pairs(b["bucket_seq"], b["bucket_len"], k, method="brute")
    for g, w in zip(got, want):
        assert np.array_equal(g, np.asarray(w))
    assert st["walked"] == st0["walked"] and st["sided"] == st0["sided"]      # the same entries pass by, staged or streamed


@pytest.mark.parametrize("k,bc_len,staged", [(2, 25, 0), (2, 25, 1), (4, 25, 0), (3, 8, 1)])
def test_bins_split_by_target_class_find_the_same_pairs(k, bc_len, staged):
    """owns_pair with classes (big libraries): the query side of a pair is the smaller (class, index); a probe starts its walk
    at its own class.  Same pairs, fewer entries walked; query shards stay disjoint and complete."""
    lib_ = synth.make_library(n_clones=100, n_reads=1200, bc_len=bc_len, p_err=0.04, p_near=0.4, seed=80 + k)
    b = util.oracle_buckets(hostdata.FastqReads([], lib_.seq, lib_.qual, lib_.length))
    U = b["n_buckets"]
    flag = 0x40000000 | (0x80000000 if staged else 0)
    got, st = probe(b["bucket_seq"], b["bucket_len"], k, seed=flag | (k + 1))
    plain, st0 = probe(b["bucket_seq"], b["bucket_len"], k, seed=k + 1)
    want = orc.edit_pairs(b["bucket_seq"], b["bucket_len"], k, method="brute")
    for g, w in zip(got, want):
        assert np.array_equal(g, np.asarray(w))
    assert st["walked"] < 0.75 * st0["walked"]
    parts = [probe(b["bucket_seq"], b["bucket_len"], k, q_range=r, seed=flag | 9)[0] for r in ((0, U // 3), (U // 3, U))]
    assert sum(len(p[0]) for p in parts) == len(want[0])
    whole, st1 = probe(b["bucket_seq"], b["bucket_len"], k, seed=flag | 0x20000000 | 3)      # class rule over whole bins
    for g, w in zip(whole, want):
        assert np.array_equal(g, np.asarray(w))
    assert st1["walked"] == st0["walked"]
