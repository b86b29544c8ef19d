import sys; sys.path.insert(0, '.')
import ctypes as C
import numpy as np
from pacybara_b200 import synth, pipeline, hostdata
from pacybara_b200.api import Context
from tests import util
ctx = Context(0)
lib = synth.make_config("cfg1", seed=1)
ra, names = pipeline.arrays_from_library(lib, ctx)
out = pipeline.cluster_reads(ctx, ra)
E = out["n_edges"]
ei = np.empty(E, np.int32); ej = np.empty(E, np.int32); ed = np.empty(E, np.int32)
rc = ctx.lib.pcb_edit_pairs_fetch(ctx.h, 0, ei.ctypes.data_as(C.c_void_p), ej.ctypes.data_as(C.c_void_p), ed.ctypes.data_as(C.c_void_p), E)
print("fused edges", E, rc, "U", out["n_buckets"])
b = util.oracle_buckets(hostdata.FastqReads([], lib.seq, lib.qual, lib.length))
si, sj, sd = ctx.edit_pairs(b["bucket_seq"], b["bucket_len"], 4)
print("standalone edges", len(si), "U", b["n_buckets"])
A = set(zip(ei.tolist(), ej.tolist(), ed.tolist())); B = set(zip(si.tolist(), sj.tolist(), sd.tolist()))
print("only fused", len(A - B), "only standalone", len(B - A))
seqs = hostdata.matrix_to_strings(b["bucket_seq"], b["bucket_len"])
from oracle import orc
for (i, j, d) in sorted(A - B)[:15]:
    print(i, j, d, seqs[i], seqs[j], orc.lev(seqs[i].encode(), seqs[j].encode()))
print(np.bincount(ed), np.bincount(sd))
