"""ctypes binding of include/pacybara_b200.h.  There is no fallback: if the CUDA library is
missing or no device is present, importing works but every use raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("PCB_LIB") or os.path.join(HERE, "_pcb.so")     # PCB_LIB: a differently built library (experiments)

PCB_OK, PCB_E_CUDA, PCB_E_INPUT, PCB_E_CAPACITY, PCB_E_STATE, PCB_E_UPTAG = range(6)
MEM_HOST, MEM_DEVICE = 0, 1
FILL = -(1 << 31)
NEEDS_ALIGNMENT = -2
CLUSTER_FULL_TABLE = 1
CLUSTER_COMPACT_ROWS = 2
MERGE_PAIRS_SORTED = 1
MERGE_SHARD_LAYOUT = 2
STAT = dict(pairs_scored=0, cells=1, edges=2, seed_len=3, kernel_launches=4, components=5, links=6, tests=7,
            buckets=8, probe_ns=9, t_bucket_ns=10, t_index_ns=11, t_hits_ns=12, t_prep_ns=13, t_replay_ns=14,
            t_scatter_ns=15, rows=16, calls=17, candidates=18, huge_components=19, huge_rounds=20)

# every symbol include/pacybara_b200.h declares
REPLAY_SHAPES = dict(auto=0, small=0, thread=1, warp=2, huge=3)
OPT_REPLAY_SHAPE, OPT_TRACE, OPT_HUGE_READS, OPT_HUGE_CAP, OPT_HUGE_WINDOW, OPT_PROBE_STAGING, OPT_PROBE_CLASSES, OPT_SEED_BINS_LOG2, OPT_SMALL_WARPS, OPT_HUGE_BUCKETS = range(10)

SYMBOLS = ["pcb_abi_version", "pcb_last_error", "pcb_ctx_create", "pcb_ctx_destroy", "pcb_stat", "pcb_set_option", "pcb_sync",
           "pcb_stream", "pcb_bucket", "pcb_edit_pairs", "pcb_edit_pairs_fetch", "pcb_sort_edges", "pcb_merge", "pcb_cluster_reads", "pcb_compact_shard", "pcb_match_barcodes", "pcb_tag_rows", "pcb_merge_libraries", "pcb_concat_tokens", "pcb_shard_pairs", "pcb_shard_merge",
           "pcb_text_open", "pcb_text_close", "pcb_fastq_records", "pcb_fastq_fetch", "pcb_rank_strings", "pcb_geno_join",
           "pcb_geno_fetch", "pcb_uptag_join", "pcb_uptag_fetch"]


class PcbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pacybara_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """The CDLL, loaded once.  Raises if the library has not been built (python -m pacybara_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise PcbError(PCB_E_CUDA, f"{SO} is missing: build it with `python -m pacybara_b200.build` "
                                       "(there is no CPU implementation to fall back to)")
        lib = C.CDLL(SO)
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        lib.pcb_abi_version.restype = C.c_int
        lib.pcb_last_error.restype = C.c_char_p
        lib.pcb_last_error.argtypes = [vp]
        lib.pcb_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        lib.pcb_ctx_destroy.argtypes = [vp]
        lib.pcb_ctx_destroy.restype = None
        lib.pcb_stat.restype = i64
        lib.pcb_stat.argtypes = [vp, C.c_int]
        lib.pcb_set_option.argtypes = [vp, C.c_int, i64]
        lib.pcb_sync.argtypes = [vp]
        lib.pcb_stream.restype = vp
        lib.pcb_stream.argtypes = [vp]
        lib.pcb_bucket.argtypes = [vp, C.c_int, vp, vp, vp, i32, i32, vp, C.POINTER(i32), vp, vp, vp]
        lib.pcb_edit_pairs.argtypes = [vp, C.c_int, vp, vp, i32, i32, i32, i32, i32, i32, i32, C.POINTER(i64)]
        lib.pcb_edit_pairs_fetch.argtypes = [vp, C.c_int, vp, vp, vp, i64]
        lib.pcb_sort_edges.argtypes = [vp, C.c_int, vp, vp, vp, i64, i32]
        lib.pcb_merge.argtypes = ([vp, C.c_int, i32, i32] + [vp] * 8 + [i64, vp, i32, i64] + [vp] * 4 +
                                  [i32, i32, dbl, i32, vp, vp, i32, i32, i32, i32, i32] + [vp] * 9)
        lib.pcb_cluster_reads.argtypes = ([vp, C.c_int, vp, vp, vp, i32, i32] + [vp] * 5 + [i64, vp, i32, i32, i32, dbl, i32, i32] +
                                          [vp, C.POINTER(i32), C.POINTER(i64)] + [vp] * 10)
        lib.pcb_compact_shard.argtypes = ([vp, C.c_int, C.c_int, i32, i64] + [vp] * 10 + [C.POINTER(i32), vp, vp, vp, vp] +
                                          [C.POINTER(i32), C.POINTER(i64)] + [vp] * 8)
        lib.pcb_match_barcodes.argtypes = [vp, C.c_int, vp, vp, i32, i32, vp, vp, i32, i32, i32, vp, vp, vp, C.POINTER(i64)]
        lib.pcb_tag_rows.argtypes = [vp, C.c_int, i32, vp, vp, vp, i32, i32, dbl, vp, vp, vp]
        lib.pcb_merge_libraries.argtypes = [vp, C.c_int, i32, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp]
        lib.pcb_shard_pairs.argtypes = [vp, C.c_int, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, C.POINTER(i32), C.POINTER(i64)]
        lib.pcb_shard_merge.argtypes = [vp, C.c_int, vp, i32, i64, vp, vp, vp, vp, vp, vp, vp, i64, vp, i32, i32, i32, dbl, i32, i32, i32, i32,
                                        i32, C.POINTER(i64)] + [vp] * 9
        lib.pcb_concat_tokens.argtypes = [vp, C.c_int, vp, i64, vp, vp, i64, vp, vp, i64, vp, i64, C.POINTER(i64)]
        p32, p64 = C.POINTER(i32), C.POINTER(i64)
        lib.pcb_text_open.argtypes = [vp, C.c_int, vp, i64, C.POINTER(vp), p64, p64]
        lib.pcb_text_close.argtypes = [vp, vp]
        lib.pcb_text_close.restype = None
        lib.pcb_fastq_records.argtypes = [vp, vp, p32, p32, p32, p64]
        lib.pcb_fastq_fetch.argtypes = [vp, vp, C.c_int, i32, vp, vp, vp, vp, vp]
        lib.pcb_rank_strings.argtypes = [vp, vp, C.c_int, vp, vp, i32, i32, vp]
        lib.pcb_geno_join.argtypes = [vp, vp, vp, C.c_int, vp, vp, i32, p64, p32, p64]
        lib.pcb_geno_fetch.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, vp]
        lib.pcb_uptag_join.argtypes = [vp, vp, vp, C.c_int, vp, vp, i32, p64, p32, p64]
        lib.pcb_uptag_fetch.argtypes = [vp, vp, C.c_int, vp, vp, vp]
        _lib = lib
    return _lib
