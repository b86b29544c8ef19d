"""Parity at the sizes BASELINE.json names, through the C-ABI, on the GPU box.

cfg 2 (the bench configuration), cfg 4 and cfg 5 (the 2*10^7-read north-star target) run at FULL size: the oracle's own
pair search is out of reach there (hours), so the GPU's complete distance table is spot-checked against the DP and fed
to the oracle's literal restatement of pacybara_cluster.py, whose rows (order, members, consensus, calls) must equal the
GPU's for the full-table run and for the default on-demand-distance run (tests/util.check_full_size_against_oracle_merge).
cfg 3 and cfg 4 additionally run scaled against the complete oracle (its own bucketing + pair search + merge), in both
replay launch shapes and both distance modes.  Reference semantics held: src/pacybara_cluster.py:385-455,546-567.
"""
import os

import numpy as np
import pytest

from pacybara_b200 import canon, hostdata, pipeline, synth
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from pacybara_b200.api import Context
    c = Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("cfg", ["cfg2", "cfg4", "cfg5"])
def test_full_size_rows_equal_the_oracle_merge(ctx, cfg):
    lib = synth.make_config(cfg, seed=1)
    res = util.check_full_size_against_oracle_merge(ctx, lib)
    assert res["reads"] == synth.CONFIGS[cfg]["gen"]["n_reads"]
    assert res["clusters"] > 0.9 * synth.CONFIGS[cfg]["gen"]["n_clones"] * (0.9 if cfg == "cfg4" else 1.0)
    assert res["edges_on_demand"] <= res["edges_full"]


@pytest.mark.parametrize("cfg,scale", [("cfg3", 0.1), ("cfg4", 0.05)])
def test_scaled_configs_against_the_complete_oracle(ctx, cfg, scale):
    lib = synth.make_config(cfg, seed=1, scale=scale)
    ra, names = pipeline.arrays_from_library(lib, ctx)
    p = dict(max_diff=lib.params["max_diff"], min_qual=lib.params["min_qual"], min_jaccard=0.2, min_matches=1)
    want = canon.canon_table(util.oracle_table_for_library(lib, ra, names, **p))
    for shape in ("thread", "warp", "huge", None):
        if shape:
            os.environ["PCB_REPLAY"] = shape
        try:
            for full in (False, True):
                out = pipeline.cluster_reads(ctx, ra, full_table=full, **p)
                got = canon.canon_table(pipeline.table_from_outputs(out, ra, names))
                assert got == want, (shape, full, canon.diff_tables(got, want))
        finally:
            os.environ.pop("PCB_REPLAY", None)


def test_dense_barcode_space_giant_component(ctx):
    """templates/pacybara_parameters.txt:13 -- a 25-nt SWSW... alphabet: at 10^6 reads most buckets hang together in one
    component of the link graph.  It is replayed by the whole device (csrc/merge_huge.cuh: rounds of deterministic
    reservations); the rows must be those of the oracle's sequential restatement, with the default knobs and with a
    scratch / window small enough that rows wait for the big scratch and are carried from round to round."""
    lib = synth.make_config("dense", seed=1, scale=0.1)
    res = util.check_full_size_against_oracle_merge(ctx, lib)
    assert ctx.stat("huge_components") >= 1 and ctx.stat("huge_rounds") >= 2
    assert res["reads"] == lib.n_reads
    os.environ.update(PCB_HUGE_CAP="128", PCB_HUGE_WINDOW="4096", PCB_HUGE_READS="64")
    try:
        util.check_full_size_against_oracle_merge(ctx, lib, n_spot=100)
    finally:
        for k in ("PCB_HUGE_CAP", "PCB_HUGE_WINDOW", "PCB_HUGE_READS"):
            os.environ.pop(k, None)
