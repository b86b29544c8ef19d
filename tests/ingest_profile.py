"""(muscle is not on the GPU box: tied barcode votes take the first member's barcode here.)
Where the time of the one-call CLI goes, files -> clusters.csv.gz, at a given config (run on the GPU box).

    python tests/ingest_profile.py [cfg2] [scale]   -> one JSON line (seconds per stage)
"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pacybara_b200 import hostdata, ingest, pipeline, synth   # noqa: E402
from pacybara_b200.api import Context                         # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    lib = synth.make_config(cfg, seed=1, scale=scale)
    d = tempfile.mkdtemp(prefix="pcb_ingest_")
    paths = lib.write_files(d)
    ctx = Context(0)
    t = {}

    def lap(name, t0):
        t[name] = round(time.perf_counter() - t0, 4)

    for rep in ("warm", "timed"):
        t0 = time.perf_counter()
        same = paths["uptag"] == paths["bc"]
        raw = ingest.read_all([paths["bc"], paths["geno"], None if same else paths["uptag"]])
        if same:
            raw[2] = raw[0]
        lap("gunzip", t0)
        t0 = time.perf_counter()
        laps = {}
        ra, names = ingest.arrays_from_bytes(ctx, raw[0], raw[1], raw[2], laps=laps)
        lap("tokenize", t0)
        t["tokenize_steps"] = laps
        t0 = time.perf_counter()
        out = pipeline.cluster_reads(ctx, ra)
        lap("cluster", t0)
        t0 = time.perf_counter()
        table = pipeline.table_from_outputs(out, ra, names, consensus=lambda bcs: bcs[0])
        lap("table", t0)
        t0 = time.perf_counter()
        hostdata.write_clusters_csv(os.path.join(d, "clusters.csv.gz"), table)
        lap("write", t0)
    t0 = time.perf_counter()
    with hostdata.open_text(paths["bc"]) as fq, hostdata.open_text(paths["geno"]) as ge, hostdata.open_text(paths["uptag"]) as up:
        ra2, names2 = pipeline.arrays_from_text(fq, ge, up, ctx)
    lap("line_parser", t0)
    t0 = time.perf_counter()
    table2 = pipeline.table_from_outputs(pipeline.cluster_reads(ctx, ra2), ra2, names2, consensus=lambda bcs: bcs[0])
    lap("line_parser_cluster_table", t0)
    t["same_table"] = table == table2
    t["reads"] = ra.n_reads
    t["rows"] = len(table)
    t["bytes"] = [len(b) for b in raw if b is not None]
    print(json.dumps(dict(config=cfg, scale=scale, seconds=t)))


if __name__ == "__main__":
    main()
