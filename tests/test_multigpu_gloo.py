"""The N>1 host logic on CPU: 2 processes, gloo backend (the GPU-side stages are replaced by the
oracle / the host emulation of the replay, the collectives and the sharding rules are the real ones)."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, os.environ["PCB_ROOT"])
    import numpy as np, torch, torch.distributed as dist
    from pacybara_b200 import hostdata, multigpu
    from oracle import orc
    from tests import util
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    case = util.load_golden(os.environ["PCB_CASE"])
    job = util.job_from_golden(case)
    p = job.problem
    # 2. my slice of the query buckets: the pairs whose query side (csrc/myers.cuh is_query_side) is mine
    first = p.bucket_reads[p.bucket_ptr[:-1]]
    mat, lens = hostdata.strings_to_matrix([job.bc_names[i].encode() for i in p.bc_id[first]], 64)
    ei, ej, ed = orc.edit_pairs(mat, lens, 2 * p.max_diff, method="brute")
    odd = ((ei ^ ej) & 1) == 1
    query = np.where(odd, np.maximum(ei, ej), np.minimum(ei, ej))
    lo, hi = multigpu.shard_range(p.n_buckets, rank, world)
    mine = (query >= lo) & (query < hi)
    t = [torch.from_numpy(np.ascontiguousarray(x[mine])) for x in (ei, ej, ed)]
    # 3. exchange
    gi, gj, gd = (x.numpy() for x in multigpu.gather_edges(t[0], t[1], t[2], world))
    order = np.lexsort((gj, gi))
    gi, gj, gd = gi[order], gj[order], gd[order]
    assert np.array_equal(gi, ei) and np.array_equal(gj, ej) and np.array_equal(gd, ed), "gathered edges differ"
    # 4. my components, 5. element-wise max
    pairs = hostdata.pairs_from_edges(gi, gj, gd, p.max_diff)
    out = util.emul_merge(p, pairs, shard=(rank, world))
    red = {}
    for k in [n for n, _ in hostdata.MERGE_OUTPUTS] + ["call_mut"]:
        tt = torch.from_numpy(out[k].copy())
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        red[k] = tt.numpy()
    got = util.table_of(job, hostdata.rows_from_merge_outputs(red))
    assert got == util.golden_table(case), "sharded table differs from the reference"
    owned = int((out["read_root"] != -(1 << 31)).sum())
    print(f"rank {rank}: {int(mine.sum())} of {len(ei)} edges scored here, {owned} of {p.n_reads} reads clustered here")
    dist.destroy_process_group()
''')


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("case", ["rand_dense_d2", "rand_d3_combo"])
def test_two_rank_gloo(tmp_path, case):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), PCB_ROOT=ROOT, PCB_CASE=case)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    assert all("edges scored here" in o for o in outs)


UPLOAD_WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, os.environ["PCB_ROOT"])
    import numpy as np, torch, torch.distributed as dist
    from pacybara_b200 import multigpu, synth, hostdata
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    for n_reads in (1, 7, 1000):
        lib = synth.make_library(n_clones=max(1, n_reads // 8), n_reads=n_reads, seed=11)
        arrs = dict(seq=lib.seq, length=lib.length.astype(np.int32), lexrank=np.arange(lib.n_reads, dtype=np.int32)[::-1].copy(),
                    geno_ptr=lib.geno_ptr.astype(np.int32), geno_mut=lib.geno_mut.astype(np.int32), geno_q=lib.geno_q.astype(np.int32),
                    mut_rank=np.arange(len(lib.mut_codes), dtype=np.int32), up_id=None)
        t = {k: (torch.from_numpy(np.ascontiguousarray(v)) if v is not None else None) for k, v in arrs.items()}
        got = multigpu.upload_sharded(t, rank, world, "cpu")
        for k, v in arrs.items():
            if v is not None:
                assert np.array_equal(got[k].numpy(), v), (n_reads, k)
        assert "up_id" not in got
    # the compacted shards put together are the per-read arrays again
    R = 10
    full = dict(read_root=np.array([0, 0, 2, 3, 3, 3, 6, 7, 7, 9], np.int32), read_pos=np.array([0, 1, 0, 0, 2, 1, 0, 0, 1, 0], np.int32))
    shards = []
    for r in range(2):
        own = np.nonzero((full["read_root"] % 2) == r)[0].astype(np.int32)
        rep = own[full["read_root"][own] == own]
        size = np.array([(full["read_root"] == x).sum() for x in rep], np.int32)
        shards.append(dict(own_read=own, own_root=full["read_root"][own], own_pos=full["read_pos"][own], own_bucket=own * 0 + r,
                           row_rep=rep, row_size=size, row_key=np.where(size > 1, rep.astype(np.int64) << 32, -1), row_bc=rep, row_up=rep,
                           row_call_off=np.arange(len(rep), dtype=np.int64), row_call_n=np.ones(len(rep), np.int32),
                           call_mut=(rep + 100).astype(np.int32)))
    whole = hostdata.merge_outputs_from_shards(shards, R)
    assert np.array_equal(whole["read_root"], full["read_root"]) and np.array_equal(whole["read_pos"], full["read_pos"])
    rows = hostdata.rows_from_merge_outputs(whole)
    assert int(rows.row_ptr[-1]) == R
    for i in range(rows.n_rows):
        mem = rows.row_members[rows.row_ptr[i]: rows.row_ptr[i + 1]]
        assert rows.call_mut[rows.call_ptr[i]] == mem[0] + 100
    print(f"rank {rank}: upload_sharded ok")
    dist.destroy_process_group()
''')


def test_upload_sharded_two_ranks(tmp_path):
    worker = tmp_path / "upload_worker.py"
    worker.write_text(UPLOAD_WORKER)
    env = dict(os.environ, PCB_ROOT=ROOT, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(worker)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count("upload_sharded ok") == 2
