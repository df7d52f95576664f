"""World-size-2 `gloo` tests (CPU) of the multi-GPU host logic: slice partitioning and the scalar
all-reduce behind the global norms.  The per-rank engine is the CPU oracle (injected by the test; the
product's default engine is the CUDA `Wavelets`)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT
from pypwt_b200.sharded import partition


def test_partition_covers_every_slice_once():
    for S in (1, 2, 7, 8, 64, 511, 512):
        for G in (1, 2, 3, 4, 8):
            b = partition(S, G)
            assert len(b) == G and b[0][0] == 0 and b[-1][1] == S
            assert all(b[i][1] == b[i + 1][0] for i in range(G - 1))
            sizes = [hi - lo for lo, hi in b]
            assert sum(sizes) == S and max(sizes) == -(-S // G)


WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    sys.path.insert(0, %(root)r)
    import torch.distributed as dist
    from pypwt_b200.sharded import ShardedWavelets
    from oracle import pdwt_oracle as O

    class Engine:                      # CPU stand-in with the Wavelets interface, slice by slice
        def __init__(self, shard, wname, levels, **kw):
            self.ws = [O.OracleWavelets(s, wname, levels, **kw) for s in shard]
        def forward(self): [w.forward() for w in self.ws]
        def inverse(self): [w.inverse() for w in self.ws]
        def soft_threshold(self, *a, **k): [w.soft_threshold(*a, **k) for w in self.ws]
        def norms(self): return sum(w.norm1() for w in self.ws), sum(w.norm2sq() for w in self.ws)
        @property
        def image(self): return np.stack([w.image for w in self.ws])

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    stack = np.random.default_rng(3).integers(0, 256, size=(%(slices)d, 32, 48)).astype(np.float32)
    S = ShardedWavelets(stack, "db2", 2, rank=rank, world_size=world, dist=dist, engine=Engine)
    assert S.init_nccl() is False      # collective even on ranks with an empty block; CPU engine: no NCCL
    S.forward()
    g1 = S.global_norms()
    S.soft_threshold(10.0)
    g2 = S.global_norms()
    S.inverse()
    err = float(np.abs(S.local_image - stack[S.local_slices[0]:S.local_slices[1]]).max()) if S.W else 0.0
    print(json.dumps({"rank": rank, "slices": S.local_slices, "g1": g1, "g2": g2, "err": err}), flush=True)
    dist.destroy_process_group()
""")


def _run(tmp_path, slices, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "slices": slices})
    import socket
    with socket.socket() as sk:            # a port that is free right now
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    recs = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(recs) == world
    recs.sort(key=lambda r: r["rank"])
    return recs


def test_empty_blocks_take_part_in_the_collectives(tmp_path):
    """S < G and S % G != 0: 1 slice over 2 ranks leaves rank 1 empty; 5 slices over 4 ranks give blocks
    of 2, 2, 1, 0.  Every rank must reach every collective and report the same global norms."""
    from oracle import pdwt_oracle as O
    for slices, world, expect in ((1, 2, [[0, 1], [1, 1]]), (5, 4, [[0, 2], [2, 4], [4, 5], [5, 5]])):
        recs = _run(tmp_path, slices, world)
        assert [r["slices"] for r in recs] == expect
        stack = np.random.default_rng(3).integers(0, 256, size=(slices, 32, 48)).astype(np.float32)
        n1 = 0.0
        for s in stack:
            W = O.OracleWavelets(s, "db2", 2)
            W.forward()
            n1 += W.norm1()
        for r in recs:
            assert abs(r["g1"][0] - n1) <= 1e-9 * n1
            assert r["g1"] == recs[0]["g1"] and r["g2"] == recs[0]["g2"]


def test_global_norms_world_size_2(tmp_path):
    recs = _run(tmp_path, 5, 2)
    assert recs[0]["slices"] == [0, 3] and recs[1]["slices"] == [3, 5]
    # both ranks see the same global value, equal to the single-process value
    from oracle import pdwt_oracle as O
    stack = np.random.default_rng(3).integers(0, 256, size=(5, 32, 48)).astype(np.float32)
    n1 = n2 = 0.0
    for s in stack:
        W = O.OracleWavelets(s, "db2", 2)
        W.forward()
        n1 += W.norm1()
        n2 += W.norm2sq()
    for r in recs:
        assert abs(r["g1"][0] - n1) <= 1e-9 * n1 and abs(r["g1"][1] - n2) <= 1e-9 * n2
        assert r["g2"][0] < r["g1"][0]          # thresholding shrinks the L1 norm, on every rank alike
        assert r["err"] < 30.0
    assert recs[0]["g2"] == recs[1]["g2"]
