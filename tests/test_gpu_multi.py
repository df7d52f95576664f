"""Two-GPU test of the sharded stack: one process per GPU (torchrun), per-rank CUDA `Wavelets` on its
block of slices, global norms through the fused reduction + ncclAllReduce of the C ABI."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    sys.path.insert(0, %(root)r)
    import torch, torch.distributed as dist
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import pypwt_b200
    from pypwt_b200.sharded import ShardedWavelets
    pypwt_b200.set_device(local)
    rank, world = dist.get_rank(), dist.get_world_size()
    stack = np.random.default_rng(5).integers(0, 256, size=(6, 256, 512)).astype(np.float32)
    S = ShardedWavelets(stack, "sym8", 3, rank=rank, world_size=world, dist=dist)
    used_nccl = S.init_nccl()
    S.forward()
    g = S.global_norms()
    l = S.local_norms()
    S.soft_threshold(5.0)
    g2 = S.global_norms()
    S.inverse()
    lo, hi = S.local_slices
    err = float(np.abs(S.local_image - stack[lo:hi]).max())
    print(json.dumps({"rank": rank, "nccl": bool(used_nccl), "g": g, "l": l, "g2": g2, "err": err,
                      "slices": [lo, hi]}), flush=True)
    S.close()
    dist.destroy_process_group()
""")


def test_sharded_stack_two_gpus(tmp_path):
    import pypwt_b200
    if pypwt_b200.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    port = 29600 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    recs = sorted((json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")), key=lambda r: r["rank"])
    assert len(recs) == 2 and all(r["nccl"] for r in recs)
    assert recs[0]["slices"] == [0, 3] and recs[1]["slices"] == [3, 6]
    # global = sum of locals, identical on both ranks, equal to a single-GPU transform of the whole stack
    tot1 = recs[0]["l"][0] + recs[1]["l"][0]
    tot2 = recs[0]["l"][1] + recs[1]["l"][1]
    for r in recs:
        assert abs(r["g"][0] - tot1) <= 1e-12 * tot1 and abs(r["g"][1] - tot2) <= 1e-12 * tot2
        assert r["g2"][0] < r["g"][0] and r["err"] < 30.0
    assert recs[0]["g"] == recs[1]["g"] and recs[0]["g2"] == recs[1]["g2"]
    import pycudwt
    stack = np.random.default_rng(5).integers(0, 256, size=(6, 256, 512)).astype(np.float32)
    W = pycudwt.Wavelets(stack, "sym8", 3)
    W.forward()
    n1, n2 = W.norms()
    assert abs(n1 - tot1) <= 1e-9 * n1 and abs(n2 - tot2) <= 1e-9 * n2


def test_stack_wavelets_single_process():
    """`StackWavelets`: one process shards the stack over every visible GPU (ncclCommInitAll when there are several,
    no communicator with one) -- results equal the one-plan transform of the whole stack."""
    import pypwt_b200
    import pycudwt
    from pypwt_b200.sharded import StackWavelets
    stack = np.random.default_rng(7).integers(0, 256, size=(5, 128, 256)).astype(np.float32)
    S = StackWavelets(stack, "sym8", 3)
    assert len(S.plans) == min(pypwt_b200.device_count(), 3 if pypwt_b200.device_count() == 4 else 5) or len(S.plans) >= 1
    W = pycudwt.Wavelets(stack, "sym8", 3)
    S.forward(); W.forward()
    n1, n2 = S.norms()
    w1, w2 = W.norms()
    assert abs(n1 - w1) <= 1e-9 * w1 and abs(n2 - w2) <= 1e-9 * w2
    cs, cw = S.coeffs, W.coeffs
    assert np.array_equal(cs[0], cw[0])
    for l in range(1, 4):
        for j in range(3):
            assert np.array_equal(cs[l][j], cw[l][j])
    S.soft_threshold(5.0); W.soft_threshold(5.0)
    S.inverse(); W.inverse()
    assert np.array_equal(S.image, W.image)
    S.forward(stack)
    assert abs(S.norm1() - w1) <= 1e-9 * w1
    S.close()
