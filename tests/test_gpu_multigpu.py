"""N > 1 on real GPUs (skipped on a 1-GPU box): two ranks shard the pair triangle, gather over NCCL and
reproduce the single-GPU matrix bit for bit."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

_WORKER = r'''
import os, sys, ctypes as C
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from prographmsa_b200 import alphabet, model, synth, sharding, distance, _lib
from prographmsa_b200.options import Options
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
opts = Options(mldist_flag=True, nwdist_flag=True)
mf = model.getDefault(opts)
seqs, _ = synth.simulate_family(60, 200, mf, seed=3, indel_rate=0.03)
names = sorted(seqs); N = len(names)
lst = [seqs[k] for k in names]
fac = distance.DistanceFactoryAlign(mf, opts, _lib.Context(local))
pb, pe = sharding.pair_shard(N, rank, world); per = sharding.shard_size(N, world)
d, v = fac.pair_range(lst, pb, pe)
loc = torch.zeros(2, per, dtype=torch.float64, device="cuda")
loc[0, :pe - pb] = torch.from_numpy(d).cuda(); loc[1, :pe - pb] = torch.from_numpy(v).cuda()
D, V = sharding.scatter_to_matrices(sharding.all_gather_shards(dist, loc, world).cpu().numpy(), N)
if rank == 0:
    dm = fac.computePwDistances(seqs, names)
    assert np.array_equal(D, dm.distances) and np.array_equal(V, dm.variances)
    print("multigpu ok")
dist.barrier(); dist.destroy_process_group()
'''


def test_two_gpus_reproduce_single_gpu_matrix(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29621", str(script), ROOT],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "multigpu ok" in r.stdout


def test_two_contexts_in_one_process():
    """two pgm_ctx in one process (on two devices when the box has them, else both on device 0): the kernels'
    opted-in dynamic shared memory sizes are per device, so the second context must not inherit a skipped opt-in.  Both
    contexts run the n = 20 ML kernel (> 48 KB of dynamic shared memory), the NW profile kernel and a graph DP job."""
    import numpy as np
    import torch
    from prographmsa_b200 import model, synth, distance, graphalign, _lib
    from prographmsa_b200.graph import sequence_graph
    from prographmsa_b200.options import Options
    opts = Options(mldist_flag=True, nwdist_flag=True)
    mf = model.getDefault(opts)
    seqs, _ = synth.simulate_family(24, 150, mf, seed=5, indel_rate=0.03)
    names = sorted(seqs)
    second = 1 if torch.cuda.device_count() >= 2 else 0
    ctxs = [_lib.Context(0), _lib.Context(second)]
    try:
        mats = [distance.DistanceFactoryAlign(mf, opts, c).computePwDistances(seqs, names) for c in ctxs]
        assert np.array_equal(mats[0].distances, mats[1].distances) and np.array_equal(mats[0].variances, mats[1].variances)
        g1, g2 = sequence_graph(seqs[names[0]], 20), sequence_graph(seqs[names[1]], 20)
        m = mf.getModel(0.4, 0.4)
        res = [graphalign.alignGraphs(g1, g2, m, opts, ctx=c) for c in ctxs]
        assert res[0]["score"] == res[1]["score"] and np.array_equal(res[0]["mapping1"], res[1]["mapping1"])
    finally:
        for c in ctxs:
            c.close()
