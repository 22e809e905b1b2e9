"""World-size-2 test of the multi-GPU path's host logic on CPU (gloo): the sensor frame lives on rank 0 and is broadcast every
step, every rank integrates only the voxel blocks it owns (hp.m_dummy = {rank, world}: the spatial shard of SURVEY.md section 8e),
no collective follows the kernels.  The CPU oracle stands in for the device kernels (same owner function as tsdf.cu); what is
checked is the distributed contract: shards are disjoint, their union is bit-identical to the unsharded run, and the load is
balanced.  bench.py --gpus N runs exactly this sequence with NCCL and the CUDA library."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bundlefusion_b200 import synth
from bundlefusion_b200.scene_rep import camera_params, default_hash_params
from oracle import oracle as orc

W, H, N_FRAMES = 160, 120, 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_stream(hp, frames, cam):
    o = orc.OracleSceneRepHashSDF(hp)
    for d, c, T in frames:
        o.integrate(T, d, c, cam)
    d, c, T = frames[1]                                       # one re-integration at a nudged pose + GC
    T2 = T.copy(); T2[:3, 3] += np.array([0.02, -0.01, 0.015], np.float32)
    o.deIntegrate(T, d, c, cam); o.integrate(T2, d, c, cam); o.garbageCollect()
    return orc.canonical_blocks(o.download())


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cam = camera_params(W, H)
    hp = default_hash_params(num_buckets=20011, num_sdf_blocks=30000)
    hp.m_dummy = (world << 32) | rank
    frames = []
    for i in range(N_FRAMES):
        if rank == 0:
            d, c, T = synth.make_frame(35 * i, W, H)
            td, tc, tT = torch.from_numpy(d.copy()), torch.from_numpy(c.copy()), torch.from_numpy(T.copy())
        else:
            td, tc, tT = torch.empty(H, W), torch.empty(H, W, 4, dtype=torch.uint8), torch.empty(4, 4)
        dist.broadcast(td, 0); dist.broadcast(tc, 0); dist.broadcast(tT, 0)          # the one real exchange of the path
        frames.append((td.numpy(), tc.numpy(), tT.numpy()))
    blocks, vox = _run_stream(hp, frames, cam)
    gathered = [None] * world
    dist.all_gather_object(gathered, (blocks, vox))
    if rank == 0:
        hp1 = default_hash_params(num_buckets=20011, num_sdf_blocks=30000)
        ref_blocks, ref_vox = _run_stream(hp1, frames, cam)
        np.savez(out_path, ref_blocks=ref_blocks, ref_vox=ref_vox, **{f"b{r}": g[0] for r, g in enumerate(gathered)},
                 **{f"v{r}": g[1] for r, g in enumerate(gathered)})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_stream_world_size_2(tmp_path):
    world, port, out = 2, _free_port(), str(tmp_path / "shards.npz")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    z = np.load(out)
    sets = [set(map(tuple, z[f"b{r}"])) for r in range(world)]
    assert not (sets[0] & sets[1])                                            # shards are disjoint
    ref = z["ref_blocks"]
    assert sets[0] | sets[1] == set(map(tuple, ref))                          # and cover exactly the unsharded block set
    assert min(len(s) for s in sets) > 0.35 * len(ref)                        # balanced owner function
    merged_b = np.concatenate([z["b0"], z["b1"]]); merged_v = np.concatenate([z["v0"], z["v1"]])
    order = np.lexsort((merged_b[:, 2], merged_b[:, 1], merged_b[:, 0]))
    ref_order = np.lexsort((ref[:, 2], ref[:, 1], ref[:, 0]))
    np.testing.assert_array_equal(merged_b[order], ref[ref_order])
    np.testing.assert_array_equal(merged_v[order], z["ref_vox"][ref_order])   # every voxel word bit-identical
