"""The N>1 path on real GPUs: one process per GPU (torchrun), NCCL over NVLink -- sample sharding with one reduce,
progressive rendering with a periodic reduce, tile sharding (bit-identical), and the sharded upload + all-gather that
bench.py's e2e leg uses.  Skipped on a box with fewer than two GPUs (tests/test_distributed_cpu.py covers the logic on gloo)."""
import os
import subprocess
import sys

import pytest

import dogeray_b200 as drb
from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import dogeray_b200 as drb
from dogeray_b200 import synth
from dogeray_b200.distributed import render_progressive, render_sharded, shard_tiles

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
objs, st = synth.heightfield_scene(n=40, width=160, height=90, spp=7, max_depth=5)
hs = drb.HostScene.from_objects(objs, st)

# the scene through a sharded upload: this rank's share of the object lines, one all-gather, build from device memory
rec = drb.OBJECT_DTYPE.itemsize
n = hs.num_objects
chunk = (n + world - 1) // world
lo, hi = min(n, rank * chunk), min(n, (rank + 1) * chunk)
mine = torch.zeros(chunk * rec, dtype=torch.uint8).pin_memory()
mine[: (hi - lo) * rec] = torch.from_numpy(hs.objects()[lo:hi].view(np.uint8).copy())
gathered = torch.empty(world * chunk * rec, dtype=torch.uint8, device=dev)
dist.all_gather_into_tensor(gathered, mine.to(dev, non_blocking=True))
stream = torch.cuda.current_stream().cuda_stream
scene = drb.Scene.from_device_objects(hs, gathered.data_ptr(), device=local, stream=stream)
plain = drb.Scene.from_host(hs, device=local)
wa, wb = scene.wide(), plain.wide()
assert np.array_equal(wa["child"], wb["child"]) and np.array_equal(wa["boxes"], wb["boxes"])

def render(base, count):
    acc = torch.zeros(st.height, st.width, 3, device=dev)
    scene.render_device(acc.data_ptr(), st, seed=4, sample_base=base, sample_count=count, stream=stream)
    return acc

full, sf = plain.render(st, seed=4)
acc, count = render_sharded(render, st.spp, rank, world)
torch.cuda.synchronize()
if rank == 0:
    err = float(np.abs(acc.cpu().numpy() - full).max())
    print("SHARDED_MAXERR %g" % err); assert err < 1e-4, err
snaps = []
for total, done in render_progressive(render, st.spp, 2, rank, world):
    snaps.append(done); last = total
torch.cuda.synchronize()
assert snaps == [2, 4, 6, 7], snaps
if rank == 0:
    err = float(np.abs(last.cpu().numpy() - full).max())
    print("PROGRESSIVE_MAXERR %g" % err); assert err < 1e-4, err
# chunk of one sample over two ranks: every other share is empty and must add nothing
for total, done in render_progressive(render, 3, 1, rank, world):
    last1 = total
torch.cuda.synchronize()
if rank == 0:
    three, _ = plain.render(st, seed=4, sample_count=3)
    err = float(np.abs(last1.cpu().numpy() - three).max())
    print("EMPTY_SHARE_MAXERR %g" % err); assert err < 1e-4, err
# interleaved tiles: the reduce of disjoint partial images IS the one-GPU image
tr, tc = shard_tiles(rank, world)
part = torch.zeros(st.height, st.width, 3, device=dev)
scene.render_device(part.data_ptr(), st, seed=4, stream=stream, tile_rank=tr, tile_count=tc)
dist.reduce(part, dst=0)
torch.cuda.synchronize()
if rank == 0:
    assert np.array_equal(part.cpu().numpy(), full)
    print("TILES_BIT_IDENTICAL")
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_nccl_sharding_progressive_and_gathered_upload(tmp_path):
    if drb.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", str(script), ROOT]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    for tag in ("SHARDED_MAXERR", "PROGRESSIVE_MAXERR", "EMPTY_SHARE_MAXERR", "TILES_BIT_IDENTICAL"):
        assert tag in p.stdout, tag
