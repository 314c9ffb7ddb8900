"""The N>1 path on CPU: world_size-2 gloo, samples sharded, one reduce -- with the C oracle standing in
for the GPU renderer (the sharding / reduce logic is the code under test, not the oracle)."""
import os
import sys

import numpy as np
import pytest

from dogeray_b200.distributed import shard_samples
from conftest import ROOT


def test_shard_samples_partitions_exactly():
    for spp in (0, 1, 7, 8, 256, 1023):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                b, c = shard_samples(spp, r, world, base=5)
                got += list(range(b, b + c))
            assert got == list(range(5, 5 + spp))
            counts = [shard_samples(spp, r, world)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        shard_samples(4, 2, 2)


def test_tile_owner_masks_partition_the_image():
    from dogeray_b200.distributed import shard_tiles, tile_owner_mask
    for (w, h) in ((64, 40), (70, 45), (7, 3), (1920, 1080)):
        for world in (1, 2, 3, 8):
            masks = [tile_owner_mask(w, h, r, world) for r in range(world)]
            assert np.sum(masks, axis=0).min() == 1 and np.sum(masks, axis=0).max() == 1      # every pixel exactly once
            if w >= 16 and world > 1:
                assert not masks[0][0, 8] and masks[1][0, 8] and masks[0][0, :8].all()         # neighbouring tiles alternate
            counts = [int(m.sum()) for m in masks]
            if (w, h) == (1920, 1080):
                assert max(counts) - min(counts) <= 32 * 2                                     # balanced to a couple of tiles
    assert shard_tiles(1, 4) == (1, 4)
    with pytest.raises(ValueError):
        shard_tiles(4, 4)


WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import dogeray_b200 as drb
from dogeray_b200 import synth
from dogeray_b200.distributed import init_from_env, render_progressive, render_sharded
from oracle import restated

rank, world = init_from_env("gloo")
objs, st = synth.heightfield_scene(n=10, width=24, height=16, spp=5, max_depth=4)
path = os.path.join(sys.argv[2], "s%d.rts" % rank)
drb.write_rts(path, st, objs)
r = restated.Restated(path)

def render(base, count):
    s = st.replace(spp=max(count, 1))
    r.apply(s); r.set_seed(3)
    f, _, _ = r.frame(1, base, threads=1)                      # 255 * mean over `count` samples, (W,H,3)
    return torch.from_numpy((f.astype(np.float64) * count / 255.0).astype(np.float32))          # count == 0 (an empty share) -> zeros

acc, count = render_sharded(render, st.spp, rank, world)
if rank == 0:
    r.apply(st); r.set_seed(3)
    full, _, _ = r.frame(1, 0, threads=1)
    got = acc.numpy() / st.spp * 255.0
    err = float(np.abs(got - full).max())
    print("MAXERR %g" % err)
    assert err < 2e-3, err
# progressive: 5 samples in chunks of 2 -> three snapshots; the last one is the whole frame again
snaps = []
for total, done in render_progressive(render, st.spp, 2, rank, world):
    snaps.append(done)
    last = total
assert snaps == [2, 4, 5], snaps
if rank == 0:
    err = float(np.abs(last.numpy() / st.spp * 255.0 - full).max())
    print("PROGRESSIVE_MAXERR %g" % err)
    assert err < 2e-3, err
# fewer samples than ranks: rank 1's share of every 1-sample chunk is empty and must contribute nothing
calls = []
def render_logged(base, count):
    calls.append((base, count))
    return render(base, count)
for total, done in render_progressive(render_logged, 3, 1, rank, world):
    last1 = total
assert calls == ([(0, 1), (1, 1), (2, 1)] if rank == 0 else [(1, 0), (2, 0), (3, 0)]), calls
if rank == 0:
    r.apply(st.replace(spp=3)); r.set_seed(3)
    three, _, _ = r.frame(1, 0, threads=1)
    err = float(np.abs(last1.numpy() / 3 * 255.0 - three).max())
    print("EMPTY_SHARE_MAXERR %g" % err)
    assert err < 2e-3, err
# tile sharding: each rank contributes only the pixels of its tiles; the reduced image IS the full frame, bit for bit
from dogeray_b200.distributed import tile_owner_mask
r.apply(st); r.set_seed(3)
whole, _, _ = r.frame(1, 0, threads=1)                          # (W, H, 3)
mine = tile_owner_mask(st.width, st.height, rank, world).T      # (W, H)
part = torch.from_numpy(np.where(mine[:, :, None], whole, np.float32(0)))
dist.reduce(part, dst=0)
if rank == 0:
    assert np.array_equal(part.numpy(), whole)
    print("TILES_BIT_IDENTICAL")
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_gloo_sample_sharding(tmp_path):
    import subprocess
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script), ROOT, str(tmp_path)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "MAXERR" in p.stdout and "PROGRESSIVE_MAXERR" in p.stdout and "TILES_BIT_IDENTICAL" in p.stdout and "EMPTY_SHARE_MAXERR" in p.stdout
