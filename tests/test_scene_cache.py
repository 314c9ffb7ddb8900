"""Binary scene cache: a hit returns exactly what the text parse returns; any change of the
scene text, of the texture list or of the cache file itself falls back to the parse."""
import os
import shutil

import numpy as np
import pytest

import dogeray_b200 as drb
from dogeray_b200 import synth
from conftest import HAVE_REF, SAMPLES, all_sample_scenes, sample


def same_scene(a, b):
    assert a.num_objects == b.num_objects
    assert a.num_skipped == b.num_skipped
    assert bytes(a.settings) == bytes(b.settings)
    assert a.texture_paths == b.texture_paths
    assert np.array_equal(a.objects().view(np.uint8), b.objects().view(np.uint8))


def write_scene(tmp_path, n=64, seed=3):
    objs, st = synth.heightfield_scene(n, seed=seed)
    path = str(tmp_path / "scene.rts")
    drb.write_rts(path, st, objs)
    return path


def test_miss_then_hit_is_identical(tmp_path):
    path = write_scene(tmp_path)
    plain = drb.HostScene.load(path, str(tmp_path))
    first = drb.HostScene.load(path, str(tmp_path), cache=True)
    assert not first.cache_hit and os.path.exists(path + ".drbcache")
    second = drb.HostScene.load(path, str(tmp_path), cache=True)
    assert second.cache_hit
    same_scene(first, plain)
    same_scene(second, plain)
    assert not [n for n in os.listdir(tmp_path) if ".tmp." in n]          # written through rename


def test_explicit_cache_path_and_edited_scene(tmp_path):
    path = write_scene(tmp_path)
    cache = str(tmp_path / "elsewhere.bin")
    assert not drb.HostScene.load(path, str(tmp_path), cache=cache).cache_hit
    assert drb.HostScene.load(path, str(tmp_path), cache=cache).cache_hit
    assert not os.path.exists(path + ".drbcache")
    text = open(path, "rb").read()
    i = text.index(b"\n", text.index(b"\n*") + 1) + 1                    # first object line: a digit of its first coordinate
    while not text[i:i + 1].isdigit():
        i += 1
    edited = text[:i] + (b"7" if text[i:i + 1] != b"7" else b"8") + text[i + 1:]
    assert len(edited) == len(text)
    open(path, "wb").write(edited)
    again = drb.HostScene.load(path, str(tmp_path), cache=cache)
    assert not again.cache_hit                                            # same length, same mtime granularity: the hash decides
    same_scene(again, drb.HostScene.load(path, str(tmp_path)))
    assert drb.HostScene.load(path, str(tmp_path), cache=cache).cache_hit


def test_texture_list_is_part_of_the_key(tmp_path):
    tex = synth.write_test_textures(str(tmp_path))
    objs, st, tp = synth.materials_scene(tex, width=32, height=32, spp=1, max_depth=2, nu=8, nv=4)
    path = str(tmp_path / "mats.rts")
    drb.write_rts(path, st, objs, tex_names=[os.path.basename(t) for t in tp], backtex_name=os.path.basename(tp[0]))
    a = drb.HostScene.load(path, str(tmp_path), cache=True)
    assert not a.cache_hit and (a.objects()["texnum"] >= 0).any()
    assert drb.HostScene.load(path, str(tmp_path), cache=True).cache_hit
    shutil.copy(str(tmp_path / os.path.basename(a.texture_paths[0])), str(tmp_path / "0000_first.ppm"))   # shifts every index
    b = drb.HostScene.load(path, str(tmp_path), cache=True)
    assert not b.cache_hit
    same_scene(b, drb.HostScene.load(path, str(tmp_path)))
    assert not np.array_equal(a.objects()["texnum"], b.objects()["texnum"])


@pytest.mark.parametrize("damage", ["truncate", "magic", "garbage", "empty"])
def test_damaged_cache_falls_back_to_the_parse(tmp_path, damage):
    path = write_scene(tmp_path, n=16)
    plain = drb.HostScene.load(path, str(tmp_path))
    drb.HostScene.load(path, str(tmp_path), cache=True)
    cache = path + ".drbcache"
    data = open(cache, "rb").read()
    bad = {"truncate": data[:-100], "magic": b"X" + data[1:], "garbage": os.urandom(300), "empty": b""}[damage]
    open(cache, "wb").write(bad)
    hs = drb.HostScene.load(path, str(tmp_path), cache=True)
    assert not hs.cache_hit
    same_scene(hs, plain)
    assert drb.HostScene.load(path, str(tmp_path), cache=True).cache_hit  # and the cache was repaired


def test_unwritable_cache_is_not_an_error(tmp_path):
    path = write_scene(tmp_path, n=8)
    hs = drb.HostScene.load(path, str(tmp_path), cache=str(tmp_path / "no" / "such" / "dir" / "c.bin"))
    assert not hs.cache_hit
    same_scene(hs, drb.HostScene.load(path, str(tmp_path)))


def test_missing_scene_is_an_io_error(tmp_path):
    with pytest.raises(drb.DogerayError) as e:
        drb.HostScene.load(str(tmp_path / "absent.rts"), str(tmp_path), cache=True)
    assert e.value.status < 0 and "absent.rts" in str(e.value)


def test_warnings_and_skips_survive_the_cache(tmp_path):
    path = str(tmp_path / "odd.rts")
    open(path, "w").write("/ comment\n2,0,0,0,1,1,1\n\n0,0,0,0,1,1,1,0,0,1\n")
    plain = drb.HostScene.load(path, str(tmp_path))
    drb.HostScene.load(path, str(tmp_path), cache=True)
    hit = drb.HostScene.load(path, str(tmp_path), cache=True)
    assert hit.cache_hit and hit.num_skipped == plain.num_skipped > 0
    assert b"line" in drb._lib.drb_last_error()                            # the first warning is reported again
    same_scene(hit, plain)


def test_hash_is_stable_and_sensitive():
    # chunk boundaries (1 MiB) and tails of every length modulo 8
    rng = np.random.default_rng(1)
    data = rng.integers(0, 256, (1 << 20) * 2 + 13, dtype=np.uint8).tobytes()
    h = drb.hash_bytes(data)
    assert h == drb.hash_bytes(bytes(data)) and 0 <= h < 1 << 64
    seen = {h}
    for cut in (0, 1, 7, 8, 9, (1 << 20) - 1, 1 << 20, (1 << 20) + 1, len(data) - 1):
        seen.add(drb.hash_bytes(data[:cut]))
        flipped = bytearray(data); flipped[min(cut, len(data) - 1)] ^= 1
        seen.add(drb.hash_bytes(bytes(flipped)))
    assert len(seen) == 1 + 9 + 9 - 0                                     # no collisions among these
    swapped = data[1 << 20:2 << 20] + data[:1 << 20] + data[2 << 20:]     # chunk order matters
    assert drb.hash_bytes(swapped) != h
    assert drb.hash_bytes(b"") == drb.hash_bytes(b"")


@pytest.mark.skipif(not os.path.isdir(SAMPLES), reason="sample scenes live under /root/reference")
@pytest.mark.parametrize("name", all_sample_scenes() if os.path.isdir(SAMPLES) else [])
def test_every_shipped_scene_round_trips_through_the_cache(tmp_path, name):
    cache = str(tmp_path / "c.bin")
    plain = drb.HostScene.load(sample(name), SAMPLES)
    assert not drb.HostScene.load(sample(name), SAMPLES, cache=cache).cache_hit
    hit = drb.HostScene.load(sample(name), SAMPLES, cache=cache)
    assert hit.cache_hit
    same_scene(hit, plain)
