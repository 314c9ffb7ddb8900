"""Sampler contract: product Philox == oracle Philox == Random123 known-answer vectors."""
import json
import os

import numpy as np

import dogeray_b200 as drb
from oracle import restated

HERE = os.path.dirname(os.path.abspath(__file__))


def test_known_answer_vectors():
    kat = json.load(open(os.path.join(HERE, "golden", "philox_kat.json")))
    for v in kat:
        seed = v["key"][0] | (v["key"][1] << 32)
        x, y, s, blk = v["ctr"]
        for lane in range(4):
            n = (blk << 2) | lane
            if n >= 2 ** 32:      # the block index is n >> 2: only 30 bits are addressable through the API
                continue
            assert drb.philox_word(seed, x, y, s, n) == v["out"][lane]
            assert restated.philox_word(seed, x, y, s, n) == v["out"][lane]


def test_product_matches_oracle_on_random_counters():
    rng = np.random.default_rng(0)
    for _ in range(300):
        seed = int(rng.integers(0, 2 ** 63)); x, y, s = (int(v) for v in rng.integers(0, 2 ** 32, 3)); n = int(rng.integers(0, 2 ** 20))
        assert drb.philox_word(seed, x, y, s, n) == restated.philox_word(seed, x, y, s, n)


def test_uniform_range_and_exactness_of_2u_minus_1():
    u32 = np.array([drb.philox_uniform(3, 1, 2, 0, n) for n in range(2000)], np.float32)
    u = u32.astype(np.float64)
    assert (u > 0).all() and (u <= 1).all()                  # (0, 1], like curand_uniform_double
    k = u * 2 ** 25
    assert np.array_equal(k, np.round(k))                    # on the 2^-25 grid
    # the reference computes u*2-1 in double and rounds to float; in float it is the same number
    assert np.array_equal((u * 2 - 1).astype(np.float32), u32 * np.float32(2) - np.float32(1))
    assert np.array_equal((u * 2 - 1).astype(np.float32).astype(np.float64), u * 2 - 1)
    assert abs(u.mean() - 0.5) < 0.03
