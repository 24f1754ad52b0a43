"""bench.py's solar-angle schedule (host logic, no GPU): N in {1, 2, 4, 8} ranks partition the 8 x 8 sweep exactly,
the ranks of one step share the elevation and are 360/N degrees apart, and N = 1 walks every angle."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_angles_partition_the_sweep(bench, world):
    per_rank = [bench.angles_for(r, world, 64 // world) for r in range(world)]
    flat = [a for seq in per_rank for a in seq]
    assert len(set(flat)) == 64 and len(flat) == 64
    for s in range(64 // world):
        step = [per_rank[r][s] for r in range(world)]
        assert len({e for e, _ in step}) == 1                      # same elevation -> comparable work
        az = sorted(a for _, a in step)
        gaps = {(az[(k + 1) % world] - az[k]) % 360 for k in range(world)} if world > 1 else {0}
        assert gaps == ({360 // world} if world > 1 else {0})


def test_single_rank_mixes_elevations_and_azimuths(bench):
    seq = bench.angles_for(0, 1, 23)[3:]                          # the default run: 3 warm-up + 20 timed steps
    assert len({e for e, _ in seq}) == 8 and len({a for _, a in seq}) == 8
