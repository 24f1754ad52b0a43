"""bench.py's solar-angle schedule (host logic, no GPU): whole angles per rank; N in {1, 2, 4, 8} ranks partition the
8 x 8 sweep exactly; every window of the walk mixes elevations and azimuths; with the committed per-angle profile the
timed slots are dealt so that the ranks' predicted totals agree within one per cent."""
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
    per_rank = bench.schedule(world, 64 // world)
    flat = [a for seq in per_rank for a in seq]
    assert len(set(flat)) == 64 and len(flat) == 64 and all(len(seq) == 64 // world for seq in per_rank)
    assert [bench.angles_for(r, world, 64 // world) for r in range(world)] == per_rank


def test_single_rank_mixes_elevations_and_azimuths(bench):
    seq = bench.angles_for(0, 1, 25, 5)[5:]                        # the driver's run: 5 warm-up + 20 timed steps
    assert len({e for e, _ in seq}) == 8 and len({a for _, a in seq}) == 8
    from collections import Counter
    assert max(Counter(e for e, _ in seq).values()) <= 3           # 20 steps over 8 elevations: 2 or 3 of each


@pytest.mark.parametrize("world", [2, 4, 8])
def test_timed_slots_are_balanced_by_cost(bench, world):
    cost = bench.angle_costs()
    if cost is None:
        pytest.skip("profiles/r02_cast_rays_profile.json not present")
    from pyqsm_b200 import synthetic as syn
    index = {a: i for i, a in enumerate(syn.hemisphere_sweep())}
    plan = bench.schedule(world, 25, 5)
    totals = [sum(cost[index[a]] for a in seq[5:]) for seq in plan]
    assert all(len(seq) == 25 for seq in plan)
    assert (max(totals) - min(totals)) / max(totals) < 0.01
    # the multiset of timed angles is the walk's, whatever the deal
    walk = [(27 * g) % 64 for g in range(world * 25)][world * 5:]
    assert sorted(index[a] for seq in plan for a in seq[5:]) == sorted(walk)
