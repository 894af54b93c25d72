"""Host-side logic of bench.py that needs no GPU: how a rank's batched replay is spread over host threads (configs[4]) and
the rank's slice of the host cores."""
import os

import bench


def test_batch_schedule():
    # enough cores: one sequence per thread
    assert bench.batch_schedule(8, 64, 16) == (8, 1, 8)
    assert bench.batch_schedule(8, 8, 8) == (8, 1, 8)
    # short of cores: fewer threads, each with several sequences in flight (1.5 x the wanted number, never more than there are)
    assert bench.batch_schedule(8, 64, 4) == (4, 3, 12)
    assert bench.batch_schedule(8, 8, 4) == (4, 3, 8)
    assert bench.batch_schedule(8, 32, 6) == (6, 2, 12)
    assert bench.batch_schedule(8, 64, 1) == (1, 12, 12)
    # fewer sequences than wanted, unknown core count, forced group
    assert bench.batch_schedule(8, 3, 16) == (3, 1, 3)
    assert bench.batch_schedule(8, 64, 0) == (8, 1, 8)
    assert bench.batch_schedule(2, 8, 16, forced_group=3) == (2, 3, 6)
    for want in range(1, 12):
        for cpus in range(0, 12):
            for n in (1, 5, 64):
                t, g, f = bench.batch_schedule(want, n, cpus)
                assert t >= 1 and g >= 1 and 1 <= f <= n and (cpus == 0 or t <= cpus)


def test_rank_cpus_is_computed_once(monkeypatch):
    bench._RANK_CPUS = None
    cores = set(range(32))
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: cores, raising=False)
    first = bench.rank_cpus(2, 8)
    assert first == [8, 9, 10, 11]
    # main() then restricts the process to that slice; asking again must not slice the slice
    cores = set(first)
    assert bench.rank_cpus(2, 8) == first
    bench._RANK_CPUS = None
    assert bench.rank_cpus(0, 1) == sorted(cores)
    bench._RANK_CPUS = None
