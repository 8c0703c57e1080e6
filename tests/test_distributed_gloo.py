"""world_size-2 gloo test of the population sharding + residual gather (CPU; the evaluator
stand-in is the oracle, the GPU evaluator is exercised by bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from helpers import ROOT


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyxrd_oracle as orc
    from pyxrd_b200 import synthetic as syn
    from pyxrd_b200.distributed import evaluate_population_sharded

    def evaluate(mixes):
        rows = []
        for m in mixes:
            orc.calculate_mixture(m)
            rows.append(np.array(m.residuals, dtype=float))
        return np.array(rows)

    mixes = syn.CONFIGS["c1"].population(5, 5)          # 5 candidates over 2 ranks: 3 + 2
    rng = np.random.default_rng(0)
    obs = rng.uniform(10, 20, (2100, 1))
    for m in mixes:
        m.specimens[0].observed_intensity = obs.copy()
    table = evaluate_population_sharded(mixes, evaluate)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), table)
    # the template path: solution rows instead of DataObject graphs
    from pyxrd_b200.distributed import evaluate_solutions_sharded
    x = np.arange(7 * 3, dtype=float).reshape(7, 3)                  # 7 solutions over 2 ranks: 4 + 3
    one = evaluate_solutions_sharded(x, lambda block: block.sum(axis=1) * 2.0)
    two = evaluate_solutions_sharded(x, lambda block: np.stack([block[:, 0], block[:, 2] - block[:, 1]], axis=1))
    # an evaluator that hands back a torch tensor (on the GPU path: the zero-copy view of the device buffer) is gathered
    # as it is; `columns` saves the broadcast of the column count
    import torch
    three = evaluate_solutions_sharded(x, lambda block: torch.from_numpy(np.stack([block[:, 1], block[:, 0] * 0.5], axis=1)),
                                       columns=2)
    np.save(os.path.join(out_dir, "sol%d.npy" % rank), np.column_stack([one, two, three]))
    dist.destroy_process_group()


def test_shard_bounds():
    from pyxrd_b200.distributed import shard_bounds
    assert shard_bounds(5, 2) == [(0, 3), (3, 5)]
    assert shard_bounds(256, 8) == [(32 * r, 32 * r + 32) for r in range(8)]
    assert shard_bounds(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    assert shard_bounds(0, 2) == [(0, 0), (0, 0)]


def test_two_rank_gather_equals_single_rank(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "rank0.npy")
    b = np.load(tmp_path / "rank1.npy")
    assert a.shape == (5, 2) and np.array_equal(a, b)
    from oracle import pyxrd_oracle as orc
    from pyxrd_b200 import synthetic as syn
    mixes = syn.CONFIGS["c1"].population(5, 5)
    rng = np.random.default_rng(0)
    obs = rng.uniform(10, 20, (2100, 1))
    want = []
    for m in mixes:
        m.specimens[0].observed_intensity = obs.copy()
        orc.calculate_mixture(m)
        want.append(m.residuals)
    assert np.array_equal(a, np.array(want, dtype=float))
    x = np.arange(7 * 3, dtype=float).reshape(7, 3)
    want = np.column_stack([x.sum(axis=1) * 2.0, x[:, 0], x[:, 2] - x[:, 1], x[:, 1], x[:, 0] * 0.5])
    assert np.array_equal(np.load(tmp_path / "sol0.npy"), want) and np.array_equal(np.load(tmp_path / "sol1.npy"), want)
