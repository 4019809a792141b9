"""Multi-GPU host logic on CPU: regions are partitioned across ranks (no data-path
collective) and gathered on the host in target-name order."""
import json
import os
import subprocess
import sys

from conftest import ROOT
from breakmer_b200 import shard


def test_lpt_assignment_is_a_balanced_partition():
    costs = [50, 1, 1, 30, 20, 20, 9, 9, 9, 1]
    owned = shard.assign_lpt(costs, 3)
    assert sorted(i for o in owned for i in o) == list(range(len(costs)))
    loads = [sum(costs[i] for i in o) for o in owned]
    assert max(loads) == 50 and min(loads) >= 49
    assert shard.assign_lpt(costs, 3) == owned          # deterministic
    assert shard.assign_lpt([], 2) == [[], []]
    assert shard.assign_lpt([3, 2], 1) == [[0, 1]]


def test_gather_single_rank_orders_by_name():
    assert list(shard.gather_by_name({"b": 1, "a": 2}, 0, 1)) == ["a", "b"]


def test_two_ranks_gloo():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "_shard_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"] and res["n"] == 46 and sum(res["sizes"]) == 46 and min(res["sizes"]) > 0
    assert res["queue_ok"]


def test_call_queue_single_rank_is_a_local_counter():
    q = shard.CallQueue("x", 3)
    assert [q.take(), q.take(), q.take(), q.take(), q.take()] == [0, 1, 2, None, None]
    assert shard.CallQueue("x", 0).take() is None


def test_chunks_cover_the_shard_and_respect_the_call_limit():
    assert shard.chunk_indices([], 5) == []
    idx = list(range(0, 46, 2))
    chunks = shard.chunk_indices(idx, 5)
    assert [i for c in chunks for i in c] == idx and max(len(c) for c in chunks) <= 5
    assert max(len(c) for c in chunks) - min(len(c) for c in chunks) <= 4
    assert shard.chunk_indices(idx, 1000) == [idx]


def test_static_cost_grows_with_reads_and_bytes():
    from breakmer_b200 import synth
    small = synth.config_region("C5", 1)
    big = synth.config_region("C2", 1)
    assert shard.region_cost(big) > shard.region_cost(small) > 0
