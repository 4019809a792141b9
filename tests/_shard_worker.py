"""Worker for tests/test_sharding.py: world_size-2 run of the multi-GPU host logic
on the gloo backend (no GPU here, so each rank's per-region work is done by the
oracle; what is under test is the partition + host-side gather)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch.distributed as dist   # noqa: E402

from breakmer_b200 import shard, synth   # noqa: E402
from oracle import assembler_py   # noqa: E402
from oracle.make_golden import oracle_sample_only   # noqa: E402


def work(region):
    _r, _c, _s, only = oracle_sample_only(region)
    ctg = assembler_py.init_assembly(only, region.reads, region.k, region.rc_thresh, region.read_len)
    return {"n_only": len(only), "contigs": [c["seq"] for c in ctg]}


def main():
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    regions = list(synth.config_regions("C5", n=40)) + list(synth.config_regions("C2", n=6))
    owned = shard.assign_lpt([shard.region_cost(r) for r in regions], world)
    local = {regions[i].name: work(regions[i]) for i in owned[rank]}
    merged = shard.gather_by_name(local, rank, world)
    # the cross-rank call queue: every item is taken exactly once, by whichever rank asks first; two queues of the same
    # name in a row do not share a counter
    takes = []
    for n_items in (37, 5):
        q = shard.CallQueue("t", n_items, world)
        mine = []
        while True:
            t = q.take()
            if t is None:
                break
            mine.append(t)
        assert q.take() is None
        dist.barrier()
        takes.append(mine)
    all_takes = [None] * world if rank == 0 else None
    dist.gather_object(takes, all_takes, dst=0)
    if rank == 0:
        serial = {r.name: work(r) for r in regions}
        ok = merged == dict(sorted(serial.items())) and list(merged) == sorted(merged)
        sizes = [len(o) for o in owned]
        q_ok = all(sorted(t for r in all_takes for t in r[i]) == list(range(n)) for i, n in enumerate((37, 5)))
        print(json.dumps({"ok": bool(ok), "sizes": sizes, "n": len(merged), "queue_ok": bool(q_ok)}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
