"""Region sharding across GPUs.

Target regions are independent: the reference's region loop
(sv_processor.py:185-201) touches no shared mutable state.  Work is therefore
partitioned BY REGION, one process per GPU, with no data-path collective; the
only cross-rank step is a host-side gather of per-region results in target-name
order (sv_processor.py:175-176 iterates sorted names).
"""


def region_cost(region):
    """Static cost estimate used for balancing (before anything ran on the device):
    overlap DP work grows with reads x read length x contig length."""
    n = len(region.reads)
    return n * max(1, n) + 1


def assign_lpt(costs, world_size):
    """Longest-processing-time-first assignment.  Returns, per rank, the sorted list
    of region indices it owns.  Deterministic: ties broken by index."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world_size
    owned = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda w: (load[w], w))
        owned[r].append(i)
        load[r] += costs[i]
    return [sorted(o) for o in owned]


def gather_by_name(local_results, rank, world_size, group=None):
    """Host-side gather of {region name: result} dicts onto rank 0, merged and
    ordered by target name.  Uses torch.distributed only as plumbing."""
    if world_size == 1:
        return dict(sorted(local_results.items()))
    import torch.distributed as dist
    gathered = [None] * world_size if rank == 0 else None
    dist.gather_object(local_results, gathered, dst=0, group=group)
    if rank != 0:
        return None
    merged = {}
    for part in gathered:
        overlap = set(merged) & set(part)
        if overlap:
            raise ValueError("region assigned to two ranks: %s" % sorted(overlap)[:3])
        merged.update(part)
    return dict(sorted(merged.items()))
