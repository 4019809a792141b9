"""Full-size parity of the five BASELINE.json configurations (north_star: "bit-exact contigs and
sample-only k-mer sets versus the reference on all five configs").

Every region of each configuration goes through bk_compare_kmers_batch on the device and is compared,
region by region, with the oracle (which is pinned to the reference's own output, tests/golden/): the
sample-only {mer: count} set and the complete contig records (sequence, both count vectors, read ids,
k-mer 5-tuples, kmer_locs).  The oracle side runs on all host cores with its C restatement of olc.nw.

    C1  1 region, 20 kb, 1.5 kb deletion, k=15
    C2  500-target panel, k=15
    C3  500 targets, tumour/normal with normal-k-mer subtraction (K4)
    C4  100 amplicons at 2000x, k=21
    C5  20,000 exome-scale regions, in calls of 2,500 (n_regions <= 65535 per call)
"""
import multiprocessing as mp
import os

import pytest

from breakmer_b200 import synth
from oracle import assembler_py
from oracle.make_golden import digest, oracle_sample_only

pytestmark = pytest.mark.gpu

FULL = {"C1": 1, "C2": 500, "C3": 500, "C4": 100, "C5": 20000}
CALL = 2500


def _oracle_one(args):
    cfg, i = args
    r = synth.config_region(cfg, i)
    _a, _b, _c, only = oracle_sample_only(r)
    ctg = assembler_py.init_assembly(only, r.reads, r.k, r.rc_thresh, r.read_len)
    return i, digest(sorted(only.items())), digest(ctg), len(only), len(ctg)


@pytest.fixture(scope="module")
def pool():
    # forked before this process touches CUDA; the workers never do
    p = mp.get_context("fork").Pool(os.cpu_count() or 1)
    yield p
    p.close()
    p.join()


@pytest.fixture(scope="module")
def handle(pool):
    from breakmer_b200 import _lib
    h = _lib.Handle(0)
    yield h
    h.close()


def _verify(cfg, pool, handle):
    from breakmer_b200 import batch
    n = FULL[cfg]
    pending = pool.map_async(_oracle_one, [(cfg, i) for i in range(n)], chunksize=max(1, min(64, n // 64)))
    got = {}
    status_bad = []
    for a in range(0, n, CALL):
        regions = [synth.config_region(cfg, i) for i in range(a, min(n, a + CALL))]
        out = batch.run(handle, batch.PackedBatch(regions))
        assert out.n_regions == len(regions)
        for j in range(len(regions)):
            got[a + j] = (digest(sorted(out.sample_only(j).items())), digest(out.contig_records(j)))
            if out.region_status[j] != 0:
                status_bad.append(a + j)
    exp = {r[0]: r[1:] for r in pending.get()}
    assert not status_bad, "%s: region_status != 0 for regions %s" % (cfg, status_bad[:10])
    bad_only = [i for i in range(n) if got[i][0] != exp[i][0]]
    bad_ctg = [i for i in range(n) if got[i][1] != exp[i][1]]
    n_only = sum(e[2] for e in exp.values())
    n_ctg = sum(e[3] for e in exp.values())
    print("%s: %d regions, %d sample-only k-mers, %d contigs, %d / %d mismatching (k-mer sets / contigs)" %
          (cfg, n, n_only, n_ctg, len(bad_only), len(bad_ctg)))
    assert not bad_only, "%s: sample-only k-mer sets differ from the oracle in regions %s" % (cfg, bad_only[:10])
    assert not bad_ctg, "%s: contigs differ from the oracle in regions %s" % (cfg, bad_ctg[:10])
    assert n_only > 0
    return n_only, n_ctg


def test_c1_single_region_full_size(pool, handle):
    _only, n_ctg = _verify("C1", pool, handle)
    assert n_ctg >= 1


def test_c2_panel_500_targets_full_size(pool, handle):
    _only, n_ctg = _verify("C2", pool, handle)
    assert n_ctg > 500


def test_c3_tumour_normal_500_targets_full_size(pool, handle):
    _only, n_ctg = _verify("C3", pool, handle)
    assert n_ctg > 500


def test_c4_amplicons_2000x_k21_full_size(pool, handle):
    _only, n_ctg = _verify("C4", pool, handle)
    assert n_ctg > 100


def test_c5_exome_20000_regions_full_size(pool, handle):
    _only, n_ctg = _verify("C5", pool, handle)
    assert n_ctg > 0
