"""GPU parity of the whole hot path (bk_compare_kmers_batch) against the oracle and
against the reference's own outputs (golden)."""
import pytest

from conftest import golden
from breakmer_b200 import synth
from oracle import assembler_py
from oracle.make_golden import digest, oracle_sample_only, region_scenarios

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    from breakmer_b200 import _lib
    h = _lib.Handle(0)
    yield h
    h.close()


def oracle_region(region):
    _r, _c, _s, only = oracle_sample_only(region)
    ctg = assembler_py.init_assembly(only, region.reads, region.k, region.rc_thresh, region.read_len)
    return only, ctg


def check_batch(handle, regions, with_mers=False):
    from breakmer_b200 import batch
    pk = batch.PackedBatch(regions)
    exp = [oracle_region(r) for r in regions]
    if with_mers:
        pk.set_mers([e[0] for e in exp])
    out = batch.run(handle, pk)
    assert out.n_regions == len(regions)
    assert all(s == 0 for s in out.region_status)
    for i, r in enumerate(regions):
        only, ctg = exp[i]
        assert out.sample_only(i) == only, r.name
        got = out.contig_records(i)
        assert got == ctg, r.name
    return out


def test_golden_regions_by_k(handle):
    cases = golden("assembly_golden.json")["cases"]
    by_k = {}
    for c in cases:
        kw = dict(c["kwargs"]); kw["event"] = tuple(kw["event"])
        by_k.setdefault(kw["k"], []).append((c, synth.make_region(c["name"], **kw)))
    for k, items in by_k.items():
        from breakmer_b200 import batch
        regions = [it[1] for it in items]
        out = batch.run(handle, batch.PackedBatch(regions))
        for i, (c, region) in enumerate(items):
            got = out.contig_records(i)
            assert len(out.sample_only(i)) == c["n_sample_only"]
            assert digest(got) == c["contigs_sha256"], c["name"]       # the reference's own output
            if "contigs" in c:
                assert got == c["contigs"]


def test_init_assembly_shape_with_given_mers(handle):
    regions = [synth.make_region(n, **kw) for n, kw in region_scenarios()[:20] if kw["k"] == 15]
    check_batch(handle, regions, with_mers=True)


def test_config_slices(handle):
    check_batch(handle, [synth.config_region("C1", 0)])
    check_batch(handle, list(synth.config_regions("C2", n=24)))
    check_batch(handle, list(synth.config_regions("C3", n=12)))       # normal subtraction
    check_batch(handle, list(synth.config_regions("C4", n=4)))
    check_batch(handle, list(synth.config_regions("C5", n=200, start=0)))


def test_empty_and_ragged_regions(handle):
    r0 = synth.make_region("e0", seed=5, L=800, cov=0, k=15, e=0.0, event=("none",))
    assert r0.reads == []
    r1 = synth.make_region("e1", seed=6, L=900, cov=200, k=15, e=0.01, event=("del", 100, None))
    r2 = synth.Region(name="e2", k=15, ref_fwd="ACGT" * 100, reads=[("@a:1:1:1:1/1_0", "ACGTN", "IIIII", False)],
                      sc_records=[("a", "AC")])
    out = check_batch(handle, [r0, r1, r2, r0])
    assert out.ctg_reg_off[1] == 0 and out.sample_only(0) == {}
    from breakmer_b200 import batch
    empty = batch.run(handle, batch.PackedBatch([]))
    assert empty.n_regions == 0 and empty.n_contigs == 0


def test_long_reads_use_the_blocked_dp(handle):
    regions = [synth.make_region("lr%d" % i, seed=700 + i, L=3000, cov=120, k=21, e=0.01,
                                 event=("del", 200, None), rl=300, rl_jitter=40) for i in range(3)]
    check_batch(handle, regions)


def test_results_do_not_depend_on_speculation_width(handle):
    from breakmer_b200 import batch
    regions = list(synth.config_regions("C2", n=16, start=40)) + [synth.config_region("C1", 0)]
    pk = batch.PackedBatch(regions)
    ref = None
    try:
        for w in (1, 2, 4, 8):
            handle.set_option("spec_width", w)
            out = batch.run(handle, pk)
            got = [(out.sample_only(i), out.contig_records(i)) for i in range(len(regions))]
            if ref is None:
                ref = got
                for i, r in enumerate(regions):
                    only, ctg = oracle_region(r)
                    assert got[i] == (only, ctg)
            else:
                assert got == ref, "spec_width %d changed the result" % w
    finally:
        handle.set_option("spec_width", 0)


def test_capacity_is_reported_per_region_and_the_rest_of_the_batch_completes(handle):
    """A region with a read beyond the DP's length limit gets region_status = BK_ERR_CAPACITY; every other region of
    the same call is processed normally, with read indices in the caller's numbering."""
    from breakmer_b200 import _lib, batch
    long_r = synth.Region(name="long", k=15, ref_fwd="ACGT" * 50, reads=[("@a:1:1:1:1/1_0", "ACGT" * 1100, "I" * 4400, False)],
                          sc_records=[("a", "ACGT" * 10)])
    normal = list(synth.config_regions("C2", n=6, start=3))
    regions = normal[:2] + [long_r] + normal[2:4] + [long_r] + normal[4:]
    out = batch.run(handle, batch.PackedBatch(regions))
    for i, r in enumerate(regions):
        if r is long_r:
            assert out.region_status[i] == _lib.BK_ERR_CAPACITY
            assert out.contig_records(i) == []
        else:
            only, ctg = oracle_region(r)
            assert out.region_status[i] == 0
            assert out.sample_only(i) == only, r.name
            assert out.contig_records(i) == ctg, r.name
    # alone in a call: same report, and the handle stays usable
    out = batch.run(handle, batch.PackedBatch([long_r]))
    assert list(out.region_status) == [_lib.BK_ERR_CAPACITY]
    check_batch(handle, [synth.config_region("C2", 7)])


def test_submit_and_wait_keep_batches_in_flight_from_one_thread(handle):
    """bk_batch_submit / bk_batch_wait: three handles, one host thread, results identical to the one-call form."""
    from breakmer_b200 import _lib, batch
    sets = [list(synth.config_regions("C2", n=8, start=10 * j)) for j in range(3)]
    packed = [batch.PackedBatch(s) for s in sets]
    handles = [_lib.Handle(0) for _ in sets]
    try:
        for rep in range(2):
            for hh, pk in zip(handles, packed):
                batch.submit(hh, pk)
            with pytest.raises(_lib.BreakmerError):
                batch.submit(handles[0], packed[0])          # one batch in flight per handle
            outs = [batch.wait(hh, pk) for hh, pk in zip(handles, packed)]
            for regions, pk, out in zip(sets, packed, outs):
                ref = batch.run(handle, pk)
                for i in range(len(regions)):
                    assert out.sample_only(i) == ref.sample_only(i)
                    assert out.contig_records(i) == ref.contig_records(i)
        with pytest.raises(_lib.BreakmerError):
            batch.wait(handles[0])                           # nothing in flight
    finally:
        for hh in handles:
            hh.close()


def _mutated(region, fn):
    reads = [(rid, fn(i, seq), qual, io) for i, (rid, seq, qual, io) in enumerate(region.reads)]
    return synth.Region(name=region.name + "_m", k=region.k, ref_fwd=region.ref_fwd, reads=reads,
                        sc_records=region.sc_records, normal_reads=region.normal_reads)


def test_odd_inputs_lowercase_n_rich_and_duplicates(handle):
    base = [synth.make_region("odd%d" % i, seed=1200 + i, L=1500, cov=150, k=15, e=0.01,
                              event=[("del", 120, None), ("ins", 40), ("tdup", 200)][i], indel_p=0.3) for i in range(3)]
    # lower-case reads: jellyfish folds case when counting, str.find/== in the assembler do not
    lower = _mutated(base[0], lambda i, s: s.lower() if i % 5 == 0 else s)
    # N-rich reads: 'N' == 'N' scores as a match in olc.nw, windows with N never match a k-mer
    nrich = _mutated(base[1], lambda i, s: (s[:20] + "NNNN" + s[24:]) if i % 3 == 0 else s)
    # every record four times: multiplicities, counts and support vectors scale
    dup = synth.Region(name="dup", k=15, ref_fwd=base[2].ref_fwd,
                       reads=[(rid + "x%d" % j if j else rid, s, q, io) for (rid, s, q, io) in base[2].reads for j in range(4)],
                       sc_records=base[2].sc_records)
    check_batch(handle, [lower, nrich, dup] + base)


@pytest.mark.parametrize("k", [11, 25, 31])
def test_other_k(handle, k):
    regions = [synth.make_region("k%d_%d" % (k, i), seed=1300 + i, L=1200, cov=120, k=k, e=0.005,
                                 event=("del", 150, None), indel_p=0.2) for i in range(3)]
    check_batch(handle, regions)


def test_reference_kmer_cache_gives_the_same_result(handle):
    from breakmer_b200 import _lib, batch
    regions = list(synth.config_regions("C3", n=10)) + list(synth.config_regions("C2", n=10, start=100))
    exp = [oracle_region(r) for r in regions]
    handle.ref_cache_build([r.ref_fwd for r in regions], regions[0].k)
    try:
        out = batch.run(handle, batch.PackedBatch(regions, with_ref=False))
        for i in range(len(regions)):
            assert out.sample_only(i) == exp[i][0]
            assert out.contig_records(i) == exp[i][1]
        # a batch of a different shape must be refused, not silently mis-served
        with pytest.raises(_lib.BreakmerError):
            batch.run(handle, batch.PackedBatch(regions[:5], with_ref=False))
    finally:
        handle.ref_cache_clear()
    with pytest.raises(_lib.BreakmerError):
        batch.run(handle, batch.PackedBatch(regions, with_ref=False))
    check_batch(handle, regions[:4])
