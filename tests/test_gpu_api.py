"""GPU tests of the reference-shaped Python API (the drop-in boundary): olc.nw,
utils.run_jellyfish/load_kmers, sv_assembly.init_assembly, target.compare_kmers."""
import logging
import os

import pytest

from conftest import golden
from breakmer_b200 import synth
from oracle import assembler_py, kmers_py
from oracle.make_golden import oracle_sample_only, region_scenarios

pytestmark = pytest.mark.gpu


def test_olc_nw_signature_and_values():
    from breakmer_b200 import olc
    cases = golden("nw_golden.json")["cases"][:40]
    for c in cases[:6]:
        assert list(olc.nw(c["seq1"], c["seq2"])) == c["out"]
    got = olc.nw_batch([(c["seq1"], c["seq2"]) for c in cases])
    assert [list(g) for g in got] == [c["out"] for c in cases]
    with pytest.raises(NameError):
        olc.nw("", "ACGT")
    assert (olc.match_award, olc.mismatch_penalty, olc.gap_penalty) == (1, -2, -2)


def test_olc_nw_has_no_length_limit():
    """olc.py:40-52 allocates its tables for any length; the drop-in routes a pair beyond the packed-cell kernels' 4095
    bases to nw_long_kernel and returns the same 7-tuple"""
    import random
    from breakmer_b200 import olc
    from oracle import nw_py
    rng = random.Random(5)
    contig = "".join(rng.choice("ACGT") for _ in range(4500))
    read = contig[4420:] + "".join(rng.choice("ACGT") for _ in range(70))
    assert olc.nw(contig, read) == nw_py.nw_fast(contig, read)
    assert olc.nw(read, contig) == nw_py.nw_fast(read, contig)


def _write_region_files(region, d):
    ref_f = os.path.join(d, region.name + "_forward_refseq.fa")
    ref_r = os.path.join(d, region.name + "_reverse_refseq.fa")
    with open(ref_f, "w") as f:
        f.write(">%s\n%s\n" % (region.name, region.ref_fwd))
    with open(ref_r, "w") as f:
        f.write(">%s\n%s\n" % (region.name, kmers_py.revcomp(region.ref_fwd)))
    fq = os.path.join(d, region.name + "_sv_reads_cleaned_filtered.fastq")
    with open(fq, "w") as f:
        for rid, seq, qual, _io in region.reads:
            f.write("%s\n%s\n+\n%s\n" % (rid, seq, qual))
    sc = os.path.join(d, region.name + "_sv_sc_seqs.fa")
    with open(sc, "w") as f:
        for name, seq in region.sc_records:
            f.write(">%s\n%s\n" % (name, seq))
    return ref_f, ref_r, fq, sc


def test_run_jellyfish_and_load_kmers_like_the_reference(tmp_path):
    from breakmer_b200 import utils
    name, kw = region_scenarios()[4]
    region = synth.make_region(name, **kw)
    ref_f, ref_r, fq, sc = _write_region_files(region, str(tmp_path))
    k = region.k
    # sv_processor.py:613-620, verbatim shape
    ref = {}
    for fn in (ref_f, ref_r):
        ref = utils.load_kmers(utils.run_jellyfish(fn, "/no/such/jellyfish", k), ref)
    case = utils.load_kmers(utils.run_jellyfish(fq, "jellyfish", k), {})
    case_sc = utils.load_kmers(utils.run_jellyfish(sc, "jellyfish", k), {})
    oref, ocase, osc, only = oracle_sample_only(region)
    assert ref == oref and case == ocase and case_sc == osc
    sample_only = list((set(case) & set(case_sc)).difference(set(ref)))
    assert {m: case[m] for m in sample_only} == only
    # marker-file cache (utils.py:157): a second call must not recount
    dump = fq + "_%dmers_dump" % k
    assert os.path.isfile(utils.get_marker_fn(dump))
    os.remove(dump)
    assert utils.run_jellyfish(fq, "jellyfish", k) == dump and not os.path.isfile(dump)
    # comma separated list accumulates (utils.py:288-295)
    both = utils.load_kmers(",".join([ref_f + "_%dmers_dump" % k, ref_r + "_%dmers_dump" % k]), {})
    assert both == oref


def _fq_recs(region):
    from breakmer_b200.utils import fq_read
    fq_recs = {}
    for rid, seq, qual, io in region.reads:
        fr = fq_read(rid, seq, qual, io)
        fq_recs.setdefault(fr.seq, []).append(fr)
    return fq_recs


def test_init_assembly_drop_in():
    from breakmer_b200 import sv_assembly
    cases = [c for c in golden("assembly_golden.json")["cases"] if "contigs" in c][:12]
    for c in cases:
        kw = dict(c["kwargs"]); kw["event"] = tuple(kw["event"])
        region = synth.make_region(c["name"], **kw)
        _r, _c, _s, only = oracle_sample_only(region)
        ctgs = sv_assembly.init_assembly(only, _fq_recs(region), region.k, region.rc_thresh, region.read_len)
        got = [{"seq": ct.get_contig_seq(), "indel_only": ct.get_contig_counts().indel_only,
                "others": ct.get_contig_counts().others, "reads": sorted(r.id for r in ct.reads),
                "kmers": [list(t) for t in ct.kmers], "kmer_locs": ct.get_kmer_locs()} for ct in ctgs]
        assert got == c["contigs"], c["name"]               # the reference's own output
        for ct in ctgs:
            assert ct.get_contig_len() == len(ct.aseq.seq)
            assert ct.get_total_read_support() == max(ct.aseq.counts.indel_only) + max(ct.aseq.counts.others)
            n = len(ct.aseq.counts.others)
            assert ct.aseq.counts.get_counts(0, min(5, n), 'indel') == \
                [a + b for a, b in zip(ct.aseq.counts.indel_only[:5], ct.aseq.counts.others[:5])]
    assert sv_assembly.init_assembly({}, {}, 15, 2, 100) == []


class _Params:
    def __init__(self, k):
        self.k = k
        self.opts = {"jellyfish": "jellyfish"}

    def get_kmer_size(self):
        return self.k

    def get_sr_thresh(self, kind):
        assert kind == 'min'
        return 2


class _Target:
    """The slice of sv_processor.target that compare_kmers touches."""

    def __init__(self, region, d):
        ref_f, ref_r, fq, sc = _write_region_files(region, d)
        self.name = region.name
        self.params = _Params(region.k)
        self.files = {"target_ref_fn": [ref_f, ref_r], "cleaned_fq": fq, "sv_sc_unmapped_fa": sc}
        self.paths = {"kmers": d}
        self.kmers = {}
        self.cleaned_read_recs = _fq_recs(region)
        self.read_len = region.read_len
        self.logger = logging.getLogger("root")


def test_compare_kmers_single_and_batched(tmp_path):
    from breakmer_b200 import sv_processor
    regions = [synth.make_region(n, **kw) for n, kw in region_scenarios()[:8] if kw["k"] == 15]
    targets = [_Target(r, str(tmp_path)) for r in regions]
    sv_processor.compare_kmers(targets[0])
    sv_processor.compare_kmers_batch(targets[1:])
    for r, t in zip(regions, targets):
        _a, _b, _c, only = oracle_sample_only(r)
        exp = assembler_py.init_assembly(only, r.reads, r.k, r.rc_thresh, r.read_len)
        with open(t.files["sample_kmers"]) as f:
            lines = dict(l.split("\t") for l in f.read().splitlines())
        assert {m: int(c) for m, c in lines.items()} == only
        assert [ct.get_contig_seq() for ct in t.kmers["clusters"]] == [e["seq"] for e in exp]
        assert [sorted(x.id for x in ct.reads) for ct in t.kmers["clusters"]] == [e["reads"] for e in exp]
        assert t.cleaned_read_recs is None and t.kmers["case_only"] == {} and t.kmers["ref"] == {}
        assert t.files["kmer_clusters"].endswith("_sample_kmers_merged.out")


def test_compare_kmers_native_ingest(tmp_path):
    """ingest="native": the same targets through bk_ingest_files."""
    from breakmer_b200 import sv_processor
    regions = [synth.make_region(n, **kw) for n, kw in region_scenarios()[:10] if kw["k"] == 15]
    targets = [_Target(r, str(tmp_path)) for r in regions]
    for t in targets:
        t.paths["contigs"] = os.path.join(str(tmp_path), t.name, "contigs")
    sv_processor.compare_kmers_batch(targets, ingest="native", write_contigs=True)
    for r, t in zip(regions, targets):
        for n, ct in enumerate(t.kmers["clusters"], 1):
            with open(os.path.join(t.paths["contigs"], "contig%d" % n, "contig%d.fa" % n)) as f:
                assert f.read() == ">contig1\n" + ct.get_contig_seq()
            with open(os.path.join(t.paths["contigs"], "contig%d" % n, "contig%d.fq" % n)) as f:
                assert sorted(f.read().splitlines()[0::4]) == sorted(x.id for x in ct.reads)
        _a, _b, _c, only = oracle_sample_only(r)
        exp = assembler_py.init_assembly(only, r.reads, r.k, r.rc_thresh, r.read_len)
        with open(t.files["sample_kmers"]) as f:
            lines = dict(l.split("\t") for l in f.read().splitlines())
        assert {m: int(c) for m, c in lines.items()} == only
        got = [{"seq": ct.get_contig_seq(), "indel_only": ct.get_contig_counts().indel_only,
                "others": ct.get_contig_counts().others, "reads": sorted(x.id for x in ct.reads),
                "kmers": [list(k) for k in ct.kmers], "kmer_locs": ct.get_kmer_locs()} for ct in t.kmers["clusters"]]
        assert got == exp, r.name
        by_id = {rec[0]: rec for rec in r.reads}
        for ct in t.kmers["clusters"]:
            for fr in ct.reads:
                rid, seq, qual, io = by_id[fr.id]
                assert (fr.seq, fr.qual, bool(fr.indel_only)) == (seq, qual, bool(io))
        assert t.cleaned_read_recs is None


def test_ingested_batch_equals_packed_batch(tmp_path):
    """Files -> Ingest -> bk_compare_kmers_batch gives the records PackedBatch gives, normal sample included."""
    from breakmer_b200 import batch, get_handle, ingest
    regions = list(synth.config_regions("C3", 6)) + list(synth.config_regions("C2", 4, start=17))
    d = str(tmp_path)
    refs, fqs, scs, nms = [], [], [], []
    for r in regions:
        ref_f, _ref_r, fq, sc = _write_region_files(r, d)
        nm = None
        if r.normal_reads:
            nm = os.path.join(d, r.name + "_normal.fastq")
            with open(nm, "w") as f:
                for rec in r.normal_reads:
                    f.write("%s\n%s\n+\n%s\n" % (rec[0] if rec[0].startswith("@") else "@" + rec[0], rec[1], "I" * len(rec[1])))
        refs.append(ref_f); fqs.append(fq); scs.append(sc); nms.append(nm)
    h = get_handle()
    want = batch.run(h, batch.PackedBatch(regions, with_normal=True))
    want_rec = [(want.sample_only(i), want.contig_records(i)) for i in range(len(regions))]
    g = ingest.Ingest(n_threads=4)                      # pinned buffer
    pk = g.files(refs, fqs, scs, normal=nms, k=15, rc_thresh=regions[0].rc_thresh)
    got = batch.run(h, pk)
    assert [(got.sample_only(i), got.contig_records(i)) for i in range(len(regions))] == want_rec
    assert sum(len(c) for _s, c in want_rec) > 0
    g.close()


def test_sharded_compare_kmers_batch_over_devices(tmp_path):
    """compare_kmers_batch(targets, devices=[...]): region-sharded multi-GPU entry (one host thread + its own handles
    per device entry, chunks handed out dynamically).  Two workers on device 0 exercise the same code on a one-GPU box."""
    from breakmer_b200 import sv_processor
    regions = list(synth.config_regions("C2", 14, start=60)) + list(synth.config_regions("C5", 30, start=90))
    d1, d2 = os.path.join(str(tmp_path), "a"), os.path.join(str(tmp_path), "b")
    os.makedirs(d1); os.makedirs(d2)
    single = [_Target(r, d1) for r in regions]
    sharded = [_Target(r, d2) for r in regions]
    sv_processor.compare_kmers_batch(single, ingest="native")
    sv_processor.compare_kmers_batch(sharded, ingest="native", devices=[0, 0], max_targets=7)
    n_ctg = 0
    for a, b in zip(single, sharded):
        with open(a.files["sample_kmers"]) as fa, open(b.files["sample_kmers"]) as fb:
            assert fa.read() == fb.read()
        assert len(a.kmers["clusters"]) == len(b.kmers["clusters"])
        for x, y in zip(a.kmers["clusters"], b.kmers["clusters"]):
            assert x.get_contig_seq() == y.get_contig_seq() and x.kmers == y.kmers and x.get_kmer_locs() == y.get_kmer_locs()
            assert x.get_contig_counts().indel_only == y.get_contig_counts().indel_only
            assert sorted(r.id for r in x.reads) == sorted(r.id for r in y.reads)
            n_ctg += 1
        assert b.cleaned_read_recs is None
    assert n_ctg > 10
    # python marshalling through the same sharded path
    third = [_Target(r, d2) for r in regions[:10]]
    sv_processor.compare_kmers_batch(third, devices=[0, 0], max_targets=3)
    for a, b in zip(single, third):
        assert [c.get_contig_seq() for c in a.kmers["clusters"]] == [c.get_contig_seq() for c in b.kmers["clusters"]]


def test_compare_kmers_batch_with_an_applying_thread(tmp_path):
    """apply_thread=True: one host thread parses and submits, a second waits for the results and applies them.  Same
    state on the targets and the same contig files as the single-thread pass; an over-limit target is still isolated."""
    from breakmer_b200 import sv_processor
    regions = list(synth.config_regions("C2", 18, start=120)) + list(synth.config_regions("C5", 25, start=300))
    long_r = synth.Region(name="zz_long", k=15, ref_fwd="ACGT" * 50, reads=[("@a:1:1:1:1/1_0", "ACGT" * 1100, "I" * 4400, False)],
                          sc_records=[("a", "ACGT" * 10)])
    d1, d2 = os.path.join(str(tmp_path), "a"), os.path.join(str(tmp_path), "b")
    os.makedirs(d1); os.makedirs(d2)
    plain = [_Target(r, d1) for r in regions]
    piped = [_Target(r, d2) for r in regions + [long_r]]
    for ts in (plain, piped):
        for t in ts:
            t.paths["contigs"] = os.path.join(t.paths["kmers"], t.name, "contigs")
    sv_processor.compare_kmers_batch(plain, ingest="native", write_contigs=True, max_targets=1000)
    with pytest.raises(sv_processor.CapacityError) as e:
        sv_processor.compare_kmers_batch(piped, ingest="native", write_contigs=True, max_targets=6, inflight=3, apply_thread=True)
    assert e.value.targets == ["zz_long"]
    assert piped[-1].cleaned_read_recs is not None and "clusters" not in piped[-1].kmers
    n_ctg = 0
    for a, b in zip(plain, piped):
        with open(a.files["sample_kmers"]) as fa, open(b.files["sample_kmers"]) as fb:
            assert fa.read() == fb.read()
        assert len(a.kmers["clusters"]) == len(b.kmers["clusters"])
        for n, (x, y) in enumerate(zip(a.kmers["clusters"], b.kmers["clusters"]), 1):
            assert x.get_contig_seq() == y.get_contig_seq() and x.kmers == y.kmers and x.get_kmer_locs() == y.get_kmer_locs()
            assert x.get_contig_counts().others == y.get_contig_counts().others
            assert sorted(r.id for r in x.reads) == sorted(r.id for r in y.reads)
            for ext in ("fa", "fq"):
                with open(os.path.join(a.paths["contigs"], "contig%d" % n, "contig%d.%s" % (n, ext))) as fa, \
                        open(os.path.join(b.paths["contigs"], "contig%d" % n, "contig%d.%s" % (n, ext))) as fb:
                    assert fa.read() == fb.read()
            n_ctg += 1
        assert b.cleaned_read_recs is None
    assert n_ctg > 10
    # python marshalling through the same two-thread pass, twice in a row (the pipeline and its writer are kept)
    for _ in range(2):
        third = [_Target(r, d2) for r in regions[:12]]
        sv_processor.compare_kmers_batch(third, max_targets=5, apply_thread=True)
        for a, b in zip(plain, third):
            assert [c.get_contig_seq() for c in a.kmers["clusters"]] == [c.get_contig_seq() for c in b.kmers["clusters"]]


def test_run_sharded_matches_one_call():
    from breakmer_b200 import _lib, batch, shard
    regions = list(synth.config_regions("C2", 20, start=200))
    h = _lib.Handle(0)
    try:
        pk = batch.PackedBatch(regions)
        want = shard.region_digests(batch.run(h, pk), pk)
    finally:
        h.close()
    chunks = {}

    def on_result(idx, out, packed):
        chunks[tuple(idx)] = shard.region_digests(out, packed)

    got = shard.run_sharded(regions, devices=(0, 0), max_regions=6, inflight=2, on_result=on_result)
    assert list(got) == sorted(r.name for r in regions)
    assert sorted(i for idx in chunks for i in idx) == list(range(len(regions))) and len(chunks) >= 4
    merged = {}
    for d in chunks.values():
        merged.update(d)
    assert merged == want
    assert shard.digest_of_digests(merged) == shard.digest_of_digests(want)
    for name, (out, j) in got.items():
        assert out.region_status[j] == 0


def test_one_target_over_the_limit_does_not_disturb_the_batch(tmp_path):
    from breakmer_b200 import sv_processor
    regions = list(synth.config_regions("C2", 5, start=30))
    long_r = synth.Region(name="zz_long", k=15, ref_fwd="ACGT" * 50, reads=[("@a:1:1:1:1/1_0", "ACGT" * 1100, "I" * 4400, False)],
                          sc_records=[("a", "ACGT" * 10)])
    targets = [_Target(r, str(tmp_path)) for r in regions[:2] + [long_r] + regions[2:]]
    for ingest in ("python", "native"):
        for t, r in zip(targets, regions[:2] + [long_r] + regions[2:]):
            t.cleaned_read_recs = _fq_recs(r)
            t.kmers = {}
        with pytest.raises(sv_processor.CapacityError) as e:
            sv_processor.compare_kmers_batch(targets, ingest=ingest)
        assert e.value.targets == ["zz_long"]
        for t in targets:
            if t.name == "zz_long":
                assert t.cleaned_read_recs is not None and "clusters" not in t.kmers     # untouched
            else:
                r = next(x for x in regions if x.name == t.name)
                _a, _b, _c, only = oracle_sample_only(r)
                exp = assembler_py.init_assembly(only, r.reads, r.k, r.rc_thresh, r.read_len)
                assert [c.get_contig_seq() for c in t.kmers["clusters"]] == [x["seq"] for x in exp]
                assert t.cleaned_read_recs is None


def test_init_assembly_rejects_mers_it_cannot_represent():
    from breakmer_b200 import sv_assembly
    r = synth.config_region("C2", 3)
    with pytest.raises(ValueError):
        sv_assembly.init_assembly({"ACGTN" * 3: 3}, _fq_recs(r), 15, 2, 100)
    with pytest.raises(ValueError):
        sv_assembly.init_assembly({"ACGT": 3}, _fq_recs(r), 15, 2, 100)
