"""Contig hand-off row: the files sv_processor.contig.setup writes
(sv_processor.py:749-782).  CPU part: the restatement oracle/handoff_py.py against golden files
produced by the reference's own writer methods (tests/golden/handoff_cases.json).  GPU part: the
native writer bk_write_contigs on a real batch result against the restatement."""
import json
import os

import pytest

from oracle import handoff_py

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_writers_match_reference():
    with open(os.path.join(HERE, "golden", "handoff_cases.json")) as f:
        gold = json.load(f)["targets"]
    assert sum(len(t["files"]) for t in gold) > 0
    for t in gold:
        contigs = [(seq, [k[0] for k in kmers], [tuple(r) for r in reads]) for seq, kmers, reads in t["contigs"]]
        files, cluster = handoff_py.target_files(contigs)
        if cluster is not None:
            files["clusters.out"] = cluster
        assert files == t["files"]


@pytest.mark.gpu
def test_native_writer_on_a_batch_result(tmp_path):
    from breakmer_b200 import batch, get_handle, ingest, synth
    from test_gpu_api import _write_region_files
    regions = list(synth.config_regions("C2", 12, start=40))
    d = str(tmp_path)
    refs, fqs, scs = [], [], []
    for r in regions:
        ref_f, _rr, fq, sc = _write_region_files(r, d)
        refs.append(ref_f); fqs.append(fq); scs.append(sc)
    g = ingest.Ingest(n_threads=4)
    pk = g.files(refs, fqs, scs, k=15, rc_thresh=regions[0].rc_thresh)
    h = get_handle()
    res = batch.run(h, pk, decode=False)
    dirs = [os.path.join(d, "out", r.name, "contigs") for r in regions]
    dirs[3] = None                                              # a skipped target
    clusters = [os.path.join(d, "out", r.name + "_clusters.out") for r in regions]
    n_files = g.write_contigs(res, pk, dirs, clusters)
    out = batch.BatchOutput(res, pk)
    ids, seqs, quals = pk.read_ids, pk.read_seqs(), pk.read_quals()
    want_files = 0
    n_contigs = 0
    for i, r in enumerate(regions):
        if dirs[i] is None:
            assert not os.path.exists(os.path.join(d, "out", r.name))
            continue
        contigs = []
        for j, rec in enumerate(out.contig_records(i)):
            c = int(out.ctg_reg_off[i]) + j
            ro, nr = out.reads_off[c]
            reads = [(ids[int(x)], seqs[int(x)], quals[int(x)]) for x in out.reads[ro:ro + nr]]
            contigs.append((rec["seq"], [k[0] for k in rec["kmers"]], reads))
        files, cluster = handoff_py.target_files(contigs)
        n_contigs += len(contigs)
        for rel, text in files.items():
            with open(os.path.join(dirs[i], rel), newline="\n") as f:
                assert f.read() == text, (r.name, rel)
        want_files += len(files)
        if cluster is None:
            assert not os.path.exists(clusters[i])
        else:
            with open(clusters[i], newline="\n") as f:
                assert f.read() == cluster
            want_files += 1
        have = sorted(os.listdir(dirs[i])) if os.path.isdir(dirs[i]) else []
        assert have == sorted("contig%d" % (n + 1) for n in range(len(contigs)))
    assert n_contigs > 5 and n_files == want_files
    g.close()
