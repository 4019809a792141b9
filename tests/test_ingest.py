"""Ingest row: the native text readers bk_ingest_* against
(1) golden vectors produced by the reference's own FastqFile / get_fastq_reads
(tests/golden/ingest_cases.json, oracle/make_golden_ingest.py) and (2) the CPU
restatement oracle/ingest_py.py.  Host code only -- runs without a GPU (pinned=False)."""
import json
import os
import random
from collections import OrderedDict

import numpy as np
import pytest

from breakmer_b200 import ingest, utils
from oracle import ingest_py

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ing():
    g = ingest.Ingest(n_threads=4, pinned=False)
    yield g
    g.close()


def golden_cases():
    with open(os.path.join(HERE, "golden", "ingest_cases.json")) as f:
        return json.load(f)["cases"]


CASES = golden_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_fastq_reader(case):
    if "error" in case:
        with pytest.raises(ValueError):
            ingest_py.fastq_records(case["text"])
        return
    assert [list(r) for r in ingest_py.fastq_records(case["text"])] == case["records"]
    if "fq_recs" in case:
        recs, read_len = ingest_py.fq_recs(case["text"])
        assert [[s, ids] for s, ids in recs.items()] == [[s, ids] for s, ids, _f in case["fq_recs"]]
        assert read_len == case["read_len"]
        for s, ids, flags in case["fq_recs"]:
            assert [ingest_py.indel_only_suffix(i) for i in ids] == flags


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_native_matches_reference_fastq_reader(ing, case):
    if "error" in case:
        with pytest.raises(ValueError):
            ing.texts([None], [case["text"]], [None])
        return
    pk = ing.texts([None], [case["text"]], [None])
    got = [list(t) for t in zip(pk.read_ids, pk.read_seqs(), pk.read_quals())]
    assert got == case["records"]
    assert pk.n_reads == len(case["records"])
    if "fq_recs" in case:
        # the record model of get_fastq_reads: group by sequence in first-occurrence order
        recs = OrderedDict()
        for i, (h, s, _q) in enumerate(got):
            recs.setdefault(s, []).append((h, bool(pk.read_flags[i])))
        assert [[s, [h for h, _ in g], [f for _, f in g]] for s, g in recs.items()] == case["fq_recs"]
        assert int(pk.read_len[0]) == case["read_len"]


def test_dropin_fastqfile_matches_reference(tmp_path):
    for case in CASES:
        fn = tmp_path / (case["name"] + ".fastq")
        with open(fn, "w", newline="\n") as f:
            f.write(case["text"])
        if "error" in case:
            with pytest.raises((ValueError, NameError)):
                list(utils.FastqFile(str(fn)))
        else:
            assert [list(r) for r in utils.FastqFile(str(fn))] == case["records"]


KMER_TEXTS = [
    "",
    ">r1\nACGT\nTTGA\n>r2\nCC\n",
    ">only\nACGTACGT",
    "junk before\n>r\nAC\n\nGT\n>empty\n>last\n  TT  \n",
    ">crlf\r\nACGT\r\nGG\r\n",
    "@a:1:2:3:4/1_0\nACGT\n+\nIIII\n@a:1:2:3:5/1_0\nGGCC\n+\nIIII\n",
    "@a:1:2:3:4/1_0\nACGT\n+\nIIII\n@partial\nTT\n",
    "no header at all\nACGT\n",
    ">n\nacgtnNACGT\n",
]


@pytest.mark.parametrize("text", KMER_TEXTS)
def test_native_kmer_input_reader_matches_oracle(ing, text, tmp_path):
    pk = ing.texts([text], [None], [text], normal=[text])
    assert pk.sequences("sc") == [ingest_py.kmer_sequences(text)]
    assert pk.sequences("normal") == [ingest_py.kmer_sequences(text)]
    ref = ingest_py.kmer_sequences(text, first_only=True)
    assert pk.sequences("ref") == [ref if ref else [""]]
    fn = tmp_path / "x.fa"
    with open(fn, "w", newline="\n") as f:
        f.write(text)
    assert utils.read_sequences(str(fn)) == ingest_py.kmer_sequences(text)


def _random_region_texts(rng, i):
    dna = lambda n: "".join(rng.choice("ACGTN" if rng.random() < 0.05 else "ACGT") for _ in range(n))  # noqa: E731
    ref = ">chr%d:1-100\n" % i + "\n".join(dna(60) for _ in range(rng.randint(0, 5))) + "\n"
    n = rng.randint(0, 40)
    pool = [dna(rng.randint(30, 120)) for _ in range(max(1, n // 3))]
    reads = "".join("@X%d:1:%d:%d:%d/%d_%d\n%s\n+\n%s\n" % (i, j, j + 3, j + 9, 1 + (j & 1), rng.randint(0, 1), s, "I" * len(s))
                    for j, s in ((j, rng.choice(pool)) for j in range(n)))
    sc = "".join(">q%d\n%s\n" % (j, dna(rng.randint(5, 40))) for j in range(rng.randint(0, 10)))
    normal = "".join("@N%d:1:%d:2:3/1_0\n%s\n+\n%s\n" % (i, j, s, "I" * len(s))
                     for j, s in ((j, dna(rng.randint(30, 90))) for j in range(rng.randint(0, 12))))
    return ref, reads, sc, normal


def _check_batch(pk, texts):
    n = len(texts)
    ids, seqs, quals = pk.read_ids, pk.read_seqs(), pk.read_quals()
    sc, nm, ref = pk.sequences("sc"), pk.sequences("normal"), pk.sequences("ref")
    assert int(pk.read_reg_off[0]) == 0 and int(pk.read_reg_off[n]) == pk.n_reads
    for r, (t_ref, t_reads, t_sc, t_nm) in enumerate(texts):
        a, b = int(pk.read_reg_off[r]), int(pk.read_reg_off[r + 1])
        want = ingest_py.fastq_records(t_reads or "")
        assert [tuple(x) for x in zip(ids[a:b], seqs[a:b], quals[a:b])] == want
        assert [bool(f) for f in pk.read_flags[a:b]] == [ingest_py.indel_only_suffix(h) for h, _s, _q in want]
        assert int(pk.read_len[r]) == max([len(s) for _h, s, _q in want] + [0])
        assert sc[r] == ingest_py.kmer_sequences(t_sc or "")
        assert nm[r] == ingest_py.kmer_sequences(t_nm or "")
        assert ref[r] == (ingest_py.kmer_sequences(t_ref or "", first_only=True) or [""])


@pytest.mark.parametrize("threads", [1, 3, 16])
def test_native_batch_layout_texts_and_files(threads, tmp_path):
    rng = random.Random(99 + threads)
    texts = [_random_region_texts(rng, i) for i in range(37)]
    texts[5] = (None, None, None, None)                      # a region with no inputs at all
    texts[11] = (texts[11][0], "", "", "")
    g = ingest.Ingest(n_threads=threads, pinned=False)
    cols = list(zip(*texts))
    pk = g.texts(cols[0], cols[1], cols[2], normal=cols[3], k=15, rc_thresh=2)
    assert pk.n == 37 and pk.k == 15 and pk.struct().rc_thresh == 2 and pk.has_normal
    _check_batch(pk, texts)
    # the same through files
    paths = []
    for r, row in enumerate(texts):
        p = []
        for s, t in enumerate(row):
            if t is None:
                p.append(None)
                continue
            fn = tmp_path / ("r%d_%d.txt" % (r, s))
            with open(fn, "w", newline="\n") as f:
                f.write(t)
            p.append(str(fn))
        paths.append(p)
    pc = list(zip(*paths))
    pk2 = g.files(pc[0], pc[1], pc[2], normal=pc[3])
    _check_batch(pk2, texts)
    # a second, smaller call reuses the buffer
    pk3 = g.texts(cols[0][:3], cols[1][:3], cols[2][:3])
    assert not pk3.has_normal and pk3.struct().normal_bases is None
    _check_batch(pk3, [(a, b, c, None) for a, b, c, _d in texts[:3]])
    g.close()


def test_missing_file_and_flag_override(ing, tmp_path):
    with pytest.raises(IOError):
        ing.files([str(tmp_path / "absent.fa")], [None], [None])
    pk = ing.texts([None], ["@a:1:2:3:4/1_0\nACGT\n+\nIIII\n@a:1:2:3:5/1_1\nACGA\n+\nIIII\n"], [None])
    assert pk.read_flags.tolist() == [0, 1]
    pk.read_flags[:] = np.array([1, 0], np.uint8)             # the caller's own fq_read.indel_only values
    fl = pk.struct().read_flags
    import ctypes
    assert list((ctypes.c_uint8 * 2).from_address(fl)) == [1, 0]


def test_parsers_under_address_sanitizer():
    """tests/sim/ingest_fuzz.cpp: the readers of csrc/ingest.cuh compiled by g++ with ASan + UBSan and driven with
    random and corrupted texts (layout invariants checked, memory errors abort)."""
    import shutil
    import subprocess
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if shutil.which("g++") is None or not os.path.isdir(os.path.join(cuda, "include")):
        pytest.skip("g++ / CUDA headers not available")
    sim = os.path.join(HERE, "sim")
    exe = os.path.join(sim, "ingest_fuzz")
    src = os.path.join(sim, "ingest_fuzz.cpp")
    dep = os.path.join(HERE, "..", "breakmer_b200", "csrc", "ingest.cuh")
    if not os.path.isfile(exe) or max(os.path.getmtime(src), os.path.getmtime(dep)) > os.path.getmtime(exe):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-I", os.path.join(cuda, "include"),
                               "-o", exe, src, "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lpthread"])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(cuda, "lib64") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    out = subprocess.run([exe, "1500"], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "parsed" in out.stdout
