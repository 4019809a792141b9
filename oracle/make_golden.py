"""Generate tests/golden/*.json by running the REFERENCE ITSELF (via ref_shim).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Run in the build
container, where /root/reference exists:

    python -m oracle.make_golden            # writes tests/golden/
    python -m oracle.make_golden --fuzz N   # additionally cross-checks the oracle
                                            # against the reference on N extra
                                            # regions / 20*N extra nw pairs (not stored)

Fixtures written:
  nw_golden.json        inputs + the reference's `olc.nw` 7-tuples: the three
                        example pairs that sit in comments at olc.py:11-16 plus
                        seeded random pairs (overlaps, containments, identical,
                        unrelated, low-complexity tie cases, reads with N).
  assembly_golden.json  for seeded synthetic regions (breakmer_b200.synth): the
                        reference's `init_assembly` output (contig sequence,
                        both count vectors, read ids, k-mer 5-tuples,
                        kmer_locs) -- in full for small regions, as a SHA-256
                        digest for the rest.  Inputs are named by their
                        generator arguments and pinned by an input digest.
  kmers_golden.json     oracle-derived known-answer vectors for the k-mer stage
                        ("parity unpinned": jellyfish is not available; see
                        oracle/kmers_py.py).

For the bulk assembly cases the reference's `olc.nw` is replaced by the C
restatement for speed, *after* this script has proven the two equal on every nw
pair it generates; a subset of regions (`literal_nw: true`) runs with the
reference's own pure-Python `olc.nw`.
"""
import argparse
import hashlib
import json
import os
import random
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_shim, nw_py, kmers_py, assembler_py   # noqa: E402
from breakmer_b200 import synth                               # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


# --------------------------------------------------------------------------
def olc_example_pairs():
    """The commented example strings at olc.py:11-16 (inputs only)."""
    with open(os.path.join(ref_shim.REFERENCE_ROOT, "olc.py")) as f:
        lines = f.read().splitlines()[10:16]
    seqs = {}
    order = []
    for ln in lines:
        m = re.match(r"#(seq\d\d)\s*=\s*['\"]([ACGTN]+)['\"]", ln.strip())
        if m:
            order.append((m.group(1), m.group(2)))
    pairs = []
    for i in range(0, len(order), 2):
        d = dict(order[i:i + 2])
        pairs.append((d["seq11"], d["seq22"]))
    return pairs


def _mut(rng, s, e):
    out = []
    for c in s:
        r = rng.random()
        if r < e / 3:
            continue                                   # deletion
        if r < 2 * e / 3:
            out.append(rng.choice("ACGT"))             # insertion
        if r < e:
            out.append(rng.choice("ACGTN"))
        else:
            out.append(c)
    return "".join(out) or "A"


def random_nw_pairs(rng, n):
    pairs = []
    for t in range(n):
        kind = t % 8
        la, lb = rng.randint(15, 260), rng.randint(15, 160)
        alpha = "ACGT" if kind != 5 else "AC"
        g = "".join(rng.choice(alpha) for _ in range(la + lb + 60))
        if kind == 0:      # suffix/prefix overlap
            ov = rng.randint(5, min(la, lb))
            a, b = g[:la], g[la - ov:la - ov + lb]
        elif kind == 1:    # b contained in a
            p = rng.randint(0, max(0, la - lb))
            a, b = g[:la], g[p:p + min(lb, la)]
        elif kind == 2:    # identical
            a = g[:la]
            b = a
        elif kind == 3:    # unrelated
            a, b = g[:la], "".join(rng.choice("ACGT") for _ in range(lb))
        elif kind == 4:    # prefix/suffix overlap the other way
            ov = rng.randint(5, min(la, lb))
            b, a = g[:lb], g[lb - ov:lb - ov + la]
        elif kind == 5:    # low complexity (many score ties)
            a, b = g[:la], g[rng.randint(0, 20):][:lb]
        elif kind == 6:    # homopolymer runs
            a = "A" * rng.randint(5, 40) + g[:la] + "T" * rng.randint(1, 30)
            b = g[la // 2:la] + "T" * rng.randint(1, 40)
        else:              # very short
            a, b = g[:rng.randint(1, 12)], g[3:3 + rng.randint(1, 12)]
        e = rng.choice([0.0, 0.0, 0.01, 0.03, 0.1])
        a, b = _mut(rng, a, e), _mut(rng, b, e)
        if rng.random() < 0.5:
            a, b = b, a
        pairs.append((a, b))
    return pairs


# --------------------------------------------------------------------------
def region_scenarios():
    """(name, make_region kwargs).  Small enough that the whole set runs in a
    couple of minutes through the reference."""
    sc = []
    ev = [("del", 300, None), ("ins", 40), ("inv", 400), ("tdup", 300), ("trl",)]
    i = 0
    for k in (15, 21):
        for e in (0.0, 0.005, 0.02):
            for event in ev:
                for cov in (60, 200):
                    i += 1
                    sc.append(("g%03d" % i, dict(seed=9000 + i, L=1500 + 37 * i, cov=cov, k=k, e=e, event=event,
                                                vaf=0.5 if i % 3 else 1.0, indel_p=0.3 if i % 2 else 0.0,
                                                rl_jitter=(8 if i % 4 == 0 else 0),
                                                spurious_frac=(0.01 if i % 5 == 0 else 0.0))))
    # deep amplicon (config 4 shape), translocation with spurious reads (config 5 shape), no event
    sc.append(("amp1", dict(seed=9501, L=400, cov=1200, k=21, e=0.005, event=("del", 60, None), indel_p=0.3)))
    sc.append(("amp2", dict(seed=9502, L=350, cov=800, k=15, e=0.01, event=("ins", 40), indel_p=0.5)))
    sc.append(("spur", dict(seed=9503, L=3000, cov=100, k=15, e=0.005, event=("none",), spurious_frac=0.02)))
    sc.append(("trl5", dict(seed=9504, L=2500, cov=100, k=15, e=0.005, event=("trl",), spurious_frac=0.01)))
    sc.append(("short", dict(seed=9505, L=600, cov=150, k=15, e=0.01, event=("del", 45, None), rl=60, rl_jitter=20)))
    sc.append(("c1", dict(seed=1, L=20000, cov=200, k=15, e=0.002, event=("del", 1500, 9450), indel_p=0.3)))
    return sc


def region_inputs_digest(r):
    h = hashlib.sha256()
    h.update(r.ref_fwd.encode())
    for rec in r.reads:
        h.update(("%s|%s|%d;" % (rec[0], rec[1], int(rec[3]))).encode())
    for rec in r.sc_records:
        h.update(("%s|%s;" % rec).encode())
    for rec in r.normal_reads:
        h.update(("%s|%s;" % rec).encode())
    return h.hexdigest()


def digest(obj):
    return hashlib.sha256(json.dumps(obj, sort_keys=True, separators=(",", ":")).encode()).hexdigest()


def reference_assembly(asm_mod, region, mers):
    fq_recs = {}
    for rid, seq, qual, io in region.reads:
        fr = ref_shim.fq_read(rid, seq, qual, io)
        fq_recs.setdefault(fr.seq, []).append(fr)            # utils.py:239-244
    cts = asm_mod.init_assembly(dict(mers), fq_recs, region.k, region.rc_thresh, region.read_len)
    out = []
    for ct in cts:
        out.append({
            "seq": ct.get_contig_seq(),
            "indel_only": list(ct.get_contig_counts().indel_only),
            "others": list(ct.get_contig_counts().others),
            "reads": sorted(r.id for r in ct.reads),
            "kmers": [list(t) for t in ct.kmers],
            "kmer_locs": list(ct.get_kmer_locs()),
        })
    return out


def oracle_sample_only(region):
    normal = [x[1] for x in region.normal_reads] if region.normal_reads else None
    return kmers_py.sample_only(region.ref_fwd, [x[1] for x in region.reads],
                                [x[1] for x in region.sc_records], region.k, normal)


# --------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fuzz", type=int, default=0)
    args = ap.parse_args()
    if not ref_shim.available():
        sys.exit("reference tree not present; golden vectors can only be generated where it is")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    olc, asm_mod = ref_shim.load()
    if nw_py.c_lib() is None:
        sys.exit("build oracle/c first (make -C oracle/c)")

    # ---- nw ---------------------------------------------------------------
    rng = random.Random(20261017)
    pairs = []
    for a, b in olc_example_pairs():
        pairs += [(a, b), (b, a)]
    n_examples = len(pairs)
    pairs += random_nw_pairs(rng, 400)
    cases = []
    for a, b in pairs:
        ref = list(olc.nw(a, b))
        assert ref == list(nw_py.nw(a, b)) == list(nw_py.nw_fast(a, b)), (a, b)
        cases.append({"seq1": a, "seq2": b, "out": ref})
    with open(os.path.join(GOLDEN_DIR, "nw_golden.json"), "w") as f:
        json.dump({"source": "olc.nw of /root/reference run through oracle/ref_shim.py",
                   "n_reference_examples": n_examples, "cases": cases}, f, separators=(",", ":"))
    print("nw_golden.json: %d cases (first %d are olc.py:11-16)" % (len(cases), n_examples))
    for a, b in random_nw_pairs(random.Random(7), 20 * args.fuzz):
        assert list(olc.nw(a, b)) == list(nw_py.nw(a, b)) == list(nw_py.nw_fast(a, b)), (a, b)

    # ---- assembly ---------------------------------------------------------------
    ref_nw_literal = olc.nw
    out_cases = []
    for name, kw in region_scenarios():
        region = synth.make_region(name, **kw)
        _ref, _case, _sc, only = oracle_sample_only(region)
        literal = (len(region.reads) <= 120)
        olc.nw = ref_nw_literal if literal else nw_py.nw_fast
        ref_out = reference_assembly(asm_mod, region, only)
        mine = assembler_py.init_assembly(only, region.reads, region.k, region.rc_thresh, region.read_len)
        if mine != ref_out:
            sys.exit("ORACLE != REFERENCE on scenario %s" % name)
        entry = {"name": name, "kwargs": kw, "inputs_sha256": region_inputs_digest(region),
                 "literal_nw": literal, "n_reads": len(region.reads), "n_sample_only": len(only),
                 "n_contigs": len(ref_out), "contigs_sha256": digest(ref_out)}
        if len(region.reads) <= 260 and len(json.dumps(ref_out)) < 60000:
            entry["contigs"] = ref_out
        out_cases.append(entry)
        print("  %-6s reads=%4d only=%4d contigs=%d %s" % (name, len(region.reads), len(only), len(ref_out),
                                                          "literal" if literal else ""))
    olc.nw = ref_nw_literal
    with open(os.path.join(GOLDEN_DIR, "assembly_golden.json"), "w") as f:
        json.dump({"source": "init_assembly of /root/reference run through oracle/ref_shim.py",
                   "cases": out_cases}, f, separators=(",", ":"))
    print("assembly_golden.json: %d regions" % len(out_cases))

    # ---- k-mers (oracle derived) ----------------------------------------------
    kcases = []
    for name, kw in region_scenarios()[:12] + region_scenarios()[-6:-1]:
        region = synth.make_region(name, **kw)
        ref, case, sc, only = oracle_sample_only(region)
        kcases.append({"name": name, "kwargs": kw, "inputs_sha256": region_inputs_digest(region),
                       "n_ref": len(ref), "n_case": len(case), "n_sc": len(sc),
                       "ref_sha256": digest(sorted(ref.items())), "case_sha256": digest(sorted(case.items())),
                       "sc_sha256": digest(sorted(sc.items())),
                       "sample_only": sorted(only.items())})
    with open(os.path.join(GOLDEN_DIR, "kmers_golden.json"), "w") as f:
        json.dump({"source": "oracle/kmers_py.py (jellyfish 1.1.11 semantics restated; parity unpinned)",
                   "cases": kcases}, f, separators=(",", ":"))
    print("kmers_golden.json: %d regions" % len(kcases))

    # ---- extra cross-check, not stored ------------------------------------------
    if args.fuzz:
        olc.nw = nw_py.nw_fast
        rng = random.Random(99)
        bad = 0
        for t in range(args.fuzz):
            ev = rng.choice([("del", rng.randint(20, 800), None), ("ins", rng.randint(10, 60)),
                             ("inv", rng.randint(80, 600)), ("tdup", rng.randint(60, 400)), ("trl",), ("none",)])
            kw = dict(seed=100000 + t, L=rng.randint(400, 6000), cov=rng.choice([40, 100, 200, 500]),
                      k=rng.choice([15, 21, 11, 25]), e=rng.choice([0, 0.002, 0.01, 0.03]), event=ev,
                      vaf=rng.choice([1.0, 0.5, 0.25]), indel_p=rng.choice([0, 0.3, 1.0]),
                      rl=rng.choice([100, 75, 150]), rl_jitter=rng.choice([0, 0, 10, 30]),
                      spurious_frac=rng.choice([0, 0, 0.01, 0.03]))
            region = synth.make_region("f%d" % t, **kw)
            _r, _c, _s, only = oracle_sample_only(region)
            a = reference_assembly(asm_mod, region, only)
            b = assembler_py.init_assembly(only, region.reads, region.k, region.rc_thresh, region.read_len)
            if a != b:
                bad += 1
                print("MISMATCH", kw)
        print("fuzz: %d regions, %d mismatches" % (args.fuzz, bad))
        if bad:
            sys.exit(1)


if __name__ == "__main__":
    main()
