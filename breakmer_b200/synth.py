"""Synthetic target regions at the post-extraction boundary of the reference.

BreaKmer's `extract_bam_reads` / `clean_reads` need pysam, samtools and
cutadapt, none of which exist in this environment, so
benchmark and test inputs are generated directly in the formats those two
steps leave behind for `compare_kmers`:

  * the forward reference string of [start-200, end+200)  (utils.py:367)
  * cleaned FASTQ records  @inst:lane:tile:x:y/<1|2>_<0|1>  (utils.py:442,
    utils.py:704-712 requires exactly five ':' fields and a '/')
  * soft-clip FASTA records  seq[0:s+k]  for a leading clip and
    seq[e-k:len] for a trailing clip  (sv_processor.py:499,502)

Everything is driven by `random.Random(seed)`, so a (config, seed) pair names
one exact input on any machine.
"""
import random
from dataclasses import dataclass, field

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
_OTHER = {"A": "CGT", "C": "AGT", "G": "ACT", "T": "ACG", "N": "ACGT"}


def revcomp(s):
    return "".join(_COMP[c] for c in reversed(s))


def _rand_seq(rng, n):
    return "".join(rng.choices("ACGT", k=n))


@dataclass
class Region:
    """One target region, as `target.compare_kmers` sees it."""
    name: str
    k: int
    ref_fwd: str                      # [start-200, end+200) forward strand
    reads: list                       # [(id, seq, qual, indel_only)], cleaned FASTQ order
    sc_records: list                  # [(name, seq)] soft-clip / unmapped FASTA
    normal_reads: list = field(default_factory=list)   # [(id, seq)] normal sample (config 3)
    rc_thresh: int = 2                # params.get_sr_thresh('min') default (utils.py:665-674)
    event: str = "none"

    @property
    def read_len(self):               # utils.py:236 (max over kept records)
        return max((len(r[1]) for r in self.reads), default=0)

    def input_bases(self):
        n = len(self.ref_fwd)
        n += sum(len(r[1]) for r in self.reads)
        n += sum(len(r[1]) for r in self.sc_records)
        n += sum(len(r[1]) for r in self.normal_reads)
        return n


def _poisson(rng, lam):
    # Knuth; lam is small (<= ~25) everywhere this is used
    if lam <= 0:
        return 0
    import math
    limit = math.exp(-lam)
    n, p = 0, rng.random()
    while p > limit:
        n += 1
        p *= rng.random()
    return n


def _mutate(rng, s, e, n_rate):
    if e <= 0 and n_rate <= 0:
        return s
    out = list(s)
    for i, c in enumerate(out):
        r = rng.random()
        if r < n_rate:
            out[i] = "N"
        elif r < n_rate + e:
            out[i] = rng.choice(_OTHER[c])
    return "".join(out)


def _plant(rng, w, event):
    """Return (sample_genome, [junction positions in sample coords])."""
    n = len(w)
    lo, hi = 200 + 20, n - 200 - 20           # keep the event inside the target proper
    if hi - lo < 120:
        lo, hi = n // 4, 3 * n // 4
    kind = event[0]
    if kind == "none":
        return w, []
    if kind == "del":
        d = event[1]
        d = min(d, max(10, (hi - lo) // 2))
        p = rng.randint(lo, hi - d) if event[2] is None else event[2]
        return w[:p] + w[p + d:], [p]
    if kind == "ins":
        d = event[1]
        p = rng.randint(lo, hi)
        return w[:p] + _rand_seq(rng, d) + w[p:], [p, p + d]
    if kind == "inv":
        d = min(event[1], max(60, (hi - lo) // 2))
        p = rng.randint(lo, hi - d)
        return w[:p] + revcomp(w[p:p + d]) + w[p + d:], [p, p + d]
    if kind == "tdup":
        d = min(event[1], max(60, (hi - lo) // 2))
        p = rng.randint(lo, hi - d)
        return w[:p + d] + w[p:p + d] + w[p + d:], [p + d]
    if kind == "trl":
        p = rng.randint(lo, hi)
        return w[:p] + _rand_seq(rng, 300), [p]
    raise ValueError("unknown event %r" % (event,))


def _junction_reads(rng, g, junctions, cov, rl, e, n_rate, k, indel_p, tag, min_side=5, rl_jitter=0):
    """Reads of length rl from genome g that span a junction with >= min_side
    bases on the minor side; returns [(id, seq, qual, indel_only, sc_seq)]."""
    out = []
    lam = cov / float(rl)
    serial = 0
    for jn, J in enumerate(junctions):
        for left in range(min_side, rl - min_side + 1):
            a = J - left
            if a < 0 or a + rl > len(g):
                continue
            for _ in range(_poisson(rng, lam)):
                this_rl = rl - (rng.randint(0, rl_jitter) if rl_jitter else 0)
                seq = _mutate(rng, g[a:a + this_rl], e, n_rate)
                ll = len(seq)
                if left >= ll:
                    continue
                if left >= ll - left:            # trailing soft clip: seq[e-k:len]
                    ee = left
                    sc = seq[(ee - k):ll]        # python slice, as sv_processor.py:502
                else:                            # leading soft clip: seq[0:s+k]
                    sc = seq[0:left + k]
                io = rng.random() < indel_p
                serial += 1
                rid = "@SYN%s:%d:%d:%d:%d/%d_%d" % (tag, 1, jn + 1, a, serial, 1 + (serial & 1), int(io))
                out.append((rid, seq, "I" * ll, io, sc))
    return out


def _spurious_reads(rng, w, n_reads, rl, e, n_rate, k, tag):
    out = []
    for s in range(n_reads):
        a = rng.randint(0, len(w) - rl)
        seq = _mutate(rng, w[a:a + rl], e, n_rate)
        c = rng.randint(5, 40)
        if rng.random() < 0.5:                   # leading clip of c random bases
            seq = _rand_seq(rng, c) + seq[c:]
            sc = seq[0:c + k]
        else:
            seq = seq[:rl - c] + _rand_seq(rng, c)
            ee = rl - c
            sc = seq[(ee - k):rl]
        rid = "@SPU%s:%d:%d:%d:%d/%d_%d" % (tag, 2, 1, a, s + 1, 1 + (s & 1), 0)
        out.append((rid, seq, "I" * len(seq), False, sc))
    return out


def make_region(name, seed, L, cov, k, e, event, vaf=1.0, rl=100, n_rate=0.001,
                indel_p=0.0, spurious_frac=0.0, normal_cov=0.0, germline=False, rl_jitter=0):
    """Build one Region.  `event` is a tuple, e.g. ("del", 1500, None)."""
    rng = random.Random(seed)
    w = _rand_seq(rng, L + 400)
    germ_g = w
    if germline:
        # a germline 12 bp deletion shared by tumour and normal (config 3)
        q = rng.randint(230, max(231, len(w) - 260))
        germ_g = w[:q] + w[q + 12:]
        germ_j = [q]
    g, junctions = _plant(rng, germ_g, event)
    recs = _junction_reads(rng, g, junctions, cov * vaf, rl, e, n_rate, k, indel_p, "T", rl_jitter=rl_jitter)
    if germline:
        # tumour reads over the germline junction (those are clipped too); the
        # planted event may have shifted/removed it, so re-locate by content
        recs += _junction_reads(rng, germ_g, germ_j, cov * vaf, rl, e, n_rate, k, indel_p, "G")
    if spurious_frac > 0:
        n_sp = int(round(spurious_frac * cov * len(w) / float(rl)))
        recs += _spurious_reads(rng, w, n_sp, rl, e, n_rate, k, "S")
    rng.shuffle(recs)
    reads = [(r[0], r[1], r[2], r[3]) for r in recs]
    sc_records = [(r[0].lstrip("@").rsplit("_", 1)[0], r[4]) for r in recs]
    normal_reads = []
    if normal_cov > 0 and germline:
        nrecs = _junction_reads(rng, germ_g, germ_j, normal_cov, rl, e, n_rate, k, 0.0, "N")
        rng.shuffle(nrecs)
        normal_reads = [(r[0], r[1]) for r in nrecs]
    return Region(name=name, k=k, ref_fwd=w, reads=reads, sc_records=sc_records,
                  normal_reads=normal_reads, event=event[0])


# ---------------------------------------------------------------------------
# The five BASELINE.json configurations
# ---------------------------------------------------------------------------
_C2_EVENTS = [("ins", 40), ("del", None), ("inv", None), ("tdup", None)]


def config_regions(cfg, n=None, start=0):
    """Yield the regions of configuration `cfg` ("C1".."C5").  `n` limits the
    count (default: the configuration's full size); `start` offsets the index,
    so region i of a config is the same object however the config is sliced."""
    full = {"C1": 1, "C2": 500, "C3": 500, "C4": 100, "C5": 20000}[cfg]
    n = full if n is None else n
    for i in range(start, start + n):
        yield config_region(cfg, i)


def config_region(cfg, i):
    if cfg == "C1":
        # 1 region, L=20,000, RL=100, 200x, 1.5 kb deletion at centre, k=15, e=0.002, seed 1
        L = 20000
        return make_region("C1_%05d" % i, 1 + i, L, 200, 15, 0.002,
                           ("del", 1500, (L + 400) // 2 - 750), indel_p=0.3)
    if cfg in ("C2", "C3"):
        prng = random.Random(777000 + i)
        L = prng.randint(2000, 20000)
        ev = _C2_EVENTS[i % 4]
        if ev[0] == "ins":
            event = ("ins", 40)
        elif ev[0] == "del":
            event = ("del", prng.randint(30, 1500), None)
        else:
            event = (ev[0], prng.randint(100, 1200))
        return make_region("%s_%05d" % (cfg, i), 1000 + i, L, 200, 15, 0.01, event, vaf=0.5,
                           indel_p=0.3 if ev[0] in ("ins", "del") else 0.0,
                           normal_cov=100.0 if cfg == "C3" else 0.0, germline=(cfg == "C3"))
    if cfg == "C4":
        prng = random.Random(778000 + i)
        L = prng.randint(300, 500)
        return make_region("C4_%05d" % i, 4000 + i, L, 2000, 21, 0.005,
                           ("del", prng.randint(30, 120), None), indel_p=0.3)
    if cfg == "C5":
        prng = random.Random(779000 + i)
        L = prng.randint(500, 5000)
        event = ("trl",) if prng.random() < 0.02 else ("none",)
        return make_region("C5_%05d" % i, 50000 + i, L, 100, 15, 0.005, event,
                           spurious_frac=0.01)
    raise ValueError(cfg)
