"""Drop-ins for the hot-path pieces of the reference's utils.py.

  fq_read, FastqFile        utils.py:681-720  (record model, T1)
  run_jellyfish             utils.py:151-179  k-mer counting, on the GPU instead of
                                              `jellyfish count` + `jellyfish dump -c`
  load_kmers                utils.py:287-297  dump file(s) -> {mer: count}
  get_marker_fn             utils.py:146

`run_jellyfish` keeps the reference's signature, file naming and marker-file
cache: it writes "<fa_fn>_<k>mers_dump" ("<MER> <count>" per line, the format
`dump -c` produces) next to the input and touches ".<dump name>".  The
`jellyfish` argument (path of the binary) is accepted and ignored.

These are the drop-in's INTERFACE, so a few small pieces follow the reference's
closely on purpose: `fq_read` has the reference's attributes, `get_marker_fn` and
the head of `run_jellyfish` its file naming and marker cache, and `load_kmers`
is the same ten-line dump reader (a dump line is "<mer> <count>"; counts of a
mer seen in several files add up).  Everything that computes is new.
"""
import logging
import os
import re

from . import _lib, get_handle


class fq_read:
    def __init__(self, header, seq, qual, indel_only):
        self.id = header
        self.seq = str(seq)
        self.qual = str(qual)
        self.used = False
        self.dup = False
        self.indel_only = indel_only


_WS = " \t\n\r\x0b\x0c"                       # what str.strip() removes on CPython 2.7
_INT = re.compile(r"^[ \t\n\r\x0b\x0c]*[+-]?[0-9]+[ \t\n\r\x0b\x0c]*$")


class FastqFile(object):
    """Iterator over (header, seq, qual) with the reference's checks (utils.py:692-720): five
    ':'-separated header fields, exactly one '/' (and at most one '#') in the fifth, integer lane /
    tile / x / y; a trailing group of fewer than four lines ends the iteration.  Lines end at "\\n"
    only, as in CPython 2.7 text mode on Linux."""

    def __init__(self, f):
        if isinstance(f, str):
            f = open(f, newline="\n")
        self._f = f

    def __iter__(self):
        return self

    def __next__(self):
        header, seq, _qh, qual = [next(self._f) for _ in range(4)]
        header = header.strip(_WS)
        inst, lane, tile, x, y_end = header.split(':')
        if y_end.count('/') != 1:
            raise ValueError("FASTQ header %r: the last field needs exactly one '/'" % header)
        y = y_end.split('/')[0]
        if y.count('#') > 1:
            raise ValueError("FASTQ header %r: more than one '#'" % header)
        y = y.split('#')[0]
        for v in (lane, tile, x, y):
            if not _INT.match(v):
                raise ValueError("FASTQ header %r: lane, tile, x and y must be integers" % header)
        return (header, seq.strip(_WS), qual.strip(_WS))

    next = __next__


def get_marker_fn(fn):
    return os.path.join(os.path.split(fn)[0], "." + os.path.basename(fn))


def read_sequences(fn):
    """Record sequences of a FASTA or FASTQ file (format sniffed from the first
    byte, as jellyfish does); multi-line FASTA records are joined.  The batched path parses the
    same way natively (bk_ingest_files, csrc/ingest.cuh)."""
    seqs = []
    with open(fn, newline="\n") as f:
        text = f.read()
    if not text:
        return seqs
    lines = text.split("\n")
    if lines[-1] == "":
        lines.pop()
    if text[0] == "@":
        return [lines[i].strip(_WS) for i in range(1, len(lines), 4)]
    cur = None
    for line in lines:
        line = line.strip(_WS)
        if line.startswith(">"):
            if cur is not None:
                seqs.append("".join(cur))
            cur = []
        elif cur is not None:
            cur.append(line)
    if cur is not None:
        seqs.append("".join(cur))
    return seqs


def run_jellyfish(fa_fn, jellyfish, kmer_size):
    logger = logging.getLogger('root')
    file_path = os.path.split(fa_fn)[0]
    file_base = os.path.basename(fa_fn)
    dump_fn = os.path.join(file_path, file_base + "_" + str(kmer_size) + "mers_dump")
    dump_marker_fn = get_marker_fn(dump_fn)
    if not os.path.isfile(dump_marker_fn):
        logger.info('Counting %d-mers of %s on the GPU' % (kmer_size, fa_fn))
        mers, counts = get_handle().count_kmers(read_sequences(fa_fn), int(kmer_size))
        with open(dump_fn, "w") as out:
            for m, c in zip(_lib.codes_to_mers(mers, int(kmer_size)), counts):
                out.write("%s %d\n" % (m, int(c)))
        open(dump_marker_fn, "a").close()
        logger.info('Completed k-mer dump %s, touching marker file %s' % (dump_fn, dump_marker_fn))
    else:
        logger.info('Kmers already generated for target.')
    return dump_fn


def load_kmers(fns, kmers):
    fns = fns.split(",")
    for fn in fns:
        with open(fn) as f:
            for line in f.readlines():
                line = line.strip()
                mer, count = line.split()
                if mer not in kmers:
                    kmers[mer] = 0
                kmers[mer] += int(count)
    return kmers
