"""CPU restatement of the readers on the way into compare_kmers (ingest row,
SURVEY.md section 8.7 f.1).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the
product path is bk_ingest_* in breakmer_b200/csrc/ingest.cuh.

  fastq_records(text)     FastqFile            utils.py:692-720
  fq_recs(text)           the record model of  get_fastq_reads utils.py:230-244
                          (with every read kept; the sv_reads filter of :213-236 needs the
                          BAM-derived dictionary and stays with the caller)
  kmer_sequences(text)    the record sequences `jellyfish count` sees in a FASTA / FASTQ input
                          (utils.py:160).  jellyfish 1.1.11 is a third-party binary that is absent
                          here: this part restates its documented behaviour (format sniffed on the
                          first byte; '>' header lines, the following lines joined; FASTQ = the
                          second line of every four) and is PARITY UNPINNED; the FastqFile and
                          get_fastq_reads parts are pinned against the reference's own code by
                          oracle/make_golden_ingest.py -> tests/golden/ingest_cases.json.

Text is bytes-as-str with Python-2 semantics: lines end at "\\n" only, strip() removes
space, \\t \\n \\v \\f \\r.
"""
import re
from collections import OrderedDict

_WS = " \t\n\r\x0b\x0c"
_INT = re.compile(r"^[ \t\n\r\x0b\x0c]*[+-]?[0-9]+[ \t\n\r\x0b\x0c]*$")


def _lines(text):
    if not text:
        return []
    parts = text.split("\n")
    if parts[-1] == "":
        parts.pop()
    return parts


def fastq_records(text):
    """[(header, seq, qual)] -- raises ValueError where FastqFile.next raises (utils.py:704-719)."""
    lines = _lines(text)
    out = []
    for i in range(0, len(lines) - 3, 4):            # a trailing group of < 4 lines ends the iteration (:703)
        header = lines[i].strip(_WS)
        seq = lines[i + 1].strip(_WS)
        qual = lines[i + 3].strip(_WS)
        fields = header.split(":")
        if len(fields) != 5:                          # :705 tuple unpacking
            raise ValueError("record %d: header does not split into five fields" % (len(out) + 1))
        inst, lane, tile, x, y = fields
        if y.count("/") != 1:                         # :711-712 (no '/': `end` is unbound at :719)
            raise ValueError("record %d: fifth field needs exactly one '/'" % (len(out) + 1))
        y = y.split("/")[0]
        if y.count("#") > 1:                          # :713-714
            raise ValueError("record %d: more than one '#'" % (len(out) + 1))
        y = y.split("#")[0]
        for v in (lane, tile, x, y):                  # :716-719 int()
            if not _INT.match(v):
                raise ValueError("record %d: non-integer lane/tile/x/y" % (len(out) + 1))
        out.append((header, seq, qual))
    return out


def fq_recs(text):
    """(OrderedDict seq -> [header, ...] in file order, read_len) -- utils.py:230-244 with add == True."""
    recs = OrderedDict()
    read_len = 0
    for header, seq, qual in fastq_records(text):
        read_len = max(read_len, len(seq))
        recs.setdefault(seq, []).append(header)
    return recs, read_len


def indel_only_suffix(header):
    """The flag fq_line wrote behind the last '_' (utils.py:436-443)."""
    return header.rsplit("_", 1)[-1] in ("1", "True") if "_" in header else False


def kmer_sequences(text, first_only=False):
    out = []
    if not text:
        return out
    lines = _lines(text)
    if text[0] == "@":
        out = [lines[i].strip(_WS) for i in range(1, len(lines), 4)]
    else:
        cur = None
        for line in lines:
            line = line.strip(_WS)
            if line.startswith(">"):
                if cur is not None:
                    out.append("".join(cur))
                cur = []
            elif cur is not None:
                cur.append(line)
        if cur is not None:
            out.append("".join(cur))
    return out[:1] if first_only else out
