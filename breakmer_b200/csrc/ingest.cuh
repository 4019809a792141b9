// Ingest: FASTA / FASTQ text of a batch of target regions
// -> the packed arrays of bk_batch_input, parsed on host threads straight into page-locked
// memory, so the reference's file round trips on the way into compare_kmers
// (FastqFile utils.py:692-720, get_fastq_reads utils.py:203-246, the readers inside
// `jellyfish count` utils.py:160) cost one pass over the text.  Host code only: none of the
// hot path's arithmetic happens here, and nothing here is a fallback for a kernel.
//
// Text rules (restated from the reference's readers, checked against the Python drop-ins
// breakmer_b200.utils.FastqFile / read_sequences in tests/test_ingest.py):
//   * lines end at "\n" only (CPython 2.7 text mode on Linux does no newline translation); a "\r" before it is
//     white space and goes with the strip;
//   * reads file = strict 4-line FASTQ records; header, sequence and quality are stripped of
//     surrounding white space; a trailing group of fewer than four lines is dropped
//     silently (the StopIteration inside FastqFile.next ends the iteration);
//     the header must split on ':' into exactly five fields, the fifth must hold exactly
//     one '/' and at most one '#', and lane / tile / x / y must be integers (utils.py:704-719
//     raises otherwise) -> BK_ERR_FORMAT;
//   * k-mer inputs (reference window, soft-clip FASTA, normal sample) are sniffed on the
//     first byte like jellyfish does: '@' = FASTQ (the sequence is the second line of every
//     group of four), anything else = FASTA ('>' starts a record, the following lines are
//     joined after stripping; text before the first '>' is ignored);
//   * of the reference window file only the first record is used (one window per target).
#pragma once
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <exception>
#include <mutex>
#include <thread>
#include <vector>

#include "host_util.cuh"

namespace bk {

struct TextView { const char* p; size_t n; };

struct LineReader {
  const char* p; const char* end;
  explicit LineReader(TextView t) : p(t.p), end(t.p + t.n) {}
  // next line without its terminator; false at end of text
  bool next(const char*& a, const char*& b) {
    if (p >= end) return false;
    a = p;
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    b = nl ? nl : end;
    p = nl ? nl + 1 : end;
    return true;
  }
};

// str.strip() of CPython 2.7: space, \t \n \v \f \r
inline bool py_space(unsigned char c) { return c == ' ' || (c >= 0x09 && c <= 0x0d); }
inline void strip(const char*& a, const char*& b) {
  while (a < b && py_space((unsigned char)*a)) ++a;
  while (b > a && py_space((unsigned char)b[-1])) --b;
}
// Python int(): optional white space, optional sign, one or more digits
inline bool py_int(const char* a, const char* b) {
  strip(a, b);
  if (a < b && (*a == '+' || *a == '-')) ++a;
  if (a >= b) return false;
  for (; a < b; ++a) if (*a < '0' || *a > '9') return false;
  return true;
}

struct ParsedSet {               // the records of one input of one region
  std::string bases;
  std::vector<int32_t> len;
  void add(const char* a, const char* b) { bases.append(a, b); len.push_back((int32_t)(b - a)); }
};
// The reads of a region are kept as slices of the source text (which outlives the layout step) and copied ONCE,
// straight into the batch buffer.
struct ReadSlice { const char* id; const char* seq; const char* qual; int32_t id_len, seq_len, qual_len; uint8_t flag; };
struct ParsedReads {
  std::vector<ReadSlice> recs;
  int64_t id_bytes = 0, seq_bytes = 0, qual_bytes = 0;
  int32_t max_len = 0;
};

// utils.py:704-719
inline const char* check_fastq_header(const char* a, const char* b) {
  const char* f[6];
  int nf = 0;
  f[nf++] = a;
  for (const char* q = a; q < b; ++q)
    if (*q == ':') {
      if (nf == 5) return "header does not split into five ':'-separated fields";
      f[nf++] = q + 1;
    }
  if (nf != 5) return "header does not split into five ':'-separated fields";
  f[5] = b + 1;
  const char* y0 = f[4];
  const char* y1 = b;
  int slashes = 0, hashes = 0;
  const char* slash = nullptr;
  for (const char* q = y0; q < y1; ++q) if (*q == '/') { ++slashes; if (!slash) slash = q; }
  if (slashes != 1) return "fifth header field needs exactly one '/'";
  y1 = slash;
  const char* hash = nullptr;
  for (const char* q = y0; q < y1; ++q) if (*q == '#') { ++hashes; if (!hash) hash = q; }
  if (hashes > 1) return "fifth header field holds more than one '#'";
  if (hash) y1 = hash;
  if (!py_int(f[1], f[2] - 1) || !py_int(f[2], f[3] - 1) || !py_int(f[3], f[4] - 1) || !py_int(y0, y1))
    return "lane, tile, x and y of the header must be integers";
  return nullptr;
}

// FastqFile (utils.py:692-720) + the record model of get_fastq_reads (utils.py:230-244)
inline void parse_reads_fastq(TextView t, ParsedReads& out, std::string& err) {
  LineReader lr(t);
  int64_t rec = 0;
  out.recs.reserve(t.n / 200 + 4);
  for (;;) {
    const char *h0, *h1, *s0, *s1, *p0, *p1, *q0, *q1;
    if (!lr.next(h0, h1) || !lr.next(s0, s1) || !lr.next(p0, p1) || !lr.next(q0, q1)) break;
    ++rec;
    strip(h0, h1); strip(s0, s1); strip(q0, q1);
    if (const char* why = check_fastq_header(h0, h1)) {
      char buf[256];
      snprintf(buf, sizeof buf, "FASTQ record %lld: %s (utils.py:704-719)", (long long)rec, why);
      err = buf;
      return;
    }
    // the extraction step writes "@<qname>/<end>_<0|1>", 1 = indel_only (fq_line, utils.py:436-443)
    const char* u = h1;
    while (u > h0 && u[-1] != '_') --u;
    const size_t sl = (size_t)(h1 - u);
    const bool flag = u > h0 && ((sl == 4 && memcmp(u, "True", 4) == 0) || (sl == 1 && *u == '1'));
    out.recs.push_back(ReadSlice{h0, s0, q0, (int32_t)(h1 - h0), (int32_t)(s1 - s0), (int32_t)(q1 - q0), (uint8_t)(flag ? 1 : 0)});
    out.id_bytes += h1 - h0; out.seq_bytes += s1 - s0; out.qual_bytes += q1 - q0;
    if ((int32_t)(s1 - s0) > out.max_len) out.max_len = (int32_t)(s1 - s0);
  }
}

// record sequences of a k-mer input, format sniffed on the first byte
inline void parse_sequences(TextView t, ParsedSet& out, bool first_only) {
  if (t.n == 0) return;
  LineReader lr(t);
  const char *a, *b;
  if (t.p[0] == '@') {
    int64_t i = 0;
    while (lr.next(a, b)) {
      if ((i & 3) == 1) {
        strip(a, b);
        out.add(a, b);
        if (first_only) return;
      }
      ++i;
    }
    return;
  }
  bool open = false;
  std::string cur;
  while (lr.next(a, b)) {
    strip(a, b);
    if (a < b && *a == '>') {
      if (open) {
        out.add(cur.data(), cur.data() + cur.size());
        if (first_only) return;
      }
      open = true;
      cur.clear();
    } else if (open) {
      cur.append(a, b);
    }
  }
  if (open) out.add(cur.data(), cur.data() + cur.size());
}

struct IngestRegion {
  ParsedSet ref, sc, normal;
  ParsedReads reads;
  std::string err;
};

// one host buffer (page-locked when the ingest object was created with pinned = 1), grown on demand
struct HostBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool pinned = false;
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    release();
    size_t want = bytes + bytes / 4 + 4096;
    if (pinned) BK_CUDA(cudaHostAlloc(&p, want, cudaHostAllocDefault));
    else { p = malloc(want); if (!p) throw std::bad_alloc(); }
    cap = want;
  }
  void release() {
    if (p) { if (pinned) cudaFreeHost(p); else free(p); }
    p = nullptr; cap = 0;
  }
  ~HostBuf() { release(); }
};

struct Ingest {
  int n_threads = 1;
  bool pinned = false;
  std::string err;
  HostBuf buf;
  std::vector<IngestRegion> regions;
  std::vector<std::string> file_text;      // bk_ingest_files: the text of every file, 4 per region

  template <typename F>
  void parallel_for(int n, F&& f) {
    const int nt = std::max(1, std::min(n_threads, n));
    if (nt == 1) { for (int i = 0; i < n; ++i) f(i); return; }
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    th.reserve(nt);
    // an exception inside a worker (bad_alloc from a growing string, fail()) must not escape its thread -- that would
    // terminate the host process: the first one is kept and rethrown on the calling thread after the join
    std::exception_ptr first;
    std::mutex first_mu;
    for (int t = 0; t < nt; ++t)
      th.emplace_back([&] {
        try {
          for (int i; (i = next.fetch_add(1)) < n;) f(i);
        } catch (...) {
          std::lock_guard<std::mutex> lk(first_mu);
          if (!first) first = std::current_exception();
          next.store(n);                       // stop handing out work
        }
      });
    for (auto& x : th) x.join();
    if (first) std::rethrow_exception(first);
  }
};

struct IngestText {              // what bk_ingest_* hands back besides the bk_batch_input arrays
  const char* id_bytes; const int64_t* id_off;
  const char* qual_bytes; const int64_t* qual_off;
  int64_t n_reads;
  uint8_t* read_flags;
};

inline bool read_whole_file(const char* path, std::string& out) {
  // plain POSIX calls: a batch reads four small files per target, so the per-file call count matters
  const int fd = open(path, O_RDONLY | O_CLOEXEC);
  if (fd < 0) return false;
  out.clear();
  struct stat st;
  memset(&st, 0, sizeof st);
  size_t have = 0;
  const bool is_reg = fstat(fd, &st) == 0 && S_ISREG(st.st_mode);
  if (is_reg && st.st_size > 0) out.resize((size_t)st.st_size);
  else out.resize(1 << 16);
  bool ok = true;
  for (;;) {
    if (have == out.size()) out.resize(out.size() + (out.size() >> 1) + (1 << 16));      // the file grew, or it is a pipe
    const ssize_t got = read(fd, &out[have], out.size() - have);
    if (got < 0) { if (errno == EINTR) continue; ok = false; break; }
    if (got == 0) break;
    have += (size_t)got;
    if (is_reg && have == (size_t)st.st_size) break;                        // the common case: one read
  }
  close(fd);
  out.resize(have);
  return ok;
}

// Parses the 4 x n texts (null pointer = absent input) and lays the batch out in g.buf.
template <typename BatchInput>
void ingest_texts(Ingest& g, int n, const TextView* ref, const TextView* reads, const TextView* sc, const TextView* normal,
                  BatchInput* in, IngestText* text) {
  g.regions.assign((size_t)n, IngestRegion());
  g.parallel_for(n, [&](int r) {
    IngestRegion& R = g.regions[r];
    if (ref && ref[r].p) parse_sequences(ref[r], R.ref, true);
    if (reads && reads[r].p) parse_reads_fastq(reads[r], R.reads, R.err);
    if (sc && sc[r].p) parse_sequences(sc[r], R.sc, false);
    if (normal && normal[r].p) parse_sequences(normal[r], R.normal, false);
  });
  for (int r = 0; r < n; ++r)
    if (!g.regions[r].err.empty()) fail(BK_ERR_FORMAT, "region %d: %s", r, g.regions[r].err.c_str());

  // ---- layout: [int64 tables][int32 read_len][bases ...][ids][quals][flags], every array 64-byte aligned
  struct Tot { int64_t rec = 0, bytes = 0; };
  std::vector<Tot> t_ref(n + 1), t_rd(n + 1), t_sc(n + 1), t_nm(n + 1);
  std::vector<int64_t> t_id(n + 1, 0), t_q(n + 1, 0);
  for (int r = 0; r < n; ++r) {
    const IngestRegion& R = g.regions[r];
    t_ref[r + 1] = {t_ref[r].rec + 1, t_ref[r].bytes + (R.ref.len.empty() ? 0 : R.ref.len[0])};
    t_rd[r + 1] = {t_rd[r].rec + (int64_t)R.reads.recs.size(), t_rd[r].bytes + R.reads.seq_bytes};
    t_sc[r + 1] = {t_sc[r].rec + (int64_t)R.sc.len.size(), t_sc[r].bytes + (int64_t)R.sc.bases.size()};
    t_nm[r + 1] = {t_nm[r].rec + (int64_t)R.normal.len.size(), t_nm[r].bytes + (int64_t)R.normal.bases.size()};
    t_id[r + 1] = t_id[r] + R.reads.id_bytes;
    t_q[r + 1] = t_q[r] + R.reads.qual_bytes;
  }
  const int64_t n_rd = t_rd[n].rec, n_sc = t_sc[n].rec, n_nm = t_nm[n].rec;
  size_t total = 0;
  auto place = [&](size_t bytes) { size_t o = total; total += (bytes + 63) & ~size_t(63); return o; };
  const size_t o_ref_off = place((n + 1) * 8), o_rd_off = place((n_rd + 1) * 8), o_rd_reg = place((n + 1) * 8);
  const size_t o_sc_off = place((n_sc + 1) * 8), o_sc_reg = place((n + 1) * 8);
  const size_t o_nm_off = place((n_nm + 1) * 8), o_nm_reg = place((n + 1) * 8);
  const size_t o_id_off = place((n_rd + 1) * 8), o_q_off = place((n_rd + 1) * 8);
  const size_t o_rlen = place((size_t)(n + 1) * 4);
  const size_t o_ref = place(t_ref[n].bytes + 1), o_rd = place(t_rd[n].bytes + 1), o_sc = place(t_sc[n].bytes + 1);
  const size_t o_nm = place(t_nm[n].bytes + 1), o_id = place(t_id[n] + 1), o_q = place(t_q[n] + 1), o_fl = place(n_rd + 1);
  g.buf.pinned = g.pinned;
  g.buf.reserve(total);
  char* base = (char*)g.buf.p;
  int64_t* ref_off = (int64_t*)(base + o_ref_off);
  int64_t* rd_off = (int64_t*)(base + o_rd_off);   int64_t* rd_reg = (int64_t*)(base + o_rd_reg);
  int64_t* sc_off = (int64_t*)(base + o_sc_off);   int64_t* sc_reg = (int64_t*)(base + o_sc_reg);
  int64_t* nm_off = (int64_t*)(base + o_nm_off);   int64_t* nm_reg = (int64_t*)(base + o_nm_reg);
  int64_t* id_off = (int64_t*)(base + o_id_off);   int64_t* q_off = (int64_t*)(base + o_q_off);
  int32_t* rlen = (int32_t*)(base + o_rlen);
  uint8_t* fl = (uint8_t*)(base + o_fl);
  ref_off[n] = t_ref[n].bytes; rd_off[n_rd] = t_rd[n].bytes; sc_off[n_sc] = t_sc[n].bytes; nm_off[n_nm] = t_nm[n].bytes;
  rd_reg[n] = n_rd; sc_reg[n] = n_sc; nm_reg[n] = n_nm; id_off[n_rd] = t_id[n]; q_off[n_rd] = t_q[n];
  rlen[n] = 0; fl[n_rd] = 0;
  g.parallel_for(n, [&](int r) {
    const IngestRegion& R = g.regions[r];
    ref_off[r] = t_ref[r].bytes;
    if (!R.ref.len.empty()) memcpy(base + o_ref + t_ref[r].bytes, R.ref.bases.data(), (size_t)R.ref.len[0]);
    auto put = [&](const ParsedSet& S, const Tot& t0, int64_t* off, int64_t* reg, size_t o_bytes) {
      reg[r] = t0.rec;
      int64_t o = t0.bytes;
      for (size_t i = 0; i < S.len.size(); ++i) { off[t0.rec + (int64_t)i] = o; o += S.len[i]; }
      if (!S.bases.empty()) memcpy(base + o_bytes + t0.bytes, S.bases.data(), S.bases.size());
    };
    put(R.sc, t_sc[r], sc_off, sc_reg, o_sc);
    put(R.normal, t_nm[r], nm_off, nm_reg, o_nm);
    rd_reg[r] = t_rd[r].rec;
    int64_t os = t_rd[r].bytes, oi = t_id[r], oq = t_q[r];
    char* d_seq = base + o_rd;
    char* d_id = base + o_id;
    char* d_q = base + o_q;
    for (size_t i = 0; i < R.reads.recs.size(); ++i) {
      const ReadSlice& x = R.reads.recs[i];
      const int64_t gi = t_rd[r].rec + (int64_t)i;
      rd_off[gi] = os; id_off[gi] = oi; q_off[gi] = oq; fl[gi] = x.flag;
      memcpy(d_seq + os, x.seq, (size_t)x.seq_len);   os += x.seq_len;
      memcpy(d_id + oi, x.id, (size_t)x.id_len);      oi += x.id_len;
      memcpy(d_q + oq, x.qual, (size_t)x.qual_len);   oq += x.qual_len;
    }
    rlen[r] = R.reads.max_len;
  });
  memset(in, 0, sizeof *in);
  in->n_regions = n;
  in->ref_bases = base + o_ref; in->ref_off = ref_off;
  in->read_bases = base + o_rd; in->read_off = rd_off; in->read_reg_off = rd_reg; in->read_flags = fl;
  in->sc_bases = base + o_sc; in->sc_off = sc_off; in->sc_reg_off = sc_reg;
  if (normal) { in->normal_bases = base + o_nm; in->normal_off = nm_off; in->normal_reg_off = nm_reg; }
  in->read_len = rlen;
  if (text) {
    text->id_bytes = base + o_id; text->id_off = id_off;
    text->qual_bytes = base + o_q; text->qual_off = q_off;
    text->n_reads = n_rd; text->read_flags = fl;
  }
  g.regions.clear();
}

}  // namespace bk

// ---- contig hand-off -------------------------------------------
// The files sv_processor.contig.setup writes for every contig before blat (sv_processor.py:749-782),
// for all contigs of a batch result at once, on host threads:
//   <contigs_dir>/<id>/<id>.fq   reads of the contig: id, seq, "+", qual             (write_read_fq, :767-772)
//   <contigs_dir>/<id>/<id>.fa   ">contig1\n" + sequence, no trailing newline         (write_contig_fa, :776-781)
//   <cluster_fn>                 "<id> <n kmers>\n<mers,>\n<read ids,>\n\n"; opened with 'w' for every contig
//                                (:758-763), so what survives is the LAST contig of the target -- only that is written
// id = "contig<n>", n = 1, 2, ... in acceptance order (resolve_sv, sv_processor.py:649-653).
// Order policy: the reference iterates a Python set of reads (arbitrary order); here reads are written in the
// order of ctg_reads.
#include <sys/stat.h>
#include <sys/types.h>
#include <errno.h>

namespace bk {

inline bool make_dirs(const std::string& path) {
  if (mkdir(path.c_str(), 0777) == 0 || errno == EEXIST) {           // common case: the parent exists already
    struct stat st0;
    return stat(path.c_str(), &st0) == 0 && S_ISDIR(st0.st_mode);
  }
  std::string cur;
  size_t i = 0;
  while (i <= path.size()) {
    const size_t j = path.find('/', i);
    const size_t e = j == std::string::npos ? path.size() : j;
    cur = path.substr(0, e);
    if (!cur.empty() && mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST) return false;
    if (j == std::string::npos) break;
    i = j + 1;
  }
  struct stat st;
  return stat(path.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
}

inline bool write_whole_file(const std::string& path, const std::string& data) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const bool ok = data.empty() || fwrite(data.data(), 1, data.size(), f) == data.size();
  return (fclose(f) == 0) && ok;
}

inline void append_mer(std::string& s, uint64_t code, int k) {
  for (int t = k - 1; t >= 0; --t) s.push_back("ACGT"[(code >> (2 * t)) & 3u]);
}

// "<mer>\t<case count>\n" per sample-only k-mer of a target: the file target.compare_kmers writes
// (sv_processor.py:625-632).  The reference walks a Python set (arbitrary order); here ascending mer order.
template <typename Result>
int64_t write_sample_kmer_files(Ingest& g, const Result* res, const char* const* paths, int k) {
  const int R = res->n_regions;
  std::atomic<int64_t> n_files{0};
  std::atomic<int> failed{-1};
  g.parallel_for(R, [&](int r) {
    if (!paths[r] || !paths[r][0]) return;
    const int64_t a = res->so_off[r], b = res->so_off[r + 1];
    std::string txt;
    txt.reserve((size_t)(b - a) * (size_t)(k + 8));
    char num[16];
    for (int64_t i = a; i < b; ++i) {
      append_mer(txt, res->so_mers[i], k);
      txt.push_back('\t');
      const int nn = snprintf(num, sizeof num, "%u", res->so_counts[i]);
      txt.append(num, (size_t)nn);
      txt.push_back('\n');
    }
    if (!write_whole_file(paths[r], txt)) { failed = r; return; }
    n_files += 1;
  });
  if (failed >= 0) fail(BK_ERR_IO, "cannot write %s", paths[failed]);
  return n_files.load();
}

template <typename Result, typename BatchInput>
int64_t write_contig_files(Ingest& g, const Result* res, const BatchInput* in, const IngestText* text,
                           const char* const* contigs_dir, const char* const* cluster_fn, int k) {
  const int R = res->n_regions;
  struct Job { int region; int64_t contig; int ordinal; bool last; };
  std::vector<Job> jobs;
  for (int r = 0; r < R; ++r) {
    if (!contigs_dir[r] || !contigs_dir[r][0]) continue;
    const int64_t a = res->ctg_reg_off[r], b = res->ctg_reg_off[r + 1];
    for (int64_t c = a; c < b; ++c) jobs.push_back({r, c, (int)(c - a) + 1, c + 1 == b});
  }
  std::atomic<int64_t> n_files{0};
  std::atomic<int> failed{-1};
  g.parallel_for((int)jobs.size(), [&](int ji) {
    const Job& J = jobs[ji];
    const std::string id = "contig" + std::to_string(J.ordinal);
    const std::string dir = std::string(contigs_dir[J.region]) + "/" + id;
    if (!make_dirs(dir)) { failed = ji; return; }
    const int64_t so = res->ctg_seq_off[2 * J.contig], sl = res->ctg_seq_off[2 * J.contig + 1];
    const int64_t ro = res->ctg_reads_off[2 * J.contig], nr = res->ctg_reads_off[2 * J.contig + 1];
    const int64_t ko = res->ctg_kmers_off[2 * J.contig], nk = res->ctg_kmers_off[2 * J.contig + 1];
    std::string fq, ids;
    for (int64_t i = 0; i < nr; ++i) {
      const int64_t rec = res->ctg_reads[ro + i];
      const char* idp = text->id_bytes + text->id_off[rec];
      const size_t idn = (size_t)(text->id_off[rec + 1] - text->id_off[rec]);
      fq.append(idp, idn); fq.push_back('\n');
      fq.append(in->read_bases + in->read_off[rec], (size_t)(in->read_off[rec + 1] - in->read_off[rec]));
      fq.append("\n+\n");
      fq.append(text->qual_bytes + text->qual_off[rec], (size_t)(text->qual_off[rec + 1] - text->qual_off[rec]));
      fq.push_back('\n');
      if (J.last && cluster_fn && cluster_fn[J.region]) { if (i) ids.push_back(','); ids.append(idp, idn); }
    }
    std::string fa = ">contig1\n";
    fa.append(res->ctg_seq + so, (size_t)sl);
    if (!write_whole_file(dir + "/" + id + ".fq", fq) || !write_whole_file(dir + "/" + id + ".fa", fa)) { failed = ji; return; }
    n_files += 2;
    if (J.last && cluster_fn && cluster_fn[J.region] && cluster_fn[J.region][0]) {
      std::string cl = id + " " + std::to_string(nk) + "\n";
      for (int64_t e = 0; e < nk; ++e) { if (e) cl.push_back(','); append_mer(cl, res->ctg_kmer_mer[ko + e], k); }
      cl.push_back('\n');
      cl += ids;
      cl += "\n\n";
      if (!write_whole_file(cluster_fn[J.region], cl)) { failed = ji; return; }
      n_files += 1;
    }
  });
  if (failed >= 0) fail(BK_ERR_IO, "cannot write the files of contig %d of region %d under %s", jobs[failed].ordinal,
                        jobs[failed].region, contigs_dir[jobs[failed].region]);
  return n_files.load();
}

}  // namespace bk
