// Host orchestration of the batched whole-path pipeline (bk_compare_kmers_batch,
// bk_batch_upload + bk_compare_kmers_resident).  Included into api.cu.
//
// Stage order on the handle's stream (every stage is a kernel in this directory):
//   1. group identical reads            prep.cuh   read_hash -> radix sort -> leaders -> scan -> scatter
//   2. k-mer stage                      kmers.cuh  emit (ref fwd+rc, reads, soft clips, normal) -> radix sort
//                                                  -> run_select (count + set algebra) -> run_scatter
//   3. seed order, liveness             prep.cuh   mer_prep -> radix sort
//   4. k-mer -> read inverted index     prep.cuh   index_emit -> radix sort -> post_off / post_split
//   5. assembly                         assemble.cuh  one warp per region, dynamic region queue
//   6. results to pinned host memory
// Host work is limited to sizing allocations from a few device counters,
// compacting away empty records, and ordering the contig descriptor table.
#pragma once

#include <numeric>

#include "assemble.cuh"
#include "prep.cuh"

namespace bk {

struct RecordSet {          // one of: reads, soft clips, normal reads (device copies)
  const uint8_t* bases = nullptr;
  int64_t n_bases = 0;
  const int64_t* off = nullptr;       // n_rec + 1 (all records, empties included)
  int64_t n_rec = 0;
  const int32_t* seg = nullptr;       // region of every record
  // compacted view without empty records (k-mer emit needs distinct starts)
  const int64_t* koff = nullptr;
  const int32_t* kseg = nullptr;
  int64_t kn_rec = 0;
  // host copies of the region boundaries (bases, compacted records): the k-mer stage may run in region chunks
  std::vector<int64_t> reg_base, reg_krec;
};

struct Pipeline {
  int n_regions = 0, k = 0, rc_thresh = 0, have_mers = 0;
  bool use_ref_cache = false;
  RecordSet ref, reads, sc, normal;
  const uint8_t* read_flags = nullptr;
  const int64_t* read_reg_off = nullptr;  // device
  const int32_t* read_len = nullptr;      // device, per region
  const uint64_t* in_mers = nullptr;      // have_mers
  const uint32_t* in_counts = nullptr;
  const int64_t* in_mers_off = nullptr;   // device
  int64_t n_in_mers = 0;
  int max_read_len = 0;
  int64_t total_read_bytes = 0;
  int64_t h2d_bytes = 0;
};

}  // namespace bk

namespace {

void upload_record_set(bk_handle_t h, Arena<false>& A, const char* bases, const int64_t* off, const int64_t* reg_off,
                       int n_regions, RecordSet& rs, int64_t& h2d, const char* what) {
  rs = RecordSet();
  if (!off || !reg_off) {
    if (bases) fail(BK_ERR_ARG, "batch: %s offsets missing", what);
    int64_t zero = 0;
    rs.off = to_device(h, A, &zero, 1);
    rs.koff = rs.off;
    return;
  }
  const int64_t n_rec = reg_off[n_regions];
  if (reg_off[0] != 0) fail(BK_ERR_ARG, "batch: %s region offsets must start at 0", what);
  if (off[0] != 0) fail(BK_ERR_ARG, "batch: %s record offsets must start at 0", what);
  std::vector<int32_t> seg(n_rec ? n_rec : 1), kseg;
  std::vector<int64_t> koff;
  koff.reserve(n_rec + 1);
  kseg.reserve(n_rec + 1);
  rs.reg_base.assign(n_regions + 1, 0);
  rs.reg_krec.assign(n_regions + 1, 0);
  for (int r = 0; r < n_regions; ++r) {
    if (reg_off[r + 1] < reg_off[r]) fail(BK_ERR_ARG, "batch: %s region offsets not monotone", what);
    rs.reg_base[r] = off[reg_off[r]];
    rs.reg_krec[r] = (int64_t)koff.size();
    for (int64_t i = reg_off[r]; i < reg_off[r + 1]; ++i) {
      seg[i] = r;
      if (off[i + 1] < off[i]) fail(BK_ERR_ARG, "batch: %s record offsets not monotone", what);
      if (off[i + 1] > off[i]) { koff.push_back(off[i]); kseg.push_back(r); }
    }
  }
  const int64_t n_bases = off[n_rec];
  rs.reg_base[n_regions] = n_bases;
  rs.reg_krec[n_regions] = (int64_t)koff.size();
  koff.push_back(n_bases);
  if (n_bases > 0 && !bases) fail(BK_ERR_ARG, "batch: %s bases missing", what);
  rs.n_bases = n_bases;
  rs.n_rec = n_rec;
  rs.bases = (const uint8_t*)to_device(h, A, bases, (size_t)n_bases);
  rs.off = to_device(h, A, off, (size_t)n_rec + 1);
  rs.seg = to_device(h, A, seg.data(), (size_t)n_rec);
  rs.kn_rec = (int64_t)koff.size() - 1;
  rs.koff = to_device(h, A, koff.data(), koff.size());
  rs.kseg = to_device(h, A, kseg.data(), kseg.size());
  h2d += n_bases + (n_rec + 1) * 8;
}

void pipeline_upload_into(bk_handle_t h, Arena<false>& A, const bk_batch_input* in, Pipeline& p) {
  p = Pipeline();
  if (in->n_regions < 0 || in->n_regions > 65535) fail(BK_ERR_ARG, "batch: n_regions must be in 0..65535 per call");
  if (in->k < 2 || in->k > 31) fail(BK_ERR_ARG, "batch: k must be in 2..31");
  const int R = in->n_regions;
  p.n_regions = R; p.k = in->k; p.rc_thresh = in->rc_thresh; p.have_mers = in->have_mers;
  if (!in->read_off || !in->read_reg_off) fail(BK_ERR_ARG, "batch: read arrays missing");
  upload_record_set(h, A, in->read_bases, in->read_off, in->read_reg_off, R, p.reads, p.h2d_bytes, "read");
  p.read_reg_off = to_device(h, A, in->read_reg_off, (size_t)R + 1);
  p.total_read_bytes = p.reads.n_bases;
  {
    std::vector<int32_t> rl(R ? R : 1, 0);
    int mx = 0;
    for (int r = 0; r < R; ++r) {
      int m = 0;
      for (int64_t i = in->read_reg_off[r]; i < in->read_reg_off[r + 1]; ++i)
        m = std::max<int>(m, (int)(in->read_off[i + 1] - in->read_off[i]));
      mx = std::max(mx, m);
      rl[r] = in->read_len ? in->read_len[r] : m;                    // utils.py:236
    }
    if (mx > NW_MAX_LEN) fail(BK_ERR_CAPACITY, "batch: a read is longer than %d bases", NW_MAX_LEN);
    p.max_read_len = mx;
    p.read_len = to_device(h, A, rl.data(), (size_t)(R ? R : 1));
  }
  if (in->read_flags) p.read_flags = to_device(h, A, in->read_flags, (size_t)p.reads.n_rec);
  p.h2d_bytes += p.reads.n_rec;
  if (in->have_mers) {
    if (!in->in_mers_off) fail(BK_ERR_ARG, "batch: in_mers_off missing");
    p.n_in_mers = in->in_mers_off[R];
    p.in_mers = to_device(h, A, in->in_mers, (size_t)p.n_in_mers);
    p.in_counts = to_device(h, A, in->in_counts, (size_t)p.n_in_mers);
    p.in_mers_off = to_device(h, A, in->in_mers_off, (size_t)R + 1);
    p.h2d_bytes += p.n_in_mers * 12;
  } else {
    if (in->ref_off) {
      std::vector<int64_t> ident(R + 1);
      std::iota(ident.begin(), ident.end(), 0);
      upload_record_set(h, A, in->ref_bases, in->ref_off, ident.data(), R, p.ref, p.h2d_bytes, "ref");
    } else {
      // no reference sequence in this batch: the handle's reference k-mer cache stands in for it
      if (!h->ref_cache_mers) fail(BK_ERR_ARG, "batch: ref arrays missing and no reference k-mer cache on the handle");
      if (h->ref_cache_regions != R || h->ref_cache_k != in->k)
        fail(BK_ERR_ARG, "batch: the reference k-mer cache was built for %d regions, k=%d", h->ref_cache_regions, h->ref_cache_k);
      p.use_ref_cache = true;
    }
    upload_record_set(h, A, in->sc_bases, in->sc_off, in->sc_reg_off, R, p.sc, p.h2d_bytes, "soft-clip");
    upload_record_set(h, A, in->normal_bases, in->normal_off, in->normal_reg_off, R, p.normal, p.h2d_bytes, "normal");
  }
}

void pipeline_upload(bk_handle_t h, const bk_batch_input* in) {
  h->resident.reset();
  h->pipe.reset(new Pipeline());
  pipeline_upload_into(h, h->resident, in, *h->pipe);
  BK_CUDA(stream_wait(h));
}

template <typename T>
T* dev_zero(bk_handle_t h, size_t n) {
  T* p = h->dev.get<T>(n ? n : 1);
  BK_CUDA(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), h->st));
  return p;
}

template <typename T>
const T* to_host(bk_handle_t h, const T* d, size_t n) {
  T* p = h->pin.get<T>(n ? n : 1);
  if (n) BK_CUDA(cudaMemcpyAsync(p, d, n * sizeof(T), cudaMemcpyDeviceToHost, h->st));
  return p;
}

inline unsigned nblk(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

// k-mer windows of the records of regions [r0, r1) of one input set
void emit_set(bk_handle_t h, const RecordSet& rs, int r0, int r1, int k, int tag, bool rc, uint64_t* keys, uint32_t* vals,
              int64_t base, int64_t base_rc, const ProbeTable* probe = nullptr) {
  if (rs.reg_base.empty()) return;
  const int64_t b0 = rs.reg_base[r0], b1 = rs.reg_base[r1];
  if (b1 == b0) return;
  const int64_t kr0 = rs.reg_krec[r0], kr1 = rs.reg_krec[r1];
  EmitParams E{};
  E.bases = rs.bases + b0; E.n_bases = b1 - b0; E.rec_off = rs.koff + kr0; E.n_rec = kr1 - kr0; E.rec_seg = rs.kseg + kr0;
  E.rec_mult = nullptr; E.k = k; E.tag = tag; E.emit_rc = rc ? 1 : 0; E.off_shift = b0; E.seg_shift = r0;
  E.keys = keys; E.vals = vals; E.out_base = base; E.out_base_rc = base_rc;
  if (probe) { E.probe_keys = probe->keys; E.probe_idx = probe->idx; E.probe_mask = probe->mask; E.dead = probe->dead; }
  TimedLaunch t(h->timers, h->st, KF_EMIT);
  kmer_emit_kernel<<<nblk(E.n_bases, EMIT_TILE), EMIT_THREADS, 0, h->st>>>(E);
}

// Persistent reference k-mer cache: forward + reverse-complement k-mers of every target window, counted once,
// kept as sorted per-region mer arrays (the analogue of the marker-file cache of the reference dumps, utils.py:157).
void ref_cache_build(bk_handle_t h, const char* ref_bases, const int64_t* ref_off, int R, int k) {
  cudaStream_t st = h->st;
  h->dev.reset();
  h->pin.reset();
  h->cache.reset();
  h->ref_cache_mers = nullptr; h->ref_cache_koff = nullptr; h->ref_cache_regions = 0; h->ref_cache_k = 0;
  std::vector<int64_t> ident(R + 1);
  std::iota(ident.begin(), ident.end(), 0);
  RecordSet ref;
  int64_t h2d = 0;
  upload_record_set(h, h->dev, ref_bases, ref_off, ident.data(), R, ref, h2d, "ref");
  const int max_seg_bits = 63 - 2 * k;
  const int chunk = (int)std::min<int64_t>(R > 0 ? R : 1, int64_t(1) << std::min(max_seg_bits, 16));
  std::vector<SelectOut> outs;
  std::vector<uint32_t> counts_host;
  uint32_t* seg_counts_all = dev_zero<uint32_t>(h, (size_t)R + 1);
  int64_t total = 0;
  for (int r0 = 0; r0 < R; r0 += chunk) {
    const int r1 = std::min(R, r0 + chunk);
    const int64_t nr = ref.reg_base[r1] - ref.reg_base[r0];
    const int64_t nk = 2 * nr;
    if (nk >= (int64_t(1) << 31)) fail(BK_ERR_CAPACITY, "ref cache: more than 2^31 k-mer windows in one chunk");
    uint64_t* keys = h->dev.get<uint64_t>(nk);
    uint32_t* vals = h->dev.get<uint32_t>(nk);
    emit_set(h, ref, r0, r1, k, TAG_REF, true, keys, vals, 0, nr);
    SelectOut so = sort_and_select(h, keys, vals, nk, k, bits_for((uint64_t)(r1 - r0)), SELECT_ALL, r1 - r0);
    BK_CUDA(cudaMemcpyAsync(seg_counts_all + r0, so.seg_counts, (size_t)(r1 - r0) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    outs.push_back(so);
    total += so.n;
  }
  uint64_t* mers = h->cache.get<uint64_t>(total);
  int64_t* koff = h->cache.get<int64_t>(R + 1);
  int64_t at = 0;
  for (auto& so : outs) {
    if (so.n) BK_CUDA(cudaMemcpyAsync(mers + at, so.mers, so.n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    at += so.n;
  }
  uint32_t* seg_excl = h->dev.get<uint32_t>(R + 1);
  uint32_t* d_tot = h->dev.get<uint32_t>(1);
  uint32_t* stmp = h->dev.get<uint32_t>(scan_tmp_elems(R));
  exclusive_scan_u32(seg_counts_all, seg_excl, R, stmp, d_tot, st);
  widen_scan_kernel<<<nblk(R + 1, 256), 256, 0, st>>>(seg_excl, d_tot, R, koff);
  BK_CUDA(cudaGetLastError());
  BK_CUDA(stream_wait(h));
  h->ref_cache_mers = mers; h->ref_cache_koff = koff; h->ref_cache_regions = R; h->ref_cache_k = k;
}

void pipeline_run(bk_handle_t h, const bk_batch_input* in, bool resident, bk_batch_result* out) {
  cudaStream_t st = h->st;
  h->dev.reset();
  h->pin.reset();
  memset(out, 0, sizeof *out);
  Pipeline local;
  Pipeline* pp;
  struct EventPair {                       // destroyed on every exit path, including exceptions
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  } evs;
  BK_CUDA(cudaEventCreate(&evs.a));
  BK_CUDA(cudaEventCreate(&evs.b));
  cudaEvent_t ev0 = evs.a, ev1 = evs.b;
  BK_CUDA(cudaEventRecord(ev0, st));
  if (resident) {
    if (!h->pipe) fail(BK_ERR_ARG, "bk_compare_kmers_resident: no batch uploaded");
    pp = h->pipe.get();
  } else {
    pipeline_upload_into(h, h->dev, in, local);
    pp = &local;
  }
  const Pipeline& p = *pp;
  const int R = p.n_regions;
  const int k = p.k;
  out->n_regions = R;

  // ---- 1. group identical reads -------------------------------------------------------
  const int64_t n_rec = p.reads.n_rec;
  int64_t NU = 0;
  int32_t* u_rec = nullptr; uint32_t* u_mult = nullptr; uint8_t* u_io = nullptr; int32_t* u_len = nullptr;
  int64_t* u_off = h->dev.get<int64_t>(R + 1);
  if (n_rec > 0) {
    uint64_t* hk = h->dev.get<uint64_t>(n_rec);
    uint32_t* hv = h->dev.get<uint32_t>(n_rec);
    {
      TimedLaunch t(h->timers, st, KF_GROUP);
      read_hash_kernel<<<nblk(n_rec * 32, 128), 128, 0, st>>>(p.reads.bases, p.reads.off, p.reads.seg, n_rec, hk, hv);
    }
    const int64_t tiles = rs_num_tiles(n_rec);
    RadixSortScratch sc;
    sc.keys_alt = h->dev.get<uint64_t>(n_rec);
    sc.vals_alt = h->dev.get<uint32_t>(n_rec);
    sc.table = h->dev.get<uint32_t>(256 * tiles);
    sc.scan_tmp = h->dev.get<uint32_t>(scan_tmp_elems(256 * tiles));
    uint64_t* sk; uint32_t* sv;
    radix_sort_pairs(hk, hv, n_rec, 32, sc, st, &sk, &sv, h->timers, true);   // low 32 hash bits; runs are verified byte-wise
    int32_t* leader_of = h->dev.get<int32_t>(n_rec);
    uint32_t* mult_by_rec = dev_zero<uint32_t>(h, n_rec);
    uint32_t* flag = h->dev.get<uint32_t>(n_rec);
    uint32_t* u_index = h->dev.get<uint32_t>(n_rec);
    uint32_t* d_total = h->dev.get<uint32_t>(1);
    uint32_t* stmp = h->dev.get<uint32_t>(scan_tmp_elems(n_rec));
    {
      TimedLaunch t(h->timers, st, KF_GROUP, 2);
      group_leader_kernel<<<nblk(n_rec, 128), 128, 0, st>>>(sk, sv, n_rec, p.reads.bases, p.reads.off, p.reads.seg, leader_of,
                                                            mult_by_rec);
      leader_flag_kernel<<<nblk(n_rec, 256), 256, 0, st>>>(leader_of, n_rec, flag);
    }
    {
      TimedLaunch t(h->timers, st, KF_SCAN, 3);
      exclusive_scan_u32(flag, u_index, n_rec, stmp, d_total, st);
    }
    const uint32_t* h_total = to_host(h, d_total, 1);
    BK_CUDA(stream_wait(h));
    NU = *h_total;
    u_rec = h->dev.get<int32_t>(NU);
    u_mult = h->dev.get<uint32_t>(NU);
    u_io = h->dev.get<uint8_t>(NU);
    u_len = h->dev.get<int32_t>(NU);
    {
      TimedLaunch t(h->timers, st, KF_GROUP, 2);
      unique_scatter_kernel<<<nblk(n_rec, 256), 256, 0, st>>>(leader_of, u_index, n_rec, mult_by_rec, p.read_flags, p.reads.off, u_rec,
                                                              u_mult, u_io, u_len);
      region_uoff_kernel<<<nblk(R + 1, 256), 256, 0, st>>>(p.read_reg_off, R, u_index, n_rec, d_total, u_off);
    }
  } else {
    BK_CUDA(cudaMemsetAsync(u_off, 0, (R + 1) * sizeof(int64_t), st));
    u_rec = h->dev.get<int32_t>(1); u_mult = h->dev.get<uint32_t>(1); u_io = h->dev.get<uint8_t>(1); u_len = h->dev.get<int32_t>(1);
  }

  // ---- 2. k-mer stage ---------------------------------------------------------------------
  int64_t S_total = 0;
  const uint64_t* so_mer = nullptr; const uint32_t* so_cnt = nullptr;
  int64_t* so_off = h->dev.get<int64_t>(R + 1);
  if (p.have_mers) {
    S_total = p.n_in_mers;
    so_mer = p.in_mers; so_cnt = p.in_counts;
    BK_CUDA(cudaMemcpyAsync(so_off, p.in_mers_off, (R + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
  } else {
    const int64_t n_keys = 2 * p.ref.n_bases + p.reads.n_bases + p.sc.n_bases + p.normal.n_bases;
    out->n_kmer_occurrences = n_keys;
    int64_t n_sorted = 0;
    // The sort key is [0 | region | mer]: 2k + region bits + 1 <= 64.  Large k leaves few region bits, so the stage runs
    // over chunks of regions (one chunk for the usual k); selected mers come out in (region, mer) order either way.
    const int max_seg_bits = 63 - 2 * k;
    const int chunk = (int)std::min<int64_t>(R > 0 ? R : 1, int64_t(1) << std::min(max_seg_bits, 16));
    struct ChunkOut { SelectOut so; int r0, r1; };
    std::vector<ChunkOut> chunks;
    uint32_t* seg_counts_all = dev_zero<uint32_t>(h, (size_t)R + 1);
    auto span = [&](const RecordSet& rs, int r0, int r1) { return rs.reg_base.empty() ? int64_t(0) : rs.reg_base[r1] - rs.reg_base[r0]; };
    for (int r0 = 0; r0 < R; r0 += chunk) {
      const int r1 = std::min(R, r0 + chunk);
      const int64_t nr = span(p.ref, r0, r1), nd = span(p.reads, r0, r1), ns = span(p.sc, r0, r1), nn = span(p.normal, r0, r1);
      // Only the sample's windows are sorted.  The reference (forward + reverse FASTA, Q2) and normal (K4) windows are
      // streamed past the candidates (case & case_sc) afterwards: a hit removes the candidate.
      const bool probe_ref = nr > 0 && !p.use_ref_cache, probe_normal = nn > 0;
      const int64_t nk = nd + ns;
      if (2 * nr + nk + nn >= (int64_t(1) << 31)) fail(BK_ERR_CAPACITY, "batch: more than 2^31 k-mer windows in one chunk; use fewer regions per call");
      n_sorted += nk;
      uint64_t* keys = h->dev.get<uint64_t>(nk);
      uint32_t* vals = h->dev.get<uint32_t>(nk);
      emit_set(h, p.reads, r0, r1, k, TAG_CASE, false, keys, vals, 0, 0);         // every record, duplicates included (Q3)
      emit_set(h, p.sc, r0, r1, k, TAG_SC, false, keys, vals, nd, 0);
      std::function<void(const ProbeTable&)> probe;
      if (probe_ref || probe_normal)
        probe = [&, r0, r1](const ProbeTable& T) {
          if (probe_ref) emit_set(h, p.ref, r0, r1, k, TAG_REF, true, nullptr, nullptr, 0, 0, &T);
          if (probe_normal) emit_set(h, p.normal, r0, r1, k, TAG_NORMAL, false, nullptr, nullptr, 0, 0, &T);
        };
      ChunkOut co;
      co.so = sort_and_select(h, keys, vals, nk, k, bits_for((uint64_t)(r1 - r0)), SELECT_SAMPLE_ONLY, r1 - r0, p.use_ref_cache, r0, probe);
      co.r0 = r0; co.r1 = r1;
      BK_CUDA(cudaMemcpyAsync(seg_counts_all + r0, co.so.seg_counts, (size_t)(r1 - r0) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
      chunks.push_back(co);
    }
    out->n_sorted_keys = n_sorted;
    if (chunks.size() == 1) {
      S_total = chunks[0].so.n;
      so_mer = chunks[0].so.mers; so_cnt = chunks[0].so.counts;
    } else {
      for (auto& c : chunks) S_total += c.so.n;
      uint64_t* all_m = h->dev.get<uint64_t>(S_total);
      uint32_t* all_c = h->dev.get<uint32_t>(S_total);
      int64_t at = 0;
      for (auto& c : chunks) {
        if (c.so.n) {
          BK_CUDA(cudaMemcpyAsync(all_m + at, c.so.mers, c.so.n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
          BK_CUDA(cudaMemcpyAsync(all_c + at, c.so.counts, c.so.n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        }
        at += c.so.n;
      }
      so_mer = all_m; so_cnt = all_c;
    }
    uint32_t* seg_excl = h->dev.get<uint32_t>(R + 1);
    uint32_t* d_tot = h->dev.get<uint32_t>(1);
    uint32_t* stmp = h->dev.get<uint32_t>(scan_tmp_elems(R));
    TimedLaunch t(h->timers, st, KF_SCAN, 4);
    exclusive_scan_u32(seg_counts_all, seg_excl, R, stmp, d_tot, st);
    widen_scan_kernel<<<nblk(R + 1, 256), 256, 0, st>>>(seg_excl, d_tot, R, so_off);
  }

  // per-region table sizes decide the field widths of the packed sort keys below
  const int64_t* h_so_off = to_host(h, so_off, (size_t)R + 1);
  const int64_t* h_u_off = to_host(h, u_off, (size_t)R + 1);
  BK_CUDA(stream_wait(h));
  int64_t max_s = 1, max_u = 1;
  for (int r = 0; r < R; ++r) {
    max_s = std::max<int64_t>(max_s, h_so_off[r + 1] - h_so_off[r]);
    max_u = std::max<int64_t>(max_u, h_u_off[r + 1] - h_u_off[r]);
  }
  const int s_bits = std::max(1, bits_for((uint64_t)max_s)), u_bits = std::max(1, bits_for((uint64_t)max_u));
  if (s_bits > 24 || u_bits > 24) fail(BK_ERR_CAPACITY, "a region has more than 2^24 reads or sample-only k-mers");

  // ---- 3. liveness + seed order -----------------------------------------------------------------
  uint8_t* m_alive = h->dev.get<uint8_t>(S_total);
  int32_t* seed_order = h->dev.get<int32_t>(S_total);
  int* d_overflow = dev_zero<int>(h, 1);
  if (S_total > 0) {
    uint64_t* sk0 = h->dev.get<uint64_t>(S_total);
    uint32_t* sv0 = h->dev.get<uint32_t>(S_total);
    {
      TimedLaunch t(h->timers, st, KF_PREP);
      mer_prep_kernel<<<nblk(S_total, 256), 256, 0, st>>>(so_mer, so_cnt, so_off, R, S_total, k, s_bits, m_alive, sk0, sv0, d_overflow);
    }
    const int64_t tiles = rs_num_tiles(S_total);
    RadixSortScratch sc;
    sc.keys_alt = h->dev.get<uint64_t>(S_total);
    sc.vals_alt = h->dev.get<uint32_t>(S_total);
    sc.table = h->dev.get<uint32_t>(256 * tiles);
    sc.scan_tmp = h->dev.get<uint32_t>(scan_tmp_elems(256 * tiles));
    uint64_t* sk; uint32_t* sv;
    radix_sort_pairs(sk0, sv0, S_total, bits_for((uint64_t)R) + 24 + s_bits, sc, st, &sk, &sv, h->timers, true);
    TimedLaunch t(h->timers, st, KF_PREP);
    unpack_u32_to_i32_kernel<<<nblk(S_total, 256), 256, 0, st>>>(sv, S_total, seed_order);
  }

  // ---- 4. inverted index ------------------------------------------------------------------------------
  int64_t n_post = 0;
  int64_t* post_off = h->dev.get<int64_t>(S_total + 1);
  int32_t* post_read = nullptr; int32_t* post_pos = nullptr;
  int64_t* rk_off = h->dev.get<int64_t>(NU + 1);
  int32_t* rk_s = nullptr; int32_t* rk_pos = nullptr;
  if (S_total > 0 && NU > 0) {
    const int64_t cap = p.reads.n_bases;               // at most one entry per window
    uint64_t* ik = h->dev.get<uint64_t>(cap);
    uint32_t* iv = h->dev.get<uint32_t>(cap);
    uint64_t* ik2 = h->dev.get<uint64_t>(cap);
    unsigned long long* d_n = dev_zero<unsigned long long>(h, 1);
    {
      TimedLaunch t(h->timers, st, KF_INDEX);
      const int ws_stride = ((p.max_read_len + 31) / 32) * 32 + 32;
      const size_t idx_smem = (size_t)IDX_WARPS * ws_stride * sizeof(int32_t);
      if (idx_smem > 48 * 1024) BK_CUDA(cudaFuncSetAttribute(index_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)idx_smem));
      index_emit_kernel<<<nblk(NU, IDX_WARPS), 32 * IDX_WARPS, idx_smem, st>>>(p.reads.bases, p.reads.off, u_off, u_rec, R, NU, so_off,
                                                                               so_mer, k, ik, iv, ik2, u_bits, s_bits, ws_stride, d_n,
                                                                               (unsigned long long)cap);
    }
    const unsigned long long* h_n = to_host(h, d_n, 1);
    BK_CUDA(stream_wait(h));
    n_post = (int64_t)*h_n;
    if (n_post > cap) fail(BK_ERR_CAPACITY, "index: posting overflow");
    post_read = h->dev.get<int32_t>(n_post);
    post_pos = h->dev.get<int32_t>(n_post);
    rk_s = h->dev.get<int32_t>(n_post);
    rk_pos = h->dev.get<int32_t>(n_post);
    uint64_t* sk = ik; uint32_t* sv = iv;
    uint64_t* sk2 = ik2; uint32_t* sv2 = nullptr;
    if (n_post > 0) {
      // the read -> k-mers copy gets its own value array (same positions) before either sort permutes it
      uint32_t* iv2 = h->dev.get<uint32_t>(n_post);
      BK_CUDA(cudaMemcpyAsync(iv2, iv, n_post * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
      sv2 = iv2;
      const int64_t tiles = rs_num_tiles(n_post);
      RadixSortScratch sc;
      sc.keys_alt = h->dev.get<uint64_t>(n_post);
      sc.vals_alt = h->dev.get<uint32_t>(n_post);
      sc.table = h->dev.get<uint32_t>(256 * tiles);
      sc.scan_tmp = h->dev.get<uint32_t>(scan_tmp_elems(256 * tiles));
      radix_sort_pairs(ik, iv, n_post, u_bits + bits_for((uint64_t)S_total + 1), sc, st, &sk, &sv, h->timers, true);
      RadixSortScratch sc2 = sc;
      sc2.keys_alt = h->dev.get<uint64_t>(n_post);
      sc2.vals_alt = h->dev.get<uint32_t>(n_post);
      radix_sort_pairs(ik2, iv2, n_post, s_bits + bits_for((uint64_t)NU + 1), sc2, st, &sk2, &sv2, h->timers, true);
    }
    TimedLaunch t(h->timers, st, KF_INDEX, 4);
    post_off_kernel<<<nblk(S_total + 1, 256), 256, 0, st>>>(sk, n_post, S_total, u_bits, post_off);
    post_off_kernel<<<nblk(NU + 1, 256), 256, 0, st>>>(sk2, n_post, NU, s_bits, rk_off);
    if (n_post > 0) {
      post_split_kernel<<<nblk(n_post, 256), 256, 0, st>>>(sk, sv, n_post, u_bits, post_read, post_pos);
      post_split_kernel<<<nblk(n_post, 256), 256, 0, st>>>(sk2, sv2, n_post, s_bits, rk_s, rk_pos);
    }
  } else {
    BK_CUDA(cudaMemsetAsync(post_off, 0, (S_total + 1) * sizeof(int64_t), st));
    BK_CUDA(cudaMemsetAsync(rk_off, 0, (NU + 1) * sizeof(int64_t), st));
    post_read = h->dev.get<int32_t>(1); post_pos = h->dev.get<int32_t>(1);
    rk_s = h->dev.get<int32_t>(1); rk_pos = h->dev.get<int32_t>(1);
  }

  // work order
  const int* h_overflow = to_host(h, d_overflow, 1);
  BK_CUDA(stream_wait(h));
  if (*h_overflow) fail(BK_ERR_CAPACITY, "seed order: a k-mer count or region size exceeds 2^24");
  std::vector<int32_t> order(R);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int a, int b) {     // most expensive regions first (longest processing time)
    const int64_t ca = (h_so_off[a + 1] - h_so_off[a]) * (h_u_off[a + 1] - h_u_off[a]);
    const int64_t cb = (h_so_off[b + 1] - h_so_off[b]) * (h_u_off[b + 1] - h_u_off[b]);
    return ca != cb ? ca > cb : a < b;
  });

  // ---- 5. assembly ----------------------------------------------------------------------------------------
  AsmParams A;
  memset(&A, 0, sizeof A);
  A.n_regions = R; A.k = k; A.rc_thresh = p.rc_thresh;
  A.rbases = p.reads.bases; A.roff = p.reads.off;
  A.u_off = u_off; A.u_rec = u_rec; A.u_mult = u_mult; A.u_io = u_io; A.u_len = u_len; A.read_len = p.read_len;
  A.rk_off = rk_off; A.rk_s = rk_s; A.rk_pos = rk_pos;
  A.so_off = so_off; A.so_mer = so_mer; A.so_cnt = so_cnt; A.seed_order = seed_order;
  A.post_off = post_off; A.post_read = post_read; A.post_pos = post_pos;
  A.work_order = to_device(h, h->dev, order.data(), (size_t)R);
  // speculation width and residency of the assembler (see assemble_kernel).  Auto = 4: one aligning warp per SM
  // sub-partition; 8 warps share the four ALU pipes of the SM and finish a round no sooner (measured on C4).
  int spec_w = h->spec_width;
  if (const char* e = getenv("BK_SPEC_W")) spec_w = atoi(e);
  if (spec_w <= 0) spec_w = 4;
  spec_w = spec_w >= 8 ? 8 : (spec_w >= 4 ? 4 : (spec_w >= 2 ? 2 : 1));
  int ctas_per_sm = spec_w == 8 ? 1 : (spec_w == 4 ? 3 : (spec_w == 2 ? 6 : 8));
  if (const char* e = getenv("BK_ASM_CTAS_PER_SM")) ctas_per_sm = std::max(1, atoi(e));
  int grid = std::min<int64_t>(R, (int64_t)h->sm_count * ctas_per_sm);
  // dynamic shared memory: exactly what the kernel needs.  (An earlier version padded it to 1/ctas_per_sm of the SM to bound
  // residency; that pushed the shared-memory carve-out to the maximum and left the assembler's bookkeeping -- dependent
  // loads of per-region tables -- ~29 KB of L1: 15 % slower.  BK_ASM_PAD=1 restores it for experiments.)
  A.read_cap = (int)std::min<int64_t>(ASM_CAP, std::max<int64_t>(64, (p.max_read_len + 2 + 15) / 16 * 16));
  const size_t need_smem = assemble_smem_bytes(spec_w, A.read_cap);
  int dyn_smem = (int)need_smem;
  if (getenv("BK_ASM_PAD")) dyn_smem = std::max(dyn_smem, (int)((227 * 1024) / ctas_per_sm) - 1024);
  if (grid < 1) grid = 1;
  A.w_cseq = h->dev.get<uint8_t>((size_t)grid * ASM_BUF);
  A.w_cnt = h->dev.get<int32_t>((size_t)grid * 4 * ASM_BUF);
  A.w_K = h->dev.get<int32_t>((size_t)grid * 4 * ASM_KCAP);
  A.w_NK = h->dev.get<int32_t>((size_t)grid * 4 * ASM_KCAP);
  A.w_wcode = h->dev.get<uint64_t>((size_t)grid * ASM_CAP);
  A.w_diff = h->dev.get<int32_t>((size_t)grid * (ASM_CAP + 1));
  A.w_edge = p.max_read_len > 256 ? h->dev.get<int2>((size_t)grid * spec_w * 2 * ASM_CAP) : nullptr;
  A.w_lastcol = h->dev.get<uint2>((size_t)grid * spec_w * ASM_LASTCOL);
  A.region_status = h->dev.get<int32_t>(R ? R : 1);
  A.region_ncontigs = h->dev.get<int32_t>(R ? R : 1);
  // a mutable copy of the liveness flags per attempt, everything else zeroed
  uint8_t* alive_run = h->dev.get<uint8_t>(S_total);
  A.m_alive = alive_run;
  uint8_t* zero_lo = nullptr;
  size_t zero_bytes = 0;
  auto zalloc = [&](size_t bytes) { bytes = (bytes + 255) & ~size_t(255); size_t o = zero_bytes; zero_bytes += bytes; return o; };
  const size_t o_mused = zalloc(S_total), o_checked = zalloc(S_total * 4), o_taken = zalloc(S_total * 4), o_first = zalloc(S_total * 4);
  const size_t o_rused = zalloc(NU), o_rdel = zalloc(NU), o_rq = zalloc(NU), o_rbuf = zalloc(NU * 4), o_rin = zalloc(NU * 4);
  const size_t o_work = zalloc(sizeof(int)), o_cursor = zalloc(5 * sizeof(unsigned long long)), o_stats = zalloc(16 * sizeof(unsigned long long));
  zero_lo = h->dev.get<uint8_t>(zero_bytes);
  A.m_used = zero_lo + o_mused; A.m_checked = (uint32_t*)(zero_lo + o_checked); A.m_taken = (uint32_t*)(zero_lo + o_taken); A.m_first = (uint32_t*)(zero_lo + o_first);
  A.r_used = zero_lo + o_rused; A.r_deleted = zero_lo + o_rdel; A.r_queued = zero_lo + o_rq;
  A.r_buf = (uint32_t*)(zero_lo + o_rbuf); A.r_inreads = (uint32_t*)(zero_lo + o_rin);
  A.work_counter = (int*)(zero_lo + o_work);
  A.out_cursor = (unsigned long long*)(zero_lo + o_cursor);
  A.stats = (unsigned long long*)(zero_lo + o_stats);
  A.q_read = h->dev.get<int32_t>(NU); A.q_seed = h->dev.get<int32_t>(NU);
  A.l_alt = h->dev.get<int32_t>(NU); A.l_del = h->dev.get<int32_t>(NU);
  A.hit_u = h->dev.get<int32_t>(NU); A.hit_pos = h->dev.get<int32_t>(NU);
  A.hit2_u = h->dev.get<int32_t>(NU); A.hit2_pos = h->dev.get<int32_t>(NU);

  const bool want_prof = getenv("BK_PHASE_PRINT") != nullptr;
  A.prof_regions = want_prof ? dev_zero<unsigned long long>(h, (size_t)R * 12 + 12) : nullptr;
  unsigned long long cap_seq = (unsigned long long)std::max<int64_t>(1 << 20, 8 * p.total_read_bytes);
  const unsigned long long* h_cursor = nullptr;
  const unsigned long long* h_stats = nullptr;
  const int32_t* h_status = nullptr;
  for (int attempt = 0;; ++attempt) {
    A.cap_seq = cap_seq; A.cap_cnt = cap_seq; A.cap_reads = cap_seq; A.cap_kmers = 2 * cap_seq;
    A.cap_ctg = std::max<unsigned long long>(1024, cap_seq / 64);
    A.o_seq = h->dev.get<uint8_t>(A.cap_seq); A.o_locs = h->dev.get<int32_t>(A.cap_seq);
    A.o_io = h->dev.get<int32_t>(A.cap_cnt); A.o_ot = h->dev.get<int32_t>(A.cap_cnt);
    A.o_reads = h->dev.get<int32_t>(A.cap_reads);
    A.o_kmer_mer = h->dev.get<uint64_t>(A.cap_kmers); A.o_kmer_pos = h->dev.get<int32_t>(A.cap_kmers);
    A.o_kmer_meta = h->dev.get<int32_t>(A.cap_kmers);
    A.o_desc = h->dev.get<int64_t>(A.cap_ctg * 10);
    BK_CUDA(cudaMemsetAsync(zero_lo, 0, zero_bytes, st));
    if (S_total) BK_CUDA(cudaMemcpyAsync(alive_run, m_alive, S_total, cudaMemcpyDeviceToDevice, st));
    if (R > 0) {
      TimedLaunch t(h->timers, st, KF_ASSEMBLE);
      // shared-memory carve-out: just enough for the resident CTAs, the rest of the 256 KB stays L1
      int carve = (int)((100 * (size_t)ctas_per_sm * ((size_t)dyn_smem + 1024) + 228 * 1024 - 1) / (228 * 1024));
      if (const char* e = getenv("BK_ASM_CARVEOUT")) carve = atoi(e);
      carve = std::min(100, std::max(0, carve));
      auto prep = [&](const void* fn) {
        BK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem));
        BK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
      };
      if (spec_w == 8) {
        prep((const void*)assemble_kernel<8>);
        assemble_kernel<8><<<grid, 256, dyn_smem, st>>>(A);
      } else if (spec_w == 4) {
        prep((const void*)assemble_kernel<4>);
        assemble_kernel<4><<<grid, 128, dyn_smem, st>>>(A);
      } else if (spec_w == 2) {
        prep((const void*)assemble_kernel<2>);
        assemble_kernel<2><<<grid, 64, dyn_smem, st>>>(A);
      } else {
        prep((const void*)assemble_kernel<1>);
        assemble_kernel<1><<<grid, 32, dyn_smem, st>>>(A);
      }
    }
    BK_CUDA(cudaGetLastError());
    h_cursor = to_host(h, A.out_cursor, 5);
    h_stats = to_host(h, A.stats, 16);
    h_status = to_host(h, A.region_status, (size_t)(R ? R : 1));
    BK_CUDA(stream_wait(h));
    const bool overflow = h_cursor[0] > A.cap_seq || h_cursor[1] > A.cap_cnt || h_cursor[2] > A.cap_reads ||
                          h_cursor[3] > A.cap_kmers || h_cursor[4] > A.cap_ctg;
    if (!overflow) break;
    if (attempt >= 3) fail(BK_ERR_CAPACITY, "assembly output arena overflow");
    cap_seq *= 8;                                          // rare: rerun the assembly with a larger arena
  }

  // ---- 6. results -----------------------------------------------------------------------------------------------
  const int64_t n_ctg = (int64_t)h_cursor[4];
  out->n_contigs = n_ctg;
  out->n_check_align = (int64_t)h_stats[0];
  out->n_dp_cells = (int64_t)h_stats[1];
  if (getenv("BK_PHASE_PRINT")) {
    static const char* nm[] = {"nw", "find_reads", "kmers", "finalize", "emit", "predict", "total", "max_region"};
    for (int i = 0; i < 8; ++i) fprintf(stderr, "phase %-10s %12llu cycles\n", nm[i], h_stats[8 + i]);
    fprintf(stderr, "find_reads calls %llu seeds %llu check_align %llu rounds %llu slots %llu\n", h_stats[2], h_stats[3], h_stats[0],
            h_stats[4], h_stats[5]);
    if (A.prof_regions) {
      std::vector<unsigned long long> pr((size_t)R * 12);
      BK_CUDA(cudaMemcpy(pr.data(), A.prof_regions, pr.size() * 8, cudaMemcpyDeviceToHost));
      std::vector<int> ids(R);
      std::iota(ids.begin(), ids.end(), 0);
      std::sort(ids.begin(), ids.end(), [&](int a, int b) { return pr[(size_t)a * 12 + 6] > pr[(size_t)b * 12 + 6]; });
      {
        unsigned long long t0 = ~0ull, t1 = 0, sum = 0;
        for (int r = 0; r < R; ++r) {
          const unsigned long long a = pr[(size_t)r * 12 + 8], b = pr[(size_t)r * 12 + 9];
          if (a < t0) t0 = a;
          if (b > t1) t1 = b;
          sum += b - a;
        }
        fprintf(stderr, "TIMELINE handle %p regions_start_ns %llu end_ns %llu span_ms %.3f sum_region_ms %.3f avg_concurrency %.1f\n", (void*)h,
                t0, t1, (t1 - t0) / 1e6, sum / 1e6, (double)sum / (double)(t1 - t0 + 1));
        if (getenv("BK_REGION_DUMP")) {
          FILE* f = fopen(getenv("BK_REGION_DUMP"), "w");
          if (f) {
            for (int r = 0; r < R; ++r)
              fprintf(f, "%d %lld %lld %llu %llu %llu\n", r, (long long)(h_u_off[r + 1] - h_u_off[r]), (long long)(h_so_off[r + 1] - h_so_off[r]),
                      pr[(size_t)r * 12 + 6], pr[(size_t)r * 12 + 0], pr[(size_t)r * 12 + 8] - t0);
            fclose(f);
          }
        }
        if (getenv("BK_TIMELINE_DUMP")) {
          FILE* f = fopen(getenv("BK_TIMELINE_DUMP"), "a");
          if (f) {
            for (int r = 0; r < R; ++r)
              fprintf(f, "%p %d %llu %llu %llu\n", (void*)h, r, pr[(size_t)r * 12 + 8], pr[(size_t)r * 12 + 9], pr[(size_t)r * 12 + 10]);
            fclose(f);
          }
        }
      }
      for (int t = 0; t < 4 && t < R; ++t) {
        const unsigned long long* q = &pr[(size_t)ids[t] * 12];
        fprintf(stderr, "region %d (U=%lld S=%lld): total %llu nw %llu find %llu kmers %llu finalize %llu emit %llu predict %llu\n", ids[t],
                (long long)(h_u_off[ids[t] + 1] - h_u_off[ids[t]]), (long long)(h_so_off[ids[t] + 1] - h_so_off[ids[t]]), q[6], q[0], q[1],
                q[2], q[3], q[4], q[5]);
      }
    }
  }
  out->so_off = h_so_off;
  out->so_mers = to_host(h, so_mer, (size_t)S_total);
  out->so_counts = to_host(h, so_cnt, (size_t)S_total);
  out->uniq_reg_off = h_u_off;
  out->uniq_rec = to_host(h, u_rec, (size_t)NU);
  out->uniq_mult = to_host(h, u_mult, (size_t)NU);
  out->region_status = h_status;
  const int64_t* h_desc = to_host(h, A.o_desc, (size_t)n_ctg * 10);
  out->ctg_seq = (const char*)to_host(h, A.o_seq, (size_t)h_cursor[0]);
  out->ctg_kmer_locs = to_host(h, A.o_locs, (size_t)h_cursor[0]);
  out->ctg_indel_only = to_host(h, A.o_io, (size_t)h_cursor[1]);
  out->ctg_others = to_host(h, A.o_ot, (size_t)h_cursor[1]);
  out->ctg_reads = to_host(h, A.o_reads, (size_t)h_cursor[2]);
  out->ctg_kmer_mer = to_host(h, A.o_kmer_mer, (size_t)h_cursor[3]);
  out->ctg_kmer_pos = to_host(h, A.o_kmer_pos, (size_t)h_cursor[3]);
  const int32_t* h_meta = to_host(h, A.o_kmer_meta, (size_t)h_cursor[3]);
  BK_CUDA(cudaEventRecord(ev1, st));
  BK_CUDA(stream_wait(h));
  float ms = 0;
  BK_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
  out->gpu_ms = ms;

  // contigs left the device in completion order; the table below puts them in
  // (region, acceptance order) -- payload arrays stay where they are
  std::vector<int64_t> idx(n_ctg);
  std::iota(idx.begin(), idx.end(), 0);
  std::sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) {
    const int64_t* da = h_desc + a * 10; const int64_t* db = h_desc + b * 10;
    return da[0] != db[0] ? da[0] < db[0] : da[1] < db[1];
  });
  int64_t* t_reg = h->pin.get<int64_t>(R + 1);
  int64_t* t_seq = h->pin.get<int64_t>(2 * n_ctg + 2);
  int64_t* t_cnt = h->pin.get<int64_t>(2 * n_ctg + 2);
  int64_t* t_rd = h->pin.get<int64_t>(2 * n_ctg + 2);
  int64_t* t_km = h->pin.get<int64_t>(2 * n_ctg + 2);
  for (int r = 0; r <= R; ++r) t_reg[r] = 0;
  for (int64_t c = 0; c < n_ctg; ++c) {
    const int64_t* d = h_desc + idx[c] * 10;
    t_reg[d[0] + 1] += 1;
    t_seq[2 * c] = d[2]; t_seq[2 * c + 1] = d[3];
    t_cnt[2 * c] = d[4]; t_cnt[2 * c + 1] = d[5];
    t_rd[2 * c] = d[6]; t_rd[2 * c + 1] = d[7];
    t_km[2 * c] = d[8]; t_km[2 * c + 1] = d[9];
  }
  for (int r = 0; r < R; ++r) t_reg[r + 1] += t_reg[r];
  out->ctg_reg_off = t_reg;
  out->ctg_seq_off = t_seq; out->ctg_cnt_off = t_cnt; out->ctg_reads_off = t_rd; out->ctg_kmers_off = t_km;
  // split the packed tuple meta word into the three fields of the 5-tuple
  int32_t* lth = h->pin.get<int32_t>(h_cursor[3] + 1);
  int32_t* dist = h->pin.get<int32_t>(h_cursor[3] + 1);
  int32_t* ord = h->pin.get<int32_t>(h_cursor[3] + 1);
  for (unsigned long long e = 0; e < h_cursor[3]; ++e) {
    lth[e] = h_meta[e] & 1; ord[e] = (h_meta[e] >> 1) & 3; dist[e] = h_meta[e] >> 3;
  }
  out->ctg_kmer_lth = lth; out->ctg_kmer_dist = dist; out->ctg_kmer_order = ord;
}

}  // namespace
