// K-mer stage kernels (kernels G1-G4).
//
//   kmer_emit_kernel   G1+G2: ASCII bases -> 2-bit codes in shared memory ->
//                      one 64-bit key per window position.  Jellyfish 1.1.11
//                      semantics as restated in oracle/kmers_py.py: forward
//                      strand only (utils.py:160 has no -C), a window holding a
//                      non-ACGT base or crossing a record boundary is dropped,
//                      lower case folded.  With emit_rc the reverse complement
//                      of every window is emitted too -- that is the second
//                      FASTA the reference writes and counts for the target
//                      (utils.py:367-371, sv_processor.py:613-615).
//   (radix_sort.cuh)   G3: sort by (region, mer); the set tag travels in the value
//   run_select_kernel  G3+G4: one pass over the sorted keys does the run-length
//                      count AND the set algebra of sv_processor.py:621-622:
//                      all occurrences of a (region, mer) are adjacent whatever
//                      set they came from, so (case & case_sc) - ref [- normal]
//                      is a property of the run.
//   run_scatter_kernel compaction of the selected runs to (mer, count) arrays.
//
// These kernels back bk_count_kmers, bk_sample_only and the reference k-mer cache (sorted outputs over arbitrary
// inputs).  The batched pipeline itself counts per region in shared-memory hash tables: region_kmers.cuh.
//
// Key layout:    [ 0 | region | mer (2k bits) ]          (one spare top bit: the all-ones invalid key sorts last)
// Value layout:  [ set tag : 2 | multiplicity : 30 ]
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace bk {

enum : int { TAG_CASE = 0, TAG_SC = 1, TAG_REF = 2, TAG_NORMAL = 3 };

constexpr int EMIT_THREADS = 256;
constexpr int EMIT_PER_THREAD = 4;
constexpr int EMIT_TILE = EMIT_THREADS * EMIT_PER_THREAD;

struct EmitParams {
  const uint8_t* bases;
  int64_t n_bases;
  const int64_t* rec_off;     // n_rec + 1, strictly increasing (no empty records)
  int64_t n_rec;
  const int32_t* rec_seg;     // region of each record, or null (all 0)
  const uint32_t* rec_mult;   // multiplicity of each record, or null (all 1)
  int k;
  int tag;
  int emit_rc;
  int64_t off_shift;          // rec_off values are relative to `bases` after subtracting this (region chunks)
  int seg_shift;              // region ids are stored relative to this (region chunks)
  uint64_t* keys;             // [out_base + p] forward, [out_base_rc + p] reverse complement
  uint32_t* vals;
  int64_t out_base, out_base_rc;
};

__device__ __forceinline__ uint64_t cand_hash(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 29;
  return x;
}

__global__ void __launch_bounds__(EMIT_THREADS) kmer_emit_kernel(EmitParams P) {
  __shared__ uint8_t code[EMIT_TILE + 32];
  __shared__ int64_t s_rfirst;
  const int tid = threadIdx.x;
  const int64_t p0 = (int64_t)blockIdx.x * EMIT_TILE;
  const int k = P.k;
  if (tid == 0) {
    // last record whose start is <= p0
    int64_t lo = 0, hi = P.n_rec;           // rec_off[lo] <= p0 < rec_off[hi]
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (P.rec_off[mid] - P.off_shift <= p0) lo = mid; else hi = mid;
    }
    s_rfirst = lo;
  }
  for (int x = tid; x < EMIT_TILE + k - 1; x += EMIT_THREADS) {
    const int64_t p = p0 + x;
    code[x] = (p < P.n_bases) ? (uint8_t)base_code(P.bases[p]) : (uint8_t)4;
  }
  __syncthreads();
  const int64_t rfirst = s_rfirst;
  const int64_t lim = p0 + EMIT_TILE + k - 1;
  for (int64_t r = rfirst + 1 + tid; r < P.n_rec; r += EMIT_THREADS) {
    const int64_t o = P.rec_off[r] - P.off_shift;
    if (o >= lim) break;
    code[o - p0] |= 8;                       // a record starts here
  }
  __syncthreads();
  // record index of each position = rfirst + number of starts at or before it
  const int xb = tid * EMIT_PER_THREAD;
  uint32_t cnt = 0;
#pragma unroll
  for (int q = 0; q < EMIT_PER_THREAD; ++q) cnt += (code[xb + q] >> 3) & 1u;
  uint32_t total;
  uint32_t run = block_excl_scan(cnt, &total);
  // the thread's EMIT_PER_THREAD consecutive windows share k-1 bases: build the first
  // window (k loads), then roll (one load per further window)
  const uint64_t kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1ull);
  uint64_t fwd = 0, rc = 0;
  int bad = 0;
#pragma unroll
  for (int q = 0; q < EMIT_PER_THREAD; ++q) {
    const int x = xb + q;
    run += (code[x] >> 3) & 1u;
    const int64_t p = p0 + x;
    if (p >= P.n_bases) break;
    // `bad` = how many consecutive windows, starting with the current one, are ruled out by positions seen so
    // far: an invalid base at relative position t rules out windows 0..t, a record start at t >= 1 windows 0..t-1
    if (q == 0) {
      for (int t = 0; t < k; ++t) {
        const unsigned c = code[x + t];
        fwd = (fwd << 2) | (c & 3u);
        rc |= (uint64_t)(3u - (c & 3u)) << (2 * t);
        const int b = (c & 4u) ? t + 1 : ((t > 0 && (c & 8u)) ? t : 0);
        bad = b > bad ? b : bad;
      }
    } else {
      const unsigned c = code[x + k - 1];
      fwd = ((fwd << 2) | (c & 3u)) & kmask;
      rc = (rc >> 2) | ((uint64_t)(3u - (c & 3u)) << (2 * (k - 1)));
      bad = bad > 0 ? bad - 1 : 0;
      const int b = (c & 4u) ? k : (((c & 8u) && k > 1) ? k - 1 : 0);
      bad = b > bad ? b : bad;
    }
    const bool ok = bad == 0;
    const int64_t rec = rfirst + run;
    uint64_t seg = 0;
    uint32_t mult = 1;
    if (ok) {
      if (P.rec_seg) seg = (uint64_t)(P.rec_seg[rec] - P.seg_shift);
      if (P.rec_mult) mult = P.rec_mult[rec];
    }
    const uint64_t hi = seg << (2 * k);
    const uint32_t v = (mult & 0x3FFFFFFFu) | ((uint32_t)P.tag << 30);
    P.keys[P.out_base + p] = ok ? (hi | fwd) : KEY_INVALID;
    P.vals[P.out_base + p] = v;
    if (P.emit_rc) {
      P.keys[P.out_base_rc + p] = ok ? (hi | rc) : KEY_INVALID;
      P.vals[P.out_base_rc + p] = v;
    }
  }
}

// ---------------------------------------------------------------------------
enum : int { SELECT_ALL = 0, SELECT_SAMPLE_ONLY = 1 };

struct RunParams {
  const uint64_t* keys;       // sorted
  const uint32_t* vals;
  int64_t n;
  int k;
  int mode;
  uint32_t* flags;            // n : 1 at the head of a selected run
  uint32_t* run_count;        // n : count of the run (valid at selected heads)
  // scatter stage
  const uint32_t* pos;        // exclusive scan of flags
  uint64_t* out_mers;
  uint32_t* out_counts;
  uint32_t* seg_counts;       // per region number of selected runs (atomic), may be null
};

__global__ void __launch_bounds__(256) run_select_kernel(RunParams P) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  const uint64_t key = P.keys[i];
  uint32_t flag = 0, total = 0;
  if (key != KEY_INVALID) {
    const bool head = (i == 0) || (P.keys[i - 1] != key);
    if (head) {
      uint32_t c_case = 0;
      unsigned seen = 0;
      for (int64_t j = i; j < P.n; ++j) {
        if (P.keys[j] != key) break;
        const uint32_t vv = P.vals[j];
        const unsigned tag = vv >> 30;
        const uint32_t v = vv & 0x3FFFFFFFu;
        seen |= 1u << tag;
        total += v;
        if (tag == TAG_CASE) c_case += v;
      }
      if (P.mode == SELECT_ALL) {
        flag = 1;
      } else {
        // (case & case_sc) - ref - normal ; reported count is the case count (sv_processor.py:621-631)
        bool sel = (seen & (1u << TAG_CASE)) && (seen & (1u << TAG_SC)) && !(seen & (1u << TAG_REF)) &&
                   !(seen & (1u << TAG_NORMAL));
        flag = sel ? 1u : 0u;
        total = c_case;
      }
    }
  }
  P.flags[i] = flag;
  P.run_count[i] = total;
}

__global__ void __launch_bounds__(256) run_scatter_kernel(RunParams P) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  if (!P.flags[i]) return;
  const uint64_t g = P.keys[i];
  const uint64_t mer = g & ((P.k == 32) ? ~0ull : ((1ull << (2 * P.k)) - 1ull));
  const uint32_t dst = P.pos[i];
  P.out_mers[dst] = mer;
  P.out_counts[dst] = P.run_count[i];
  if (P.seg_counts) atomicAdd(&P.seg_counts[(uint32_t)(g >> (2 * P.k))], 1u);
}

__global__ void set_u32_kernel(uint32_t* __restrict__ dst, uint32_t v) { *dst = v; }
__global__ void add_u32_kernel(const uint32_t* __restrict__ a, uint32_t* __restrict__ acc) { *acc += *a; }

}  // namespace bk
