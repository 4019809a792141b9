// Host build of the decision-chain replay of bk_dedup_reads (breakmer_b200/csrc/dedup.cuh), so that the container
// without a GPU can check it against the golden vectors: the caller supplies the aligner (the oracle's nw) that the
// device kernel is in the product.  Test infrastructure only.
#include "../../include/breakmer_b200.h"
#include "../../breakmer_b200/csrc/dedup.cuh"

typedef int (*align_fn)(const int32_t* pair_a, const int32_t* pair_b, int64_t n, int32_t* out);

extern "C" int dedup_sim_run(const int64_t* seq_off, const int32_t* mer_pos, const int64_t* batch_off, int64_t n_batches,
                             double frac, uint8_t* check, uint8_t* flags, int64_t* n_pairs, int* n_rounds, align_fn align) {
  return dedup_run(seq_off, mer_pos, batch_off, n_batches, frac, check, flags, n_pairs, n_rounds, align);
}
