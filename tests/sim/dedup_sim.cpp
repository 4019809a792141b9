// Host build of the decision-chain replay of bk_dedup_reads (breakmer_b200/csrc/dedup.cuh), so that the container
// without a GPU can check it against the golden vectors: the caller supplies the score table (from the oracle's nw)
// that the device kernel produces in the product.  Test infrastructure only.
#include "../../include/breakmer_b200.h"
#include "../../breakmer_b200/csrc/dedup.cuh"

extern "C" void dedup_sim_replay(const int64_t* seq_off, const int32_t* mer_pos, int64_t lo, int64_t hi, const int32_t* tab,
                                 double frac, uint8_t* check, uint8_t* flags) {
  dedup_replay(seq_off, mer_pos, lo, hi, tab, frac, check, flags);
}
