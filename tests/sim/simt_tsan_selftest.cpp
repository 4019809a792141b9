// tests/sim: self-test of the race detector (simt_host.h under -fsanitize=thread -DSIMT_TSAN): the pattern that was
// found three times in assemble.cuh -- every lane reads a flag, lane 0 sets it -- with and without the barrier in between,
// behind earlier barriers and in the last warp of a four-warp block.  argv[1] = "racy" | "fixed".
#include <stdio.h>
#include <string.h>

#include <vector>

#include "simt_host.h"

int sink[256];

int main(int argc, char** argv) {
  const bool racy = argc > 1 && strcmp(argv[1], "racy") == 0;
  std::vector<unsigned char> flags(64, 0);
  unsigned char* r_used = flags.data();
  for (int rep = 0; rep < 3; ++rep)
    simt::run_block(4, 0, [&]() { __syncwarp(); __syncthreads(); });
  simt::run_block(4, 0, [&]() {
    const int l = simt::lane_id(), w = simt::warp_id();
    if (w == 3) {
      const int u = 7;
      __syncwarp();
      const bool fresh = !r_used[u];          // every lane reads the flag ...
      if (!racy) __syncwarp();
      if (l == 0 && fresh) r_used[u] = 1;     // ... lane 0 sets it
      sink[simt::thread_id()] = fresh;
      __syncwarp();
    }
    __syncthreads();
  });
  printf("done\n");
  return 0;
}
