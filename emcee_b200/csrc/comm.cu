// Multi-GPU plumbing (see comm.h).  Phase 1: single rank only.
#include "comm.h"

#include <string.h>

namespace eb {

int comm_unique_id(char* id128) {
  memset(id128, 0, EB_COMM_ID_BYTES);
  return 1;
}

int comm_init(Comm& c, const char*, int rank, int nranks, int mode, int64_t N, int D, double* coords,
              cudaStream_t) {
  if (nranks != 1 || rank != 0) {
    c.err = "multi-GPU communicator not built yet";
    return 1;
  }
  c.rank = 0;
  c.nranks = 1;
  c.mode = mode;
  c.N = N;
  c.D = D;
  c.coords = coords;
  c.rows_per_rank = N;
  return 0;
}

int comm_export(Comm& c, char*) {
  c.err = "peer memory not built yet";
  return 1;
}
int comm_import(Comm& c, const char*) {
  c.err = "peer memory not built yet";
  return 1;
}
void comm_destroy(Comm&) {}

void comm_fill_args(const Comm&, HalfStepArgs& a) {
  a.peer_coords = nullptr;
  a.rows_per_rank = a.N;
}

int comm_active_range(Comm&, cudaStream_t, HalfStepArgs& a, const int32_t*) {
  a.i_lo = 0;
  a.i_hi = a.a_count;
  return 0;
}

int comm_after_split(Comm&, cudaStream_t, const HalfStepArgs&, uint64_t&) { return 0; }

}  // namespace eb
