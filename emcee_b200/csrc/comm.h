// Multi-GPU plumbing of the engine: one process per GPU, walkers sharded by
// contiguous row block (rank r owns walkers [r*N/R, (r+1)*N/R)).
//   EB_COMM_ALLGATHER: every rank keeps a full replica of coords; after each
//     split one in-place ncclAllGather of the owned row blocks (NCCL is
//     dlopen-ed, so single-GPU use needs no NCCL at all).
//   EB_COMM_P2P: no replica traffic; the half-step kernel loads partner rows
//     straight from the owner's HBM over NVLink (cudaIpc-mapped peer pointers)
//     and ranks meet at a flag barrier in peer memory between splits.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "engine.cuh"

namespace eb {

constexpr int MAX_RANKS = 16;

struct Comm {
  int rank = 0, nranks = 1, mode = EB_COMM_ALLGATHER;
  int64_t N = 0;
  int D = 0;
  int64_t rows_per_rank = 0;
  double* coords = nullptr;   // this rank's [N, D] buffer (+ barrier flags in its tail)
  unsigned* flags = nullptr;  // this rank's barrier flags [MAX_RANKS], inside the coords allocation
  int2* ranges = nullptr;     // [table_cap, MAX_SPLITS] active-rank range of this rank per (step, set)
  int32_t* aperm = nullptr;   // [table_cap, N] P2P: owned active ranks per (step, set), partner-local first
  void* nccl = nullptr;       // ncclComm_t
  // P2P
  void* peer_base[MAX_RANKS] = {nullptr};
  const double** peer_coords_dev = nullptr;  // device array [nranks]
  unsigned** peer_flags_dev = nullptr;       // device array [nranks]
  unsigned epoch = 0;
  unsigned* done = nullptr;  // CTA-completion counter for the kernel-fused barrier
  bool imported = false;
  std::string err;
};

int comm_unique_id(char* id128);
int comm_init(Comm& c, const char* id128, int rank, int nranks, int mode, int64_t N, int D, double* coords,
              unsigned* flags, size_t table_cap, cudaStream_t st);
int comm_export(Comm& c, char* blob);
int comm_import(Comm& c, const char* blobs);
void comm_destroy(Comm& c);
int comm_probe(Comm& c, int peer, int what, int row_doubles, cudaStream_t st, double* gbs);
// peer pointers / ownership for the kernels
void comm_fill_args(const Comm& c, HalfStepArgs& a);
// [i_lo, i_hi): the active ranks of this split owned by this rank (device resident when nranks > 1)
void comm_active_range(const Comm& c, HalfStepArgs& a, size_t step_in_chunk);
// P2P: all ranks rendezvous before the first split of a call
int comm_begin(Comm& c, cudaStream_t st, int* status, uint64_t& launches);
// P2P + dense_dmma: the barrier rides inside the half-step kernel (wait at its start, signal from
// its last CTA, and between the `nhalf` half-steps of a persistent launch); fills the p2p_* fields and
// advances the epoch by nhalf.  Returns false if not applicable.
bool comm_fuse_barrier(Comm& c, HalfStepArgs& a, int nhalf);
// P2P: an explicit rendezvous on the stream (publish the next epoch, wait for every peer's); needed
// between a kernel that carried the barrier itself and a consumer that does not wait on its own
int comm_barrier(Comm& c, cudaStream_t st, int* status, uint64_t& launches);
// make the rows updated in this split visible to every rank
int comm_after_split(Comm& c, cudaStream_t st, int* status, uint64_t& launches);
// replicate log_prob / accept mask / counters (nacc may be null) -- and coords in P2P mode -- on every rank
int comm_sync_state(Comm& c, cudaStream_t st, int* status, double* logp, uint8_t* accepted,
                    unsigned long long* nacc, uint64_t& launches);
// one in-place all-gather of the owned row blocks of coords (used after a sharded eb_set_state in
// EB_COMM_ALLGATHER mode, where kernels read partner rows from the local replica)
int comm_gather_coords(Comm& c, cudaStream_t st, uint64_t& launches);

}  // namespace eb
