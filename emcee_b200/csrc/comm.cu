// Multi-GPU plumbing (see comm.h): one process per GPU, walkers sharded by row block.
//
// The reference's only parallel construct is pool.map over the proposals of one
// split (ensemble.py:492-496); what makes that legal -- the complement is frozen
// while a split is updated (red_blue.py:85-104) -- is what makes row-block
// sharding legal here: a rank updates the active walkers it owns and only READS
// complement rows, which other ranks do not write during the split.
//
//   EB_COMM_ALLGATHER  every rank holds a replica of coords; after each split the
//                      owned row blocks are all-gathered in place (ncclAllGather,
//                      NCCL dlopen-ed: single-GPU use never loads it).
//   EB_COMM_P2P        no replica traffic: the half-step kernel reads partner rows
//                      straight from the owner's HBM over NVLink through
//                      cudaIpc-mapped pointers (HalfStepArgs::peer_coords), and the
//                      ranks meet at a flag barrier in peer memory between splits.
#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>  // types only; every entry point is resolved with dlsym
#include <stdio.h>
#include <string.h>

namespace eb {

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi& nccl_api() {
  static NcclApi api;
  if (api.lib || !api.err.empty()) return api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) {
    api.err = std::string("cannot load libnccl.so.2: ") + dlerror();
    return api;
  }
#define LOAD(field, sym)                                                    \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym));   \
  if (!api.field) api.err = std::string("libnccl lacks ") + sym;
  LOAD(GetUniqueId, "ncclGetUniqueId")
  LOAD(CommInitRank, "ncclCommInitRank")
  LOAD(AllGather, "ncclAllGather")
  LOAD(CommDestroy, "ncclCommDestroy")
  LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
  return api;
}

static_assert(sizeof(ncclUniqueId) == EB_COMM_ID_BYTES, "EB_COMM_ID_BYTES must match ncclUniqueId");

struct IpcBlob {
  cudaIpcMemHandle_t mem;  // the coords allocation (flags live in its tail)
  int32_t rank;
  int32_t pad;
};
static_assert(sizeof(IpcBlob) <= EB_IPC_BLOB_BYTES, "IPC blob too large");

// cross-GPU barrier: publish my epoch into every peer's flag array, then wait until
// every peer has published theirs into mine.  Launched on the engine's stream after
// the half-step kernel, so all of that kernel's writes precede the release.
__global__ void p2p_barrier_kernel(unsigned* const* peer_flags, volatile unsigned* my_flags, int rank, int nranks,
                                   unsigned epoch, int* status) {
  const int t = threadIdx.x;
  if (t < nranks) {
    __threadfence_system();
    atomicExch_system(peer_flags[t] + rank, epoch);
    const long long t0 = clock64();
    while ((int)(my_flags[t] - epoch) < 0) {
      __nanosleep(64);
      if (clock64() - t0 > 60000000000ll) {  // ~30 s: a peer died; report instead of hanging the GPU
        atomicOr(status, FLAG_COMM_TIMEOUT);
        break;
      }
    }
    __threadfence_system();
  }
}

// ---- peer-memory read probes (eb_comm_probe): what NVLink delivers for our access patterns ----
__global__ void probe_stream_kernel(const double2* __restrict__ src, size_t n16, double* sink) {
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const double2 v = __ldcg(src + i);
    acc += v.x + v.y;
  }
  if (acc == 123.456) *sink = acc;
}
// each warp reads whole rows of `row16` 16-byte chunks at pseudo-random row indices
__global__ void probe_rows_kernel(const double2* __restrict__ src, size_t nrows, int row16, int rows_per_warp,
                                  double* sink) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double acc = 0.0;
  for (int k = 0; k < rows_per_warp; ++k) {
    const size_t r = (size_t)fmix32((uint32_t)(warp * 7919u + k * 104729u + 17u)) % nrows;
    for (int c = lane; c < row16; c += 32) {
      const double2 v = __ldcg(src + r * row16 + c);
      acc += v.x + v.y;
    }
  }
  if (acc == 123.456) *sink = acc;
}
// same rows through TMA bulk copies into shared memory (what dense_dmma's producers do)
__global__ void probe_bulk_kernel(const double* __restrict__ src, size_t nrows, int row_doubles, int rows_per_warp,
                                  double* sink) {
  extern __shared__ __align__(16) unsigned char psm[];
  const int lane = threadIdx.x & 31, warp_in = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  double* buf = reinterpret_cast<double*>(psm) + (size_t)warp_in * 16 * row_doubles;
  uint64_t* bar = reinterpret_cast<uint64_t*>(psm + (size_t)nw * 16 * row_doubles * sizeof(double)) + warp_in;
  const unsigned bar_s = (unsigned)__cvta_generic_to_shared(bar);
  if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const size_t warp = (size_t)blockIdx.x * nw + warp_in;
  const unsigned bytes = (unsigned)(row_doubles * sizeof(double));
  for (int k = 0; k < rows_per_warp / 16; ++k) {
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(16u * bytes) : "memory");
    __syncwarp();
    if (lane < 16) {
      const size_t r = (size_t)fmix32((uint32_t)(warp * 7919u + (k * 16 + lane) * 104729u + 17u)) % nrows;
      const unsigned dst = (unsigned)__cvta_generic_to_shared(buf + (size_t)lane * row_doubles);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "l"(src + r * row_doubles), "r"(bytes), "r"(bar_s)
                   : "memory");
    }
    unsigned ok = 0;
    while (!ok) {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(ok)
          : "r"(bar_s), "r"((unsigned)(k & 1))
          : "memory");
    }
  }
  if (buf[lane] == 123.456) *sink = buf[lane];
}

int fail(Comm& c, const std::string& msg) {
  c.err = msg;
  return 1;
}

#define CCK(c, call)                                                                        \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      cudaGetLastError();                                                                   \
      return fail(c, std::string(#call) + " failed: " + cudaGetErrorString(_e));            \
    }                                                                                       \
  } while (0)

#define NCK(c, call)                                                                        \
  do {                                                                                      \
    ncclResult_t _r = (call);                                                               \
    if (_r != ncclSuccess) return fail(c, std::string(#call) + " failed: " + nccl_api().GetErrorString(_r)); \
  } while (0)

}  // namespace

int comm_unique_id(char* id128) {
  NcclApi& api = nccl_api();
  if (!api.err.empty()) return 1;
  ncclUniqueId id;
  if (api.GetUniqueId(&id) != ncclSuccess) return 1;
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int comm_init(Comm& c, const char* id128, int rank, int nranks, int mode, int64_t N, int D, double* coords,
              unsigned* flags, size_t table_cap, cudaStream_t st) {
  if (nranks < 1 || nranks > MAX_RANKS || rank < 0 || rank >= nranks) return fail(c, "comm_init: bad rank / nranks");
  if (N % nranks != 0) return fail(c, "comm_init: nwalkers must be divisible by the number of ranks");
  if (mode != EB_COMM_ALLGATHER && mode != EB_COMM_P2P) return fail(c, "comm_init: unknown mode");
  c.rank = rank;
  c.nranks = nranks;
  c.mode = mode;
  c.N = N;
  c.D = D;
  c.coords = coords;
  c.flags = flags;
  c.rows_per_rank = N / nranks;
  if (nranks == 1) return 0;
  NcclApi& api = nccl_api();
  if (!api.err.empty()) return fail(c, api.err);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  NCK(c, api.CommInitRank(&comm, nranks, id, rank));
  c.nccl = comm;
  CCK(c, cudaMalloc(&c.ranges, table_cap * MAX_SPLITS * sizeof(int2)));
  if (mode == EB_COMM_P2P) CCK(c, cudaMalloc(&c.aperm, table_cap * (size_t)N * sizeof(int32_t)));
  CCK(c, cudaMalloc(&c.done, sizeof(unsigned)));
  CCK(c, cudaMemsetAsync(c.done, 0, sizeof(unsigned), st));
  CCK(c, cudaMemsetAsync(c.flags, 0, MAX_RANKS * sizeof(unsigned), st));
  CCK(c, cudaStreamSynchronize(st));
  return 0;
}

int comm_export(Comm& c, char* blob) {
  memset(blob, 0, EB_IPC_BLOB_BYTES);
  IpcBlob b;
  memset(&b, 0, sizeof(b));
  CCK(c, cudaIpcGetMemHandle(&b.mem, c.coords));
  b.rank = c.rank;
  memcpy(blob, &b, sizeof(b));
  return 0;
}

int comm_import(Comm& c, const char* blobs) {
  if (c.nranks == 1) return 0;
  const double* coords_host[MAX_RANKS];
  unsigned* flags_host[MAX_RANKS];
  const size_t flag_off = (size_t)c.N * c.D;  // flags sit right behind the coords in the same allocation
  for (int r = 0; r < c.nranks; ++r) {
    IpcBlob b;
    memcpy(&b, blobs + (size_t)r * EB_IPC_BLOB_BYTES, sizeof(b));
    if (b.rank != r) return fail(c, "comm_import: blobs are not ordered by rank");
    void* base = nullptr;
    if (r == c.rank) {
      base = c.coords;
    } else {
      CCK(c, cudaIpcOpenMemHandle(&base, b.mem, cudaIpcMemLazyEnablePeerAccess));
      c.peer_base[r] = base;
    }
    coords_host[r] = static_cast<const double*>(base);
    flags_host[r] = reinterpret_cast<unsigned*>(static_cast<double*>(base) + flag_off);
  }
  CCK(c, cudaMalloc(&c.peer_coords_dev, c.nranks * sizeof(double*)));
  CCK(c, cudaMalloc(&c.peer_flags_dev, c.nranks * sizeof(unsigned*)));
  CCK(c, cudaMemcpy(c.peer_coords_dev, coords_host, c.nranks * sizeof(double*), cudaMemcpyHostToDevice));
  CCK(c, cudaMemcpy(c.peer_flags_dev, flags_host, c.nranks * sizeof(unsigned*), cudaMemcpyHostToDevice));
  c.imported = true;
  return 0;
}

// GB/s of reading `peer`'s coords buffer (peer == own rank: local HBM) with pattern `what`:
// 0 streaming 16-byte loads, 1 random whole rows with 16-byte loads, 2 random whole rows with TMA bulk copies
int comm_probe(Comm& c, int peer, int what, int row_doubles, cudaStream_t st, double* gbs) {
  if (peer < 0 || peer >= c.nranks) return fail(c, "comm_probe: bad peer");
  if (peer != c.rank && !c.imported) return fail(c, "comm_probe: peer memory not imported");
  const double* src = peer == c.rank ? c.coords : static_cast<const double*>(c.peer_base[peer]);
  const size_t total = (size_t)c.N * c.D;  // doubles
  double* sink = nullptr;
  CCK(c, cudaMalloc(&sink, 8));
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double bytes = 0;
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, st);
    if (what == 0) {
      probe_stream_kernel<<<sms * 8, 512, 0, st>>>(reinterpret_cast<const double2*>(src), total / 2, sink);
      bytes = (double)total * 8;
    } else {
      const size_t nrows = total / row_doubles;
      const int rows_per_warp = 64, warps = sms * 16;
      if (what == 1) {
        probe_rows_kernel<<<sms * 2, 256, 0, st>>>(reinterpret_cast<const double2*>(src), nrows, row_doubles / 2,
                                                  rows_per_warp, sink);
      } else {
        const size_t smem = (size_t)8 * 16 * row_doubles * sizeof(double) + 8 * sizeof(uint64_t);
        cudaFuncSetAttribute(probe_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe_bulk_kernel<<<sms * 2, 256, smem, st>>>(src, nrows, row_doubles, rows_per_warp, sink);
      }
      bytes = (double)warps * rows_per_warp * row_doubles * 8;
    }
    cudaEventRecord(e1, st);
    CCK(c, cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  CCK(c, cudaGetLastError());
  *gbs = bytes / (best * 1e-3) / 1e9;
  return 0;
}

void comm_destroy(Comm& c) {
  for (int r = 0; r < MAX_RANKS; ++r)
    if (c.peer_base[r]) cudaIpcCloseMemHandle(c.peer_base[r]);
  cudaFree(c.peer_coords_dev);
  cudaFree(c.peer_flags_dev);
  cudaFree(c.ranges);
  cudaFree(c.aperm);
  cudaFree(c.done);
  if (c.nccl) nccl_api().CommDestroy(static_cast<ncclComm_t>(c.nccl));
  c = Comm();
}

void comm_fill_args(const Comm& c, HalfStepArgs& a) {
  a.rows_per_rank = c.rows_per_rank > 0 ? c.rows_per_rank : a.N;
  a.peer_coords = (c.nranks > 1 && c.mode == EB_COMM_P2P) ? c.peer_coords_dev : nullptr;
}

void comm_active_range(const Comm& c, HalfStepArgs& a, size_t step_in_chunk) {
  a.i_lo = 0;
  if (c.nranks == 1) {
    a.i_hi = a.a_count;
    a.range = nullptr;
    return;
  }
  // the exact range is device resident (split_table_kernel); the host only needs a bound for the grid
  a.i_hi = (int)(a.a_count < c.rows_per_rank ? a.a_count : c.rows_per_rank);
  a.range = c.ranges + step_in_chunk * MAX_SPLITS + a.split;
}

static int p2p_barrier(Comm& c, cudaStream_t st, int* status, uint64_t& launches) {
  if (!c.imported) return fail(c, "EB_COMM_P2P: eb_comm_import has not been called");
  c.epoch += 1;
  p2p_barrier_kernel<<<1, 32, 0, st>>>(c.peer_flags_dev, c.flags, c.rank, c.nranks, c.epoch, status);
  CCK(c, cudaGetLastError());
  ++launches;
  return 0;
}

bool comm_fuse_barrier(Comm& c, HalfStepArgs& a, int nhalf) {
  if (c.nranks == 1 || c.mode != EB_COMM_P2P || !c.imported) return false;
  a.p2p_peer_flags = c.peer_flags_dev;
  a.p2p_my_flags = c.flags;
  a.p2p_done = c.done;
  a.p2p_rank = c.rank;
  a.p2p_nranks = c.nranks;
  a.p2p_wait = c.epoch;        // everybody has finished the previous barrier event
  a.p2p_signal = c.epoch + 1;  // ... the completion of this launch's half-step h is event epoch + 1 + h
  c.epoch += (unsigned)nhalf;
  return true;
}

int comm_barrier(Comm& c, cudaStream_t st, int* status, uint64_t& launches) {
  if (c.nranks == 1 || c.mode != EB_COMM_P2P) return 0;
  return p2p_barrier(c, st, status, launches);
}

int comm_begin(Comm& c, cudaStream_t st, int* status, uint64_t& launches) {
  if (c.nranks == 1 || c.mode != EB_COMM_P2P) return 0;
  return p2p_barrier(c, st, status, launches);
}

int comm_after_split(Comm& c, cudaStream_t st, int* status, uint64_t& launches) {
  if (c.nranks == 1) return 0;
  if (c.mode == EB_COMM_ALLGATHER) {
    const size_t cnt = (size_t)c.rows_per_rank * c.D;
    NCK(c, nccl_api().AllGather(c.coords + (size_t)c.rank * cnt, c.coords, cnt, ncclDouble,
                                static_cast<ncclComm_t>(c.nccl), st));
    ++launches;
    return 0;
  }
  return p2p_barrier(c, st, status, launches);
}

int comm_sync_state(Comm& c, cudaStream_t st, int* status, double* logp, uint8_t* accepted,
                    unsigned long long* nacc, uint64_t& launches) {
  if (c.nranks == 1) return 0;
  NcclApi& api = nccl_api();
  ncclComm_t comm = static_cast<ncclComm_t>(c.nccl);
  const size_t R = (size_t)c.rows_per_rank;
  if (c.mode == EB_COMM_P2P) {  // replicas were not maintained during the run
    // nobody may still be pulling rows of the last half-step when the gathers start rewriting replicas
    if (p2p_barrier(c, st, status, launches)) return 1;
    NCK(c, api.AllGather(c.coords + (size_t)c.rank * R * c.D, c.coords, R * c.D, ncclDouble, comm, st));
    ++launches;
  }
  NCK(c, api.AllGather(logp + c.rank * R, logp, R, ncclDouble, comm, st));
  NCK(c, api.AllGather(accepted + c.rank * R, accepted, R, ncclUint8, comm, st));
  launches += 2;
  if (nacc) {
    NCK(c, api.AllGather(nacc + c.rank * R, nacc, R, ncclUint64, comm, st));
    ++launches;
  }
  if (c.mode == EB_COMM_P2P) return p2p_barrier(c, st, status, launches);  // gathers landed everywhere
  return 0;
}

int comm_gather_coords(Comm& c, cudaStream_t st, uint64_t& launches) {
  if (c.nranks == 1) return 0;
  const size_t cnt = (size_t)c.rows_per_rank * c.D;
  NCK(c, nccl_api().AllGather(c.coords + (size_t)c.rank * cnt, c.coords, cnt, ncclDouble,
                              static_cast<ncclComm_t>(c.nccl), st));
  ++launches;
  return 0;
}

}  // namespace eb
