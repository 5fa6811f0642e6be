// Fused half-step for StretchMove + dense Gaussian on the FP64 tensor pipe.
//
// Reference semantics: moves/red_blue.py:82-104 with moves/stretch.py:26-33 as the
// proposal and log_prob(x) = -0.5 (x-mu)^T A (x-mu) (document/plots/oned.py:17-18).
//
// Design (DESIGN.md "dense_dmma"):
//   * A = L L^T is factored once on the host; lp = -0.5 |L^T (q - mu)|^2, i.e. a
//     [walkers x D] x [D x D lower-triangular] product: only the blocks on or
//     below the diagonal are multiplied (D(D+1) instead of 2 D^2 flops).
//   * one warp owns a tile of 8 active walkers; mma.sync.m8n8k4.f64 (DMMA).  The
//     A operand (the proposal rows q) lives in REGISTERS for the whole tile --
//     lane (g, t) holds row g, physical columns {8j+2t, 8j+2t+1}, j < D/8, which
//     it reads from HBM/L2 as 16-byte vectors; the contraction index is permuted
//     accordingly when L is packed, which is free.  The same registers are the
//     values written back if the proposal is accepted, so q is formed exactly
//     once (bit-exact sub/mul/sub, no FMA contraction).
//   * L (packed per 4x8 fragment, 256 B per DMMA, conflict-free LDS.64) is staged
//     in shared memory once per CTA by cp.async; CTAs are persistent (one per SM)
//     and tiles are dealt SM-major so every SM sub-partition gets the same count.
//   * per-row |y|^2 is reduced over the 4 lanes of a row with two shuffles.
#include <math.h>

#include "engine.cuh"

namespace eb {

namespace {

constexpr int DMMA_THREADS = 256;  // 8 warps, 2 per SM sub-partition (the tensor pipe needs no more)
constexpr int DMMA_WARPS = DMMA_THREADS / 32;
constexpr int NI = 4;  // column tiles in flight per warp (independent accumulator chains)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__host__ __device__ constexpr int packed_blocks(int KB) { return KB * (KB + 1); }  // 2 * KB(KB+1)/2
// landing buffer of one warp: [s|c][8 rows][D + 8] doubles; the +8 (64 B) row skew
// makes the 16-byte fragment reads of 8 consecutive lanes hit 32 distinct banks
__host__ __device__ constexpr int row_stride(int KB) { return 8 * KB + 8; }
__host__ __device__ constexpr size_t dmma_smem_bytes(int KB) {
  return ((size_t)packed_blocks(KB) * 32 + 8 * KB + (size_t)DMMA_WARPS * 2 * 8 * row_stride(KB)) * sizeof(double);
}

// what a lane must know about one tile before its rows can be fetched.  The two
// walker ids are kept as the raw 32-bit words the index loads return: their first
// consumer is pinned (below) behind the previous tile's tensor-pipe phase, so the
// L2 latency of the order[] lookups is never waited for.
struct TileMeta {
  int32_t w, wp;  // this lane's active walker and its partner (order[] entries)
  double zz, u;   // stretch factor, accept uniform
  bool valid;
};

// an opaque move: volatile asm statements keep their program order, so whatever
// reads the result cannot be scheduled ahead of the DMMA block that precedes it
__device__ __forceinline__ int32_t pin(int32_t x) {
  int32_t y;
  asm volatile("mov.b32 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

template <int KB, bool HAS_MEAN>
__global__ void __launch_bounds__(DMMA_THREADS, 1) half_step_dense_dmma_kernel(const HalfStepArgs a) {
  constexpr int D = 8 * KB;
  constexpr int RS = row_stride(KB);
  extern __shared__ double smem[];
  double* sL = smem;                            // packed_blocks(KB) * 32 doubles
  double* sMu = sL + packed_blocks(KB) * 32;    // D doubles
  double* sRows = sMu + D;                      // DMMA_WARPS * 2 * 8 * RS doubles

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  double* myS = sRows + (size_t)warp * 2 * 8 * RS + (size_t)g * RS + 2 * t;  // this lane's chunks of row g
  double* myC = myS + 8 * RS;

  // ---- stage the packed factor (and the mean) in shared memory, asynchronously
  {
    const double* src = a.model.chol;
    constexpr int n16 = packed_blocks(KB) * 32 / 2;
    for (int k = tid; k < n16; k += DMMA_THREADS) cp_async16(sL + 2 * k, src + 2 * k);
    if (HAS_MEAN)
      for (int k = tid; k < D; k += DMMA_THREADS) sMu[k] = a.model.params[k];
  }

  const int i_lo = a.range ? a.range->x : a.i_lo;
  const int i_hi = a.range ? a.range->y : a.i_hi;
  const int64_t count = (int64_t)i_hi - i_lo;
  const int64_t ntiles = (count + 7) >> 3;
  const int64_t Nc = a.N - a.a_count;
  const double dm1 = (double)a.D - 1.0;
  const int64_t tstride = (int64_t)gridDim.x * DMMA_WARPS;

  // draws + index lookups of one tile (stretch.py:30-32, red_blue.py:82-87,100)
  auto prep = [&](int64_t tile) -> TileMeta {
    TileMeta m;
    int64_t i = (int64_t)i_lo + tile * 8 + g;
    m.valid = i < i_hi;
    if (!m.valid) i = (int64_t)i_hi - 1;
    const u32x4 A = draw_words(a.seed, a.step, (uint32_t)a.split, TAG_PROP_A, (uint32_t)i);
    const double tt = __dadd_rn(__dmul_rn(__dsub_rn(a.p0, 1.0), u53(A.x, A.y)), 1.0);
    m.zz = __ddiv_rn(__dmul_rn(tt, tt), a.p0);
    const int64_t r = (int64_t)bounded64(A.z, A.w, (uint64_t)Nc);
    m.w = __ldg(a.order + a.a_start + i);
    m.wp = __ldg(a.order + (r < a.a_start ? r : r + a.a_count));
    const u32x4 U = draw_words(a.seed, a.step, (uint32_t)a.split, TAG_ACCEPT, (uint32_t)i);
    m.u = u53(U.x, U.y);
    return m;
  };
  // asynchronous gather of this lane's 16-byte chunks of both rows into the landing buffer
  auto fetch = [&](int64_t w, int64_t wp) {
    const double* s_row = a.coords + (size_t)w * D + 2 * t;
    const double* c_row =
        (a.peer_coords != nullptr ? a.peer_coords[wp / a.rows_per_rank] : a.coords) + (size_t)wp * D + 2 * t;
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      cp_async16(myS + 8 * j, s_row + 8 * j);
      cp_async16(myC + 8 * j, c_row + 8 * j);
    }
  };

  // tiles dealt SM-major: CTA b takes b, b+G, b+2G, ...; its warps take them round-robin.
  // Software pipeline per warp: rows of tile k+1 stream into shared memory (LDGSTS, no
  // registers) and the indices of tile k+2 are looked up while tile k is on the tensor pipe.
  int64_t tile = (int64_t)blockIdx.x + (int64_t)gridDim.x * warp;
  TileMeta cur{}, nxt{};
  if (tile < ntiles) {
    cur = prep(tile);
    fetch(cur.w, cur.wp);
    if (tile + tstride < ntiles) nxt = prep(tile + tstride);
  }
  cp_async_wait_all();  // the factor chunks this thread copied (and tile 0's rows)
  __syncthreads();      // ... and everybody else's

  for (; tile < ntiles; tile += tstride) {
    cp_async_wait_all();
    // ---- form the proposal in registers (stretch.py:33, each op rounded once)
    double q[2 * KB];
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      const double2 s2 = *reinterpret_cast<const double2*>(myS + 8 * j);
      const double2 c2 = *reinterpret_cast<const double2*>(myC + 8 * j);
      q[2 * j + 0] = __dsub_rn(c2.x, __dmul_rn(__dsub_rn(c2.x, s2.x), cur.zz));
      q[2 * j + 1] = __dsub_rn(c2.y, __dmul_rn(__dsub_rn(c2.y, s2.y), cur.zz));
    }
    const int64_t w_cur = cur.w;
    const double lp_old = a.logp[w_cur];
    // ---- keep the pipeline full: rows of the next tile, indices of the one after
    const bool has_next = tile + tstride < ntiles;
    TileMeta ready = nxt;
    if (has_next) {
      ready.w = pin(ready.w);  // looked up one tensor-pipe phase ago
      ready.wp = pin(ready.wp);
      fetch(ready.w, ready.wp);
      if (tile + 2 * tstride < ntiles) nxt = prep(tile + 2 * tstride);
    }

    // ---- y = L^T (q - mu) block by block on the tensor pipe; rs = sum_n y_n^2.
    // A dependent DMMA chain issues only every ~64 cycles, the pipe takes one every
    // 16 per sub-partition: NI column tiles x 2 k-halves = 2*NI independent
    // accumulator chains per warp keep it fed with 2 warps per sub-partition.
    double rs = 0.0;
    const double* bptr = sL + lane;
#pragma unroll
    for (int nb0 = 0; nb0 < KB; nb0 += NI) {
      double c[NI][2][2];
#pragma unroll
      for (int n = 0; n < NI; ++n) c[n][0][0] = c[n][0][1] = c[n][1][0] = c[n][1][1] = 0.0;
#pragma unroll
      for (int j = nb0; j < KB; ++j) {
        double x0 = q[2 * j + 0], x1 = q[2 * j + 1];
        if (HAS_MEAN) {
          const double2 m2 = *reinterpret_cast<const double2*>(sMu + 8 * j + 2 * t);
          x0 -= m2.x;
          x1 -= m2.y;
        }
#pragma unroll
        for (int n = 0; n < NI; ++n) {
          if (nb0 + n < KB && j >= nb0 + n) {
            dmma884(c[n][0][0], c[n][0][1], x0, bptr[0]);
            dmma884(c[n][1][0], c[n][1][1], x1, bptr[32]);
            bptr += 64;
          }
        }
      }
#pragma unroll
      for (int n = 0; n < NI; ++n) {
        const double y0 = c[n][0][0] + c[n][1][0], y1 = c[n][0][1] + c[n][1][1];
        rs = fma(y0, y0, rs);
        rs = fma(y1, y1, rs);
      }
    }
    rs += __shfl_xor_sync(0xffffffffu, rs, 1);
    rs += __shfl_xor_sync(0xffffffffu, rs, 2);
    const double lp_new = -0.5 * rs;

    // ---- guards (ensemble.py:476-479, 550-551): a non-finite lp is the only way
    // a non-finite coordinate can show, so the element scan is off the fast path
    if (!isfinite(lp_new)) {
      bool any_inf = false, any_nan = false;
#pragma unroll
      for (int k = 0; k < 2 * KB; ++k) {
        any_inf |= isinf(q[k]);
        any_nan |= isnan(q[k]);
      }
      if (any_inf) atomicOr(a.status, FLAG_INF_PARAM);
      if (any_nan) atomicOr(a.status, FLAG_NAN_PARAM);
      if (isnan(lp_new)) atomicOr(a.status, FLAG_NAN_LOGPROB);
    }

    // ---- Metropolis accept + in-place update (red_blue.py:96-104, move.py:29-34)
    const double factor = __dmul_rn(dm1, log(cur.zz));  // stretch.py:31
    const double lnpdiff = __dsub_rn(__dadd_rn(factor, lp_new), lp_old);
    const bool acc = cur.valid && (lnpdiff > log(cur.u));
    if (acc) {
      double* dst = a.coords + (size_t)w_cur * D + 2 * t;
#pragma unroll
      for (int j = 0; j < KB; ++j) *reinterpret_cast<double2*>(dst + 8 * j) = make_double2(q[2 * j], q[2 * j + 1]);
    }
    if (cur.valid && t == 0) {
      if (acc) {
        a.logp[w_cur] = lp_new;
        atomicAdd(a.nacc + w_cur, 1ull);  // RED: fire and forget, no load to wait for
      }
      a.accepted[w_cur] = acc ? 1 : 0;
    }
    cur = ready;
  }
}

template <int KB>
cudaError_t launch_t(const HalfStepArgs& a, int sm_count, cudaStream_t st) {
  const size_t smem = dmma_smem_bytes(KB);
  const bool has_mean = a.model.s0 != 0.0;  // set by eb_model_set when mu != 0
  auto kern = has_mean ? half_step_dense_dmma_kernel<KB, true> : half_step_dense_dmma_kernel<KB, false>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int64_t count = (int64_t)a.i_hi - a.i_lo;
  if (count <= 0) return cudaSuccess;
  const int64_t ntiles = (count + 7) / 8;
  const int grid = (int)(ntiles < sm_count ? ntiles : sm_count);
  kern<<<grid, DMMA_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace

bool dense_dmma_supported(int D) {
  switch (D) {
    case 16:
    case 32:
    case 48:
    case 64:
    case 96:
    case 128:
      return true;
  }
  return false;
}

size_t dense_dmma_factor_doubles(int D) { return (size_t)packed_blocks(D / 8) * 32; }

// L: row-major lower-triangular factor (A = L L^T).  Packed in the order the
// kernel consumes it: for each group of NI 8-column tiles, for each 8-row group
// j, for each tile nb of the group with j >= nb, two 4x8 fragments (half = 0, 1)
// whose lane (g, t) element is L[8j + 2t + half][8nb + g].
void dense_dmma_pack_factor(const double* L, int D, double* packed) {
  const int KB = D / 8;
  size_t idx = 0;
  for (int nb0 = 0; nb0 < KB; nb0 += NI)
    for (int j = nb0; j < KB; ++j)
      for (int n = 0; n < NI; ++n) {
        const int nb = nb0 + n;
        if (nb >= KB || j < nb) continue;
        for (int half = 0; half < 2; ++half)
          for (int lane = 0; lane < 32; ++lane) {
            const int g = lane >> 2, t = lane & 3;
            packed[idx++] = L[(size_t)(8 * j + 2 * t + half) * D + (8 * nb + g)];
          }
      }
}

cudaError_t launch_half_step_dense_dmma(const HalfStepArgs& a, int sm_count, cudaStream_t st) {
  switch (a.D) {
    case 16:
      return launch_t<2>(a, sm_count, st);
    case 32:
      return launch_t<4>(a, sm_count, st);
    case 48:
      return launch_t<6>(a, sm_count, st);
    case 64:
      return launch_t<8>(a, sm_count, st);
    case 96:
      return launch_t<12>(a, sm_count, st);
    case 128:
      return launch_t<16>(a, sm_count, st);
  }
  return cudaErrorNotSupported;
}

}  // namespace eb
