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

constexpr int DMMA_THREADS = 512;  // 16 warps, 4 per SM sub-partition

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__host__ __device__ constexpr int packed_blocks(int KB) { return KB * (KB + 1); }  // 2 * KB(KB+1)/2

template <int KB, bool HAS_MEAN>
__global__ void __launch_bounds__(DMMA_THREADS, 1) half_step_dense_dmma_kernel(const HalfStepArgs a) {
  constexpr int D = 8 * KB;
  extern __shared__ double smem[];
  double* sL = smem;                               // packed_blocks(KB) * 32 doubles
  double* sMu = smem + packed_blocks(KB) * 32;     // D doubles

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;

  // ---- stage the packed factor (and the mean) in shared memory, asynchronously
  {
    const double* src = a.model.chol;
    constexpr int n16 = packed_blocks(KB) * 32 / 2;
    for (int k = tid; k < n16; k += DMMA_THREADS) cp_async16(sL + 2 * k, src + 2 * k);
    if (HAS_MEAN)
      for (int k = tid; k < D; k += DMMA_THREADS) sMu[k] = a.model.params[k];
  }

  const int64_t count = (int64_t)a.i_hi - a.i_lo;
  const int64_t ntiles = (count + 7) >> 3;
  const int64_t Nc = a.N - a.a_count;
  const double dm1 = (double)a.D - 1.0;
  bool staged = false;

  // tiles dealt SM-major: CTA b takes b, b+G, b+2G, ...; its warps take them round-robin
  for (int64_t tile = (int64_t)blockIdx.x + (int64_t)gridDim.x * warp; tile < ntiles;
       tile += (int64_t)gridDim.x * (DMMA_THREADS / 32)) {
    int64_t i = (int64_t)a.i_lo + tile * 8 + g;
    const bool valid = i < a.i_hi;
    if (!valid) i = (int64_t)a.i_hi - 1;

    // ---- draws for this lane's row (stretch.py:30-32, red_blue.py:100) ----
    const u32x4 A = draw_words(a.seed, a.step, (uint32_t)a.split, TAG_PROP_A, (uint32_t)i);
    const double tt = __dadd_rn(__dmul_rn(__dsub_rn(a.p0, 1.0), u53(A.x, A.y)), 1.0);
    const double zz = __ddiv_rn(__dmul_rn(tt, tt), a.p0);
    const int64_t r = (int64_t)bounded64(A.z, A.w, (uint64_t)Nc);
    const int64_t w = a.order[a.a_start + i];
    const int64_t wp = a.order[r < a.a_start ? r : r + a.a_count];
    const double* s_row = a.coords + (size_t)w * D + 2 * t;
    const double* c_row =
        (a.peer_coords != nullptr ? a.peer_coords[wp / a.rows_per_rank] : a.coords) + (size_t)wp * D + 2 * t;

    // ---- gather both rows as 16-byte vectors and form the proposal in registers
    double q[2 * KB];
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      // L2-only (.cg) 16-byte loads: each row is read once per half-step, and
      // a partner row may live in a peer GPU's HBM
      const double2 s2 = __ldcg(reinterpret_cast<const double2*>(s_row + 8 * j));
      const double2 c2 = __ldcg(reinterpret_cast<const double2*>(c_row + 8 * j));
      // stretch.py:33  q = c - (c - s) * zz, each op rounded once
      q[2 * j + 0] = __dsub_rn(c2.x, __dmul_rn(__dsub_rn(c2.x, s2.x), zz));
      q[2 * j + 1] = __dsub_rn(c2.y, __dmul_rn(__dsub_rn(c2.y, s2.y), zz));
    }
    const double lp_old = a.logp[w];
    const u32x4 U = draw_words(a.seed, a.step, (uint32_t)a.split, TAG_ACCEPT, (uint32_t)i);
    const double log_u = log(u53(U.x, U.y));
    const double factor = __dmul_rn(dm1, log(zz));  // stretch.py:31

    if (!staged) {  // first tile of this warp: the factor must have landed
      cp_async_wait_all();
      __syncthreads();
      staged = true;
    }

    // ---- y = L^T (q - mu) block by block on the tensor pipe; rs = sum_n y_n^2
    double rs = 0.0;
    const double* bptr = sL + lane;
#pragma unroll
    for (int nb = 0; nb < KB; ++nb) {
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int j = nb; j < KB; ++j) {
        double x0 = q[2 * j + 0], x1 = q[2 * j + 1];
        if (HAS_MEAN) {
          const double2 m2 = *reinterpret_cast<const double2*>(sMu + 8 * j + 2 * t);
          x0 -= m2.x;
          x1 -= m2.y;
        }
        dmma884(c0, c1, x0, bptr[0]);
        dmma884(c0, c1, x1, bptr[32]);
        bptr += 64;
      }
      rs = fma(c0, c0, rs);
      rs = fma(c1, c1, rs);
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
    const double lnpdiff = __dsub_rn(__dadd_rn(factor, lp_new), lp_old);
    const bool acc = valid && (lnpdiff > log_u);
    if (acc) {
      double* dst = a.coords + (size_t)w * D + 2 * t;
#pragma unroll
      for (int j = 0; j < KB; ++j) *reinterpret_cast<double2*>(dst + 8 * j) = make_double2(q[2 * j], q[2 * j + 1]);
    }
    if (valid && t == 0) {
      if (acc) {
        a.logp[w] = lp_new;
        a.nacc[w] += 1ull;
      }
      a.accepted[w] = acc ? 1 : 0;
    }
  }
  if (!staged) {  // warps without a tile still own part of the async copy
    cp_async_wait_all();
    __syncthreads();
  }
}

template <int KB>
cudaError_t launch_t(const HalfStepArgs& a, int sm_count, cudaStream_t st) {
  const size_t smem = ((size_t)packed_blocks(KB) * 32 + 8 * KB) * sizeof(double);
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
// kernel consumes it: for each 8-column tile nb, for each 8-row group j >= nb,
// two 4x8 fragments (half = 0, 1) whose lane (g, t) element is
// L[8j + 2t + half][8nb + g].
void dense_dmma_pack_factor(const double* L, int D, double* packed) {
  const int KB = D / 8;
  size_t idx = 0;
  for (int nb = 0; nb < KB; ++nb)
    for (int j = nb; j < KB; ++j)
      for (int half = 0; half < 2; ++half)
        for (int lane = 0; lane < 32; ++lane) {
          const int g = lane >> 2, t = lane & 3;
          packed[idx++] = L[(size_t)(8 * j + 2 * t + half) * D + (8 * nb + g)];
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
