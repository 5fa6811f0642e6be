// Device micro-benchmarks used to anchor the FP64 roofline (MEASURED_PEAKS.json
// only carries HBM copy and bf16 GEMM peaks): DFMA (CUDA-core fp64) and DMMA
// (fp64 tensor core, mma.sync m8n8k4 / m16n8k8 / m16n8k16) issue rates, and a
// plain HBM copy.  Exposed as eb_microbench(); never on the product path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/emcee_b200.h"

namespace {

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
      "{%12,%13,%14,%15}, {%0,%1,%2,%3};"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
        "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double x, double y) {
  double acc[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) acc[k] = (double)(threadIdx.x + k);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = fma(acc[k], x, y);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma884_kernel(double* out, int iters, double x, double y) {
  double d0[ILP], d1[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) d0[k] = d1[k] = (double)k;
  const double a = x + threadIdx.x * 1e-9, b = y;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) dmma884(d0[k], d1[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += d0[k] + d1[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma1688_kernel(double* out, int iters, double x, double y) {
  double d[ILP][4];
#pragma unroll
  for (int k = 0; k < ILP; ++k)
    for (int j = 0; j < 4; ++j) d[k][j] = (double)(k + j);
  double a[4] = {x, x + 1e-9 * threadIdx.x, x, x};
  double b[2] = {y, y};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) dmma1688(d[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += d[k][0] + d[k][1] + d[k][2] + d[k][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma16816_kernel(double* out, int iters, double x, double y) {
  double d[ILP][4];
#pragma unroll
  for (int k = 0; k < ILP; ++k)
    for (int j = 0; j < 4; ++j) d[k][j] = (double)(k + j);
  double a[8] = {x, x + 1e-9 * threadIdx.x, x, x, x, x, x, x};
  double b[4] = {y, y, y, y};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) dmma16816(d[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += d[k][0] + d[k][1] + d[k][2] + d[k][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the consumer loop of dense_dmma.cu in isolation: B fragments streamed from shared
// memory (one LDS per DMMA), A fragments from 32 different registers, CH independent
// accumulator chains
template <int CH>
__global__ void dmma884_smem_kernel(double* out, int iters, double x) {
  extern __shared__ double sb[];
  for (int k = threadIdx.x; k < 8704; k += blockDim.x) sb[k] = 1e-9 * k;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double a[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) a[k] = x + 1e-9 * (k + lane);
  double c0[CH], c1[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) c0[k] = c1[k] = 0.0;
  for (int it = 0; it < iters; ++it) {
    const double* bp = sb + lane;
#pragma unroll
    for (int j = 0; j < 256; ++j) {
      dmma884(c0[j % CH], c1[j % CH], a[j % 32], bp[32 * j]);
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CH; ++k) s += c0[k] + c1[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// same consumer loop with the B fragments double-buffered in registers: the G 16-byte
// loads of group i+1 are issued before the 2G DMMAs of group i
template <int G>
__global__ void dmma884_smem_pf_kernel(double* out, int iters, double x) {
  extern __shared__ double sb[];
  for (int k = threadIdx.x; k < 8704; k += blockDim.x) sb[k] = 1e-9 * k;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double a[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) a[k] = x + 1e-9 * (k + lane);
  double c0[4], c1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) c0[k] = c1[k] = 0.0;
  constexpr int NG = 128 / G;  // groups per pass (128 LDS.128 = 256 DMMAs per pass)
  for (int it = 0; it < iters; ++it) {
    const double2* bp = reinterpret_cast<const double2*>(sb) + lane;
    double2 cur[G], nxt[G];
#pragma unroll
    for (int k = 0; k < G; ++k) cur[k] = bp[32 * k];
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
      if (gi + 1 < NG) {
#pragma unroll
        for (int k = 0; k < G; ++k) nxt[k] = bp[32 * ((gi + 1) * G + k)];
      }
#pragma unroll
      for (int k = 0; k < G; ++k) {
        const int j = gi * G + k;
        dmma884(c0[(2 * j) % 4], c1[(2 * j) % 4], a[(2 * j) % 32], cur[k].x);
        dmma884(c0[(2 * j + 1) % 4], c1[(2 * j + 1) % 4], a[(2 * j + 1) % 32], cur[k].y);
      }
#pragma unroll
      for (int k = 0; k < G; ++k) cur[k] = nxt[k];
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) s += c0[k] + c1[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void dmma1688_smem_kernel(double* out, int iters, double x) {
  extern __shared__ double sb[];
  for (int k = threadIdx.x; k < 8704; k += blockDim.x) sb[k] = 1e-9 * k;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double a[16][4];
#pragma unroll
  for (int k = 0; k < 16; ++k)
    for (int m = 0; m < 4; ++m) a[k][m] = x + 1e-9 * (k + m + lane);
  double c[CH][4];
#pragma unroll
  for (int k = 0; k < CH; ++k) c[k][0] = c[k][1] = c[k][2] = c[k][3] = 0.0;
  for (int it = 0; it < iters; ++it) {
    const double* bp = sb + 2 * lane;
#pragma unroll
    for (int j = 0; j < 128; ++j) {
      const double2 b2 = *reinterpret_cast<const double2*>(bp + 64 * j);
      const double b[2] = {b2.x, b2.y};
      dmma1688(c[j % CH], a[j % 16], b);
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < CH; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

template <class F>
double time_ms(F&& launch, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}

}  // namespace

extern "C" int eb_microbench(int what, int warps_per_sm, double* result) {
  if (!result) return EB_ERR_INVALID;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return EB_ERR_CUDA;
  const int sms = prop.multiProcessorCount;
  if (warps_per_sm <= 0) warps_per_sm = 16;
  const int threads = 32 * (warps_per_sm > 32 ? 32 : warps_per_sm);
  const int blocks = sms * (warps_per_sm > 32 ? warps_per_sm / 32 : 1);
  const int iters = 4096;
  double* out = nullptr;
  if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return EB_ERR_CUDA;
  constexpr int ILP = 8;
  const double nwarps = (double)blocks * threads / 32.0;
  double ms = 0, flops = 0;
  switch (what) {
    case 0:  // DFMA
      ms = time_ms([&] { dfma_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
      flops = nwarps * 32.0 * ILP * iters * 2.0;
      break;
    case 1:  // DMMA m8n8k4: 8*8*4*2 flop per warp instruction
      ms = time_ms([&] { dmma884_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
      flops = nwarps * ILP * iters * 512.0;
      break;
    case 2:  // DMMA m16n8k8
      ms = time_ms([&] { dmma1688_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
      flops = nwarps * ILP * iters * 2048.0;
      break;
    case 3:  // DMMA m16n8k16
      ms = time_ms([&] { dmma16816_kernel<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
      flops = nwarps * ILP * iters * 4096.0;
      break;
    case 5:  // DMMA m8n8k4, B from shared memory, 8 chains (dense_dmma consumer loop)
    case 6:  // same, 4 chains
    case 7:  // DMMA m16n8k8, B from shared memory, 4 chains
    case 8: {  // same, 2 chains
      const int it2 = 64;
      const size_t smem = 8704 * sizeof(double);
      cudaFuncSetAttribute(dmma884_smem_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(dmma884_smem_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(dmma1688_smem_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(dmma1688_smem_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (what == 5)
        ms = time_ms([&] { dmma884_smem_kernel<8><<<blocks, threads, smem>>>(out, it2, 1.0000001); }, 5);
      else if (what == 6)
        ms = time_ms([&] { dmma884_smem_kernel<4><<<blocks, threads, smem>>>(out, it2, 1.0000001); }, 5);
      else if (what == 7)
        ms = time_ms([&] { dmma1688_smem_kernel<4><<<blocks, threads, smem>>>(out, it2, 1.0000001); }, 5);
      else
        ms = time_ms([&] { dmma1688_smem_kernel<2><<<blocks, threads, smem>>>(out, it2, 1.0000001); }, 5);
      flops = nwarps * it2 * (what <= 6 ? 256 * 512.0 : 128 * 2048.0);
      break;
    }
    case 9:     // DMMA m8n8k4, B from smem, double-buffered 4 x LDS.128 ahead
    case 10:    // ... 8 ahead
    case 11: {  // ... 16 ahead
      const int it2 = 64;
      const size_t smem = 8704 * sizeof(double);
      cudaFuncSetAttribute(dmma884_smem_pf_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(dmma884_smem_pf_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(dmma884_smem_pf_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (what == 9)
        ms = time_ms([&] { dmma884_smem_pf_kernel<4><<<blocks, threads, smem>>>(out, it2, 1.0000001); }, 5);
      else if (what == 10)
        ms = time_ms([&] { dmma884_smem_pf_kernel<8><<<blocks, threads, smem>>>(out, it2, 1.0000001); }, 5);
      else
        ms = time_ms([&] { dmma884_smem_pf_kernel<16><<<blocks, threads, smem>>>(out, it2, 1.0000001); }, 5);
      flops = nwarps * it2 * 256 * 512.0;
      break;
    }
    case 4: {  // HBM copy, 1 GiB read + 1 GiB write
      const size_t n = (size_t)1 << 26;  // double2 elements = 1 GiB
      double2 *a = nullptr, *b = nullptr;
      if (cudaMalloc(&a, n * 16) != cudaSuccess || cudaMalloc(&b, n * 16) != cudaSuccess) {
        cudaFree(a);
        cudaFree(out);
        return EB_ERR_CUDA;
      }
      cudaMemset(a, 1, n * 16);
      ms = time_ms([&] { copy_kernel<<<sms * 16, 512>>>(a, b, n); }, 5);
      cudaFree(a);
      cudaFree(b);
      cudaFree(out);
      *result = 2.0 * (double)n * 16.0 / (ms * 1e-3) / 1e9;  // GB/s
      return cudaGetLastError() == cudaSuccess ? EB_OK : EB_ERR_CUDA;
    }
    default:
      cudaFree(out);
      return EB_ERR_INVALID;
  }
  cudaFree(out);
  if (cudaGetLastError() != cudaSuccess) return EB_ERR_CUDA;
  *result = flops / (ms * 1e-3) / 1e12;  // TFLOP/s
  return EB_OK;
}
