// microbench.cu — measured roofs for the similarity kernels and the streamed path ("report against the
// box's measured L2 read bandwidth").  MEASURED_PEAKS.json only holds the HBM copy figure; the linear memories the
// similarity kernels gather from are L2/L1-resident, so their denominators are measured here, on the device the bench
// runs on, with the same load instructions the kernels use (128-bit read-only loads):
//   LMB200_MB_L2_READ : every CTA streams 16 B per thread through a buffer that fits L2 but not L1 (ld.global.cg: L1 bypassed)
//   LMB200_MB_L1_READ : every CTA re-reads its own L1-resident window (ld.global.nc, L1 allocating): the L1 data-pipe roof
//   LMB200_MB_HBM_READ: the same stream over a buffer far larger than L2
//   LMB200_MB_H2D     : cudaMemcpyAsync from pinned host memory (the e2e path's roof; run on N ranks at once for N>1)
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lmb200.h"

namespace {

// Each thread sums uint4 loads.  Four independent loads are in flight per thread and the index arithmetic is 32-bit
// (per_cta16 is a multiple of 256 * 4), so the loop is bounded by the load path, not by address computation.
template <int MODE>  // 0: L2 (ld.cg), 1: L1 (ld.nc)
__global__ void __launch_bounds__(256) read_kernel(const uint4* __restrict__ buf, unsigned span16, unsigned per_cta16, int reps,
                                                   unsigned* __restrict__ sink) {
  unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  const unsigned base = (unsigned)(((unsigned long long)blockIdx.x * per_cta16) % span16);
  for (int r = 0; r < reps; ++r) {
    for (unsigned i = threadIdx.x; i < per_cta16; i += 256 * 4) {
      uint4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unsigned idx = base + i + 256u * k;
        if (idx >= span16) idx -= span16;
        v[k] = MODE == 0 ? __ldcg(buf + idx) : __ldg(buf + idx);
      }
      a0 += v[0].x ^ v[0].w; a1 += v[1].y ^ v[1].z; a2 += v[2].x ^ v[2].y; a3 += v[3].z ^ v[3].w;
    }
  }
  const unsigned acc = a0 ^ a1 ^ a2 ^ a3;
  if (acc == 0x9E3779B9u) sink[0] = acc;  // never true for the zero-filled buffer: keeps the loads alive
}

// The plain form of the same stream: one load per iteration, compiler-unrolled by four.  Which of the two forms is
// faster differs between the L2 and the L1 test (and between driver versions): a roof is the BEST any of them reaches.
template <int MODE>
__global__ void __launch_bounds__(256) read_kernel_plain(const uint4* __restrict__ buf, size_t span16, size_t per_cta16, int reps,
                                                         unsigned* __restrict__ sink) {
  unsigned acc = 0;
  const size_t base = ((size_t)blockIdx.x * per_cta16) % span16;
  for (int r = 0; r < reps; ++r) {
#pragma unroll 4
    for (size_t i = threadIdx.x; i < per_cta16; i += 256) {
      size_t k = base + i;
      if (k >= span16) k -= span16;
      const uint4 v = MODE == 0 ? __ldcg(buf + k) : __ldg(buf + k);
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  }
  if (acc == 0x9E3779B9u) sink[0] = acc;
}

}  // namespace

extern "C" int lmb200_microbench(int kind, size_t bytes, int iters, double* gbps) {
  if (!gbps || iters < 1) return LMB200_E_INVALID;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return LMB200_E_NODEVICE; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best_ms = 1e30f;
  double moved = 0;
  int rc = LMB200_OK;
  if (kind == LMB200_MB_H2D) {
    if (bytes == 0) bytes = (size_t)96 * 640 * 480 * 5;  // one bench step of frames
    void *h = nullptr, *d = nullptr;
    if (cudaHostAlloc(&h, bytes, cudaHostAllocDefault) != cudaSuccess || cudaMalloc(&d, bytes) != cudaSuccess) rc = LMB200_E_CUDA;
    if (rc == LMB200_OK) {
      for (size_t i = 0; i < bytes; i += 4096) ((volatile char*)h)[i] = 1;  // first touch on this thread's NUMA node
      cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice);
      cudaEventRecord(a);
      for (int it = 0; it < iters; ++it) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, 0);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      cudaEventElapsedTime(&best_ms, a, b);
      moved = (double)bytes * iters;
    }
    if (h) cudaFreeHost(h);
    if (d) cudaFree(d);
  } else if (kind == LMB200_MB_L2_READ || kind == LMB200_MB_L1_READ || kind == LMB200_MB_HBM_READ) {
    if (bytes == 0) bytes = kind == LMB200_MB_HBM_READ ? ((size_t)2 << 30) : kind == LMB200_MB_L2_READ ? ((size_t)48 << 20) : ((size_t)16 << 10);
    const int ctas = sms * 8;
    // L1: the buffer is `bytes` PER CTA window (all CTAs of an SM share windows of the same small buffer set)
    const size_t total = kind == LMB200_MB_L1_READ ? bytes * (size_t)ctas : bytes;
    const size_t span16 = total / 16;
    size_t per_cta16 = kind == LMB200_MB_L1_READ ? bytes / 16 : span16 / ctas * (kind == LMB200_MB_L2_READ ? 16 : 1);
    per_cta16 = per_cta16 / 1024 * 1024;                      // whole unrolled iterations of the CTA
    if (per_cta16 == 0 || span16 >= (1ull << 32)) { cudaEventDestroy(a); cudaEventDestroy(b); return LMB200_E_INVALID; }
    const int reps = kind == LMB200_MB_L1_READ ? 256 : 1;
    void* d = nullptr;
    unsigned* sink = nullptr;
    if (cudaMalloc(&d, total) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) rc = LMB200_E_CUDA;
    if (rc == LMB200_OK) {
      cudaMemset(d, 0, total);
      for (int it = 0; it < 2 * (iters + 2); ++it) {  // both kernel forms, alternating; the first two passes of each are warm-up (bring the buffer into L2)
        const bool plain = (it & 1) != 0;
        cudaEventRecord(a);
        if (kind == LMB200_MB_L1_READ) {
          if (plain) read_kernel_plain<1><<<ctas, 256>>>((const uint4*)d, span16, per_cta16, reps, sink);
          else read_kernel<1><<<ctas, 256>>>((const uint4*)d, (unsigned)span16, (unsigned)per_cta16, reps, sink);
        } else {
          if (plain) read_kernel_plain<0><<<ctas, 256>>>((const uint4*)d, span16, per_cta16, reps, sink);
          else read_kernel<0><<<ctas, 256>>>((const uint4*)d, (unsigned)span16, (unsigned)per_cta16, reps, sink);
        }
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (it >= 4 && ms < best_ms) best_ms = ms;
      }
      moved = (double)per_cta16 * 16.0 * ctas * reps;  // per launch (best of iters)
      if (cudaGetLastError() != cudaSuccess) rc = LMB200_E_CUDA;
    }
    if (d) cudaFree(d);
    if (sink) cudaFree(sink);
  } else {
    rc = LMB200_E_INVALID;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  if (rc == LMB200_OK) *gbps = moved / (best_ms * 1e-3) / 1e9;
  return rc;
}
