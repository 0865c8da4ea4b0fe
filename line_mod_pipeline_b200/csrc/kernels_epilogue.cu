// kernels_epilogue.cu — Detector::match's epilogue on the device, for the template-sharded step (sm_100a).
// Upstream: the last lines of Detector::match (opencv_contrib rgbd/linemod.cpp):
//     std::sort(matches.begin(), matches.end());  matches.erase(std::unique(matches.begin(), matches.end()), matches.end());
// Input: the match buffers of every rank's template shard as ncclAllGather left them in device memory (layout:
// gather_pack_kernel in kernels_match.cu).  One CTA per frame:
//   1. the ranks' lists are merged back into the reference's GENERATION order (contiguous shards: rank-ordered
//      concatenation; interleaved shards: by selection position — every rank's list is already ordered by it, so an
//      element's place is its own index plus a binary search in each other rank's list);
//   2. one thread runs libstdc++'s std::sort on (key, source index) pairs in shared memory (sort_emul.h: the sort is not
//      stable, so only the same algorithm leaves tying matches in the reference's order);
//   3. std::unique (adjacent duplicates: comparing with the predecessor is equivalent, equality being transitive), block
//      scan, and the finished records go to pinned host memory with coalesced stores.
// Every rank finishes every frame itself: ~25 k records per 128-frame step are nothing next to a second collective plus
// a host sort on the critical path (round 1: ~3.6 ms per step; this kernel: well under 0.3 ms on a lane of its own).
// Frames with more than EPI_SORT_CAP records, and steps whose headers carry an overflow flag, are flagged for the host path.
#include <cstring>

#include "../../include/lmb200.h"
#include "kernels.cuh"
#include "sort_emul.h"

namespace lmk {

constexpr int EPI_MAX_WORLD = 64;

__global__ void __launch_bounds__(256) shard_epilogue_kernel(EpilogueArgs a) {
  extern __shared__ __align__(16) unsigned char epi_smem[];
  uint64_t* key = reinterpret_cast<uint64_t*>(epi_smem);                       // [EPI_SORT_CAP]
  u32* src = reinterpret_cast<u32*>(epi_smem + (size_t)EPI_SORT_CAP * 8);      // [EPI_SORT_CAP] absolute record index in a.gathered
  u32* pos = src + EPI_SORT_CAP;                                               // [EPI_SORT_CAP] selection position (merge key)
  __shared__ int s_base[EPI_MAX_WORLD + 1];   // first concatenated index of rank r's list
  __shared__ u32 s_rec0[EPI_MAX_WORLD];       // absolute index of rank r's first record of this frame
  __shared__ int s_scan[256];
  __shared__ int s_flags, s_offset;
  const int f = blockIdx.x, tid = threadIdx.x;
  const size_t stride = (size_t)2 * a.frames + a.gcap;   // records per rank in a.gathered

  if (tid == 0) {
    int flags = 0, n = 0;
    for (int r = 0; r < a.world; ++r) {
      const Cand h0 = a.gathered[r * stride + 2 * f];
      flags |= (h0.x & 3);
      s_base[r] = n;
      s_rec0[r] = (u32)(r * stride + 2 * (size_t)a.frames + (size_t)h0.y);
      n += max(h0.tsel, 0);
    }
    s_base[a.world] = n;
    s_flags = flags;
  }
  // output offset: records of all earlier frames (before std::unique), so every frame's region is known without a scan over results
  int part = 0;
  for (int i = tid; i < f * a.world; i += 256) {
    const int r = i / f, ff = i - r * f;
    part += max(a.gathered[r * stride + 2 * ff].tsel, 0);
  }
  s_scan[tid] = part;
  __syncthreads();
  for (int d = 128; d > 0; d >>= 1) {
    if (tid < d) s_scan[tid] += s_scan[tid + d];
    __syncthreads();
  }
  const int n = s_base[a.world];
  if (tid == 0) {
    s_offset = s_scan[0];
    if (n > EPI_SORT_CAP) s_flags |= 4;
    if ((long long)s_scan[0] + n > (long long)a.out_cap) s_flags |= 8;
  }
  __syncthreads();
  const int flags = s_flags, offset = s_offset;
  if (tid == 0) {
    const Cand h1 = a.gathered[a.rank * stride + 2 * f + 1];   // this rank's device counters ride along (profile)
    a.hdr[2 * f] = make_int4(0, offset, flags, n);
    a.hdr[2 * f + 1] = make_int4(h1.tsel, h1.x, h1.y, __float_as_int(h1.sim));
  }
  if (flags) return;   // the host takes this step (lmb200_fetch_resident_allgather, synchronous path)

  // ---- 1. generation order
  constexpr int PER = EPI_SORT_CAP / 256;
  for (int i = tid; i < n; i += 256) {
    int r = 0;
    while (i >= s_base[r + 1]) ++r;
    const u32 ai = s_rec0[r] + (u32)(i - s_base[r]);
    src[i] = ai;
    pos[i] = a.pos_of_g ? (u32)a.pos_of_g[a.gathered[ai].tsel] : (u32)i;
  }
  __syncthreads();
  u32 dst[PER], asrc[PER]; uint64_t k64[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int i = tid + q * 256;
    dst[q] = 0xFFFFFFFFu;
    if (i >= n) continue;
    int r = 0;
    while (i >= s_base[r + 1]) ++r;
    u32 d = (u32)(i - s_base[r]);
    if (a.pos_of_g) {
      const u32 p = pos[i];
      for (int r2 = 0; r2 < a.world; ++r2) {
        if (r2 == r) continue;
        int lo = s_base[r2], hi = s_base[r2 + 1];      // lower_bound(pos[lo..hi), p): a template lives on one rank, so no ties across ranks
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (pos[mid] < p) lo = mid + 1; else hi = mid; }
        d += (u32)(lo - s_base[r2]);
      }
    } else {
      d = (u32)i;
    }
    const Cand c = a.gathered[src[i]];
    dst[q] = d; asrc[q] = src[i];
    k64[q] = lmsort::match_key(c.sim, a.g_tid[c.tsel]);
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < PER; ++q)
    if (dst[q] != 0xFFFFFFFFu) { key[dst[q]] = k64[q]; src[dst[q]] = asrc[q]; }
  __syncthreads();

  // ---- 2. std::sort, exactly
  if (tid == 0) {
    lmsort::Arr arr; arr.k = key; arr.v = src;
    lmsort::std_sort(arr, n);
  }
  __syncthreads();

  // ---- 3. std::unique + compaction; thread t owns the contiguous run [t*chunk, (t+1)*chunk)
  const int chunk = (n + 255) / 256;
  const int k0 = min(n, tid * chunk), k1 = min(n, k0 + chunk);
  u32 keep = 0;   // chunk <= PER <= 32
  for (int k = k0; k < k1; ++k) {
    bool kp = true;
    if (k > 0) {
      const Cand c = a.gathered[src[k]], p = a.gathered[src[k - 1]];
      kp = !(c.x == p.x && c.y == p.y && c.sim == p.sim && a.g_class[c.tsel] == a.g_class[p.tsel]);
    }
    if (kp) keep |= 1u << (k - k0);
  }
  const int mine = __popc(keep);
  s_scan[tid] = mine;
  __syncthreads();
  for (int d = 1; d < 256; d <<= 1) {   // inclusive Hillis-Steele scan
    const int v = tid >= d ? s_scan[tid - d] : 0;
    __syncthreads();
    s_scan[tid] += v;
    __syncthreads();
  }
  const int total = s_scan[255];
  int w = s_scan[tid] - mine;
  EpiMatch* dev = a.out_dev + offset;
  for (int k = k0; k < k1; ++k) {
    if (!((keep >> (k - k0)) & 1u)) continue;
    const Cand c = a.gathered[src[k]];
    EpiMatch m; m.x = c.x; m.y = c.y; m.sim = c.sim; m.class_index = a.g_class[c.tsel]; m.template_id = a.g_tid[c.tsel];
    dev[w++] = m;
  }
  __syncthreads();   // the CTA's own global writes are visible to it after the barrier
  const u32* dw = reinterpret_cast<const u32*>(dev);
  u32* hw = reinterpret_cast<u32*>(a.out_host + offset);
  for (int i = tid; i < total * 5; i += 256) hw[i] = dw[i];   // coalesced stores into pinned host memory
  if (tid == 0) a.hdr[2 * f].x = total;
}

void launch_shard_epilogue(const EpilogueArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)EPI_SORT_CAP * 16;   // 64 KB: above the default limit, opted in per device (the attribute is per context)
  cudaFuncSetAttribute(shard_epilogue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (a.frames > 0) shard_epilogue_kernel<<<a.frames, 256, smem, st>>>(a);
}

}  // namespace lmk

// Test hook (include/lmb200.h): runs shard_epilogue_kernel on host-made "gathered" buffers, through the same pinned,
// device-mapped result area as the product path, so the single-GPU suite covers the device epilogue of the sharded step.
extern "C" int lmb200_debug_shard_epilogue(const int32_t* gathered, int world, int rank, int frames, int gcap, const int32_t* pos_of_g,
                                           const int32_t* g_class, const int32_t* g_tid, int ntpl, lmb200_match_rec* out, size_t out_cap,
                                           int32_t* hdr) {
  using namespace lmk;
  if (!gathered || !g_class || !g_tid || !out || !hdr || world < 1 || world > EPI_MAX_WORLD || rank < 0 || rank >= world || frames < 1 ||
      gcap < 0 || ntpl < 1 || out_cap < 1 || out_cap > 0x7fffffff)
    return LMB200_E_INVALID;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return LMB200_E_NODEVICE; }
  const size_t recs = (size_t)world * ((size_t)2 * frames + gcap);
  Cand* d_g = nullptr; int *d_pos = nullptr, *d_cls = nullptr, *d_tid = nullptr; EpiMatch* d_out = nullptr;
  EpiMatch* h_out = nullptr; int4* h_hdr = nullptr;
  int rc = LMB200_OK;
  auto ok = [&](cudaError_t e) { if (e != cudaSuccess && rc == LMB200_OK) rc = LMB200_E_CUDA; return e == cudaSuccess; };
  ok(cudaMalloc(&d_g, recs * sizeof(Cand)));
  ok(cudaMalloc(&d_cls, (size_t)ntpl * 4)); ok(cudaMalloc(&d_tid, (size_t)ntpl * 4));
  if (pos_of_g) ok(cudaMalloc(&d_pos, (size_t)ntpl * 4));
  ok(cudaMalloc(&d_out, out_cap * sizeof(EpiMatch)));
  ok(cudaHostAlloc((void**)&h_out, out_cap * sizeof(EpiMatch), cudaHostAllocMapped));
  ok(cudaHostAlloc((void**)&h_hdr, (size_t)2 * frames * sizeof(int4), cudaHostAllocMapped));
  if (rc == LMB200_OK) {
    ok(cudaMemcpy(d_g, gathered, recs * sizeof(Cand), cudaMemcpyHostToDevice));
    ok(cudaMemcpy(d_cls, g_class, (size_t)ntpl * 4, cudaMemcpyHostToDevice));
    ok(cudaMemcpy(d_tid, g_tid, (size_t)ntpl * 4, cudaMemcpyHostToDevice));
    if (pos_of_g) ok(cudaMemcpy(d_pos, pos_of_g, (size_t)ntpl * 4, cudaMemcpyHostToDevice));
    std::memset(h_hdr, 0xFF, (size_t)2 * frames * sizeof(int4));
    EpilogueArgs ea;
    ea.gathered = d_g; ea.world = world; ea.rank = rank; ea.frames = frames; ea.gcap = gcap;
    ea.pos_of_g = d_pos; ea.g_class = d_cls; ea.g_tid = d_tid;
    ea.out_dev = d_out; ea.out_cap = (int)out_cap;
    void* dp = nullptr;
    ok(cudaHostGetDevicePointer(&dp, h_out, 0)); ea.out_host = (EpiMatch*)dp;
    ok(cudaHostGetDevicePointer(&dp, h_hdr, 0)); ea.hdr = (int4*)dp;
    if (rc == LMB200_OK) {
      launch_shard_epilogue(ea, 0);
      ok(cudaDeviceSynchronize());
    }
    if (rc == LMB200_OK) {
      std::memcpy(out, h_out, out_cap * sizeof(EpiMatch));
      std::memcpy(hdr, h_hdr, (size_t)2 * frames * sizeof(int4));
    }
  }
  cudaFree(d_g); cudaFree(d_cls); cudaFree(d_tid); cudaFree(d_pos); cudaFree(d_out);
  if (h_out) cudaFreeHost(h_out);
  if (h_hdr) cudaFreeHost(h_hdr);
  return rc;
}
