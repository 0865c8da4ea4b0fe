// kernels_match.cu — template-side kernels (sm_100a): the hot loops of Detector::matchClass
// (opencv_contrib rgbd/linemod.cpp):
//   build_offsets_kernel        accessLinearMemory() for every feature, once per frame geometry  (a11)
//   similarity_coarse_kernel    similarity + addSimilarities + threshold scan, fused            (a12-a14)
//   similarity_local_kernel     similarityLocal + addSimilarities + argmax + update + filter    (a15,a16)
//   pack_kernel                 ordered compaction into reference generation order              (a17)
// These are sparse u8 gather-accumulates over L2-resident linear memories: no tensor cores.
// Packed byte sums never carry between bytes (<= 63 features x 4 per modality = 252), so one 32-bit
// IADD adds four responses; unaligned feature rows are read as aligned 128-bit (coarse) / 64-bit
// (local) loads and realigned with funnel shifts.
#include "kernels.cuh"

namespace lmk {

// ---------------------------------------------------------------------------------------------
// Plan kernel: one warp per template.  For every modality the valid feature offsets
//   off = label*per_label + grid_index*W*H + lm_index            (accessLinearMemory)
// become 8-byte plan rows.  Coarsest level: rows SORTED by the word shift (off>>3)&3 of their start in the
// nibble-packed memory (warp-level counting sort), every bucket padded to a multiple of 3 with rows of the zero
// tail, the four bucket sizes in hdr.bkt[m]; a row = (byte offset of its 16 B chunk, funnel shift in bits).
// Finer levels (strip layout): a row = (offset of label/cell/grid-row, grid column).
// Features upstream's similarity() would skip are dropped.
// hdr.flags: bit0 local-safe  : similarityLocal can never skip a feature or leave its plane
//            bit1 coarse-safe : every feature row + template_positions stays inside its label's block
//            bit2 one-P       : template_positions is the same for every modality
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) build_offsets_kernel(const u32* __restrict__ feat, u32* __restrict__ offs,
                                                            TplHdr* __restrict__ hdr, int ntpl, int M, LevelGeom g, int coarsest) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= ntpl) return;
  TplHdr h = hdr[warp];
  const int T = g.T, W = g.W, H = g.H;
  const u32 plane = (u32)W * H;
  const u32 lt = (1u << lane) - 1u;
  bool local_safe = h.width[0] >= 0 && h.height[0] >= 0 && h.width[0] <= g.cols - 16 * T && h.height[0] <= g.rows - 16 * T;
  bool coarse_safe = true;
  int pmin = 0x7FFFFFFF, pmax = -0x7FFFFFFF;
  for (int m = 0; m < M; ++m) {
    u32 off2[2], gx2[2];
    int key2[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = lane + 32 * j;
      u32 off = OFF_INVALID, gx0 = 0;
      if (k < (int)hdr[warp].nf[m]) {
        u32 f = feat[((size_t)warp * M + m) * FEAT_SLOTS + k];
        int x = f & 0x3FFF, y = (f >> 14) & 0x3FFF, label = (f >> 28) & 7;
        bool valid = (f >> 31) && x < g.cols && y < g.rows;
        if (valid) {
          u32 base = (u32)((y % T) * T + (x % T)) * plane + (u32)(y / T) * W + (u32)(x / T);
          off = (u32)label * g.per_label + base;
          if (!coarsest) {  // strip layout: byte offset of (label, phase, row y/T) without the column part, and the column
            off = (u32)label * g.per_label + (u32)((y % T) * T + (x % T)) * g.plane + (u32)(y / T) * 16u;
            gx0 = (u32)(x / T);
          }
          int wf = (hdr[warp].width[m] - 1) / T + 1, hf = (hdr[warp].height[m] - 1) / T + 1;
          long long P = (long long)(H - hf) * W + (W - wf) + 1;
          if (P > (long long)plane) P = plane;
          if (P > 0 && (long long)base + P > (long long)g.per_label) coarse_safe = false;
        }
        if (!(f >> 31) || x > h.width[0] || y > h.height[0]) local_safe = false;
      }
      off2[j] = off; gx2[j] = gx0;
      key2[j] = off == OFF_INVALID ? 4 : (coarsest ? (int)((off >> 3) & 3) : 0);
    }
    // counting sort by key (0..3 valid buckets, 4 = dropped)
    const int slots = coarsest ? COARSE_SLOTS : FEAT_SLOTS;
    u32 start = 0, packed = 0;
    u32 dst[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
    // entries are pairs: coarsest level (16 B chunk byte offset in the nibble-packed LM, funnel shift in bits);
    // finer levels (strip-layout byte offset without the column part, column)
    u32* o = offs + ((size_t)warp * M + m) * slots * 2;
    for (int s0 = lane; s0 < slots * 2; s0 += 32) o[s0] = OFF_INVALID;
    __syncwarp();
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      u32 b0 = __ballot_sync(0xffffffffu, key2[0] == w), b1 = __ballot_sync(0xffffffffu, key2[1] == w);
      u32 c0 = __popc(b0), c1 = __popc(b1);
      if (key2[0] == w) dst[0] = start + __popc(b0 & lt);
      if (key2[1] == w) dst[1] = start + c0 + __popc(b1 & lt);
      u32 cnt = c0 + c1;
      if (coarsest) {  // pad to a multiple of 3 with rows of the zero tail that have this bucket's word shift
        u32 padded = (cnt + 2) / 3 * 3;
        if (lane < (int)(padded - cnt)) { o[2 * (start + cnt + lane)] = 4u * g.per_label; o[2 * (start + cnt + lane) + 1] = 0u; }
        cnt = padded;
      }
      packed |= cnt << (8 * w);
      start += cnt;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (dst[j] != 0xFFFFFFFFu) {
        if (coarsest) { o[2 * dst[j]] = (off2[j] >> 1) & ~15u; o[2 * dst[j] + 1] = (off2[j] & 7u) * 4u; }
        else { o[2 * dst[j]] = off2[j]; o[2 * dst[j] + 1] = gx2[j]; }
      }
    {
      int wf = (hdr[warp].width[m] - 1) / T + 1, hf = (hdr[warp].height[m] - 1) / T + 1;
      int P = (H - hf) * W + (W - wf) + 1;
      P = P > (int)plane ? (int)plane : P;
      pmin = min(pmin, P); pmax = max(pmax, P);
      if (lane == 0) { hdr[warp].bkt[m] = packed; hdr[warp].P[m] = P; }
    }
  }
  local_safe = __all_sync(0xffffffffu, local_safe);
  coarse_safe = __all_sync(0xffffffffu, coarse_safe);
  if (lane == 0) hdr[warp].flags = (local_safe ? 1u : 0u) | (coarse_safe ? 2u : 0u) | (pmin == pmax ? 4u : 0u);
}

void launch_build_offsets(const u32* feat, u32* offs, TplHdr* hdr, int ntpl, int M, LevelGeom g, bool coarsest, cudaStream_t st) {
  if (ntpl <= 0) return;
  int blocks = (ntpl * 32 + 127) / 128;
  build_offsets_kernel<<<blocks, 128, 0, st>>>(feat, offs, hdr, ntpl, M, g, coarsest ? 1 : 0);
}

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
// zero the bytes at index >= nv (nv in [0,16]) of a 16-byte value held in 4 little-endian words
__device__ __forceinline__ void mask16(u32& w0, u32& w1, u32& w2, u32& w3, int nv) {
  u32 w[4] = {w0, w1, w2, w3};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = nv - 4 * i;
    u32 m = r >= 4 ? 0xFFFFFFFFu : (r <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - r))));
    w[i] &= m;
  }
  w0 = w[0]; w1 = w[1]; w2 = w[2]; w3 = w[3];
}

__device__ __forceinline__ int raw_threshold(int nf, float threshold) {
  // (int)(2*nf + (threshold/100.f)*(2*nf) + 0.5f), float arithmetic without contraction
  float two_nf = (float)(2 * nf);
  return (int)__fadd_rn(__fadd_rn(two_nf, __fmul_rn(__fdiv_rn(threshold, 100.f), two_nf)), 0.5f);
}

// ---------------------------------------------------------------------------------------------
// Coarse similarity on the NIBBLE-PACKED linear memories of the coarsest level (responses are 0..4:
// two positions per byte halve the L1/L2 traffic of the gather).  One WARP per (template, frame), no
// block barriers.  A lane owns 32 consecutive positions = one aligned 16 B chunk; a pass covers 1024
// positions (cfg-A: 1200 positions, typical template_positions 800-1000 -> one pass).
// Per feature: two aligned LDG.128 + four funnel shifts realign the row (the word shift is a compile-
// time constant per plan bucket), four 32-bit IADDs add 32 responses into nibble accumulators; every
// third feature (3*4 = 12 <= 15) they are spilled into byte accumulators.  Hits are appended in raster
// order to a per-warp queue; one atomicAdd per template with hits reserves a contiguous block of the
// frame's candidate store (templates with more hits than the queue holds re-sweep and write directly).
// ---------------------------------------------------------------------------------------------
constexpr int CW_WARPS = 4;          // warps (templates) per CTA
constexpr int CW_QCAP = 96;          // queued hits per warp
#ifndef CW_TPW
#define CW_TPW 8                     // most templates per warp (consecutive groups of CW_WARPS); chosen per launch
#endif
#ifndef CW_MINB
#define CW_MINB 7                    // resident CTAs per SM the register allocation targets (7 -> <= 72 registers; measured best of 5..8)
#endif

// zero the nibbles at index >= nv (nv < 32) of 32 nibbles held in 4 little-endian words
__device__ __forceinline__ void mask_nibbles(u32& w0, u32& w1, u32& w2, u32& w3, int nv) {
  u32 w[4] = {w0, w1, w2, w3};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = nv - 8 * i;
    u32 m = r >= 8 ? 0xFFFFFFFFu : (r <= 0 ? 0u : ((1u << (4 * r)) - 1u));
    w[i] &= m;
  }
  w0 = w[0]; w1 = w[1]; w2 = w[2]; w3 = w[3];
}

struct NibAcc {
  u32 n0, n1, n2, n3;   // nibble accumulators (32 positions)
  u32 b[8];             // byte accumulators: b[2i+par] byte t <-> position 8i + 2t + par
  __device__ __forceinline__ void spill() {
    b[0] += n0 & 0x0F0F0F0Fu; b[1] += (n0 >> 4) & 0x0F0F0F0Fu;
    b[2] += n1 & 0x0F0F0F0Fu; b[3] += (n1 >> 4) & 0x0F0F0F0Fu;
    b[4] += n2 & 0x0F0F0F0Fu; b[5] += (n2 >> 4) & 0x0F0F0F0Fu;
    b[6] += n3 & 0x0F0F0F0Fu; b[7] += (n3 >> 4) & 0x0F0F0F0Fu;
    n0 = n1 = n2 = n3 = 0u;
  }
  __device__ __forceinline__ void mask_tail(int rem) { mask_tail8(b, rem); }
  // zero the byte sums of positions >= rem (0 < rem < 32): v[2i+par] byte t <-> position 8i + 2t + par
  static __device__ __forceinline__ void mask_tail8(u32* v, int rem) {
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      int nb = (rem - 8 * (w >> 1) - (w & 1) + 1) >> 1;  // valid bytes in this word
      u32 m = nb >= 4 ? 0xFFFFFFFFu : (nb <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - nb))));
      v[w] &= m;
    }
  }
};

template <int WS, bool MASK>
__device__ __forceinline__ void realign_add(const uint4& A, const uint4& B, u32 sh, int nv, NibAcc& acc) {
  u32 w0, w1, w2, w3;
  if (WS == 0) {
    w0 = __funnelshift_r(A.x, A.y, sh); w1 = __funnelshift_r(A.y, A.z, sh);
    w2 = __funnelshift_r(A.z, A.w, sh); w3 = __funnelshift_r(A.w, B.x, sh);
  } else if (WS == 1) {
    w0 = __funnelshift_r(A.y, A.z, sh); w1 = __funnelshift_r(A.z, A.w, sh);
    w2 = __funnelshift_r(A.w, B.x, sh); w3 = __funnelshift_r(B.x, B.y, sh);
  } else if (WS == 2) {
    w0 = __funnelshift_r(A.z, A.w, sh); w1 = __funnelshift_r(A.w, B.x, sh);
    w2 = __funnelshift_r(B.x, B.y, sh); w3 = __funnelshift_r(B.y, B.z, sh);
  } else {
    w0 = __funnelshift_r(A.w, B.x, sh); w1 = __funnelshift_r(B.x, B.y, sh);
    w2 = __funnelshift_r(B.y, B.z, sh); w3 = __funnelshift_r(B.z, B.w, sh);
  }
  if (MASK && nv < 32) mask_nibbles(w0, w1, w2, w3, nv);
  acc.n0 += w0; acc.n1 += w1; acc.n2 += w2; acc.n3 += w3;
}

// One word-shift bucket: its (padded) row count n is a multiple of 3, so the loop is groups of 3 rows only:
// all loads first (6 x LDG.128 in flight per lane), then realign + one 3-input add per word (3*4 = 12 fits a
// nibble), then the spill into byte sums.
// CW_SHFL (experiment, -DCW_SHFL): a lane's second chunk is its right neighbour's first one, so it is taken with four
// shuffles instead of a second LDG.128; only the last active lane loads it.  act = ballot of the lanes inside this call.
template <int WS, bool SAFE>
__device__ __forceinline__ void accum_bucket(const u8* __restrict__ lmb, const uint2* __restrict__ lst, int k0, int n,
                                             int pos0, int P, u32 per_label, NibAcc& acc, u32 act = 0xffffffffu, bool last = false) {
  for (int k = k0; k < k0 + n; k += 3) {
    uint2 e[3];  // (chunk byte offset, funnel shift) straight from the plan: no per-row address arithmetic
    uint4 A[3], B[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) e[j] = lst[k + j];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const uint4* p = reinterpret_cast<const uint4*>(lmb + e[j].x);
#ifdef CW_SHFL
      A[j] = __ldg(p);
      B[j].x = __shfl_down_sync(act, A[j].x, 1); B[j].y = __shfl_down_sync(act, A[j].y, 1);
      B[j].z = __shfl_down_sync(act, A[j].z, 1); B[j].w = __shfl_down_sync(act, A[j].w, 1);
      if (last) B[j] = __ldg(p + 1);
#else
      A[j] = __ldg(p); B[j] = __ldg(p + 1);
#endif
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      int nv = 32;
      if (!SAFE) {  // positions past the label block contribute 0
        const u32 off = 2u * e[j].x + 8u * WS + (e[j].y >> 2);  // the row's nibble offset
        nv = min(P, (int)(per_label - off % per_label)) - pos0;
      }
      if (SAFE || nv > 0) realign_add<WS, !SAFE>(A[j], B[j], e[j].y, nv, acc);
    }
    acc.spill();
  }
}

struct CoarseCtx {
  HdrR hdr;
  const uint2* lst;                  // this warp's staged plan rows [M][COARSE_SLOTS] (shared memory)
  int M, T, W, H, HW, raw_thr, nf_total;
  u32 per_label;
  const int* Pm;                     // plan: upstream's template_positions per modality
  const int* ord;                    // this frame's modality order (shared memory)
  int early_exit;
  mutable u32 chunks;                // 16 B chunk loads issued for this template (lane-uniform)
  u16* dump;                         // DUMP instantiation: the template's full u16 similarity map [H*W] (pre-zeroed)
  __device__ __forceinline__ int P(int m) const { return __ldg(Pm + m); }
};

// value of position p (compile-time) from the byte accumulators / the u16 totals
template <bool WIDE>
__device__ __forceinline__ int hit_val(const u32* t, int p) {
  const int w = 2 * (p >> 3) + (p & 1), tt = (p & 7) >> 1;
  if (WIDE) return (int)((t[2 * w + (tt & 1)] >> (16 * (tt >> 1))) & 0xFFFFu);
  return (int)((t[w] >> (8 * tt)) & 0xFFu);
}

// WIDE=false: 4*nf_total <= 255, every modality accumulates into the same byte sums.
// WIDE=true : per-modality byte sums are widened into u16 totals (upstream's addSimilarities).
// direct=false: count hits and queue them (entries beyond CW_QCAP are dropped, the count continues);
// direct=true : write Cand records at out[0..) in raster order.
// DUMP=true (lmb200_debug_fetch(LMB200_DBG_SIMILARITY)): the same accumulation with the early exit switched off; every
// position's total is written to cx.dump instead of being thresholded (upstream's similarity() + addSimilarities()).
template <bool SAFE, bool WIDE, bool DUMP = false>
__device__ __forceinline__ int coarse_sweep(const CoarseCtx& cx, const u8* const* s_lm, u32* queue, Cand* out, bool direct, int isel, int lane) {
  int nhit = 0;
  int Pmax = 0;
  for (int m = 0; m < cx.M; ++m) Pmax = max(Pmax, cx.P(m));
  // All modalities of a cropped pyramid share one template_positions (plan flag bit2): the row tail is then cut once,
  // on the totals, after the early exit (which may look at the uncut sums: it only errs towards continuing).
#ifdef CW_NO_TAIL_ONCE
#define tail_once false
#else
#define tail_once (!WIDE && (cx.hdr.flags & 4u))
#endif
  const int offset = cx.T / 2 + (cx.T % 2 - 1);
  const float denom = (float)(4 * cx.nf_total);
  const u32 K = (u32)(0x7FFF - cx.raw_thr) * 0x00010001u;  // field + K sets bit 15 iff field > raw_thr
  for (int c0 = 0; 32 * c0 < Pmax; c0 += 32) {
    const int pos0 = 32 * (c0 + lane);
    NibAcc acc;
    acc.n0 = acc.n1 = acc.n2 = acc.n3 = 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.b[i] = 0u;
    u32 tot[WIDE ? 16 : 8];  // WIDE: u16 pairs (see wide_val); else byte sums laid out like acc.b
#pragma unroll
    for (int i = 0; i < (WIDE ? 16 : 8); ++i) tot[i] = 0u;
    bool dead = false;
    for (int mi = 0; mi < cx.M; ++mi) {
      const int m = cx.ord[mi];  // most selective modality of this frame first (see the early exit below)
      const int P = cx.P(m);
      if (32 * c0 >= P) continue;  // warp-uniform
      const int rem = P - pos0;
      const u8* lmb = s_lm[m] + (pos0 >> 1);
      const u32 bk = cx.hdr.bkt(m);
      const int n0 = bk & 255, n1 = (bk >> 8) & 255, n2 = (bk >> 16) & 255, n3 = bk >> 24;
      const uint2* lst = cx.lst + m * COARSE_SLOTS;
      const u32 act_m = __ballot_sync(0xffffffffu, rem > 0);   // a prefix of the warp: rem falls with the lane
      cx.chunks += 2u * (u32)(n0 + n1 + n2 + n3) * (u32)__popc(act_m);
      if (rem > 0) {  // lanes past template_positions sit the modality out (one branch, not one per group)
#ifdef CW_SHFL
        const u32 act = act_m;
        const bool last = lane == 31 - __clz(act);
        accum_bucket<0, SAFE>(lmb, lst, 0, n0, pos0, P, cx.per_label, acc, act, last);
        accum_bucket<1, SAFE>(lmb, lst, n0, n1, pos0, P, cx.per_label, acc, act, last);
        accum_bucket<2, SAFE>(lmb, lst, n0 + n1, n2, pos0, P, cx.per_label, acc, act, last);
        accum_bucket<3, SAFE>(lmb, lst, n0 + n1 + n2, n3, pos0, P, cx.per_label, acc, act, last);
#else
        accum_bucket<0, SAFE>(lmb, lst, 0, n0, pos0, P, cx.per_label, acc);
        accum_bucket<1, SAFE>(lmb, lst, n0, n1, pos0, P, cx.per_label, acc);
        accum_bucket<2, SAFE>(lmb, lst, n0 + n1, n2, pos0, P, cx.per_label, acc);
        accum_bucket<3, SAFE>(lmb, lst, n0 + n1 + n2, n3, pos0, P, cx.per_label, acc);
#endif
      }
      if (SAFE && !tail_once && rem > 0 && rem < 32) acc.mask_tail(rem);  // at most one lane per modality: the row tail
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        if (WIDE) {
          tot[2 * w] += acc.b[w] & 0x00FF00FFu;
          tot[2 * w + 1] += (acc.b[w] >> 8) & 0x00FF00FFu;
        } else {
          tot[w] += acc.b[w];
        }
        acc.b[w] = 0u;
      }
      // exact early exit: when no position of this pass can still exceed raw_thr, the remaining modalities
      // cannot produce a hit (responses are <= 4 per feature) and the pass ends here
      int later = 0;  // features of the modalities still to come
      for (int k = mi + 1; k < cx.M; ++k) later += cx.hdr.nf(cx.ord[k]);
      const int bound = cx.raw_thr - 4 * later;
      if (!DUMP && cx.early_exit && mi + 1 < cx.M && bound >= 0) {  // warp-uniform
        const u32 Kb = (u32)(0x7FFF - bound) * 0x00010001u;
        u32 alive = 0;
        if (WIDE) {
#pragma unroll
          for (int i = 0; i < 16; ++i) alive |= tot[i] + Kb;
        } else {
#pragma unroll
          for (int w = 0; w < 8; ++w) alive |= ((tot[w] & 0x00FF00FFu) + Kb) | (((tot[w] >> 8) & 0x00FF00FFu) + Kb);
        }
        if (!__any_sync(0xffffffffu, (alive & 0x80008000u) != 0u)) { dead = true; break; }
      }
    }
    if (dead) continue;
    if (SAFE && tail_once) {
      const int rem = Pmax - pos0;
      if (rem > 0 && rem < 32) NibAcc::mask_tail8(tot, rem);
    }
    if (DUMP) {
#pragma unroll
      for (int p = 0; p < 32; ++p)
        if (pos0 + p < cx.HW) cx.dump[pos0 + p] = (u16)hit_val<WIDE>(tot, p);
      continue;
    }
    u32 any = 0;
    if (WIDE) {
#pragma unroll
      for (int i = 0; i < 16; ++i) any |= tot[i] + K;
    } else {
#pragma unroll
      for (int w = 0; w < 8; ++w) any |= ((tot[w] & 0x00FF00FFu) + K) | (((tot[w] >> 8) & 0x00FF00FFu) + K);
    }
    if (!__any_sync(0xffffffffu, (any & 0x80008000u) != 0u)) continue;
    // rare path: exact per-position test in raster order
    u32 mask = 0;
#pragma unroll
    for (int p = 0; p < 32; ++p) {
      const int v = hit_val<WIDE>(tot, p);
      mask |= (v > cx.raw_thr ? 1u : 0u) << p;
    }
    const int mine = __popc(mask);
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int slot = nhit + incl - mine;
#pragma unroll
    for (int p = 0; p < 32; ++p) {
      if (mask & (1u << p)) {
        const int raw = hit_val<WIDE>(tot, p);
        const int j = pos0 + p;
        if (direct) {
          const int row = j / cx.W, col = j - row * cx.W;
          Cand cd;
          cd.tsel = isel;
          cd.x = col * cx.T + offset;
          cd.y = row * cx.T + offset;
          cd.sim = __fadd_rn(__fdiv_rn(__fmul_rn((float)raw, 100.f), denom), 0.5f);
          out[slot] = cd;
        } else if (slot < CW_QCAP) {
          queue[slot] = ((u32)j << 11) | (u32)raw;  // raw <= 4*63*MAX_MOD = 1008 < 2048; j < 2^21
        }
        ++slot;
      }
    }
    nhit += total;
  }
  return nhit;
}

#undef tail_once

template <bool SAFE, bool WIDE>
__device__ __forceinline__ void coarse_template(const MatchParams& mp, const CoarseCtx& cx, const u8* const* s_lm, u32* queue,
                                                int isel, int frame, int lane) {
  const int nhit = coarse_sweep<SAFE, WIDE>(cx, s_lm, queue, nullptr, false, isel, lane);
  int base = 0, run = nhit;
  if (lane == 0) {
    if (nhit > 0) {
      base = atomicAdd(&mp.ctr[frame].cand_count, nhit);
      if (base + nhit > mp.cand_cap) { mp.ctr[frame].overflow = 1; run = 0; }
    }
    mp.tpl_start[(size_t)frame * mp.nsel_stride + isel] = base;
    mp.tpl_cnt[(size_t)frame * mp.nsel_stride + isel] = run;
    mp.tpl_alive[(size_t)frame * mp.nsel_stride + isel] = run;
    atomicAdd(&mp.ctr[frame].coarse_chunks, (unsigned long long)cx.chunks);
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  run = __shfl_sync(0xffffffffu, run, 0);
  if (run == 0) return;
  Cand* out = mp.cand + (size_t)frame * mp.cand_cap + base;
  if (nhit <= CW_QCAP) {
    __syncwarp();
    const int offset = cx.T / 2 + (cx.T % 2 - 1);
    const float denom = (float)(4 * cx.nf_total);
    for (int q = lane; q < nhit; q += 32) {
      const u32 e = queue[q];
      const int j = (int)(e >> 11), raw = (int)(e & 2047u);
      const int row = j / cx.W, col = j - row * cx.W;
      Cand cd;
      cd.tsel = isel;
      cd.x = col * cx.T + offset;
      cd.y = row * cx.T + offset;
      cd.sim = __fadd_rn(__fdiv_rn(__fmul_rn((float)raw, 100.f), denom), 0.5f);
      out[q] = cd;
    }
  } else {
    coarse_sweep<SAFE, WIDE>(cx, s_lm, nullptr, out, true, isel, lane);
  }
}

template <bool WIDE, bool DUMP = false>
__global__ void __launch_bounds__(CW_WARPS * 32, CW_MINB) similarity_coarse_kernel(MatchParams mp, LevelParams lp, int tpw, u16* dump = nullptr) {
  __shared__ const u8* s_lm[MAX_MOD];
  __shared__ int s_ord[MAX_MOD];
  __shared__ u32 s_queue[CW_WARPS][CW_QCAP];
  __shared__ uint2 s_off[CW_WARPS][MAX_MOD * COARSE_SLOTS];
  const int frame = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {  // lp.lm = nibble-packed linear memories of the coarsest level
    s_lm[0] = lp.lm[0] + (size_t)frame * lp.lm_stride[0];
    s_lm[1] = lp.lm[1] + (size_t)frame * lp.lm_stride[1];
    s_lm[2] = lp.lm[2] + (size_t)frame * lp.lm_stride[2];
    s_lm[3] = lp.lm[3] + (size_t)frame * lp.lm_stride[3];
    // Modalities are summed in ascending order of this frame's mean response: the modality with the lowest
    // responses bounds the scores hardest, so the early exit of coarse_sweep fires for the most templates.
    // Any order gives the same sums; the order only decides how much work the exit saves.
    u32 key[MAX_MOD];
    for (int m = 0; m < MAX_MOD; ++m) { s_ord[m] = m; key[m] = m < mp.M ? __ldg(lp.resp_sum + (size_t)frame * MAX_MOD + m) : 0xFFFFFFFFu; }
    for (int i = 1; i < mp.M; ++i)
      for (int j = i; j > 0 && key[s_ord[j]] < key[s_ord[j - 1]]; --j) { int t = s_ord[j]; s_ord[j] = s_ord[j - 1]; s_ord[j - 1] = t; }
  }
  __syncthreads();
  for (int it = 0; it < tpw; ++it) {  // a warp scores tpw templates: early exits even out over the sequence
    const int isel = (blockIdx.x * tpw + it) * CW_WARPS + warp;
    if (isel >= mp.nsel) return;
    const int g = mp.sel[isel];
    CoarseCtx cx;
    cx.hdr = load_hdr(lp.hdr + g);
    cx.M = mp.M; cx.T = lp.g.T; cx.W = lp.g.W; cx.H = lp.g.H; cx.HW = lp.g.W * lp.g.H;
    cx.per_label = lp.g.per_label;
    cx.nf_total = cx.hdr.nf_total();
    cx.raw_thr = raw_threshold(cx.nf_total, mp.threshold);
    if (cx.raw_thr > 0x7FFE) cx.raw_thr = 0x7FFE;  // nothing can exceed it anyway (scores <= 1008)
    if (cx.raw_thr < 0) cx.raw_thr = -1;
    cx.lst = s_off[warp];
    cx.Pm = lp.hdr[g].P;
    cx.ord = s_ord;
    cx.dump = dump;
    cx.early_exit = mp.early_exit;
    cx.chunks = 0u;
    __syncwarp();  // the previous template's reads of s_off / s_queue are done
    for (int s = lane; s < cx.M * COARSE_SLOTS; s += 32)
      s_off[warp][s] = __ldg(reinterpret_cast<const uint2*>(lp.offs) + (size_t)g * cx.M * COARSE_SLOTS + s);
    __syncwarp();
    if (DUMP) {  // one template, no candidate bookkeeping
      if (cx.hdr.flags & 2u) coarse_sweep<true, WIDE, true>(cx, s_lm, nullptr, nullptr, false, isel, lane);
      else coarse_sweep<false, WIDE, true>(cx, s_lm, nullptr, nullptr, false, isel, lane);
      return;
    }
    if (cx.hdr.flags & 2u) coarse_template<true, WIDE>(mp, cx, s_lm, s_queue[warp], isel, frame, lane);
    else coarse_template<false, WIDE>(mp, cx, s_lm, s_queue[warp], isel, frame, lane);
  }
}

// Debug path: the production accumulation of ONE template (mp.sel points at its entry, mp.nsel == 1, one frame) with the
// early exit disabled, totals written to map[H*W] (u16, zero where upstream's similarity() leaves dst untouched).
void launch_similarity_map(const MatchParams& mp, const LevelParams& lp, bool wide, u16* map, cudaStream_t st) {
  cudaMemsetAsync(map, 0, (size_t)lp.g.W * lp.g.H * sizeof(u16), st);
  dim3 grid(1, 1);
  if (wide) similarity_coarse_kernel<true, true><<<grid, CW_WARPS * 32, 0, st>>>(mp, lp, 1, map);
  else similarity_coarse_kernel<false, true><<<grid, CW_WARPS * 32, 0, st>>>(mp, lp, 1, map);
}

// wide: some template has 4*nf_total > 255 at the coarsest level (byte sums across modalities could carry)
void launch_similarity_coarse(const MatchParams& mp, const LevelParams& lp, bool wide, cudaStream_t st) {
  if (mp.nsel <= 0 || mp.frames <= 0) return;
  // templates per warp: 1 while the grid is a few waves of the machine (latency matters), up to CW_TPW for big batches
  static int sm_count[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!sm_count[dev]) { cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev); if (sm_count[dev] <= 0) sm_count[dev] = 148; }
  const long long ctas1 = (long long)((mp.nsel + CW_WARPS - 1) / CW_WARPS) * mp.frames;
  const long long wave = (long long)sm_count[dev] * CW_MINB;
  int tpw = (int)(ctas1 / (4 * wave));
  tpw = tpw < 1 ? 1 : (tpw > CW_TPW ? CW_TPW : tpw);
  dim3 grid((mp.nsel + CW_WARPS * tpw - 1) / (CW_WARPS * tpw), mp.frames);
  if (wide) similarity_coarse_kernel<true><<<grid, CW_WARPS * 32, 0, st>>>(mp, lp, tpw);
  else similarity_coarse_kernel<false><<<grid, CW_WARPS * 32, 0, st>>>(mp, lp, tpw);
}

// Packs a byte linear memory (values 0..4) two positions per byte: out[q] = in[2q] | in[2q+1] << 4.  The first warp
// of every block also adds the responses it packs into resp_sum[frame]: a 1/8 sample of the linear memory, the
// per-frame statistic that orders the modalities in the coarse kernel (one atomic per block).
__global__ void __launch_bounds__(256) pack_nibbles_kernel(const u8* __restrict__ lm, size_t lm_stride, u8* __restrict__ lmn,
                                                           size_t lmn_stride, u32 n_out8 /* 8-byte output groups */,
                                                           u32* __restrict__ resp_sum, int resp_stride) {
  const u32 i = blockIdx.x * 256 + threadIdx.x;
  u32 sum = 0;
  if (i < n_out8) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(lm + (size_t)blockIdx.y * lm_stride) + i);
    const u32 t0 = v.x | (v.x >> 4), t1 = v.y | (v.y >> 4), t2 = v.z | (v.z >> 4), t3 = v.w | (v.w >> 4);
    uint2 o;
    o.x = __byte_perm(t0, t1, 0x6420);
    o.y = __byte_perm(t2, t3, 0x6420);
    reinterpret_cast<uint2*>(lmn + (size_t)blockIdx.y * lmn_stride)[i] = o;
    if (threadIdx.x < 32) sum = __dp4a(v.x, 0x01010101u, __dp4a(v.y, 0x01010101u, __dp4a(v.z, 0x01010101u, __dp4a(v.w, 0x01010101u, 0u))));
  }
  if (threadIdx.x >= 32) return;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  if (threadIdx.x == 0 && sum) atomicAdd(resp_sum + (size_t)blockIdx.y * resp_stride, sum);
}

void launch_pack_nibbles(const u8* lm, size_t lm_stride, u8* lmn, size_t lmn_stride, LevelGeom g, int frames,
                         u32* resp_sum, int resp_stride, cudaStream_t st) {
  u32 n_out8 = g.per_label / 2;  // 8*per_label positions / 16 per thread
  dim3 grid((n_out8 + 255) / 256, frames);
  pack_nibbles_kernel<<<grid, 256, 0, st>>>(lm, lm_stride, lmn, lmn_stride, n_out8, resp_sum, resp_stride);
}

// ---------------------------------------------------------------------------------------------
// Local refinement: one CTA (4 warps) per candidate, the template's features are split across the warps and the
// four partial 16x16 maps are added through shared memory (a 4x shorter dependent-load chain per candidate:
// small batches are latency-bound); lane = (row 0..15, half 0..1) owns 8 of the 16x16 sums.
// Fast path (hdr.flags bit0): plan entries (strip-layout offset, column) shifted by the candidate's patch origin.
// Slow path: upstream's per-feature bounds checks with guarded byte loads (malformed / oversized
// templates — N4 in SURVEY.md — where upstream itself is undefined; semantics = oracle's).
// ---------------------------------------------------------------------------------------------
// WPC = false: one CTA per candidate, its features split over the four warps (shortest dependent chain: single frames, few
//              candidates).  WPC = true: one WARP per candidate, no block barriers and no shared-memory reduction — four
//              independent candidates per CTA keep the load pipeline busier when there are thousands of them per launch.
template <bool WPC>
__global__ void __launch_bounds__(128) similarity_local_kernel(MatchParams mp, LevelParams lp) {
  const int frame = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (mp.ctr[frame].overflow) return;  // candidate store overflowed: the host grows it and redoes this frame
  const int n = min(mp.ctr[frame].cand_count, mp.cand_cap);
  const int T = lp.g.T, W = lp.g.W, M = mp.M;
  const int border = 8 * T, offset = T / 2 + (T % 2 - 1);
  const u32 plane = (u32)W * lp.g.H;   // upstream's flat plane (slow path semantics)
  const u32 H16 = (u32)lp.g.H * 16u;
  Cand* cands = mp.cand + (size_t)frame * mp.cand_cap;
  const int rr = lane >> 1, hh = lane & 1;
  __shared__ const u8* s_lm[MAX_MOD];
  if (threadIdx.x == 0) {
    s_lm[0] = lp.lm[0] + (size_t)frame * lp.lm_stride[0];
    s_lm[1] = lp.lm[1] + (size_t)frame * lp.lm_stride[1];
    s_lm[2] = lp.lm[2] + (size_t)frame * lp.lm_stride[2];
    s_lm[3] = lp.lm[3] + (size_t)frame * lp.lm_stride[3];
  }
  __shared__ uint4 s_part[4][32];
  __shared__ uint2 s_ent[4][32];
  __syncthreads();

  for (int c = WPC ? blockIdx.x * 4 + warp : blockIdx.x; c < n; c += WPC ? gridDim.x * 4 : gridDim.x) {
    Cand rec = cands[c];
    if (rec.sim < 0.f || rec.tsel < 0 || rec.tsel >= mp.nsel) continue;  // CTA-uniform (WPC: warp-uniform)
    const int g = mp.sel[rec.tsel];
    const HdrR hdr = load_hdr(lp.hdr + g);
    int x = rec.x * 2 + 1, y = rec.y * 2 + 1;
    x = max(x, border); y = max(y, border);
    x = min(x, lp.g.cols - hdr.width(0) - border);
    y = min(y, lp.g.rows - hdr.height(0) - border);
    const int cxT = x / T - 8, cyT = y / T - 8;
    u32 t0 = 0, t1 = 0, t2 = 0, t3 = 0;  // u16 pairs: positions (0,2) (1,3) (4,6) (5,7) of this lane's 8
    int nfl = 0;
    for (int m = 0; m < M; ++m) {
      const int nf = hdr.nf(m);
      nfl += nf;
      u32 a0 = 0, a1 = 0;
      const u8* lmb = s_lm[m];
      if (hdr.flags & 1u) {
        // Strip layout: the lane's 8 bytes of patch row rr (columns g .. g+7, g = feature column + 8*hh) lie in the
        // 16 B window of two 8 B-aligned chunks: chunk k8 = g >> 3 sits in strip k8 >> 1 at byte (k8 & 1) * 8 of
        // row cyT + rr, the next chunk 8 B further or at the start of the next strip's row.  Two LDG.64, one
        // word select and two funnel shifts realign them; a warp's load touches the 4-6 lines of the patch.
        const u32 yrow = (u32)(cyT + rr) * 16u, xh = (u32)cxT + 8u * (u32)hh;
        const uint2* offp = reinterpret_cast<const uint2*>(lp.offs) + (size_t)g * M * FEAT_SLOTS + m * FEAT_SLOTS;
        const int q = WPC ? nf : (nf + 3) >> 2, kb0 = WPC ? 0 : warp * q, ke = min(nf, kb0 + q);  // this warp's share of the features
        for (int kb = kb0; kb < ke; kb += 32) {
        const int kn = min(32, ke - kb);
        __syncwarp();  // the previous reads of s_ent are done
        if (lane < kn) s_ent[warp][lane] = __ldg(offp + kb + lane);
        __syncwarp();
#pragma unroll 4
        for (int kk = 0; kk < kn; ++kk) {
          const uint2 e = s_ent[warp][kk];  // (byte offset of label/phase/row, column) of the feature: broadcast read
          const u32 gcol = e.y + xh, k8 = gcol >> 3, odd = k8 & 1u;
          const u8* p0 = lmb + (e.x + yrow + (k8 >> 1) * H16 + odd * 8u);
          const uint2 lo = __ldg(reinterpret_cast<const uint2*>(p0));
          const uint2 hi = __ldg(reinterpret_cast<const uint2*>(p0 + (odd ? H16 - 8u : 8u)));
          const bool s1 = (gcol & 4u) != 0;
          const u32 v0 = s1 ? lo.y : lo.x, v1 = s1 ? hi.x : lo.y, v2 = s1 ? hi.y : hi.x;
          const u32 sh = (gcol & 3u) * 8u;
          a0 += __funnelshift_r(v0, v1, sh);
          a1 += __funnelshift_r(v1, v2, sh);
        }
        }
      } else if (WPC || warp == 0) {
        const u32* fp = lp.feat + (size_t)g * M * FEAT_SLOTS + m * FEAT_SLOTS;
        for (int k = 0; k < nf; ++k) {
          u32 f = __ldg(fp + k);
          if (!(f >> 31)) continue;
          int fx = (int)(f & 0x3FFF) + cxT * T, fy = (int)((f >> 14) & 0x3FFF) + cyT * T;
          int label = (f >> 28) & 7;
          if (fx < 0 || fy < 0 || fx >= lp.g.cols || fy >= lp.g.rows) continue;
          long long base = (long long)((fy % T) * T + (fx % T)) * plane + (long long)(fy / T) * W + fx / T;
          const u8* lml = lmb + (size_t)label * lp.g.per_label;
          long long rb = base + (long long)rr * W + hh * 8;
          const long long flat_label = (long long)T * T * plane;  // upstream's flat index space of one label
#pragma unroll
          for (int b = 0; b < 8; ++b) {
            u32 v = 0u;
            if (rb + b < flat_label) {  // flat index -> (phase, row, column) -> strip layout
              const u32 f = (u32)(rb + b), ph = f / plane, rem = f - ph * plane, gy = rem / (u32)W, gx = rem - gy * (u32)W;
              v = (u32)lml[(size_t)ph * lp.g.plane + (gx >> 4) * H16 + gy * 16u + (gx & 15u)];
            }
            if (b < 4) a0 += v << (8 * b); else a1 += v << (8 * (b - 4));
          }
        }
      }
      t0 += a0 & 0x00FF00FFu; t1 += (a0 >> 8) & 0x00FF00FFu;
      t2 += a1 & 0x00FF00FFu; t3 += (a1 >> 8) & 0x00FF00FFu;
    }
    if (!WPC) {
      s_part[warp][lane] = make_uint4(t0, t1, t2, t3);
      __syncthreads();
    }
    if (WPC || warp == 0) {
      if (!WPC) {
        const uint4 p1 = s_part[1][lane], p2 = s_part[2][lane], p3 = s_part[3][lane];
        t0 += p1.x + p2.x + p3.x; t1 += p1.y + p2.y + p3.y; t2 += p1.z + p2.z + p3.z; t3 += p1.w + p2.w + p3.w;
      }
    // argmax, strict '>' in raster order from best = 0
    int v[8] = {(int)(t0 & 0xFFFF), (int)(t1 & 0xFFFF), (int)(t0 >> 16), (int)(t1 >> 16),
                (int)(t2 & 0xFFFF), (int)(t3 & 0xFFFF), (int)(t2 >> 16), (int)(t3 >> 16)};
    int best = 0, bi = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b)
      if (v[b] > best) { best = v[b]; bi = b; }
    u32 key = best > 0 ? ((u32)best << 8) | (u32)(255 - (rr * 16 + hh * 8 + bi)) : 0u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) key = max(key, __shfl_xor_sync(0xffffffffu, key, d));
    if (lane == 0) {
      int br = -1, bc = -1, bs = 0;
      if (key) { int idx = 255 - (int)(key & 255u); br = idx >> 4; bc = idx & 15; bs = (int)(key >> 8); }
      rec.x = (cxT + bc) * T + offset;
      rec.y = (cyT + br) * T + offset;
      float sim = __fdiv_rn(__fmul_rn((float)bs, 100.f), (float)(4 * nfl));
      rec.sim = sim < mp.threshold ? -1.0f : sim;
      cands[c] = rec;
      if (rec.sim < 0.f) atomicSub(&mp.tpl_alive[(size_t)frame * mp.nsel_stride + rec.tsel], 1);
      atomicAdd(&mp.ctr[frame].local_bytes, (unsigned long long)nfl * 256ull);
    }
    }
    if (!WPC) __syncthreads();  // s_part is reused by the next candidate
  }
}

void launch_similarity_local(const MatchParams& mp, const LevelParams& lp, cudaStream_t st) {
  if (mp.nsel <= 0 || mp.frames <= 0) return;
  // persistent grid: the candidate count lives on the device, warps stride over it
  int bx = 148 * 8;
  if (mp.frames >= 8) bx = 148 * 2;
  dim3 grid(bx, mp.frames);
  if (mp.frames >= 8) similarity_local_kernel<true><<<grid, 128, 0, st>>>(mp, lp);   // batches: a warp per candidate
  else similarity_local_kernel<false><<<grid, 128, 0, st>>>(mp, lp);                 // latency path: a CTA per candidate
}

// ---------------------------------------------------------------------------------------------
// Ordered compaction: surviving candidates of frame f in (selection order, raster order).
// ---------------------------------------------------------------------------------------------
// The per-template alive counts are maintained by the coarse/local kernels, so the order-defining prefix
// sum needs no candidate reads.  Phase 1 scans the alive counts (1024 templates per round) and lists the
// templates that still own matches with their output base; phase 2 copies those blocks warp-parallel
// (ballot ranks keep the raster order), so no thread ever walks a candidate block serially.
constexpr int PK_LIST = 2048;
__global__ void __launch_bounds__(1024) pack_kernel(MatchParams mp, Cand* __restrict__ out, int out_cap) {
  __shared__ int wsum[32];
  __shared__ int s_run, s_n;
  __shared__ int s_tpl[PK_LIST], s_base[PK_LIST];
  const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (mp.ctr[frame].overflow) return;  // host will grow the store and redo this frame
  const Cand* cands = mp.cand + (size_t)frame * mp.cand_cap;
  Cand* o = out + (size_t)frame * out_cap;
  const int* ts = mp.tpl_start + (size_t)frame * mp.nsel_stride;
  const int* tc = mp.tpl_cnt + (size_t)frame * mp.nsel_stride;
  const int* ta = mp.tpl_alive + (size_t)frame * mp.nsel_stride;
  if (tid == 0) { s_run = 0; s_n = 0; }
  __syncthreads();
  for (int i0 = 0; i0 < mp.nsel; i0 += 1024) {
    const int i = i0 + tid;
    const int alive = i < mp.nsel ? ta[i] : 0;
    int incl = alive;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int v = wsum[lane], inc2 = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, inc2, d);
        if (lane >= d) inc2 += u;
      }
      wsum[lane] = inc2 - v;  // exclusive warp offsets
    }
    __syncthreads();
    const int run = s_run;
    if (alive > 0) {
      const int base = run + wsum[wid] + incl - alive;
      const int slot = atomicAdd(&s_n, 1);
      if (slot < PK_LIST) { s_tpl[slot] = i; s_base[slot] = base; }
      else {  // list full (thousands of matching templates): this thread copies its block itself
        int pos = base;
        const int st = ts[i], cn = tc[i];
        for (int j = 0; j < cn; ++j) {
          Cand cd = cands[st + j];
          if (cd.sim >= 0.f) {
            cd.tsel = mp.sel[cd.tsel];
            if (pos < out_cap) o[pos] = cd;
            ++pos;
          }
        }
      }
    }
    __syncthreads();
    if (tid == 1023) s_run = run + wsum[31] + incl;  // last thread: exclusive offset of its warp + its inclusive sum = round total
    __syncthreads();
  }
  if (tid == 0) mp.ctr[frame].out_count = s_run;
  const int nlist = min(s_n, PK_LIST);
  const u32 lt = (1u << lane) - 1u;
  for (int e = wid; e < nlist; e += 32) {
    const int i = s_tpl[e], st = ts[i], cn = tc[i];
    int pos = s_base[e];
    for (int j0 = 0; j0 < cn; j0 += 32) {
      Cand cd;
      bool keep = false;
      if (j0 + lane < cn) {
        cd = cands[st + j0 + lane];
        keep = cd.sim >= 0.f;
      }
      const u32 bal = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int p = pos + __popc(bal & lt);
        cd.tsel = mp.sel[cd.tsel];  // selection index -> global template index (rank-independent)
        if (p < out_cap) o[p] = cd;
      }
      pos += __popc(bal);
    }
  }
}

// Send buffer of the template-sharded match gather: a COMPACT variable-length packing (one frame with thousands of
// matches must not size every frame's slot), one block per frame:
//   send[2f]     = {out_count, flags (1: candidate store overflowed, 2: the record area is too small), offset, 0}
//   send[2f + 1] = {local_bytes lo, hi, coarse_chunks lo, hi}
//   send[2*frames + offset ...] = the frame's packed matches; offset = sum of the counts of the frames before it.
__global__ void __launch_bounds__(256) gather_pack_kernel(const SlotCtr* __restrict__ ctr, const Cand* __restrict__ out, int out_cap,
                                                          Cand* __restrict__ send, int frames, int rec_cap) {
  __shared__ int s_part[8];
  const int f = blockIdx.x;
  int part = 0;
  for (int i = threadIdx.x; i < f; i += 256) part += min(ctr[i].out_count, out_cap);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
  __syncthreads();
  int offset = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) offset += s_part[w];
  const SlotCtr c = ctr[f];
  const int n = min(c.out_count, out_cap);
  const bool fits = offset + n <= rec_cap;
  if (threadIdx.x == 0) {
    Cand h0, h1;
    h0.tsel = c.out_count; h0.x = ((c.overflow != 0 || c.out_count > out_cap) ? 1 : 0) | (fits ? 0 : 2); h0.y = offset; h0.sim = 0.f;
    h1.tsel = (int)(u32)c.local_bytes; h1.x = (int)(u32)(c.local_bytes >> 32);
    h1.y = (int)(u32)c.coarse_chunks; h1.sim = __int_as_float((int)(u32)(c.coarse_chunks >> 32));
    send[2 * f] = h0; send[2 * f + 1] = h1;
  }
  if (!fits) return;
  const uint4* src = reinterpret_cast<const uint4*>(out + (size_t)f * out_cap);
  uint4* d4 = reinterpret_cast<uint4*>(send + 2 * (size_t)frames + offset);
  for (int i = threadIdx.x; i < n; i += 256) d4[i] = src[i];
}

void launch_gather_pack(const SlotCtr* ctr, const Cand* out, int out_cap, Cand* send, int rec_cap, int frames, cudaStream_t st) {
  if (frames > 0) gather_pack_kernel<<<frames, 256, 0, st>>>(ctr, out, out_cap, send, frames, rec_cap);
}

void launch_pack(const MatchParams& mp, Cand* out, int out_cap, cudaStream_t st) {
  if (mp.frames <= 0) return;
  pack_kernel<<<mp.frames, 1024, 0, st>>>(mp, out, out_cap);
}

}  // namespace lmk
