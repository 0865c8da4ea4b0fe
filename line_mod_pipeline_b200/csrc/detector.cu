// detector.cu — host orchestration of the match path: device template tables, frame plan,
// kernel sequencing on CUDA streams, result fetch + the reference's sort/unique epilogue.
// Mirrors cv::linemod::Detector::match / matchClass (opencv_contrib rgbd/linemod.cpp; SURVEY.md
// Appendix A.6-A.7) as called from the reference at src/HighLevelLinemod.cpp:152.
// There is NO CPU fallback: without a CUDA device every compute entry point fails with
// LMB200_E_NODEVICE.
#include "detector.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <thread>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost nothing unless a profiler is attached

using namespace lmk;

namespace lmh {

// NVTX range per stage of the path (Nsight Systems / ncu --nvtx): frame side, template side, fetch + epilogue, gather
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

void finalize_matches(std::vector<Match>& m) {
  std::sort(m.begin(), m.end());
  m.erase(std::unique(m.begin(), m.end()), m.end());
}

int DevBuf::alloc(size_t n) {
  if (n <= bytes && p && !view) return 0;
  release();
  cudaError_t e = cudaMalloc(&p, n);
  if (e != cudaSuccess) { p = nullptr; bytes = 0; return (int)e; }
  bytes = n;
  return 0;
}
void DevBuf::release() {
  if (p && !view) cudaFree(p);
  p = nullptr; bytes = 0; view = false;
}
void DevBuf::alias(void* ptr, size_t n) {
  release();
  p = ptr; bytes = n; view = true;
}

int set_error(lmb200_detector* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
int cuda_fail(lmb200_detector* h, cudaError_t e, const char* what) {
  return set_error(h, LMB200_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CU(call)                                                   \
  do {                                                             \
    cudaError_t e__ = (call);                                      \
    if (e__ != cudaSuccess) return cuda_fail(h, e__, #call);       \
  } while (0)
#define ALLOC(buf, n)                                              \
  do {                                                             \
    int e__ = (buf).alloc(n);                                      \
    if (e__) return cuda_fail(h, (cudaError_t)e__, "cudaMalloc");  \
  } while (0)

int ensure_device(lmb200_detector* h) {
  if (h->device_ready) {
    cudaSetDevice(h->device);
    return LMB200_OK;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return set_error(h, LMB200_E_NODEVICE, "no CUDA device available (this library has no CPU fallback)");
  }
  int dev = h->cfg.device;
  if (dev < 0) CU(cudaGetDevice(&dev));
  if (dev >= n) return set_error(h, LMB200_E_INVALID, "device ordinal out of range");
  CU(cudaSetDevice(dev));
  h->device = dev;
  int prio_lo = 0, prio_hi = 0;
  CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  for (int i = 0; i < LMB200_LANES; ++i) {
    // lane 1 (copies, NCCL, result fetches) runs at the highest priority: its small kernels and collectives must not queue
    // behind the compute lanes' grids of the next step
    // lane 5 (quantisers + NCCL all-gather of the sharded step's frame side) too: a collective that waits for CTA slots
    // behind a 0.8 ms template-side grid on ONE rank stalls the same collective on every other rank
    CU(cudaStreamCreateWithPriority(&h->lanes[i].stream, cudaStreamNonBlocking, (i == 1 || i == 5) ? prio_hi : prio_lo));
  }
  h->device_ready = true;
  return LMB200_OK;
}

// ---------------------------------------------------------------- profiling helpers
struct ProfScope {
  lmb200_detector* h; int family; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(lmb200_detector* h_, int fam, cudaStream_t s) : h(h_), family(fam), st(s) {
    h->prof.launches[fam]++;
    if (!h->profiling) return;
    auto take = [&]() {
      cudaEvent_t ev;
      if (!h->event_pool.empty()) { ev = h->event_pool.back(); h->event_pool.pop_back(); }
      else cudaEventCreate(&ev);
      return ev;
    };
    a = take(); b = take();
    cudaEventRecord(a, st);
  }
  void finish() {
    if (!a) return;
    cudaEventRecord(b, st);
    h->prof_pending.push_back(ProfRec{family, a, b});
    a = nullptr;
  }
  ~ProfScope() { finish(); }
};

static void collect_profile(lmb200_detector* h) {
  std::vector<ProfRec> keep;
  for (auto& r : h->prof_pending) {
    if (cudaEventQuery(r.b) != cudaSuccess) {  // still in flight on another stream: resolve at a later collect
      cudaGetLastError();
      keep.push_back(r);
      continue;
    }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) h->prof.ms[r.family] += ms;
    h->event_pool.push_back(r.a);
    h->event_pool.push_back(r.b);
  }
  h->prof_pending.swap(keep);
}

// ---------------------------------------------------------------- tables
// First use of a DepthNormal modality on the built-in stand-in NORMAL_LUT: say so (ADVICE r1 — the labels cannot
// equal cv::linemod's until upstream's normal_lut.i is supplied through lmb200_load_normal_lut / _set_normal_lut).
static void warn_standin_normal_lut(lmb200_detector* h) {
  if (h->standin_warned || !h->normal_lut_standin) return;
  bool dn = false;
  for (int m = 0; m < h->cfg.num_modalities; ++m) dn |= h->cfg.modalities[m].type == LMB200_DEPTH_NORMAL;
  if (!dn) return;
  h->standin_warned = true;
  const char* msg = "DepthNormal modality is running on the built-in stand-in NORMAL_LUT: quantized normals (and templates "
                    "extracted from them) differ from cv::linemod's; supply upstream's normal_lut.i with lmb200_load_normal_lut";
  if (!h->warnings.empty()) h->warnings += "; ";
  h->warnings += msg;
  if (!std::getenv("LMB200_QUIET")) std::fprintf(stderr, "[lmb200 warning] %s\n", msg);
}

static int upload_luts(lmb200_detector* h) {
  warn_standin_normal_lut(h);
  if (!h->luts_dirty) return LMB200_OK;
  uint2 table[256];
  for (int s = 0; s < 256; ++s) {
    u32 lo = 0, hi = 0;
    for (int o = 0; o < 8; ++o) {
      u32 v = std::max(h->sim_lut[32 * o + (s & 15)], h->sim_lut[32 * o + 16 + (s >> 4)]);
      if (o < 4) lo |= v << (8 * o); else hi |= v << (8 * (o - 4));
    }
    table[s] = make_uint2(lo, hi);
  }
  ALLOC(h->d_table, sizeof(table));
  ALLOC(h->d_normal_lut, 8000);
  CU(cudaMemcpy(h->d_table.p, table, sizeof(table), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->d_normal_lut.p, h->normal_lut, 8000, cudaMemcpyHostToDevice));
  h->luts_dirty = false;
  h->plan_epoch++;
  return LMB200_OK;
}

static void shard_jobs_quiesce(lmb200_detector* h);
static int rebuild_templates(lmb200_detector* h) {
  if (!h->templates_dirty) return LMB200_OK;
  shard_jobs_quiesce(h);   // the epilogue thread reads g_class / g_tid
  const int M = h->cfg.num_modalities, L = h->cfg.pyramid_levels;
  h->class_list.clear(); h->g_class.clear(); h->g_tid.clear();
  int ci = 0;
  for (auto& kv : h->classes) {
    h->class_list.push_back(kv.first);
    for (size_t t = 0; t < kv.second.size(); ++t) { h->g_class.push_back(ci); h->g_tid.push_back((int)t); }
    ++ci;
  }
  h->ntpl = (int)h->g_class.size();
  h->max_nf_coarse = 0;
  std::vector<TplHdr> hdr((size_t)std::max(1, h->ntpl));
  std::vector<u32> feat((size_t)std::max(1, h->ntpl) * M * FEAT_SLOTS);
  for (int l = 0; l < L; ++l) {
    std::fill(feat.begin(), feat.end(), 0u);
    int g = 0;
    for (auto& kv : h->classes)
      for (auto& tp : kv.second) {
        TplHdr& hd = hdr[g];
        std::memset(&hd, 0, sizeof(hd));
        int nf_sum = 0;
        for (int m = 0; m < M; ++m) {
          const Template& t = tp[(size_t)l * M + m];
          if (t.features.size() > 63)
            return set_error(h, LMB200_E_FEATURES, "template has more than 63 features (upstream CV_Assert)");
          hd.width[m] = (short)std::max(-32768, std::min(32767, t.width));
          hd.height[m] = (short)std::max(-32768, std::min(32767, t.height));
          hd.nf[m] = (u8)t.features.size();
          nf_sum += (int)t.features.size();
          for (size_t k = 0; k < t.features.size(); ++k)
            feat[((size_t)g * M + m) * FEAT_SLOTS + k] = pack_feature(t.features[k].x, t.features[k].y, t.features[k].label);
        }
        if (l == L - 1) h->max_nf_coarse = std::max(h->max_nf_coarse, nf_sum);
        ++g;
      }
    ALLOC(h->d_hdr[l], hdr.size() * sizeof(TplHdr));
    ALLOC(h->d_feat[l], feat.size() * sizeof(u32));
    ALLOC(h->d_offs[l], (size_t)std::max(1, h->ntpl) * M * 2 * (l == L - 1 ? COARSE_SLOTS : FEAT_SLOTS) * sizeof(u32));  // (offset, shift | column) pairs
    CU(cudaMemcpy(h->d_hdr[l].p, hdr.data(), hdr.size() * sizeof(TplHdr), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_feat[l].p, feat.data(), feat.size() * sizeof(u32), cudaMemcpyHostToDevice));
  }
  h->templates_dirty = false;
  h->plan_dirty = true;
  h->sel_key.clear();  // force selection rebuild
  h->h_sel.clear();
  return LMB200_OK;
}

static void free_host_mirrors(lmb200_detector* h) {
  if (h->h_ctr) cudaFreeHost(h->h_ctr);
  if (h->h_out) cudaFreeHost(h->h_out);
  h->h_ctr = nullptr; h->h_out = nullptr;
}

static int alloc_match_buffers(lmb200_detector* h) {
  h->plan_epoch++;
  const int S = h->slots;
  h->nsel_stride = std::max(1, h->ntpl);
  ALLOC(h->d_cand, (size_t)S * h->cand_cap * sizeof(Cand));
  ALLOC(h->d_out, (size_t)S * h->out_cap * sizeof(Cand));
  ALLOC(h->d_ctr, (size_t)S * sizeof(SlotCtr));
  ALLOC(h->d_resp_sum, (size_t)S * MAX_MOD * sizeof(u32));
  ALLOC(h->d_tpl_start, (size_t)S * h->nsel_stride * sizeof(int));
  ALLOC(h->d_tpl_cnt, (size_t)S * h->nsel_stride * sizeof(int));
  ALLOC(h->d_tpl_alive, (size_t)S * h->nsel_stride * sizeof(int));
  free_host_mirrors(h);
  h->h_head = std::min(h->out_cap, 1024);
  CU(cudaHostAlloc((void**)&h->h_ctr, (size_t)S * sizeof(SlotCtr), cudaHostAllocDefault));
  CU(cudaHostAlloc((void**)&h->h_out, (size_t)S * h->out_cap * sizeof(Cand), cudaHostAllocDefault));
  // thresholds survive a reallocation (grow_capacity): ranges matched earlier are re-run with them at fetch time
  if ((int)h->slot_threshold.size() != S) h->slot_threshold.assign(S, 0.f);
  h->slot_gen.assign(S, -1);              // every slot's device results are gone
  return LMB200_OK;
}

static int ensure_plan(lmb200_detector* h, int rows, int cols);

// lmb200_upload_templates: device tables now; the plan too when a frame size is already known
int upload_templates_now(lmb200_detector* h) {
  int rc = ensure_device(h);
  if (rc) return rc;
  if (h->rows > 0 && h->cols > 0) return ensure_plan(h, h->rows, h->cols);
  return rebuild_templates(h);
}

// Frame plan: per-level geometry + every per-slot device buffer.  Re-planned when the frame size changes.
static int ensure_plan(lmb200_detector* h, int rows, int cols) {
  const int M = h->cfg.num_modalities, L = h->cfg.pyramid_levels;
  int rc = upload_luts(h);
  if (rc) return rc;
  rc = rebuild_templates(h);
  if (rc) return rc;
  const int S = h->cfg.max_batch > 0 ? h->cfg.max_batch : 64;
  {
    const char* e = std::getenv("LMB200_GENERIC_FRAME");
    const bool gen = e && *e && *e != '0';
    if (gen != h->generic_frame_side) { h->generic_frame_side = gen; h->rows = 0; }  // re-plan
  }
  bool geom_changed = rows != h->rows || cols != h->cols || S != h->slots || (int)h->levels.size() != L;
  if (geom_changed) {
    // validate like upstream's CV_Asserts (linearize: rows%T, cols%T; computeResponseMaps: (rows*cols)%16)
    int r = rows, c = cols;
    for (int l = 0; l < L; ++l) {
      int T = h->cfg.T[l];
      if (T <= 0 || r <= 0 || c <= 0 || r % T || c % T || (r * c) % 16)
        return set_error(h, LMB200_E_SIZE, "image size at pyramid level " + std::to_string(l) + " (" + std::to_string(c) +
                                                "x" + std::to_string(r) + ") must be divisible by T and rows*cols by 16");
      if (c > 16383 || r > 16383) return set_error(h, LMB200_E_SIZE, "image larger than 16383 pixels per side");
      if (l == L - 1 && (long long)(r / T) * (c / T) >= (1 << 21))
        return set_error(h, LMB200_E_SIZE, "coarsest level has more than 2^21 positions (add a pyramid level or use a larger T)");
      r /= 2; c /= 2;
    }
    CU(cudaDeviceSynchronize());
    for (auto& lb : h->levels)
      for (int m = 0; m < LMB200_MAX_MODALITIES; ++m) { lb.bgr[m].release(); lb.q[m].release(); lb.mask[m].release(); lb.lm[m].release(); lb.lmn[m].release(); }
    h->levels.assign(L, LevelBuffers());
    r = rows; c = cols;
    auto up256 = [](size_t v) { return (v + 255) & ~(size_t)255; };
    // One frame buffer per slot holding the modality sources back to back exactly like a tightly packed host
    // frame (BGR then depth for the reference wiring), so a chunk of host-contiguous frames is ONE H2D copy.
    h->frame_bytes = 0;
    for (int m = 0; m < M; ++m) {
      h->src_off[m] = h->frame_bytes;
      h->frame_bytes += (size_t)rows * cols * (h->cfg.modalities[m].type == LMB200_COLOR_GRADIENT ? 3 : 2);
    }
    ALLOC(h->d_frames, h->frame_bytes * S + 256);
    for (int l = 0; l < L; ++l) {
      LevelBuffers& lb = h->levels[l];
      int T = h->cfg.T[l];
      lb.g.T = T; lb.g.rows = r; lb.g.cols = c; lb.g.W = c / T; lb.g.H = r / T;
      // coarsest level: upstream's flat linear memories; finer levels: 16-column strips (see LevelGeom)
      lb.g.strips = l == L - 1 ? 0 : (lb.g.W + 15) / 16;
      lb.g.plane = lb.g.strips ? (u32)lb.g.strips * lb.g.H * 16u : (u32)lb.g.W * lb.g.H;
      lb.g.per_label = (u32)T * T * lb.g.plane;
      lb.fast_spread = spread_fast_covers(lb.g) && !h->generic_frame_side;
      // DepthNormalPyramid::pyrDown is resize(INTER_NEAREST) to (cols/2, rows/2): src[2y][2x] exactly while the sizes stay even
      lb.dn_inplace = l == 0 ? true : (h->levels[l - 1].dn_inplace && h->levels[l - 1].g.rows % 2 == 0 && h->levels[l - 1].g.cols % 2 == 0);
      lb.q_stride = up256((size_t)r * c);
      lb.lm_stride = up256((size_t)8 * lb.g.per_label + LM_PAD + (lb.g.strips ? (size_t)lb.g.H * 16 : 0));  // + one strip column: the
                                                                 // second chunk of a patch row in the last strip is loaded, never used
      lb.bgr_stride = l == 0 ? h->frame_bytes : up256((size_t)r * c * 3);
      for (int m = 0; m < M; ++m) {
        ALLOC(lb.q[m], lb.q_stride * S);
        ALLOC(lb.lm[m], lb.lm_stride * S);
        CU(cudaMemset(lb.lm[m].p, 0, lb.lm_stride * S));  // the pad bytes stay zero forever
        if (l == L - 1) {  // nibble-packed copy read by similarity_coarse_kernel
          lb.lmn_stride = up256((size_t)4 * r * c + LM_PAD) + coarse_zero_tail(lb.g);  // zero tail: padded plan rows read it
          ALLOC(lb.lmn[m], lb.lmn_stride * S);
          CU(cudaMemset(lb.lmn[m].p, 0, lb.lmn_stride * S));
        }
        if (h->cfg.modalities[m].type == LMB200_COLOR_GRADIENT) {
          if (l == 0) lb.bgr[m].alias((u8*)h->d_frames.p + h->src_off[m], h->frame_bytes * S);
          else ALLOC(lb.bgr[m], lb.bgr_stride * S);
        }
      }
      r /= 2; c /= 2;
    }
    h->dn_materialize = false;
    for (int l = 1; l < L; ++l) h->dn_materialize |= !(h->levels[l].fast_spread && h->levels[l].dn_inplace);
    h->depth_stride = h->frame_bytes / 2;   // u16 elements between the depth images of consecutive slots
    for (int m = 0; m < M; ++m)
      if (h->cfg.modalities[m].type == LMB200_DEPTH_NORMAL) {
        h->d_depth[m].alias((u8*)h->d_frames.p + h->src_off[m], h->frame_bytes * S);
        ALLOC(h->d_dnraw[m], h->levels[0].q_stride * S);
      }
    h->rows = rows; h->cols = cols; h->slots = S;
    h->plan_epoch++;
    if (h->cand_cap <= 0) h->cand_cap = h->cfg.candidate_capacity > 0 ? h->cfg.candidate_capacity : 16384;
    if (h->out_cap <= 0) h->out_cap = h->cand_cap;
    h->nsel_stride = 0;
    h->plan_dirty = true;
    h->masks_in_use = false;
  }
  if (h->nsel_stride < std::max(1, h->ntpl) || !h->d_cand.p) {
    rc = alloc_match_buffers(h);
    if (rc) return rc;
  }
  if (h->plan_dirty) {
    cudaStream_t st = h->lanes[0].stream;
    for (int l = 0; l < L; ++l)
      launch_build_offsets(h->d_feat[l].as<u32>(), h->d_offs[l].as<u32>(), h->d_hdr[l].as<TplHdr>(), h->ntpl, M,
                           h->levels[l].g, l == L - 1, st);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    // algorithmic coarse bytes per template: sum_m nf * P  
    h->tpl_cost.assign(h->ntpl, 0.0);
    const LevelGeom& g = h->levels[L - 1].g;
    int gi = 0;
    for (auto& kv : h->classes)
      for (auto& tp : kv.second) {
        double cost = 0;
        for (int m = 0; m < M; ++m) {
          const Template& t = tp[(size_t)(L - 1) * M + m];
          int wf = (t.width - 1) / g.T + 1, hf = (t.height - 1) / g.T + 1;
          long long P = (long long)(g.H - hf) * g.W + (g.W - wf) + 1;
          if (P > (long long)g.W * g.H) P = (long long)g.W * g.H;
          if (P > 0) cost += (double)P * (double)t.features.size();
        }
        h->tpl_cost[gi++] = cost;
      }
    h->plan_dirty = false;
    h->plan_epoch++;
    h->sel_key.clear();
  }
  return LMB200_OK;
}

}  // namespace lmh

using namespace lmh;

extern "C" int lmb200_shard_plan(const double* costs, int n, int world, int* begin);

// Selection list = templates this handle scores, in generation order (class order as requested,
// template_id ascending), then restricted to this rank's cost-balanced shard.
static int ensure_selection(lmb200_detector* h, const char* const* class_ids, int n_class_ids) {
  std::string key = std::to_string(h->shard_rank) + "/" + std::to_string(h->shard_world) + ":";
  for (int i = 0; i < n_class_ids; ++i) { key += class_ids[i]; key.push_back('\x1f'); }
  if (key == h->sel_key && h->d_sel.p) return LMB200_OK;
  shard_jobs_quiesce(h);   // the epilogue thread reads pos_of_g
  std::vector<int> sel;
  if (n_class_ids <= 0) {
    sel.resize(h->ntpl);
    for (int i = 0; i < h->ntpl; ++i) sel[i] = i;
  } else {
    // first global index of each class
    std::vector<int> first(h->class_list.size() + 1, 0);
    for (int i = 0; i < h->ntpl; ++i) first[h->g_class[i] + 1] = i + 1;
    for (size_t c = 1; c < first.size(); ++c) first[c] = std::max(first[c], first[c - 1]);
    for (int i = 0; i < n_class_ids; ++i) {
      auto it = std::lower_bound(h->class_list.begin(), h->class_list.end(), std::string(class_ids[i]));
      if (it == h->class_list.end() || *it != class_ids[i]) continue;  // unknown ids are skipped like upstream
      int c = (int)(it - h->class_list.begin());
      for (int g = first[c]; g < first[c + 1]; ++g) sel.push_back(g);
    }
  }
  h->shard_interleaved = false;
  if (h->shard_world > 1) {
    // Interleaved shards (rank r scores selection positions r, r+world, ...): neighbouring templates of a class
    // (viewpoints/rotations of one object, the ones that fire together) spread over all ranks, which balances
    // both the coarse gather and the refinement work.  The merge restores generation order by selection position.
    // With a template listed twice (duplicate class ids) positions are ambiguous: fall back to contiguous
    // cost-balanced blocks, whose rank-ordered concatenation is generation order by construction.
    std::vector<int> pos(std::max(1, h->ntpl), -1);
    bool dup = false;
    for (size_t i = 0; i < sel.size(); ++i) { if (pos[sel[i]] >= 0) dup = true; else pos[sel[i]] = (int)i; }
    if (!dup) {
      std::vector<int> mine;
      for (size_t i = (size_t)h->shard_rank; i < sel.size(); i += (size_t)h->shard_world) mine.push_back(sel[i]);
      sel.swap(mine);
      h->pos_of_g.swap(pos);
      h->shard_interleaved = true;
    } else {
      std::vector<double> costs(sel.size());
      for (size_t i = 0; i < sel.size(); ++i) costs[i] = h->tpl_cost[sel[i]] + 1.0;
      std::vector<int> begin(h->shard_world + 1);
      lmb200_shard_plan(costs.data(), (int)sel.size(), h->shard_world, begin.data());
      sel = std::vector<int>(sel.begin() + begin[h->shard_rank], sel.begin() + begin[h->shard_rank + 1]);
    }
  }
  h->h_sel = sel;
  h->sel_bytes_coarse = 0;
  for (int g : sel) h->sel_bytes_coarse += (long long)h->tpl_cost[g];
  ALLOC(h->d_sel, std::max<size_t>(1, sel.size()) * sizeof(int));
  if (!sel.empty()) CU(cudaMemcpy(h->d_sel.p, sel.data(), sel.size() * sizeof(int), cudaMemcpyHostToDevice));
  h->sel_key = key;
  h->plan_epoch++;
  return LMB200_OK;
}

static int check_sources(lmb200_detector* h, const lmb200_image* srcs, int n_sources, int* rows, int* cols) {
  const int M = h->cfg.num_modalities;
  if (n_sources != M) return set_error(h, LMB200_E_SOURCES, "sources.size() != modalities.size()");
  for (int m = 0; m < M; ++m) {
    int want = h->cfg.modalities[m].type == LMB200_COLOR_GRADIENT ? LMB200_8UC3 : LMB200_16UC1;
    if (!srcs[m].data || srcs[m].type != want)
      return set_error(h, LMB200_E_SOURCES, "source " + std::to_string(m) + " must be " + (want == LMB200_8UC3 ? "8UC3" : "16UC1"));
    if (srcs[m].rows != srcs[0].rows || srcs[m].cols != srcs[0].cols)
      return set_error(h, LMB200_E_SOURCES, "sources differ in size");
  }
  *rows = srcs[0].rows; *cols = srcs[0].cols;
  return LMB200_OK;
}

// H2D of one frame's sources into `slot` on `st`.
static int upload_one(lmb200_detector* h, const lmb200_image* srcs, int slot, cudaStream_t st) {
  const int M = h->cfg.num_modalities;
  for (int m = 0; m < M; ++m) {
    const lmb200_image& im = srcs[m];
    if (h->cfg.modalities[m].type == LMB200_COLOR_GRADIENT) {
      size_t rowb = (size_t)im.cols * 3, step = im.step ? im.step : rowb;
      u8* dst = h->levels[0].bgr[m].as<u8>() + (size_t)slot * h->levels[0].bgr_stride;
      if (step == rowb) CU(cudaMemcpyAsync(dst, im.data, rowb * im.rows, cudaMemcpyHostToDevice, st));
      else CU(cudaMemcpy2DAsync(dst, rowb, im.data, step, rowb, im.rows, cudaMemcpyHostToDevice, st));
    } else {
      size_t rowb = (size_t)im.cols * 2, step = im.step ? im.step : rowb;
      u16* dst = h->d_depth[m].as<u16>() + (size_t)slot * h->depth_stride;
      if (step == rowb) CU(cudaMemcpyAsync(dst, im.data, rowb * im.rows, cudaMemcpyHostToDevice, st));
      else CU(cudaMemcpy2DAsync(dst, rowb, im.data, step, rowb, im.rows, cudaMemcpyHostToDevice, st));
    }
  }
  return LMB200_OK;
}

// True when the chunk's host images are tightly packed and laid out frame after frame exactly like the device
// frame buffer (source 0, source 1, ... of frame f, then frame f+1): the chunk can then be copied by one memcpy.
static bool chunk_is_contiguous(lmb200_detector* h, const lmb200_image* frames, int cnt, int n_sources) {
  const char* base = (const char*)frames[0].data;
  for (int i = 0; i < cnt; ++i)
    for (int m = 0; m < n_sources; ++m) {
      const lmb200_image& im = frames[(size_t)i * n_sources + m];
      size_t rowb = (size_t)im.cols * (im.type == LMB200_8UC3 ? 3 : 2);
      if (im.step != 0 && im.step != rowb) return false;
      if ((const char*)im.data != base + (size_t)i * h->frame_bytes + h->src_off[m]) return false;
    }
  return true;
}

// Frame side for slots [first, first+count): quantise every modality at every level, then
// spread + response + linearize.  (Detector::match, first half.)
// phase: FS_ALL, or one half of it — FS_QUANTIZE (pyrDown + the quantizers: everything that writes quantized maps) /
// FS_SPREAD (spread + response + linearize).  The template-sharded multi-GPU step quantises only the rank's own frame
// block, all-gathers the quantized maps over NCCL and then spreads every frame (lmb200_match_resident_sharded).
enum { FS_ALL = 0, FS_QUANTIZE = 1, FS_SPREAD = 2 };
// only_m >= 0: that modality's chain alone (level order), without the response-sum memset — the single-frame graph runs
// the modalities as parallel branches (lmb200_match).
static int run_frame_side(lmb200_detector* h, int first, int count, cudaStream_t st, int phase = FS_ALL, int only_m = -1) {
  NvtxRange nvtx(phase == FS_QUANTIZE ? "lmb200:quantize" : phase == FS_SPREAD ? "lmb200:spread_linearize" : "lmb200:frame_side");
  const int M = h->cfg.num_modalities, L = h->cfg.pyramid_levels;
  if (phase != FS_QUANTIZE && only_m < 0)
    CU(cudaMemsetAsync(h->d_resp_sum.as<u32>() + (size_t)first * MAX_MOD, 0, (size_t)count * MAX_MOD * sizeof(u32), st));
  for (int l = 0; l < L; ++l) {
    LevelBuffers& lb = h->levels[l];
    for (int m = 0; m < M; ++m) {
      if (only_m >= 0 && m != only_m) continue;
      const lmb200_modality& mod = h->cfg.modalities[m];
      u8* q = lb.q[m].as<u8>() + (size_t)first * lb.q_stride;
      if (phase == FS_SPREAD) {
      } else if (mod.type == LMB200_COLOR_GRADIENT) {
        u8* bgr = lb.bgr[m].as<u8>() + (size_t)first * lb.bgr_stride;
        const bool gen = h->generic_frame_side;   // round 1's kernels keep the coarser levels interleaved, round 2's as byte planes
        if (l > 0) {
          LevelBuffers& pb = h->levels[l - 1];
          ProfScope ps(h, LMB200_K_PYRDOWN, st);
          if (gen)
            launch_pyrdown_bgr(pb.bgr[m].as<u8>() + (size_t)first * pb.bgr_stride, pb.bgr_stride, bgr, lb.bgr_stride,
                               pb.g.rows, pb.g.cols, count, st);
          else
            launch_pyrdown_planar(pb.bgr[m].as<u8>() + (size_t)first * pb.bgr_stride, pb.bgr_stride, l - 1 > 0, bgr, lb.bgr_stride,
                                  pb.g.rows, pb.g.cols, count, st);
        }
        ProfScope ps(h, LMB200_K_CG_QUANTIZE, st);
        if (gen)
          launch_cg_quantize(bgr, lb.bgr_stride, q, lb.q_stride, nullptr, 0, lb.g.rows, lb.g.cols,
                             mod.weak_threshold * mod.weak_threshold, count, st);
        else
          launch_cg_quantize2(bgr, lb.bgr_stride, l > 0, q, lb.q_stride, nullptr, 0, lb.g.rows, lb.g.cols,
                              mod.weak_threshold * mod.weak_threshold, count, st);
      } else {
        if (l == 0 && !h->generic_frame_side) {
          ProfScope ps(h, LMB200_K_DN_QUANTIZE, st);
          launch_dn_median(h->d_depth[m].as<u16>() + (size_t)first * h->depth_stride, h->depth_stride, q, lb.q_stride,
                           lb.g.rows, lb.g.cols, mod.distance_threshold, mod.difference_threshold, h->d_normal_lut.as<u8>(), count, st);
        } else if (l == 0) {
          u8* raw = h->d_dnraw[m].as<u8>() + (size_t)first * lb.q_stride;
          {
            ProfScope ps(h, LMB200_K_DN_QUANTIZE, st);
            launch_dn_quantize(h->d_depth[m].as<u16>() + (size_t)first * h->depth_stride, h->depth_stride, raw,
                               lb.q_stride, nullptr, lb.g.rows, lb.g.cols, mod.distance_threshold,
                               mod.difference_threshold, h->d_normal_lut.as<u8>(), count, st);
          }
          ProfScope ps(h, LMB200_K_MEDIAN, st);
          launch_median5(raw, lb.q_stride, q, lb.q_stride, lb.g.rows, lb.g.cols, count, st);
        } else if (h->dn_materialize) {  // some level needs the decimated maps in memory: write the whole chain
          LevelBuffers& pb = h->levels[l - 1];
          ProfScope ps(h, LMB200_K_DECIMATE, st);
          launch_resize_nn(pb.q[m].as<u8>() + (size_t)first * pb.q_stride, pb.q_stride, pb.g.rows, pb.g.cols, q,
                           lb.q_stride, lb.g.rows, lb.g.cols, count, st);
        }
      }
      if (phase == FS_QUANTIZE) continue;
      const u8* mask = nullptr;
      if (h->masks_in_use && lb.mask[m].p) mask = lb.mask[m].as<u8>() + (size_t)first * lb.q_stride;
      ProfScope ps(h, LMB200_K_LINEARIZE, st);
      if (lb.fast_spread) {
        SpreadArgs a;
        a.q = q; a.q_stride = lb.q_stride; a.q_pitch = lb.g.cols; a.q_step = 1;
        if (mod.type == LMB200_DEPTH_NORMAL && l > 0 && !h->dn_materialize) {  // read the level-0 map with stride 2^l
          LevelBuffers& l0 = h->levels[0];
          a.q = l0.q[m].as<u8>() + (size_t)first * l0.q_stride; a.q_stride = l0.q_stride; a.q_pitch = l0.g.cols; a.q_step = 1 << l;
        }
        a.mask = mask; a.mask_stride = lb.q_stride;
        a.lm = lb.lm[m].as<u8>() + (size_t)first * lb.lm_stride; a.lm_stride = lb.lm_stride;
        a.lmn = l == L - 1 ? lb.lmn[m].as<u8>() + (size_t)first * lb.lmn_stride : nullptr; a.lmn_stride = lb.lmn_stride;
        a.resp_sum = h->d_resp_sum.as<u32>() + (size_t)first * MAX_MOD + m; a.resp_stride = MAX_MOD;
        a.g = lb.g; a.table = h->d_table.as<uint2>();
        if (!launch_spread_fast(a, count, st)) return set_error(h, LMB200_E_INVALID, "internal: fast spread kernel does not cover this level");
        continue;
      }
      launch_spread_linearize(q, lb.q_stride, mask, lb.q_stride, lb.lm[m].as<u8>() + (size_t)first * lb.lm_stride,
                              lb.lm_stride, lb.g, h->d_table.as<uint2>(), count, st);
      if (l == L - 1) h->prof.launches[LMB200_K_LINEARIZE]++;  // the nibble packer is a launch of its own
      if (l == L - 1)
        launch_pack_nibbles(lb.lm[m].as<u8>() + (size_t)first * lb.lm_stride, lb.lm_stride,
                            lb.lmn[m].as<u8>() + (size_t)first * lb.lmn_stride, lb.lmn_stride, lb.g, count,
                            h->d_resp_sum.as<u32>() + (size_t)first * MAX_MOD + m, MAX_MOD, st);
    }
  }
  CU(cudaGetLastError());
  return LMB200_OK;
}

// The fast spread kernels read DepthNormal's coarser levels out of the level-0 map in place; callers that want those
// quantized maps themselves (quantized_images of match(), lmb200_debug_fetch) get them materialised here.
static int materialize_dn_levels(lmb200_detector* h, int first, int count, cudaStream_t st) {
  const int M = h->cfg.num_modalities, L = h->cfg.pyramid_levels;
  for (int m = 0; m < M; ++m) {
    if (h->cfg.modalities[m].type != LMB200_DEPTH_NORMAL) continue;
    for (int l = 1; l < L; ++l) {
      LevelBuffers& lb = h->levels[l];
      LevelBuffers& pb = h->levels[l - 1];
      if (h->dn_materialize) continue;  // run_frame_side wrote the chain already
      launch_resize_nn(pb.q[m].as<u8>() + (size_t)first * pb.q_stride, pb.q_stride, pb.g.rows, pb.g.cols,
                       lb.q[m].as<u8>() + (size_t)first * lb.q_stride, lb.q_stride, lb.g.rows, lb.g.cols, count, st);
    }
  }
  CU(cudaGetLastError());
  return LMB200_OK;
}

static MatchParams make_match_params(lmb200_detector* h, int first, int count, float threshold) {
  MatchParams mp;
  mp.M = h->cfg.num_modalities;
  mp.nsel = (int)h->h_sel.size();
  mp.frames = count;
  mp.sel = h->d_sel.as<int>();
  mp.threshold = threshold;
  mp.early_exit = h->early_exit ? 1 : 0;
  mp.cand = h->d_cand.as<Cand>() + (size_t)first * h->cand_cap;
  mp.cand_cap = h->cand_cap;
  mp.ctr = h->d_ctr.as<SlotCtr>() + first;
  mp.nsel_stride = h->nsel_stride;
  mp.tpl_start = h->d_tpl_start.as<int>() + (size_t)first * h->nsel_stride;
  mp.tpl_cnt = h->d_tpl_cnt.as<int>() + (size_t)first * h->nsel_stride;
  mp.tpl_alive = h->d_tpl_alive.as<int>() + (size_t)first * h->nsel_stride;
  return mp;
}

// nibble = true: the coarsest level's nibble-packed linear memories (similarity_coarse_kernel)
static LevelParams make_level_params(lmb200_detector* h, int l, int first, bool nibble = false) {
  LevelParams lp;
  LevelBuffers& lb = h->levels[l];
  lp.g = lb.g;
  for (int m = 0; m < MAX_MOD; ++m) {
    int mm = m < h->cfg.num_modalities ? m : 0;
    if (nibble) {
      lp.lm[m] = lb.lmn[mm].as<u8>() + (size_t)first * lb.lmn_stride;
      lp.lm_stride[m] = lb.lmn_stride;
    } else {
      lp.lm[m] = lb.lm[mm].as<u8>() + (size_t)first * lb.lm_stride;
      lp.lm_stride[m] = lb.lm_stride;
    }
  }
  lp.hdr = h->d_hdr[l].as<TplHdr>();
  lp.offs = h->d_offs[l].as<u32>();
  lp.feat = h->d_feat[l].as<u32>();
  lp.resp_sum = h->d_resp_sum.as<u32>() + (size_t)first * MAX_MOD;
  return lp;
}

// Template side for slots [first, first+count): coarse similarity + threshold, refinement per level,
// ordered compaction.  (Detector::matchClass.)  stop_after_coarse: debug path.
static int run_matching(lmb200_detector* h, int first, int count, float threshold, cudaStream_t st,
                        bool stop_after_coarse = false) {
  NvtxRange nvtx("lmb200:matchClass(similarity, similarityLocal, pack)");
  const int L = h->cfg.pyramid_levels;
  MatchParams mp = make_match_params(h, first, count, threshold);
  CU(cudaMemsetAsync(mp.ctr, 0, sizeof(SlotCtr) * count, st));
  if (mp.nsel > 0) {
    {
      ProfScope ps(h, LMB200_K_SIM_COARSE, st);
      launch_similarity_coarse(mp, make_level_params(h, L - 1, first, true), 4 * h->max_nf_coarse > 255, st);
    }
    if (!stop_after_coarse)
      for (int l = L - 2; l >= 0; --l) {
        ProfScope ps(h, LMB200_K_SIM_LOCAL, st);
        launch_similarity_local(mp, make_level_params(h, l, first), st);
      }
    ProfScope ps(h, LMB200_K_PACK, st);
    launch_pack(mp, h->d_out.as<Cand>() + (size_t)first * h->out_cap, h->out_cap, st);
  }
  CU(cudaGetLastError());
  for (int i = 0; i < count; ++i) { h->slot_threshold[first + i] = threshold; h->slot_gen[first + i] = h->buffer_generation; }
  h->prof.frames += count;
  h->prof.bytes_coarse += (long long)count * h->sel_bytes_coarse;
  return LMB200_OK;
}

static int grow_capacity(lmb200_detector* h) {
  shard_jobs_quiesce(h);
  CU(cudaDeviceSynchronize());
  if (h->cand_cap > (1 << 24)) return set_error(h, LMB200_E_CUDA, "candidate buffer overflow persists after growing");
  h->cand_cap *= 4; h->out_cap = h->cand_cap;
  h->buffer_generation++;  // batches in flight were computed into the old stores: their tickets rerun at collect
  return alloc_match_buffers(h);
}

// D2H of counts + list heads, then per-frame record vectors in generation order (global template
// index in .tsel).  Returns +1 (nothing consumed) when a device-side store overflowed: the caller
// grows the stores (grow_capacity) and redoes the template side — linear memories stay resident.
// D2H of the counters and list heads of slots [first, first+count) (also captured into the single-frame CUDA graph)
static int enqueue_result_copies(lmb200_detector* h, int first, int count, cudaStream_t st) {
  CU(cudaMemcpyAsync(h->h_ctr + first, h->d_ctr.as<SlotCtr>() + first, sizeof(SlotCtr) * count, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpy2DAsync(h->h_out + (size_t)first * h->out_cap, (size_t)h->out_cap * sizeof(Cand),
                       h->d_out.as<Cand>() + (size_t)first * h->out_cap, (size_t)h->out_cap * sizeof(Cand),
                       (size_t)h->h_head * sizeof(Cand), count, cudaMemcpyDeviceToHost, st));
  return LMB200_OK;
}

static int fetch_raw(lmb200_detector* h, int first, int count, cudaStream_t st, std::vector<std::vector<Cand>>& out, bool copies_enqueued = false) {
  NvtxRange nvtx("lmb200:fetch(D2H match lists)");
  if (!copies_enqueued) {
    int rc = enqueue_result_copies(h, first, count, st);
    if (rc) return rc;
  }
  CU(cudaStreamSynchronize(st));
  for (int i = 0; i < count; ++i)
    if (h->h_ctr[first + i].overflow != 0 || h->h_ctr[first + i].out_count > h->out_cap) return 1;
  bool any_tail = false;
  for (int i = 0; i < count; ++i) {
    int n = h->h_ctr[first + i].out_count;
    if (n > h->h_head) {
      any_tail = true;
      CU(cudaMemcpyAsync(h->h_out + (size_t)(first + i) * h->out_cap + h->h_head,
                         h->d_out.as<Cand>() + (size_t)(first + i) * h->out_cap + h->h_head,
                         (size_t)(n - h->h_head) * sizeof(Cand), cudaMemcpyDeviceToHost, st));
    }
  }
  if (any_tail) CU(cudaStreamSynchronize(st));
  out.resize(count);
  for (int i = 0; i < count; ++i) {
    int n = h->h_ctr[first + i].out_count;
    const Cand* src = h->h_out + (size_t)(first + i) * h->out_cap;
    out[i].assign(src, src + n);
    h->prof.bytes_local += (long long)h->h_ctr[first + i].local_bytes;
    h->prof.chunks_coarse += (long long)h->h_ctr[first + i].coarse_chunks;
  }
  if (h->profiling) collect_profile(h);
  return LMB200_OK;
}

// Single-lane fetch: on overflow grow the stores and redo the template side of [first, first+count).
// Slots whose results were computed into stores that have since been reallocated (another range overflowed and
// grew them) are matched again before they are read: the linear memories are still resident.
static int revalidate_slots(lmb200_detector* h, int first, int count) {
  for (int i = 0; i < count;) {
    if (h->slot_gen[first + i] == h->buffer_generation) { ++i; continue; }
    int j = i + 1;  // maximal stale run with one threshold
    while (j < count && h->slot_gen[first + j] != h->buffer_generation && h->slot_threshold[first + j] == h->slot_threshold[first + i]) ++j;
    int rc = run_matching(h, first + i, j - i, h->slot_threshold[first + i], h->lanes[0].stream);
    if (rc) return rc;
    CU(cudaStreamSynchronize(h->lanes[0].stream));
    i = j;
  }
  return LMB200_OK;
}

static int fetch_grow(lmb200_detector* h, int first, int count, cudaStream_t st, std::vector<std::vector<Cand>>& out, bool copies_enqueued = false) {
  {
    int rc = revalidate_slots(h, first, count);
    if (rc) return rc;
  }
  for (;;) {
    int rc = fetch_raw(h, first, count, st, out, copies_enqueued);
    copies_enqueued = false;
    if (rc <= 0) return rc;
    float thr = h->slot_threshold[first];
    rc = grow_capacity(h);
    if (rc) return rc;
    rc = run_matching(h, first, count, thr, h->lanes[0].stream);
    if (rc) return rc;
    CU(cudaStreamSynchronize(h->lanes[0].stream));
  }
}

static void to_matches(lmb200_detector* h, const std::vector<Cand>& raw, std::vector<Match>& m) {
  m.resize(raw.size());
  for (size_t i = 0; i < raw.size(); ++i) {
    int g = raw[i].tsel;
    if (g < 0 || g >= h->ntpl) { m[i] = Match{raw[i].x, raw[i].y, raw[i].sim, -1, -1}; continue; }  // never index with a bad record
    m[i] = Match{raw[i].x, raw[i].y, raw[i].sim, h->g_class[g], h->g_tid[g]};
  }
}

static int emit(lmb200_detector* h, const std::vector<Match>& m, lmb200_match_rec* out, size_t cap, size_t base,
                size_t* written) {
  size_t n = m.size();
  size_t room = cap > base ? cap - base : 0;
  size_t w = std::min(n, room);
  for (size_t i = 0; i < w; ++i) {
    out[base + i].x = m[i].x; out[base + i].y = m[i].y; out[base + i].similarity = m[i].similarity;
    out[base + i].class_index = m[i].class_index; out[base + i].template_id = m[i].template_id;
  }
  *written = n;
  (void)h;
  return w < n ? LMB200_E_TRUNCATED : LMB200_OK;
}

// The slot-addressed entry points (match, upload/match_resident) share slots with batches in flight.
static void drain_tickets(lmb200_detector* h) {
  if (!h->device_ready || !(h->tickets[0].active || h->tickets[1].active)) return;
  for (int i = 0; i < LMB200_LANES; ++i) cudaStreamSynchronize(h->lanes[i].stream);
}

// Host epilogue of n independent frames (record -> Match, std::sort, std::unique) on a few threads.
static void finalize_frames(lmb200_detector* h, const std::vector<std::vector<Cand>>& raws, std::vector<std::vector<Match>>& outs, unsigned max_threads = 0u) {
  if (max_threads == 0u) max_threads = (unsigned)std::max(1, h->host_threads);
  NvtxRange nvtx("lmb200:epilogue(std::sort, std::unique)");
  const int n = (int)raws.size();
  outs.resize(n);
  auto work = [&](int lo, int hi) {
    for (int i = lo; i < hi; ++i) { to_matches(h, raws[i], outs[i]); finalize_matches(outs[i]); }
  };
  size_t total = 0;
  for (auto& r : raws) total += r.size();
  int nt = (int)std::min<size_t>(std::min<unsigned>(max_threads, std::max(1u, std::thread::hardware_concurrency())), total / 2048 + 1);
  if (nt <= 1 || n < 2) { work(0, n); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; ++t) pool.emplace_back(work, (int)((long long)n * t / nt), (int)((long long)n * (t + 1) / nt));
  for (auto& th : pool) th.join();
}

static int prepare(lmb200_detector* h, const lmb200_image* frames, int n_frames, int n_sources,
                   const char* const* class_ids, int n_class_ids) {
  if (!h || !frames || n_frames <= 0) return set_error(h, LMB200_E_INVALID, "bad arguments");
  int rc = ensure_device(h);
  if (rc) return rc;
  int rows = 0, cols = 0;
  for (int f = 0; f < n_frames; ++f) {
    int r, c;
    rc = check_sources(h, frames + (size_t)f * n_sources, n_sources, &r, &c);
    if (rc) return rc;
    if (f == 0) { rows = r; cols = c; }
    else if (r != rows || c != cols) return set_error(h, LMB200_E_SOURCES, "frames of one batch differ in size");
  }
  rc = ensure_plan(h, rows, cols);
  if (rc) return rc;
  return ensure_selection(h, class_ids, n_class_ids);
}

static_assert(sizeof(lmk::EpiMatch) == sizeof(lmb200_match_rec), "the device epilogue writes lmb200_match_rec records");

namespace lmh {
// ---------------------------------------------------------------- template-sharded steps: host epilogue off the critical path
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Gathered match buffers of all ranks (host memory; per rank [2*count header records][gcap match records], see
// gather_pack_kernel) -> the finished lists of every frame: reference generation order restored (rank-ordered
// concatenation for contiguous shards; ordered by selection position for interleaved shards: every template lives on
// exactly one rank and its records are already in raster order), then the 1-GPU epilogue (record -> Match, std::sort,
// std::unique), so the result is sequence-identical to the unsharded run.
static void merge_gathered(lmb200_detector* h, const Cand* G, int gcap, int count, unsigned max_threads,
                           std::vector<lmb200_match_rec>& recs, std::vector<size_t>& offs, long long* candidates, long long* matches,
                           double* t_reorder, double* t_sort) {
  const int world = h->comm_world;
  const double t0 = now_ms();
  auto rank_base = [&](int r) { return G + (size_t)r * ((size_t)2 * count + gcap); };
  std::vector<std::vector<Cand>> alls(count);
  auto gather_frames = [&](int a, int b) {
    for (int i = a; i < b; ++i) {
      std::vector<Cand>& all = alls[i];
      size_t tot = 0;
      for (int r = 0; r < world; ++r) tot += (size_t)rank_base(r)[2 * i].tsel;
      all.reserve(tot);
      for (int r = 0; r < world; ++r) {
        const Cand* rb = rank_base(r);
        const Cand* rec = rb + 2 * (size_t)count + rb[2 * i].y;
        all.insert(all.end(), rec, rec + rb[2 * i].tsel);
      }
      if (h->shard_interleaved)
        std::stable_sort(all.begin(), all.end(), [h](const Cand& a, const Cand& b) { return h->pos_of_g[a.tsel] < h->pos_of_g[b.tsel]; });
    }
  };
  const int nt = count >= 16 ? (int)std::min(4u, std::max(1u, max_threads)) : 1;
  if (nt == 1) gather_frames(0, count);
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t) pool.emplace_back(gather_frames, count * t / nt, count * (t + 1) / nt);
    for (auto& th : pool) th.join();
  }
  const double t1 = now_ms();
  std::vector<std::vector<Match>> ms;
  finalize_frames(h, alls, ms, max_threads);
  size_t total = 0;
  for (auto& m : ms) total += m.size();
  recs.resize(total); offs.resize((size_t)count + 1);
  size_t base = 0;
  long long nc = 0;
  for (int i = 0; i < count; ++i) {
    offs[i] = base;
    nc += (long long)alls[i].size();
    for (const Match& mt : ms[i]) {
      lmb200_match_rec& r = recs[base++];
      r.x = mt.x; r.y = mt.y; r.similarity = mt.similarity; r.class_index = mt.class_index; r.template_id = mt.template_id;
    }
  }
  offs[count] = base;
  if (candidates) *candidates = nc;
  if (matches) *matches = (long long)total;
  if (t_reorder) *t_reorder = t1 - t0;
  if (t_sort) *t_sort = now_ms() - t1;
}

// any header flag set (1: a candidate store overflowed on some rank, 2: the record area was too small) -> synchronous path
static bool gathered_needs_redo(const Cand* G, int gcap, int count, int world) {
  for (int r = 0; r < world; ++r)
    for (int i = 0; i < count; ++i)
      if (G[(size_t)r * ((size_t)2 * count + gcap) + 2 * i].x) return true;
  return false;
}

static void shard_epilogue_loop(lmb200_detector* h) {
  cudaSetDevice(h->device);
  for (;;) {
    ShardJob* job = nullptr;
    {
      std::unique_lock<std::mutex> lk(h->epi_mu);
      h->epi_cv.wait(lk, [h] { return h->epi_stop || !h->epi_queue.empty(); });
      if (h->epi_queue.empty()) return;       // stop requested and nothing left
      job = h->epi_queue.front();
    }
    const double t0 = now_ms();
    int state = 3;
    if (cudaEventSynchronize(job->ev) == cudaSuccess && !gathered_needs_redo(job->host, job->cap, job->count, h->comm_world)) {
      job->t_wait = now_ms() - t0;
      const Cand* mine = job->host + (size_t)h->comm_rank * ((size_t)2 * job->count + job->cap);
      job->bytes_local = 0; job->chunks_coarse = 0;
      for (int i = 0; i < job->count; ++i) {
        const Cand& st1 = mine[2 * i + 1];
        job->bytes_local += (long long)(((unsigned long long)(u32)st1.x << 32) | (u32)st1.tsel);
        job->chunks_coarse += (long long)(((unsigned long long)(u32)__builtin_bit_cast(int, st1.sim) << 32) | (u32)st1.y);
      }
      // few helper threads: one process per GPU shares the host with its peers
      merge_gathered(h, job->host, job->cap, job->count, 3u, job->recs, job->offs, &job->candidates, &job->matches, &job->t_reorder, &job->t_sort);
      state = 2;
    }
    {
      std::lock_guard<std::mutex> lk(h->epi_mu);
      job->state = state;
      h->epi_queue.pop_front();
    }
    h->epi_done_cv.notify_all();
  }
}

static void shard_job_submit(lmb200_detector* h, ShardJob* job) {
  {
    std::lock_guard<std::mutex> lk(h->epi_mu);
    if (!h->epi_thread.joinable()) { h->epi_stop = false; h->epi_thread = std::thread(shard_epilogue_loop, h); }
    job->state = 1;
    h->epi_queue.push_back(job);
  }
  h->epi_cv.notify_one();
}

// wait until the epilogue thread is done with this entry (no-op when it is idle or already finished)
static void shard_job_wait(lmb200_detector* h, ShardJob* job) {
  std::unique_lock<std::mutex> lk(h->epi_mu);
  h->epi_done_cv.wait(lk, [job] { return job->state != 1; });
}
static void shard_job_retire(lmb200_detector* h, ShardJob* job) {
  shard_job_wait(h, job);
  job->state = 0;
}
static void shard_jobs_quiesce(lmb200_detector* h) {
  if (!h->epi_thread.joinable()) return;
  std::unique_lock<std::mutex> lk(h->epi_mu);
  h->epi_done_cv.wait(lk, [h] { return h->epi_queue.empty(); });
}
void shard_jobs_stop(lmb200_detector* h) {
  if (!h->epi_thread.joinable()) return;
  {
    std::lock_guard<std::mutex> lk(h->epi_mu);
    h->epi_stop = true;
  }
  h->epi_cv.notify_all();
  h->epi_thread.join();
  for (auto& j : h->jobs) j.state = 0;
}

}  // namespace lmh

extern "C" {

int lmb200_upload_frames(lmb200_handle h, const lmb200_image* frames, int n_frames, int n_sources, int first_slot) {
  if (h) drain_tickets(h);
  int rc = prepare(h, frames, n_frames, n_sources, nullptr, 0);
  if (rc) return rc;
  if (first_slot < 0 || first_slot + n_frames > h->slots) return set_error(h, LMB200_E_INVALID, "slot range exceeds max_batch");
  // "upload_async": the copies run on a lane of their own, behind the last match that read these slots, and the next
  // match on the compute lane waits for them — H2D of step k+1 overlaps the kernels of step k
  cudaStream_t st = h->upload_async ? h->lanes[2].stream : h->lanes[0].stream;
  if (h->upload_async)
    for (auto& mk : h->resident_marks)
      if (mk.ev && mk.first < first_slot + n_frames && first_slot < mk.first + mk.count) CU(cudaStreamWaitEvent(st, mk.ev, 0));
  for (int f = 0; f < n_frames; ++f) {
    ProfScope ps(h, LMB200_K_UPLOAD, st);
    rc = upload_one(h, frames + (size_t)f * n_sources, first_slot + f, st);
    if (rc) return rc;
  }
  if (h->upload_async) {
    if (!h->upload_ev) CU(cudaEventCreateWithFlags(&h->upload_ev, cudaEventDisableTiming));
    CU(cudaEventRecord(h->upload_ev, st));
    h->upload_pending = true;
  } else {
    CU(cudaStreamSynchronize(st));
  }
  return LMB200_OK;
}

// The ring keeps the completion marks of the last four resident steps; lanes that run AHEAD of the compute lane (frame
// lane, asynchronous upload lane) order themselves behind earlier readers of a slot range through these marks.  A mark
// that leaves the ring (a fifth distinct slot range in flight) can no longer be found by range, so both lanes are made to
// wait for it here: they execute in order, hence everything they do later is behind that step as well.
static cudaError_t evict_mark(lmb200_detector* h, const ResidentMark& mk) {
  cudaError_t e = cudaStreamWaitEvent(h->lanes[5].stream, mk.ev, 0);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(h->lanes[2].stream, mk.ev, 0);
  return e;
}

// the compute lane picks up frames uploaded asynchronously
static int wait_async_upload(lmb200_detector* h) {
  if (h->upload_pending) {
    CU(cudaStreamWaitEvent(h->lanes[0].stream, h->upload_ev, 0));
    h->upload_pending = false;
  }
  return LMB200_OK;
}

int lmb200_match_resident(lmb200_handle h, int first_slot, int count, float threshold,
                          const char* const* class_ids, int n_class_ids) {
  if (!h || !h->device_ready || h->rows == 0) return set_error(h, LMB200_E_INVALID, "no frames uploaded");
  if (first_slot < 0 || count <= 0 || first_slot + count > h->slots) return set_error(h, LMB200_E_INVALID, "bad slot range");
  cudaSetDevice(h->device);
  drain_tickets(h);
  int rc = ensure_plan(h, h->rows, h->cols);
  if (rc) return rc;
  rc = ensure_selection(h, class_ids, n_class_ids);
  if (rc) return rc;
  cudaStream_t st = h->lanes[0].stream;
  // "resident_overlap" (default on): the frame side runs on the high-priority frame lane and the template side on the
  // compute lane, chained by an event — consecutive calls on DIFFERENT slot ranges (double buffering) then overlap the
  // issue-bound quantisers of step k+1 with the latency-bound similarity kernels of step k.  Calls on the same range
  // serialise through the range's completion mark, exactly as before.
  // (profiling keeps one lane: per-kernel CUDA-event times are only meaningful without a concurrent kernel)
  cudaStream_t sf = h->resident_overlap && !h->profiling && count >= 8 ? h->lanes[5].stream : st;
  if (sf != st) {
    if (h->upload_pending) {
      CU(cudaStreamWaitEvent(sf, h->upload_ev, 0));
      h->upload_pending = false;
    }
    for (auto& mk : h->resident_marks)   // earlier steps still reading (or, for the frame buffers, uploads ordered behind) these slots
      if (mk.ev && mk.first < first_slot + count && first_slot < mk.first + mk.count) CU(cudaStreamWaitEvent(sf, mk.ev, 0));
    if (!h->fs_done) CU(cudaEventCreateWithFlags(&h->fs_done, cudaEventDisableTiming));
  } else {
    rc = wait_async_upload(h);
    if (rc) return rc;
  }
  rc = run_frame_side(h, first_slot, count, sf);
  if (rc) return rc;
  if (sf != st) {
    CU(cudaEventRecord(h->fs_done, sf));
    CU(cudaStreamWaitEvent(st, h->fs_done, 0));
  }
  rc = run_matching(h, first_slot, count, threshold, st);
  if (rc) return rc;
  // completion event of this slot range: the fetch calls wait on it from the copy stream, so a fetch of step k does
  // not queue behind the kernels of step k+1 that were already enqueued on the compute stream
  ResidentMark& mk = h->resident_marks[h->resident_next++ & 3];
  if (!mk.ev) CU(cudaEventCreateWithFlags(&mk.ev, cudaEventDisableTiming));
  else CU(evict_mark(h, mk));
  mk.first = first_slot; mk.count = count;
  CU(cudaEventRecord(mk.ev, st));
  return LMB200_OK;
}

// Template-sharded step with the frame side sharded too (one process per GPU, after lmb200_set_template_shard +
// lmb200_comm_init; every rank holds the same frames in slots [first, first+count) — or at least its own block of them):
// rank r quantises frames [first + r*n, first + (r+1)*n), n = count / world; the quantized maps (1 B per pixel, level and
// modality: 0.7 MB per cfg-A frame) are all-gathered in place over NVLink in one NCCL group; every rank then spreads all
// frames and scores its template shard.  The expensive half of the frame side (pyrDown + quantisers) is thereby divided
// by the number of GPUs instead of replicated — it was what capped round 1's template-sharded scaling at ~29 %.
// Falls back to the replicated frame side when count is not a multiple of the world size.
int lmb200_match_resident_sharded(lmb200_handle h, int first_slot, int count, float threshold,
                                  const char* const* class_ids, int n_class_ids) {
  if (!h || !h->device_ready || h->rows == 0) return set_error(h, LMB200_E_INVALID, "no frames uploaded");
  if (first_slot < 0 || count <= 0 || first_slot + count > h->slots) return set_error(h, LMB200_E_INVALID, "bad slot range");
  if (!h->nccl_comm || h->comm_world != h->shard_world || h->comm_rank != h->shard_rank)
    return set_error(h, LMB200_E_COMM, "lmb200_match_resident_sharded needs lmb200_comm_init and lmb200_set_template_shard with the same rank/world");
  const int world = h->comm_world;
  if (world == 1 || count % world != 0) return lmb200_match_resident(h, first_slot, count, threshold, class_ids, n_class_ids);
  cudaSetDevice(h->device);
  drain_tickets(h);
  int rc = ensure_plan(h, h->rows, h->cols);
  if (rc) return rc;
  rc = ensure_selection(h, class_ids, n_class_ids);
  if (rc) return rc;
  h->masks_in_use = false;
  const int M = h->cfg.num_modalities, L = h->cfg.pyramid_levels;
  const int n = count / world, own = first_slot + h->comm_rank * n;
  const bool trace = std::getenv("LMB200_TRACE") != nullptr;
  const double t_begin = trace ? now_ms() : 0.0;
  double tp[6] = {0, 0, 0, 0, 0, 0};
  int tpi = 0;
  auto lap = [&]() { if (trace && tpi < 6) tp[tpi++] = now_ms(); };
  cudaStream_t st = h->lanes[0].stream;
  // Lanes.  Quantisers and the all-gather of the quantized maps (shard_overlap = 1, the default; 2: spread + linearize
  // too — measured slower at 8 GPUs: both sides are L2-heavy) run on lane 5, so that the frame side of step k+1 overlaps
  // the template side of step k on the compute lane; the compute lane picks the step up through ev_q.
  // The map all-gather uses the main communicator, the match gather the second one: NCCL serialises the collectives of one
  // communicator in issue order, which would chain the maps of step k+1 behind the matches of step k.
  cudaStream_t sq = h->shard_overlap ? h->lanes[5].stream : st;
  const bool spread_on_sq = h->shard_overlap >= 2;
  ShardJob* job = nullptr;
  for (auto& g : h->jobs) if (g.first == first_slot && g.count == count) job = &g;
  if (!job) job = &h->jobs[h->job_next++ & 3];
  shard_job_retire(h, job);          // an unfetched step on this entry: let the epilogue thread finish with it first
  if (!job->ev) CU(cudaEventCreateWithFlags(&job->ev, cudaEventDisableTiming));
  if (!job->ev_q) CU(cudaEventCreateWithFlags(&job->ev_q, cudaEventDisableTiming));
  if (h->upload_pending) {
    CU(cudaStreamWaitEvent(sq, h->upload_ev, 0));
    h->upload_pending = false;
  }
  if (sq != st)   // the maps of these slots may still be read by an earlier step on the compute lane
    for (auto& mk : h->resident_marks)
      if (mk.ev && mk.first < first_slot + count && first_slot < mk.first + mk.count) CU(cudaStreamWaitEvent(sq, mk.ev, 0));
  lap();
  rc = run_frame_side(h, own, n, sq, FS_QUANTIZE);
  if (rc) return rc;
  lap();
  {
    ProfScope ps_comm(h, LMB200_K_COMM, sq);
    rc = comm_group_begin(h);
    if (rc) return rc;
    for (int l = 0; l < L; ++l)
      for (int m = 0; m < M; ++m) {
        // maps the quantise phase writes: ColorGradient at every level; DepthNormal at level 0 (+ the decimation chain when materialised)
        const bool written = h->cfg.modalities[m].type == LMB200_COLOR_GRADIENT || l == 0 || h->dn_materialize;
        if (!written) continue;
        LevelBuffers& lb = h->levels[l];
        u8* base = lb.q[m].as<u8>() + (size_t)first_slot * lb.q_stride;
        rc = comm_allgather(h, base + (size_t)h->comm_rank * n * lb.q_stride, base, (size_t)n * lb.q_stride, sq);
        if (rc) { comm_group_end(h); return rc; }
      }
    rc = comm_group_end(h);
    if (rc) return rc;
  }
  if (sq != st && !spread_on_sq) {
    CU(cudaEventRecord(job->ev_q, sq));
    CU(cudaStreamWaitEvent(st, job->ev_q, 0));
  }
  lap();
  rc = run_frame_side(h, first_slot, count, spread_on_sq ? sq : st, FS_SPREAD);
  if (rc) return rc;
  if (sq != st && spread_on_sq) {
    CU(cudaEventRecord(job->ev_q, sq));
    CU(cudaStreamWaitEvent(st, job->ev_q, 0));
  }
  lap();
  rc = run_matching(h, first_slot, count, threshold, st);
  if (rc) return rc;
  lap();
  ResidentMark& mk = h->resident_marks[h->resident_next++ & 3];
  if (!mk.ev) CU(cudaEventCreateWithFlags(&mk.ev, cudaEventDisableTiming));
  else CU(evict_mark(h, mk));
  mk.first = first_slot; mk.count = count;
  CU(cudaEventRecord(mk.ev, st));
  // The match gather rides on the compute lane right behind the kernels — pack, ncclAllGather and the copy to pinned host
  // memory need no host decision — and the epilogue thread takes over from there (ShardJob in detector.h).  Overflowing
  // stores / lists longer than the buffer show in the gathered headers and send the fetch down the synchronous path.
  {
    if (h->gather_cap <= 0) h->gather_cap = 256;   // average records per frame the buffer holds; doubles when a step needs more
    const int rec_cap = h->gather_cap * count;
    const size_t bytes = ((size_t)2 * count + rec_cap) * sizeof(Cand);
    job->first = first_slot; job->count = count;
    ALLOC(job->send, bytes);
    ALLOC(job->recv, bytes * world);
    const bool dev_epi = h->shard_device_epilogue && world <= 64;
    if (!dev_epi && job->host_bytes < bytes * world) {   // host epilogue thread: the gathered buffers are copied to pinned memory
      if (job->host) { cudaFreeHost(job->host); job->host = nullptr; }
      CU(cudaHostAlloc((void**)&job->host, bytes * world, cudaHostAllocDefault));
      job->host_bytes = bytes * world;
    }
    if (dev_epi) {
      // finished lists come from the device (kernels_epilogue.cu): tables, scratch and the pinned, device-mapped result area
      if (h->epi_tables_epoch != h->plan_epoch) {
        const size_t nb = (size_t)std::max(1, h->ntpl) * sizeof(int);
        ALLOC(h->d_gclass, nb); ALLOC(h->d_gtid, nb); ALLOC(h->d_posg, nb);
        if (h->ntpl > 0) {
          CU(cudaMemcpyAsync(h->d_gclass.p, h->g_class.data(), (size_t)h->ntpl * sizeof(int), cudaMemcpyHostToDevice, st));
          CU(cudaMemcpyAsync(h->d_gtid.p, h->g_tid.data(), (size_t)h->ntpl * sizeof(int), cudaMemcpyHostToDevice, st));
          if (h->shard_interleaved) CU(cudaMemcpyAsync(h->d_posg.p, h->pos_of_g.data(), (size_t)h->ntpl * sizeof(int), cudaMemcpyHostToDevice, st));
          CU(cudaStreamSynchronize(st));   // the sources are pageable vectors: finish the copies before anything can change them
        }
        h->epi_tables_epoch = h->plan_epoch;
      }
      const size_t fin_cap = (size_t)rec_cap * world;
      ALLOC(job->fin_dev, fin_cap * sizeof(EpiMatch));
      if (job->fin_cap < fin_cap) {
        if (job->fin_host) { cudaFreeHost(job->fin_host); job->fin_host = nullptr; }
        CU(cudaHostAlloc((void**)&job->fin_host, fin_cap * sizeof(EpiMatch), cudaHostAllocMapped));
        job->fin_cap = fin_cap;
      }
      if (job->hdr_frames < count) {
        if (job->hdr_host) { cudaFreeHost(job->hdr_host); job->hdr_host = nullptr; }
        CU(cudaHostAlloc((void**)&job->hdr_host, (size_t)2 * count * sizeof(int4), cudaHostAllocMapped));
        job->hdr_frames = count;
      }
      if (!job->ev_g) CU(cudaEventCreateWithFlags(&job->ev_g, cudaEventDisableTiming));
      if (job->generation >= 0) CU(cudaStreamWaitEvent(st, job->ev, 0));   // the previous epilogue on this entry still reads recv
    }
    launch_gather_pack(h->d_ctr.as<SlotCtr>() + first_slot, h->d_out.as<Cand>() + (size_t)first_slot * h->out_cap, h->out_cap,
                       job->send.as<Cand>(), rec_cap, count, st);
    rc = comm_allgather(h, job->send.p, job->recv.p, bytes, st, true);
    if (rc) return rc;
    job->cap = rec_cap; job->generation = h->buffer_generation;
    if (dev_epi) {
      // std::sort + std::unique of every frame on the high-priority lane: the serial sorts (one thread per frame) must not
      // hold the next step's kernels back on the compute lane
      cudaStream_t se = h->lanes[1].stream;
      CU(cudaEventRecord(job->ev_g, st));
      CU(cudaStreamWaitEvent(se, job->ev_g, 0));
      EpilogueArgs ea;
      ea.gathered = job->recv.as<Cand>(); ea.world = world; ea.rank = h->comm_rank; ea.frames = count; ea.gcap = rec_cap;
      ea.pos_of_g = h->shard_interleaved ? h->d_posg.as<int>() : nullptr;
      ea.g_class = h->d_gclass.as<int>(); ea.g_tid = h->d_gtid.as<int>();
      ea.out_dev = job->fin_dev.as<EpiMatch>(); ea.out_cap = (int)std::min<size_t>(job->fin_cap, 0x7fffffff);
      void* dp = nullptr;
      CU(cudaHostGetDevicePointer(&dp, job->fin_host, 0)); ea.out_host = (EpiMatch*)dp;
      CU(cudaHostGetDevicePointer(&dp, job->hdr_host, 0)); ea.hdr = (int4*)dp;
      {
        ProfScope ps(h, LMB200_K_EPILOGUE, se);
        launch_shard_epilogue(ea, se);
      }
      CU(cudaEventRecord(job->ev, se));
      job->state = 4;
    } else {
      CU(cudaMemcpyAsync(job->host, job->recv.p, bytes * world, cudaMemcpyDeviceToHost, st));
      CU(cudaEventRecord(job->ev, st));
      shard_job_submit(h, job);
    }
  }
  if (trace && h->comm_rank == 0) {
    const double t_end = now_ms();
    std::fprintf(stderr, "[lmb200 trace] sharded submit: enqueue %.3f ms", t_end - t_begin);
    if (t_end - t_begin > 1.0)
      std::fprintf(stderr, " (prepare %.3f, quantise %.3f, map all-gather %.3f, spread %.3f, matching %.3f, gather + epilogue %.3f)",
                   tp[0] - t_begin, tp[1] - tp[0], tp[2] - tp[1], tp[3] - tp[2], tp[4] - tp[3], t_end - tp[4]);
    std::fprintf(stderr, "\n");
  }
  return LMB200_OK;
}

// Stream on which the results of slots [first, first+count) can be read: the copy stream, made to wait for the
// matching lmb200_match_resident call (falls back to the compute stream when no mark covers the range).
static cudaStream_t resident_fetch_stream(lmb200_detector* h, int first, int count) {
  for (int i = 1; i <= 4; ++i) {
    ResidentMark& mk = h->resident_marks[(h->resident_next - i) & 3];
    if (mk.ev && mk.first <= first && first + count <= mk.first + mk.count) {
      if (cudaStreamWaitEvent(h->lanes[1].stream, mk.ev, 0) == cudaSuccess) return h->lanes[1].stream;
      break;
    }
  }
  return h->lanes[0].stream;
}

int lmb200_fetch_resident(lmb200_handle h, int first_slot, int count, lmb200_match_rec* out, size_t cap, size_t* offsets) {
  if (!h || !h->device_ready) return set_error(h, LMB200_E_INVALID, "nothing to fetch");
  if (first_slot < 0 || count <= 0 || first_slot + count > h->slots) return set_error(h, LMB200_E_INVALID, "bad slot range");
  cudaSetDevice(h->device);
  std::vector<std::vector<Cand>> raw;
  int rc = fetch_grow(h, first_slot, count, resident_fetch_stream(h, first_slot, count), raw);
  if (rc) return rc;
  size_t base = 0;
  int status = LMB200_OK;
  std::vector<std::vector<Match>> ms;
  finalize_frames(h, raw, ms);
  for (int i = 0; i < count; ++i) {
    h->prof.candidates += (long long)raw[i].size();
    h->prof.matches += (long long)ms[i].size();
    size_t n = 0;
    if (offsets) offsets[i] = base;
    if (emit(h, ms[i], out, cap, base, &n) != LMB200_OK) status = LMB200_E_TRUNCATED;
    base += n;
  }
  if (offsets) offsets[count] = base;
  return status;
}

int lmb200_synchronize(lmb200_handle h) {
  if (!h || !h->device_ready) return LMB200_OK;
  cudaSetDevice(h->device);
  for (int i = 0; i < LMB200_LANES; ++i) CU(cudaStreamSynchronize(h->lanes[i].stream));
  if (h->profiling) collect_profile(h);
  return LMB200_OK;
}

void* lmb200_stream(lmb200_handle h) { return h && h->device_ready ? (void*)h->lanes[0].stream : nullptr; }

int lmb200_timer_record(lmb200_handle h, int which) {
  if (!h || which < 0 || which > 1) return LMB200_E_INVALID;
  int rc = ensure_device(h);
  if (rc) return rc;
  if (!h->timer[which]) CU(cudaEventCreate(&h->timer[which]));
  CU(cudaEventRecord(h->timer[which], h->lanes[0].stream));
  return LMB200_OK;
}
int lmb200_timer_elapsed_ms(lmb200_handle h, float* ms) {
  if (!h || !ms || !h->timer[0] || !h->timer[1]) return LMB200_E_INVALID;
  CU(cudaEventSynchronize(h->timer[1]));
  CU(cudaEventElapsedTime(ms, h->timer[0], h->timer[1]));
  return LMB200_OK;
}

static int upload_masks(lmb200_detector* h, const lmb200_image* masks, cudaStream_t st) {
  const int M = h->cfg.num_modalities, L = h->cfg.pyramid_levels;
  bool any = false;
  for (int m = 0; m < M; ++m) any |= masks[m].data != nullptr;
  h->masks_in_use = any;
  if (!any) return LMB200_OK;
  for (int m = 0; m < M; ++m) {
    for (int l = 0; l < L; ++l) {
      LevelBuffers& lb = h->levels[l];
      ALLOC(lb.mask[m], lb.q_stride * h->slots);
    }
    LevelBuffers& l0 = h->levels[0];
    if (!masks[m].data) {  // no mask for this modality = all ones
      for (int l = 0; l < L; ++l) CU(cudaMemsetAsync(h->levels[l].mask[m].p, 0xFF, h->levels[l].q_stride, st));
      continue;
    }
    if (masks[m].type != LMB200_8UC1 || masks[m].rows != h->rows || masks[m].cols != h->cols)
      return set_error(h, LMB200_E_SOURCES, "mask must be 8UC1 of the source size");
    size_t step = masks[m].step ? masks[m].step : (size_t)h->cols;
    CU(cudaMemcpy2DAsync(l0.mask[m].p, h->cols, masks[m].data, step, h->cols, h->rows, cudaMemcpyHostToDevice, st));
    for (int l = 1; l < L; ++l) {
      LevelBuffers& pb = h->levels[l - 1];
      LevelBuffers& lb = h->levels[l];
      launch_resize_nn(pb.mask[m].as<u8>(), pb.q_stride, pb.g.rows, pb.g.cols, lb.mask[m].as<u8>(), lb.q_stride, lb.g.rows,
                       lb.g.cols, 1, st);
    }
  }
  return LMB200_OK;
}

int lmb200_match(lmb200_handle h, const lmb200_image* sources, int n_sources, float threshold,
                 const char* const* class_ids, int n_class_ids, lmb200_match_rec* out, size_t cap, size_t* n_out,
                 lmb200_image* quantized_out, const lmb200_image* masks) {
  if (n_out) *n_out = 0;
  if (h) drain_tickets(h);
  int rc = prepare(h, sources, 1, n_sources, class_ids, n_class_ids);
  if (rc) return rc;
  cudaStream_t st = h->lanes[0].stream;
  {
    ProfScope ps(h, LMB200_K_UPLOAD, st);
    rc = upload_one(h, sources, 0, st);
    if (rc) return rc;
  }
  if (masks) {
    rc = upload_masks(h, masks, st);
    if (rc) return rc;
  } else {
    h->masks_in_use = false;
  }
  // Single-frame latency path (BASELINE configs[1] literally): the ~15 launches + memsets + result copies of one frame are
  // captured once into a CUDA graph and replayed with one launch; any change of plan, template set, selection, stores or
  // threshold re-captures (plan_epoch).  Masks, profiling and LMB200_NO_GRAPH=1 take the plain stream path.
  bool graphed = false;
  if (!masks && !h->profiling && h->use_graph) {
    u32 thr_bits; std::memcpy(&thr_bits, &threshold, 4);
    if (!h->match_graph || h->graph_epoch != h->plan_epoch || h->graph_thr_bits != thr_bits || h->graph_early_exit != h->early_exit) {
      if (h->match_graph) { cudaGraphExecDestroy(h->match_graph); h->match_graph = nullptr; }
      lmb200_profile before = h->prof;
      cudaGraph_t g = nullptr;
      CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      // The modalities' chains (ColorGradient: quantise / pyrDown / spread per level; DepthNormal: quantise + spreads)
      // are independent until the template side: one graph branch each, forked and joined through events, so a single
      // frame's small grids run side by side instead of one after the other.
      int r1 = LMB200_OK;
      const int Mm = h->cfg.num_modalities;
      if (Mm > 1 && Mm <= 4) {
        for (auto& e : h->fork_ev) if (!e) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        cudaMemsetAsync(h->d_resp_sum.as<u32>(), 0, MAX_MOD * sizeof(u32), st);
        cudaEventRecord(h->fork_ev[0], st);
        for (int m = 1; m < Mm && !r1; ++m) {
          cudaStream_t sb = h->lanes[1 + m].stream;     // lanes 2..4
          cudaStreamWaitEvent(sb, h->fork_ev[0], 0);
          r1 = run_frame_side(h, 0, 1, sb, FS_ALL, m);
          cudaEventRecord(h->fork_ev[m], sb);
        }
        if (!r1) r1 = run_frame_side(h, 0, 1, st, FS_ALL, 0);
        for (int m = 1; m < Mm; ++m) cudaStreamWaitEvent(st, h->fork_ev[m], 0);
      } else {
        r1 = run_frame_side(h, 0, 1, st);
      }
      int r2 = r1 ? r1 : run_matching(h, 0, 1, threshold, st);
      int r3 = r2 ? r2 : enqueue_result_copies(h, 0, 1, st);
      cudaError_t ce = cudaStreamEndCapture(st, &g);
      if (r3 || ce != cudaSuccess) { if (g) cudaGraphDestroy(g); cudaGetLastError(); return r3 ? r3 : cuda_fail(h, ce, "cudaStreamEndCapture"); }
      ce = cudaGraphInstantiate(&h->match_graph, g, 0);
      cudaGraphDestroy(g);
      if (ce != cudaSuccess) { h->match_graph = nullptr; return cuda_fail(h, ce, "cudaGraphInstantiate"); }
      for (int k = 0; k < LMB200_K_COUNT; ++k) h->graph_launches[k] = h->prof.launches[k] - before.launches[k];
      h->prof = before;   // the capture itself launched nothing
      h->graph_epoch = h->plan_epoch; h->graph_thr_bits = thr_bits; h->graph_early_exit = h->early_exit;
    }
    CU(cudaGraphLaunch(h->match_graph, st));
    for (int k = 0; k < LMB200_K_COUNT; ++k) h->prof.launches[k] += h->graph_launches[k];
    h->prof.frames += 1; h->prof.bytes_coarse += h->sel_bytes_coarse;
    h->slot_threshold[0] = threshold; h->slot_gen[0] = h->buffer_generation;
    graphed = true;
  } else {
    rc = run_frame_side(h, 0, 1, st);
    if (rc) return rc;
    rc = run_matching(h, 0, 1, threshold, st);
    if (rc) return rc;
  }
  std::vector<std::vector<Cand>> raw;
  rc = fetch_grow(h, 0, 1, st, raw, graphed);
  h->masks_in_use = false;
  if (rc) return rc;
  if (quantized_out) {
    const int M = h->cfg.num_modalities, L = h->cfg.pyramid_levels;
    rc = materialize_dn_levels(h, 0, 1, st);
    if (rc) return rc;
    CU(cudaStreamSynchronize(st));
    for (int l = 0; l < L; ++l)
      for (int m = 0; m < M; ++m) {
        lmb200_image& qi = quantized_out[l * M + m];
        LevelBuffers& lb = h->levels[l];
        if (!qi.data) continue;
        if (qi.rows != lb.g.rows || qi.cols != lb.g.cols) return set_error(h, LMB200_E_INVALID, "quantized_out image has the wrong size");
        size_t step = qi.step ? qi.step : (size_t)lb.g.cols;
        // quantize() output is the map after masking
        CU(cudaMemcpy2D((void*)qi.data, step, lb.q[m].p, lb.g.cols, lb.g.cols, lb.g.rows, cudaMemcpyDeviceToHost));
        if (masks && masks[m].data) {
          std::vector<u8> mk((size_t)lb.g.rows * lb.g.cols);
          CU(cudaMemcpy(mk.data(), lb.mask[m].p, mk.size(), cudaMemcpyDeviceToHost));
          u8* q = (u8*)qi.data;
          for (int r = 0; r < lb.g.rows; ++r)
            for (int c = 0; c < lb.g.cols; ++c)
              if (!mk[(size_t)r * lb.g.cols + c]) q[(size_t)r * step + c] = 0;
        }
      }
  }
  std::vector<Match> m;
  to_matches(h, raw[0], m);
  h->prof.candidates += (long long)raw[0].size();
  finalize_matches(m);
  h->prof.matches += (long long)m.size();
  size_t n = 0;
  rc = emit(h, m, out, cap, 0, &n);
  if (n_out) *n_out = n;
  return rc;
}

// Streaming batch.  Stream C (copy) brings chunks of frames into one of G slot groups, the compute
// streams run the whole path on them (round-robin) and copy counters + list heads into per-frame pinned
// staging; copy and compute are chained by events only, so the host enqueues the ENTIRE batch without
// blocking (submit) and later finalises chunk after chunk (collect: sort/unique) while later chunks — or the
// next batch — are still copying/computing.  Slot groups are recycled through one rolling "done" event per
// group, so consecutive batches pipeline into each other.
//   C:  [wait group_done(g)] H2D(k) rec h2d(k)          X:  [wait h2d(k)] kernels(k) D2H(k) rec group_done(g), done(k)
static int batch_enqueue(lmb200_detector* h, BatchTicket& tk) {
  NvtxRange nvtx("lmb200:batch_enqueue(H2D chunks + kernels)");
  const int n_frames = tk.n_frames, n_sources = tk.n_sources;
  const lmb200_image* frames = tk.frames.data();
  cudaStream_t Xs[4] = {h->lanes[0].stream, h->lanes[2].stream, h->lanes[3].stream, h->lanes[4].stream}, Cs = h->lanes[1].stream;
  int nx = 3;  // compute streams used round-robin by consecutive chunks (measured: 1 -> 5.4 ms, 2 -> 4.33, 3 -> 4.25 per 96 frames)
  if (const char* e = std::getenv("LMB200_XSTREAMS")) nx = std::max(1, std::min(4, std::atoi(e)));
  int G = std::max(1, std::min(4, h->slots / 12));  // slot groups of >= 12 frames
  if (h->slots >= 2 && G < 2) G = 2;
  if (const char* e = std::getenv("LMB200_GROUPS")) G = std::max(1, std::min(h->slots, std::atoi(e)));
  const int gs = h->slots / G;
  // A stream of submitted batches keeps the GPU busy across batch boundaries, so it takes large chunks (kernel
  // efficiency: 96 frames/launch run 1.6x faster per frame than 12); the blocking call takes small ones (its first
  // copy and its last chunk are exposed).  Measured, 96 frames x 3 000 templates: streaming 32.1 k -> 33.7 k frames/s.
  int chunk = std::min(gs, tk.streaming ? 24 : 12);
  if (const char* e = std::getenv("LMB200_CHUNK")) chunk = std::max(1, std::min(gs, std::atoi(e)));
  if (G != h->b_groups) {  // (re)create the rolling per-group events
    CU(cudaDeviceSynchronize());
    for (auto e : h->group_done) cudaEventDestroy(e);
    h->group_done.assign(G, nullptr);
    for (auto& e : h->group_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->group_used.assign(G, 0);
    h->b_groups = G;
  }
  // chunk schedule.  Blocking call: ramp up (chunk/3, 2*chunk/3, chunk, ...) so compute starts after a short first copy.
  // Streaming (submit/collect): the previous batch is still computing when this one starts copying, so there is nothing to
  // hide — equal chunks, as large as the slot groups allow (kernels are more efficient on larger launches).  The tail is
  // balanced instead of leaving a runt chunk.
  tk.cf0.clear(); tk.ccnt.clear();
  {
    int f = 0;
    if (!tk.streaming)
      for (int k = 0; k < 2 && f < n_frames; ++k) {
        int cnt = std::min(std::max(1, (k + 1) * chunk / 3), n_frames - f);
        tk.cf0.push_back(f); tk.ccnt.push_back(cnt);
        f += cnt;
      }
    const int rest = n_frames - f, nc = (rest + chunk - 1) / chunk;
    for (int k = 0; k < nc; ++k) {
      int cnt = rest / nc + (k < rest % nc ? 1 : 0);
      tk.cf0.push_back(f); tk.ccnt.push_back(cnt);
      f += cnt;
    }
  }
  const int nchunks = (int)tk.cf0.size();
  // per-frame pinned staging for the results of this batch
  if (tk.cap_frames < n_frames || tk.head != h->h_head) {
    if (tk.b_ctr) cudaFreeHost(tk.b_ctr);
    if (tk.b_out) cudaFreeHost(tk.b_out);
    tk.b_ctr = nullptr; tk.b_out = nullptr;
    CU(cudaHostAlloc((void**)&tk.b_ctr, (size_t)n_frames * sizeof(SlotCtr), cudaHostAllocDefault));
    CU(cudaHostAlloc((void**)&tk.b_out, (size_t)n_frames * h->h_head * sizeof(Cand), cudaHostAllocDefault));
    tk.cap_frames = n_frames; tk.head = h->h_head;
  }
  while ((int)tk.events.size() < 2 * nchunks) {
    cudaEvent_t ev;
    CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    tk.events.push_back(ev);
  }
  // resident steps (lmb200_match_resident[_sharded]) still in flight read the same frame slots: the copy stream orders
  // itself behind their completion marks (free when they have long finished)
  for (auto& mk : h->resident_marks)
    if (mk.ev) CU(cudaStreamWaitEvent(Cs, mk.ev, 0));
  const bool trace = std::getenv("LMB200_TRACE") != nullptr;
  std::vector<cudaEvent_t> tev;  // trace: 4 timing events per chunk (copy start/end, compute start/end)
  if (trace) { tev.resize(4 * (size_t)nchunks); for (auto& e : tev) cudaEventCreate(&e); }
  for (int k = 0; k < nchunks; ++k) {
    const int f0 = tk.cf0[k], cnt = tk.ccnt[k];
    const long long seq = h->chunk_seq++;
    const int g = (int)(seq % G), slot0 = g * gs;
    cudaEvent_t ev_h2d = tk.events[2 * k], ev_done = tk.events[2 * k + 1];
    cudaStream_t X = Xs[seq % nx];  // adjacent chunks overlap on separate compute streams (fills the low-parallelism tails)
    if (h->group_used[g]) CU(cudaStreamWaitEvent(Cs, h->group_done[g], 0));  // slot group free again
    if (trace) cudaEventRecord(tev[4 * k], Cs);
    if (chunk_is_contiguous(h, frames + (size_t)f0 * n_sources, cnt, n_sources)) {
      CU(cudaMemcpyAsync((u8*)h->d_frames.p + (size_t)slot0 * h->frame_bytes, frames[(size_t)f0 * n_sources].data,
                         h->frame_bytes * cnt, cudaMemcpyHostToDevice, Cs));   // one DMA for the whole chunk
    } else {
      for (int i = 0; i < cnt; ++i) {
        int rc = upload_one(h, frames + (size_t)(f0 + i) * n_sources, slot0 + i, Cs);
        if (rc) return rc;
      }
    }
    h->prof.launches[LMB200_K_UPLOAD]++;
    CU(cudaEventRecord(ev_h2d, Cs));
    if (trace) cudaEventRecord(tev[4 * k + 1], Cs);
    CU(cudaStreamWaitEvent(X, ev_h2d, 0));
    if (trace) cudaEventRecord(tev[4 * k + 2], X);
    int rc = run_frame_side(h, slot0, cnt, X);
    if (rc) return rc;
    rc = run_matching(h, slot0, cnt, tk.threshold, X);
    if (rc) return rc;
    CU(cudaMemcpyAsync(tk.b_ctr + f0, h->d_ctr.as<SlotCtr>() + slot0, sizeof(SlotCtr) * cnt, cudaMemcpyDeviceToHost, X));
    CU(cudaMemcpy2DAsync(tk.b_out + (size_t)f0 * h->h_head, (size_t)h->h_head * sizeof(Cand),
                         h->d_out.as<Cand>() + (size_t)slot0 * h->out_cap, (size_t)h->out_cap * sizeof(Cand),
                         (size_t)h->h_head * sizeof(Cand), cnt, cudaMemcpyDeviceToHost, X));
    CU(cudaEventRecord(ev_done, X));
    CU(cudaEventRecord(h->group_done[g], X));
    h->group_used[g] = 1;
    if (trace) cudaEventRecord(tev[4 * k + 3], X);
  }
  if (trace) {
    cudaDeviceSynchronize();
    for (int k = 0; k < nchunks; ++k) {
      float t[4];
      for (int j = 0; j < 4; ++j) cudaEventElapsedTime(&t[j], tev[0], tev[4 * k + j]);
      std::fprintf(stderr, "[lmb200 trace] chunk %2d: h2d %.3f..%.3f  compute %.3f..%.3f ms\n", k, t[0], t[1], t[2], t[3]);
    }
    for (auto& e : tev) cudaEventDestroy(e);
  }
  tk.head_used = h->h_head;
  tk.generation = h->buffer_generation;
  return LMB200_OK;
}

// Finalise a submitted batch in frame order.  Returns +1 when a device-side store (or the staged list head) was
// too small, or when the stores were reallocated under the ticket: the caller grows and re-runs the batch.
static int batch_finalize(lmb200_detector* h, BatchTicket& tk, lmb200_match_rec* out, size_t cap, size_t* offsets, int* status) {
  const int nchunks = (int)tk.cf0.size();
  if (tk.generation != h->buffer_generation) return 1;
  size_t base = 0;
  std::vector<Cand> raw;
  std::vector<Match> m;
  *status = LMB200_OK;
  for (int k = 0; k < nchunks; ++k) {
    const int f0 = tk.cf0[k], cnt = tk.ccnt[k];
    CU(cudaEventSynchronize(tk.events[2 * k + 1]));
    for (int i = 0; i < cnt; ++i) {
      const int f = f0 + i, n = tk.b_ctr[f].out_count;
      if (tk.b_ctr[f].overflow || n > h->out_cap || n > tk.head_used) return 1;
      raw.assign(tk.b_out + (size_t)f * tk.head_used, tk.b_out + (size_t)f * tk.head_used + n);
      h->prof.bytes_local += (long long)tk.b_ctr[f].local_bytes;
      h->prof.chunks_coarse += (long long)tk.b_ctr[f].coarse_chunks;
      to_matches(h, raw, m);
      h->prof.candidates += (long long)raw.size();
      finalize_matches(m);
      h->prof.matches += (long long)m.size();
      size_t w = 0;
      if (offsets) offsets[f] = base;
      if (emit(h, m, out, cap, base, &w) != LMB200_OK) *status = LMB200_E_TRUNCATED;
      base += w;
    }
  }
  if (offsets) offsets[tk.n_frames] = base;
  return LMB200_OK;
}

int lmb200_match_batch_submit(lmb200_handle h, const lmb200_image* frames, int n_frames, int n_sources, float threshold,
                              const char* const* class_ids, int n_class_ids, int* ticket) {
  if (!h || !ticket) return LMB200_E_INVALID;
  int t = -1;
  for (int i = 0; i < 2; ++i)
    if (!h->tickets[i].active) { t = i; break; }
  if (t < 0) return set_error(h, LMB200_E_INVALID, "two batches are already in flight: collect one first");
  // a different class selection than the batch in flight would overwrite the shared selection list under it
  std::string key;
  for (int i = 0; i < n_class_ids; ++i) { key += class_ids[i]; key.push_back('\x1f'); }
  BatchTicket& other = h->tickets[t ^ 1];
  if (other.active && (other.sel_key != key) && h->device_ready) {
    cudaSetDevice(h->device);
    for (int i = 0; i < LMB200_LANES; ++i) cudaStreamSynchronize(h->lanes[i].stream);
  }
  int rc = prepare(h, frames, n_frames, n_sources, class_ids, n_class_ids);
  if (rc) return rc;
  h->masks_in_use = false;
  BatchTicket& tk = h->tickets[t];
  tk.frames.assign(frames, frames + (size_t)n_frames * n_sources);
  tk.n_frames = n_frames; tk.n_sources = n_sources; tk.threshold = threshold; tk.sel_key = key;
  tk.streaming = !h->blocking_submit;
  tk.class_ids.clear();
  for (int i = 0; i < n_class_ids; ++i) tk.class_ids.push_back(class_ids[i]);
  rc = batch_enqueue(h, tk);
  if (rc) return rc;
  tk.active = true;
  *ticket = t;
  return LMB200_OK;
}

int lmb200_match_batch_collect(lmb200_handle h, int ticket, lmb200_match_rec* out, size_t cap, size_t* offsets) {
  if (!h || ticket < 0 || ticket > 1 || !h->tickets[ticket].active) return set_error(h, LMB200_E_INVALID, "no such batch in flight");
  cudaSetDevice(h->device);
  BatchTicket& tk = h->tickets[ticket];
  for (;;) {
    int status = LMB200_OK;
    int rc = batch_finalize(h, tk, out, cap, offsets, &status);
    if (rc == LMB200_OK) {
      if (h->profiling) { for (int i = 0; i < LMB200_LANES; ++i) CU(cudaStreamSynchronize(h->lanes[i].stream)); collect_profile(h); }
      tk.active = false;
      return status;
    }
    if (rc < 0) { tk.active = false; return rc; }
    // grow (unless another collect already did) and rerun this batch; the caller's frames must still be alive
    CU(cudaDeviceSynchronize());
    if (tk.generation == h->buffer_generation) {
      bool store = false;
      for (int f = 0; f < tk.n_frames; ++f) store |= tk.b_ctr[f].overflow != 0 || tk.b_ctr[f].out_count > h->out_cap;
      if (store) { rc = grow_capacity(h); if (rc) { tk.active = false; return rc; } }
      else { h->h_head = std::min(h->out_cap, h->h_head * 4); h->buffer_generation++; h->plan_epoch++; }
    }
    std::vector<const char*> ids;
    for (auto& s : tk.class_ids) ids.push_back(s.c_str());
    rc = ensure_selection(h, ids.empty() ? nullptr : ids.data(), (int)ids.size());
    if (rc) { tk.active = false; return rc; }
    rc = batch_enqueue(h, tk);
    if (rc) { tk.active = false; return rc; }
  }
}

int lmb200_match_batch(lmb200_handle h, const lmb200_image* frames, int n_frames, int n_sources, float threshold,
                       const char* const* class_ids, int n_class_ids, lmb200_match_rec* out, size_t cap, size_t* offsets) {
  int ticket = -1;
  if (h) h->blocking_submit = true;
  int rc = lmb200_match_batch_submit(h, frames, n_frames, n_sources, threshold, class_ids, n_class_ids, &ticket);
  if (h) h->blocking_submit = false;
  if (rc) return rc;
  return lmb200_match_batch_collect(h, ticket, out, cap, offsets);
}

int lmb200_host_alloc(size_t bytes, void** out) {
  cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) { cudaGetLastError(); return LMB200_E_CUDA; }
  return LMB200_OK;
}
int lmb200_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? LMB200_OK : LMB200_E_CUDA; }

// ---------------------------------------------------------------- post-match colour check
int lmb200_postmatch_color(lmb200_handle h, int slot, const uint8_t* lower_hsv, const uint8_t* upper_hsv,
                           const lmb200_match_rec* matches, size_t n, int* inside, int* total) {
  if (!h || !lower_hsv || !upper_hsv || (n && (!matches || !inside || !total))) return set_error(h, LMB200_E_INVALID, "bad arguments");
  if (!h->device_ready || h->rows == 0 || slot < 0 || slot >= h->slots) return set_error(h, LMB200_E_INVALID, "no frame resident in that slot (call lmb200_match / lmb200_upload_frames first)");
  cudaSetDevice(h->device);
  int rc = ensure_plan(h, h->rows, h->cols);
  if (rc) return rc;
  const int M = h->cfg.num_modalities;
  int mc = -1;
  for (int m = 0; m < M && mc < 0; ++m) if (h->cfg.modalities[m].type == LMB200_COLOR_GRADIENT) mc = m;
  if (mc < 0) return set_error(h, LMB200_E_INVALID, "the colour check needs a ColorGradient modality (its BGR source image)");
  if (n == 0) return LMB200_OK;
  cudaStream_t st = h->lanes[0].stream;
  const int wpr = (h->cols + 31) / 32;
  ALLOC(h->d_hue_bits, (size_t)h->rows * wpr * sizeof(u32));
  LevelBuffers& l0 = h->levels[0];
  launch_hsv_inrange_bits(l0.bgr[mc].as<u8>() + (size_t)slot * l0.bgr_stride, h->rows, h->cols, lower_hsv, upper_hsv, h->d_hue_bits.as<u32>(), st);
  // (class_index, template_id) -> global template index
  std::vector<int> first(h->class_list.size() + 1, 0);
  for (int i = 0; i < h->ntpl; ++i) first[h->g_class[i] + 1] = i + 1;
  for (size_t c = 1; c < first.size(); ++c) first[c] = std::max(first[c], first[c - 1]);
  std::vector<int2> xy(n);
  std::vector<int> gi(n);
  for (size_t i = 0; i < n; ++i) {
    xy[i] = make_int2(matches[i].x, matches[i].y);
    const int c = matches[i].class_index, t = matches[i].template_id;
    gi[i] = (c >= 0 && c < (int)h->class_list.size() && t >= 0 && first[c] + t < first[c + 1]) ? first[c] + t : -1;
  }
  ScopedDevBuf d_xy, d_g, d_res;
  ALLOC(d_xy, n * sizeof(int2)); ALLOC(d_g, n * sizeof(int)); ALLOC(d_res, n * sizeof(int2));
  CU(cudaMemcpyAsync(d_xy.p, xy.data(), n * sizeof(int2), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_g.p, gi.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
  launch_template_mask_count(d_xy.as<int2>(), (int)n, d_g.as<int>(), h->d_hdr[0].as<TplHdr>(), h->d_feat[0].as<u32>(), M, h->rows, h->cols,
                             h->d_hue_bits.as<u32>(), d_res.as<int2>(), st);
  CU(cudaGetLastError());
  std::vector<int2> res(n);
  CU(cudaMemcpyAsync(res.data(), d_res.p, n * sizeof(int2), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (size_t i = 0; i < n; ++i) { inside[i] = res[i].x; total[i] = res[i].y; }
  return LMB200_OK;
}

// HighLevelLineMOD::groupSimilarMatches + discardSmallMatchGroups (src/HighLevelLinemod.cpp:206-253), host code: the greedy
// grouping depends on the order of `matches` (the first member anchors its group), so it is restated sequentially.
int lmb200_group_matches(const lmb200_match_rec* matches, size_t n, float radius_threshold, float discard_group_ratio,
                         int* group_of_match, int* n_groups) {
  if ((n && (!matches || !group_of_match)) || !n_groups) return LMB200_E_INVALID;
  struct G { int x, y; std::vector<size_t> idx; };
  std::vector<G> groups;
  for (size_t i = 0; i < n; ++i) {
    bool found = false;
    for (size_t q = 0; q < groups.size() && !found; ++q) {
      const double dx = (double)matches[i].x - groups[q].x, dy = (double)matches[i].y - groups[q].y;
      if (std::sqrt(dx * dx + dy * dy) < radius_threshold) { groups[q].idx.push_back(i); found = true; }  // cv::norm(Point) < float
    }
    if (!found) { G g; g.x = matches[i].x; g.y = matches[i].y; g.idx.push_back(i); groups.push_back(g); }
  }
  size_t biggest = 0;
  for (auto& g : groups) biggest = std::max(biggest, g.idx.size());
  for (size_t i = 0; i < n; ++i) group_of_match[i] = -1;
  int kept = 0;
  for (auto& g : groups) {
    const float ratio = (float)(g.idx.size() * 100 / biggest);   // integer division first, like the reference
    if (ratio > discard_group_ratio) { for (size_t i : g.idx) group_of_match[i] = kept; ++kept; }
  }
  *n_groups = kept;
  return LMB200_OK;
}

int lmb200_postmatch_median_depth(const uint16_t* depth, int rows, int cols, size_t step_bytes, const int* bb4, int median_position,
                                  uint16_t* median) {
  if (!depth || !bb4 || !median || rows <= 0 || cols <= 0 || median_position <= 0) return LMB200_E_INVALID;
  const int x = bb4[0], y = bb4[1], w = bb4[2], hgt = bb4[3];
  if (x < 0 || y < 0 || w <= 0 || hgt <= 0 || x + w > cols || y + hgt > rows) return LMB200_E_INVALID;   // cv::Mat ROI assertion
  const size_t step = step_bytes ? step_bytes : (size_t)cols * sizeof(uint16_t);
  std::vector<uint16_t> v;
  v.reserve((size_t)w * hgt);
  for (int r = 0; r < hgt; ++r) {
    const uint16_t* row = reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(depth) + (size_t)(y + r) * step) + x;
    // threshold(in, 1, 65535, THRESH_BINARY) keeps values > 1; 65535 - that, added with saturation: 0 and 1 become 65535
    for (int c = 0; c < w; ++c) v.push_back(row[c] > 1 ? row[c] : (uint16_t)65535);
  }
  std::nth_element(v.begin(), v.begin() + v.size() / 4, v.end());
  *median = v[v.size() / (size_t)median_position];
  return LMB200_OK;
}

int lmb200_set_option(lmb200_handle h, const char* name, int value) {
  if (!h || !name) return LMB200_E_INVALID;
  if (std::strcmp(name, "early_exit") == 0) { h->early_exit = value != 0; return LMB200_OK; }
  if (std::strcmp(name, "upload_async") == 0) { h->upload_async = value != 0; return LMB200_OK; }
  if (std::strcmp(name, "cuda_graph") == 0) { h->use_graph = value != 0; return LMB200_OK; }
  if (std::strcmp(name, "shard_overlap") == 0) { h->shard_overlap = value < 0 ? 0 : value > 2 ? 2 : value; return LMB200_OK; }
  if (std::strcmp(name, "resident_overlap") == 0) { h->resident_overlap = value != 0; return LMB200_OK; }
  if (std::strcmp(name, "host_threads") == 0) { h->host_threads = value < 1 ? 1 : value > 64 ? 64 : value; return LMB200_OK; }
  if (std::strcmp(name, "shard_device_epilogue") == 0) { h->shard_device_epilogue = value != 0; return LMB200_OK; }
  return set_error(h, LMB200_E_INVALID, std::string("unknown option ") + name);
}

int lmb200_set_profiling(lmb200_handle h, int enabled) {
  if (!h) return LMB200_E_INVALID;
  h->profiling = enabled != 0;
  return LMB200_OK;
}
int lmb200_get_profile(lmb200_handle h, lmb200_profile* out, int reset) {
  if (!h || !out) return LMB200_E_INVALID;
  if (h->device_ready && h->profiling) {
    cudaSetDevice(h->device);
    for (int i = 0; i < LMB200_LANES; ++i) cudaStreamSynchronize(h->lanes[i].stream);
    collect_profile(h);
  }
  *out = h->prof;
  if (reset) std::memset(&h->prof, 0, sizeof(h->prof));
  return LMB200_OK;
}

int lmb200_debug_fetch(lmb200_handle h, int kind, int slot, int index, void* dst, size_t* n_bytes) {
  if (!h || !h->device_ready || !n_bytes || slot < 0 || slot >= h->slots) return set_error(h, LMB200_E_INVALID, "bad debug_fetch arguments");
  cudaSetDevice(h->device);
  const int M = h->cfg.num_modalities, L = h->cfg.pyramid_levels;
  cudaStream_t st = h->lanes[0].stream;
  CU(cudaDeviceSynchronize());
  auto give = [&](const void* dev, size_t n) -> int {
    size_t capb = *n_bytes;
    *n_bytes = n;
    if (!dst || capb < n) return LMB200_E_TRUNCATED;
    CU(cudaMemcpy(dst, dev, n, cudaMemcpyDeviceToHost));
    return LMB200_OK;
  };
  if (kind == LMB200_DBG_QUANTIZED || kind == LMB200_DBG_LINMEM) {
    if (index < 0 || index >= L * M) return set_error(h, LMB200_E_INVALID, "bad map index");
    LevelBuffers& lb = h->levels[index / M];
    int m = index % M;
    if (kind == LMB200_DBG_QUANTIZED) {
      int rc = materialize_dn_levels(h, slot, 1, st);
      if (rc) return rc;
      CU(cudaStreamSynchronize(st));
      return give(lb.q[m].as<u8>() + (size_t)slot * lb.q_stride, (size_t)lb.g.rows * lb.g.cols);
    }
    if (!lb.g.strips && lb.fast_spread) {  // coarsest level, fast path: only the nibble-packed memory exists; unpack it
      const size_t flat = (size_t)8 * lb.g.rows * lb.g.cols, capb = *n_bytes;
      *n_bytes = flat;
      if (!dst || capb < flat) return LMB200_E_TRUNCATED;
      std::vector<u8> tmp(flat / 2);
      CU(cudaMemcpy(tmp.data(), lb.lmn[m].as<u8>() + (size_t)slot * lb.lmn_stride, tmp.size(), cudaMemcpyDeviceToHost));
      u8* o = (u8*)dst;
      for (size_t i = 0; i < tmp.size(); ++i) { o[2 * i] = tmp[i] & 15; o[2 * i + 1] = tmp[i] >> 4; }
      return LMB200_OK;
    }
    if (!lb.g.strips) return give(lb.lm[m].as<u8>() + (size_t)slot * lb.lm_stride, (size_t)8 * lb.g.rows * lb.g.cols);
    // strip layout -> upstream's flat layout
    const size_t flat = (size_t)8 * lb.g.rows * lb.g.cols, capb = *n_bytes;
    *n_bytes = flat;
    if (!dst || capb < flat) return LMB200_E_TRUNCATED;
    std::vector<u8> tmp((size_t)8 * lb.g.per_label);
    CU(cudaMemcpy(tmp.data(), lb.lm[m].as<u8>() + (size_t)slot * lb.lm_stride, tmp.size(), cudaMemcpyDeviceToHost));
    const int W = lb.g.W, H = lb.g.H, TT = lb.g.T * lb.g.T;
    u8* o = (u8*)dst;
    for (int lp = 0; lp < 8 * TT; ++lp)  // (label, phase) planes
      for (int gy = 0; gy < H; ++gy)
        for (int gx = 0; gx < W; ++gx)
          o[((size_t)lp * H + gy) * W + gx] = tmp[(size_t)lp * lb.g.plane + (size_t)(gx >> 4) * H * 16 + (size_t)gy * 16 + (gx & 15)];
    return LMB200_OK;
  }
  if (kind == LMB200_DBG_COARSE || kind == LMB200_DBG_UNSORTED) {
    int rc = run_matching(h, slot, 1, h->slot_threshold[slot], st, kind == LMB200_DBG_COARSE);
    if (rc) return rc;
    std::vector<std::vector<Cand>> raw;
    rc = fetch_raw(h, slot, 1, st, raw);
    if (rc) return rc > 0 ? set_error(h, LMB200_E_TRUNCATED, "candidate store overflow in debug fetch") : rc;
    size_t n = raw[0].size() * sizeof(lmb200_match_rec), capb = *n_bytes;
    *n_bytes = n;
    if (!dst || capb < n) return LMB200_E_TRUNCATED;
    lmb200_match_rec* o = (lmb200_match_rec*)dst;
    for (size_t i = 0; i < raw[0].size(); ++i) {
      int g = raw[0][i].tsel;
      o[i].x = raw[0][i].x; o[i].y = raw[0][i].y; o[i].similarity = raw[0][i].sim;
      o[i].class_index = h->g_class[g]; o[i].template_id = h->g_tid[g];
    }
    return LMB200_OK;
  }
  if (kind == LMB200_DBG_MAGNITUDE) {
    if (index < 0 || index >= M || h->cfg.modalities[index].type != LMB200_COLOR_GRADIENT) return set_error(h, LMB200_E_INVALID, "not a ColorGradient modality");
    LevelBuffers& lb = h->levels[0];
    size_t n = (size_t)lb.g.rows * lb.g.cols;
    ALLOC(h->d_mag, n * sizeof(float));
    ScopedDevBuf tmpq;
    ALLOC(tmpq, n);
    const lmb200_modality& mod = h->cfg.modalities[index];
    launch_cg_quantize(lb.bgr[index].as<u8>() + (size_t)slot * lb.bgr_stride, lb.bgr_stride, tmpq.as<u8>(), n,
                       h->d_mag.as<float>(), n, lb.g.rows, lb.g.cols, mod.weak_threshold * mod.weak_threshold, 1, st);
    CU(cudaStreamSynchronize(st));
    int rc = give(h->d_mag.p, n * sizeof(float));
    tmpq.release();
    return rc;
  }
  if (kind == LMB200_DBG_DN_INDICES) {
    if (index < 0 || index >= M || h->cfg.modalities[index].type != LMB200_DEPTH_NORMAL) return set_error(h, LMB200_E_INVALID, "not a DepthNormal modality");
    LevelBuffers& lb = h->levels[0];
    size_t n = (size_t)lb.g.rows * lb.g.cols;
    ALLOC(h->d_dnidx, 3 * n);
    ScopedDevBuf tmpq;
    ALLOC(tmpq, n);
    const lmb200_modality& mod = h->cfg.modalities[index];
    launch_dn_quantize(h->d_depth[index].as<u16>() + (size_t)slot * h->depth_stride, h->depth_stride, tmpq.as<u8>(), n,
                       h->d_dnidx.as<int8_t>(), lb.g.rows, lb.g.cols, mod.distance_threshold, mod.difference_threshold,
                       h->d_normal_lut.as<u8>(), 1, st);
    CU(cudaStreamSynchronize(st));
    int rc = give(h->d_dnidx.p, 3 * n);
    tmpq.release();
    return rc;
  }
  if (kind == LMB200_DBG_SIMILARITY) {
    // index = global template index (classes in map order, template_id ascending); the map is computed by the
    // production coarse kernel (DUMP instantiation: same plan, realignment and packed adds, early exit off)
    if (index < 0 || index >= h->ntpl) return set_error(h, LMB200_E_INVALID, "bad template index");
    if (h->plan_dirty || h->templates_dirty) return set_error(h, LMB200_E_INVALID, "template set changed since the last match");
    const LevelGeom& g = h->levels[L - 1].g;
    const size_t n = (size_t)g.W * g.H * sizeof(u16);
    ScopedDevBuf d_map, d_one;
    ALLOC(d_map, n);
    ALLOC(d_one, sizeof(int));
    CU(cudaMemcpyAsync(d_one.p, &index, sizeof(int), cudaMemcpyHostToDevice, st));
    MatchParams mp = make_match_params(h, slot, 1, h->slot_threshold[slot]);
    mp.sel = d_one.as<int>(); mp.nsel = 1;
    launch_similarity_map(mp, make_level_params(h, L - 1, slot, true), 4 * h->max_nf_coarse > 255, d_map.as<u16>(), st);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    return give(d_map.p, n);
  }
  return set_error(h, LMB200_E_INVALID, "unknown debug kind");
}

}  // extern "C"

// ---------------------------------------------------------------- addTemplate (GPU quantisation + host selection)
namespace lmh {

// Quantise `sources` (any size) on the GPU at every pyramid level and hand the maps to the host
// extractor.  Uses a private single-frame workspace so the match plan is not disturbed.
int add_template_gpu(lmb200_detector* h, const char* class_id, const lmb200_image* sources, int n_sources,
                     const lmb200_image* object_mask, int* bb4, int* template_id) {
  int rc = ensure_device(h);
  if (rc) return rc;
  rc = upload_luts(h);
  if (rc) return rc;
  int rows, cols;
  rc = check_sources(h, sources, n_sources, &rows, &cols);
  if (rc) return rc;
  const int M = h->cfg.num_modalities, L = h->cfg.pyramid_levels;
  std::vector<u8> mask0;
  if (object_mask && object_mask->data) {
    if (object_mask->type != LMB200_8UC1 || object_mask->rows != rows || object_mask->cols != cols)
      return set_error(h, LMB200_E_SOURCES, "object_mask must be 8UC1 of the source size");
    size_t step = object_mask->step ? object_mask->step : (size_t)cols;
    mask0.resize((size_t)rows * cols);
    for (int r = 0; r < rows; ++r) std::memcpy(&mask0[(size_t)r * cols], (const u8*)object_mask->data + (size_t)r * step, cols);
  }
  cudaStream_t st = h->lanes[0].stream;
  // upstream default-inserts the class entry before extraction can fail
  std::vector<TemplatePyramid>& tps = h->classes[class_id];
  h->templates_dirty = true;
  TemplatePyramid tp((size_t)M * L);
  *template_id = -1;

  size_t n0 = (size_t)rows * cols;
  ScopedDevBuf d_src, d_src2, d_q, d_raw, d_mag;
  auto cleanup = [&]() { d_src.release(); d_src2.release(); d_q.release(); d_raw.release(); d_mag.release(); };
  std::vector<u8> hq(n0), hmask, hmask_next;
  std::vector<float> hmag(n0);
  for (int m = 0; m < M; ++m) {
    const lmb200_modality& mod = h->cfg.modalities[m];
    const lmb200_image& im = sources[m];
    int r = rows, c = cols;
    int num_features = mod.num_features, extract_threshold = mod.extract_threshold;
    hmask = mask0;
    if (mod.type == LMB200_COLOR_GRADIENT) {
      ALLOC(d_src, n0 * 3); ALLOC(d_src2, n0 * 3); ALLOC(d_q, n0); ALLOC(d_mag, n0 * sizeof(float));
      size_t rowb = (size_t)cols * 3, step = im.step ? im.step : rowb;
      CU(cudaMemcpy2DAsync(d_src.p, rowb, im.data, step, rowb, rows, cudaMemcpyHostToDevice, st));
      u8* cur = d_src.as<u8>();
      u8* nxt = d_src2.as<u8>();
      for (int l = 0; l < L; ++l) {
        if (l > 0) {
          launch_pyrdown_bgr(cur, 0, nxt, 0, r, c, 1, st);
          std::swap(cur, nxt);
          num_features /= 2;
          int nr = r / 2, nc = c / 2;
          if (!hmask.empty()) { hmask_next.resize((size_t)nr * nc); resize_nn_host(hmask.data(), r, c, hmask_next.data(), nr, nc); hmask.swap(hmask_next); }
          r = nr; c = nc;
        }
        launch_cg_quantize(cur, 0, d_q.as<u8>(), 0, d_mag.as<float>(), 0, r, c, mod.weak_threshold * mod.weak_threshold, 1, st);
        CU(cudaMemcpyAsync(hq.data(), d_q.p, (size_t)r * c, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(hmag.data(), d_mag.p, (size_t)r * c * sizeof(float), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (!extract_color_gradient(hq.data(), hmag.data(), hmask.empty() ? nullptr : hmask.data(), r, c, num_features,
                                    mod.strong_threshold, l, tp[(size_t)l * M + m])) { cleanup(); return LMB200_OK; }
      }
    } else {
      ALLOC(d_src, n0 * 2); ALLOC(d_q, n0); ALLOC(d_raw, n0);
      size_t rowb = (size_t)cols * 2, step = im.step ? im.step : rowb;
      CU(cudaMemcpy2DAsync(d_src.p, rowb, im.data, step, rowb, rows, cudaMemcpyHostToDevice, st));
      launch_dn_quantize(d_src.as<u16>(), 0, d_raw.as<u8>(), 0, nullptr, rows, cols, mod.distance_threshold,
                         mod.difference_threshold, h->d_normal_lut.as<u8>(), 1, st);
      launch_median5(d_raw.as<u8>(), 0, d_q.as<u8>(), 0, rows, cols, 1, st);
      CU(cudaMemcpyAsync(hq.data(), d_q.p, n0, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      std::vector<u8> cur(hq.begin(), hq.begin() + n0), nxt;
      for (int l = 0; l < L; ++l) {
        if (l > 0) {
          num_features /= 2; extract_threshold /= 2;
          int nr = r / 2, nc = c / 2;
          nxt.resize((size_t)nr * nc);
          resize_nn_host(cur.data(), r, c, nxt.data(), nr, nc);
          cur.swap(nxt);
          if (!hmask.empty()) { hmask_next.resize((size_t)nr * nc); resize_nn_host(hmask.data(), r, c, hmask_next.data(), nr, nc); hmask.swap(hmask_next); }
          r = nr; c = nc;
        }
        if (!extract_depth_normal(cur.data(), hmask.empty() ? nullptr : hmask.data(), r, c, num_features, extract_threshold, l,
                                  tp[(size_t)l * M + m])) { cleanup(); return LMB200_OK; }
      }
    }
  }
  cleanup();
  int bb[4];
  crop_templates(tp, bb);
  if (bb4) std::memcpy(bb4, bb, sizeof(bb));
  *template_id = (int)tps.size();
  tps.push_back(tp);
  return LMB200_OK;
}


// Bulk form of addTemplate: the reference's generateTemplates loop renders one view after the
// other and calls detector->addTemplate on each (HighLevelLinemod.cpp:68-110).  Here `n` rendered views of one
// size are quantised by the batched frame-side kernels, `BULK_CHUNK` views per launch, while a pool of host
// threads runs feature selection on the previous chunk (two pinned arenas).  Template ids are assigned in input
// order to the views whose extraction succeeds, exactly as n successive addTemplate calls would.
static const int BULK_CHUNK = 32;

int add_templates_bulk(lmb200_detector* h, const char* class_id, int n, const lmb200_image* sources, int n_sources,
                       const lmb200_image* masks, int* bb4, int* template_ids) {
  int rc = ensure_device(h);
  if (rc) return rc;
  rc = upload_luts(h);
  if (rc) return rc;
  const int M = h->cfg.num_modalities, L = h->cfg.pyramid_levels;
  int rows = 0, cols = 0;
  for (int i = 0; i < n; ++i) {
    int r, c;
    rc = check_sources(h, sources + (size_t)i * n_sources, n_sources, &r, &c);
    if (rc) return rc;
    if (i == 0) { rows = r; cols = c; }
    else if (r != rows || c != cols) return set_error(h, LMB200_E_SOURCES, "bulk addTemplate needs views of one size");
    if (masks && masks[i].data && (masks[i].type != LMB200_8UC1 || masks[i].rows != rows || masks[i].cols != cols))
      return set_error(h, LMB200_E_SOURCES, "object_mask must be 8UC1 of the source size");
  }
  std::vector<int> lr(L), lc(L);
  for (int l = 0; l < L; ++l) { lr[l] = l ? lr[l - 1] / 2 : rows; lc[l] = l ? lc[l - 1] / 2 : cols; }
  const size_t n0 = (size_t)rows * cols;
  // per-view record in the pinned arena: ColorGradient keeps (quantized, magnitude) of every level, DepthNormal the
  // level-0 map (coarser levels are decimations, done by the worker)
  std::vector<size_t> off_q((size_t)M * L, 0), off_mag((size_t)M * L, 0);
  size_t rec = 0;
  for (int m = 0; m < M; ++m) {
    bool cg = h->cfg.modalities[m].type == LMB200_COLOR_GRADIENT;
    for (int l = 0; l < (cg ? L : 1); ++l) {
      size_t nl = (size_t)lr[l] * lc[l];
      off_q[(size_t)m * L + l] = rec; rec += (nl + 15) & ~(size_t)15;
      if (cg) { off_mag[(size_t)m * L + l] = rec; rec += (nl * sizeof(float) + 15) & ~(size_t)15; }
    }
  }
  const int B = std::min(n, BULK_CHUNK);
  const int nchunks = (n + B - 1) / B;
  u8* arena[2] = {nullptr, nullptr};
  DevBuf d_a, d_b, d_q, d_mag;
  auto cleanup = [&]() {
    for (int a = 0; a < 2; ++a) if (arena[a]) { cudaFreeHost(arena[a]); arena[a] = nullptr; }
    d_a.release(); d_b.release(); d_q.release(); d_mag.release();
  };
  struct Guard { decltype(cleanup)& f; ~Guard() { f(); } } guard{cleanup};
  for (int a = 0; a < (nchunks > 1 ? 2 : 1); ++a) CU(cudaHostAlloc((void**)&arena[a], rec * B, cudaHostAllocDefault));
  ALLOC(d_a, n0 * 3 * B); ALLOC(d_b, n0 * 3 * B); ALLOC(d_q, n0 * B); ALLOC(d_mag, n0 * sizeof(float) * B);
  cudaStream_t st = h->lanes[0].stream;

  auto enqueue = [&](int k, u8* ar) -> int {
    const int first = k * B, cnt = std::min(B, n - first);
    for (int m = 0; m < M; ++m) {
      const lmb200_modality& mod = h->cfg.modalities[m];
      const bool cg = mod.type == LMB200_COLOR_GRADIENT;
      const size_t rowb = (size_t)cols * (cg ? 3 : 2);
      for (int i = 0; i < cnt; ++i) {
        const lmb200_image& im = sources[(size_t)(first + i) * n_sources + m];
        CU(cudaMemcpy2DAsync(d_a.as<u8>() + (size_t)i * rowb * rows, rowb, im.data, im.step ? im.step : rowb, rowb, rows,
                             cudaMemcpyHostToDevice, st));
      }
      if (cg) {
        u8* cur = d_a.as<u8>();
        u8* nxt = d_b.as<u8>();
        for (int l = 0; l < L; ++l) {
          const size_t nl = (size_t)lr[l] * lc[l];
          if (l > 0) {
            launch_pyrdown_bgr(cur, (size_t)lr[l - 1] * lc[l - 1] * 3, nxt, nl * 3, lr[l - 1], lc[l - 1], cnt, st);
            std::swap(cur, nxt);
          }
          launch_cg_quantize(cur, nl * 3, d_q.as<u8>(), nl, d_mag.as<float>(), nl, lr[l], lc[l],
                             mod.weak_threshold * mod.weak_threshold, cnt, st);
          CU(cudaMemcpy2DAsync(ar + off_q[(size_t)m * L + l], rec, d_q.p, nl, nl, cnt, cudaMemcpyDeviceToHost, st));
          CU(cudaMemcpy2DAsync(ar + off_mag[(size_t)m * L + l], rec, d_mag.p, nl * sizeof(float), nl * sizeof(float), cnt,
                               cudaMemcpyDeviceToHost, st));
        }
      } else {
        launch_dn_quantize(d_a.as<u16>(), n0, d_b.as<u8>(), n0, nullptr, rows, cols, mod.distance_threshold,
                           mod.difference_threshold, h->d_normal_lut.as<u8>(), cnt, st);
        launch_median5(d_b.as<u8>(), n0, d_q.as<u8>(), n0, rows, cols, cnt, st);
        CU(cudaMemcpy2DAsync(ar + off_q[(size_t)m * L], rec, d_q.p, n0, n0, cnt, cudaMemcpyDeviceToHost, st));
      }
    }
    CU(cudaGetLastError());
    return LMB200_OK;
  };

  struct View { TemplatePyramid tp; int bb[4]; bool ok; };
  auto extract_view = [&](int idx, const u8* recp, View& v) {
    v.tp.assign((size_t)M * L, Template());
    v.ok = true;
    // mask pyramid (nearest-neighbour halvings, as both quantizers' pyrDown do)
    std::vector<std::vector<u8>> mk;
    if (masks && masks[idx].data) {
      mk.resize(L);
      const lmb200_image& mi = masks[idx];
      size_t step = mi.step ? mi.step : (size_t)cols;
      mk[0].resize(n0);
      for (int r = 0; r < rows; ++r) std::memcpy(&mk[0][(size_t)r * cols], (const u8*)mi.data + (size_t)r * step, cols);
      for (int l = 1; l < L; ++l) {
        mk[l].resize((size_t)lr[l] * lc[l]);
        resize_nn_host(mk[l - 1].data(), lr[l - 1], lc[l - 1], mk[l].data(), lr[l], lc[l]);
      }
    }
    for (int m = 0; m < M && v.ok; ++m) {
      const lmb200_modality& mod = h->cfg.modalities[m];
      int num_features = mod.num_features, extract_threshold = mod.extract_threshold;
      std::vector<u8> cur, nxt;
      for (int l = 0; l < L && v.ok; ++l) {
        if (l > 0) { num_features /= 2; extract_threshold /= 2; }
        const u8* mask_l = mk.empty() ? nullptr : mk[l].data();
        if (mod.type == LMB200_COLOR_GRADIENT) {
          v.ok = extract_color_gradient(recp + off_q[(size_t)m * L + l], (const float*)(recp + off_mag[(size_t)m * L + l]), mask_l,
                                        lr[l], lc[l], num_features, mod.strong_threshold, l, v.tp[(size_t)l * M + m]);
        } else {
          const u8* q = recp + off_q[(size_t)m * L];
          if (l > 0) {
            nxt.resize((size_t)lr[l] * lc[l]);
            resize_nn_host(l == 1 ? q : cur.data(), lr[l - 1], lc[l - 1], nxt.data(), lr[l], lc[l]);
            cur.swap(nxt);
            q = cur.data();
          }
          v.ok = extract_depth_normal(q, mask_l, lr[l], lc[l], num_features, extract_threshold, l, v.tp[(size_t)l * M + m]);
        }
      }
    }
    if (v.ok) crop_templates(v.tp, v.bb);
  };

  // upstream default-inserts the class entry before extraction can fail
  h->classes[class_id];
  h->templates_dirty = true;
  for (int i = 0; i < n; ++i) template_ids[i] = -1;
  const int hw = (int)std::thread::hardware_concurrency();
  const int nthreads = std::max(1, std::min(std::min(hw > 0 ? hw : 1, 16), B));
  rc = enqueue(0, arena[0]);
  if (rc) return rc;
  std::vector<View> views(B);
  for (int k = 0; k < nchunks; ++k) {
    CU(cudaStreamSynchronize(st));
    if (k + 1 < nchunks) { rc = enqueue(k + 1, arena[(k + 1) & 1]); if (rc) return rc; }
    const int first = k * B, cnt = std::min(B, n - first);
    const u8* ar = arena[k & 1];
    std::atomic<int> next(0);
    auto work = [&]() { for (int i; (i = next.fetch_add(1)) < cnt;) extract_view(first + i, ar + (size_t)i * rec, views[i]); };
    std::vector<std::thread> pool;
    for (int t = 1; t < std::min(nthreads, cnt); ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    std::vector<TemplatePyramid>& tps = h->classes[class_id];
    for (int i = 0; i < cnt; ++i) {
      if (!views[i].ok) continue;
      template_ids[first + i] = (int)tps.size();
      if (bb4) std::memcpy(bb4 + (size_t)(first + i) * 4, views[i].bb, sizeof(views[i].bb));
      tps.push_back(std::move(views[i].tp));
    }
  }
  return LMB200_OK;
}

}  // namespace lmh

// ---------------------------------------------------------------- template-sharded multi-GPU fetch
extern "C" int lmb200_fetch_resident_allgather(lmb200_handle h, int first_slot, int count, lmb200_match_rec* out,
                                                size_t cap, size_t* offsets) {
  if (!h || !h->device_ready) return set_error(h, LMB200_E_INVALID, "nothing to fetch");
  if (first_slot < 0 || count <= 0 || first_slot + count > h->slots) return set_error(h, LMB200_E_INVALID, "bad slot range");
  if (!h->nccl_comm) return set_error(h, LMB200_E_COMM, "communicator not initialised (lmb200_comm_init)");
  cudaSetDevice(h->device);
  const int world = h->comm_world;
  const bool trace = std::getenv("LMB200_TRACE") != nullptr;
  const double t_begin = now_ms();
  auto emit_lists = [&](const std::vector<lmb200_match_rec>& recs, const std::vector<size_t>& offs) {
    const size_t total = offs[count], w = std::min(total, cap);
    if (w && out) std::memcpy(out, recs.data(), w * sizeof(lmb200_match_rec));
    if (offsets) for (int i = 0; i <= count; ++i) offsets[i] = offs[i];
    return w < total ? LMB200_E_TRUNCATED : LMB200_OK;
  };
  // 1. A step submitted through lmb200_match_resident_sharded: its gather rode on the compute lane and the epilogue thread
  //    has merged it (or is about to): pick the finished lists up.
  for (auto& job : h->jobs) {
    if (!(job.state != 0 && job.first == first_slot && job.count == count)) continue;
    if (job.state == 4) {   // device epilogue: the finished lists are in pinned memory once ev fires
      CU(cudaEventSynchronize(job.ev));
      const double t_dev = now_ms() - t_begin;
      bool ok = job.generation == h->buffer_generation;
      for (int i = 0; i < count && ok; ++i) ok = job.hdr_host[2 * i].z == 0;
      job.state = 0;
      if (!ok) break;       // a flag (store overflow, record area too small, list too long for the device sort): synchronous path below
      size_t base = 0;
      int status = LMB200_OK;
      for (int i = 0; i < count; ++i) {
        const int4 hd = job.hdr_host[2 * i], c4 = job.hdr_host[2 * i + 1];
        const size_t nfin = (size_t)hd.x, room = cap > base ? cap - base : 0, w = std::min(nfin, room);
        if (offsets) offsets[i] = base;
        if (w && out) std::memcpy(out + base, job.fin_host + hd.y, w * sizeof(lmb200_match_rec));
        if (w < nfin) status = LMB200_E_TRUNCATED;
        base += nfin;
        h->prof.candidates += hd.w; h->prof.matches += hd.x;
        h->prof.bytes_local += (long long)(((unsigned long long)(u32)c4.y << 32) | (u32)c4.x);
        h->prof.chunks_coarse += (long long)(((unsigned long long)(u32)c4.w << 32) | (u32)c4.z);
      }
      if (offsets) offsets[count] = base;
      if (h->profiling) collect_profile(h);
      if (trace && h->comm_rank == 0)
        std::fprintf(stderr, "[lmb200 trace] allgather fetch: waited %.3f ms for the device (kernels + gather + device epilogue), copy-out %.3f ms\n",
                     t_dev, now_ms() - t_begin - t_dev);
      return status;
    }
    shard_job_wait(h, &job);
    const bool ready = job.state == 2 && job.generation == h->buffer_generation;
    job.state = 0;
    if (!ready) break;      // flags in the gathered headers, or the stores grew meanwhile: synchronous path below
    h->prof.bytes_local += job.bytes_local; h->prof.chunks_coarse += job.chunks_coarse;
    h->prof.candidates += job.candidates; h->prof.matches += job.matches;
    if (h->profiling) collect_profile(h);
    const int status = emit_lists(job.recs, job.offs);
    if (trace && h->comm_rank == 0)
      std::fprintf(stderr, "[lmb200 trace] allgather fetch: waited %.3f ms for the epilogue thread (its wait for the device %.3f ms, reorder %.3f ms, sort/unique %.3f ms)\n",
                   now_ms() - t_begin, job.t_wait, job.t_reorder, job.t_sort);
    return status;
  }
  // 2. Synchronous path (plain lmb200_match_resident with a template shard, or a step whose headers carry a flag).  One round
  //    trip: a kernel packs {count, flags, counters, matches} of every frame, the buffers are all-gathered and copied to
  //    pinned host memory.  Every rank sees every flag and count, so all ranks take the same decisions: a rank that
  //    overflowed its candidate store grows it and redoes its template side, a record area that is too small doubles, and
  //    the gather is repeated.
  cudaStream_t st = resident_fetch_stream(h, first_slot, count);
  {
    int rc = revalidate_slots(h, first_slot, count);
    if (rc) return rc;
  }
  if (h->gather_cap <= 0) h->gather_cap = 256;    // average records per frame the buffer holds; doubles when a step needs more
  int gcap = 0;
  for (;;) {
    const int rec_cap = h->gather_cap * count;
    const size_t bytes = ((size_t)2 * count + rec_cap) * sizeof(Cand);
    ALLOC(h->d_gather_send, bytes);
    ALLOC(h->d_gather_recv, bytes * world);
    if (h->h_gather_bytes < bytes * world) {
      if (h->h_gather) { cudaFreeHost(h->h_gather); h->h_gather = nullptr; }
      CU(cudaHostAlloc((void**)&h->h_gather, bytes * world, cudaHostAllocDefault));
      h->h_gather_bytes = bytes * world;
    }
    launch_gather_pack(h->d_ctr.as<SlotCtr>() + first_slot, h->d_out.as<Cand>() + (size_t)first_slot * h->out_cap, h->out_cap,
                       h->d_gather_send.as<Cand>(), rec_cap, count, st);
    int rc = comm_allgather(h, h->d_gather_send.p, h->d_gather_recv.p, bytes, st, true);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h->h_gather, h->d_gather_recv.p, bytes * world, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    bool store_over = false, my_over = false, small = false;
    for (int r = 0; r < world; ++r)
      for (int i = 0; i < count; ++i) {
        const int fl = h->h_gather[(size_t)r * ((size_t)2 * count + rec_cap) + 2 * i].x;
        if (fl & 1) { store_over = true; if (r == h->comm_rank) my_over = true; }
        if (fl & 2) small = true;
      }
    if (my_over) {
      float thr = h->slot_threshold[first_slot];
      rc = grow_capacity(h);
      if (rc) return rc;
      rc = run_matching(h, first_slot, count, thr, h->lanes[0].stream);
      if (rc) return rc;
      CU(cudaStreamSynchronize(h->lanes[0].stream));
      st = h->lanes[0].stream;
    }
    if (store_over) continue;
    if (small) { h->gather_cap *= 2; continue; }   // every rank sees the flag: all double together
    gcap = rec_cap;
    break;
  }
  const Cand* mine = h->h_gather + (size_t)h->comm_rank * ((size_t)2 * count + gcap);
  for (int i = 0; i < count; ++i) {
    const Cand& st1 = mine[2 * i + 1];
    h->prof.bytes_local += (long long)(((unsigned long long)(u32)st1.x << 32) | (u32)st1.tsel);
    h->prof.chunks_coarse += (long long)(((unsigned long long)(u32)__builtin_bit_cast(int, st1.sim) << 32) | (u32)st1.y);
  }
  if (h->profiling) collect_profile(h);
  const double t_gather = now_ms();
  std::vector<lmb200_match_rec> recs; std::vector<size_t> offs;
  long long nc = 0, nm = 0;
  double t_reorder = 0, t_sort = 0;
  merge_gathered(h, h->h_gather, gcap, count, (unsigned)std::max(1, h->host_threads), recs, offs, &nc, &nm, &t_reorder, &t_sort);
  h->prof.candidates += nc; h->prof.matches += nm;
  const int status = emit_lists(recs, offs);
  if (trace && h->comm_rank == 0)
    std::fprintf(stderr, "[lmb200 trace] allgather fetch (synchronous): gather+D2H %.3f ms, reorder %.3f ms, sort/unique %.3f ms\n",
                 t_gather - t_begin, t_reorder, t_sort);
  return status;
}

// Test hook (include/lmb200.h): the HOST merge of a template-sharded step — the code the synchronous fetch path and the
// epilogue thread run (merge_gathered) — on caller-made gathered buffers; no device is touched, so the CPU suite and the
// gloo multi-rank test cover the wire format and the generation-order merge.
extern "C" int lmb200_debug_merge_gathered(const int32_t* gathered, int world, int frames, int gcap, const int32_t* pos_of_g,
                                           const int32_t* g_class, const int32_t* g_tid, int ntpl, lmb200_match_rec* out, size_t cap,
                                           size_t* offsets) {
  if (!gathered || !g_class || !g_tid || !offsets || world < 1 || frames < 1 || gcap < 0 || ntpl < 1) return LMB200_E_INVALID;
  static_assert(sizeof(Cand) == 4 * sizeof(int32_t), "a gathered record is four 32-bit words");
  lmb200_detector tmp;
  tmp.comm_world = world; tmp.comm_rank = 0;
  tmp.ntpl = ntpl;
  tmp.g_class.assign(g_class, g_class + ntpl);
  tmp.g_tid.assign(g_tid, g_tid + ntpl);
  tmp.shard_interleaved = pos_of_g != nullptr;
  if (pos_of_g) tmp.pos_of_g.assign(pos_of_g, pos_of_g + ntpl);
  tmp.host_threads = 2;
  const Cand* G = reinterpret_cast<const Cand*>(gathered);
  if (gathered_needs_redo(G, gcap, frames, world)) return set_error(nullptr, LMB200_E_INVALID, "a gathered header carries a flag");
  for (int r = 0; r < world; ++r)     // never index past a rank's record area
    for (int i = 0; i < frames; ++i) {
      const Cand& h0 = G[(size_t)r * ((size_t)2 * frames + gcap) + 2 * i];
      if (h0.tsel < 0 || h0.y < 0 || (long long)h0.y + h0.tsel > gcap) return LMB200_E_INVALID;
    }
  std::vector<lmb200_match_rec> recs; std::vector<size_t> offs;
  merge_gathered(&tmp, G, gcap, frames, 2u, recs, offs, nullptr, nullptr, nullptr, nullptr);
  const size_t total = offs[frames], w = std::min(total, cap);
  if (w && out) std::memcpy(out, recs.data(), w * sizeof(lmb200_match_rec));
  for (int i = 0; i <= frames; ++i) offsets[i] = offs[i];
  return w < total ? LMB200_E_TRUNCATED : LMB200_OK;
}
