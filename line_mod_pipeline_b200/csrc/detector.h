// detector.h — host-side state of one lmb200 handle (internal; the public surface is include/lmb200.h).
#pragma once
#include <cuda_runtime.h>

#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lmb200.h"
#include "kernels.cuh"

#define LMB200_LANES 6

namespace lmh {

struct Feature { int x, y, label; };
struct Template {
  int width = 0, height = 0, pyramid_level = 0;
  std::vector<Feature> features;
};
typedef std::vector<Template> TemplatePyramid;  // index = level * num_modalities + modality

struct Match {  // cv::linemod::Match with class_id as index into the sorted class list
  int x, y;
  float similarity;
  int class_index;
  int template_id;
  bool operator<(const Match& r) const {
    if (similarity != r.similarity) return similarity > r.similarity;
    return template_id < r.template_id;
  }
  bool operator==(const Match& r) const {
    return x == r.x && y == r.y && similarity == r.similarity && class_index == r.class_index;
  }
};

// sort + unique, exactly upstream's epilogue of Detector::match
void finalize_matches(std::vector<Match>& m);

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int alloc(size_t n);  // returns cudaError
  bool view = false;    // alias of another allocation: never freed here
  void release();
  void alias(void* ptr, size_t n);
  template <typename T> T* as() const { return (T*)p; }
};

// function-local device buffer: released on every return path (the CU/ALLOC macros return early on errors)
struct ScopedDevBuf : DevBuf {
  ScopedDevBuf() = default;
  ScopedDevBuf(const ScopedDevBuf&) = delete;
  ScopedDevBuf& operator=(const ScopedDevBuf&) = delete;
  ~ScopedDevBuf() { release(); }
};

struct LevelBuffers {
  lmk::LevelGeom g;
  size_t q_stride = 0, lm_stride = 0, bgr_stride = 0, lmn_stride = 0;
  bool fast_spread = false;            // round-2 spread kernels cover this geometry (kernels.cuh: spread_fast_covers)
  bool dn_inplace = false;             // DepthNormal at this level = exact 2^l decimation of the level-0 map, read in place
  DevBuf bgr[LMB200_MAX_MODALITIES];   // CG source at this level (level 0 = uploaded frame)
  DevBuf q[LMB200_MAX_MODALITIES];     // quantized map
  DevBuf mask[LMB200_MAX_MODALITIES];  // optional mask pyramid
  DevBuf lm[LMB200_MAX_MODALITIES];    // linear memories [slots][8*rows*cols + pad]
  DevBuf lmn[LMB200_MAX_MODALITIES];   // coarsest level only: nibble-packed copy [slots][4*rows*cols + pad]
};

struct Lane {
  cudaStream_t stream = nullptr;
};

struct ProfRec { int family; cudaEvent_t a, b; };
// One template-sharded step in flight (lmb200_match_resident_sharded).  The match gather — pack kernel, ncclAllGather and
// the copy to pinned host memory — rides on the compute lane right behind the step's kernels, and the host epilogue
// (reference generation order, record -> Match, std::sort, std::unique of every frame) runs on the handle's epilogue
// thread as soon as `ev` fires: neither costs the submitting thread anything, and lmb200_fetch_resident_allgather only
// picks the finished lists up.  No second collective, no host-synchronous rendezvous between the ranks.
struct ShardJob {
  int first = -1, count = 0, cap = 0;
  long long generation = -1;
  int state = 0;        // 0 idle, 1 queued for the epilogue thread, 2 finished lists ready, 3 the gathered flags ask for the synchronous path,
                        // 4 device epilogue enqueued (finished lists land in fin_host when ev fires)
  DevBuf send, recv;
  lmk::Cand* host = nullptr; size_t host_bytes = 0;
  cudaEvent_t ev = nullptr, ev_q = nullptr;      // results landed in pinned memory / quantized maps of the step gathered
  cudaEvent_t ev_g = nullptr;                    // match gather done on the compute lane (the device epilogue waits for it on lane 1)
  DevBuf fin_dev;                                // device epilogue: finished records, device scratch
  lmk::EpiMatch* fin_host = nullptr; size_t fin_cap = 0;   // ... and their pinned, device-mapped host mirror
  int4* hdr_host = nullptr; int hdr_frames = 0;  // pinned, device-mapped: per frame {n_final, offset, flags, n_in}, {counters}
  std::vector<lmb200_match_rec> recs; std::vector<size_t> offs;   // finished lists, frames back to back
  long long candidates = 0, matches = 0, bytes_local = 0, chunks_coarse = 0;
  double t_wait = 0, t_reorder = 0, t_sort = 0;
};
struct ResidentMark { cudaEvent_t ev = nullptr; int first = 0, count = 0; };  // completion of one lmb200_match_resident call

// One submitted lmb200_match_batch: chunk schedule, completion events and per-frame pinned result staging.
struct BatchTicket {
  bool active = false;
  std::vector<lmb200_image> frames;
  int n_frames = 0, n_sources = 0;
  float threshold = 0.f;
  std::string sel_key;
  std::vector<std::string> class_ids;
  std::vector<int> cf0, ccnt;
  std::vector<cudaEvent_t> events;        // per chunk: h2d done, chunk done
  lmk::SlotCtr* b_ctr = nullptr; lmk::Cand* b_out = nullptr;
  int cap_frames = 0, head = 0, head_used = 0;
  long long generation = 0;               // buffer_generation at enqueue time
  bool streaming = false;                 // submitted through submit/collect (throughput) rather than the blocking call
};

}  // namespace lmh

struct lmb200_detector {
  lmb200_config cfg;
  std::string err;
  uint8_t sim_lut[256];
  uint8_t normal_lut[8000];
  bool normal_lut_standin = true;   // still the built-in stand-in table (capi.cpp:default_normal_lut)
  bool standin_warned = false;
  std::string warnings;

  // ---- template store (host truth) ----
  std::map<std::string, std::vector<lmh::TemplatePyramid>> classes;
  bool templates_dirty = true;  // device tables stale
  bool plan_dirty = true;       // offsets stale

  // ---- device ----
  bool device_ready = false;
  int device = -1;
  lmh::Lane lanes[LMB200_LANES];   // 0: compute, 1: copy / device epilogue (high priority), 2..4: extra compute lanes of the batch path,
                                   // 5: frame side of the template-sharded step (quantisers + map all-gather; high priority)
  lmh::DevBuf d_table, d_normal_lut;
  bool luts_dirty = true;

  // template tables (per level)
  int ntpl = 0;
  int max_nf_coarse = 0;                           // max over templates of sum_m nf at the coarsest level
  std::vector<int> g_class, g_tid;                 // global template index -> (class index, template id)
  std::vector<std::string> class_list;             // sorted class ids (index = class_index)
  lmh::DevBuf d_hdr[LMB200_MAX_LEVELS], d_feat[LMB200_MAX_LEVELS], d_offs[LMB200_MAX_LEVELS];
  std::vector<double> tpl_cost;                    // coarse bytes per template for the current plan

  // selection
  std::string sel_key;
  std::vector<int> h_sel;                          // global indices scored by this handle (after sharding)
  lmh::DevBuf d_sel;
  long long sel_bytes_coarse = 0;
  int shard_rank = 0, shard_world = 1;
  bool shard_interleaved = false; std::vector<int> pos_of_g;  // interleaved shards: global template index -> selection position

  // frame plan
  int rows = 0, cols = 0, slots = 0;
  int cand_cap = 0, out_cap = 0;
  bool masks_in_use = false;
  bool dn_materialize = false;          // some DepthNormal level needs its decimated map in memory (generic spread path or odd sizes):
                                        // then every level's map is written (resize chain); else coarser levels are read in place
  bool generic_frame_side = false;      // LMB200_GENERIC_FRAME=1: round 1's generic frame-side kernels (A/B runs, fallback tests)
  std::vector<lmh::LevelBuffers> levels;
  lmh::DevBuf d_depth[LMB200_MAX_MODALITIES], d_dnraw[LMB200_MAX_MODALITIES], d_mag, d_dnidx;
  size_t depth_stride = 0;
  lmh::DevBuf d_frames; size_t frame_bytes = 0, src_off[LMB200_MAX_MODALITIES] = {0, 0, 0, 0};  // [slots][sources back to back]
  lmh::DevBuf d_hue_bits;                           // post-match colour check: bit mask of the in-range pixels of one frame
  lmh::DevBuf d_resp_sum;                           // [slots][MAX_MOD] response sums of the coarsest linear memories
  lmh::DevBuf d_cand, d_ctr, d_tpl_start, d_tpl_cnt, d_tpl_alive, d_out;
  int nsel_stride = 0;
  // pinned host mirrors
  lmk::SlotCtr* h_ctr = nullptr;
  lmk::Cand* h_out = nullptr; int h_head = 0;      // first h_head records of every slot
  std::vector<float> slot_threshold;
  std::vector<long long> slot_gen;        // buffer_generation the slot's match results were computed under (-1: none)
  // batches in flight (lmb200_match_batch_submit / _collect)
  lmh::BatchTicket tickets[2];
  std::vector<cudaEvent_t> group_done; std::vector<char> group_used; int b_groups = 0;
  long long chunk_seq = 0, buffer_generation = 0;
  bool blocking_submit = false;           // lmb200_match_batch is running its own submit
  lmh::ResidentMark resident_marks[4]; unsigned resident_next = 0;

  // profiling
  bool profiling = false;
  cudaEvent_t upload_ev = nullptr; bool upload_pending = false;
  // single-frame CUDA graph (lmb200_match)
  cudaGraphExec_t match_graph = nullptr; long long plan_epoch = 0, graph_epoch = -1; uint32_t graph_thr_bits = 0; bool graph_early_exit = true;
  long long graph_launches[LMB200_K_COUNT] = {0}; bool use_graph = true;
  cudaEvent_t fork_ev[4] = {nullptr, nullptr, nullptr, nullptr};   // fork / join events of the graph's per-modality branches
  bool upload_async = false;             // lmb200_set_option("upload_async"): lmb200_upload_frames returns without synchronising
  bool resident_overlap = true;          // lmb200_set_option("resident_overlap"): lmb200_match_resident runs the frame side on the frame lane
  cudaEvent_t fs_done = nullptr;
  int host_threads = 8;                  // lmb200_set_option("host_threads"): threads of the host epilogue (sort/unique) of one fetch/collect;
                                         // one process per GPU shares the host's cores with its peers
  bool early_exit = true;                // lmb200_set_option("early_exit"): measurement runs switch the coarse kernel's exact exit off
  std::vector<lmh::ProfRec> prof_pending;
  std::vector<cudaEvent_t> event_pool;
  lmb200_profile prof;
  cudaEvent_t timer[2] = {nullptr, nullptr};

  // comm
  void* nccl_comm = nullptr; int comm_rank = 0, comm_world = 1;
  void* nccl_comm_fetch = nullptr;          // second communicator (ncclCommSplit) for the result-fetch collectives
  lmh::DevBuf d_gather_send, d_gather_recv; int gather_cap = 0;
  lmk::Cand* h_gather = nullptr; size_t h_gather_bytes = 0;
  lmh::ShardJob jobs[4]; unsigned job_next = 0;
  std::thread epi_thread; std::mutex epi_mu; std::condition_variable epi_cv, epi_done_cv;   // epilogue thread of the sharded steps
  std::deque<lmh::ShardJob*> epi_queue; bool epi_stop = false;
  bool shard_device_epilogue = true;        // lmb200_set_option("shard_device_epilogue"): std::sort/std::unique of the sharded step on the device
                                            // (kernels_epilogue.cu); 0: on the handle's epilogue thread
  lmh::DevBuf d_gclass, d_gtid, d_posg; long long epi_tables_epoch = -1;   // device copies of g_class / g_tid / pos_of_g
  int shard_overlap = 1;                // lmb200_set_option("shard_overlap"): quantise + map all-gather of step k+1 on a lane of their own

  // scratch for lmb200_get_template
  std::vector<lmb200_feature> tmp_features;
};

namespace lmh {
// detector.cu
int ensure_device(lmb200_detector* h);
int set_error(lmb200_detector* h, int code, const std::string& msg);
int upload_templates_now(lmb200_detector* h);
int add_template_gpu(lmb200_detector* h, const char* class_id, const lmb200_image* sources, int n_sources,
                     const lmb200_image* object_mask, int* bb4, int* template_id);
int add_templates_bulk(lmb200_detector* h, const char* class_id, int n, const lmb200_image* sources, int n_sources,
                       const lmb200_image* masks, int* bb4, int* template_ids);
int cuda_fail(lmb200_detector* h, cudaError_t e, const char* what);
// extract.cpp (host feature selection; images are level-sized row-major arrays)
bool extract_color_gradient(const uint8_t* quant, const float* magnitude, const uint8_t* mask /*nullable*/,
                            int rows, int cols, int num_features, float strong_threshold, int level, Template& out);
bool extract_depth_normal(const uint8_t* quant, const uint8_t* mask /*nullable*/, int rows, int cols,
                          int num_features, int extract_threshold, int level, Template& out);
void crop_templates(TemplatePyramid& tp, int bb[4]);
void resize_nn_host(const uint8_t* src, int rows, int cols, uint8_t* dst, int drows, int dcols);
// persistence.cpp
int write_detector_file(lmb200_detector* h, const char* path);
int read_detector_file(const char* path, int device, lmb200_handle* out, std::string& err);
int write_class_file(lmb200_detector* h, const std::string& class_id, const char* path);
int read_class_file(lmb200_detector* h, const char* path, std::string& err, const char* override_id = nullptr);
int write_cache_file(lmb200_detector* h, const char* path);
int read_cache_file(const char* path, int device, lmb200_handle* out, std::string& err);
void set_create_error(const std::string& msg);
// comm.cpp
int comm_unique_id(uint8_t* id128, std::string& err);
int comm_init(lmb200_detector* h, const uint8_t* id128, int rank, int world);
int comm_destroy(lmb200_detector* h);
void shard_jobs_stop(lmb200_detector* h);   // detector.cu: drains and joins the epilogue thread
int comm_allgather(lmb200_detector* h, const void* send, void* recv, size_t bytes, cudaStream_t st, bool fetch_path = false);
int comm_group_begin(lmb200_detector* h);
int comm_group_end(lmb200_detector* h);
void default_normal_lut(uint8_t* out8000);
void default_similarity_lut(uint8_t* out256);
}  // namespace lmh
