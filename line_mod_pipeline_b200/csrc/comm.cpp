// comm.cpp — multi-GPU plumbing: NCCL (dlopen'ed, so the library loads without it), the
// cost-balanced template shard plan and the deterministic merge of per-rank match lists
// (SURVEY.md §8e).  The reference has no multi-GPU path; this is the template-sharded design
// BASELINE.json's north_star names: each GPU scores a contiguous block of the generation-ordered
// template list, match buffers are combined with one small ncclAllGather over NVLink, and the
// rank-ordered concatenation (= reference generation order) goes through the same std::sort /
// std::unique epilogue, so the result is identical to the 1-GPU run by construction.
#include <dlfcn.h>

#include <cmath>
#include <cstring>
#include <mutex>

#include "detector.h"

namespace lmh {

namespace {
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int (*fn_get_uid)(ncclUniqueId_t*);
typedef int (*fn_init_rank)(void**, int, ncclUniqueId_t, int);
typedef int (*fn_destroy)(void*);
typedef int (*fn_allgather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*fn_errstr)(int);
typedef int (*fn_group)(void);
typedef int (*fn_split)(void*, int, int, void**, void*);

struct Nccl {
  void* lib = nullptr;
  fn_get_uid get_uid = nullptr; fn_init_rank init_rank = nullptr; fn_destroy destroy = nullptr;
  fn_allgather allgather = nullptr; fn_errstr errstr = nullptr;
  fn_group group_start = nullptr, group_end = nullptr;
  fn_split split = nullptr;
  std::string err;
};
Nccl g_nccl;
std::once_flag g_once;

void load_nccl() {
  // RTLD_NOLOAD first: reuse the copy a host process (e.g. PyTorch) already mapped
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (g_nccl.lib) break; }
  if (!g_nccl.lib) for (const char* n : names) { g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
  if (!g_nccl.lib) { g_nccl.err = "cannot dlopen libnccl.so.2"; return; }
  g_nccl.get_uid = (fn_get_uid)dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.init_rank = (fn_init_rank)dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.destroy = (fn_destroy)dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.allgather = (fn_allgather)dlsym(g_nccl.lib, "ncclAllGather");
  g_nccl.errstr = (fn_errstr)dlsym(g_nccl.lib, "ncclGetErrorString");
  g_nccl.group_start = (fn_group)dlsym(g_nccl.lib, "ncclGroupStart");
  g_nccl.group_end = (fn_group)dlsym(g_nccl.lib, "ncclGroupEnd");
  g_nccl.split = (fn_split)dlsym(g_nccl.lib, "ncclCommSplit");
  if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.destroy || !g_nccl.allgather) g_nccl.err = "libnccl lacks a required symbol";
}
bool nccl_ok(std::string& err) {
  std::call_once(g_once, load_nccl);
  if (!g_nccl.err.empty()) { err = g_nccl.err; return false; }
  return true;
}
std::string nccl_msg(const char* what, int rc) {
  return std::string(what) + ": " + (g_nccl.errstr ? g_nccl.errstr(rc) : "nccl error") + " (" + std::to_string(rc) + ")";
}
}  // namespace

int comm_unique_id(uint8_t* id128, std::string& err) {
  if (!nccl_ok(err)) return LMB200_E_COMM;
  ncclUniqueId_t id;
  int rc = g_nccl.get_uid(&id);
  if (rc) { err = nccl_msg("ncclGetUniqueId", rc); return LMB200_E_COMM; }
  std::memcpy(id128, id.internal, 128);
  return LMB200_OK;
}

int comm_init(lmb200_detector* h, const uint8_t* id128, int rank, int world) {
  std::string err;
  if (!nccl_ok(err)) return set_error(h, LMB200_E_COMM, err);
  int rc = ensure_device(h);
  if (rc) return rc;
  if (h->nccl_comm) comm_destroy(h);
  ncclUniqueId_t id;
  std::memcpy(id.internal, id128, 128);
  void* comm = nullptr;
  rc = g_nccl.init_rank(&comm, world, id, rank);
  if (rc) return set_error(h, LMB200_E_COMM, nccl_msg("ncclCommInitRank", rc));
  h->nccl_comm = comm; h->comm_rank = rank; h->comm_world = world;
  // A second communicator over the same ranks for the result-fetch collectives (copy lane): NCCL serialises the
  // operations of ONE communicator in issue order, which would make the match gather of step k-1 wait for the quantized-map
  // all-gather of step k+1 that the compute lane has already queued.  Optional: older NCCL without ncclCommSplit shares one.
  h->nccl_comm_fetch = nullptr;
  if (g_nccl.split && world > 1) {
    void* c2 = nullptr;
    if (g_nccl.split(comm, 0, rank, &c2, nullptr) == 0 && c2) h->nccl_comm_fetch = c2;
  }
  return LMB200_OK;
}

int comm_destroy(lmb200_detector* h) {
  if (h->nccl_comm_fetch && g_nccl.destroy) g_nccl.destroy(h->nccl_comm_fetch);
  h->nccl_comm_fetch = nullptr;
  if (h->nccl_comm && g_nccl.destroy) g_nccl.destroy(h->nccl_comm);
  h->nccl_comm = nullptr; h->comm_world = 1; h->comm_rank = 0;
  return LMB200_OK;
}

int comm_allgather(lmb200_detector* h, const void* send, void* recv, size_t bytes, cudaStream_t st, bool fetch_path) {
  if (!h->nccl_comm) return set_error(h, LMB200_E_COMM, "communicator not initialised (lmb200_comm_init)");
  void* comm = (fetch_path && h->nccl_comm_fetch) ? h->nccl_comm_fetch : h->nccl_comm;
  int rc = g_nccl.allgather(send, recv, bytes, /*ncclInt8*/ 0, comm, st);
  if (rc) return set_error(h, LMB200_E_COMM, nccl_msg("ncclAllGather", rc));
  return LMB200_OK;
}

// several collectives fused into one NCCL launch (optional: absent symbols make these no-ops)
int comm_group_begin(lmb200_detector* h) {
  if (g_nccl.group_start) { int rc = g_nccl.group_start(); if (rc) return set_error(h, LMB200_E_COMM, nccl_msg("ncclGroupStart", rc)); }
  return LMB200_OK;
}
int comm_group_end(lmb200_detector* h) {
  if (g_nccl.group_end) { int rc = g_nccl.group_end(); if (rc) return set_error(h, LMB200_E_COMM, nccl_msg("ncclGroupEnd", rc)); }
  return LMB200_OK;
}

}  // namespace lmh

using namespace lmh;

extern "C" {

int lmb200_comm_unique_id(uint8_t* unique_id128) {
  if (!unique_id128) return LMB200_E_INVALID;
  std::string err;
  int rc = comm_unique_id(unique_id128, err);
  if (rc) set_error(nullptr, rc, err);
  return rc;
}
int lmb200_comm_init(lmb200_handle h, const uint8_t* unique_id128, int rank, int world) {
  if (!h || !unique_id128 || world < 1 || rank < 0 || rank >= world) return LMB200_E_INVALID;
  return comm_init(h, unique_id128, rank, world);
}
int lmb200_comm_destroy(lmb200_handle h) { return h ? comm_destroy(h) : LMB200_E_INVALID; }

// Contiguous split with prefix-cost boundaries at k/world of the total (deterministic, rank-independent).
int lmb200_shard_plan(const double* costs, int n, int world, int* begin) {
  if (n < 0 || world < 1 || !begin || (n > 0 && !costs)) return LMB200_E_INVALID;
  double total = 0;
  for (int i = 0; i < n; ++i) total += costs[i] > 0 ? costs[i] : 0;
  begin[0] = 0;
  int i = 0;
  double run = 0;
  for (int r = 1; r < world; ++r) {
    double target = total * r / world;
    while (i < n && run + (costs[i] > 0 ? costs[i] : 0) * 0.5 < target) { run += costs[i] > 0 ? costs[i] : 0; ++i; }
    begin[r] = i;
  }
  begin[world] = n;
  return LMB200_OK;
}

int lmb200_merge_matches(const lmb200_match_rec* const* parts, const size_t* counts, int world,
                         lmb200_match_rec* out, size_t cap, size_t* n_out) {
  if (!parts || !counts || world < 1 || !n_out) return LMB200_E_INVALID;
  std::vector<Match> m;
  for (int r = 0; r < world; ++r)
    for (size_t i = 0; i < counts[r]; ++i) {
      const lmb200_match_rec& p = parts[r][i];
      m.push_back(Match{p.x, p.y, p.similarity, p.class_index, p.template_id});
    }
  finalize_matches(m);
  *n_out = m.size();
  size_t w = m.size() < cap ? m.size() : cap;
  for (size_t i = 0; i < w && out; ++i) {
    out[i].x = m[i].x; out[i].y = m[i].y; out[i].similarity = m[i].similarity;
    out[i].class_index = m[i].class_index; out[i].template_id = m[i].template_id;
  }
  return (w < m.size() || !out) && m.size() ? LMB200_E_TRUNCATED : LMB200_OK;
}

}  // extern "C"
