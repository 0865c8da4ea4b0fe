// lmb200_detector.hpp — header-only C++ surface over the lmb200 C ABI with the names and semantics of
// cv::linemod::Detector / Modality / Feature / Template / Match (opencv_contrib rgbd/linemod.hpp), so the
// reference's HighLevelLineMOD (include/HighLevelLinemod.h:102 `cv::Ptr<cv::linemod::Detector> detector`,
// src/HighLevelLinemod.cpp:26-43,:93,:152,:256-320) can be pointed at it.  Errors that upstream raises as
// cv::Exception (CV_Assert) are thrown as lm::Error (std::runtime_error).
//
// Without OpenCV, images are lm::ImageView (non-owning).  Define LM_WITH_OPENCV before including this header
// to get cv::Mat overloads.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "lmb200.h"
#ifdef LM_WITH_OPENCV
#include <opencv2/core.hpp>
#endif

namespace lm {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error("lmb200 error " + std::to_string(c) + ": " + m), code(c) {}
};

struct Feature { int x, y, label; Feature() : x(0), y(0), label(0) {} Feature(int x_, int y_, int l_) : x(x_), y(y_), label(l_) {} };
struct Template { int width = 0, height = 0, pyramid_level = 0; std::vector<Feature> features; };
struct Rect { int x = 0, y = 0, width = 0, height = 0; };

struct Match {
  int x = 0, y = 0;
  float similarity = 0;
  std::string class_id;
  int template_id = 0;
  bool operator<(const Match& r) const { return similarity != r.similarity ? similarity > r.similarity : template_id < r.template_id; }
  bool operator==(const Match& r) const { return x == r.x && y == r.y && similarity == r.similarity && class_id == r.class_id; }
};

struct ImageView {
  const void* data = nullptr; int rows = 0, cols = 0, type = LMB200_8UC1; size_t step = 0;
  ImageView() {}
  ImageView(const void* d, int r, int c, int t, size_t s = 0) : data(d), rows(r), cols(c), type(t), step(s) {}
#ifdef LM_WITH_OPENCV
  ImageView(const cv::Mat& m) : data(m.data), rows(m.rows), cols(m.cols), type(m.type()), step(m.step) {}
#endif
  bool empty() const { return data == nullptr; }
  lmb200_image c() const { lmb200_image i; i.data = data; i.rows = rows; i.cols = cols; i.type = type; i.step = step; return i; }
};

// Modality descriptors (upstream: Modality::create("ColorGradient"|"DepthNormal"), ColorGradient(), DepthNormal()).
struct Modality {
  lmb200_modality m;
  std::string name() const { return m.type == LMB200_COLOR_GRADIENT ? "ColorGradient" : "DepthNormal"; }
  static Modality create(const std::string& type) {
    Modality r;
    if (type == "ColorGradient") lmb200_default_modality(LMB200_COLOR_GRADIENT, &r.m);
    else if (type == "DepthNormal") lmb200_default_modality(LMB200_DEPTH_NORMAL, &r.m);
    else throw Error(LMB200_E_INVALID, "unsupported modality type " + type);
    return r;
  }
};
inline Modality ColorGradient(float weak_threshold = 10.0f, size_t num_features = 63, float strong_threshold = 55.0f) {
  Modality r = Modality::create("ColorGradient");
  r.m.weak_threshold = weak_threshold; r.m.num_features = (int)num_features; r.m.strong_threshold = strong_threshold;
  return r;
}
inline Modality DepthNormal(int distance_threshold = 2000, int difference_threshold = 50, size_t num_features = 63, int extract_threshold = 2) {
  Modality r = Modality::create("DepthNormal");
  r.m.distance_threshold = distance_threshold; r.m.difference_threshold = difference_threshold;
  r.m.num_features = (int)num_features; r.m.extract_threshold = extract_threshold;
  return r;
}

class Detector {
 public:
  Detector(const std::vector<Modality>& modalities, const std::vector<int>& T_pyramid, int device = -1, int max_batch = 0) {
    lmb200_config cfg = lmb200_config();
    if (modalities.size() > LMB200_MAX_MODALITIES || T_pyramid.size() > LMB200_MAX_LEVELS) throw Error(LMB200_E_INVALID, "too many modalities/levels");
    cfg.num_modalities = (int)modalities.size();
    for (size_t i = 0; i < modalities.size(); ++i) cfg.modalities[i] = modalities[i].m;
    cfg.pyramid_levels = (int)T_pyramid.size();
    for (size_t i = 0; i < T_pyramid.size(); ++i) cfg.T[i] = T_pyramid[i];
    cfg.device = device; cfg.max_batch = max_batch;
    int rc = lmb200_create(&cfg, &h_);
    if (rc) throw Error(rc, lmb200_last_error(nullptr));
  }
  explicit Detector(lmb200_handle adopted) : h_(adopted) {}
  ~Detector() { if (h_) lmb200_destroy(h_); }
  Detector(const Detector&) = delete;
  Detector& operator=(const Detector&) = delete;

  // Detector::match(sources, threshold, matches, class_ids, quantized_images, masks)
  void match(const std::vector<ImageView>& sources, float threshold, std::vector<Match>& matches,
             const std::vector<std::string>& class_ids = std::vector<std::string>(),
             std::vector<std::vector<unsigned char>>* quantized_images = nullptr,
             const std::vector<ImageView>& masks = std::vector<ImageView>()) const {
    matches.clear();
    std::vector<lmb200_image> src, msk, qout;
    for (auto& s : sources) src.push_back(s.c());
    for (auto& m : masks) msk.push_back(m.c());
    if (!masks.empty() && masks.size() != sources.size()) throw Error(LMB200_E_SOURCES, "masks.size() != modalities.size()");
    std::vector<const char*> ids;
    for (auto& s : class_ids) ids.push_back(s.c_str());
    if (quantized_images) {
      int M = lmb200_num_modalities(h_), L = lmb200_pyramid_levels(h_);
      quantized_images->assign((size_t)M * L, std::vector<unsigned char>());
      int r = sources.empty() ? 0 : sources[0].rows, c = sources.empty() ? 0 : sources[0].cols;
      for (int l = 0; l < L; ++l) {
        for (int m = 0; m < M; ++m) {
          auto& q = (*quantized_images)[(size_t)l * M + m];
          q.assign((size_t)r * c, 0);
          lmb200_image qi; qi.data = q.data(); qi.rows = r; qi.cols = c; qi.type = LMB200_8UC1; qi.step = 0;
          qout.push_back(qi);
        }
        r /= 2; c /= 2;
      }
    }
    std::vector<lmb200_match_rec> rec(4096);
    size_t n = 0;
    for (;;) {
      int rc = lmb200_match(h_, src.data(), (int)src.size(), threshold, ids.empty() ? nullptr : ids.data(), (int)ids.size(),
                            rec.data(), rec.size(), &n, qout.empty() ? nullptr : qout.data(), msk.empty() ? nullptr : msk.data());
      if (rc == LMB200_E_TRUNCATED) { rec.resize(n); continue; }
      check(rc);
      break;
    }
    matches.resize(n);
    for (size_t i = 0; i < n; ++i) {
      matches[i].x = rec[i].x; matches[i].y = rec[i].y; matches[i].similarity = rec[i].similarity;
      matches[i].class_id = lmb200_class_id(h_, rec[i].class_index); matches[i].template_id = rec[i].template_id;
    }
  }

  // Detector::addTemplate — returns the template id or -1 (reference checks this: HighLevelLinemod.cpp:97)
  int addTemplate(const std::vector<ImageView>& sources, const std::string& class_id, const ImageView& object_mask, Rect* bounding_box = nullptr) {
    std::vector<lmb200_image> src;
    for (auto& s : sources) src.push_back(s.c());
    lmb200_image mk = object_mask.c();
    int bb[4] = {0, 0, 0, 0}, tid = -1;
    check(lmb200_add_template(h_, class_id.c_str(), src.data(), (int)src.size(), object_mask.empty() ? nullptr : &mk, bb, &tid));
    if (bounding_box && tid >= 0) { bounding_box->x = bb[0]; bounding_box->y = bb[1]; bounding_box->width = bb[2]; bounding_box->height = bb[3]; }
    return tid;
  }
  // n successive addTemplate calls in one batched GPU pass (views[i] = that view's sources; masks may be empty or hold
  // empty entries); returns the template ids (-1 = extraction failed), bounding boxes optional
  std::vector<int> addTemplates(const std::vector<std::vector<ImageView>>& views, const std::string& class_id,
                                const std::vector<ImageView>& object_masks = {}, std::vector<Rect>* bounding_boxes = nullptr) {
    std::vector<lmb200_image> src, mk;
    for (auto& v : views) for (auto& s : v) src.push_back(s.c());
    for (auto& m : object_masks) mk.push_back(m.c());
    if (!mk.empty() && mk.size() != views.size()) throw Error(LMB200_E_INVALID, "one mask per view expected");
    std::vector<int> tids(views.size(), -1), bb(views.size() * 4, 0);
    int per = views.empty() ? 0 : (int)views[0].size();
    check(lmb200_add_templates(h_, class_id.c_str(), (int)views.size(), src.data(), per, mk.empty() ? nullptr : mk.data(), bb.data(), tids.data()));
    if (bounding_boxes) {
      bounding_boxes->resize(views.size());
      for (size_t i = 0; i < views.size(); ++i) { Rect r; r.x = bb[4 * i]; r.y = bb[4 * i + 1]; r.width = bb[4 * i + 2]; r.height = bb[4 * i + 3]; (*bounding_boxes)[i] = r; }
    }
    return tids;
  }
  int addSyntheticTemplate(const std::vector<Template>& templates, const std::string& class_id) {
    std::vector<lmb200_template> t(templates.size());
    std::vector<std::vector<lmb200_feature>> f(templates.size());
    for (size_t i = 0; i < templates.size(); ++i) {
      for (auto& ft : templates[i].features) { lmb200_feature x; x.x = ft.x; x.y = ft.y; x.label = ft.label; f[i].push_back(x); }
      t[i].width = templates[i].width; t[i].height = templates[i].height; t[i].pyramid_level = templates[i].pyramid_level;
      t[i].num_features = (int)f[i].size(); t[i].features = f[i].data();
    }
    int tid = -1;
    check(lmb200_add_synthetic_template(h_, class_id.c_str(), t.data(), (int)t.size(), &tid));
    return tid;
  }

  std::vector<std::string> getModalities() const {
    std::vector<std::string> v;
    for (int i = 0; i < lmb200_num_modalities(h_); ++i) v.push_back(lmb200_modality_name(h_, i));
    return v;
  }
  int getT(int pyramid_level) const { return lmb200_get_T(h_, pyramid_level); }
  int pyramidLevels() const { return lmb200_pyramid_levels(h_); }
  std::vector<Template> getTemplates(const std::string& class_id, int template_id) const {
    int n = lmb200_num_modalities(h_) * lmb200_pyramid_levels(h_);
    std::vector<Template> out((size_t)n);
    for (int i = 0; i < n; ++i) {
      lmb200_template t;
      check(lmb200_get_template(h_, class_id.c_str(), template_id, i, &t));
      out[i].width = t.width; out[i].height = t.height; out[i].pyramid_level = t.pyramid_level;
      for (int k = 0; k < t.num_features; ++k) out[i].features.push_back(Feature(t.features[k].x, t.features[k].y, t.features[k].label));
    }
    return out;
  }
  int numTemplates() const { return lmb200_num_templates(h_, nullptr); }
  int numTemplates(const std::string& class_id) const { return lmb200_num_templates(h_, class_id.c_str()); }
  int numClasses() const { return lmb200_num_classes(h_); }
  std::vector<std::string> classIds() const {
    std::vector<std::string> v;
    for (int i = 0; i < lmb200_num_classes(h_); ++i) v.push_back(lmb200_class_id(h_, i));
    return v;
  }

  // Persistence in the reference's file layout (HighLevelLinemod.cpp:256-270 / :288-300)
  void write(const std::string& path) const { check(lmb200_write(h_, path.c_str())); }
  static std::unique_ptr<Detector> read(const std::string& path, int device = -1) {
    lmb200_handle h = nullptr;
    int rc = lmb200_read(path.c_str(), device, &h);
    if (rc) throw Error(rc, lmb200_last_error(nullptr));
    return std::unique_ptr<Detector>(new Detector(h));
  }
  // Detector::writeClass / readClass (upstream takes a FileStorage / FileNode; here the class has a file of its own)
  void writeClass(const std::string& class_id, const std::string& path) const { check(lmb200_write_class(h_, class_id.c_str(), path.c_str())); }
  void readClass(const std::string& path, const std::string& class_id_override = "") {
    check(lmb200_read_class(h_, path.c_str(), class_id_override.empty() ? nullptr : class_id_override.c_str()));
  }
  void writeClasses(const std::string& format = "templates_%s.yml.gz") const { check(lmb200_write_classes(h_, format.c_str())); }
  void readClasses(const std::vector<std::string>& class_ids, const std::string& format = "templates_%s.yml.gz") {
    std::vector<const char*> ids;
    for (auto& s : class_ids) ids.push_back(s.c_str());
    check(lmb200_read_classes(h_, ids.data(), (int)ids.size(), format.c_str()));
  }

  // ---- throughput / multi-GPU entry points (not in upstream; INTEGRATION.md sections 4 and 6) --------------------------
  // Device-resident steps: upload `n` frames (views[i] = the frame's sources) into slots [first_slot, first_slot + n),
  // match them (enqueue only; steps on different slot ranges overlap), fetch the per-frame lists later.
  void uploadFrames(const std::vector<std::vector<ImageView>>& frames, int first_slot = 0) {
    std::vector<lmb200_image> src;
    for (auto& f : frames) for (auto& s : f) src.push_back(s.c());
    check(lmb200_upload_frames(h_, src.data(), (int)frames.size(), frames.empty() ? 0 : (int)frames[0].size(), first_slot));
  }
  void matchResident(int first_slot, int count, float threshold, const std::vector<std::string>& class_ids = {}) {
    std::vector<const char*> ids;
    for (auto& s : class_ids) ids.push_back(s.c_str());
    check(lmb200_match_resident(h_, first_slot, count, threshold, ids.empty() ? nullptr : ids.data(), (int)ids.size()));
  }
  // per-frame match lists of slots [first_slot, first_slot + count); allgather = the collective fetch of a template-sharded step
  std::vector<std::vector<Match>> fetchResident(int first_slot, int count, bool allgather = false, size_t capacity = 0) {
    if (!capacity) capacity = (size_t)4096 * (size_t)count;
    std::vector<lmb200_match_rec> rec(capacity);
    std::vector<size_t> offs((size_t)count + 1);
    int rc = allgather ? lmb200_fetch_resident_allgather(h_, first_slot, count, rec.data(), capacity, offs.data())
                       : lmb200_fetch_resident(h_, first_slot, count, rec.data(), capacity, offs.data());
    check(rc);
    const std::vector<std::string> ids = classIds();
    std::vector<std::vector<Match>> out((size_t)count);
    for (int i = 0; i < count; ++i)
      for (size_t k = offs[i]; k < offs[i + 1]; ++k) {
        Match m;
        m.x = rec[k].x; m.y = rec[k].y; m.similarity = rec[k].similarity; m.template_id = rec[k].template_id;
        m.class_id = rec[k].class_index >= 0 && rec[k].class_index < (int)ids.size() ? ids[rec[k].class_index] : std::string();
        out[i].push_back(m);
      }
    return out;
  }
  // One process per GPU: this handle scores shard `rank` of `world` (interleaved over the selection list) ...
  void setTemplateShard(int rank, int world) { check(lmb200_set_template_shard(h_, rank, world)); }
  // ... over an NCCL communicator (unique id from commUniqueId() on one rank, shipped to the others by the application)
  static std::vector<uint8_t> commUniqueId() {
    std::vector<uint8_t> id(128);
    int rc = lmb200_comm_unique_id(id.data());
    if (rc) throw Error(rc, lmb200_last_error(nullptr));
    return id;
  }
  void commInit(const std::vector<uint8_t>& unique_id128, int rank, int world) {
    if (unique_id128.size() != 128) throw Error(LMB200_E_INVALID, "the NCCL unique id is 128 bytes");
    check(lmb200_comm_init(h_, unique_id128.data(), rank, world));
  }
  // template-sharded step: quantisers sharded by frame block + all-gather of the maps, spread, match, match all-gather and
  // the sort/unique epilogue on the device; follow with fetchResident(first_slot, count, true)
  void matchResidentSharded(int first_slot, int count, float threshold, const std::vector<std::string>& class_ids = {}) {
    std::vector<const char*> ids;
    for (auto& s : class_ids) ids.push_back(s.c_str());
    check(lmb200_match_resident_sharded(h_, first_slot, count, threshold, ids.empty() ? nullptr : ids.data(), (int)ids.size()));
  }
  void setOption(const std::string& name, int value) { check(lmb200_set_option(h_, name.c_str(), value)); }

  lmb200_handle handle() const { return h_; }

 private:
  void check(int rc) const { if (rc) throw Error(rc, lmb200_last_error(h_)); }
  lmb200_handle h_ = nullptr;
};

// cv::linemod::getDefaultLINE / getDefaultLINEMOD
inline std::unique_ptr<Detector> getDefaultLINE() { return std::unique_ptr<Detector>(new Detector({ColorGradient()}, {5, 8})); }
inline std::unique_ptr<Detector> getDefaultLINEMOD() { return std::unique_ptr<Detector>(new Detector({ColorGradient(), DepthNormal()}, {5, 8})); }

// Headless stand-in for the reference's OpenGLRender (src/OpenglRender.cpp): model + pinhole camera -> depth (u16, mm)
// and colour (BGR8, white on black) images of n views, rendered on the host threads.
struct Mesh {
  std::vector<double> vertices;  // xyz per vertex (mm)
  std::vector<int> triangles;    // three indices per triangle
  static Mesh loadPly(const std::string& path) {
    double* v = nullptr; int* t = nullptr; int nv = 0, nt = 0;
    int rc = lmb200_load_ply(path.c_str(), &v, &nv, &t, &nt);
    if (rc) throw Error(rc, "cannot load " + path);
    Mesh m;
    m.vertices.assign(v, v + 3 * (size_t)nv);
    m.triangles.assign(t, t + 3 * (size_t)nt);
    lmb200_free(v); lmb200_free(t);
    return m;
  }
};
struct RenderedViews {
  int n = 0, width = 0, height = 0;
  std::vector<uint16_t> depth;   // [n][height][width]
  std::vector<uint8_t> colour;   // [n][height][width][3]
};
inline lmb200_camera referenceCamera(int width = 640, int height = 480, double f = 1045.69141) {
  lmb200_camera c; c.width = width; c.height = height; c.fx = f; c.fy = f; c.cx = width / 2.0; c.cy = height / 2.0; c.near_mm = 100.0; c.far_mm = 10000.0;
  return c;
}
// OpenGLRender::renderDepthToFrontBuff / renderColorToFrontBuff(model, camPosition) for many camera positions (xyz each)
inline RenderedViews renderLookAt(const Mesh& mesh, const lmb200_camera& cam, const std::vector<double>& eyes_xyz, int threads = 0) {
  RenderedViews out;
  out.n = (int)(eyes_xyz.size() / 3); out.width = cam.width; out.height = cam.height;
  out.depth.resize((size_t)out.n * cam.width * cam.height);
  out.colour.resize(out.depth.size() * 3);
  lmb200_mesh m; m.vertices = mesh.vertices.data(); m.n_vertices = (int)(mesh.vertices.size() / 3);
  m.triangles = mesh.triangles.data(); m.n_triangles = (int)(mesh.triangles.size() / 3);
  int rc = lmb200_render_lookat(&m, &cam, eyes_xyz.data(), out.n, out.depth.data(), out.colour.data(), threads);
  if (rc) throw Error(rc, "lmb200_render_lookat");
  return out;
}

}  // namespace lm
