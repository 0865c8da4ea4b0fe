// lmb200_opencv.hpp — DROP-IN replacement for <opencv2/rgbd/linemod.hpp>: declares namespace cv::linemod with the
// upstream names, signatures and semantics (opencv_contrib modules/rgbd/include/opencv2/rgbd/linemod.hpp), implemented
// over the lmb200 C ABI.  The reference's HighLevelLineMOD compiles against it unchanged:
//   include/HighLevelLinemod.h:102      cv::Ptr<cv::linemod::Detector> detector
//   src/HighLevelLinemod.cpp:26-43      cv::makePtr<cv::linemod::ColorGradient>() / DepthNormal() / Detector(modality, T)
//   src/HighLevelLinemod.cpp:55-65      detector->classIds() / numClasses() / numTemplates()
//   src/HighLevelLinemod.cpp:93         detector->addTemplate(templateImgs, in_modelName, maskRotated, &boundingBox)
//   src/HighLevelLinemod.cpp:115-126    detector->getTemplates(class_id, template_id), getModalities().size(), Feature x/y
//   src/HighLevelLinemod.cpp:152        detector->match(in_imgs, detectorThreshold, matches, currentClass)
//   src/HighLevelLinemod.cpp:260-267    detector->write(fs); detector->writeClass(id, fs)
//   src/HighLevelLinemod.cpp:294-299    detector->read(fs.root()); detector->readClass(i)
// Integration = replace `#include <opencv2/rgbd.hpp>` (include/HighLevelLinemod.h:7, include/utility.h:12) by this header
// and link liblmb200.so.  Needs only <opencv2/core.hpp> (Mat, Ptr, Rect, FileStorage, FileNode).
// Upstream CV_Assert failures surface as lm::Error (a std::runtime_error), which callers that catch cv::Exception by
// std::exception& still catch.
#pragma once
#include <opencv2/core.hpp>

#include <map>
#include <string>
#include <utility>
#include <vector>

#ifndef LM_WITH_OPENCV
#define LM_WITH_OPENCV 1
#endif
#include "lmb200_detector.hpp"

namespace cv {
namespace linemod {

// ---------------------------------------------------------------------------------------------- data model (a1)
struct Feature {
  int x, y, label;
  Feature() : x(0), y(0), label(0) {}
  Feature(int x_, int y_, int label_) : x(x_), y(y_), label(label_) {}
  void read(const FileNode& fn) { FileNodeIterator it = fn.begin(); x = (int)*it; ++it; y = (int)*it; ++it; label = (int)*it; }
  void write(FileStorage& fs) const { fs << "[:" << x << y << label << "]"; }
};

struct Template {
  int width, height, pyramid_level;
  std::vector<Feature> features;
  Template() : width(0), height(0), pyramid_level(0) {}
  void read(const FileNode& fn) {
    width = (int)fn["width"]; height = (int)fn["height"]; pyramid_level = (int)fn["pyramid_level"];
    FileNode f = fn["features"];
    features.clear();
    for (FileNodeIterator it = f.begin(); it != f.end(); ++it) { Feature ft; ft.read(*it); features.push_back(ft); }
  }
  void write(FileStorage& fs) const {
    fs << "width" << width; fs << "height" << height; fs << "pyramid_level" << pyramid_level;
    fs << "features" << "[";
    for (size_t i = 0; i < features.size(); ++i) features[i].write(fs);
    fs << "]";
  }
};

struct Match {
  Match() : x(0), y(0), similarity(0), template_id(0) {}
  Match(int x_, int y_, float s, const String& c, int t) : x(x_), y(y_), similarity(s), class_id(c), template_id(t) {}
  bool operator<(const Match& rhs) const { return similarity != rhs.similarity ? similarity > rhs.similarity : template_id < rhs.template_id; }
  bool operator==(const Match& rhs) const { return x == rhs.x && y == rhs.y && similarity == rhs.similarity && class_id == rhs.class_id; }
  int x, y;
  float similarity;
  String class_id;
  int template_id;
};

// ---------------------------------------------------------------------------------------------- modalities
class Modality {
 public:
  virtual ~Modality() {}
  virtual String name() const = 0;
  virtual void read(const FileNode& fn) = 0;
  virtual void write(FileStorage& fs) const = 0;
  virtual lmb200_modality params() const = 0;   // the C-ABI descriptor of this modality
  static Ptr<Modality> create(const String& modality_type);
  static Ptr<Modality> create(const FileNode& fn);
};

class ColorGradient : public Modality {
 public:
  ColorGradient() : weak_threshold(10.0f), num_features(63), strong_threshold(55.0f) {}
  ColorGradient(float weak, size_t nf, float strong) : weak_threshold(weak), num_features(nf), strong_threshold(strong) {}
  static Ptr<ColorGradient> create(float weak, size_t nf, float strong) { return makePtr<ColorGradient>(weak, nf, strong); }
  String name() const override { return "ColorGradient"; }
  void read(const FileNode& fn) override {
    weak_threshold = (float)fn["weak_threshold"]; num_features = (size_t)(int)fn["num_features"]; strong_threshold = (float)fn["strong_threshold"];
  }
  void write(FileStorage& fs) const override {
    fs << "type" << "ColorGradient"; fs << "weak_threshold" << weak_threshold; fs << "num_features" << (int)num_features;
    fs << "strong_threshold" << strong_threshold;
  }
  lmb200_modality params() const override {
    lmb200_modality m; lmb200_default_modality(LMB200_COLOR_GRADIENT, &m);
    m.weak_threshold = weak_threshold; m.num_features = (int)num_features; m.strong_threshold = strong_threshold;
    return m;
  }
  float weak_threshold; size_t num_features; float strong_threshold;
};

class DepthNormal : public Modality {
 public:
  DepthNormal() : distance_threshold(2000), difference_threshold(50), num_features(63), extract_threshold(2) {}
  DepthNormal(int dist, int diff, size_t nf, int ext) : distance_threshold(dist), difference_threshold(diff), num_features(nf), extract_threshold(ext) {}
  static Ptr<DepthNormal> create(int dist, int diff, size_t nf, int ext) { return makePtr<DepthNormal>(dist, diff, nf, ext); }
  String name() const override { return "DepthNormal"; }
  void read(const FileNode& fn) override {
    distance_threshold = (int)fn["distance_threshold"]; difference_threshold = (int)fn["difference_threshold"];
    num_features = (size_t)(int)fn["num_features"]; extract_threshold = (int)fn["extract_threshold"];
  }
  void write(FileStorage& fs) const override {
    fs << "type" << "DepthNormal"; fs << "distance_threshold" << distance_threshold; fs << "difference_threshold" << difference_threshold;
    fs << "num_features" << (int)num_features; fs << "extract_threshold" << extract_threshold;
  }
  lmb200_modality params() const override {
    lmb200_modality m; lmb200_default_modality(LMB200_DEPTH_NORMAL, &m);
    m.distance_threshold = distance_threshold; m.difference_threshold = difference_threshold; m.num_features = (int)num_features;
    m.extract_threshold = extract_threshold;
    return m;
  }
  int distance_threshold, difference_threshold; size_t num_features; int extract_threshold;
};

inline Ptr<Modality> Modality::create(const String& modality_type) {
  if (modality_type == "ColorGradient") return makePtr<ColorGradient>();
  if (modality_type == "DepthNormal") return makePtr<DepthNormal>();
  throw lm::Error(LMB200_E_INVALID, "Unsupported modality type " + modality_type);
}
inline Ptr<Modality> Modality::create(const FileNode& fn) {
  Ptr<Modality> m = create((String)fn["type"]);
  m->read(fn);
  return m;
}

// ---------------------------------------------------------------------------------------------- Detector
class Detector {
 public:
  Detector() {}
  Detector(const std::vector<Ptr<Modality>>& modalities_, const std::vector<int>& T_pyramid) : modalities(modalities_), T_at_level(T_pyramid) { create(); }
  ~Detector() { if (h_) lmb200_destroy(h_); }
  Detector(const Detector&) = delete;
  Detector& operator=(const Detector&) = delete;

  // upstream: match(sources, threshold, matches, class_ids = {}, quantized_images = noArray(), masks = {})
  void match(const std::vector<Mat>& sources, float threshold, std::vector<Match>& matches,
             const std::vector<String>& class_ids = std::vector<String>()) const {
    do_match(sources, threshold, matches, class_ids, nullptr, std::vector<Mat>());
  }
  void match(const std::vector<Mat>& sources, float threshold, std::vector<Match>& matches, const std::vector<String>& class_ids,
             std::vector<Mat>& quantized_images, const std::vector<Mat>& masks = std::vector<Mat>()) const {
    do_match(sources, threshold, matches, class_ids, &quantized_images, masks);
  }

  int addTemplate(const std::vector<Mat>& sources, const String& class_id, const Mat& object_mask, Rect* bounding_box = NULL) {
    need();
    std::vector<lmb200_image> src;
    for (size_t i = 0; i < sources.size(); ++i) src.push_back(lm::ImageView(sources[i]).c());
    lmb200_image mk = lm::ImageView(object_mask).c();
    int bb[4] = {0, 0, 0, 0}, tid = -1;
    check(lmb200_add_template(h_, class_id.c_str(), src.data(), (int)src.size(), object_mask.empty() ? nullptr : &mk, bb, &tid));
    cache_.clear();
    if (bounding_box && tid >= 0) *bounding_box = Rect(bb[0], bb[1], bb[2], bb[3]);
    return tid;
  }
  int addSyntheticTemplate(const std::vector<Template>& templates, const String& class_id) {
    need();
    std::vector<lmb200_template> t(templates.size());
    std::vector<std::vector<lmb200_feature>> f(templates.size());
    for (size_t i = 0; i < templates.size(); ++i) {
      for (size_t k = 0; k < templates[i].features.size(); ++k) {
        lmb200_feature x; x.x = templates[i].features[k].x; x.y = templates[i].features[k].y; x.label = templates[i].features[k].label;
        f[i].push_back(x);
      }
      t[i].width = templates[i].width; t[i].height = templates[i].height; t[i].pyramid_level = templates[i].pyramid_level;
      t[i].num_features = (int)f[i].size(); t[i].features = f[i].data();
    }
    int tid = -1;
    check(lmb200_add_synthetic_template(h_, class_id.c_str(), t.data(), (int)t.size(), &tid));
    cache_.clear();
    return tid;
  }

  const std::vector<Ptr<Modality>>& getModalities() const { return modalities; }
  int getT(int pyramid_level) const { return T_at_level[(size_t)pyramid_level]; }
  int pyramidLevels() const { return (int)T_at_level.size(); }
  // upstream returns a reference into its own store; the copy fetched through the ABI is kept until the set changes
  const std::vector<Template>& getTemplates(const String& class_id, int template_id) const {
    need();
    std::pair<String, int> key(class_id, template_id);
    std::map<std::pair<String, int>, std::vector<Template>>::iterator it = cache_.find(key);
    if (it != cache_.end()) return it->second;
    const int n = lmb200_num_modalities(h_) * lmb200_pyramid_levels(h_);
    std::vector<Template> out((size_t)n);
    for (int i = 0; i < n; ++i) {
      lmb200_template t;
      check(lmb200_get_template(h_, class_id.c_str(), template_id, i, &t));
      out[(size_t)i].width = t.width; out[(size_t)i].height = t.height; out[(size_t)i].pyramid_level = t.pyramid_level;
      for (int k = 0; k < t.num_features; ++k) out[(size_t)i].features.push_back(Feature(t.features[k].x, t.features[k].y, t.features[k].label));
    }
    return cache_[key] = out;
  }
  int numTemplates() const { return h_ ? lmb200_num_templates(h_, nullptr) : 0; }
  int numTemplates(const String& class_id) const { return h_ ? lmb200_num_templates(h_, class_id.c_str()) : 0; }
  int numClasses() const { return h_ ? lmb200_num_classes(h_) : 0; }
  std::vector<String> classIds() const {
    std::vector<String> v;
    for (int i = 0; i < numClasses(); ++i) v.push_back(lmb200_class_id(h_, i));
    return v;
  }

  // ---- persistence, upstream's node layout (SURVEY.md 8c "Template file layout")
  void read(const FileNode& fn) {
    modalities.clear(); T_at_level.clear();
    const int levels = (int)fn["pyramid_levels"];
    fn["T"] >> T_at_level;
    if ((int)T_at_level.size() != levels) throw lm::Error(LMB200_E_IO, "pyramid_levels does not match T");
    FileNode m = fn["modalities"];
    for (FileNodeIterator it = m.begin(); it != m.end(); ++it) modalities.push_back(Modality::create(*it));
    create();
  }
  void write(FileStorage& fs) const {
    fs << "pyramid_levels" << pyramidLevels();
    fs << "T" << T_at_level;
    fs << "modalities" << "[";
    for (size_t i = 0; i < modalities.size(); ++i) { fs << "{"; modalities[i]->write(fs); fs << "}"; }
    fs << "]";
  }
  String readClass(const FileNode& fn, const String& class_id_override = "") {
    need();
    // upstream: CV_Assert on modality names and pyramid_levels, then on "class not already present"
    FileNode mn = fn["modalities"];
    if (mn.size() != modalities.size()) throw lm::Error(LMB200_E_CLASS, "readClass: modality count differs from the detector's");
    size_t i = 0;
    for (FileNodeIterator it = mn.begin(); it != mn.end(); ++it, ++i)
      if ((String)*it != modalities[i]->name()) throw lm::Error(LMB200_E_CLASS, "readClass: modality names differ from the detector's");
    if ((int)fn["pyramid_levels"] != pyramidLevels()) throw lm::Error(LMB200_E_CLASS, "readClass: pyramid_levels differs from the detector's");
    String class_id = class_id_override.empty() ? (String)fn["class_id"] : class_id_override;
    if (class_id_override.empty() && numTemplates(class_id) > 0) throw lm::Error(LMB200_E_CLASS, "readClass: class already present");
    if (!class_id_override.empty() && numTemplates(class_id) > 0) return class_id;   // std::map::insert semantics: the existing entry wins
    FileNode tps = fn["template_pyramids"];
    int expected = 0;
    for (FileNodeIterator it = tps.begin(); it != tps.end(); ++it, ++expected) {
      if ((int)(*it)["template_id"] != expected) throw lm::Error(LMB200_E_IO, "readClass: template_id is not consecutive");
      FileNode tn = (*it)["templates"];
      std::vector<Template> tp;
      for (FileNodeIterator jt = tn.begin(); jt != tn.end(); ++jt) { Template t; t.read(*jt); tp.push_back(t); }
      addSyntheticTemplate(tp, class_id);
    }
    return class_id;
  }
  void writeClass(const String& class_id, FileStorage& fs) const {
    need();
    const int n = numTemplates(class_id);
    if (n <= 0 && !has_class(class_id)) throw lm::Error(LMB200_E_CLASS, "writeClass: unknown class " + class_id);
    fs << "class_id" << class_id;
    fs << "modalities" << "[:";
    for (size_t i = 0; i < modalities.size(); ++i) fs << modalities[i]->name();
    fs << "]";
    fs << "pyramid_levels" << pyramidLevels();
    fs << "template_pyramids" << "[";
    for (int t = 0; t < n; ++t) {
      const std::vector<Template>& tp = getTemplates(class_id, t);
      fs << "{";
      fs << "template_id" << t;
      fs << "templates" << "[";
      for (size_t j = 0; j < tp.size(); ++j) { fs << "{"; tp[j].write(fs); fs << "}"; }
      fs << "]";
      fs << "}";
    }
    fs << "]";
  }
  void readClasses(const std::vector<String>& class_ids, const String& fmt = "templates_%s.yml.gz") {
    for (size_t i = 0; i < class_ids.size(); ++i) {
      FileStorage fs(format(fmt.c_str(), class_ids[i].c_str()), FileStorage::READ);
      if (!fs.isOpened()) throw lm::Error(LMB200_E_IO, "readClasses: cannot open the file of class " + class_ids[i]);
      readClass(fs.root());
    }
  }
  void writeClasses(const String& fmt = "templates_%s.yml.gz") const {
    const std::vector<String> ids = classIds();
    for (size_t i = 0; i < ids.size(); ++i) {
      FileStorage fs(format(fmt.c_str(), ids[i].c_str()), FileStorage::WRITE);
      writeClass(ids[i], fs);
    }
  }

  lmb200_handle handle() const { return h_; }   // the C ABI underneath (batch / resident / multi-GPU entry points)

 protected:
  std::vector<Ptr<Modality>> modalities;
  std::vector<int> T_at_level;

 private:
  void create() {
    if (h_) { lmb200_destroy(h_); h_ = nullptr; }
    cache_.clear();
    lmb200_config cfg;
    lmb200_default_config(&cfg, 0);
    if (modalities.size() > LMB200_MAX_MODALITIES || T_at_level.size() > LMB200_MAX_LEVELS) throw lm::Error(LMB200_E_INVALID, "too many modalities/levels");
    cfg.num_modalities = (int)modalities.size();
    for (size_t i = 0; i < modalities.size(); ++i) cfg.modalities[i] = modalities[i]->params();
    cfg.pyramid_levels = (int)T_at_level.size();
    for (size_t i = 0; i < T_at_level.size(); ++i) cfg.T[i] = T_at_level[i];
    int rc = lmb200_create(&cfg, &h_);
    if (rc) throw lm::Error(rc, lmb200_last_error(nullptr));
  }
  void need() const { if (!h_) throw lm::Error(LMB200_E_INVALID, "empty Detector: construct it with modalities or read() it first"); }
  void check(int rc) const { if (rc) throw lm::Error(rc, lmb200_last_error(h_)); }
  bool has_class(const String& id) const { const std::vector<String> ids = classIds(); for (size_t i = 0; i < ids.size(); ++i) if (ids[i] == id) return true; return false; }
  void do_match(const std::vector<Mat>& sources, float threshold, std::vector<Match>& matches, const std::vector<String>& class_ids,
                std::vector<Mat>* quantized_images, const std::vector<Mat>& masks) const {
    need();
    matches.clear();
    std::vector<lmb200_image> src, msk, qout;
    for (size_t i = 0; i < sources.size(); ++i) src.push_back(lm::ImageView(sources[i]).c());
    if (!masks.empty() && masks.size() != sources.size()) throw lm::Error(LMB200_E_SOURCES, "masks.size() != modalities.size()");
    for (size_t i = 0; i < masks.size(); ++i) msk.push_back(lm::ImageView(masks[i]).c());
    std::vector<const char*> ids;
    for (size_t i = 0; i < class_ids.size(); ++i) ids.push_back(class_ids[i].c_str());
    if (quantized_images) {
      const int M = (int)modalities.size(), L = pyramidLevels();
      quantized_images->resize((size_t)M * L);
      int r = sources.empty() ? 0 : sources[0].rows, c = sources.empty() ? 0 : sources[0].cols;
      for (int l = 0; l < L; ++l) {
        for (int m = 0; m < M; ++m) {
          Mat& q = (*quantized_images)[(size_t)l * M + m];
          q.create(r, c, CV_8UC1);
          qout.push_back(lm::ImageView(q).c());
        }
        r /= 2; c /= 2;
      }
    }
    std::vector<lmb200_match_rec> rec(4096);
    size_t n = 0;
    for (;;) {
      int rc = lmb200_match(h_, src.data(), (int)src.size(), threshold, ids.empty() ? nullptr : ids.data(), (int)ids.size(), rec.data(),
                            rec.size(), &n, qout.empty() ? nullptr : qout.data(), msk.empty() ? nullptr : msk.data());
      if (rc == LMB200_E_TRUNCATED) { rec.resize(n); continue; }
      check(rc);
      break;
    }
    matches.reserve(n);
    for (size_t i = 0; i < n; ++i)
      matches.push_back(Match(rec[i].x, rec[i].y, rec[i].similarity, lmb200_class_id(h_, rec[i].class_index), rec[i].template_id));
  }

  lmb200_handle h_ = nullptr;
  mutable std::map<std::pair<String, int>, std::vector<Template>> cache_;
};

inline Ptr<Detector> getDefaultLINE() {
  std::vector<Ptr<Modality>> m;
  m.push_back(makePtr<ColorGradient>());
  static const int T_DEFAULTS[] = {5, 8};
  return makePtr<Detector>(m, std::vector<int>(T_DEFAULTS, T_DEFAULTS + 2));
}
inline Ptr<Detector> getDefaultLINEMOD() {
  std::vector<Ptr<Modality>> m;
  m.push_back(makePtr<ColorGradient>());
  m.push_back(makePtr<DepthNormal>());
  static const int T_DEFAULTS[] = {5, 8};
  return makePtr<Detector>(m, std::vector<int>(T_DEFAULTS, T_DEFAULTS + 2));
}

}  // namespace linemod
}  // namespace cv
