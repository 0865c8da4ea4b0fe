// tests/cpp_dropin_check.cpp — the reference's cv::linemod call expressions, VERBATIM, compiled against the drop-in header
// include/lmb200_opencv.hpp (instead of <opencv2/rgbd.hpp>) and run through the C ABI.
// Every block marked [ref file:lines] is copied character for character from /root/reference/src/HighLevelLinemod.cpp
// (only the surrounding scaffolding — this cut-down class — is new).  <opencv2/core.hpp> resolves to tests/cv_stub here
// because OpenCV's C++ headers are not installed in this image; against a real OpenCV the same header is used unchanged.
#include <unistd.h>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>

#include "lmb200_opencv.hpp"   // the ONE line that changes in include/HighLevelLinemod.h:7 (was: #include <opencv2/rgbd.hpp>)

class HighLevelLineMOD {
 public:
  explicit HighLevelLineMOD(bool in_onlyColorModality);
  ~HighLevelLineMOD();
  std::vector<cv::String> getClassIds();
  uint16_t getNumClasses();
  uint32_t getNumTemplates();
  bool addTemplateOnce(std::vector<cv::Mat>& templateImgs, const std::string& in_modelName, cv::Mat maskRotated, cv::Rect& boundingBoxOut);
  void templatePoints(cv::linemod::Match const& in_match, std::vector<cv::Point>& points);
  bool detectTemplate(std::vector<cv::Mat>& in_imgs, uint16_t in_classNumber);
  void writeLinemod();
  void readLinemod();

  cv::Ptr<cv::linemod::Detector> detector;            // [ref include/HighLevelLinemod.h:102]
  std::vector<cv::linemod::Match> matches;            // [ref include/HighLevelLinemod.h:166]
  bool onlyColorModality;
  float detectorThreshold = 80.f;
};

HighLevelLineMOD::HighLevelLineMOD(bool in_onlyColorModality) : onlyColorModality(in_onlyColorModality) {
  // [ref src/HighLevelLinemod.cpp:26-43]
	if (!onlyColorModality)
	{
		std::vector<cv::Ptr<cv::linemod::Modality>> modality;
		modality.emplace_back(cv::makePtr<cv::linemod::ColorGradient>());
		modality.emplace_back(cv::makePtr<cv::linemod::DepthNormal>());

		static const int T_DEFAULTS[] = {5, 8};
		detector = cv::makePtr<cv::linemod::Detector>(
			modality, std::vector<int>(T_DEFAULTS, T_DEFAULTS + 2));
	}
	else
	{
		std::vector<cv::Ptr<cv::linemod::Modality>> modality;
		modality.emplace_back(cv::makePtr<cv::linemod::ColorGradient>());
		static const int T_DEFAULTS[] = {2, 8};
		detector = cv::makePtr<cv::linemod::Detector>(
			modality, std::vector<int>(T_DEFAULTS, T_DEFAULTS + 2));
	}
}

HighLevelLineMOD::~HighLevelLineMOD()
{
	detector.release();                                 // [ref :50]
}

std::vector<cv::String> HighLevelLineMOD::getClassIds()
{
	return detector->classIds();                        // [ref :55]
}

uint16_t HighLevelLineMOD::getNumClasses()
{
	return detector->numClasses();                      // [ref :60]
}

uint32_t HighLevelLineMOD::getNumTemplates()
{
	return detector->numTemplates();                    // [ref :65]
}

bool HighLevelLineMOD::addTemplateOnce(std::vector<cv::Mat>& templateImgs, const std::string& in_modelName, cv::Mat maskRotated, cv::Rect& boundingBoxOut) {
  // [ref src/HighLevelLinemod.cpp:92-101]
		cv::Rect boundingBox;
		uint64_t template_id = detector->addTemplate(templateImgs, in_modelName, maskRotated,
		                                             &boundingBox);
		templateImgs.clear();

		if (template_id == -1)
		{
			std::cout << "ERROR::Cant create Template" << std::endl;
			return false;
		}
  boundingBoxOut = boundingBox;
  return true;
}

void HighLevelLineMOD::templatePoints(cv::linemod::Match const& in_match, std::vector<cv::Point>& points) {
  // [ref src/HighLevelLinemod.cpp:115-126]
	const std::vector<cv::linemod::Template>& templates = detector->getTemplates(
		in_match.class_id, in_match.template_id);
	cv::Point offset(in_match.x, in_match.y);
	uint16_t num_modalities = detector->getModalities().size();
	for (int m = 0; m < num_modalities; ++m)
	{
		for (cv::linemod::Feature f : templates[m].features)
		{
			points.push_back(cv::Point(f.x, f.y) + offset);
		}
	}
}

bool HighLevelLineMOD::detectTemplate(std::vector<cv::Mat>& in_imgs, uint16_t in_classNumber) {
  // [ref src/HighLevelLinemod.cpp:142-156]
	cv::Mat tmpDepth;
	bool depthCheckForColorDetector = false;

	const std::vector<std::string> currentClass(1, detector->classIds()[in_classNumber]);
	if (onlyColorModality && in_imgs.size() == 2)
	{
		tmpDepth = in_imgs[1];
		in_imgs.pop_back();
		depthCheckForColorDetector = true;
	}
	detector->match(in_imgs, detectorThreshold, matches, currentClass);
	if (depthCheckForColorDetector)
	{
		in_imgs.push_back(tmpDepth);
	}
  return !matches.empty();
}

void HighLevelLineMOD::writeLinemod()
{
  // [ref src/HighLevelLinemod.cpp:258-270]
	std::string filename = "linemod_templates.yml.gz";
	cv::FileStorage fs(filename, cv::FileStorage::WRITE);
	detector->write(fs);

	std::vector<cv::String> ids = detector->classIds();
	fs << "classes" << "[";
	for (const auto& id : ids)
	{
		fs << "{";
		detector->writeClass(id, fs);
		fs << "}";
	}
	fs << "]";
}

void HighLevelLineMOD::readLinemod()
{
  // [ref src/HighLevelLinemod.cpp:292-300]
	std::string filename = "linemod_templates.yml.gz";
	cv::FileStorage fs(filename, cv::FileStorage::READ);
	detector->read(fs.root());

	cv::FileNode fn = fs["classes"];
	for (auto&& i : fn)
	{
		detector->readClass(i);
	}
}

// ---------------------------------------------------------------------------------------------- scaffolding
static std::vector<cv::linemod::Template> some_pyramid(int seed, int n_mod) {
  std::vector<cv::linemod::Template> tp((size_t)2 * n_mod);
  for (int i = 0; i < 2 * n_mod; ++i) {
    const int level = i / n_mod;
    tp[i].width = 120 >> level; tp[i].height = 90 >> level; tp[i].pyramid_level = level;
    for (int k = 0; k < (63 >> level); ++k)
      tp[i].features.push_back(cv::linemod::Feature((7 * k + seed) % (tp[i].width + 1), (11 * k + 3 * seed) % (tp[i].height + 1), (k + seed + i) % 8));
  }
  return tp;
}

static bool same_templates(const cv::linemod::Detector& a, const cv::linemod::Detector& b) {
  if (a.classIds() != b.classIds() || a.numTemplates() != b.numTemplates() || a.pyramidLevels() != b.pyramidLevels()) return false;
  if (a.getModalities().size() != b.getModalities().size()) return false;
  for (size_t m = 0; m < a.getModalities().size(); ++m) if (a.getModalities()[m]->name() != b.getModalities()[m]->name()) return false;
  for (int l = 0; l < a.pyramidLevels(); ++l) if (a.getT(l) != b.getT(l)) return false;
  for (const cv::String& id : a.classIds())
    for (int t = 0; t < a.numTemplates(id); ++t) {
      const std::vector<cv::linemod::Template>& x = a.getTemplates(id, t);
      const std::vector<cv::linemod::Template>& y = b.getTemplates(id, t);
      if (x.size() != y.size()) return false;
      for (size_t i = 0; i < x.size(); ++i) {
        if (x[i].width != y[i].width || x[i].height != y[i].height || x[i].pyramid_level != y[i].pyramid_level || x[i].features.size() != y[i].features.size()) return false;
        for (size_t k = 0; k < x[i].features.size(); ++k)
          if (x[i].features[k].x != y[i].features[k].x || x[i].features[k].y != y[i].features[k].y || x[i].features[k].label != y[i].features[k].label) return false;
      }
    }
  return true;
}

int main(int argc, char** argv) {
  if (argc > 1 && chdir(argv[1]) != 0) return 90;        // the reference writes into the working directory
  for (int only_color = 0; only_color < 2; ++only_color) {
    HighLevelLineMOD line(only_color != 0);
    const int n_mod = only_color ? 1 : 2;
    if ((int)line.detector->getModalities().size() != n_mod || line.detector->getT(0) != (only_color ? 2 : 5)) return 1;
    for (int k = 0; k < 5; ++k) line.detector->addSyntheticTemplate(some_pyramid(k, n_mod), k % 2 ? "lagergehaeuse.ply" : "other obj");
    if (line.getNumClasses() != 2 || line.getNumTemplates() != 5 || line.getClassIds()[0] != "lagergehaeuse.ply") return 2;
    line.writeLinemod();
    // (1) the reference's own reader on what its own writer produced
    HighLevelLineMOD back(only_color != 0);
    back.readLinemod();
    if (!same_templates(*line.detector, *back.detector)) return 3;
    // (2) the file on disk is the on-disk contract: the product's C-ABI reader must load it to the same detector
    lmb200_handle h = nullptr;
    if (lmb200_read("linemod_templates.yml.gz", -1, &h) != LMB200_OK) { std::printf("lmb200_read: %s\n", lmb200_last_error(nullptr)); return 4; }
    if (lmb200_num_templates(h, nullptr) != 5 || lmb200_num_classes(h) != 2 || lmb200_get_T(h, 0) != (only_color ? 2 : 5)) return 5;
    lmb200_destroy(h);
    // (3) readClass refuses a class that is already present, like upstream's CV_Assert
    try { back.readLinemod(); /* read() recreates the detector, so this succeeds */ } catch (const std::exception&) { return 6; }
    try {
      cv::FileStorage fs("linemod_templates.yml.gz", cv::FileStorage::READ);
      cv::FileNode fn = fs["classes"];
      for (auto&& i : fn) back.detector->readClass(i);
      return 7;
    } catch (const lm::Error& e) { if (e.code != LMB200_E_CLASS) return 8; }
    // (4) writeClasses / readClasses through FileStorage with the default "templates_%s.yml.gz" pattern
    line.detector->writeClasses();
    std::vector<cv::Ptr<cv::linemod::Modality>> mods = line.detector->getModalities();
    std::vector<int> Ts; for (int l = 0; l < line.detector->pyramidLevels(); ++l) Ts.push_back(line.detector->getT(l));
    cv::linemod::Detector fresh(mods, Ts);
    fresh.readClasses(line.detector->classIds());
    if (!same_templates(*line.detector, fresh)) return 9;
    // (5) the calls that need the GPU: verbatim addTemplate / match / getTemplates; without a device they must fail loudly
    cv::Mat color(480, 640, CV_8UC3), depth(480, 640, CV_16UC1), mask(480, 640, CV_8UC1);
    for (int y = 0; y < 480; ++y)
      for (int x = 0; x < 640; ++x) {
        const bool in = (x - 320) * (x - 320) / 4 + (y - 240) * (y - 240) < 90 * 90 && !((x / 16 + y / 16) & 1 && (x - 320) * (x - 320) + (y - 240) * (y - 240) < 60 * 60);
        unsigned char* p = color.ptr<unsigned char>(y) + 3 * x;
        p[0] = in ? 200 : 30; p[1] = in ? (unsigned char)(40 + x / 4) : 30; p[2] = in ? (unsigned char)(60 + y / 3) : 35;
        depth.at<unsigned short>(y, x) = in ? (unsigned short)(700 + (x - 320) / 2 + (y - 240) / 3) : 1200;
        mask.at<unsigned char>(y, x) = (x - 320) * (x - 320) / 4 + (y - 240) * (y - 240) < 95 * 95 ? 255 : 0;
      }
    std::vector<cv::Mat> templateImgs;
    templateImgs.push_back(color);
    if (!only_color) templateImgs.push_back(depth);
    cv::Rect bb;
    try {
      const bool added = line.addTemplateOnce(templateImgs, "planted", mask, bb);
      std::vector<cv::Mat> in_imgs;
      in_imgs.push_back(color);
      in_imgs.push_back(depth);                           // with the colour-only wiring detectTemplate pops it, as the reference does
      uint16_t cls = 0;
      for (uint16_t c = 0; c < line.getNumClasses(); ++c) if (line.getClassIds()[c] == "planted") cls = c;
      const bool found = line.detectTemplate(in_imgs, cls);
      if (in_imgs.size() != 2) return 10;
      if (added) {
        if (!found || line.matches[0].class_id != "planted" || line.matches[0].similarity < 99.f) { std::printf("planted template not found again\n"); return 11; }
        std::vector<cv::Point> pts;
        line.templatePoints(line.matches[0], pts);
        if (pts.empty() || bb.width <= 0) return 12;
        std::printf("GPU: wiring %d: template added (bb %d,%d %dx%d), %zu matches, best %.1f at (%d,%d), %zu hull points\n", only_color, bb.x, bb.y,
                    bb.width, bb.height, line.matches.size(), line.matches[0].similarity, line.matches[0].x, line.matches[0].y, pts.size());
      } else {
        std::printf("GPU: wiring %d: extraction failed on the synthetic view (allowed), %zu matches\n", only_color, line.matches.size());
      }
    } catch (const lm::Error& e) {
      if (e.code != LMB200_E_NODEVICE) { std::printf("unexpected: %s\n", e.what()); return 13; }
      std::printf("no GPU: %s\n", e.what());
    }
  }
  std::printf("DROPIN_OK\n");
  return 0;
}
