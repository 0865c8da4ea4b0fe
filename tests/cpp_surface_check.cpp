// Compile-and-run check of include/lmb200_detector.hpp against liblmb200.so (host-only calls: no GPU needed).
#include <cstdio>
#include <cstring>
#include "lmb200_detector.hpp"

int main(int argc, char** argv) {
  auto det = lm::getDefaultLINEMOD();
  if (det->getModalities().size() != 2 || det->getT(0) != 5 || det->getT(1) != 8 || det->pyramidLevels() != 2) return 1;
  std::vector<lm::Template> tp(4);
  for (int i = 0; i < 4; ++i) {
    tp[i].width = 100 >> (i / 2); tp[i].height = 80 >> (i / 2); tp[i].pyramid_level = i / 2;
    for (int k = 0; k < 10; ++k) tp[i].features.push_back(lm::Feature(k, 2 * k, k % 8));
  }
  if (det->addSyntheticTemplate(tp, "lagergehaeuse.ply") != 0 || det->addSyntheticTemplate(tp, "lagergehaeuse.ply") != 1) return 2;
  if (det->numTemplates() != 2 || det->numClasses() != 1 || det->classIds()[0] != "lagergehaeuse.ply") return 3;
  std::string path = std::string(argc > 1 ? argv[1] : "/tmp") + "/cpp_surface.yml.gz";
  det->write(path);
  auto back = lm::Detector::read(path);
  auto t = back->getTemplates("lagergehaeuse.ply", 1);
  if (t.size() != 4 || t[3].features.size() != 10 || t[3].features[9].y != 18 || t[2].width != 50) return 4;
  // writeClass / readClass with a class_id override (host only)
  std::string cpath = std::string(argc > 1 ? argv[1] : "/tmp") + "/cpp_surface_class.yml";
  det->writeClass("lagergehaeuse.ply", cpath);
  back->readClass(cpath, "copy");
  if (back->numClasses() != 2 || back->numTemplates("copy") != 2) return 6;
  try { back->readClass(cpath); return 7; } catch (const lm::Error& e) { if (e.code != LMB200_E_CLASS) return 8; }
  // headless renderer: a tetrahedron seen from +z
  {
    lm::Mesh mesh;
    mesh.vertices = {-50, -50, 0, 50, -50, 0, 0, 60, 0, 0, 0, 80};
    mesh.triangles = {0, 1, 2, 0, 1, 3, 1, 2, 3, 2, 0, 3};
    lm::RenderedViews rv = lm::renderLookAt(mesh, lm::referenceCamera(), {0, 0, 600, 0, 0, 900});
    const int d0 = rv.depth[(size_t)240 * 640 + 320], d1 = rv.depth[(size_t)640 * 480 + (size_t)240 * 640 + 320];  // just off the apex
    if (rv.n != 2 || d0 < 520 || d0 > 523 || d1 < 820 || d1 > 823) { std::printf("render: %d %d\n", d0, d1); return 11; }
    if (rv.colour[3 * ((size_t)240 * 640 + 320)] != 255 || rv.depth[0] != 0) return 12;
  }
  // match without a GPU must throw the loud no-device error; with a GPU it must return an (empty-ish) list
  std::vector<unsigned char> bgr(480 * 640 * 3, 0);
  std::vector<unsigned short> depth(480 * 640, 0);
  std::vector<lm::Match> matches;
  try {
    det->match({lm::ImageView(bgr.data(), 480, 640, LMB200_8UC3), lm::ImageView(depth.data(), 480, 640, LMB200_16UC1)}, 80.f, matches,
               {"lagergehaeuse.ply"});
    std::printf("match ran on a GPU: %zu matches\n", matches.size());
  } catch (const lm::Error& e) {
    if (e.code != LMB200_E_NODEVICE) { std::printf("unexpected: %s\n", e.what()); return 5; }
    std::printf("no GPU: %s\n", e.what());
  }
  // bulk addTemplates: same rule (two views of the blank frame: extraction fails, ids are -1, on a GPU)
  try {
    std::vector<lm::ImageView> view = {lm::ImageView(bgr.data(), 480, 640, LMB200_8UC3), lm::ImageView(depth.data(), 480, 640, LMB200_16UC1)};
    std::vector<int> ids = det->addTemplates({view, view}, "blank");
    if (ids.size() != 2 || ids[0] != -1 || ids[1] != -1) return 9;
  } catch (const lm::Error& e) {
    if (e.code != LMB200_E_NODEVICE) { std::printf("unexpected: %s\n", e.what()); return 10; }
  }
  // throughput / multi-GPU wrappers: host-side state changes work anywhere, device work obeys the same no-device rule
  det->setOption("host_threads", 2);
  det->setTemplateShard(0, 1);
  try { det->setOption("no_such_option", 1); return 13; } catch (const lm::Error& e) { if (e.code != LMB200_E_INVALID) return 14; }
  try {
    std::vector<lm::ImageView> view = {lm::ImageView(bgr.data(), 480, 640, LMB200_8UC3), lm::ImageView(depth.data(), 480, 640, LMB200_16UC1)};
    det->uploadFrames({view, view}, 0);
    det->matchResident(0, 2, 80.f);
    std::vector<std::vector<lm::Match>> lists = det->fetchResident(0, 2);
    if (lists.size() != 2) return 15;
    std::printf("resident step ran on a GPU: %zu + %zu matches\n", lists[0].size(), lists[1].size());
  } catch (const lm::Error& e) {
    if (e.code != LMB200_E_NODEVICE) { std::printf("unexpected: %s\n", e.what()); return 16; }
  }
  std::printf("CPP_SURFACE_OK\n");
  return 0;
}
