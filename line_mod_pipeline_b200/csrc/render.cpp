// render.cpp — headless depth / silhouette rasteriser: replaces the SDL + OpenGL offscreen passes
// the reference uses to synthesise template views and benchmark renders (src/OpenglRender.cpp:49-141 with
// shader/depth.fs:1-18; callers src/HighLevelLinemod.cpp:68-110 via TemplateGenerator, src/Benchmark.cpp:18-38,:156-170).
// Host code: a z-buffer rasteriser over the model's triangles, one view per thread.  What it reproduces of the GL passes:
//   * pinhole projection u = cx + fx*x/z, v = cy - fy*y/z of camera-space points (camera looks down -z, +y up), i.e. the
//     image after the reference's vertical flip of the read-back (OpenglRender.cpp:33-47);
//   * the view matrix of glm::lookAt(eye, 0, +Y) (OpenglRender.cpp:334-345) or an explicit rotation + translation
//     (the benchmark's overloads, OpenglRender.cpp:69-94,:116-141); modelMat is never applied upstream (:88,:135);
//   * near plane: triangles with a vertex closer than `near_mm` are dropped (GL clips them; template views never get there);
//   * depth pass: perspective-correct camera-space depth of the nearest surface, rounded to the nearest millimetre into a
//     16-bit image, 0 = background (depth.fs writes linear z scaled so that a u16 read-back is millimetres);
//   * colour pass: models without vertex colours are drawn white (ModelImporter.cpp:53-71); the pipeline thresholds the
//     colour image to binary right away (HighLevelLinemod.cpp:77), so the silhouette is what matters.
// Pixel-exact equality with a GPU's rasteriser is neither possible nor needed; the tests compare with
// the numpy rasteriser that produced the committed config-1 template set, depth within 1 mm.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lmb200.h"

namespace {

struct View { double R[9]; double t[3]; };  // x_cam = R * x + t

void look_at(const double* eye_in, View& v) {
  double eye[3] = {eye_in[0], eye_in[1], eye_in[2]};
  if (eye[0] == 0 && eye[2] == 0) eye[0] = eye[2] = 1e-6;  // straight up/down: the cross product with +Y vanishes
  const double n = std::sqrt(eye[0] * eye[0] + eye[1] * eye[1] + eye[2] * eye[2]);
  const double f[3] = {-eye[0] / n, -eye[1] / n, -eye[2] / n};
  double s[3] = {f[1] * 0.0 - f[2] * 1.0, f[2] * 0.0 - f[0] * 0.0, f[0] * 1.0 - f[1] * 0.0};  // f x (0,1,0)
  const double sn = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
  for (double& c : s) c /= sn;
  const double u[3] = {s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]};  // s x f
  const double R[9] = {s[0], s[1], s[2], u[0], u[1], u[2], -f[0], -f[1], -f[2]};
  std::memcpy(v.R, R, sizeof R);
  for (int i = 0; i < 3; ++i) v.t[i] = -(R[3 * i] * eye[0] + R[3 * i + 1] * eye[1] + R[3 * i + 2] * eye[2]);
}

void render_one(const lmb200_mesh& mesh, const lmb200_camera& cam, const View& view, uint16_t* depth_out, uint8_t* colour_out,
                std::vector<double>& zbuf, std::vector<double>& proj) {
  const int W = cam.width, H = cam.height;
  const double inf = std::numeric_limits<double>::infinity();
  zbuf.assign((size_t)W * H, inf);
  proj.resize((size_t)mesh.n_vertices * 3);  // u, v, z per vertex
  for (int i = 0; i < mesh.n_vertices; ++i) {
    const double* p = mesh.vertices + 3 * (size_t)i;
    const double x = view.R[0] * p[0] + view.R[1] * p[1] + view.R[2] * p[2] + view.t[0];
    const double y = view.R[3] * p[0] + view.R[4] * p[1] + view.R[5] * p[2] + view.t[1];
    const double zc = view.R[6] * p[0] + view.R[7] * p[1] + view.R[8] * p[2] + view.t[2];
    const double z = -zc;
    proj[3 * (size_t)i] = cam.cx + cam.fx * x / z;
    proj[3 * (size_t)i + 1] = cam.cy - cam.fy * y / z;
    proj[3 * (size_t)i + 2] = z;
  }
  for (int k = 0; k < mesh.n_triangles; ++k) {
    const int a = mesh.triangles[3 * (size_t)k], b = mesh.triangles[3 * (size_t)k + 1], c = mesh.triangles[3 * (size_t)k + 2];
    const double ua = proj[3 * (size_t)a], va = proj[3 * (size_t)a + 1], za = proj[3 * (size_t)a + 2];
    const double ub = proj[3 * (size_t)b], vb = proj[3 * (size_t)b + 1], zb = proj[3 * (size_t)b + 2];
    const double uc = proj[3 * (size_t)c], vc = proj[3 * (size_t)c + 1], zc = proj[3 * (size_t)c + 2];
    if (za < cam.near_mm || zb < cam.near_mm || zc < cam.near_mm) continue;
    int x0 = (int)std::floor(std::min(ua, std::min(ub, uc))), x1 = (int)std::ceil(std::max(ua, std::max(ub, uc)));
    int y0 = (int)std::floor(std::min(va, std::min(vb, vc))), y1 = (int)std::ceil(std::max(va, std::max(vb, vc)));
    x0 = std::max(x0, 0); y0 = std::max(y0, 0); x1 = std::min(x1, W - 1); y1 = std::min(y1, H - 1);
    if (x0 > x1 || y0 > y1) continue;
    const double d = (vb - vc) * (ua - uc) + (uc - ub) * (va - vc);
    if (std::fabs(d) < 1e-12) continue;
    for (int y = y0; y <= y1; ++y) {
      const double ys = y + 0.5;
      double* zrow = zbuf.data() + (size_t)y * W;
      for (int x = x0; x <= x1; ++x) {
        const double xs = x + 0.5;
        const double l0 = ((vb - vc) * (xs - uc) + (uc - ub) * (ys - vc)) / d;
        const double l1 = ((vc - va) * (xs - uc) + (ua - uc) * (ys - vc)) / d;
        const double l2 = 1 - l0 - l1;
        if (!(l0 >= 0 && l1 >= 0 && l2 >= 0)) continue;
        const double zi = 1.0 / (l0 / za + l1 / zb + l2 / zc);  // perspective-correct depth
        if (zi < zrow[x]) zrow[x] = zi;
      }
    }
  }
  for (size_t i = 0; i < (size_t)W * H; ++i) {
    const bool hit = zbuf[i] < inf;
    uint16_t dmm = 0;
    if (hit) {
      double r = std::nearbyint(zbuf[i]);  // round half to even, like the numpy harness
      dmm = (uint16_t)std::min(65535.0, std::max(0.0, r));
    }
    if (depth_out) depth_out[i] = dmm;
    if (colour_out) { const uint8_t cval = hit ? 255 : 0; colour_out[3 * i] = cval; colour_out[3 * i + 1] = cval; colour_out[3 * i + 2] = cval; }
  }
}

int check(const lmb200_mesh* mesh, const lmb200_camera* cam) {
  if (!mesh || !cam || !mesh->vertices || !mesh->triangles || mesh->n_vertices <= 0 || mesh->n_triangles <= 0) return LMB200_E_INVALID;
  if (cam->width <= 0 || cam->height <= 0 || !(cam->fx > 0) || !(cam->fy > 0) || !(cam->near_mm > 0)) return LMB200_E_INVALID;
  for (int i = 0; i < 3 * mesh->n_triangles; ++i)
    if (mesh->triangles[i] < 0 || mesh->triangles[i] >= mesh->n_vertices) return LMB200_E_INVALID;
  return LMB200_OK;
}

int render_many(const lmb200_mesh* mesh, const lmb200_camera* cam, const std::vector<View>& views, uint16_t* depth_out,
                uint8_t* colour_out, int threads) {
  const int n = (int)views.size();
  const size_t px = (size_t)cam->width * cam->height;
  int hw = (int)std::thread::hardware_concurrency();
  int nt = threads > 0 ? threads : (hw > 0 ? hw : 1);
  nt = std::max(1, std::min(nt, n));
  std::atomic<int> next(0);
  auto work = [&]() {
    std::vector<double> zbuf, proj;
    for (int i; (i = next.fetch_add(1)) < n;)
      render_one(*mesh, *cam, views[i], depth_out ? depth_out + px * i : nullptr, colour_out ? colour_out + 3 * px * i : nullptr, zbuf, proj);
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  return LMB200_OK;
}

}  // namespace

extern "C" {

int lmb200_render_lookat(const lmb200_mesh* mesh, const lmb200_camera* cam, const double* eyes, int n_views, uint16_t* depth_out,
                         uint8_t* colour_out, int threads) {
  int rc = check(mesh, cam);
  if (rc) return rc;
  if (!eyes || n_views < 0 || (!depth_out && !colour_out)) return LMB200_E_INVALID;
  std::vector<View> views((size_t)n_views);
  for (int i = 0; i < n_views; ++i) look_at(eyes + 3 * (size_t)i, views[i]);
  return render_many(mesh, cam, views, depth_out, colour_out, threads);
}

int lmb200_render_pose(const lmb200_mesh* mesh, const lmb200_camera* cam, const double* rotations, const double* translations,
                       int n_views, uint16_t* depth_out, uint8_t* colour_out, int threads) {
  int rc = check(mesh, cam);
  if (rc) return rc;
  if (!rotations || !translations || n_views < 0 || (!depth_out && !colour_out)) return LMB200_E_INVALID;
  std::vector<View> views((size_t)n_views);
  for (int i = 0; i < n_views; ++i) {
    std::memcpy(views[i].R, rotations + 9 * (size_t)i, sizeof views[i].R);
    std::memcpy(views[i].t, translations + 3 * (size_t)i, sizeof views[i].t);
  }
  return render_many(mesh, cam, views, depth_out, colour_out, threads);
}

// ASCII PLY (what the reference ships in models/): x y z first on every vertex line, polygons triangulated as fans.
int lmb200_load_ply(const char* path, double** vertices, int* n_vertices, int** triangles, int* n_triangles) {
  if (!path || !vertices || !n_vertices || !triangles || !n_triangles) return LMB200_E_INVALID;
  FILE* fp = std::fopen(path, "r");
  if (!fp) return LMB200_E_IO;
  char line[1024];
  int nv = -1, nf = -1;
  bool ascii = false, header_done = false;
  if (!std::fgets(line, sizeof line, fp) || std::strncmp(line, "ply", 3) != 0) { std::fclose(fp); return LMB200_E_IO; }
  while (std::fgets(line, sizeof line, fp)) {
    if (std::strncmp(line, "format ascii", 12) == 0) ascii = true;
    else if (std::sscanf(line, "element vertex %d", &nv) == 1) {}
    else if (std::sscanf(line, "element face %d", &nf) == 1) {}
    else if (std::strncmp(line, "end_header", 10) == 0) { header_done = true; break; }
  }
  if (!header_done || !ascii || nv <= 0 || nf < 0) { std::fclose(fp); return LMB200_E_IO; }
  std::vector<double> v((size_t)nv * 3);
  for (int i = 0; i < nv; ++i) {
    if (!std::fgets(line, sizeof line, fp) || std::sscanf(line, "%lf %lf %lf", &v[3 * (size_t)i], &v[3 * (size_t)i + 1], &v[3 * (size_t)i + 2]) != 3) {
      std::fclose(fp);
      return LMB200_E_IO;
    }
  }
  std::vector<int> tri;
  for (int i = 0; i < nf; ++i) {
    if (!std::fgets(line, sizeof line, fp)) { std::fclose(fp); return LMB200_E_IO; }
    char* p = line;
    long cnt = std::strtol(p, &p, 10);
    if (cnt < 3 || cnt > 64) { std::fclose(fp); return LMB200_E_IO; }
    std::vector<int> idx((size_t)cnt);
    for (long k = 0; k < cnt; ++k) {
      char* q = p;
      idx[(size_t)k] = (int)std::strtol(p, &q, 10);
      if (q == p || idx[(size_t)k] < 0 || idx[(size_t)k] >= nv) { std::fclose(fp); return LMB200_E_IO; }
      p = q;
    }
    for (long k = 1; k + 1 < cnt; ++k) { tri.push_back(idx[0]); tri.push_back(idx[(size_t)k]); tri.push_back(idx[(size_t)k + 1]); }
  }
  std::fclose(fp);
  *vertices = (double*)std::malloc(v.size() * sizeof(double));
  *triangles = (int*)std::malloc(std::max<size_t>(1, tri.size()) * sizeof(int));
  if (!*vertices || !*triangles) { std::free(*vertices); std::free(*triangles); return LMB200_E_INVALID; }
  std::memcpy(*vertices, v.data(), v.size() * sizeof(double));
  std::memcpy(*triangles, tri.data(), tri.size() * sizeof(int));
  *n_vertices = nv;
  *n_triangles = (int)(tri.size() / 3);
  return LMB200_OK;
}

void lmb200_free(void* p) { std::free(p); }

// Benchmark::calculateErrorHodan (src/Benchmark.cpp:18-38) with calculateVisibilityMasks (:133-154), restated per pixel.
// The reference works on CV_16U images: Mat subtraction saturates at 0, cv::threshold(..., 65536, ...) yields 65535, and the
// masks are combined with bitwise and/or of those 16-bit words, which is what the nested conditions below spell out:
//   gtVis  = gt > 1 and not (gt - in > visThr)                      (:138-142)
//   estVis = (est > 1 and not (est - in > visThr)) | (gtVis & est)  (:144-150; the `&` is bitwise on the depth VALUE)
//   inter  = gtVis & estVis, comb = gtVis | estVis                  (:152-153)
//   ok     = inter & (|gt - est| <= errThr)                         (:28-30)
//   error  = 1 - countNonZero(ok) / countNonZero(comb)              (:31-32; 0/0 -> NaN like the reference's float division)
int lmb200_hodan_error(const uint16_t* input_depth, const uint16_t* gt_render, const uint16_t* est_render, int rows, int cols,
                       int visibility_threshold, int error_threshold, float* error, long long* n_ok, long long* n_comb) {
  if (!input_depth || !gt_render || !est_render || !error || rows <= 0 || cols <= 0) return LMB200_E_INVALID;
  long long ok = 0, comb = 0;
  const size_t n = (size_t)rows * cols;
  for (size_t i = 0; i < n; ++i) {
    const int in = input_depth[i], gt = gt_render[i], est = est_render[i];
    const unsigned gt_occl = (gt > in && gt - in > visibility_threshold) ? 65535u : 0u;    // saturating gt - in, then THRESH_BINARY
    const unsigned est_occl = (est > in && est - in > visibility_threshold) ? 65535u : 0u;
    const unsigned gt_bin = gt > 1 ? 65535u : 0u, est_bin = est > 1 ? 65535u : 0u;
    const unsigned gt_vis = gt_bin > gt_occl ? gt_bin - gt_occl : 0u;                        // saturating subtraction
    unsigned est_vis = est_bin > est_occl ? est_bin - est_occl : 0u;
    est_vis |= (gt_vis & (unsigned)est);
    const unsigned inter = gt_vis & est_vis, cmb = gt_vis | est_vis;
    const int ad = gt > est ? gt - est : est - gt;
    const unsigned close = ad > error_threshold ? 0u : 65535u;                               // THRESH_BINARY_INV
    if (inter & close) ++ok;
    if (cmb) ++comb;
  }
  *error = 1.0f - (float)ok / (float)comb;
  if (n_ok) *n_ok = ok;
  if (n_comb) *n_comb = comb;
  return LMB200_OK;
}

// The whole call of PoseDetection.cpp:99: render the model at the ground-truth and at the estimated pose (Benchmark::renderPose
// -> OpenGLRender::renderDepthToFrontBuff, here the headless rasteriser) and score them against the input depth image.
// rotations: two 3x3 row-major model-view rotations (ground truth, estimate), translations: two xyz (see lmb200_render_pose).
int lmb200_hodan_error_poses(const lmb200_mesh* mesh, const lmb200_camera* cam, const double* rotations, const double* translations,
                             const uint16_t* input_depth, int visibility_threshold, int error_threshold, float* error) {
  if (!mesh || !cam || !rotations || !translations || !input_depth || !error) return LMB200_E_INVALID;
  std::vector<uint16_t> depth((size_t)2 * cam->width * cam->height);
  int rc = lmb200_render_pose(mesh, cam, rotations, translations, 2, depth.data(), nullptr, 2);
  if (rc) return rc;
  return lmb200_hodan_error(input_depth, depth.data(), depth.data() + (size_t)cam->width * cam->height, cam->height, cam->width,
                            visibility_threshold, error_threshold, error, nullptr, nullptr);
}

}  // extern "C"
