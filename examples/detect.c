/* examples/detect.c — the reference's detection loop (src/detector.cpp:17-45 -> PoseDetection -> HighLevelLineMOD::
 * detectTemplate, src/HighLevelLinemod.cpp:138-156) on the C ABI: load linemod_templates.yml.gz, match one RGB-D frame.
 *
 *   gcc -std=c99 -I include examples/detect.c -L line_mod_pipeline_b200 -llmb200 -Wl,-rpath,$PWD/line_mod_pipeline_b200 -o detect
 *   ./detect linemod_templates.yml.gz frame.bgr frame.depth 640 480 80
 *
 * frame.bgr = rows*cols*3 bytes (BGR, as cv::VideoCapture delivers), frame.depth = rows*cols u16 millimetres. */
#include <stdio.h>
#include <stdlib.h>
#include "lmb200.h"

static void* slurp(const char* path, size_t bytes) {
  FILE* f = fopen(path, "rb");
  void* p = NULL;
  if (!f) return NULL;
  if (lmb200_host_alloc(bytes, &p) != LMB200_OK) p = malloc(bytes); /* pinned when a GPU is there */
  if (p && fread(p, 1, bytes, f) != bytes) { p = NULL; }
  fclose(f);
  return p;
}

int main(int argc, char** argv) {
  if (argc < 7) { fprintf(stderr, "usage: %s templates.yml.gz frame.bgr frame.depth cols rows threshold\n", argv[0]); return 2; }
  const int cols = atoi(argv[4]), rows = atoi(argv[5]);
  const float threshold = (float)atof(argv[6]);
  lmb200_handle det = NULL;
  int rc = lmb200_read(argv[1], -1, &det); /* detector + every class of the file (HighLevelLinemod.cpp:288-300) */
  if (rc) { fprintf(stderr, "read: %d %s\n", rc, lmb200_last_error(NULL)); return 1; }
  printf("%d classes, %d templates, %d modalities, T0 = %d\n", lmb200_num_classes(det), lmb200_num_templates(det, NULL),
         lmb200_num_modalities(det), lmb200_get_T(det, 0));
  lmb200_image src[2];
  src[0].data = slurp(argv[2], (size_t)rows * cols * 3); src[0].rows = rows; src[0].cols = cols; src[0].type = LMB200_8UC3; src[0].step = 0;
  src[1].data = slurp(argv[3], (size_t)rows * cols * 2); src[1].rows = rows; src[1].cols = cols; src[1].type = LMB200_16UC1; src[1].step = 0;
  if (!src[0].data || !src[1].data) { fprintf(stderr, "cannot read the frame\n"); return 1; }
  size_t cap = 4096, n = 0;
  lmb200_match_rec* m = (lmb200_match_rec*)malloc(cap * sizeof *m);
  /* detector->match(in_imgs, detectorThreshold, matches, currentClass): all classes here (no class list) */
  rc = lmb200_match(det, src, lmb200_num_modalities(det), threshold, NULL, 0, m, cap, &n, NULL, NULL);
  if (rc == LMB200_E_TRUNCATED) { /* n = required count */
    cap = n; m = (lmb200_match_rec*)realloc(m, cap * sizeof *m);
    rc = lmb200_match(det, src, lmb200_num_modalities(det), threshold, NULL, 0, m, cap, &n, NULL, NULL);
  }
  if (rc) { fprintf(stderr, "match: %d %s\n", rc, lmb200_last_error(det)); return 1; }
  printf("%zu matches\n", n);
  for (size_t i = 0; i < n && i < 10; ++i)
    printf("  %-24s template %4d at (%4d, %4d) similarity %.2f\n", lmb200_class_id(det, m[i].class_index), m[i].template_id, m[i].x,
           m[i].y, m[i].similarity);
  free(m);
  lmb200_destroy(det);
  return 0;
}
