"""CPU-side checks of the product library: it loads without a GPU, exports every symbol the header
declares, its host logic (template store, tables, shard plan, merge) works, and every compute entry
point fails LOUDLY without a CUDA device (there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import numpy as np
import pytest

import line_mod_pipeline_b200 as lm
from line_mod_pipeline_b200 import capi as K, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_all_exported():
    hdr = open(os.path.join(ROOT, "include", "lmb200.h")).read()
    declared = set(re.findall(r"\b(lmb200_[a-zA-Z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 40
    out = subprocess.run(["nm", "-D", "--defined-only", K.SO_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (lmb200_[a-zA-Z0-9_]+)", out))
    assert declared <= exported, "declared but not exported: %r" % sorted(declared - exported)
    assert declared == set(K.SIGNATURES), "ctypes table out of sync: %r" % sorted(declared ^ set(K.SIGNATURES))
    L = K.lib()
    assert L.lmb200_version().startswith(b"lmb200")
    assert C.sizeof(K.MatchRec) == 20 and C.sizeof(K.Feature) == 12


def test_detector_surface_host_side():
    d = lm.getDefaultLINEMOD()
    assert d.getModalities() == ["ColorGradient", "DepthNormal"] and d.pyramidLevels() == 2
    assert (d.getT(0), d.getT(1)) == (5, 8) and d.numClasses() == 0 and d.numTemplates() == 0
    line = lm.getDefaultLINE()
    assert line.getModalities() == ["ColorGradient"]
    tps = synth.random_templates(5)
    for i, tp in enumerate(tps):
        assert d.addSyntheticTemplate(tp, "b_obj") == i
    assert d.addSyntheticTemplate(tps[0], "a_obj") == 0
    assert d.classIds() == ["a_obj", "b_obj"]            # std::map key order
    assert d.numTemplates() == 6 and d.numTemplates("b_obj") == 5 and d.numTemplates("zzz") == 0
    got = d.getTemplates("b_obj", 3)
    for a, b in zip(got, tps[3]):
        assert (a["width"], a["height"], a["pyramid_level"]) == (b["width"], b["height"], b["pyramid_level"])
        assert np.array_equal(a["features"], b["features"])
    with pytest.raises(lm.LinemodError) as e:
        d.getTemplates("b_obj", 99)
    assert e.value.code == K.E_CLASS
    bad = [dict(t, features=np.zeros((64, 3), np.int32)) for t in tps[0]]
    with pytest.raises(lm.LinemodError) as e:
        d.addSyntheticTemplate(bad, "x")                 # upstream CV_Assert(features.size() <= 63)
    assert e.value.code == K.E_FEATURES
    with pytest.raises(lm.LinemodError):
        d.addSyntheticTemplate(tps[0][:3], "x")          # wrong pyramid size
    for opt, val in (("early_exit", 0), ("upload_async", 1), ("cuda_graph", 0), ("shard_overlap", 2), ("shard_device_epilogue", 0),
                     ("host_threads", 3)):
        d.setOption(opt, val)                            # measurement / deployment knobs are host state: no device needed
    with pytest.raises(lm.LinemodError):
        d.setOption("no_such_option", 1)
    d.close()                                            # a handle that never touched the device (no epilogue thread to join)


def test_tables():
    d = lm.getDefaultLINEMOD()
    lut = d.getSimilarityLut()
    # default = circular distance (sum 628); the linear variant is explicit in lmb200_config
    assert lut[:16].tolist() == [0, 4, 3, 4, 2, 4, 3, 4, 1, 4, 3, 4, 2, 4, 3, 4] and int(lut.sum()) == 628
    assert lut[16:32].tolist() == [0, 0, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3]
    dl = lm.getDefaultLINEMOD(similarity_lut=K.SIMLUT_LINEAR)
    assert int(dl.getSimilarityLut().sum()) == 528 and not dl.getSimilarityLut()[16:32].any()
    assert np.array_equal(d.getNormalLut(), synth.default_normal_lut())
    with pytest.raises(lm.LinemodError):
        d.setNormalLut(np.full(8000, 3, np.uint8))       # not one-hot
    with pytest.raises(lm.LinemodError):
        d.setSimilarityLut(np.full(256, 5, np.uint8))    # 63 * 5 would overflow a byte
    from oracle import oracle as O
    assert np.array_equal(d.getSimilarityLut(), O.similarity_lut(1))
    d.setSimilarityLut(O.similarity_lut(0))
    assert np.array_equal(d.getSimilarityLut(), O.similarity_lut(0))


def test_normal_lut_provenance_and_loader(tmp_path):
    """The stand-in NORMAL_LUT is never silent: the handle says so, and upstream's normal_lut.i (C initialiser text) or a
    raw 8000-byte file replaces it."""
    d = lm.getDefaultLINEMOD()
    assert d.normalLutIsStandin()
    rng = np.random.default_rng(0)
    lut = (1 << rng.integers(0, 8, 8000)).astype(np.uint8)
    lut[::97] = 0
    rows = lut.reshape(20, 20, 20)
    txt = "// generated\n#define GRANULARITY 20\nstatic unsigned char NORMAL_LUT[20][20][20] = {\n"
    txt += ",\n".join("{" + ", ".join("{" + ", ".join(str(int(v)) for v in r) + "}" for r in plane) + "}" for plane in rows) + "};\n"
    p = tmp_path / "normal_lut.i"
    p.write_text(txt)
    d.loadNormalLut(p)
    assert not d.normalLutIsStandin() and np.array_equal(d.getNormalLut(), lut)
    d2 = lm.getDefaultLINEMOD()
    (tmp_path / "lut.bin").write_bytes(lut.tobytes())
    d2.loadNormalLut(tmp_path / "lut.bin")
    assert np.array_equal(d2.getNormalLut(), lut)
    (tmp_path / "bad.i").write_text("static unsigned char NORMAL_LUT[2] = {1, 2};")
    with pytest.raises(lm.LinemodError) as e:
        lm.getDefaultLINEMOD().loadNormalLut(tmp_path / "bad.i")
    assert e.value.code == K.E_IO
    d3 = lm.getDefaultLINEMOD()
    d3.setNormalLut(synth.default_normal_lut())
    assert not d3.normalLutIsStandin()
    assert lm.getDefaultLINE().warnings() == ""


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    d = lm.getDefaultLINEMOD()
    bgr, depth = synth.make_frame(0, 96, 160, n_shapes=4)
    for call in (lambda: d.match([bgr, depth], 80.0),
                 lambda: d.addTemplate([bgr, depth], "x", None),
                 lambda: d.addTemplates([[bgr, depth], [bgr, depth]], "x", [None, None]),
                 lambda: d.uploadFrames([[bgr, depth]]),
                 lambda: d.uploadTemplates(),
                 lambda: d.matchBatch([[bgr, depth]], 80.0)):
        with pytest.raises(lm.LinemodError) as e:
            call()
        assert e.value.code == K.E_NODEVICE and "no CPU fallback" in str(e.value)


def test_shard_plan_and_merge():
    assert lm.shard_plan([1, 1, 1, 1], 2) == [0, 2, 4]
    assert lm.shard_plan([], 3) == [0, 0, 0, 0]
    plan = lm.shard_plan(np.r_[np.ones(10), 100, np.ones(10)], 4)
    assert plan[0] == 0 and plan[-1] == 21 and all(a <= b for a, b in zip(plan, plan[1:]))
    rng = np.random.default_rng(0)
    costs = rng.uniform(1, 3, 1000)
    p = lm.shard_plan(costs, 8)
    sums = [costs[p[i]:p[i + 1]].sum() for i in range(8)]
    assert max(sums) / min(sums) < 1.05
    # merge == sort+unique of the concatenation (Match::operator< / operator== of upstream)
    a = np.array([(10, 10, 90.0, 0, 1), (20, 10, 95.0, 0, 1), (10, 10, 90.0, 0, 2)], lm.MATCH_DTYPE)
    b = np.array([(10, 10, 90.0, 0, 3), (5, 5, 99.0, 1, 0), (10, 10, 90.0, 0, 3)], lm.MATCH_DTYPE)
    m = lm.merge_matches([a, b])
    assert [tuple(x) for x in m.tolist()] == [(5, 5, 99.0, 1, 0), (20, 10, 95.0, 0, 1), (10, 10, 90.0, 0, 1)]


def test_cpp_surface_header_compiles_and_runs(tmp_path):
    """include/lmb200_detector.hpp (the cv::linemod::Detector-shaped C++ surface) against liblmb200.so."""
    exe = str(tmp_path / "cpp_surface_check")
    so_dir = os.path.dirname(K.SO_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp_surface_check.cpp"),
           "-o", exe, "-L", so_dir, "-llmb200", "-Wl,-rpath," + so_dir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and "CPP_SURFACE_OK" in r.stdout, r.stdout + r.stderr


def _build_dropin(tmp_path):
    exe = str(tmp_path / "cpp_dropin_check")
    so_dir = os.path.dirname(K.SO_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "tests", "cv_stub"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp_dropin_check.cpp"), "-o", exe, "-L", so_dir, "-llmb200", "-Wl,-rpath," + so_dir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_dropin_header_compiles_reference_call_sites_verbatim(tmp_path):
    """include/lmb200_opencv.hpp declares namespace cv::linemod: the reference's call expressions (HighLevelLinemod.cpp:26-43,
    :50-65, :92-101, :115-126, :142-156, :258-270, :292-300), copied verbatim into tests/cpp_dropin_check.cpp, compile and run
    through the C ABI (FileStorage / FileNode persistence round trip included).  The YAML it writes is read by real cv2."""
    exe = _build_dropin(tmp_path)
    env = dict(os.environ, LMB200_QUIET="1")
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "DROPIN_OK" in r.stdout, r.stdout + r.stderr
    cv2 = pytest.importorskip("cv2")
    fs = cv2.FileStorage(str(tmp_path / "linemod_templates.yml.gz"), cv2.FILE_STORAGE_READ)
    assert fs.isOpened() and int(fs.getNode("pyramid_levels").real()) == 2 and fs.getNode("classes").size() == 2
    tp = fs.getNode("classes").at(0).getNode("template_pyramids").at(1).getNode("templates").at(0)
    assert int(tp.getNode("width").real()) == 120 and tp.getNode("features").size() == 63


def test_header_is_plain_c(tmp_path):
    """include/lmb200.h is the drop-in boundary: it must compile as C99 (cgo / JNI / FFI generators read it) and link."""
    src = tmp_path / "cabi.c"
    src.write_text('#include "lmb200.h"\n'
                   'int main(void) { lmb200_config c; lmb200_default_config(&c, 1);\n'
                   '  return (c.num_modalities == 2 && c.pyramid_levels == 2 && lmb200_version() != 0) ? 0 : 1; }\n')
    exe = str(tmp_path / "cabi")
    so_dir = os.path.dirname(K.SO_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                        "-o", exe, "-L", so_dir, "-llmb200", "-Wl,-rpath," + so_dir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([exe]).returncode == 0


def test_c_example_builds_and_fails_loudly_without_gpu(tmp_path):
    """examples/detect.c (the reference's detection loop on the C ABI) builds as C99, loads the committed config-1
    template file and, on a box without a GPU, stops at lmb200_match with the no-device error."""
    exe = str(tmp_path / "detect")
    so_dir = os.path.dirname(K.SO_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "detect.c"), "-o", exe, "-L", so_dir, "-llmb200", "-Wl,-rpath," + so_dir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    z = np.load(os.path.join(ROOT, "tests", "golden", "fixture_frame.npz"))
    z["bgr"].tofile(str(tmp_path / "f.bgr")); z["depth"].tofile(str(tmp_path / "f.depth"))
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "lagergehaeuse_templates.yml.gz"), str(tmp_path / "f.bgr"),
                        str(tmp_path / "f.depth"), "640", "480", "80"], capture_output=True, text=True)
    assert "1 classes, 1950 templates, 2 modalities, T0 = 5" in r.stdout
    if _has_gpu():
        assert r.returncode == 0 and "matches" in r.stdout
    else:
        assert r.returncode == 1 and "no CUDA device" in r.stderr


def test_group_matches_equals_the_reference_loop():
    """lmb200_group_matches = groupSimilarMatches + discardSmallMatchGroups (src/HighLevelLinemod.cpp:206-253)."""
    rng = np.random.default_rng(5)
    for trial in range(20):
        n = int(rng.integers(1, 300))
        centres = rng.integers(0, 600, (int(rng.integers(1, 6)), 2))
        m = np.zeros(n, lm.MATCH_DTYPE)
        c = centres[rng.integers(0, len(centres), n)] + rng.integers(-30, 31, (n, 2))
        m["x"], m["y"] = c[:, 0], c[:, 1]
        radius, ratio = float(rng.choice([10.0, 35.5, 80.0])), float(rng.choice([0.0, 20.0, 50.0, 99.0]))
        # the reference, restated literally
        groups = []
        for i in range(n):
            for g in groups:
                if np.sqrt(float(m["x"][i] - g[0][0]) ** 2 + float(m["y"][i] - g[0][1]) ** 2) < np.float32(radius):
                    g[1].append(i)
                    break
            else:
                groups.append(((int(m["x"][i]), int(m["y"][i])), [i]))
        biggest = max(len(g[1]) for g in groups)
        want = np.full(n, -1, np.int32)
        kept = 0
        for g in groups:
            if np.float32(len(g[1]) * 100 // biggest) > np.float32(ratio):
                want[g[1]] = kept
                kept += 1
        got, ng = lm.group_matches(m, radius, ratio)
        assert ng == kept and np.array_equal(got, want), trial


def test_device_sort_restatement_equals_std_sort():
    """csrc/sort_emul.h (what the device epilogue runs) must leave the records in exactly std::sort's order, ties included:
    Match::operator< looks at (similarity, template_id) only and std::sort is not stable, so the reference's sequence — and
    which duplicates std::unique then finds adjacent — depends on libstdc++'s algorithm."""
    L = lm.capi.lib()
    rng = np.random.default_rng(7)
    MR = lm.capi.MatchRec

    def check(n, mode, sims=40, tids=50):
        a = np.zeros(n, lm.MATCH_DTYPE)
        a["x"] = rng.integers(0, 640, n); a["y"] = rng.integers(0, 480, n)
        a["similarity"] = 80.0 + 0.5 * rng.integers(0, sims, n)      # few distinct values: ties everywhere
        a["class_index"] = rng.integers(0, 10, n); a["template_id"] = rng.integers(0, tids, n)
        e, s = np.zeros(n, lm.MATCH_DTYPE), np.zeros(n, lm.MATCH_DTYPE)
        rc = L.lmb200_debug_sort_check(a.ctypes.data_as(C.POINTER(MR)), n, mode, e.ctypes.data_as(C.POINTER(MR)), s.ctypes.data_as(C.POINTER(MR)))
        assert rc == 0
        assert e.tobytes() == s.tobytes(), "n=%d mode=%d: restated sort differs from libstdc++'s" % (n, mode)
        return s

    for n in list(range(0, 40)) + [63, 64, 65, 100, 257, 1000, 4096, 5000]:
        for rep in range(3):
            check(n, 0)
            check(n, 0, sims=2, tids=3)        # almost everything ties
            check(n, 1)                        # heap sort (std::partial_sort over the whole range)
    for n in (17, 100, 1000, 5000):
        s = check(n, 2)                        # McIlroy's adversary: the depth limit is hit, heap-sort fallback inside std::sort
        assert list(s["template_id"]) == sorted(s["template_id"])


def test_python_mirror_matches_the_header():
    """The ctypes mirror's profile struct and kernel-family names follow include/lmb200.h's enum (a mismatch would shift
    every counter read through lmb200_get_profile)."""
    hdr = open(os.path.join(ROOT, "include", "lmb200.h")).read()
    body = hdr[hdr.index("LMB200_K_UPLOAD"):hdr.index("LMB200_K_COUNT")]
    names = re.findall(r"LMB200_K_([A-Z_]+)", "LMB200_K_UPLOAD" + body[len("LMB200_K_UPLOAD"):])
    assert [n.lower() for n in names] == K.K_NAMES
    assert C.sizeof(K.Profile) == 8 * (2 * len(K.K_NAMES) + 6)


def test_match_array_reads_like_the_reference_match():
    """Per-frame results: structured arrays whose fields read as attributes on the array and on its elements (what
    np.recarray offered at 20 us per slice)."""
    from line_mod_pipeline_b200.detector import _split, MatchArray
    out = np.zeros(10, lm.MATCH_DTYPE)
    out["x"] = np.arange(10); out["similarity"] = 90.0 - np.arange(10); out["template_id"] = 7
    offs = (C.c_size_t * 4)(0, 3, 3, 10)
    lists = _split(out, offs, 3)
    assert [len(l) for l in lists] == [3, 0, 7] and all(isinstance(l, MatchArray) for l in lists)
    assert list(lists[2].x) == list(range(3, 10)) and float(lists[0].similarity.sum()) == 90.0 + 89.0 + 88.0
    assert lists[2][0].x == 3 and lists[2][0].template_id == 7 and [int(m.x) for m in lists[0]] == [0, 1, 2]
    out["x"] = -1                                  # _split copied: the lists do not alias the reusable call buffer
    assert lists[0][0].x == 0
    with pytest.raises(AttributeError):
        lists[0].no_such_field


@pytest.mark.parametrize("world,frames,max_per_rank,interleaved", [(1, 3, 200, False), (2, 5, 300, True), (8, 6, 120, True), (4, 6, 100, False)])
def test_host_merge_of_a_sharded_step(world, frames, max_per_rank, interleaved):
    """lmb200_debug_merge_gathered = the merge the synchronous allgather fetch and the epilogue thread run: the gathered wire
    format of `world` ranks -> reference generation order -> std::sort + std::unique, against the same lists built in numpy
    and finished by lmb200_merge_matches.  (The device epilogue is checked against the same cases in test_gpu_parity.py.)"""
    from helpers import sharded_step_case
    L = lm.capi.lib()
    rng = np.random.default_rng(1000 * world + frames)
    ntpl = 64 * world
    G, gcap, pos, g_class, g_tid, want = sharded_step_case(rng, world, frames, max_per_rank, ntpl, interleaved, sims=6, tid_mod=7)
    total = sum(n for n, _ in want)
    out = np.zeros(max(1, total), lm.MATCH_DTYPE)
    offs = (C.c_size_t * (frames + 1))()
    rc = L.lmb200_debug_merge_gathered(G.ctypes.data, world, frames, gcap, pos.ctypes.data if pos is not None else None, g_class.ctypes.data,
                                       g_tid.ctypes.data, ntpl, out.ctypes.data_as(C.POINTER(lm.capi.MatchRec)), len(out), offs)
    assert rc == 0
    for f in range(frames):
        got = out[offs[f]:offs[f + 1]]
        assert got.tobytes() == want[f][1].tobytes(), "frame %d differs" % f
    G[world - 1, 0, 1] = 1                                    # a store-overflow flag in a header: the merge must refuse
    assert L.lmb200_debug_merge_gathered(G.ctypes.data, world, frames, gcap, None, g_class.ctypes.data, g_tid.ctypes.data, ntpl,
                                         out.ctypes.data_as(C.POINTER(lm.capi.MatchRec)), len(out), offs) == K.E_INVALID
