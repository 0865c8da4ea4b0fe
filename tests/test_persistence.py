"""Template-file persistence (OpenCV FileStorage YAML 1.0 / .gz) — the on-disk contract of
HighLevelLineMOD::writeLinemod / readLinemod (reference: src/HighLevelLinemod.cpp:256-300)."""
import os
import numpy as np
import pytest

import line_mod_pipeline_b200 as lm
from line_mod_pipeline_b200 import capi as K, synth


def _fill(d, n=7):
    tps = synth.random_templates(n, seed=5)
    for i, tp in enumerate(tps):
        d.addSyntheticTemplate(tp, "lagergehaeuse.ply" if i % 2 else "other obj")
    return tps


def _same(a, b):
    assert a.classIds() == b.classIds() and a.numTemplates() == b.numTemplates()
    assert a.getModalities() == b.getModalities() and a.pyramidLevels() == b.pyramidLevels()
    assert [a.getT(l) for l in range(a.pyramidLevels())] == [b.getT(l) for l in range(b.pyramidLevels())]
    for cid in a.classIds():
        assert a.numTemplates(cid) == b.numTemplates(cid)
        for t in range(a.numTemplates(cid)):
            for x, y in zip(a.getTemplates(cid, t), b.getTemplates(cid, t)):
                assert (x["width"], x["height"], x["pyramid_level"]) == (y["width"], y["height"], y["pyramid_level"])
                assert np.array_equal(x["features"], y["features"])


@pytest.mark.parametrize("name", ["linemod_templates.yml.gz", "linemod_templates.yml"])
def test_round_trip(tmp_path, name):
    d = lm.getDefaultLINEMOD()
    _fill(d)
    path = str(tmp_path / name)
    d.write(path)
    _same(d, lm.Detector.read(path))


def test_three_level_color_only_round_trip(tmp_path):
    d = lm.Detector([lm.ColorGradient(weak_threshold=12.5, num_features=40, strong_threshold=60.0)], (4, 8, 16))
    for tp in synth.random_templates(3, n_modalities=1, levels=3, nf0=40):
        d.addSyntheticTemplate(tp, "c")
    path = str(tmp_path / "t.yml.gz")
    d.write(path)
    _same(d, lm.Detector.read(path))


def test_real_opencv_reads_our_file(tmp_path):
    cv2 = pytest.importorskip("cv2")
    d = lm.getDefaultLINEMOD()
    tps = _fill(d)
    path = str(tmp_path / "linemod_templates.yml.gz")
    d.write(path)
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
    assert fs.isOpened() and int(fs.getNode("pyramid_levels").real()) == 2
    T = fs.getNode("T")
    assert [int(T.at(i).real()) for i in range(T.size())] == [5, 8]
    mods = fs.getNode("modalities")
    assert mods.at(0).getNode("type").string() == "ColorGradient" and mods.at(0).getNode("weak_threshold").real() == 10.0
    assert mods.at(1).getNode("type").string() == "DepthNormal" and int(mods.at(1).getNode("distance_threshold").real()) == 2000
    classes = fs.getNode("classes")
    assert classes.size() == 2 and classes.at(0).getNode("class_id").string() == "lagergehaeuse.ply"
    tp0 = classes.at(0).getNode("template_pyramids").at(0)
    assert int(tp0.getNode("template_id").real()) == 0
    t = tp0.getNode("templates").at(3)
    want = tps[1][3]
    assert int(t.getNode("width").real()) == want["width"] and int(t.getNode("pyramid_level").real()) == 1
    f = t.getNode("features")
    got = np.array([[int(f.at(i).at(j).real()) for j in range(3)] for i in range(f.size())])
    assert np.array_equal(got, want["features"])


def test_we_read_a_file_written_by_real_opencv(tmp_path):
    """Emit the layout with cv2.FileStorage itself (its own indentation/quoting/wrapping) and load it."""
    cv2 = pytest.importorskip("cv2")
    tps = synth.random_templates(3, seed=8)
    path = str(tmp_path / "cv_written.yml.gz")
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_WRITE)
    fs.write("pyramid_levels", 2)
    fs.startWriteStruct("T", cv2.FileNode_SEQ | cv2.FileNode_FLOW); fs.write("", 5); fs.write("", 8); fs.endWriteStruct()
    fs.startWriteStruct("modalities", cv2.FileNode_SEQ)
    fs.startWriteStruct("", cv2.FileNode_MAP); fs.write("type", "ColorGradient"); fs.write("weak_threshold", 10.0)
    fs.write("num_features", 63); fs.write("strong_threshold", 55.0); fs.endWriteStruct()
    fs.startWriteStruct("", cv2.FileNode_MAP); fs.write("type", "DepthNormal"); fs.write("distance_threshold", 2000)
    fs.write("difference_threshold", 50); fs.write("num_features", 63); fs.write("extract_threshold", 2); fs.endWriteStruct()
    fs.endWriteStruct()
    fs.startWriteStruct("classes", cv2.FileNode_SEQ)
    fs.startWriteStruct("", cv2.FileNode_MAP)
    fs.write("class_id", "lagergehaeuse.ply")
    fs.startWriteStruct("modalities", cv2.FileNode_SEQ | cv2.FileNode_FLOW); fs.write("", "ColorGradient"); fs.write("", "DepthNormal"); fs.endWriteStruct()
    fs.write("pyramid_levels", 2)
    fs.startWriteStruct("template_pyramids", cv2.FileNode_SEQ)
    for i, tp in enumerate(tps):
        fs.startWriteStruct("", cv2.FileNode_MAP)
        fs.write("template_id", i)
        fs.startWriteStruct("templates", cv2.FileNode_SEQ)
        for t in tp:
            fs.startWriteStruct("", cv2.FileNode_MAP)
            fs.write("width", int(t["width"])); fs.write("height", int(t["height"])); fs.write("pyramid_level", int(t["pyramid_level"]))
            fs.startWriteStruct("features", cv2.FileNode_SEQ)
            for x, y, l in t["features"]:
                fs.startWriteStruct("", cv2.FileNode_SEQ | cv2.FileNode_FLOW); fs.write("", int(x)); fs.write("", int(y)); fs.write("", int(l)); fs.endWriteStruct()
            fs.endWriteStruct()
            fs.endWriteStruct()
        fs.endWriteStruct()
        fs.endWriteStruct()
    fs.endWriteStruct()
    fs.endWriteStruct()
    fs.endWriteStruct()
    fs.release()
    d = lm.Detector.read(path)
    assert d.classIds() == ["lagergehaeuse.ply"] and d.numTemplates() == 3 and d.getModalities() == ["ColorGradient", "DepthNormal"]
    for i, tp in enumerate(tps):
        for a, b in zip(d.getTemplates("lagergehaeuse.ply", i), tp):
            assert (a["width"], a["height"], a["pyramid_level"]) == (b["width"], b["height"], b["pyramid_level"])
            assert np.array_equal(a["features"], b["features"])


def test_write_read_classes(tmp_path):
    d = lm.getDefaultLINEMOD()
    _fill(d)
    fmt = str(tmp_path / "templates_%s.yml.gz")
    d.writeClasses(fmt)
    assert sorted(os.listdir(tmp_path)) == ["templates_lagergehaeuse.ply.yml.gz", "templates_other obj.yml.gz"]
    e = lm.getDefaultLINEMOD()
    e.readClasses(["other obj", "lagergehaeuse.ply"], fmt)
    _same(d, e)
    with pytest.raises(lm.LinemodError) as err:      # upstream: CV_Assert(class not already present)
        e.readClasses(["other obj"], fmt)
    assert err.value.code == K.E_CLASS
    line = lm.getDefaultLINE()
    with pytest.raises(lm.LinemodError) as err:      # upstream: CV_Assert(modalities match)
        line.readClasses(["other obj"], fmt)
    assert err.value.code == K.E_CLASS
    with pytest.raises(lm.LinemodError) as err:
        lm.Detector.read(str(tmp_path / "missing.yml.gz"))
    assert err.value.code == K.E_IO


def test_write_read_single_class(tmp_path):
    """Detector::writeClass / readClass(fn, class_id_override): one class per file, override semantics of upstream."""
    d = lm.getDefaultLINEMOD()
    _fill(d)
    path = str(tmp_path / "one_class.yml")
    d.writeClass("other obj", path)
    e = lm.getDefaultLINEMOD()
    e.readClass(path)
    assert e.classIds() == ["other obj"] and e.numTemplates() == d.numTemplates("other obj")
    for t in range(e.numTemplates("other obj")):
        for a, b in zip(e.getTemplates("other obj", t), d.getTemplates("other obj", t)):
            assert (a["width"], a["height"], a["pyramid_level"]) == (b["width"], b["height"], b["pyramid_level"])
            assert np.array_equal(a["features"], b["features"])
    with pytest.raises(lm.LinemodError) as err:      # upstream: CV_Assert(class not already present)
        e.readClass(path)
    assert err.value.code == K.E_CLASS
    e.readClass(path, "renamed")                     # override: stored under the new name
    assert e.classIds() == ["other obj", "renamed"] and e.numTemplates("renamed") == e.numTemplates("other obj")
    n = e.numTemplates()
    e.readClass(path, "renamed")                     # override onto an existing entry: std::map::insert keeps the old one
    assert e.numTemplates() == n
    with pytest.raises(lm.LinemodError) as err:
        d.writeClass("no such class", path)
    assert err.value.code == K.E_CLASS
    cv2 = pytest.importorskip("cv2")
    f = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)  # real OpenCV parses the single-class file
    assert f.getNode("class_id").string() == "other obj" and int(f.getNode("pyramid_levels").real()) == 2
    f.release()


def test_config1_template_file_loads():
    """The committed config-1 template set (reference file layout) parses into 1 950 pyramids of 63/63/31/31 features."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lagergehaeuse_templates.yml.gz")
    d = lm.Detector.read(path)
    assert d.classIds() == ["lagergehaeuse.ply"] and d.numTemplates() == 1950
    assert d.getModalities() == ["ColorGradient", "DepthNormal"] and [d.getT(0), d.getT(1)] == [5, 8]
    for t in (0, 977, 1949):
        tp = d.getTemplates("lagergehaeuse.ply", t)
        assert [len(x["features"]) for x in tp] == [63, 63, 31, 31]
        assert [x["pyramid_level"] for x in tp] == [0, 0, 1, 1]
        assert tp[0]["width"] == tp[1]["width"] and tp[2]["width"] == tp[0]["width"] >> 1


def test_binary_cache_round_trip_and_speed(tmp_path):
    import time
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lagergehaeuse_templates.yml.gz")
    t0 = time.perf_counter(); d = lm.Detector.read(path); t_yaml = time.perf_counter() - t0
    cache = str(tmp_path / "templates.lmb200")
    d.writeCache(cache)
    t0 = time.perf_counter(); e = lm.Detector.readCache(cache); t_bin = time.perf_counter() - t0
    _same(d, e)
    assert t_bin < t_yaml
    with pytest.raises(lm.LinemodError) as err:
        lm.Detector.readCache(path)          # a YAML file is not a cache
    assert err.value.code == K.E_IO


def test_cache_refuses_what_it_cannot_hold(tmp_path):
    """The binary cache stores coordinates in 16 bits: a (synthetic) feature outside that range is refused, the YAML
    writer keeps it."""
    d = lm.getDefaultLINEMOD()
    tps = [dict(width=100, height=80, pyramid_level=i // 2, features=np.array([[70000, 3, 1]], np.int32)) for i in range(4)]
    d.addSyntheticTemplate(tps, "far")
    with pytest.raises(lm.LinemodError) as err:
        d.writeCache(str(tmp_path / "c.bin"))
    assert err.value.code == K.E_INVALID
    d.write(str(tmp_path / "c.yml"))
    assert lm.Detector.read(str(tmp_path / "c.yml")).getTemplates("far", 0)[0]["features"][0][0] == 70000


def test_pose_sidecar_round_trip(tmp_path):
    """linemod_tempPosFile.bin: u32 class count, per class u64 n + n x 48-byte HighLevelLineMOD::Template records."""
    import struct
    rng = np.random.default_rng(3)
    classes = []
    for n in (5, 0, 3):
        a = np.zeros(n, lm.POSE_DTYPE)
        a["translation"] = rng.normal(size=(n, 3)); a["quaternion"] = rng.normal(size=(n, 4))
        a["bb"] = rng.integers(0, 640, (n, 4)); a["median_depth"] = rng.integers(400, 1200, n)
        classes.append(a)
    p = str(tmp_path / "linemod_tempPosFile.bin")
    lm.write_pose_sidecar(p, classes)
    raw = open(p, "rb").read()
    assert struct.unpack_from("<I", raw, 0)[0] == 3 and struct.unpack_from("<Q", raw, 4)[0] == 5
    assert len(raw) == 4 + 3 * 8 + 8 * 48
    tx, ty, tz, qx, qy, qz, qw, bx, by, bw, bh, md = struct.unpack_from("<3f4f4iH", raw, 12)   # the reference's raw struct dump
    assert (bx, by, bw, bh, md) == tuple(classes[0]["bb"][0]) + (classes[0]["median_depth"][0],)
    for i, a in enumerate(classes):
        b = lm.read_pose_sidecar(p, i)
        assert np.array_equal(a.tobytes(), b.tobytes())
    with pytest.raises(lm.LinemodError):
        lm.read_pose_sidecar(p, 7)


def test_corrupted_files_fail_cleanly(tmp_path):
    """Truncated / mutated template files and caches either load or raise LinemodError — never crash — and whatever
    loads can be queried (hand-written parsers: persistence.cpp)."""
    import random
    d = lm.getDefaultLINEMOD()
    _fill(d)
    yml, cache = str(tmp_path / "ok.yml"), str(tmp_path / "ok.bin")
    d.write(yml)
    d.writeCache(cache)
    txt, blob = open(yml, "rb").read(), open(cache, "rb").read()
    rng = random.Random(0)
    outcomes = {"yaml_ok": 0, "yaml_err": 0, "cache_ok": 0, "cache_err": 0}
    for it in range(160):
        b = bytearray(txt)
        mode = it % 4
        if mode == 0:
            b = b[:rng.randrange(len(b))]
        elif mode == 1:
            for _ in range(rng.randrange(1, 6)):
                b[rng.randrange(len(b))] = rng.randrange(32, 127)
        elif mode == 2:
            i = rng.randrange(len(b)); del b[i:min(len(b), i + rng.randrange(1, 200))]
        else:
            i = rng.randrange(len(b)); b[i:i] = bytes(rng.randrange(32, 127) for _ in range(rng.randrange(1, 50)))
        p = str(tmp_path / "m.yml")
        open(p, "wb").write(bytes(b))
        try:
            e = lm.Detector.read(p)
            for cid in e.classIds():
                for t in range(e.numTemplates(cid)):
                    e.getTemplates(cid, t)
            e.close()
            outcomes["yaml_ok"] += 1
        except lm.LinemodError:
            outcomes["yaml_err"] += 1
    for it in range(120):
        b = bytearray(blob)
        if it % 2 == 0:
            b = b[:rng.randrange(len(b))]
        else:
            for _ in range(rng.randrange(1, 6)):
                b[rng.randrange(len(b))] = rng.randrange(256)
        p = str(tmp_path / "m.bin")
        open(p, "wb").write(bytes(b))
        try:
            e = lm.Detector.readCache(p)
            e.numTemplates()
            e.close()
            outcomes["cache_ok"] += 1
        except lm.LinemodError:
            outcomes["cache_err"] += 1
    assert outcomes["yaml_err"] > 10 and outcomes["cache_err"] > 10, outcomes
    # pose sidecar: a count field beyond the file must be an error, not an allocation request
    side = str(tmp_path / "poses.bin")
    lm.write_pose_sidecar(side, [np.zeros(5, lm.POSE_DTYPE), np.zeros(3, lm.POSE_DTYPE)])
    blob = bytearray(open(side, "rb").read())
    blob[4:12] = (2 ** 60).to_bytes(8, "little")
    open(side, "wb").write(bytes(blob))
    for ci in (0, 1):
        with pytest.raises(lm.LinemodError):
            lm.read_pose_sidecar(side, ci)


def test_class_id_with_bracket_round_trips(tmp_path):
    """a quoted class_id containing '[' must not be mistaken for a wrapped flow sequence."""
    d = lm.getDefaultLINEMOD()
    for i, tp in enumerate(synth.random_templates(4, seed=8)):
        d.addSyntheticTemplate(tp, "a[b" if i % 2 else "c]d [e")
    path = str(tmp_path / "brackets.yml.gz")
    d.write(path)
    _same(d, lm.Detector.read(path))
    cv2 = pytest.importorskip("cv2")
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
    ids = sorted(fs.getNode("classes").at(i).getNode("class_id").string() for i in range(2))
    assert ids == ["a[b", "c]d [e"]


def test_bare_class_file_with_modalities_before_class_id(tmp_path):
    """key order inside a map is free: `modalities` listed before `class_id` is one class, not two."""
    d = lm.getDefaultLINEMOD()
    tps = synth.random_templates(2, seed=9)
    for tp in tps:
        d.addSyntheticTemplate(tp, "obj")
    p = str(tmp_path / "templates_obj.yml")
    d.writeClass("obj", p)
    lines = open(p).read().split("\n")
    ci = next(i for i, l in enumerate(lines) if l.startswith("class_id"))
    mi = next(i for i, l in enumerate(lines) if l.startswith("modalities"))
    assert ci < mi
    lines[ci], lines[mi] = lines[mi], lines[ci]
    q = str(tmp_path / "swapped.yml")
    open(q, "w").write("\n".join(lines))
    e = lm.getDefaultLINEMOD()
    e.readClass(q)
    _same(d, e)


def test_loaded_templates_are_validated(tmp_path):
    """> 63 features or a label outside 0..7 is refused at READ time with upstream's error class, not at the
    next match."""
    d = lm.getDefaultLINEMOD()
    d.addSyntheticTemplate(synth.random_templates(1, seed=10)[0], "obj")
    p = str(tmp_path / "t.yml")
    d.write(p)
    txt = open(p).read()
    first = txt.index("- [")
    line = txt[first:txt.index("\n", first)]
    bad_label = txt.replace(line, "- [ 3, 4, 9 ]", 1)
    open(str(tmp_path / "label.yml"), "w").write(bad_label)
    with pytest.raises(lm.LinemodError) as e:
        lm.Detector.read(str(tmp_path / "label.yml"))
    assert e.value.code == K.E_IO
    many = txt.replace(line, "\n".join([line] * 70), 1)
    open(str(tmp_path / "many.yml"), "w").write(many)
    with pytest.raises(lm.LinemodError) as e:
        lm.Detector.read(str(tmp_path / "many.yml"))
    assert e.value.code == K.E_FEATURES


def test_write_errors_are_reported(tmp_path):
    """a failing write is an error, not a silently truncated template file (/dev/full: every write fails)."""
    if not os.path.exists("/dev/full"):
        pytest.skip("no /dev/full")
    d = lm.getDefaultLINEMOD()
    _fill(d, 40)
    with pytest.raises(lm.LinemodError) as e:
        d.write("/dev/full")
    assert e.value.code == K.E_IO
    with pytest.raises(lm.LinemodError):
        d.writeCache("/dev/full")
