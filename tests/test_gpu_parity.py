"""GPU parity tests proper: the CUDA path (through the C ABI) vs the CPU oracle, bit-exact.

Everything here needs a B200 (`-m gpu`).  Sizes are the BASELINE.json configs where the oracle
finishes in seconds, plus edge cases (ragged geometry, empty sets, malformed templates, masks).
"""
import ctypes as C
import hashlib
import numpy as np
import pytest

import line_mod_pipeline_b200 as lm
from line_mod_pipeline_b200 import synth, capi as K
from oracle import oracle as O
from helpers import make_pair, sources, add_random, add_planted_from_oracle, assert_same_matches, sharded_step_case as _epilogue_case

pytestmark = pytest.mark.gpu
sha = lambda a: hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def _check_frame_side(det, ora, srcs, threshold=80.0, n_maps=4, masks=None):
    got, qimgs = det.match(srcs, threshold, quantized_images=True, masks=masks)
    ref = ora.match(srcs, threshold, masks=masks, debug=True)
    for i in range(n_maps):
        q = ref.quantized(i)
        assert qimgs[i].shape == q.shape
        bad = np.argwhere(qimgs[i] != q)
        assert bad.size == 0, "quantized map %d differs at %d px, first %r: got %d want %d" % (
            i, len(bad), tuple(bad[0]), qimgs[i][tuple(bad[0])], q[tuple(bad[0])])
        if masks is None:
            assert np.array_equal(det.debugFetch(K.DBG_QUANTIZED, 0, i).reshape(q.shape), q)
        lmg, lmo = det.debugFetch(K.DBG_LINMEM, 0, i), ref.linmem(i)
        assert lmg.shape == lmo.shape
        bad = np.flatnonzero(lmg != lmo)
        assert bad.size == 0, "linear memory %d differs at %d bytes, first %d" % (i, bad.size, bad[0])
    return got, ref


def test_frame_side_fixture_golden(fixture_frame, golden):
    """Quantized maps + linear memories on the reference's own frame, vs oracle AND the golden hashes."""
    bgr, depth = fixture_frame
    det, ora = make_pair()
    _check_frame_side(det, ora, [bgr, depth])
    q0 = det.debugFetch(K.DBG_QUANTIZED, 0, 0)
    q1 = det.debugFetch(K.DBG_QUANTIZED, 0, 2)
    assert sha(q0) == golden["G2G3_L0_T5"]["quantized_sha1"] == "f45507458cb4ef60bbf11099074023f5ddf0bc5d"
    assert sha(q1) == golden["G4_L1_T8"]["quantized_sha1"]
    # default table = the circular SIMILARITY_LUT: the survey's G3/G4 linear-memory hashes
    assert sha(det.debugFetch(K.DBG_LINMEM, 0, 0)) == golden["G2G3_L0_T5"]["lut1"]["linmem_sha1"] == "89d3edac27207763163c79a2eb205f6f0750ca75"
    assert sha(det.debugFetch(K.DBG_LINMEM, 0, 2)) == golden["G4_L1_T8"]["lut1"]["linmem_sha1"] == "b3fd1e4644625c962aa0d22da8a22f751270e4f6"
    assert sha(det.debugFetch(K.DBG_MAGNITUDE, 0, 0)) == golden["G2G3_L0_T5"]["magnitude_sha1"]
    assert sha(det.debugFetch(K.DBG_DN_INDICES, 0, 1)) == golden["G5_dn_indices"]["sha1"]
    # the linear variant (lmb200_config.similarity_lut = LMB200_SIMLUT_LINEAR) stays selectable
    det2, ora2 = make_pair(sim_lut=O.similarity_lut(0))
    det2.match([bgr, depth], 80.0)
    assert sha(det2.debugFetch(K.DBG_LINMEM, 0, 0)) == golden["G2G3_L0_T5"]["lut0"]["linmem_sha1"]
    assert sha(det2.debugFetch(K.DBG_LINMEM, 0, 2)) == golden["G4_L1_T8"]["lut0"]["linmem_sha1"]


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_frame_side_synthetic(idx):
    bgr, depth = synth.make_frame(idx)
    det, ora = make_pair()
    _check_frame_side(det, ora, [bgr, depth])


@pytest.mark.parametrize("rows,cols,T", [(240, 320, (4, 8)), (256, 384, (4, 8, 16)), (160, 240, (5, 8)), (96, 112, (2, 8)), (400, 600, (5, 4))])
def test_frame_side_ragged_geometry(rows, cols, T):
    """Sizes that are not multiples of the kernel tiles; W%4 != 0 linear memories (scalar store path)."""
    bgr, depth = synth.make_frame(3, rows, cols, n_shapes=15)
    det, ora = make_pair(T=T)
    _check_frame_side(det, ora, [bgr, depth], n_maps=2 * len(T))


def test_size_not_divisible_is_rejected():
    det, ora = make_pair()
    bgr, depth = synth.make_frame(0, 480, 636)
    with pytest.raises(lm.LinemodError) as e:
        det.match([bgr, depth], 80.0)
    assert e.value.code == K.E_SIZE
    with pytest.raises(ValueError):
        ora.match([bgr, depth], 80.0)
    with pytest.raises(lm.LinemodError) as e:
        det.match([bgr], 80.0)
    assert e.value.code == K.E_SOURCES


def test_masks_in_match():
    bgr, depth = synth.make_frame(4)
    det, ora = make_pair()
    m0 = synth.planted_masks(1, seed=3)[0]
    m1 = synth.planted_masks(1, seed=4)[0]
    add_random(det, ora, 50)
    got, ref = _check_frame_side(det, ora, [bgr, depth], masks=[m0, m1])
    assert_same_matches(got, ref.matches(0), "masked")


@pytest.fixture(scope="module")
def cfg2_small():
    """Config-2-shaped workload at 400 templates: 40 planted (oracle-extracted) + 360 random."""
    bgr, depth = synth.make_frame(0)
    det, ora = make_pair()
    n = add_planted_from_oracle(det, ora, [bgr, depth], synth.object_masks(0) + synth.planted_masks(40))
    assert n >= 20
    add_random(det, ora, 400 - n)
    return det, ora, bgr, depth


@pytest.mark.parametrize("threshold", [80.0, 55.0, 91.5, 30.0])
def test_match_parity(cfg2_small, threshold):
    det, ora, bgr, depth = cfg2_small
    got = det.match([bgr, depth], threshold)
    ref = ora.match([bgr, depth], threshold, threads=8, debug=True)
    assert_same_matches(det.debugFetch(K.DBG_COARSE, 0), ref.matches(2), "coarse candidates thr=%g" % threshold)
    assert_same_matches(det.debugFetch(K.DBG_UNSORTED, 0), ref.matches(1), "generation order thr=%g" % threshold)
    assert_same_matches(got, ref.matches(0), "final thr=%g" % threshold)
    if threshold <= 80:
        assert len(got) > 0


def test_match_other_frame_and_class_filter(cfg2_small):
    det, ora, _, _ = cfg2_small
    bgr, depth = synth.make_frame(5)
    for ids in ((), ("rand",), ("planted",), ("rand", "planted"), ("nope", "planted")):
        got = det.match([bgr, depth], 60.0, class_ids=ids)
        ref = ora.match([bgr, depth], 60.0, class_ids=ids, threads=8)
        assert_same_matches(got, ref.matches(0), "class filter %r" % (ids,))


def test_batch_and_resident_equal_single(cfg2_small):
    det, ora, _, _ = cfg2_small
    frames = [list(synth.make_frame(i)) for i in range(11)]
    singles = [det.match(f, 70.0) for f in frames]
    batch = det.matchBatch(frames, 70.0)
    for i, (a, b) in enumerate(zip(batch, singles)):
        assert_same_matches(a, b, "batch frame %d" % i)
    assert_same_matches(batch[3], ora.match(frames[3], 70.0, threads=8).matches(0), "batch vs oracle")
    det.uploadFrames(frames[:6], first_slot=2)
    det.matchResident(2, 6, 70.0)
    res = det.fetchResident(2, 6)
    for i in range(6):
        assert_same_matches(res[i], singles[i], "resident frame %d" % i)


def test_candidate_overflow_grows():
    """A tiny candidate store must grow transparently and still give the oracle's list."""
    bgr, depth = synth.make_frame(0)
    det, ora = make_pair(candidate_capacity=64, max_batch=4)
    add_planted_from_oracle(det, ora, [bgr, depth], synth.planted_masks(12))
    add_random(det, ora, 100)
    got = det.match([bgr, depth], 20.0)
    ref = ora.match([bgr, depth], 20.0, threads=8)
    assert len(got) > 64
    assert_same_matches(got, ref.matches(0), "overflow")
    frames = [list(synth.make_frame(i)) for i in range(5)]
    det2, ora2 = make_pair(candidate_capacity=64, max_batch=4)
    add_planted_from_oracle(det2, ora2, [bgr, depth], synth.planted_masks(12))
    add_random(det2, ora2, 100)
    for i, b in enumerate(det2.matchBatch(frames, 20.0)):
        assert_same_matches(b, ora2.match(frames[i], 20.0, threads=8).matches(0), "overflow batch %d" % i)


def test_color_only_T2_8(fixture_frame):
    """The reference's shipped default wiring: {ColorGradient}, T={2,8} (HighLevelLinemod.cpp:36-43)."""
    bgr, depth = fixture_frame
    det, ora = make_pair(modalities=("cg",), T=(2, 8))
    add_planted_from_oracle(det, ora, [bgr], synth.planted_masks(30, seed=11))
    add_random(det, ora, 150, n_modalities=1)
    for thr in (80.0, 50.0):
        got, ref = _check_frame_side(det, ora, [bgr], thr, n_maps=2)
        assert_same_matches(got, ref.matches(0), "CG-only thr=%g" % thr)


def test_three_level_pyramid():
    """Config-5-shaped pyramid (T={4,8,16}, features 63/31/15) at a size the oracle does quickly."""
    bgr, depth = synth.make_frame(6, 512, 768)
    det, ora = make_pair(T=(4, 8, 16))
    n = add_planted_from_oracle(det, ora, [bgr, depth], synth.object_masks(6, 512, 768))
    assert n > 5
    add_random(det, ora, 200, levels=3, wh_range=(40, 120))
    for thr in (80.0, 50.0):
        got, ref = _check_frame_side(det, ora, [bgr, depth], thr, n_maps=6)
        assert_same_matches(det.debugFetch(K.DBG_UNSORTED, 0), ref.matches(1), "3-level generation order")
        assert_same_matches(got, ref.matches(0), "3-level thr=%g" % thr)


def test_empty_and_degenerate_templates():
    bgr, depth = synth.make_frame(0)
    det, ora = make_pair()
    assert len(det.match([bgr, depth], 50.0)) == 0          # no templates at all
    rng = np.random.default_rng(5)
    weird = []
    for k in range(40):
        tp = synth.random_template_pyramid(rng, 2, 2)
        kind = k % 5
        for t in tp:
            f = t["features"]
            if kind == 0:      # features far outside the bbox (guarded rows)
                f[:, 0] += 300 >> t["pyramid_level"]
            elif kind == 1:    # template wider than the image minus 16T (N4: clamp makes max_x < border)
                t["width"] = 600 >> t["pyramid_level"]; t["height"] = 440 >> t["pyramid_level"]
            elif kind == 2:    # negative coordinates and an empty feature list at the coarse level
                f[0, 0] = -3
                if t["pyramid_level"] == 1:
                    t["features"] = f[:0]
            elif kind == 3:    # bbox larger than the whole image (P <= 0)
                t["width"] = 700 >> t["pyramid_level"]; t["height"] = 500 >> t["pyramid_level"]
            else:              # features beyond the image
                f[:, 1] += 470 >> t["pyramid_level"]
        weird.append(tp)
    for tp in weird:
        det.addSyntheticTemplate(tp, "weird"); ora.add_synthetic(tp, "weird")
    add_random(det, ora, 20)
    for thr in (10.0, 40.0):
        got = det.match([bgr, depth], thr)
        ref = ora.match([bgr, depth], thr, debug=True)
        assert_same_matches(det.debugFetch(K.DBG_COARSE, 0), ref.matches(2), "degenerate coarse thr=%g" % thr)
        assert_same_matches(got, ref.matches(0), "degenerate thr=%g" % thr)


def test_add_template_parity(fixture_frame):
    """Product addTemplate (GPU quantisation + host selection) == oracle addTemplate, incl. failures."""
    for name, (bgr, depth) in (("fixture", fixture_frame), ("synthetic", synth.make_frame(2))):
        det, ora = make_pair()
        masks = synth.planted_masks(24, seed=21) + [None, np.zeros((480, 640), np.uint8)]
        ok = 0
        for i, m in enumerate(masks):
            tid, bb = det.addTemplate([bgr, depth], "obj", m)
            otid, obb = ora.add_template([bgr, depth], "obj", m)
            assert tid == otid, "%s mask %d: template id %d vs oracle %d" % (name, i, tid, otid)
            if tid >= 0:
                ok += 1
                assert tuple(bb) == tuple(obb)
                got = det.getTemplates("obj", tid)
                want = O.decode_pyramid(ora.get_template_flat("obj", tid))
                for a, b in zip(got, want):
                    assert (a["width"], a["height"], a["pyramid_level"]) == (b["width"], b["height"], b["pyramid_level"])
                    assert np.array_equal(a["features"], b["features"])
        assert ok >= 10 and det.numTemplates("obj") == ora.num_templates("obj") == ok
        assert_same_matches(det.match([bgr, depth], 75.0), ora.match([bgr, depth], 75.0).matches(0), "after addTemplate")


def test_add_templates_bulk_parity():
    """lmb200_add_templates (batched GPU quantisation, threaded host selection, two chunks) == one oracle addTemplate
    per view, including failed views, the ids after them, and the matches the resulting set produces."""
    views, masks = [], []
    for i in range(3):
        bgr, depth = synth.make_frame(20 + i)
        for m in synth.object_masks(20 + i)[:12] + synth.planted_masks(6, seed=40 + i):
            views.append([bgr, depth]); masks.append(m)
        views.append([bgr, depth]); masks.append(np.zeros((480, 640), np.uint8))        # fails
        views.append([bgr, depth]); masks.append(None)                                   # whole image
    assert len(views) > 32                                                                # more than one chunk
    for mods, T in ((("cg", "dn"), (5, 8)), (("cg",), (2, 8))):
        det, ora = make_pair(modalities=mods, T=T)
        v = [vw[:len(mods)] for vw in views]
        res = det.addTemplates(v, "obj", masks)
        ok = 0
        for i, (tid, bb) in enumerate(res):
            otid, obb = ora.add_template(v[i], "obj", masks[i])
            assert tid == otid, "view %d: template id %d vs oracle %d" % (i, tid, otid)
            if tid < 0:
                continue
            ok += 1
            assert tuple(bb) == tuple(obb)
            want = O.decode_pyramid(ora.get_template_flat("obj", tid))
            for a, b in zip(det.getTemplates("obj", tid), want):
                assert (a["width"], a["height"], a["pyramid_level"]) == (b["width"], b["height"], b["pyramid_level"])
                assert np.array_equal(a["features"], b["features"])
        assert ok >= 20 and ok < len(views) and det.numTemplates("obj") == ora.num_templates("obj") == ok
        # a second bulk call continues the id sequence, exactly like further addTemplate calls
        more = det.addTemplates(v[:3], "obj", masks[:3])
        assert [t for t, _ in more if t >= 0] == list(range(ok, ok + sum(1 for t, _ in more if t >= 0)))
        for i in range(3):
            ora.add_template(v[i], "obj", masks[i])
        src = views[0][:len(mods)]
        assert_same_matches(det.match(src, 75.0), ora.match(src, 75.0).matches(0), "after bulk addTemplates")


def test_config2_full_size():
    """BASELINE config 2: one 640x480 frame vs 3 000 templates (300 planted + 2 700 random), thresholds 80 and 57."""
    bgr, depth = synth.make_frame(0)
    det, ora = make_pair()
    masks = synth.object_masks(0) + synth.planted_masks(300, seed=17)
    n = add_planted_from_oracle(det, ora, [bgr, depth], masks)
    assert n >= 100
    add_random(det, ora, 3000 - n)
    assert det.numTemplates() == 3000
    for thr in (80.0, 57.0):
        got = det.match([bgr, depth], thr)
        ref = ora.match([bgr, depth], thr, threads=8)
        assert_same_matches(got, ref.matches(0), "config 2 thr=%g" % thr)
    # size-independent properties: idempotence and order
    again = det.match([bgr, depth], 57.0)
    assert_same_matches(again, got, "idempotence")
    s = got.similarity
    assert np.all(s[:-1] >= s[1:]) and np.all(s >= 57.0)


def test_config4_20000_templates_multi_class():
    """BASELINE config 4 template set (20 000 templates, 10 classes) on one GPU vs the oracle."""
    bgr, depth = synth.make_frame(1)
    det, ora = make_pair(max_batch=2)
    masks = synth.object_masks(1)
    n_planted = add_planted_from_oracle(det, ora, [bgr, depth], masks, class_id="obj00")
    per_class = 2000
    tps = synth.random_templates(20000 - n_planted, seed=123)
    k = 0
    for c in range(10):
        cid = "obj%02d" % c
        want = per_class - (n_planted if c == 0 else 0)
        for tp in tps[k:k + want]:
            det.addSyntheticTemplate(tp, cid); ora.add_synthetic(tp, cid)
        k += want
    assert det.numTemplates() == 20000 and det.numClasses() == 10
    for thr, ids in ((80.0, ()), (58.0, ("obj07", "obj00", "obj03"))):
        got = det.match([bgr, depth], thr, class_ids=ids)
        ref = ora.match([bgr, depth], thr, class_ids=ids, threads=16)
        assert_same_matches(got, ref.matches(0), "config 4 thr=%g ids=%r" % (thr, ids))
    assert len(got) > 0


def test_config5_kinect_v2_size_three_levels():
    """BASELINE config 5 geometry: 1920x1080 padded to 1920x1088 (linearize needs rows % T == 0 at every level),
    T={4,8,16}, features 63/31/15; 2 000 templates here (the 10 000-template run is bench material)."""
    rows, cols = 1088, 1920
    bgr, depth = synth.make_frame(7, 1080, cols, n_shapes=60)
    bgr = np.concatenate([bgr, np.zeros((8, cols, 3), np.uint8)], 0)
    depth = np.concatenate([depth, np.zeros((8, cols), np.uint16)], 0)
    det, ora = make_pair(T=(4, 8, 16), max_batch=2)
    masks = [np.concatenate([m, np.zeros((8, cols), np.uint8)], 0) for m in synth.object_masks(7, 1080, cols, n_shapes=60, min_px=4000)]
    n = add_planted_from_oracle(det, ora, [bgr, depth], masks[:12])
    assert n >= 4
    add_random(det, ora, 2000 - n, levels=3, wh_range=(80, 300))
    for thr in (80.0, 60.0):
        got, ref = _check_frame_side(det, ora, [bgr, depth], thr, n_maps=6)
        assert_same_matches(det.debugFetch(K.DBG_COARSE, 0), ref.matches(2), "config 5 coarse thr=%g" % thr)
        assert_same_matches(got, ref.matches(0), "config 5 thr=%g" % thr)
    assert len(got) > 0


def test_config3_streamed_batch_of_64():
    """BASELINE config 3: 64 frames streamed through lmb200_match_batch vs 3 000 templates; every frame's list
    must equal the single-frame call, and sampled frames must equal the oracle."""
    det, ora = make_pair(max_batch=24)
    bgr0, depth0 = synth.make_frame(0)
    n = add_planted_from_oracle(det, ora, [bgr0, depth0], synth.object_masks(0) + synth.planted_masks(100, seed=17))
    add_random(det, ora, 3000 - n)
    frames = [list(synth.make_frame(i)) for i in range(64)]
    batch = det.matchBatch(frames, 80.0)
    assert len(batch) == 64
    for i in (0, 1, 31, 63):
        assert_same_matches(batch[i], ora.match(frames[i], 80.0, threads=16).matches(0), "config 3 frame %d vs oracle" % i)
    for i in range(0, 64, 7):
        assert_same_matches(batch[i], det.match(frames[i], 80.0), "config 3 frame %d vs single" % i)
    assert len(batch[0]) > 0


def test_config1_reference_frame_and_model_templates(fixture_frame):
    """BASELINE config 1: the reference's frame (benchmark/img0.png + depth0.png) vs the 1 950 lagergehaeuse templates
    (13 viewpoints x 15 radii x 10 in-plane rotations, tests/golden/make_config1_templates.py), {CG,DN}, T={5,8},
    threshold 80 (linemod_settings.yml:29) and 70.  Templates are loaded from the reference's file layout
    through the product's own reader."""
    import os
    bgr, depth = fixture_frame
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lagergehaeuse_templates.yml.gz")
    det = lm.Detector.read(path)
    assert det.classIds() == ["lagergehaeuse.ply"] and det.numTemplates() == 1950
    ora = O.Detector([dict(type=O.CG), dict(type=O.DN)], [5, 8], sim_lut=det.getSimilarityLut(), normal_lut=det.getNormalLut())
    for t in range(1950):
        ora.add_synthetic(det.getTemplates("lagergehaeuse.ply", t), "lagergehaeuse.ply")
    for thr in (80.0, 70.0):
        got = det.match([bgr, depth], thr, class_ids=["lagergehaeuse.ply"])
        ref = ora.match([bgr, depth], thr, class_ids=["lagergehaeuse.ply"], threads=16)
        assert_same_matches(got, ref.matches(0), "config 1 thr=%g" % thr)
    assert len(got) > 1000


def test_pipelined_submit_collect(cfg2_small):
    """Two batches in flight (submit k+1 before collecting k): every frame's list equals the blocking single-frame call."""
    det, ora, _, _ = cfg2_small
    fa = [list(synth.make_frame(i)) for i in range(0, 9)]
    fb = [list(synth.make_frame(i)) for i in range(9, 20)]
    pa, pb = det.prepareBatch(fa, cap=20000), det.prepareBatch(fb, cap=20000)
    ta = det.submitPrepared(pa, 65.0)
    tb = det.submitPrepared(pb, 65.0)
    with pytest.raises(lm.LinemodError):
        det.submitPrepared(pa, 65.0)              # at most two in flight
    def lists(prep):
        offs = prep["offs"]
        return [prep["out"][offs[i]:offs[i + 1]].copy().view(np.recarray) for i in range(prep["n"])]
    det.collectPrepared(pa, ta)
    ra = lists(pa)
    tc = det.submitPrepared(pa, 65.0)             # ticket slot is free again; batch c pipelines behind b
    det.collectPrepared(pb, tb)
    rb = lists(pb)
    det.collectPrepared(pa, tc)
    rc = lists(pa)
    for i, f in enumerate(fa):
        want = det.match(f, 65.0)
        assert_same_matches(ra[i], want, "pipelined batch a frame %d" % i)
        assert_same_matches(rc[i], want, "pipelined batch c frame %d" % i)
    for i, f in enumerate(fb):
        assert_same_matches(rb[i], det.match(f, 65.0), "pipelined batch b frame %d" % i)
    assert_same_matches(rb[4], ora.match(fb[4], 65.0, threads=8).matches(0), "pipelined vs oracle")


def test_pipelined_overflow_regrows_both_tickets():
    bgr, depth = synth.make_frame(0)
    det, ora = make_pair(candidate_capacity=64, max_batch=24)
    add_planted_from_oracle(det, ora, [bgr, depth], synth.object_masks(0))
    add_random(det, ora, 100)
    fa = [list(synth.make_frame(i)) for i in range(3)]
    fb = [list(synth.make_frame(i)) for i in range(3, 7)]
    pa, pb = det.prepareBatch(fa, cap=200000), det.prepareBatch(fb, cap=200000)
    ta = det.submitPrepared(pa, 20.0)
    tb = det.submitPrepared(pb, 20.0)
    det.collectPrepared(pa, ta)      # overflows the 64-entry store: grows, reruns a; b was computed into the old store
    det.collectPrepared(pb, tb)      # ... so b reruns too
    for prep, fr in ((pa, fa), (pb, fb)):
        offs = prep["offs"]
        for i, f in enumerate(fr):
            got = prep["out"][offs[i]:offs[i + 1]].view(np.recarray)
            assert len(got) > 64
            assert_same_matches(got, ora.match(f, 20.0, threads=8).matches(0), "overflow pipelined frame %d" % i)


def test_single_level_pyramid_wide_sums():
    """One pyramid level with two modalities: 63+63 features at the (only = coarsest) level, raw scores up to 504 —
    the u16-widening instantiation of the coarse kernel — and no refinement (scores keep upstream's +0.5f)."""
    bgr, depth = synth.make_frame(8)
    det, ora = make_pair(T=(8,))
    n = add_planted_from_oracle(det, ora, [bgr, depth], synth.object_masks(8))
    assert n >= 5
    add_random(det, ora, 150, levels=1)
    for thr in (80.0, 45.0):
        got, ref = _check_frame_side(det, ora, [bgr, depth], thr, n_maps=2)
        assert_same_matches(det.debugFetch(K.DBG_COARSE, 0), ref.matches(2), "1-level coarse thr=%g" % thr)
        assert_same_matches(got, ref.matches(0), "1-level thr=%g" % thr)
    assert len(got) > 0 and got.similarity.max() > 100.0      # 100 % + 0.5


def test_three_modalities():
    """More than two modalities (ColorGradient, DepthNormal, ColorGradient): the per-modality paths of every kernel."""
    bgr, depth = synth.make_frame(9)
    bgr2 = np.ascontiguousarray(bgr[:, ::-1])
    mods = ("cg", "dn", "cg")
    det, ora = make_pair(modalities=mods)
    srcs = [bgr, depth, bgr2]
    n = add_planted_from_oracle(det, ora, srcs, synth.object_masks(9))
    add_random(det, ora, 150, n_modalities=3)
    for thr in (75.0, 50.0):
        got, ref = _check_frame_side(det, ora, srcs, thr, n_maps=6)
        assert_same_matches(det.debugFetch(K.DBG_UNSORTED, 0), ref.matches(1), "3-modality generation order thr=%g" % thr)
        assert_same_matches(got, ref.matches(0), "3-modality thr=%g" % thr)
    frames = [[f[0], f[1], np.ascontiguousarray(f[0][:, ::-1])] for f in (synth.make_frame(i) for i in range(9, 14))]
    for i, b in enumerate(det.matchBatch(frames, 60.0)):
        assert_same_matches(b, ora.match(frames[i], 60.0, threads=8).matches(0), "3-modality batch frame %d" % i)


def test_upload_templates_eagerly():
    """lmb200_upload_templates before the first match and again after the set changed: same matches as the
    on-demand upload, i.e. as the oracle."""
    bgr, depth = synth.make_frame(3)
    det, ora = make_pair()
    add_planted_from_oracle(det, ora, [bgr, depth], synth.object_masks(3)[:10])
    add_random(det, ora, 60)
    det.uploadTemplates()                                   # no frame size known yet: tables only
    assert_same_matches(det.match([bgr, depth], 70.0), ora.match([bgr, depth], 70.0).matches(0), "eager upload")
    add_random(det, ora, 40, class_id="more", seed=5)
    det.uploadTemplates()                                   # frame size known: tables + plan
    assert_same_matches(det.match([bgr, depth], 70.0), ora.match([bgr, depth], 70.0).matches(0), "eager upload after change")


def test_non_default_modality_parameters():
    """ColorGradient(weak, num_features, strong) and DepthNormal(distance, difference, num_features, extract) away from
    their defaults: parameter plumbing end to end, fewer features per template, and the 64-bit instantiation of the
    DepthNormal kernel (difference_threshold > 200)."""
    cgp = dict(weak_threshold=20.0, num_features=40, strong_threshold=30.0)
    dnp = dict(distance_threshold=1500, difference_threshold=300, num_features=50, extract_threshold=3)
    det = lm.Detector([lm.ColorGradient(**cgp), lm.DepthNormal(**dnp)], [5, 8])
    ora = O.Detector([dict(type=O.CG, **cgp), dict(type=O.DN, **dnp)], [5, 8],
                     sim_lut=det.getSimilarityLut(), normal_lut=det.getNormalLut())
    bgr, depth = synth.make_frame(11)
    ok = 0
    for m in synth.object_masks(11)[:10]:
        tid, bb = det.addTemplate([bgr, depth], "obj", m)
        otid, obb = ora.add_template([bgr, depth], "obj", m)
        assert tid == otid
        if tid >= 0:
            ok += 1
            assert tuple(bb) == tuple(obb)
            for a, b in zip(det.getTemplates("obj", tid), O.decode_pyramid(ora.get_template_flat("obj", tid))):
                assert np.array_equal(a["features"], b["features"])
    assert ok >= 3
    assert len(det.getTemplates("obj", 0)[0]["features"]) == 40 and len(det.getTemplates("obj", 0)[1]["features"]) == 50
    for thr in (80.0, 60.0):
        got, ref = _check_frame_side(det, ora, [bgr, depth], thr, n_maps=4)
        assert_same_matches(got, ref.matches(0), "non-default parameters thr=%g" % thr)
    assert len(got) > 0


def _check_similarity_maps(det, ora, ref, picks, what):
    """SURVEY 8d parity gate: full u16 coarse similarity map (similarity() per modality + addSimilarities()) of sampled
    templates — production kernel with the early exit disabled vs the oracle."""
    for cid, tid in picks:
        got = det.similarityMap(cid, tid)
        want = ora.similarity_map(ref, cid, tid)
        assert got.shape == want.shape, "%s: map shape %r vs %r" % (what, got.shape, want.shape)
        bad = np.flatnonzero(got != want)
        assert bad.size == 0, "%s: similarity map of %s/%d differs at %d positions, first %d (%d vs %d)" % (
            what, cid, tid, bad.size, bad[0], got[bad[0]], want[bad[0]])


def _sample_templates(det, n, seed):
    rng = np.random.default_rng(seed)
    allt = [(cid, t) for cid in det.classIds() for t in range(det.numTemplates(cid))]
    idx = rng.choice(len(allt), size=min(n, len(allt)), replace=False)
    return [allt[i] for i in sorted(idx)]


def test_similarity_maps_config2(cfg2_small):
    det, ora, bgr, depth = cfg2_small
    det.match([bgr, depth], 80.0)
    ref = ora.match([bgr, depth], 80.0, threads=8, debug=True)
    picks = [("planted", t) for t in range(min(8, det.numTemplates("planted")))] + _sample_templates(det, 32, 1)
    _check_similarity_maps(det, ora, ref, picks, "config 2")
    # the planted templates must peak at 4 * nf (a perfect match on their own frame)
    m = det.similarityMap("planted", 0)
    assert int(m.max()) == 4 * sum(len(t["features"]) for t in det.getTemplates("planted", 0)[2:4])


def test_similarity_maps_config1(fixture_frame):
    import os
    bgr, depth = fixture_frame
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lagergehaeuse_templates.yml.gz")
    det = lm.Detector.read(path)
    ora = O.Detector([dict(type=O.CG), dict(type=O.DN)], [5, 8], sim_lut=det.getSimilarityLut(), normal_lut=det.getNormalLut())
    for t in range(1950):
        ora.add_synthetic(det.getTemplates("lagergehaeuse.ply", t), "lagergehaeuse.ply")
    det.match([bgr, depth], 80.0)
    ref = ora.match([bgr, depth], 80.0, threads=16, debug=True)
    _check_similarity_maps(det, ora, ref, _sample_templates(det, 32, 2), "config 1")


def test_similarity_maps_cg_only_and_wide_sums():
    """{CG} T={2,8} (the shipped yml default) and a single-level pyramid whose sums need u16 (WIDE instantiation)."""
    bgr, depth = synth.make_frame(2)
    det, ora = make_pair(modalities=("cg",), T=(2, 8))
    add_planted_from_oracle(det, ora, [bgr], synth.object_masks(2))
    add_random(det, ora, 60, n_modalities=1)
    det.match([bgr], 70.0)
    ref = ora.match([bgr], 70.0, debug=True)
    _check_similarity_maps(det, ora, ref, _sample_templates(det, 32, 3), "{CG} T={2,8}")
    det, ora = make_pair(T=(8,))
    add_random(det, ora, 40, levels=1)
    det.match([bgr, depth], 60.0)
    ref = ora.match([bgr, depth], 60.0, debug=True)
    _check_similarity_maps(det, ora, ref, _sample_templates(det, 32, 4), "single level, u16 sums")


def test_config5_full_10000_templates():
    """BASELINE config 5 at its full size: 1920x1088, T={4,8,16}, 10 000 templates; final lists + sampled similarity maps."""
    rows, cols = 1088, 1920
    bgr, depth = synth.make_frame(7, 1080, cols, n_shapes=60)
    bgr = np.concatenate([bgr, np.zeros((8, cols, 3), np.uint8)], 0)
    depth = np.concatenate([depth, np.zeros((8, cols), np.uint16)], 0)
    det, ora = make_pair(T=(4, 8, 16), max_batch=2)
    masks = [np.concatenate([m, np.zeros((8, cols), np.uint8)], 0) for m in synth.object_masks(7, 1080, cols, n_shapes=60, min_px=4000)]
    n = add_planted_from_oracle(det, ora, [bgr, depth], masks[:12])
    add_random(det, ora, 10000 - n, levels=3, wh_range=(80, 300))
    assert det.numTemplates() == 10000
    got = det.match([bgr, depth], 80.0)
    ref = ora.match([bgr, depth], 80.0, threads=16, debug=True)
    assert_same_matches(got, ref.matches(0), "config 5, 10 000 templates")
    assert len(got) > 0
    _check_similarity_maps(det, ora, ref, _sample_templates(det, 32, 5), "config 5")


def test_resident_ranges_survive_store_growth():
    """several lmb200_match_resident ranges in flight; the first fetch overflows the candidate store and grows
    it — the other ranges (computed into the old stores) must be re-matched at their own thresholds, not read stale."""
    det, ora = make_pair(candidate_capacity=64, max_batch=8)
    bgr, depth = synth.make_frame(0)
    add_planted_from_oracle(det, ora, [bgr, depth], synth.object_masks(0))
    add_random(det, ora, 100)
    frames = [list(synth.make_frame(i)) for i in range(6)]
    det.uploadFrames(frames, 0)
    det.matchResident(0, 2, 20.0)     # overflows 64 candidates
    det.matchResident(2, 2, 85.0)     # fits
    det.matchResident(4, 2, 25.0)     # overflows
    r0 = det.fetchResident(0, 2, cap=400000)     # grows the stores under ranges 2..5
    r1 = det.fetchResident(2, 2, cap=400000)
    r2 = det.fetchResident(4, 2, cap=400000)
    for res, first, thr in ((r0, 0, 20.0), (r1, 2, 85.0), (r2, 4, 25.0)):
        for i in range(2):
            want = ora.match(frames[first + i], thr, threads=8).matches(0)
            assert_same_matches(res[i], want, "resident range at slot %d thr=%g" % (first + i, thr))


def test_generic_frame_side_fallback_still_exact(monkeypatch):
    """Round 1's generic frame-side kernels stay in the library as the fallback for geometries the round-2 kernels do
    not cover (T outside {2,4,5,8,16}, coarsest W % 8 != 0); LMB200_GENERIC_FRAME=1 forces them everywhere."""
    monkeypatch.setenv("LMB200_GENERIC_FRAME", "1")
    bgr, depth = synth.make_frame(1)
    det, ora = make_pair()
    _check_frame_side(det, ora, [bgr, depth])
    monkeypatch.delenv("LMB200_GENERIC_FRAME")
    det2, ora2 = make_pair()
    _check_frame_side(det2, ora2, [bgr, depth])


@pytest.mark.parametrize("T,rows,cols", [((3, 6), 480, 672), ((7,), 448, 672), ((5, 10), 480, 640)])
def test_frame_side_uncovered_T_uses_fallback(T, rows, cols):
    bgr, depth = synth.make_frame(2, rows, cols, n_shapes=20)
    det, ora = make_pair(T=T)
    _check_frame_side(det, ora, [bgr, depth], n_maps=2 * len(T))


def test_dropin_header_runs_reference_call_sites_on_gpu(tmp_path):
    """The verbatim reference call sites (tests/cpp_dropin_check.cpp over include/lmb200_opencv.hpp) with a device present:
    addTemplate on a synthetic view, match finds it again at similarity 100, getTemplates feeds the hull points."""
    import subprocess, os
    from test_capi_host import _build_dropin
    exe = _build_dropin(tmp_path)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, env=dict(os.environ, LMB200_QUIET="1"))
    assert r.returncode == 0 and "DROPIN_OK" in r.stdout, r.stdout + r.stderr
    assert r.stdout.count("GPU: wiring") == 2 and "no GPU" not in r.stdout, r.stdout


def test_postmatch_color_check_equals_oracle_and_cv2(fixture_frame):
    """lmb200_postmatch_color (HSV inRange bit mask + per-match hull / fillPoly mask counts on the GPU) against
    oracle/postmatch.py (pinned to cv2 by tests/test_oracle_postmatch.py) and against cv2 itself, on synthetic matches and
    on the reference's own frame with the lagergehaeuse templates."""
    import os
    from oracle import postmatch as PM
    cv2 = pytest.importorskip("cv2")
    cases = []
    bgr, depth = synth.make_frame(4)
    det, ora = make_pair()
    add_planted_from_oracle(det, ora, [bgr, depth], synth.object_masks(4)[:10], class_id="obj")
    add_random(det, ora, 30)
    cases.append((det, bgr, depth, 70.0, (0, 0, 90), (180, 255, 255)))
    fb, fd = fixture_frame
    det1 = lm.Detector.read(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lagergehaeuse_templates.yml.gz"))
    cases.append((det1, fb, fd, 88.0, (0, 30, 40), (40, 255, 255)))       # HSV range of the reference's model file shape
    for d, b, dp, thr, lower, upper in cases:
        res = d.match([b, dp], thr)
        assert len(res) >= 5
        res = res[:200]
        inside, total = d.postmatchColor(res, lower, upper)
        hue_ref = cv2.inRange(cv2.cvtColor(b, cv2.COLOR_BGR2HSV), lower, upper)
        ids = d.classIds()
        checked = 0
        for i in range(len(res)):
            tp = d.getTemplates(ids[int(res.class_index[i])], int(res.template_id[i]))
            M = len(d.getModalities())
            x, y = int(res.x[i]), int(res.y[i])
            pts = np.array([(fx + x, fy + y) for m in range(M) for fx, fy, _ in tp[m]["features"]], np.int64)
            if pts[:, 0].min() < 0 or pts[:, 1].min() < 0 or pts[:, 0].max() >= b.shape[1] or pts[:, 1].max() >= b.shape[0]:
                assert inside[i] == -1 and total[i] == -1
                continue
            mask = PM.template_mask(tp, M, x, y, b.shape[0], b.shape[1])
            assert total[i] == int(np.count_nonzero(mask)), "match %d: mask area %d vs oracle %d" % (i, total[i], np.count_nonzero(mask))
            assert inside[i] == int(np.count_nonzero(hue_ref & mask)), "match %d" % i
            if i < 20:   # and the reference's own formulation with real cv2
                ref_mask = np.zeros(b.shape[:2], np.uint8)
                cv2.fillPoly(ref_mask, [cv2.convexHull(pts.astype(np.int32))[:, 0, :]], 255)
                assert total[i] == cv2.countNonZero(ref_mask) and inside[i] == cv2.countNonZero(cv2.bitwise_and(hue_ref, ref_mask))
            checked += 1
        assert checked >= 5
    with pytest.raises(lm.LinemodError):
        lm.getDefaultLINEMOD().postmatchColor(res[:1], (0, 0, 0), (1, 1, 1))   # no frame resident


@pytest.mark.parametrize("world,frames,max_per_rank,interleaved", [(1, 3, 200, False), (2, 5, 300, True), (3, 4, 150, True),
                                                                   (8, 16, 120, True), (8, 2, 500, True), (8, 3, 1300, True), (4, 6, 100, False)])
def test_device_epilogue_of_the_sharded_step(world, frames, max_per_rank, interleaved):
    """shard_epilogue_kernel (generation-order merge + libstdc++'s std::sort + std::unique on the device) against the host
    epilogue, on one GPU: ties on (similarity, template_id) everywhere, exact duplicates, empty frames, 1-8 ranks."""
    L = lm.capi.lib()
    MR = lm.capi.MatchRec
    rng = np.random.default_rng(100 * world + frames)
    ntpl = 64 * world
    G, gcap, pos, g_class, g_tid, want = _epilogue_case(rng, world, frames, max_per_rank, ntpl, interleaved, sims=6, tid_mod=7)
    total = sum(n for n, _ in want)
    out = np.zeros(max(1, total), lm.MATCH_DTYPE)
    hdr = np.zeros((2 * frames, 4), np.int32)
    rank = world - 1
    rc = L.lmb200_debug_shard_epilogue(G.ctypes.data, world, rank, frames, gcap, pos.ctypes.data if pos is not None else None,
                                       g_class.ctypes.data, g_tid.ctypes.data, ntpl, out.ctypes.data_as(C.POINTER(MR)), len(out), hdr.ctypes.data)
    assert rc == 0
    off = 0
    for f in range(frames):
        n_in, fin = want[f]
        n_final, offset, flags, n_seen = (int(v) for v in hdr[2 * f])
        if n_in > 4096:
            assert flags & 4, "frame %d: %d records must be flagged for the host path" % (f, n_in)
        else:
            assert flags == 0 and n_seen == n_in and offset == off, (f, flags, n_seen, n_in, offset, off)
            got = out[offset:offset + n_final]
            assert n_final == len(fin) and got.tobytes() == fin.tobytes(), "frame %d: device epilogue differs from std::sort + std::unique" % f
            assert tuple(int(v) for v in hdr[2 * f + 1]) == (100 * rank + f, 0, 7 * rank + f, 0)
        off += n_in


def test_device_epilogue_flags():
    """A store-overflow / too-small flag in any rank's header, and an output area that is too small, must flag the frame."""
    L = lm.capi.lib()
    MR = lm.capi.MatchRec
    rng = np.random.default_rng(5)
    G, gcap, pos, g_class, g_tid, want = _epilogue_case(rng, 2, 3, 50, 128, True, sims=4, tid_mod=5)
    G[1, 2 * 1, 1] = 2                                          # rank 1 says: my record area was too small for frame 1
    out = np.zeros(4096, lm.MATCH_DTYPE)
    hdr = np.zeros((6, 4), np.int32)
    assert L.lmb200_debug_shard_epilogue(G.ctypes.data, 2, 0, 3, gcap, pos.ctypes.data, g_class.ctypes.data, g_tid.ctypes.data, 128,
                                         out.ctypes.data_as(C.POINTER(MR)), len(out), hdr.ctypes.data) == 0
    assert hdr[2, 2] & 2 and hdr[0, 2] == 0 and hdr[4, 2] == 0
    G[1, 2 * 1, 1] = 0
    small = max(1, want[0][0])                                  # room for frame 0 only
    out = np.zeros(small, lm.MATCH_DTYPE)
    assert L.lmb200_debug_shard_epilogue(G.ctypes.data, 2, 0, 3, gcap, pos.ctypes.data, g_class.ctypes.data, g_tid.ctypes.data, 128,
                                         out.ctypes.data_as(C.POINTER(MR)), len(out), hdr.ctypes.data) == 0
    assert hdr[0, 2] == 0 and (want[1][0] == 0 or hdr[2, 2] & 8)
