"""Oracle vs the golden vectors of SURVEY.md §8c (G1..G6).

The goldens were produced with real OpenCV primitives by tests/golden/make_golden.py
(committed; runs only in the build container).  The reference itself holds no golden
vectors or tests for this path.
"""
import hashlib
import numpy as np
from oracle import oracle as O

sha = lambda a: hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()

# SURVEY.md §8c, verbatim
SURVEY = dict(
    G1="521625b4366ecc3f18d4997fe506a3eb147af327",
    G2_q="f45507458cb4ef60bbf11099074023f5ddf0bc5d", G2_mag="2cadcbdac8764da376063906e19e312caea156b4",
    G3_spread="f4562a7cde1020c053d11574372083b0b06808d4", G3_resp="39474ffa43100b98090a3d582c180b1eec929fae",
    G3_lm="89d3edac27207763163c79a2eb205f6f0750ca75",
    G4_q="2364f339f5107b94f2329ef0a3c8919d25b2ab0c", G4_mag="7eca0979dab08a982d5d1efbd2912d08b5df2d0a",
    G4_spread="a4cde94b559699646515e7c5293e75e647a06838", G4_resp="979b4c3bedc0def48910c2d638b62fc6c8acd5ed",
    G4_lm="b3fd1e4644625c962aa0d22da8a22f751270e4f6",
    G5="76413abbe0b84c80ce7e44f769b03379b0f7f98b", G6="de1dd7100335ac55c007c141c6ba5ddbf440be69",
)


def test_fixture_is_the_reference_frame(fixture_frame, golden):
    bgr, depth = fixture_frame
    assert bgr.shape == (480, 640, 3) and depth.shape == (480, 640) and depth.dtype == np.uint16
    assert sha(bgr) == golden["input"]["bgr_sha1"] and sha(bgr).startswith("ca500e5c")
    assert sha(depth) == golden["input"]["depth_sha1"] and sha(depth).startswith("b6032802")


def test_G1_label_table(golden):
    for fused in (1, 0):   # FMA and non-FMA polynomial evaluation give identical labels
        tab = O.label_table(fused)
        assert sha(tab) == SURVEY["G1"] == golden["G1_label_table"]["sha1"]
    assert np.bincount(tab.ravel(), minlength=8).tolist() == [414277, 488906, 690752, 488906, 414276, 488906, 690752, 488906]


def test_G2_G3_level0(fixture_frame, golden):
    bgr, _ = fixture_frame
    q, mag = O.cg_quantize(bgr)
    assert sha(q) == SURVEY["G2_q"] and sha(mag) == SURVEY["G2_mag"]
    assert int((q > 0).sum()) == 180903
    assert [int((q == (1 << i)).sum()) for i in range(8)] == [23827, 19053, 11751, 28911, 60258, 18080, 7554, 11469]
    sp = O.spread(q, 5)
    assert sha(sp) == SURVEY["G3_spread"]
    resp = O.response(sp, O.similarity_lut(1))       # survey hashed the circular-distance table
    assert sha(resp) == SURVEY["G3_resp"]
    lm = np.stack([O.linearize(resp[o], 5) for o in range(8)])
    assert lm.shape == (8, 25, 12288) and sha(lm) == SURVEY["G3_lm"]
    # the linear (non-circular) variant, pinned by make_golden.py's numpy restatement
    resp0 = O.response(sp, O.similarity_lut(0))
    assert sha(resp0) == golden["G2G3_L0_T5"]["lut0"]["response_sha1"]
    lm0 = np.stack([O.linearize(resp0[o], 5) for o in range(8)])
    assert sha(lm0) == golden["G2G3_L0_T5"]["lut0"]["linmem_sha1"]


def test_G4_level1(fixture_frame, golden):
    bgr, _ = fixture_frame
    q, mag = O.cg_quantize(O.pyrdown(bgr))
    assert sha(q) == SURVEY["G4_q"] and sha(mag) == SURVEY["G4_mag"] and int((q > 0).sum()) == 47304
    sp = O.spread(q, 8)
    assert sha(sp) == SURVEY["G4_spread"]
    resp = O.response(sp, O.similarity_lut(1))
    assert sha(resp) == SURVEY["G4_resp"]
    lm = np.stack([O.linearize(resp[o], 8) for o in range(8)])
    assert lm.shape == (8, 64, 1200) and sha(lm) == SURVEY["G4_lm"]
    assert sha(np.stack([O.linearize(O.response(sp, O.similarity_lut(0))[o], 8) for o in range(8)])) == \
        golden["G4_L1_T8"]["lut0"]["linmem_sha1"]


def test_G5_depth_normal_indices(fixture_frame):
    _, depth = fixture_frame
    _, idx = O.dn_quantize(depth, None, want_idx=True, median=False)
    assert sha(idx) == SURVEY["G5"]
    assert abs(float((idx[0, 5:474, 5:634] >= 0).mean()) - 0.95748) < 1e-4 and idx.max() == 19


def test_G6_similarity_lut(golden):
    circ, lin = O.similarity_lut(1), O.similarity_lut(0)
    assert sha(circ) == SURVEY["G6"] and int(circ.sum()) == 628
    assert sha(lin) == golden["G6_similarity_lut"]["linear_sha1"]
    # the linear variant: orientation 0 has an all-zero high-nibble half
    assert lin[:16].tolist() == [0, 4, 3, 4, 2, 4, 3, 4, 1, 4, 3, 4, 2, 4, 3, 4] and not lin[16:32].any()
    assert lin[32:64].tolist() == [0, 3, 4, 4, 3, 3, 4, 4, 2, 3, 4, 4, 3, 3, 4, 4] + [0, 1] * 8
    # upstream's literal table as recalled in round 2 == the circular table == the oracle's default
    import json, os
    rec = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "similarity_lut_recalled.json")))["table"]
    assert len(rec) == 256 and rec == circ.tolist()
    assert O.Detector([dict(type=O.CG)], [5, 8]).similarity_lut().tolist() == rec
