"""Shared helpers for parity tests: build the product detector and the oracle with identical templates."""
import numpy as np

import line_mod_pipeline_b200 as lm
from line_mod_pipeline_b200 import synth
from oracle import oracle as O


def make_pair(modalities=("cg", "dn"), T=(5, 8), sim_lut=None, **kw):
    """-> (product Detector, oracle Detector) with the same configuration and tables."""
    pm = [lm.ColorGradient() if m == "cg" else lm.DepthNormal() for m in modalities]
    om = [dict(type=O.CG) if m == "cg" else dict(type=O.DN) for m in modalities]
    det = lm.Detector(pm, T, **kw)
    if sim_lut is not None:
        det.setSimilarityLut(sim_lut)
    ora = O.Detector(om, list(T), sim_lut=det.getSimilarityLut(), normal_lut=det.getNormalLut())
    return det, ora


def sources(modalities, bgr, depth):
    return [bgr if m == "cg" else depth for m in modalities]


def add_random(det, ora, n, class_id="rand", n_modalities=2, levels=2, seed=99, **kw):
    for tp in synth.random_templates(n, n_modalities, levels, seed=seed, **kw):
        a = det.addSyntheticTemplate(tp, class_id)
        b = ora.add_synthetic(tp, class_id)
        assert a == b


def add_planted_from_oracle(det, ora, srcs, masks, class_id="planted"):
    """Extract with the oracle and install the very same pyramids in the product (isolates match parity)."""
    n = 0
    for m in masks:
        tid, _ = ora.add_template(srcs, class_id, m)
        if tid >= 0:
            flat = ora.get_template_flat(class_id, tid)
            det.addSyntheticTemplate(O.decode_pyramid(flat), class_id)
            n += 1
    return n


def rec_tuple(m):
    return [(int(a.x), int(a.y), float(a.similarity), int(a.class_index), int(a.template_id)) for a in m]


def assert_same_matches(got, want, what=""):
    g, w = rec_tuple(got), rec_tuple(want)
    assert len(g) == len(w), "%s: %d matches vs oracle %d" % (what, len(g), len(w))
    for i, (a, b) in enumerate(zip(g, w)):
        assert a == b, "%s: match %d differs: got %r want %r" % (what, i, a, b)
