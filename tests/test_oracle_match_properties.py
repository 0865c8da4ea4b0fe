"""Properties of the oracle's Detector::match (the checker the GPU parity tests compare against): they pin the
parts of upstream's matchClass / match epilogue that no primitive-level golden vector covers.
CPU only."""
import numpy as np
from oracle import oracle as O
from line_mod_pipeline_b200 import synth


def _detector(n_random=120, frame=4):
    ora = O.Detector([dict(type=O.CG), dict(type=O.DN)], [5, 8], normal_lut=synth.default_normal_lut())
    bgr, depth = synth.make_frame(frame)
    planted = 0
    for i, m in enumerate(synth.object_masks(frame)[:12]):
        tid, _ = ora.add_template([bgr, depth], "objB" if i % 2 else "objA", m)
        planted += tid >= 0
    for tp in synth.random_templates(n_random, 2, 2, seed=99):
        ora.add_synthetic(tp, "rand")
    return ora, [bgr, depth], planted


def _rows(m):
    return [(int(a), int(b), float(c), int(d), int(e)) for a, b, c, d, e in zip(m.x, m.y, m.similarity, m.class_index, m.template_id)]


def test_planted_templates_are_found_at_100_percent():
    ora, src, planted = _detector()
    assert planted >= 5
    m = ora.match(src, 95.0).matches(0)
    hits = {(r[3], r[4]) for r in _rows(m) if r[2] >= 99.9}
    assert len(hits) >= planted - 1          # every planted template matches its own frame (one may tie out in unique())


def test_result_is_sorted_unique_and_thresholded():
    ora, src, _ = _detector()
    for thr in (80.0, 60.0):
        rows = _rows(ora.match(src, thr).matches(0))
        assert rows, "no matches at %g" % thr
        sims = [r[2] for r in rows]
        assert all(a >= b for a, b in zip(sims, sims[1:]))                       # Match::operator< : similarity desc
        assert all(s >= thr for s in sims)                                       # remove_if(similarity < threshold)
        keys = [(r[0], r[1], r[2], r[3]) for r in rows]
        assert all(a != b for a, b in zip(keys, keys[1:]))                       # std::unique on (x, y, similarity, class)


def test_lower_threshold_only_adds_matches():
    """Every (x, y, class, template) found at 80 is found at 60 with the same score (the refinement of a candidate
    does not depend on the threshold; a lower threshold only admits more coarse candidates)."""
    ora, src, _ = _detector()
    hi = {(r[0], r[1], r[3], r[4]): r[2] for r in _rows(ora.match(src, 80.0, debug=True).matches(1))}
    lo = {(r[0], r[1], r[3], r[4]): r[2] for r in _rows(ora.match(src, 60.0, debug=True).matches(1))}
    assert hi and set(hi) <= set(lo)
    assert all(lo[k] == v for k, v in hi.items())


def test_threads_and_repeats_do_not_change_the_answer():
    ora, src, _ = _detector()
    ref = _rows(ora.match(src, 65.0, threads=1).matches(0))
    for threads in (2, 8):
        assert _rows(ora.match(src, 65.0, threads=threads).matches(0)) == ref
    assert _rows(ora.match(src, 65.0, threads=1).matches(0)) == ref


def test_class_filter_is_a_subset_in_class_order():
    ora, src, _ = _detector()
    both = _rows(ora.match(src, 70.0, debug=True).matches(1))     # generation order: classes in map order, template id asc
    only = _rows(ora.match(src, 70.0, class_ids=["objB"], debug=True).matches(1))
    cls = sorted(["objA", "objB", "rand"]).index("objB")
    assert [r[3] for r in both] == sorted(r[3] for r in both)     # classes are visited in lexicographic order
    assert only and [(r[0], r[1], r[2], r[4]) for r in only] == [(r[0], r[1], r[2], r[4]) for r in both if r[3] == cls]
