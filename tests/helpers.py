"""Shared helpers for parity tests: build the product detector and the oracle with identical templates."""
import ctypes as C
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


def sharded_step_case(rng, world, frames, max_per_rank, ntpl, interleaved, sims, tid_mod, gcap_extra=0):
    """Builds what the match all-gather leaves in device memory for a synthetic sharded step and the expected finished
    lists: generation order (by selection position for interleaved shards, rank-ordered concatenation otherwise), then the
    library's own host epilogue (lmb200_merge_matches = std::sort + std::unique)."""
    L = lm.capi.lib()
    MR = lm.capi.MatchRec
    perm = rng.permutation(ntpl).astype(np.int32)              # selection order: position p holds global template perm[p]
    pos_of_g = np.empty(ntpl, np.int32); pos_of_g[perm] = np.arange(ntpl, dtype=np.int32)
    g_class = (np.arange(ntpl) % 3).astype(np.int32)
    g_tid = (np.arange(ntpl) % tid_mod).astype(np.int32)
    lists = [[None] * frames for _ in range(world)]
    for r in range(world):
        # templates of rank r in its own generation order
        mine = perm[r::world] if interleaved else perm[r * (ntpl // world):(r + 1) * (ntpl // world)]
        if not interleaved:
            mine = np.sort(mine)                               # contiguous shards: global index order within the rank is irrelevant to the kernel
        for f in range(frames):
            n = int(rng.integers(0, max_per_rank + 1)) if rng.random() > 0.15 else 0
            k = np.sort(rng.integers(0, len(mine), n))         # several records per template, templates in generation order
            rec = np.zeros((n, 4), np.int32)
            rec[:, 0] = mine[k]
            rec[:, 1] = rng.integers(0, 8, n) * 5; rec[:, 2] = rng.integers(0, 6, n) * 5   # few positions: exact duplicates happen
            rec[:, 3] = (80.0 + 0.5 * rng.integers(0, sims, n)).astype(np.float32).view(np.int32)
            lists[r][f] = rec
    gcap = max(1, max(sum(len(lists[r][f]) for f in range(frames)) for r in range(world))) + gcap_extra
    stride = 2 * frames + gcap
    G = np.zeros((world, stride, 4), np.int32)
    for r in range(world):
        off = 0
        for f in range(frames):
            n = len(lists[r][f])
            G[r, 2 * f] = (n, 0, off, 0)
            G[r, 2 * f + 1] = (100 * r + f, 0, 7 * r + f, 0)    # counters: must come back for `rank`
            G[r, 2 * frames + off: 2 * frames + off + n] = lists[r][f]
            off += n
    want = []
    for f in range(frames):
        cat = np.concatenate([lists[r][f] for r in range(world)]) if world else np.zeros((0, 4), np.int32)
        if interleaved and len(cat):
            cat = cat[np.argsort(pos_of_g[cat[:, 0]], kind="stable")]
        m = np.zeros(len(cat), lm.MATCH_DTYPE)
        m["x"], m["y"], m["similarity"] = cat[:, 1], cat[:, 2], cat[:, 3].view(np.float32)
        m["class_index"], m["template_id"] = g_class[cat[:, 0]], g_tid[cat[:, 0]]
        out = np.zeros(max(1, len(m)), lm.MATCH_DTYPE)
        n_out = C.c_size_t(0)
        parts = (C.POINTER(MR) * 1)(m.ctypes.data_as(C.POINTER(MR)))
        counts = (C.c_size_t * 1)(len(m))
        rc = L.lmb200_merge_matches(parts, counts, 1, out.ctypes.data_as(C.POINTER(MR)), len(out), C.byref(n_out))
        assert rc == 0
        want.append((len(cat), out[:n_out.value].copy()))
    return G, gcap, pos_of_g if interleaved else None, g_class, g_tid, want
