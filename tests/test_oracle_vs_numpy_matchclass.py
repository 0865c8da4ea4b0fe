"""Second, independent restatement of upstream's matchClass in numpy, written from the
appendix's pseudo-code rather than from the C++ oracle, and compared with the oracle's coarse candidates and its
generation-order match list on a small frame.  The linear memories it reads are the oracle's (those are pinned by the
golden hashes G3/G4); what this test pins is the template side: template_positions, the flat "spill" reads,
raw_threshold, the +0.5f on coarse scores, the x*2+1 / clamp / (x/T - 8) refinement geometry, first-strictly-greater
arg-max and the order-preserving threshold filter.  CPU only."""
import numpy as np
from oracle import oracle as O
from line_mod_pipeline_b200 import synth

F32 = np.float32
ROWS, COLS, T_PYR = 480, 640, (5, 8)


def np_match_class(lms, sizes, pyramids, threshold, M=2):
    """lms[l][m] = uint8 [8][T*T*W*H] linear memories; sizes[l] = (Wi, Hi); pyramids = list of (class_index,
    template_id, [template dicts, index l*M+m]).  Returns (coarse candidates, final generation-order matches)."""
    L = len(sizes)
    coarse_all, final_all = [], []
    for cls, tid, tp in pyramids:
        Lc = L - 1
        T = T_PYR[Lc]; Wi, Hi = sizes[Lc]; W, H = Wi // T, Hi // T
        total = np.zeros(H * W, np.int32); nf = 0
        for m in range(M):
            t = tp[Lc * M + m]; feats = t["features"]; nf += len(feats)
            wf = (t["width"] - 1) // T + 1; hf = (t["height"] - 1) // T + 1
            P = (H - hf) * W + (W - wf) + 1
            sim = np.zeros(H * W, np.int32)
            for x, y, label in feats:
                if x < 0 or x >= Wi or y < 0 or y >= Hi:
                    continue
                base = ((y % T) * T + x % T) * (H * W) + (y // T) * W + x // T
                seg = lms[Lc][m][label][base:base + P].astype(np.int32)
                sim[:len(seg)] += seg
            total += sim & 0xFF                                          # u8 accumulators (never overflow: <= 63*4)
        raw_thr = int(F32(2 * nf) + (F32(threshold) / F32(100.0)) * F32(2 * nf) + F32(0.5))
        off = T // 2 + (T % 2 - 1)
        cands = []
        for j in np.nonzero(total > raw_thr)[0]:                         # raster order
            r, c = divmod(int(j), W)
            s = F32(F32(F32(int(total[j])) * F32(100.0)) / F32(4 * nf)) + F32(0.5)
            cands.append([c * T + off, r * T + off, s])
        coarse_all += [(x, y, float(s), cls, tid) for x, y, s in cands]
        for l in range(L - 2, -1, -1):
            T = T_PYR[l]; Wi, Hi = sizes[l]; W, H = Wi // T, Hi // T
            border = 8 * T; off = T // 2 + (T % 2 - 1)
            max_x = Wi - tp[l * M]["width"] - border; max_y = Hi - tp[l * M]["height"] - border
            for cd in cands:
                x, y = cd[0] * 2 + 1, cd[1] * 2 + 1
                x, y = max(x, border), max(y, border)
                x, y = min(x, max_x), min(y, max_y)
                ox, oy = (x // T - 8) * T, (y // T - 8) * T
                tot = np.zeros((16, 16), np.int32); nfl = 0
                for m in range(M):
                    t = tp[l * M + m]; nfl += len(t["features"])
                    for fx, fy, label in t["features"]:
                        fx, fy = fx + ox, fy + oy
                        if fx < 0 or fy < 0 or fx >= Wi or fy >= Hi:
                            continue
                        base = ((fy % T) * T + fx % T) * (H * W) + (fy // T) * W + fx // T
                        lm = lms[l][m][label]
                        for rr in range(16):
                            tot[rr] += lm[base + rr * W: base + rr * W + 16]
                best, br, bc = 0, -1, -1
                for rr in range(16):
                    for cc in range(16):
                        if tot[rr, cc] > best:
                            best, br, bc = int(tot[rr, cc]), rr, cc
                cd[0] = (x // T - 8 + bc) * T + off
                cd[1] = (y // T - 8 + br) * T + off
                cd[2] = F32(F32(best) * F32(100.0)) / F32(4 * nfl)
            cands = [cd for cd in cands if not (cd[2] < F32(threshold))]
        final_all += [(x, y, float(s), cls, tid) for x, y, s in cands]
    return coarse_all, final_all


def _rows(m):
    return [(int(a), int(b), float(c), int(d), int(e)) for a, b, c, d, e in zip(m.x, m.y, m.similarity, m.class_index, m.template_id)]


def test_numpy_matchclass_equals_oracle():
    ora = O.Detector([dict(type=O.CG), dict(type=O.DN)], list(T_PYR), normal_lut=synth.default_normal_lut())
    bgr, depth = synth.make_frame(6, ROWS, COLS)
    for i, m in enumerate(synth.object_masks(6, ROWS, COLS)[:10]):
        ora.add_template([bgr, depth], "objB" if i % 2 else "objA", m)
    for tp in synth.random_templates(40, 2, 2, seed=5, wh_range=(60, 200)):
        ora.add_synthetic(tp, "rand")
    classes = sorted(["objA", "objB", "rand"])
    pyramids = []
    for ci, cid in enumerate(classes):
        for t in range(ora.num_templates(cid)):
            pyramids.append((ci, t, O.decode_pyramid(ora.get_template_flat(cid, t))))
    assert sum(1 for p in pyramids if p[0] < 2) >= 4
    sizes = [(COLS, ROWS), (COLS // 2, ROWS // 2)]
    for thr in (80.0, 62.0):
        res = ora.match([bgr, depth], thr, debug=True)
        lms = [[res.linmem(l * 2 + m).reshape(8, -1) for m in range(2)] for l in range(2)]
        coarse, final = np_match_class(lms, sizes, pyramids, thr)
        assert coarse == _rows(res.matches(2)), "coarse candidates differ at threshold %g" % thr
        assert final == _rows(res.matches(1)), "generation-order matches differ at threshold %g" % thr
        assert len(final) > 0
        # A.7 epilogue: the final list is a sort + unique of the generation-order list
        fin = _rows(res.matches(0))
        assert sorted(set((x, y, s, c) for x, y, s, c, t in final)) == sorted(set((x, y, s, c) for x, y, s, c, t in fin))
        assert all(a[2] >= b[2] for a, b in zip(fin, fin[1:]))
