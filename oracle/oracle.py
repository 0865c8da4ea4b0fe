"""ctypes loader for the CPU oracle (oracle/linemod_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by line_mod_pipeline_b200.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblinemod_oracle.so")

CG, DN = 0, 1


def build(force=False):
    src = os.path.join(_HERE, "linemod_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.lmo_create.restype = C.c_void_p
        L.lmo_match.restype = C.c_void_p
        L.lmo_result_linmem.restype = C.c_long
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


# ------------------------------------------------------------------ primitives
def gauss7(img):
    img = _u8(img); ch = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    lib().lmo_gauss7(_p(img), img.shape[0], img.shape[1], ch, _p(out))
    return out


def sobel3(img):
    img = _u8(img); ch = 1 if img.ndim == 2 else img.shape[2]
    dx = np.empty(img.shape, np.int16); dy = np.empty(img.shape, np.int16)
    lib().lmo_sobel3(_p(img), img.shape[0], img.shape[1], ch, _p(dx), _p(dy))
    return dx, dy


def fast_atan2(y, x, fused=1):
    y = np.ascontiguousarray(y, np.float32); x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(y)
    lib().lmo_fast_atan2(_p(y), _p(x), _p(out), C.c_long(y.size), fused)
    return out


def label_table(fused=1):
    out = np.empty((2041, 2041), np.uint8)
    lib().lmo_label_table(_p(out), fused)
    return out


def pyrdown(img):
    img = _u8(img); ch = 1 if img.ndim == 2 else img.shape[2]
    shp = (img.shape[0] // 2, img.shape[1] // 2) + (() if img.ndim == 2 else (ch,))
    out = np.empty(shp, np.uint8)
    lib().lmo_pyrdown(_p(img), img.shape[0], img.shape[1], ch, _p(out))
    return out


def resize_nn(img, drows, dcols):
    img = _u8(img); out = np.empty((drows, dcols), np.uint8)
    lib().lmo_resize_nn(_p(img), img.shape[0], img.shape[1], _p(out), drows, dcols)
    return out


def median5(img):
    img = _u8(img); out = np.empty_like(img)
    lib().lmo_median5(_p(img), img.shape[0], img.shape[1], _p(out))
    return out


def erode3(img):
    img = _u8(img); out = np.empty_like(img)
    lib().lmo_erode3(_p(img), img.shape[0], img.shape[1], _p(out))
    return out


def dist_c(img):
    img = _u8(img); out = np.empty(img.shape, np.float32)
    lib().lmo_dist_c(_p(img), img.shape[0], img.shape[1], _p(out))
    return out


def cg_quantize(bgr, weak=10.0, fused=1):
    bgr = _u8(bgr); r, c = bgr.shape[:2]
    q = np.empty((r, c), np.uint8); mag = np.empty((r, c), np.float32)
    lib().lmo_cg_quantize(_p(bgr), r, c, C.c_float(weak), fused, _p(q), _p(mag))
    return q, mag


def dn_quantize(depth, lut, dist_thr=2000, diff_thr=50, median=True, want_idx=False):
    depth = np.ascontiguousarray(depth, np.uint16); r, c = depth.shape
    out = np.empty((r, c), np.uint8)
    idx = np.empty((3, r, c), np.int8) if want_idx else None
    lutp = _p(_u8(lut)) if lut is not None else None
    lib().lmo_dn_quantize(_p(depth), r, c, dist_thr, diff_thr, lutp, int(median), _p(out),
                          _p(idx) if want_idx else None)
    return (out, idx) if want_idx else out


def spread(q, T):
    q = _u8(q); out = np.empty_like(q)
    lib().lmo_spread(_p(q), q.shape[0], q.shape[1], T, _p(out))
    return out


def similarity_lut(circular=0):
    out = np.empty(256, np.uint8)
    lib().lmo_similarity_lut(circular, _p(out))
    return out


def response(sp, lut):
    sp = _u8(sp); out = np.empty((8,) + sp.shape, np.uint8)
    lib().lmo_response(_p(sp), C.c_long(sp.size), _p(_u8(lut)), _p(out))
    return out


def linearize(resp, T):
    resp = _u8(resp); r, c = resp.shape
    out = np.empty((T * T, (r // T) * (c // T)), np.uint8)
    lib().lmo_linearize(_p(resp), r, c, T, _p(out))
    return out


# ------------------------------------------------------------------ templates (flat int32)
def encode_pyramid(templates):
    """templates: list of dict(width,height,pyramid_level,features=[(x,y,label),...])."""
    flat = []
    for t in templates:
        f = np.asarray(t["features"], np.int32).reshape(-1, 3)
        flat += [t["width"], t["height"], t["pyramid_level"], len(f)]
        flat += f.reshape(-1).tolist()
    return np.asarray(flat, np.int32)


def decode_pyramid(flat):
    flat = np.asarray(flat, np.int32); out = []; p = 0
    while p < len(flat):
        w, h, l, nf = (int(v) for v in flat[p:p + 4]); p += 4
        out.append(dict(width=w, height=h, pyramid_level=l,
                        features=flat[p:p + 3 * nf].reshape(-1, 3).copy()))
        p += 3 * nf
    return out


class MatchResult:
    def __init__(self, L, r, n_maps):
        self._L, self._r, self._n = L, r, n_maps
        err = L.lmo_result_error(C.c_void_p(r))
        if err:
            L.lmo_result_free(C.c_void_p(r)); self._r = None
            raise ValueError("oracle match error %d (size not divisible by T / %%16)" % err)

    def matches(self, which=0):
        """which: 0 final (sorted+unique), 1 generation order (pre-sort), 2 coarse candidates."""
        L, r = self._L, C.c_void_p(self._r)
        n = L.lmo_result_count(r, which)
        x = np.empty(n, np.int32); y = np.empty(n, np.int32); s = np.empty(n, np.float32)
        c = np.empty(n, np.int32); t = np.empty(n, np.int32)
        if n:
            L.lmo_result_get(r, which, _p(x), _p(y), _p(s), _p(c), _p(t))
        return np.rec.fromarrays([x, y, s, c, t], names="x,y,similarity,class_index,template_id")

    def quantized(self, idx):
        L, r = self._L, C.c_void_p(self._r)
        rows = C.c_int(); cols = C.c_int()
        n = L.lmo_result_quantized(r, idx, None, C.byref(rows), C.byref(cols))
        out = np.empty((rows.value, cols.value), np.uint8)
        assert n == out.size, "run match(debug=True)"
        L.lmo_result_quantized(r, idx, _p(out), C.byref(rows), C.byref(cols))
        return out

    def linmem(self, idx):
        L, r = self._L, C.c_void_p(self._r)
        n = L.lmo_result_linmem(r, idx, None)
        out = np.empty(n, np.uint8)
        L.lmo_result_linmem(r, idx, _p(out))
        return out

    def stats(self):
        out = np.zeros(5, np.float64)
        self._L.lmo_result_stats(C.c_void_p(self._r), _p(out))
        return dict(t_frame=out[0], t_match=out[1], t_sort=out[2], bytes_coarse=int(out[3]), bytes_local=int(out[4]))

    def __del__(self):
        if getattr(self, "_r", None):
            self._L.lmo_result_free(C.c_void_p(self._r)); self._r = None


class Detector:
    """Oracle detector.  modalities: list of dict(type=CG|DN, ...optional params)."""

    def __init__(self, modalities, T, sim_lut=None, normal_lut=None, fused_atan=1):
        L = lib()
        self.M = len(modalities); self.levels = len(T); self.T = list(T)
        types = np.asarray([m["type"] for m in modalities], np.int32)
        fp = np.asarray([[m.get("weak_threshold", 10.0), m.get("strong_threshold", 55.0)] for m in modalities], np.float32)
        ip = np.asarray([[m.get("num_features", 63), m.get("distance_threshold", 2000),
                          m.get("difference_threshold", 50), m.get("extract_threshold", 2)] for m in modalities], np.int32)
        Tarr = np.asarray(T, np.int32)
        self._keep = (types, fp, ip, Tarr)
        sl = _u8(sim_lut) if sim_lut is not None else None
        nl = _u8(normal_lut) if normal_lut is not None else None
        if nl is not None:
            assert nl.size == 8000
        self._h = L.lmo_create(self.M, _p(types), _p(fp), _p(ip), self.levels, _p(Tarr),
                               _p(sl) if sl is not None else None, _p(nl) if nl is not None else None, fused_atan)
        self._L = L

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.lmo_destroy(C.c_void_p(self._h)); self._h = None

    @staticmethod
    def _srcs(sources):
        arrs = [np.ascontiguousarray(s) for s in sources]
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        return arrs, ptrs

    def add_template(self, sources, class_id, mask=None):
        arrs, ptrs = self._srcs(sources)
        r, c = arrs[0].shape[:2]
        bb = np.zeros(4, np.int32)
        m = _u8(mask) if mask is not None else None
        tid = self._L.lmo_add_template(C.c_void_p(self._h), class_id.encode(), ptrs, r, c,
                                       _p(m) if m is not None else None, _p(bb))
        return tid, tuple(int(v) for v in bb)

    def add_synthetic(self, templates, class_id):
        flat = templates if isinstance(templates, np.ndarray) else encode_pyramid(templates)
        flat = np.ascontiguousarray(flat, np.int32)
        tid = self._L.lmo_add_synthetic(C.c_void_p(self._h), class_id.encode(), _p(flat), len(flat))
        if tid < 0:
            raise ValueError("bad template pyramid")
        return tid

    def num_templates(self, class_id=None):
        return self._L.lmo_num_templates(C.c_void_p(self._h), class_id.encode() if class_id else None)

    def class_ids(self):
        n = self._L.lmo_num_classes(C.c_void_p(self._h)); out = []
        buf = C.create_string_buffer(1024)
        for i in range(n):
            self._L.lmo_class_id(C.c_void_p(self._h), i, buf, 1024); out.append(buf.value.decode())
        return out

    def get_template_flat(self, class_id, template_id):
        n = self._L.lmo_get_template_flat(C.c_void_p(self._h), class_id.encode(), template_id, None, 0)
        if n < 0:
            raise KeyError((class_id, template_id))
        out = np.empty(n, np.int32)
        self._L.lmo_get_template_flat(C.c_void_p(self._h), class_id.encode(), template_id, _p(out), n)
        return out

    def similarity_lut(self):
        out = np.empty(256, np.uint8)
        self._L.lmo_get_similarity_lut(C.c_void_p(self._h), _p(out))
        return out

    def similarity_map(self, result, class_id, template_id):
        """u16 [H, W] coarse-level map of `similarity` + `addSimilarities` for one template (result of match(debug=True))."""
        self._L.lmo_result_similarity_map.restype = C.c_long
        n = self._L.lmo_result_similarity_map(C.c_void_p(self._h), C.c_void_p(result._r), class_id.encode(), int(template_id), None)
        if n < 0:
            raise KeyError((class_id, template_id, n))
        out = np.empty(n, np.uint16)
        self._L.lmo_result_similarity_map(C.c_void_p(self._h), C.c_void_p(result._r), class_id.encode(), int(template_id), _p(out))
        return out

    def match(self, sources, threshold, class_ids=(), masks=None, threads=1, debug=False):
        arrs, ptrs = self._srcs(sources)
        r, c = arrs[0].shape[:2]
        cls = (C.c_char_p * max(1, len(class_ids)))(*[s.encode() for s in class_ids])
        mp = None
        if masks is not None:
            marrs = [_u8(m) for m in masks]
            mp = (C.c_void_p * len(marrs))(*[a.ctypes.data for a in marrs])
        res = self._L.lmo_match(C.c_void_p(self._h), ptrs, r, c, C.c_float(threshold), cls, len(class_ids),
                                mp, int(threads), int(debug))
        return MatchResult(self._L, res, self.levels * self.M)


def max_threads():
    return lib().lmo_max_threads()
