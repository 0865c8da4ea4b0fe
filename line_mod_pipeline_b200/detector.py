"""Python mirror of the reference-facing Detector surface, a thin layer over the C ABI.

Method names and argument meaning follow cv::linemod::Detector as the reference calls it
(/root/reference/src/HighLevelLinemod.cpp:26-43 ctor, :93 addTemplate, :152 match, :115 getTemplates,
:55 classIds, :65 numTemplates, :184 getT, :256-320 read/write), so parity tests read like the
reference's own call sites.  All compute happens in liblmb200.so (CUDA); nothing here has a CPU path.
"""
import ctypes as C
import numpy as np

from . import _capi as K


class LinemodError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("lmb200 error %d: %s" % (code, msg))
        self.code = code


MATCH_DTYPE = np.dtype([("x", np.int32), ("y", np.int32), ("similarity", np.float32),
                        ("class_index", np.int32), ("template_id", np.int32)])
assert MATCH_DTYPE.itemsize == C.sizeof(K.MatchRec)
_MATCH_REC_DTYPE = np.dtype((np.record, MATCH_DTYPE))


class MatchArray(np.ndarray):
    """Per-frame match list: a structured array whose fields also read as attributes (m.x, m.similarity, m[0].template_id).
    np.recarray offers the same but costs ~20 us per slice, which at 128 frames per step was more than a sharded step's
    whole device time; this subclass slices at plain-ndarray speed."""
    __slots__ = ()

    def __getattr__(self, name):           # only reached when normal attribute lookup fails
        if name in MATCH_DTYPE.fields:
            return self[name].view(np.ndarray)
        raise AttributeError(name)


def _as_matches(a):
    return a.view(_MATCH_REC_DTYPE).view(MatchArray)


def _split(out, offs, count, copy=True):
    """One flat result buffer + offsets -> list of per-frame MatchArray (ONE copy of the used records, then views)."""
    o = np.frombuffer(offs, dtype=np.uintp, count=count + 1).tolist()
    flat = out[:o[count]]
    flat = _as_matches(flat.copy() if copy else flat)
    return [flat[o[i]:o[i + 1]] for i in range(count)]


def ColorGradient(weak_threshold=10.0, num_features=63, strong_threshold=55.0):
    return dict(type=K.COLOR_GRADIENT, weak_threshold=weak_threshold, num_features=num_features,
                strong_threshold=strong_threshold)


def DepthNormal(distance_threshold=2000, difference_threshold=50, num_features=63, extract_threshold=2):
    return dict(type=K.DEPTH_NORMAL, distance_threshold=distance_threshold, difference_threshold=difference_threshold,
                num_features=num_features, extract_threshold=extract_threshold)


def _enc(s):
    """class ids are byte strings in the C ABI; arbitrary bytes round-trip through surrogateescape"""
    return s.encode("utf-8", "surrogateescape")


def _dec(b):
    return b.decode("utf-8", "surrogateescape")


def _image(a):
    """numpy array -> (lmb200_image, keepalive)."""
    if a is None:
        return K.Image(None, 0, 0, K.T_8UC1, 0), None
    if a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 3:
        t = K.T_8UC3
    elif a.dtype == np.uint16 and a.ndim == 2:
        t = K.T_16UC1
    elif a.dtype == np.uint8 and a.ndim == 2:
        t = K.T_8UC1
    else:
        raise TypeError("image must be uint8 HxWx3 (BGR), uint16 HxW (depth mm) or uint8 HxW (mask)")
    if a.strides[-1] != a.itemsize or (a.ndim == 3 and a.strides[1] != 3):
        a = np.ascontiguousarray(a)
    return K.Image(a.ctypes.data, a.shape[0], a.shape[1], t, a.strides[0]), a


def _cstr_array(ids):
    ids = list(ids or [])
    arr = (C.c_char_p * max(1, len(ids)))(*[_enc(s) for s in ids])
    return arr, len(ids)


class Detector:
    """Detector(modalities, T_pyramid) — see getDefaultLINE / getDefaultLINEMOD for the reference's two wirings."""

    def __init__(self, modalities=None, T_pyramid=(5, 8), device=-1, max_batch=0, candidate_capacity=0, similarity_lut=0,
                 _handle=None):
        self._L = K.lib()
        self._h = K._H()
        if _handle is not None:
            self._h = _handle
            return
        cfg = K.Config()
        cfg.num_modalities = len(modalities)
        for i, m in enumerate(modalities):
            self._L.lmb200_default_modality(m["type"], C.byref(cfg.modalities[i]))
            for k, v in m.items():
                setattr(cfg.modalities[i], k, v)
        cfg.pyramid_levels = len(T_pyramid)
        for i, t in enumerate(T_pyramid):
            cfg.T[i] = int(t)
        cfg.device = device
        cfg.max_batch = max_batch
        cfg.candidate_capacity = candidate_capacity
        cfg.similarity_lut = similarity_lut     # K.SIMLUT_CIRCULAR (default, upstream's table) | K.SIMLUT_LINEAR
        rc = self._L.lmb200_create(C.byref(cfg), C.byref(self._h))
        if rc:
            raise LinemodError(rc, (self._L.lmb200_last_error(None) or b"").decode("utf-8", "replace"))

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc, allow=()):
        if rc and rc not in allow:
            raise LinemodError(rc, (self._L.lmb200_last_error(self._h) or b"").decode("utf-8", "replace"))
        return rc

    def close(self):
        if getattr(self, "_h", None):
            self._L.lmb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # ------------------------------------------------------------------ introspection
    def getModalities(self):
        return [self._L.lmb200_modality_name(self._h, i).decode() for i in range(self._L.lmb200_num_modalities(self._h))]

    def pyramidLevels(self):
        return self._L.lmb200_pyramid_levels(self._h)

    def getT(self, level):
        return self._check_nonneg(self._L.lmb200_get_T(self._h, level))

    def _check_nonneg(self, v):
        if v < 0:
            raise LinemodError(v, "bad argument")
        return v

    def numClasses(self):
        return self._L.lmb200_num_classes(self._h)

    def classIds(self):
        return [_dec(self._L.lmb200_class_id(self._h, i)) for i in range(self.numClasses())]

    def numTemplates(self, class_id=None):
        return self._L.lmb200_num_templates(self._h, _enc(class_id) if class_id is not None else None)

    def getTemplates(self, class_id, template_id):
        """-> list (level*M+modality) of dict(width,height,pyramid_level,features=int32[n,3])."""
        n = self.pyramidLevels() * len(self.getModalities())
        out = []
        for i in range(n):
            t = K.Template()
            self._check(self._L.lmb200_get_template(self._h, _enc(class_id), template_id, i, C.byref(t)))
            f = np.zeros((t.num_features, 3), np.int32)
            if t.num_features:
                f[:] = np.ctypeslib.as_array(C.cast(t.features, C.POINTER(C.c_int)), shape=(t.num_features, 3))
            out.append(dict(width=t.width, height=t.height, pyramid_level=t.pyramid_level, features=f))
        return out

    # ------------------------------------------------------------------ templates
    def addTemplate(self, sources, class_id, object_mask=None):
        """-> (template_id, (x, y, w, h)); template_id == -1 when extraction fails (reference: HighLevelLinemod.cpp:97)."""
        imgs = [_image(s) for s in sources]
        arr = (K.Image * len(imgs))(*[i[0] for i in imgs])
        mimg, mkeep = _image(object_mask)
        bb = (C.c_int * 4)()
        tid = C.c_int(-1)
        self._check(self._L.lmb200_add_template(self._h, _enc(class_id), arr, len(imgs),
                                                C.byref(mimg) if object_mask is not None else None, bb, C.byref(tid)))
        return tid.value, tuple(bb)

    def addTemplates(self, views, class_id, object_masks=None):
        """Bulk addTemplate: `views` = list of per-view source lists, `object_masks` = list (entries may be None).
        -> list of (template_id, (x, y, w, h)), the result of len(views) successive addTemplate calls."""
        n = len(views)
        if n == 0:
            return []
        M = len(views[0])
        keep = [[_image(s) for s in v] for v in views]
        arr = (K.Image * (n * M))(*[i[0] for v in keep for i in v])
        marr = None
        if object_masks is not None:
            mkeep = [_image(m) for m in object_masks]
            marr = (K.Image * n)(*[m[0] for m in mkeep])
        bb = (C.c_int * (4 * n))()
        tids = (C.c_int * n)()
        self._check(self._L.lmb200_add_templates(self._h, _enc(class_id), n, arr, M, marr, bb, tids))
        return [(tids[i], tuple(bb[4 * i:4 * i + 4])) for i in range(n)]

    def addSyntheticTemplate(self, templates, class_id):
        arr = (K.Template * len(templates))()
        keep = []
        for i, t in enumerate(templates):
            f = np.ascontiguousarray(np.asarray(t["features"], np.int32).reshape(-1, 3))
            keep.append(f)
            arr[i] = K.Template(t["width"], t["height"], t["pyramid_level"], len(f), C.cast(f.ctypes.data, C.POINTER(K.Feature)))
        tid = C.c_int(-1)
        self._check(self._L.lmb200_add_synthetic_template(self._h, _enc(class_id), arr, len(templates), C.byref(tid)))
        return tid.value

    def uploadTemplates(self):
        """Pack the template set and copy it to the device now (the match calls do it on demand otherwise)."""
        self._check(self._L.lmb200_upload_templates(self._h))

    def clearTemplates(self):
        self._check(self._L.lmb200_clear_templates(self._h))

    # ------------------------------------------------------------------ persistence
    def write(self, path):
        self._check(self._L.lmb200_write(self._h, str(path).encode()))

    @classmethod
    def read(cls, path, device=-1):
        L = K.lib()
        h = K._H()
        rc = L.lmb200_read(str(path).encode(), device, C.byref(h))
        if rc:
            raise LinemodError(rc, (L.lmb200_last_error(None) or b"").decode("utf-8", "replace"))
        return cls(_handle=h)

    def writeCache(self, path):
        self._check(self._L.lmb200_write_cache(self._h, str(path).encode()))

    @classmethod
    def readCache(cls, path, device=-1):
        L = K.lib()
        h = K._H()
        rc = L.lmb200_read_cache(str(path).encode(), device, C.byref(h))
        if rc:
            raise LinemodError(rc, (L.lmb200_last_error(None) or b"").decode("utf-8", "replace"))
        return cls(_handle=h)

    def writeClass(self, class_id, path):
        self._check(self._L.lmb200_write_class(self._h, _enc(class_id), str(path).encode()))

    def readClass(self, path, class_id_override=""):
        self._check(self._L.lmb200_read_class(self._h, str(path).encode(), _enc(class_id_override) if class_id_override else None))

    def writeClasses(self, fmt="templates_%s.yml.gz"):
        self._check(self._L.lmb200_write_classes(self._h, fmt.encode()))

    def readClasses(self, class_ids, fmt="templates_%s.yml.gz"):
        arr, n = _cstr_array(class_ids)
        self._check(self._L.lmb200_read_classes(self._h, arr, n, fmt.encode()))

    # ------------------------------------------------------------------ matching
    def match(self, sources, threshold, class_ids=(), quantized_images=False, masks=None):
        """-> structured array (x,y,similarity,class_index,template_id) in the reference's order;
        with quantized_images=True also returns the list of quantized maps (index level*M+modality)."""
        imgs = [_image(s) for s in sources]
        arr = (K.Image * len(imgs))(*[i[0] for i in imgs])
        ids, nids = _cstr_array(class_ids)
        qarr, qimgs = None, None
        if quantized_images:
            r, c = sources[0].shape[:2]
            M, L = len(self.getModalities()), self.pyramidLevels()
            qimgs = []
            for l in range(L):
                for m in range(M):
                    qimgs.append(np.zeros((r, c), np.uint8))
                r //= 2; c //= 2
            qarr = (K.Image * len(qimgs))(*[_image(q)[0] for q in qimgs])
        marr, mkeep = None, None
        if masks is not None:
            mkeep = [_image(m) for m in masks]
            marr = (K.Image * len(mkeep))(*[i[0] for i in mkeep])
        cap = 4096
        while True:
            out = np.zeros(cap, MATCH_DTYPE)
            n = C.c_size_t(0)
            rc = self._check(self._L.lmb200_match(self._h, arr, len(imgs), C.c_float(threshold), ids, nids,
                                                  out.ctypes.data_as(C.POINTER(K.MatchRec)), cap, C.byref(n), qarr, marr),
                             allow=(K.E_TRUNCATED,))
            if rc == K.E_TRUNCATED:
                cap = int(n.value)
                continue
            res = _as_matches(out[:n.value])
            return (res, qimgs) if quantized_images else res

    def _frames(self, frames):
        keep, flat = [], []
        for fr in frames:
            for s in fr:
                im, k = _image(s)
                flat.append(im); keep.append(k)
        return (K.Image * len(flat))(*flat), keep, len(frames[0])

    def matchBatch(self, frames, threshold, class_ids=(), cap=None):
        """frames: list of [source per modality].  -> list of per-frame match arrays."""
        arr, keep, nsrc = self._frames(frames)
        ids, nids = _cstr_array(class_ids)
        cap = cap or 1024 * len(frames)
        while True:
            out = self._outbuf(cap)
            offs = (C.c_size_t * (len(frames) + 1))()
            rc = self._check(self._L.lmb200_match_batch(self._h, arr, len(frames), nsrc, C.c_float(threshold), ids, nids,
                                                        out.ctypes.data_as(C.POINTER(K.MatchRec)), cap, offs),
                             allow=(K.E_TRUNCATED,))
            if rc == K.E_TRUNCATED:
                cap = int(offs[len(frames)])
                continue
            return _split(out, offs, len(frames))

    def _outbuf(self, cap):
        """Reusable (uninitialised) result buffer: zero-filling tens of MB per call would dominate a batch call."""
        buf = getattr(self, "_out_cache", None)
        if buf is None or len(buf) < cap:
            buf = np.empty(cap, MATCH_DTYPE)
            self._out_cache = buf
        return buf

    def prepareSingle(self, sources, cap=8192):
        """Marshals ONE frame once so matchPreparedSingle() costs exactly one lmb200_match C-ABI call (what the reference's
        detector->match(...) costs a C++ caller)."""
        imgs = [_image(s) for s in sources]
        arr = (K.Image * len(imgs))(*[i[0] for i in imgs])
        return dict(arr=arr, keep=imgs, n=len(imgs), cap=cap, out=np.empty(cap, MATCH_DTYPE), count=C.c_size_t(0))

    def matchPreparedSingle(self, prep, threshold, class_ids=()):
        ids, nids = _cstr_array(class_ids)
        self._check(self._L.lmb200_match(self._h, prep["arr"], prep["n"], C.c_float(threshold), ids, nids,
                                         prep["out"].ctypes.data_as(C.POINTER(K.MatchRec)), prep["cap"], C.byref(prep["count"]), None, None))
        return int(prep["count"].value)

    def prepareBatch(self, frames, cap=None):
        """Marshals a frame list once (ctypes image array + result buffers) so repeated matchPrepared() calls
        cost exactly one lmb200_match_batch C-ABI call — what a C/C++ caller pays."""
        arr, keep, nsrc = self._frames(frames)
        cap = cap or 1024 * len(frames)
        return dict(arr=arr, keep=keep, nsrc=nsrc, n=len(frames), cap=cap, out=np.empty(cap, MATCH_DTYPE),
                    offs=(C.c_size_t * (len(frames) + 1))())

    def matchPrepared(self, prep, threshold, class_ids=()):
        ids, nids = _cstr_array(class_ids)
        rc = self._check(self._L.lmb200_match_batch(self._h, prep["arr"], prep["n"], prep["nsrc"], C.c_float(threshold), ids, nids,
                                                    prep["out"].ctypes.data_as(C.POINTER(K.MatchRec)), prep["cap"], prep["offs"]),
                         allow=(K.E_TRUNCATED,))
        if rc == K.E_TRUNCATED:
            raise LinemodError(rc, "prepared output buffer too small: need %d records" % prep["offs"][prep["n"]])
        return int(prep["offs"][prep["n"]])

    def submitPrepared(self, prep, threshold, class_ids=()):
        """Non-blocking half of matchPrepared: enqueues the whole batch, returns a ticket (max. two in flight)."""
        ids, nids = _cstr_array(class_ids)
        t = C.c_int(-1)
        self._check(self._L.lmb200_match_batch_submit(self._h, prep["arr"], prep["n"], prep["nsrc"], C.c_float(threshold), ids, nids, C.byref(t)))
        return t.value

    def collectPrepared(self, prep, ticket):
        rc = self._check(self._L.lmb200_match_batch_collect(self._h, ticket, prep["out"].ctypes.data_as(C.POINTER(K.MatchRec)),
                                                            prep["cap"], prep["offs"]), allow=(K.E_TRUNCATED,))
        if rc == K.E_TRUNCATED:
            raise LinemodError(rc, "prepared output buffer too small: need %d records" % prep["offs"][prep["n"]])
        return int(prep["offs"][prep["n"]])

    def uploadFrames(self, frames, first_slot=0):
        arr, keep, nsrc = self._frames(frames)
        self._check(self._L.lmb200_upload_frames(self._h, arr, len(frames), nsrc, first_slot))

    def matchResident(self, first_slot, count, threshold, class_ids=()):
        ids, nids = _cstr_array(class_ids)
        self._check(self._L.lmb200_match_resident(self._h, first_slot, count, C.c_float(threshold), ids, nids))

    def matchResidentSharded(self, first_slot, count, threshold, class_ids=()):
        """Template-sharded step with the quantisers sharded by frame block + NCCL all-gather of the quantized maps."""
        ids, nids = _cstr_array(class_ids)
        self._check(self._L.lmb200_match_resident_sharded(self._h, first_slot, count, C.c_float(threshold), ids, nids))

    def prepareUpload(self, frames):
        """Marshals a frame list once (ctypes image array) so uploadPrepared() costs one lmb200_upload_frames call."""
        arr, keep, nsrc = self._frames(frames)
        return dict(arr=arr, keep=keep, nsrc=nsrc, n=len(frames))

    def uploadPrepared(self, prep, first_slot=0):
        self._check(self._L.lmb200_upload_frames(self._h, prep["arr"], prep["n"], prep["nsrc"], first_slot))

    def prepareFetch(self, count, cap=None):
        """Result buffers of fetchResidentPrepared(): records of all frames back to back in prep['out'], frame i at
        prep['offs'][i] .. prep['offs'][i+1] (what a C caller of lmb200_fetch_resident[_allgather] holds)."""
        cap = cap or 1024 * count
        return dict(n=count, cap=cap, out=np.empty(cap, MATCH_DTYPE), offs=(C.c_size_t * (count + 1))())

    def fetchResidentPrepared(self, prep, first_slot, allgather=False):
        """One C-ABI call into the prepared buffers; returns the number of records.  lists(prep) makes the per-frame views."""
        fn = self._L.lmb200_fetch_resident_allgather if allgather else self._L.lmb200_fetch_resident
        rc = self._check(fn(self._h, first_slot, prep["n"], prep["out"].ctypes.data_as(C.POINTER(K.MatchRec)), prep["cap"], prep["offs"]),
                         allow=(K.E_TRUNCATED,))
        if rc == K.E_TRUNCATED:
            raise LinemodError(rc, "prepared output buffer too small: need %d records" % prep["offs"][prep["n"]])
        return int(prep["offs"][prep["n"]])

    @staticmethod
    def lists(prep, copy=True):
        return _split(prep["out"], prep["offs"], prep["n"], copy)

    def fetchResident(self, first_slot, count, allgather=False, cap=None):
        cap = cap or 1024 * count
        fn = self._L.lmb200_fetch_resident_allgather if allgather else self._L.lmb200_fetch_resident
        while True:
            out = self._outbuf(cap)
            offs = (C.c_size_t * (count + 1))()
            rc = self._check(fn(self._h, first_slot, count, out.ctypes.data_as(C.POINTER(K.MatchRec)), cap, offs),
                             allow=(K.E_TRUNCATED,))
            if rc == K.E_TRUNCATED:
                cap = int(offs[count])
                if allgather:
                    raise LinemodError(rc, "output capacity too small for a collective fetch; pass cap=")
                continue
            return _split(out, offs, count)

    def synchronize(self):
        self._check(self._L.lmb200_synchronize(self._h))

    def stream(self):
        return self._L.lmb200_stream(self._h)

    def timerRecord(self, which):
        self._check(self._L.lmb200_timer_record(self._h, which))

    def timerElapsedMs(self):
        ms = C.c_float(0)
        self._check(self._L.lmb200_timer_elapsed_ms(self._h, C.byref(ms)))
        return ms.value

    # ------------------------------------------------------------------ tables / sharding / profile / debug
    def setSimilarityLut(self, lut):
        lut = np.ascontiguousarray(lut, np.uint8); assert lut.size == 256
        self._check(self._L.lmb200_set_similarity_lut(self._h, lut.ctypes.data))

    def getSimilarityLut(self):
        out = np.zeros(256, np.uint8)
        self._check(self._L.lmb200_get_similarity_lut(self._h, out.ctypes.data))
        return out

    def setNormalLut(self, lut):
        lut = np.ascontiguousarray(lut, np.uint8); assert lut.size == 8000
        self._check(self._L.lmb200_set_normal_lut(self._h, lut.ctypes.data))

    def getNormalLut(self):
        out = np.zeros(8000, np.uint8)
        self._check(self._L.lmb200_get_normal_lut(self._h, out.ctypes.data))
        return out

    def setTemplateShard(self, rank, world):
        self._check(self._L.lmb200_set_template_shard(self._h, rank, world))

    def commInit(self, unique_id, rank, world):
        uid = np.ascontiguousarray(unique_id, np.uint8); assert uid.size == 128
        self._check(self._L.lmb200_comm_init(self._h, uid.ctypes.data, rank, world))

    def setProfiling(self, on):
        self._check(self._L.lmb200_set_profiling(self._h, int(on)))

    def getProfile(self, reset=False):
        p = K.Profile()
        self._check(self._L.lmb200_get_profile(self._h, C.byref(p), int(reset)))
        d = dict(ms={K.K_NAMES[i]: p.ms[i] for i in range(len(K.K_NAMES))}, launches={K.K_NAMES[i]: p.launches[i] for i in range(len(K.K_NAMES))},
                 bytes_coarse=p.bytes_coarse, bytes_local=p.bytes_local, frames=p.frames, candidates=p.candidates,
                 matches=p.matches, chunks_coarse=p.chunks_coarse)
        return d

    def postmatchColor(self, matches, lower_hsv, upper_hsv, slot=0):
        """GPU colour check of the reference (HighLevelLinemod.cpp:159-161, :113-135, :424-434) for `matches` (a match record
        array of this detector) on the frame resident in `slot` -> (inside, total) int32 arrays; -1 = hull leaves the image."""
        m = np.ascontiguousarray(matches, MATCH_DTYPE)
        lo = np.ascontiguousarray(lower_hsv, np.uint8); hi = np.ascontiguousarray(upper_hsv, np.uint8)
        inside = np.zeros(len(m), np.int32); total = np.zeros(len(m), np.int32)
        self._check(self._L.lmb200_postmatch_color(self._h, slot, lo.ctypes.data, hi.ctypes.data, m.ctypes.data_as(C.POINTER(K.MatchRec)), len(m),
                                                   inside.ctypes.data_as(C.POINTER(C.c_int)), total.ctypes.data_as(C.POINTER(C.c_int))))
        return inside, total

    def setOption(self, name, value):
        self._check(self._L.lmb200_set_option(self._h, name.encode(), int(value)))

    def debugFetch(self, kind, slot=0, index=0):
        n = C.c_size_t(0)
        rc = self._L.lmb200_debug_fetch(self._h, kind, slot, index, None, C.byref(n))
        if rc not in (K.OK, K.E_TRUNCATED):
            self._check(rc)
        buf = np.zeros(max(1, n.value), np.uint8)
        n2 = C.c_size_t(buf.size)
        self._check(self._L.lmb200_debug_fetch(self._h, kind, slot, index, buf.ctypes.data, C.byref(n2)))
        buf = buf[:n2.value]
        if kind in (K.DBG_COARSE, K.DBG_UNSORTED):
            return _as_matches(buf.view(MATCH_DTYPE))
        if kind == K.DBG_MAGNITUDE:
            return buf.view(np.float32)
        if kind == K.DBG_DN_INDICES:
            return buf.view(np.int8)
        if kind == K.DBG_SIMILARITY:
            return buf.view(np.uint16)
        return buf

    def similarityMap(self, class_id, template_id, slot=0):
        """u16 coarse-level similarity map (upstream similarity() + addSimilarities()) of one template on the frame last
        matched in `slot`, from the production kernel with its early exit disabled (LMB200_DBG_SIMILARITY)."""
        g = 0
        for cid in self.classIds():
            if cid == class_id:
                break
            g += self.numTemplates(cid)
        else:
            raise KeyError(class_id)
        return self.debugFetch(K.DBG_SIMILARITY, slot, g + int(template_id))

    def normalLutIsStandin(self):
        return bool(self._L.lmb200_normal_lut_is_standin(self._h))

    def loadNormalLut(self, path):
        self._check(self._L.lmb200_load_normal_lut(self._h, str(path).encode()))

    def warnings(self):
        return (self._L.lmb200_warnings(self._h) or b"").decode()


def getDefaultLINE(**kw):
    """cv::linemod::getDefaultLINE(): ColorGradient, T={5,8}."""
    return Detector([ColorGradient()], (5, 8), **kw)


def getDefaultLINEMOD(**kw):
    """cv::linemod::getDefaultLINEMOD(): ColorGradient + DepthNormal, T={5,8} (reference: HighLevelLinemod.cpp:26-35)."""
    return Detector([ColorGradient(), DepthNormal()], (5, 8), **kw)


POSE_DTYPE = np.dtype([("translation", np.float32, 3), ("quaternion", np.float32, 4), ("bb", np.int32, 4),
                       ("median_depth", np.uint16), ("pad", np.uint16)])
assert POSE_DTYPE.itemsize == 48


def read_pose_sidecar(path, class_index):
    """linemod_tempPosFile.bin (reference: HighLevelLinemod.cpp:272-284) -> structured array of one class."""
    L = K.lib()
    n = C.c_size_t(0)
    rc = L.lmb200_read_pose_sidecar(str(path).encode(), class_index, None, 0, C.byref(n))
    if rc not in (K.OK, K.E_TRUNCATED):
        raise LinemodError(rc, "cannot read pose sidecar")
    out = np.zeros(n.value, POSE_DTYPE)
    if n.value:
        rc = L.lmb200_read_pose_sidecar(str(path).encode(), class_index, out.ctypes.data, n.value, C.byref(n))
        if rc:
            raise LinemodError(rc, "cannot read pose sidecar")
    return out


def write_pose_sidecar(path, per_class):
    L = K.lib()
    arrs = [np.ascontiguousarray(a, POSE_DTYPE) for a in per_class]
    ptrs = (C.c_void_p * max(1, len(arrs)))(*[a.ctypes.data for a in arrs])
    counts = (C.c_size_t * max(1, len(arrs)))(*[len(a) for a in arrs])
    rc = L.lmb200_write_pose_sidecar(str(path).encode(), ptrs, counts, len(arrs))
    if rc:
        raise LinemodError(rc, "cannot write pose sidecar")


def comm_unique_id():
    uid = np.zeros(128, np.uint8)
    rc = K.lib().lmb200_comm_unique_id(uid.ctypes.data)
    if rc:
        raise LinemodError(rc, (K.lib().lmb200_last_error(None) or b"").decode("utf-8", "replace"))
    return uid


def merge_matches(parts):
    """Host merge of per-rank generation-ordered match arrays (rank order) -> reference-ordered list."""
    L = K.lib()
    parts = [np.ascontiguousarray(p, MATCH_DTYPE) for p in parts]
    ptrs = (C.POINTER(K.MatchRec) * len(parts))(*[p.ctypes.data_as(C.POINTER(K.MatchRec)) for p in parts])
    counts = (C.c_size_t * len(parts))(*[len(p) for p in parts])
    total = sum(len(p) for p in parts)
    out = np.zeros(max(1, total), MATCH_DTYPE)
    n = C.c_size_t(0)
    rc = L.lmb200_merge_matches(ptrs, counts, len(parts), out.ctypes.data_as(C.POINTER(K.MatchRec)), len(out), C.byref(n))
    if rc:
        raise LinemodError(rc, "merge failed")
    return _as_matches(out[:n.value])


def shard_plan(costs, world):
    L = K.lib()
    costs = np.ascontiguousarray(costs, np.float64)
    begin = (C.c_int * (world + 1))()
    rc = L.lmb200_shard_plan(costs.ctypes.data_as(C.POINTER(C.c_double)), len(costs), world, begin)
    if rc:
        raise LinemodError(rc, "shard_plan failed")
    return list(begin)


def group_matches(matches, radius_threshold, discard_group_ratio):
    """groupSimilarMatches + discardSmallMatchGroups (reference: HighLevelLinemod.cpp:206-253) -> (group index per match or -1, n_groups)."""
    m = np.ascontiguousarray(matches, MATCH_DTYPE)
    out = np.zeros(len(m), np.int32); ng = C.c_int(0)
    rc = K.lib().lmb200_group_matches(m.ctypes.data_as(C.POINTER(K.MatchRec)), len(m), C.c_float(radius_threshold), C.c_float(discard_group_ratio),
                                      out.ctypes.data_as(C.POINTER(C.c_int)), C.byref(ng))
    if rc:
        raise LinemodError(rc, "lmb200_group_matches")
    return out, ng.value
