"""Headless view synthesis (host code in liblmb200.so): the software stand-in for the reference's
SDL + OpenGL offscreen renderer (`OpenGLRender`, src/OpenglRender.cpp) that feeds template generation
(src/HighLevelLinemod.cpp:68-110) and the benchmark's error renders (src/Benchmark.cpp:18-38,:156-170)."""
import ctypes as C
import numpy as np

from . import _capi as K


class RenderError(RuntimeError):
    def __init__(self, code, what):
        super().__init__("lmb200 error %d in %s" % (code, what))
        self.code = code


def load_ply(path):
    """ASCII PLY -> (vertices float64 [n,3], triangles int32 [m,3]); polygons are triangulated as fans."""
    L = K.lib()
    v = C.POINTER(C.c_double)(); t = C.POINTER(C.c_int)()
    nv = C.c_int(); nt = C.c_int()
    rc = L.lmb200_load_ply(str(path).encode(), C.byref(v), C.byref(nv), C.byref(t), C.byref(nt))
    if rc:
        raise RenderError(rc, "lmb200_load_ply(%s)" % path)
    try:
        verts = np.ctypeslib.as_array(v, shape=(nv.value, 3)).copy()
        tris = np.ctypeslib.as_array(t, shape=(max(nt.value, 1), 3))[:nt.value].copy()
    finally:
        L.lmb200_free(v); L.lmb200_free(t)
    return verts, tris.astype(np.int32)


class Camera:
    """Pinhole camera of the reference's renderer: fy for both axes, principal point at the image centre,
    near/far 100/10000 mm (src/OpenglRender.cpp:3-12)."""

    def __init__(self, width=640, height=480, fx=1045.69141, fy=1045.69141, cx=None, cy=None, near_mm=100.0, far_mm=10000.0):
        self.c = K.Camera(width, height, fx, fy, width / 2 if cx is None else cx, height / 2 if cy is None else cy, near_mm, far_mm)

    @property
    def size(self):
        return self.c.height, self.c.width


def _mesh(vertices, triangles):
    v = np.ascontiguousarray(vertices, np.float64).reshape(-1, 3)
    t = np.ascontiguousarray(triangles, np.int32).reshape(-1, 3)
    m = K.Mesh(v.ctypes.data_as(C.POINTER(C.c_double)), len(v), t.ctypes.data_as(C.POINTER(C.c_int)), len(t))
    return m, (v, t)


def _outputs(n, cam, depth, colour):
    h, w = cam.size
    d = np.empty((n, h, w), np.uint16) if depth else None
    c = np.empty((n, h, w, 3), np.uint8) if colour else None
    dp = d.ctypes.data_as(C.POINTER(C.c_uint16)) if depth else None
    cp = c.ctypes.data_as(C.POINTER(C.c_uint8)) if colour else None
    return d, c, dp, cp


def render_lookat(vertices, triangles, eyes, camera=None, depth=True, colour=True, threads=0):
    """Views from camera positions `eyes` [n,3] looking at the origin, +Y up (OpenGLRender::render*ToFrontBuff(model,
    camPosition)).  -> (depth u16 [n,h,w] in mm or None, colour u8 [n,h,w,3] or None)."""
    cam = camera or Camera()
    m, keep = _mesh(vertices, triangles)
    e = np.ascontiguousarray(eyes, np.float64).reshape(-1, 3)
    d, c, dp, cp = _outputs(len(e), cam, depth, colour)
    rc = K.lib().lmb200_render_lookat(C.byref(m), C.byref(cam.c), e.ctypes.data_as(C.POINTER(C.c_double)), len(e), dp, cp, threads)
    if rc:
        raise RenderError(rc, "lmb200_render_lookat")
    return d, c


def render_pose(vertices, triangles, rotations, translations, camera=None, depth=True, colour=True, threads=0):
    """Views with explicit model-view transforms x_cam = R x + t (the benchmark's overloads of the reference renderer)."""
    cam = camera or Camera()
    m, keep = _mesh(vertices, triangles)
    R = np.ascontiguousarray(rotations, np.float64).reshape(-1, 9)
    t = np.ascontiguousarray(translations, np.float64).reshape(-1, 3)
    if len(R) != len(t):
        raise ValueError("one translation per rotation expected")
    d, c, dp, cp = _outputs(len(R), cam, depth, colour)
    rc = K.lib().lmb200_render_pose(C.byref(m), C.byref(cam.c), R.ctypes.data_as(C.POINTER(C.c_double)),
                                    t.ctypes.data_as(C.POINTER(C.c_double)), len(R), dp, cp, threads)
    if rc:
        raise RenderError(rc, "lmb200_render_pose")
    return d, c


def hodan_error(input_depth, gt_render, est_render, visibility_threshold=15, error_threshold=20):
    """Benchmark::calculateErrorHodan (src/Benchmark.cpp:18-38, :133-154) on three u16 depth images (mm).
    -> (error, n_ok, n_comb); the reference counts a pose as correct when error < 0.3."""
    a = [np.ascontiguousarray(x, np.uint16) for x in (input_depth, gt_render, est_render)]
    if not (a[0].shape == a[1].shape == a[2].shape and a[0].ndim == 2):
        raise ValueError("three u16 images of one size expected")
    err = C.c_float(); ok = C.c_longlong(); comb = C.c_longlong()
    P = C.POINTER(C.c_uint16)
    rc = K.lib().lmb200_hodan_error(a[0].ctypes.data_as(P), a[1].ctypes.data_as(P), a[2].ctypes.data_as(P), a[0].shape[0], a[0].shape[1],
                                    visibility_threshold, error_threshold, C.byref(err), C.byref(ok), C.byref(comb))
    if rc:
        raise RenderError(rc, "lmb200_hodan_error")
    return err.value, ok.value, comb.value


def hodan_error_poses(vertices, triangles, R_gt, t_gt, R_est, t_est, input_depth, camera=None, visibility_threshold=15, error_threshold=20):
    """The whole call of PoseDetection.cpp:99: both renders (headless rasteriser) + the error."""
    cam = camera or Camera()
    m, keep = _mesh(vertices, triangles)
    R = np.ascontiguousarray(np.stack([np.asarray(R_gt, np.float64).reshape(9), np.asarray(R_est, np.float64).reshape(9)]))
    t = np.ascontiguousarray(np.stack([np.asarray(t_gt, np.float64).reshape(3), np.asarray(t_est, np.float64).reshape(3)]))
    d = np.ascontiguousarray(input_depth, np.uint16)
    err = C.c_float()
    rc = K.lib().lmb200_hodan_error_poses(C.byref(m), C.byref(cam.c), R.ctypes.data_as(C.POINTER(C.c_double)), t.ctypes.data_as(C.POINTER(C.c_double)),
                                          d.ctypes.data_as(C.POINTER(C.c_uint16)), visibility_threshold, error_threshold, C.byref(err))
    if rc:
        raise RenderError(rc, "lmb200_hodan_error_poses")
    return err.value
