"""ctypes binding of the CPU ORACLE (oracle/vdf_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under vid_dup_finder_lib_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvdf_oracle.so")

OK, NOT_ENOUGH_FRAMES, VIDPROC = 0, 1, 2
LB_BLACKWHITE, LB_ANYCOLOUR = 0, 1


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("vdf_oracle.c", "vdf_oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libvdf_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.vdfo_hamming.restype = C.c_uint32
        L.vdfo_hamming.argtypes = [C.c_void_p, C.c_void_p]
        L.vdfo_tolerance_int.restype = C.c_uint32
        L.vdfo_tolerance_int.argtypes = [C.c_double]
        L.vdfo_self_window_thresh.restype = C.c_uint32
        L.vdfo_self_window_thresh.argtypes = [C.c_uint32]
        L.vdfo_ref_window_durations.restype = None
        L.vdfo_ref_window_durations.argtypes = [C.c_uint32, _u32p, _u32p]
        L.vdfo_path_cmp.restype = C.c_int
        L.vdfo_path_cmp.argtypes = [C.c_char_p, C.c_char_p]
        L.vdfo_sort_order.restype = None
        L.vdfo_sort_order.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_uint64, C.c_void_p]
        L.vdfo_search_self.restype = C.c_int
        L.vdfo_search_self.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(_u64p),
                                       C.POINTER(_u64p), _u64p]
        L.vdfo_self_edges.restype = C.c_int
        L.vdfo_self_edges.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(_u64p), _u64p]
        L.vdfo_group_from_edges.restype = C.c_int
        L.vdfo_group_from_edges.argtypes = [C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(_u64p), C.POINTER(_u64p),
                                            _u64p]
        L.vdfo_search_refs.restype = C.c_int
        L.vdfo_search_refs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                       C.c_uint32, C.POINTER(_u64p), C.POINTER(_u64p)]
        L.vdfo_self_window_pairs.restype = C.c_uint64
        L.vdfo_self_window_pairs.argtypes = [C.c_void_p, C.c_uint64]
        L.vdfo_free.restype = None
        L.vdfo_free.argtypes = [C.c_void_p]
        L.vdfo_letterbox_frame.restype = None
        L.vdfo_letterbox_frame.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_int, C.c_uint8,
                                           C.c_void_p]
        L.vdfo_cropdetect_letterbox.restype = C.c_int
        L.vdfo_cropdetect_letterbox.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_size_t,
                                                C.c_size_t, C.c_void_p]
        L.vdfo_resize_coeffs.restype = C.c_int
        L.vdfo_resize_coeffs.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(_u32p), C.POINTER(C.POINTER(C.c_int16)),
                                         _u32p, _u32p]
        L.vdfo_resize_lanczos3.restype = C.c_int
        L.vdfo_resize_lanczos3.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_uint32, C.c_uint32,
                                           C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.vdfo_dct2_16.restype = None
        L.vdfo_dct2_16.argtypes = [C.c_void_p]
        L.vdfo_dct3d.restype = None
        L.vdfo_dct3d.argtypes = [C.c_void_p]
        L.vdfo_hash_from_small.restype = None
        L.vdfo_hash_from_small.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.vdfo_hash_stack.restype = C.c_int
        L.vdfo_hash_stack.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_size_t, C.c_size_t,
                                      C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.vdfo_hash_stacks.restype = C.c_int
        L.vdfo_hash_stacks.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p,
                                       C.c_void_p]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _hashes(h) -> np.ndarray:
    h = np.ascontiguousarray(h, dtype=np.uint64)
    assert h.ndim == 2 and h.shape[1] == 16
    return h


def _take(ptr, n, dtype=np.uint64) -> np.ndarray:
    out = np.ctypeslib.as_array(ptr, shape=(max(int(n), 1),))[: int(n)].astype(dtype, copy=True)
    lib().vdfo_free(ptr)
    return out


# ------------------------------------------------------------------------------ search path
def hamming(x, y) -> int:
    x = np.ascontiguousarray(x, dtype=np.uint64)
    y = np.ascontiguousarray(y, dtype=np.uint64)
    return int(lib().vdfo_hamming(_p(x), _p(y)))


def tolerance_int(tol: float) -> int:
    return int(lib().vdfo_tolerance_int(float(tol)))


def self_window_thresh(d: int) -> int:
    return int(lib().vdfo_self_window_thresh(int(d)))


def ref_window_durations(d: int):
    lo, hi = C.c_uint32(), C.c_uint32()
    lib().vdfo_ref_window_durations(int(d), C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def path_cmp(a: str, b: str) -> int:
    return int(lib().vdfo_path_cmp(os.fsencode(a), os.fsencode(b)))


def sort_order(durations, paths) -> np.ndarray:
    d = np.ascontiguousarray(durations, dtype=np.uint32)
    n = len(d)
    arr = (C.c_char_p * max(n, 1))(*[os.fsencode(p) for p in paths])
    out = np.zeros(n, dtype=np.uint64)
    lib().vdfo_sort_order(_p(d), arr, n, _p(out))
    return out


def _csr(gp, mm, ng):
    ngv = int(ng.value)
    gpa = np.ctypeslib.as_array(gp, shape=(ngv + 1,)).copy()
    total = int(gpa[-1])
    lib().vdfo_free(gp)
    mma = _take(mm, total)
    return gpa, mma


def search_self(hashes_sorted, dur_sorted, tol_int: int):
    """-> (group_ptr, members): members are indices into the sorted array, target last."""
    h = _hashes(hashes_sorted)
    d = np.ascontiguousarray(dur_sorted, dtype=np.uint32)
    gp, mm, ng = _u64p(), _u64p(), C.c_uint64()
    rc = lib().vdfo_search_self(_p(h), _p(d), len(d), int(tol_int), C.byref(gp), C.byref(mm), C.byref(ng))
    assert rc == 0
    return _csr(gp, mm, ng)


def self_edges(hashes_sorted, dur_sorted, tol_int: int) -> np.ndarray:
    h = _hashes(hashes_sorted)
    d = np.ascontiguousarray(dur_sorted, dtype=np.uint32)
    e, ne = _u64p(), C.c_uint64()
    rc = lib().vdfo_self_edges(_p(h), _p(d), len(d), int(tol_int), C.byref(e), C.byref(ne))
    assert rc == 0
    return _take(e, 2 * ne.value).reshape(-1, 2)


def group_from_edges(n: int, edges):
    e = np.ascontiguousarray(edges, dtype=np.uint64).reshape(-1, 2)
    gp, mm, ng = _u64p(), _u64p(), C.c_uint64()
    rc = lib().vdfo_group_from_edges(int(n), _p(e), len(e), C.byref(gp), C.byref(mm), C.byref(ng))
    assert rc == 0
    return _csr(gp, mm, ng)


def group_components(n: int, edges):
    """Connected components of the edge graph, sequentially, in the way the app's DisjointSet accumulates pairs
    (vid_dup_finder_app/src/app/disjoint_set.rs:22-44: new set / append / merge), written out in the library's
    convention for its optional components mode: members ascending, the smallest entry last, groups by descending
    smallest entry.  -> (group_ptr, member_idx)"""
    set_of = {}
    sets = []
    for i, j in np.asarray(edges, dtype=np.uint64).reshape(-1, 2).tolist():
        a, b = set_of.get(i), set_of.get(j)
        if a is not None and a is b:
            continue
        if a is None and b is None:
            t = {i, j}
            sets.append(t)
        elif a is None or b is None:
            t = a if b is None else b
            t.update((i, j))
        else:
            t, drop = (a, b) if len(a) >= len(b) else (b, a)
            t.update(drop)
            for v in drop:
                set_of[v] = t
            drop.clear()
        set_of[i] = set_of[j] = t
    comps = sorted((sorted(t) for t in sets if t), key=lambda c: -c[0])
    gp, mm = [0], []
    for c in comps:
        mm.extend(c[1:] + c[:1])
        gp.append(len(mm))
    return np.array(gp, dtype=np.uint64), np.array(mm, dtype=np.uint64)


def search_refs(cand_sorted, cand_dur_sorted, refs, ref_dur, tol_int: int):
    c = _hashes(cand_sorted)
    cd = np.ascontiguousarray(cand_dur_sorted, dtype=np.uint32)
    r = _hashes(refs) if len(refs) else np.zeros((0, 16), np.uint64)
    rd = np.ascontiguousarray(ref_dur, dtype=np.uint32)
    rp, ci = _u64p(), _u64p()
    rc = lib().vdfo_search_refs(_p(c), _p(cd), len(cd), _p(r), _p(rd), len(rd), int(tol_int), C.byref(rp),
                                C.byref(ci))
    assert rc == 0
    rpa = np.ctypeslib.as_array(rp, shape=(len(rd) + 1,)).copy()
    lib().vdfo_free(rp)
    return rpa, _take(ci, rpa[-1])


def self_window_pairs(dur_sorted) -> int:
    d = np.ascontiguousarray(dur_sorted, dtype=np.uint32)
    return int(lib().vdfo_self_window_pairs(_p(d), len(d)))


# ------------------------------------------------------------------------------ hashing path
def letterbox_frame(img: np.ndarray, mode: int = LB_ANYCOLOUR, tol: int = 16):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    out = np.zeros(4, dtype=np.uint32)
    lib().vdfo_letterbox_frame(_p(img), w, h, w, mode, tol, _p(out))
    return tuple(int(v) for v in out)  # (left, right, top, bottom)


def cropdetect_letterbox(frames: np.ndarray):
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    n, h, w = frames.shape
    out = np.zeros(4, dtype=np.uint32)
    st = lib().vdfo_cropdetect_letterbox(_p(frames), n, w, h, w, w * h, _p(out))
    return st, tuple(int(v) for v in out)


def resize_coeffs(in_size: int, out_size: int = 16):
    b, k = _u32p(), C.POINTER(C.c_int16)()
    win, prec = C.c_uint32(), C.c_uint32()
    rc = lib().vdfo_resize_coeffs(in_size, out_size, C.byref(b), C.byref(k), C.byref(win), C.byref(prec))
    assert rc == 0
    bounds = np.ctypeslib.as_array(b, shape=(out_size * 2,)).copy().reshape(out_size, 2)
    coefs = np.ctypeslib.as_array(k, shape=(out_size * win.value,)).copy().reshape(out_size, win.value)
    lib().vdfo_free(b)
    lib().vdfo_free(k)
    return bounds, coefs, int(prec.value)


def resize_lanczos3(img: np.ndarray, crop=None, out_w: int = 16, out_h: int = 16) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    left, top, cw, ch = crop if crop is not None else (0, 0, w, h)
    dst = np.zeros((out_h, out_w), dtype=np.uint8)
    rc = lib().vdfo_resize_lanczos3(_p(img), w, h, w, left, top, cw, ch, out_w, out_h, _p(dst))
    assert rc == 0
    return dst


def dct2_16(x) -> np.ndarray:
    b = np.ascontiguousarray(x, dtype=np.float64).copy()
    assert b.shape == (16,)
    lib().vdfo_dct2_16(_p(b))
    return b


def dct3d(cube) -> np.ndarray:
    b = np.ascontiguousarray(cube, dtype=np.float64).copy()
    assert b.shape == (16, 16, 16)
    lib().vdfo_dct3d(_p(b))
    return b


def hash_from_small(small: np.ndarray, want_coefs: bool = False):
    s = np.ascontiguousarray(small, dtype=np.uint8)
    assert s.shape == (16, 16, 16)
    h = np.zeros(16, dtype=np.uint64)
    co = np.zeros((16, 16, 16), dtype=np.float64) if want_coefs else None
    lib().vdfo_hash_from_small(_p(s), _p(h), _p(co) if want_coefs else None)
    return (h, co) if want_coefs else h


def hash_stack(frames: np.ndarray, cropdetect: int = 1, frame_dims=None):
    """frames [n,h,w] u8 -> (status, hash[16] u64, crop(l,r,t,b), small[16,16,16] u8)"""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    n, h, w = frames.shape
    hv = np.zeros(16, dtype=np.uint64)
    crop = np.zeros(4, dtype=np.uint32)
    small = np.zeros((16, 16, 16), dtype=np.uint8)
    fd = None
    if frame_dims is not None:
        fd = np.ascontiguousarray(frame_dims, dtype=np.uint32)
    st = lib().vdfo_hash_stack(_p(frames), n, w, h, w, w * h, int(cropdetect), _p(fd) if fd is not None else None,
                               _p(hv), _p(crop), _p(small))
    return int(st), hv, tuple(int(v) for v in crop), small


def hash_stacks(frames: np.ndarray, cropdetect: int = 1):
    """frames [n_stacks,16,h,w] u8 -> (hash [n,16] u64, status [n] i32); single thread"""
    frames = np.ascontiguousarray(frames, dtype=np.uint8)
    n, t, h, w = frames.shape
    assert t == 16
    hv = np.zeros((n, 16), dtype=np.uint64)
    st = np.zeros(n, dtype=np.int32)
    lib().vdfo_hash_stacks(_p(frames), n, w, h, int(cropdetect), _p(hv), _p(st))
    return hv, st
