"""ctypes binding of the C ABI in include/vdf_b200.h (libvdf_b200.so, built by csrc/Makefile).

There is no CPU fallback: if the shared library is missing or no sm_100 GPU is present, every entry point
raises.  numpy arrays are host buffers; `*_device` methods take raw device pointers (e.g. torch `data_ptr()`).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("VDF_B200_SO") or os.path.join(_HERE, "libvdf_b200.so")  # the override is for kernel experiments (scripts/)

OK = 0
ERR_CUDA, ERR_ALLOC, ERR_INVALID, ERR_EDGE_OVERFLOW, ERR_NO_DEVICE = -1, -2, -3, -4, -5
STACK_OK, STACK_NOT_ENOUGH_FRAMES, STACK_VIDPROC = 0, 1, 2
CROPDETECT_NONE, CROPDETECT_LETTERBOX, CROPDETECT_MOTION = 0, 1, 2
STACK_FLAG_MIXED_SIZES = 1

EXPORTS = [
    "vdf_version", "vdf_ctx_create", "vdf_ctx_destroy", "vdf_last_error", "vdf_ctx_set_shard", "vdf_ctx_set_option", "vdf_stage_sorted", "vdf_peer_alloc", "vdf_peer_open", "vdf_peer_close",
    "vdf_ctx_stream", "vdf_ctx_counters", "vdf_ctx_kernel_time", "vdf_search_self", "vdf_group_greedy", "vdf_search_self_groups",
    "vdf_search_refs", "vdf_search_self_device", "vdf_search_refs_device", "vdf_group_greedy_device",
    "vdf_self_window_pairs", "vdf_free_edges", "vdf_free_groups", "vdf_free_csr", "vdf_hash_stacks",
    "vdf_hash_stacks_device", "vdf_hash_stacks_small_device", "vdf_hash_from_small",
    "vdf_sort_order", "vdf_search", "vdf_search_with_references", "vdf_ctx_last_phases", "vdf_group_components",
    "vdf_group_components_device", "vdf_cache_load", "vdf_cache_save", "vdf_free_cache",
    "vdf_pipeline_create", "vdf_pipeline_push", "vdf_pipeline_flush", "vdf_pipeline_poll", "vdf_pipeline_error", "vdf_pipeline_destroy",
    "vdf_ctx_create_multi", "vdf_ctx_device_count", "vdf_sort_order_device", "vdf_table_create", "vdf_table_create_device", "vdf_table_destroy", "vdf_table_len",
    "vdf_table_search_self_device", "vdf_table_search_self_groups", "vdf_table_search_refs_device",
]


class VdfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vdf_b200 error {code}: {msg}")
        self.code = code


class Edges(C.Structure):
    _fields_ = [("n", C.c_uint64), ("ij", C.POINTER(C.c_uint64))]


class Groups(C.Structure):
    _fields_ = [("n_groups", C.c_uint64), ("group_ptr", C.POINTER(C.c_uint64)), ("member_idx", C.POINTER(C.c_uint64))]


class Csr(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("row_ptr", C.POINTER(C.c_uint64)), ("col_idx", C.POINTER(C.c_uint64))]


class StackDesc(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("frame_stride", C.c_uint64), ("width", C.c_uint32), ("height", C.c_uint32),
                ("pitch", C.c_uint32), ("n_frames", C.c_uint32), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


STACK_DESC_DTYPE = np.dtype([("offset", "<u8"), ("frame_stride", "<u8"), ("width", "<u4"), ("height", "<u4"),
                             ("pitch", "<u4"), ("n_frames", "<u4"), ("flags", "<u4"), ("reserved", "<u4")])
assert STACK_DESC_DTYPE.itemsize == C.sizeof(StackDesc) == 40


def build(force: bool = False) -> str:
    """Compile the CUDA sources for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", "Makefile"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "vdf_b200.h"))
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", csrc, "-j4"] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise RuntimeError(f"{_SO} is missing: build it with `make -C vid_dup_finder_lib_b200/csrc` "
                           "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    L = C.CDLL(_SO)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    L.vdf_version.restype = C.c_char_p
    L.vdf_ctx_create.argtypes = [i32, C.POINTER(vp)]
    L.vdf_ctx_create_multi.argtypes = [C.POINTER(i32), i32, C.POINTER(vp)]
    L.vdf_ctx_device_count.argtypes = [vp]
    L.vdf_table_create.argtypes = [vp, vp, vp, u64, C.POINTER(vp)]
    L.vdf_table_create_device.argtypes = [vp, vp, vp, u64, C.POINTER(vp)]
    L.vdf_table_destroy.argtypes = [vp]
    L.vdf_table_destroy.restype = None
    L.vdf_table_len.argtypes = [vp]
    L.vdf_table_len.restype = u64
    L.vdf_table_search_self_device.argtypes = [vp, u32, vp, u64, C.POINTER(u64)]
    L.vdf_table_search_self_groups.argtypes = [vp, u32, C.POINTER(Groups)]
    L.vdf_table_search_refs_device.argtypes = [vp, u64, vp, vp, u64, u32, vp, u64, C.POINTER(u64)]
    L.vdf_ctx_destroy.argtypes = [vp]
    L.vdf_ctx_destroy.restype = None
    L.vdf_last_error.argtypes = [vp]
    L.vdf_last_error.restype = C.c_char_p
    L.vdf_ctx_set_shard.argtypes = [vp, u32, u32]
    L.vdf_ctx_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.vdf_stage_sorted.argtypes = [vp, vp, vp, vp, vp, u64, vp, vp, vp, C.POINTER(vp), C.POINTER(vp)]
    L.vdf_peer_alloc.argtypes = [vp, u64, vp]
    L.vdf_peer_open.argtypes = [vp, u32, u32, vp]
    L.vdf_peer_close.argtypes = [vp]
    L.vdf_ctx_stream.argtypes = [vp]
    L.vdf_ctx_stream.restype = vp
    L.vdf_ctx_counters.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    L.vdf_ctx_counters.restype = None
    L.vdf_ctx_kernel_time.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(u64), i32]
    L.vdf_search_self.argtypes = [vp, vp, vp, u64, u32, C.POINTER(Edges)]
    L.vdf_group_greedy.argtypes = [vp, u64, C.POINTER(Edges), C.POINTER(Groups)]
    L.vdf_group_components.argtypes = [vp, u64, C.POINTER(Edges), C.POINTER(Groups)]
    L.vdf_group_components_device.argtypes = [vp, u64, vp, u64, C.POINTER(Groups)]
    L.vdf_search_self_groups.argtypes = [vp, vp, vp, u64, u32, C.POINTER(Groups)]
    L.vdf_search_refs.argtypes = [vp, vp, vp, u64, vp, vp, u64, u32, C.POINTER(Csr)]
    L.vdf_search_self_device.argtypes = [vp, vp, vp, u64, u32, vp, u64, C.POINTER(u64)]
    L.vdf_search_refs_device.argtypes = [vp, vp, vp, u64, u64, vp, vp, u64, u32, vp, u64, C.POINTER(u64)]
    L.vdf_group_greedy_device.argtypes = [vp, u64, vp, u64, vp, C.POINTER(Groups)]
    L.vdf_self_window_pairs.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.vdf_sort_order.argtypes = [vp, vp, vp, u64, vp]
    L.vdf_sort_order_device.argtypes = [vp, vp, vp, vp, u64, vp, vp]
    L.vdf_search.argtypes = [vp, vp, vp, vp, vp, u64, C.c_double, C.POINTER(Groups)]
    L.vdf_search_with_references.argtypes = [vp, vp, vp, u64, vp, vp, vp, vp, u64, C.c_double, C.POINTER(Csr)]
    L.vdf_ctx_last_phases.argtypes = [vp, C.POINTER(C.c_double)]
    for f in ("vdf_free_edges", "vdf_free_groups", "vdf_free_csr"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = None
    L.vdf_hash_stacks.argtypes = [vp, vp, vp, u32, i32, vp, vp, vp]
    L.vdf_hash_stacks_device.argtypes = [vp, vp, vp, u32, i32, vp, vp, vp]
    L.vdf_hash_stacks_small_device.argtypes = [vp, vp, vp, u32, i32, vp, vp]
    L.vdf_hash_from_small.argtypes = [vp, vp, u32, vp]
    _lib = L
    return L


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _hash_array(h) -> np.ndarray:
    h = np.ascontiguousarray(h, dtype=np.uint64)
    if h.ndim != 2 or h.shape[1] != 16:
        raise ValueError("hashes must be [n,16] uint64")
    return h


def _copy_u64(ptr, n: int) -> np.ndarray:
    if n == 0:
        return np.zeros(0, dtype=np.uint64)
    return np.ctypeslib.as_array(ptr, shape=(n,)).copy()


def sort_order(durations, path_blob: np.ndarray, path_off: np.ndarray) -> np.ndarray:
    """vdf_sort_order: Search::sort's permutation (stable, (duration, Rust Path order)); host-only, needs no GPU"""
    d = np.ascontiguousarray(durations, dtype=np.uint32)
    out = np.empty(len(d), dtype=np.uint64)
    rc = lib().vdf_sort_order(_ptr(d), _ptr(path_blob), _ptr(path_off), len(d), _ptr(out))
    if rc != 0:
        raise VdfError(rc, "vdf_sort_order")
    return out.astype(np.int64)


def make_descs(n: int, width: int, height: int, n_frames: int = 16, pitch: Optional[int] = None) -> np.ndarray:
    """descriptors for n same-sized stacks laid out back to back ([n][n_frames][height][pitch])"""
    pitch = pitch or width
    d = np.zeros(n, dtype=STACK_DESC_DTYPE)
    fs = pitch * height
    d["offset"] = np.arange(n, dtype=np.uint64) * np.uint64(fs * n_frames)
    d["frame_stride"] = fs
    d["width"], d["height"], d["pitch"], d["n_frames"] = width, height, pitch, n_frames
    return d


class _GroupsOwner:
    """keeps a vdf_groups result alive while numpy views of it exist, then returns it to the library"""

    def __init__(self, g: Groups):
        self._g = Groups(g.n_groups, g.group_ptr, g.member_idx)

    def __del__(self):
        try:
            lib().vdf_free_groups(C.byref(self._g))
        except Exception:
            pass


class Context:
    """One GPU (device: int) or several GPUs of one node driven from this process (device: a list of ids,
    vdf_ctx_create_multi); one stream per GPU, not thread-safe (see include/vdf_b200.h)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            ids = (C.c_int * len(device))(*[int(d) for d in device])
            rc = lib().vdf_ctx_create_multi(ids, len(device), C.byref(self._h))
            self.devices, device = [int(d) for d in device], int(device[0])
        else:
            rc = lib().vdf_ctx_create(int(device), C.byref(self._h))
            self.devices = [int(device)]
        if rc != OK:
            raise VdfError(rc, "vdf_ctx_create failed (no sm_100 device?) - this library has no CPU fallback")
        self.device = device
        self.peer_capacity, self.peer_world = 0, 0  # edge exchange over peer memory (peer_alloc / peer_open)

    @property
    def device_count(self) -> int:
        return int(lib().vdf_ctx_device_count(self._h))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().vdf_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != OK:
            raise VdfError(rc, lib().vdf_last_error(self._h).decode(errors="replace"))

    # ---- configuration
    def set_shard(self, rank: int, world: int):
        self._check(lib().vdf_ctx_set_shard(self._h, rank, world))

    def set_option(self, key: str, value: int):
        self._check(lib().vdf_ctx_set_option(self._h, key.encode(), int(value)))

    @property
    def stream_ptr(self) -> int:
        return int(lib().vdf_ctx_stream(self._h) or 0)

    def counters(self) -> Tuple[int, int, int]:
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib().vdf_ctx_counters(self._h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def kernel_time(self, which: int, reset: bool = False) -> Tuple[float, int]:
        """(total device ms, launches) of kernel slot 0 hamming, 1 resize, 2 letterbox, 3 dct+pack"""
        ms, cnt = C.c_double(), C.c_uint64()
        self._check(lib().vdf_ctx_kernel_time(self._h, which, C.byref(ms), C.byref(cnt), int(reset)))
        return ms.value, cnt.value

    # ---- search path, host buffers
    def search_self(self, hash_sorted, dur_sorted, tol_int: int) -> np.ndarray:
        h = _hash_array(hash_sorted)
        d = np.ascontiguousarray(dur_sorted, dtype=np.uint32)
        e = Edges()
        self._check(lib().vdf_search_self(self._h, _ptr(h), _ptr(d), len(d), int(tol_int), C.byref(e)))
        out = _copy_u64(e.ij, 2 * e.n).reshape(-1, 2)
        lib().vdf_free_edges(C.byref(e))
        return out

    def _groups(self, g: Groups):
        """(group_ptr, member_idx) as numpy views of the library's result block (no copy: at 1 M hashes the CSR is 2.7 MB);
        the block goes back to the library's pool when both arrays are gone"""
        ng = int(g.n_groups)
        if ng == 0:
            lib().vdf_free_groups(C.byref(g))
            return np.zeros(1, dtype=np.uint64), np.zeros(0, dtype=np.uint64)
        owner = _GroupsOwner(g)
        total = int(g.group_ptr[ng])
        gp_buf = (C.c_uint64 * (ng + 1)).from_address(C.addressof(g.group_ptr.contents))
        mm_buf = (C.c_uint64 * max(total, 1)).from_address(C.addressof(g.member_idx.contents))
        gp_buf._owner, mm_buf._owner = owner, owner
        return np.frombuffer(gp_buf, dtype=np.uint64), np.frombuffer(mm_buf, dtype=np.uint64)[:total]

    def stage_sorted(self, hashes, durations, path_blob: np.ndarray, path_off: np.ndarray, d_hash_dst: int = 0, d_dur_dst: int = 0):
        """vdf_stage_sorted: -> (order [n] int64, device pointer of the sorted hashes, device pointer of the sorted durations);
        d_hash_dst / d_dur_dst: caller-owned device buffers to upload into (0: buffers owned by the context)"""
        h = _hash_array(hashes)
        d = np.ascontiguousarray(durations, dtype=np.uint32)
        order = np.empty(len(d), dtype=np.uint64)
        ph, pd = C.c_void_p(), C.c_void_p()
        self._check(lib().vdf_stage_sorted(self._h, _ptr(h), _ptr(d), _ptr(path_blob), _ptr(path_off), len(d), _ptr(order),
                                           C.c_void_p(d_hash_dst or None), C.c_void_p(d_dur_dst or None), C.byref(ph), C.byref(pd)))
        return order.view(np.int64), int(ph.value or 0), int(pd.value or 0)

    def sort_order_device(self, durations, path_blob: np.ndarray, path_off: np.ndarray, d_order_out: int, d_dur_sorted_out: int = 0):
        """vdf_sort_order_device: Search::sort's permutation (n x u32) and the sorted durations, written to device memory"""
        d = np.ascontiguousarray(durations, dtype=np.uint32)
        self._check(lib().vdf_sort_order_device(self._h, _ptr(d), _ptr(path_blob), _ptr(path_off), len(d), d_order_out,
                                                C.c_void_p(d_dur_sorted_out or None)))

    def group_greedy(self, n: int, edges):
        e = np.ascontiguousarray(edges, dtype=np.uint64).reshape(-1, 2)
        st = Edges(len(e), e.ctypes.data_as(C.POINTER(C.c_uint64)))
        g = Groups()
        self._check(lib().vdf_group_greedy(self._h, int(n), C.byref(st), C.byref(g)))
        return self._groups(g)

    def group_components(self, n: int, edges):
        """connected components of the edge graph (optional mode, NOT the reference's grouping rule)"""
        e = np.ascontiguousarray(edges, dtype=np.uint64).reshape(-1, 2)
        st = Edges(len(e), e.ctypes.data_as(C.POINTER(C.c_uint64)))
        g = Groups()
        self._check(lib().vdf_group_components(self._h, int(n), C.byref(st), C.byref(g)))
        return self._groups(g)

    def search_self_groups(self, hash_sorted, dur_sorted, tol_int: int):
        h = _hash_array(hash_sorted)
        d = np.ascontiguousarray(dur_sorted, dtype=np.uint32)
        g = Groups()
        self._check(lib().vdf_search_self_groups(self._h, _ptr(h), _ptr(d), len(d), int(tol_int), C.byref(g)))
        return self._groups(g)

    def search_refs(self, cand_sorted, cand_dur_sorted, refs, ref_dur, tol_int: int):
        c = _hash_array(cand_sorted) if len(cand_sorted) else np.zeros((0, 16), np.uint64)
        cd = np.ascontiguousarray(cand_dur_sorted, dtype=np.uint32)
        r = _hash_array(refs) if len(refs) else np.zeros((0, 16), np.uint64)
        rd = np.ascontiguousarray(ref_dur, dtype=np.uint32)
        out = Csr()
        self._check(lib().vdf_search_refs(self._h, _ptr(c), _ptr(cd), len(cd), _ptr(r), _ptr(rd), len(rd), int(tol_int),
                                          C.byref(out)))
        rp = _copy_u64(out.row_ptr, len(rd) + 1)
        ci = _copy_u64(out.col_idx, int(rp[-1]))
        lib().vdf_free_csr(C.byref(out))
        return rp, ci

    def self_window_pairs(self, dur_sorted) -> int:
        d = np.ascontiguousarray(dur_sorted, dtype=np.uint32)
        v = C.c_uint64()
        self._check(lib().vdf_self_window_pairs(self._h, _ptr(d), len(d), C.byref(v)))
        return v.value

    # ---- the crate's public search functions, whole: inputs in the caller's order, indices of the caller's arrays out
    def search(self, hashes, durations, path_blob: np.ndarray, path_off: np.ndarray, tolerance: float):
        """vdf_search -> (group_ptr, member_idx): member_idx indexes the caller's arrays"""
        h = _hash_array(hashes) if len(durations) else np.zeros((0, 16), np.uint64)
        d = np.ascontiguousarray(durations, dtype=np.uint32)
        g = Groups()
        self._check(lib().vdf_search(self._h, _ptr(h), _ptr(d), _ptr(path_blob), _ptr(path_off), len(d), float(tolerance),
                                     C.byref(g)))
        return self._groups(g)

    def search_with_references(self, ref_hashes, ref_durations, new_hashes, new_durations, new_path_blob: np.ndarray,
                               new_path_off: np.ndarray, tolerance: float):
        """vdf_search_with_references -> (row_ptr, col_idx): col_idx indexes the caller's new_hashes arrays"""
        r = _hash_array(ref_hashes) if len(ref_durations) else np.zeros((0, 16), np.uint64)
        rd = np.ascontiguousarray(ref_durations, dtype=np.uint32)
        c = _hash_array(new_hashes) if len(new_durations) else np.zeros((0, 16), np.uint64)
        cd = np.ascontiguousarray(new_durations, dtype=np.uint32)
        out = Csr()
        self._check(lib().vdf_search_with_references(self._h, _ptr(r), _ptr(rd), len(rd), _ptr(c), _ptr(cd), _ptr(new_path_blob),
                                                     _ptr(new_path_off), len(cd), float(tolerance), C.byref(out)))
        rp = _copy_u64(out.row_ptr, len(rd) + 1)
        ci = _copy_u64(out.col_idx, int(rp[-1]))
        lib().vdf_free_csr(C.byref(out))
        return rp, ci

    def last_phases(self):
        """ms of the last search()/search_with_references(): host sort, gather + H2D enqueue, device, index remap"""
        a = (C.c_double * 4)()
        self._check(lib().vdf_ctx_last_phases(self._h, a))
        return [float(x) for x in a]

    # ---- search path, device pointers
    def search_self_device(self, d_hash: int, d_dur: int, n: int, tol_int: int, d_keys_out: int, capacity: int) -> int:
        cnt = C.c_uint64()
        rc = lib().vdf_search_self_device(self._h, d_hash, d_dur, n, int(tol_int), d_keys_out, capacity, C.byref(cnt))
        if rc == ERR_EDGE_OVERFLOW:
            return -int(cnt.value)
        self._check(rc)
        return int(cnt.value)

    def search_refs_device(self, d_cand: int, d_cand_dur: int, n_cand: int, cand_base: int, d_refs: int, d_ref_dur: int,
                           n_ref: int, tol_int: int, d_keys_out: int, capacity: int) -> int:
        cnt = C.c_uint64()
        rc = lib().vdf_search_refs_device(self._h, d_cand, d_cand_dur, n_cand, cand_base, d_refs, d_ref_dur, n_ref,
                                          int(tol_int), d_keys_out, capacity, C.byref(cnt))
        if rc == ERR_EDGE_OVERFLOW:
            return -int(cnt.value)
        self._check(rc)
        return int(cnt.value)

    # ---- edge exchange over NVLink peer memory (include/vdf_b200.h: vdf_peer_*) ----
    def peer_alloc(self, capacity_keys: int) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(lib().vdf_peer_alloc(self._h, int(capacity_keys), buf))
        self.peer_capacity, self.peer_world = int(capacity_keys), 0
        return buf.raw

    def peer_open(self, rank: int, world: int, handles: bytes):
        assert len(handles) == 64 * world
        self._check(lib().vdf_peer_open(self._h, int(rank), int(world), C.c_char_p(handles)))
        self.peer_world = int(world)

    def peer_close(self):
        self._check(lib().vdf_peer_close(self._h))
        self.peer_capacity, self.peer_world = 0, 0

    def group_greedy_device(self, n: int, d_keys_sorted: int, n_edges: int, d_remap: int = 0):
        """d_remap: device pointer of n x u32, sorted position -> index to report (0: report sorted positions)"""
        g = Groups()
        self._check(lib().vdf_group_greedy_device(self._h, int(n), d_keys_sorted, int(n_edges), C.c_void_p(d_remap or None), C.byref(g)))
        return self._groups(g)

    # ---- prepared tables (`Search::from` once, `search_self` / `search_with_references` many times)
    def table_create(self, hash_sorted, dur_sorted) -> "Table":
        h = _hash_array(hash_sorted) if len(dur_sorted) else np.zeros((0, 16), np.uint64)
        d = np.ascontiguousarray(dur_sorted, dtype=np.uint32)
        t = C.c_void_p()
        self._check(lib().vdf_table_create(self._h, _ptr(h), _ptr(d), len(d), C.byref(t)))
        return Table(self, t)

    def table_create_device(self, d_hash_sorted: int, d_dur_sorted: int, n: int, keepalive=None) -> "Table":
        t = C.c_void_p()
        self._check(lib().vdf_table_create_device(self._h, d_hash_sorted, d_dur_sorted, int(n), C.byref(t)))
        return Table(self, t, keepalive)

    # ---- hashing path
    def hash_stacks(self, frames: np.ndarray, descs: np.ndarray, cropdetect: int = CROPDETECT_LETTERBOX):
        """frames: host u8 buffer; descs: STACK_DESC_DTYPE[n] -> (hash [n,16] u64, status [n] i32, crop [n,4] u32)"""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        descs = np.ascontiguousarray(descs, dtype=STACK_DESC_DTYPE)
        n = len(descs)
        out = np.zeros((n, 16), dtype=np.uint64)
        st = np.zeros(n, dtype=np.int32)
        crop = np.zeros((n, 4), dtype=np.uint32)
        self._check(lib().vdf_hash_stacks(self._h, _ptr(frames), _ptr(descs), n, int(cropdetect), _ptr(out), _ptr(st),
                                          _ptr(crop)))
        return out, st, crop

    def hash_stacks_device(self, d_frames: int, descs: np.ndarray, cropdetect: int, d_out_hash: int):
        descs = np.ascontiguousarray(descs, dtype=STACK_DESC_DTYPE)
        n = len(descs)
        st = np.zeros(n, dtype=np.int32)
        crop = np.zeros((n, 4), dtype=np.uint32)
        self._check(lib().vdf_hash_stacks_device(self._h, d_frames, _ptr(descs), n, int(cropdetect), d_out_hash, _ptr(st),
                                                 _ptr(crop)))
        return st, crop

    def hash_stacks_small_device(self, d_frames: int, descs: np.ndarray, cropdetect: int, d_out_small: int):
        descs = np.ascontiguousarray(descs, dtype=STACK_DESC_DTYPE)
        crop = np.zeros((len(descs), 4), dtype=np.uint32)
        self._check(lib().vdf_hash_stacks_small_device(self._h, d_frames, _ptr(descs), len(descs), int(cropdetect),
                                                       d_out_small, _ptr(crop)))
        return crop

    def hash_from_small(self, small: np.ndarray) -> np.ndarray:
        small = np.ascontiguousarray(small, dtype=np.uint8).reshape(-1, 4096)
        out = np.zeros((len(small), 16), dtype=np.uint64)
        self._check(lib().vdf_hash_from_small(self._h, _ptr(small), len(small), _ptr(out)))
        return out


class Table:
    """vdf_table: the sorted table resident in HBM in the pair kernels' layout (include/vdf_b200.h, "prepared tables")."""

    def __init__(self, ctx: Context, handle, keepalive=None):
        self.ctx, self._t, self._keepalive = ctx, handle, keepalive

    def close(self):
        if getattr(self, "_t", None) and self._t.value and self.ctx._h.value:
            lib().vdf_table_destroy(self._t)
        self._t = C.c_void_p()
        self._keepalive = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(lib().vdf_table_len(self._t))

    def search_self_device(self, tol_int: int, d_keys_out: int, capacity: int) -> int:
        cnt = C.c_uint64()
        rc = lib().vdf_table_search_self_device(self._t, int(tol_int), d_keys_out, capacity, C.byref(cnt))
        if rc == ERR_EDGE_OVERFLOW:
            return -int(cnt.value)
        self.ctx._check(rc)
        return int(cnt.value)

    def search_self_groups(self, tol_int: int):
        g = Groups()
        self.ctx._check(lib().vdf_table_search_self_groups(self._t, int(tol_int), C.byref(g)))
        return self.ctx._groups(g)

    def search_refs_device(self, cand_base: int, d_refs: int, d_ref_dur: int, n_ref: int, tol_int: int, d_keys_out: int, capacity: int) -> int:
        cnt = C.c_uint64()
        rc = lib().vdf_table_search_refs_device(self._t, int(cand_base), d_refs, d_ref_dur, int(n_ref), int(tol_int), d_keys_out, capacity,
                                                C.byref(cnt))
        if rc == ERR_EDGE_OVERFLOW:
            return -int(cnt.value)
        self.ctx._check(rc)
        return int(cnt.value)


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    """Process-wide context on LOCAL_RANK's GPU (one process per GPU)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx
