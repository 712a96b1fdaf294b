"""ctypes binding of libd3f.so (include/d3f.h).  Plain pointers and sizes only.

There is no fallback: if the library is missing or cannot be loaded, every entry point raises
NativeLibraryError — the field query never runs on anything but the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

from . import build as _build

D3F_MAX_VIEWS = 16
D3F_MAX_KEYS = 8
D3F_MAX_PEERS = 8
D3F_IPC_HANDLE_BYTES = 64
D3F_F32, D3F_U8 = 0, 1
FLAG_EVAL_DIST, FLAG_RECIP_NORM = 1, 2
ABI_VERSION = 2
D3F_ETIMEOUT = -4

# every symbol include/d3f.h declares (tests check the built library exports exactly these)
SYMBOLS = ('d3f_eval', 'd3f_eval_host', 'd3f_release_scratch', 'd3f_eval_ordered', 'd3f_bin_workspace_bytes',
           'd3f_bin_order', 'd3f_sweep_select', 'd3f_comm_create', 'd3f_comm_connect', 'd3f_comm_destroy',
           'd3f_comm_status', 'd3f_eval_allgather', 'd3f_comm_broadcast', 'd3f_eval_backward', 'd3f_track_loss_grad',
           'd3f_track_update', 'd3f_track_step_supported', 'd3f_track_step', 'd3f_pca_project',
           'd3f_create_grid', 'd3f_abi_version', 'd3f_last_error', 'd3f_launch_count', 'd3f_last_variant',
           'd3f_sizeof_key', 'd3f_sizeof_obs')


class NativeLibraryError(RuntimeError):
    pass


class D3FError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f'libd3f error {code}: {msg}')
        self.code = code


class D3FObs(C.Structure):
    _fields_ = [('V', C.c_int32), ('H', C.c_int32), ('W', C.c_int32),
                ('pose', C.c_void_p), ('K', C.c_void_p), ('depth', C.c_void_p)]


class D3FKey(C.Structure):
    _fields_ = [('data', C.c_void_p), ('dtype', C.c_int32), ('h', C.c_int32), ('w', C.c_int32), ('C', C.c_int32),
                ('stride_v', C.c_int64), ('stride_y', C.c_int64), ('stride_x', C.c_int64),
                ('bias', C.c_void_p)]


class D3FTrack(C.Structure):
    _fields_ = [('t_in', C.c_void_p), ('r_in', C.c_void_p), ('t_out', C.c_void_p), ('r_out', C.c_void_p),
                ('m_t', C.c_void_p), ('v_t', C.c_void_p), ('m_r', C.c_void_p), ('v_r', C.c_void_p),
                ('last_pts', C.c_void_p), ('grad_pts', C.c_void_p), ('pts', C.c_void_p),
                ('n_inst', C.c_int32), ('n_pts', C.c_int32),
                ('step', C.c_float), ('lr', C.c_float), ('beta1', C.c_float), ('beta2', C.c_float), ('eps', C.c_float),
                ('reg_w', C.c_float)]


class D3FGrid(C.Structure):
    _fields_ = [('x', C.c_void_p), ('y', C.c_void_p), ('z', C.c_void_p),
                ('nx', C.c_int32), ('ny', C.c_int32), ('nz', C.c_int32)]


_lib: Optional[C.CDLL] = None


def lib_path() -> str:
    return os.environ.get('D3F_LIBRARY', _build.LIB_PATH)


def load() -> C.CDLL:
    """Load libd3f.so (building it first if sources are present and it is stale or missing)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if 'D3F_LIBRARY' not in os.environ and not _build.is_current():
        try:
            _build.build()
        except Exception as e:  # no nvcc on this box: use a prebuilt library if there is one
            if not os.path.exists(path):
                raise NativeLibraryError(f'libd3f.so is not built and cannot be built here: {e}') from e
            import warnings
            warnings.warn(f'{path} was built from different sources than the ones in {_build.CSRC} and cannot be '
                          f'rebuilt here ({e}); loading it anyway — struct layouts are checked below', RuntimeWarning)
    if not os.path.exists(path):
        raise NativeLibraryError(f'{path} does not exist; run `python -m d3fields_b200.build`')
    try:
        lib = C.CDLL(path)
    except OSError as e:
        raise NativeLibraryError(f'cannot load {path}: {e}') from e
    vp, i32, i64, u32, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_float, C.c_double
    lib.d3f_eval.argtypes = [C.POINTER(D3FObs), vp, i64, C.POINTER(D3FKey), i32, vp, vp,
                             C.POINTER(vp), C.POINTER(vp), u32, f32, vp]
    lib.d3f_eval.restype = C.c_int
    lib.d3f_eval_host.argtypes = [C.POINTER(D3FObs), vp, i64, C.POINTER(D3FKey), i32, vp, vp,
                                  C.POINTER(vp), u32, f32, vp]
    lib.d3f_eval_host.restype = C.c_int
    lib.d3f_release_scratch.restype = C.c_int
    lib.d3f_eval_ordered.argtypes = [C.POINTER(D3FObs), vp, i64, vp, C.POINTER(D3FKey), i32, vp, vp, C.POINTER(vp), u32, f32, vp]
    lib.d3f_eval_ordered.restype = C.c_int
    lib.d3f_bin_workspace_bytes.argtypes = [i64]
    lib.d3f_bin_workspace_bytes.restype = i64
    lib.d3f_bin_order.argtypes = [vp, i64, f32, vp, vp, i64, vp]
    lib.d3f_bin_order.restype = C.c_int
    lib.d3f_sweep_select.argtypes = [C.POINTER(D3FObs), C.POINTER(D3FGrid), vp, i64, C.POINTER(D3FKey), f32, f32, vp, vp,
                                     i64, vp, vp, vp, u32, f32, vp]
    lib.d3f_sweep_select.restype = C.c_int
    lib.d3f_comm_create.argtypes = [i32, i32, i64, i64, C.POINTER(vp), vp]
    lib.d3f_comm_create.restype = C.c_int
    lib.d3f_comm_connect.argtypes = [vp, vp]
    lib.d3f_comm_connect.restype = C.c_int
    lib.d3f_comm_destroy.argtypes = [vp]
    lib.d3f_comm_destroy.restype = C.c_int
    lib.d3f_comm_status.argtypes = [vp, vp]
    lib.d3f_comm_status.restype = C.c_int
    lib.d3f_eval_allgather.argtypes = [vp, C.POINTER(D3FObs), vp, i64, C.POINTER(D3FKey), i32, C.POINTER(vp),
                                       i64, i64, i64, u32, f32, vp, C.POINTER(vp), C.POINTER(vp)]
    lib.d3f_eval_allgather.restype = C.c_int
    lib.d3f_comm_broadcast.argtypes = [vp, vp, i64, i32, vp]
    lib.d3f_comm_broadcast.restype = C.c_int
    lib.d3f_track_loss_grad.argtypes = [vp, vp, vp, vp, i64, i32, f32, vp, vp, vp, vp]
    lib.d3f_track_loss_grad.restype = C.c_int
    lib.d3f_track_update.argtypes = [C.POINTER(D3FTrack), vp]
    lib.d3f_track_update.restype = C.c_int
    lib.d3f_track_step_supported.argtypes = [C.POINTER(D3FObs), C.POINTER(D3FKey)]
    lib.d3f_track_step_supported.restype = C.c_int
    lib.d3f_track_step.argtypes = [C.POINTER(D3FObs), C.POINTER(D3FKey), vp, C.POINTER(D3FTrack), f32, vp, vp, vp, u32, f32, vp]
    lib.d3f_track_step.restype = C.c_int
    lib.d3f_sizeof_key.restype = C.c_int
    lib.d3f_sizeof_obs.restype = C.c_int
    lib.d3f_eval_backward.argtypes = [C.POINTER(D3FObs), vp, i64, C.POINTER(D3FKey), i32, C.POINTER(vp), vp, vp, u32, f32, vp]
    lib.d3f_eval_backward.restype = C.c_int
    lib.d3f_pca_project.argtypes = [vp, i64, i32, vp, vp, i32, vp, vp]
    lib.d3f_pca_project.restype = C.c_int
    lib.d3f_create_grid.argtypes = [f64, f64, f64, f64, i32, i32, i32, vp, vp]
    lib.d3f_create_grid.restype = C.c_int
    lib.d3f_abi_version.restype = C.c_int
    lib.d3f_last_error.restype = C.c_char_p
    lib.d3f_launch_count.restype = C.c_int64
    lib.d3f_last_variant.argtypes = [i32]
    lib.d3f_last_variant.restype = C.c_char_p
    got = lib.d3f_abi_version()
    if got != ABI_VERSION:
        raise NativeLibraryError(f'{path}: ABI version {got}, this package expects {ABI_VERSION}')
    if lib.d3f_sizeof_key() != C.sizeof(D3FKey) or lib.d3f_sizeof_obs() != C.sizeof(D3FObs):
        raise NativeLibraryError(f'{path}: D3FKey/D3FObs are {lib.d3f_sizeof_key()}/{lib.d3f_sizeof_obs()} bytes in the '
                                 f'library but {C.sizeof(D3FKey)}/{C.sizeof(D3FObs)} in this binding (stale library?)')
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise D3FError(rc, load().d3f_last_error().decode(errors='replace'))


def _ptr_array(ptrs: Sequence[Optional[int]]):
    arr = (C.c_void_p * max(len(ptrs), 1))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def make_key(key: tuple) -> D3FKey:
    """(data_ptr, dtype, h, w, C[, bias_ptr[, (stride_v, stride_y, stride_x)]]) -> D3FKey; strides in elements,
    omitted = contiguous."""
    sv, sy, sx = key[6] if len(key) > 6 and key[6] is not None else (0, 0, 0)
    return D3FKey(key[0], key[1], key[2], key[3], key[4], sv, sy, sx, key[5] if len(key) > 5 else None)


def _keys_array(keys: Sequence[tuple]):
    arr = (D3FKey * max(len(keys), 1))()
    for i, key in enumerate(keys):
        arr[i] = make_key(key)
    return arr


def eval_device(V: int, H: int, W: int, pose: int, K: int, depth: int, pts: int, n: int,
                keys: Sequence[tuple], dist: int, valid: int, outs: Sequence[int],
                inters: Optional[Sequence[Optional[int]]], flags: int, mu: float, stream: int) -> None:
    """d3f_eval with raw device addresses.  keys: (data_ptr, dtype, h, w, C) tuples."""
    lib = load()
    obs = D3FObs(V, H, W, pose, K, depth)
    inter_arr = _ptr_array(inters) if inters is not None else None
    _check(lib.d3f_eval(C.byref(obs), pts, n, _keys_array(keys), len(keys), dist, valid,
                        _ptr_array(outs), inter_arr, flags, mu, stream))


def eval_host(V: int, H: int, W: int, pose: int, K: int, depth: int, pts_host: int, n: int,
              keys: Sequence[tuple], dist_host: int, valid_host: int, outs_host: Sequence[int],
              flags: int, mu: float, obs_stream: int = 0) -> None:
    lib = load()
    obs = D3FObs(V, H, W, pose, K, depth)
    _check(lib.d3f_eval_host(C.byref(obs), pts_host, n, _keys_array(keys), len(keys), dist_host, valid_host,
                             _ptr_array(outs_host), flags, mu, obs_stream))


def release_scratch() -> None:
    _check(load().d3f_release_scratch())


def eval_ordered(V: int, H: int, W: int, pose: int, K: int, depth: int, pts: int, n: int, order: int,
                 keys: Sequence[tuple], dist: int, valid: int, outs: Sequence[int], flags: int, mu: float,
                 stream: int) -> None:
    lib = load()
    obs = D3FObs(V, H, W, pose, K, depth)
    _check(lib.d3f_eval_ordered(C.byref(obs), pts, n, order, _keys_array(keys), len(keys), dist, valid,
                                _ptr_array(outs), flags, mu, stream))


def bin_workspace_bytes(n: int) -> int:
    return int(load().d3f_bin_workspace_bytes(n))


def bin_order(pts: int, n: int, cell: float, order: int, workspace: int, workspace_bytes: int, stream: int) -> None:
    _check(load().d3f_bin_order(pts, n, cell, order, workspace, workspace_bytes, stream))


def sweep_select(V: int, H: int, W: int, pose: int, K: int, depth: int, grid: Optional[tuple], pts: Optional[int],
                 n: int, mask_key: Optional[tuple], dist_thr: float, mask_thr: float, dist_out: Optional[int],
                 valid_out: Optional[int], capacity: int, sel_count: Optional[int], sel_index: Optional[int],
                 sel_inst: Optional[int], flags: int, mu: float, stream: int) -> None:
    """grid: (x_ptr, y_ptr, z_ptr, nx, ny, nz) or None."""
    lib = load()
    obs = D3FObs(V, H, W, pose, K, depth)
    g = D3FGrid(*grid) if grid is not None else None
    k = make_key(mask_key) if mask_key is not None else None
    _check(lib.d3f_sweep_select(C.byref(obs), C.byref(g) if g is not None else None, pts, n,
                                C.byref(k) if k is not None else None, dist_thr, mask_thr, dist_out, valid_out,
                                capacity, sel_count, sel_index, sel_inst, flags, mu, stream))


def comm_create(rank: int, world: int, capacity_points: int, staging_bytes: int):
    """-> (comm handle (int), 64-byte IPC handle of this rank's segment)."""
    lib = load()
    comm = C.c_void_p()
    handle = (C.c_ubyte * D3F_IPC_HANDLE_BYTES)()
    _check(lib.d3f_comm_create(rank, world, capacity_points, staging_bytes, C.byref(comm), handle))
    return comm.value, bytes(handle)


def comm_connect(comm: int, handles: bytes) -> None:
    buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
    _check(load().d3f_comm_connect(comm, buf))


def comm_destroy(comm: int) -> None:
    _check(load().d3f_comm_destroy(comm))


def comm_status(comm: int, stream: int) -> None:
    _check(load().d3f_comm_status(comm, stream))


def comm_broadcast(comm: int, buf: int, nbytes: int, root: int, stream: int) -> None:
    _check(load().d3f_comm_broadcast(comm, buf, nbytes, root, stream))


def eval_allgather(comm: int, V: int, H: int, W: int, pose: int, K: int, depth: int, pts: int, n: int,
                   keys: Sequence[tuple], outs: Sequence[int], base: int, block: int, stride: int,
                   flags: int, mu: float, stream: int):
    """-> (dist_all_ptr, valid_all_ptr): local addresses of the gathered arrays of this call."""
    lib = load()
    obs = D3FObs(V, H, W, pose, K, depth)
    d, v = C.c_void_p(), C.c_void_p()
    _check(lib.d3f_eval_allgather(comm, C.byref(obs), pts, n, _keys_array(keys), len(keys), _ptr_array(outs),
                                  base, block, stride, flags, mu, stream, C.byref(d), C.byref(v)))
    return d.value, v.value


def eval_backward(V: int, H: int, W: int, pose: int, K: int, depth: int, pts: int, n: int, keys: Sequence[tuple],
                  grad_outs: Sequence[Optional[int]], grad_dist: Optional[int], grad_pts: int, flags: int, mu: float,
                  stream: int) -> None:
    lib = load()
    obs = D3FObs(V, H, W, pose, K, depth)
    _check(lib.d3f_eval_backward(C.byref(obs), pts, n, _keys_array(keys), len(keys), _ptr_array(grad_outs),
                                 grad_dist, grad_pts, flags, mu, stream))


def track_loss_grad(feat: int, src: int, dist: int, valid: int, n: int, c: int, dist_w: float, g_feat: int, g_dist: int,
                    loss_terms: Optional[int], stream: int) -> None:
    _check(load().d3f_track_loss_grad(feat, src, dist, valid, n, c, dist_w, g_feat, g_dist, loss_terms, stream))


def track_update(stream: int, **kw) -> None:
    """d3f_track_update; keyword arguments are the fields of D3FTrack (device addresses / sizes / hyper-parameters)."""
    t = D3FTrack(**kw)
    _check(load().d3f_track_update(C.byref(t), stream))


def track_step_supported(V: int, H: int, W: int, pose: int, K: int, depth: int, key: tuple) -> bool:
    """Whether d3f_track_step (the one-launch Adam iteration) takes this observation / descriptor map."""
    obs = D3FObs(V, H, W, pose, K, depth)
    return bool(load().d3f_track_step_supported(C.byref(obs), _keys_array([key])))


def track_step(V: int, H: int, W: int, pose: int, K: int, depth: int, key: tuple, src: int, dist_w: float,
               grad_scratch: int, arrivals: int, loss_terms: Optional[int], flags: int, mu: float, stream: int, **kw) -> None:
    """d3f_track_step; keyword arguments are the fields of D3FTrack."""
    obs = D3FObs(V, H, W, pose, K, depth)
    t = D3FTrack(**kw)
    _check(load().d3f_track_step(C.byref(obs), _keys_array([key]), src, C.byref(t), dist_w, grad_scratch, arrivals,
                                 loss_terms, flags, mu, stream))


def pca_project(x: int, n: int, c: int, mean: int, comp: int, n_comp: int, y: int, stream: int) -> None:
    _check(load().d3f_pca_project(x, n, c, mean, comp, n_comp, y, stream))


def create_grid(x0: float, y0: float, z0: float, step: float, nx: int, ny: int, nz: int, pts: int, stream: int) -> None:
    _check(load().d3f_create_grid(x0, y0, z0, step, nx, ny, nz, pts, stream))


def launch_count() -> int:
    return int(load().d3f_launch_count())


def last_variant(k: int = 0) -> str:
    return load().d3f_last_variant(k).decode()
