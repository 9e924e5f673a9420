"""ctypes binding of ``include/gigpower_b200.h``.

The library is the product: there is no CPU fallback.  If
``libgigpower_b200.so`` is missing the import of this module raises, and every
solve raises if no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libgigpower_b200.so')


class GpError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build the CUDA library first "
        f"(python -m gigpower_b200.csrc.build, or __graft_entry__.build()). "
        f"gigpower_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

i32p, f64p, vp = C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_void_p


class GpFbsDesc(C.Structure):
    _fields_ = [
        ('n_steps', C.c_int32), ('n_vrows', C.c_int32), ('n_irows', C.c_int32),
        ('n_srows', C.c_int32), ('n_child', C.c_int32), ('n_lines', C.c_int32),
        ('bus_vrow', i32p), ('par_vrow', i32p), ('edge_irow', i32p), ('load_srow', i32p),
        ('kind', i32p), ('child_begin', i32p), ('child_count', i32p), ('child_irow', i32p),
        ('z', f64p), ('cappu', f64p), ('vr_gamma', f64p),
        ('root_vrow', C.c_int32 * 3), ('root_load_srow', C.c_int32 * 3),
        ('root_cappu', C.c_double * 3), ('vref', C.c_double * 6), ('zip', C.c_double * 3),
        ('line_irow', i32p), ('line_tx_vrow', i32p), ('line_rx_vrow', i32p),
    ]


class GpNr3Desc(C.Structure):
    _fields_ = [
        ('n_steps', C.c_int32), ('n_vrows', C.c_int32), ('n_irows', C.c_int32),
        ('n_srows', C.c_int32), ('n_child', C.c_int32), ('n_lines', C.c_int32),
        ('bus_vrow', i32p), ('par_vrow', i32p), ('edge_irow', i32p), ('edge_orow', i32p),
        ('load_srow', i32p), ('kind', i32p),
        ('child_begin', i32p), ('child_count', i32p), ('child_step', i32p), ('child_irow', i32p),
        ('z', f64p), ('cappu', f64p), ('vr_gamma', f64p),
        ('root_vrow', C.c_int32 * 3), ('root_load_srow', C.c_int32 * 3),
        ('root_cappu', C.c_double * 3), ('vslack', C.c_double * 6), ('zip', C.c_double * 3),
        ('cA', C.c_double * 3), ('cB', C.c_double * 3), ('f0_absent', C.c_double),
        ('line_irow', i32p), ('line_tx_vrow', i32p), ('line_rx_vrow', i32p),
    ]


def _sig(name, restype, argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


gp_version = _sig('gp_version', C.c_int, [])
gp_last_error = _sig('gp_last_error', C.c_int, [C.c_char_p, C.c_size_t])
gp_launch_count = _sig('gp_launch_count', C.c_int64, [])
gp_set_option = _sig('gp_set_option', C.c_int, [C.c_char_p, C.c_int])
gp_fbs_active_kernel = _sig('gp_fbs_active_kernel', C.c_int, [vp])
gp_fbs_onchip_warps = _sig('gp_fbs_onchip_warps', C.c_int, [vp])
gp_measure_fp64_peak = _sig('gp_measure_fp64_peak', C.c_int, [C.c_int, f64p])
gp_measure_dmma = _sig('gp_measure_dmma', C.c_int, [C.c_int, f64p])
gp_fbs_specialise = _sig('gp_fbs_specialise', C.c_int, [vp])
gp_fbs_schedule = _sig('gp_fbs_schedule', C.c_int, [C.POINTER(GpFbsDesc), i32p, i32p])
gp_fbs_walk_events = _sig('gp_fbs_walk_events', C.c_int, [C.POINTER(GpFbsDesc), i32p, C.c_int64])
gp_fbs_jit_source = _sig('gp_fbs_jit_source', C.c_int64, [C.POINTER(GpFbsDesc), C.c_int32, C.c_int32, C.c_char_p, C.c_int64])
gp_fbs_jit_warps = _sig('gp_fbs_jit_warps', C.c_int, [vp])
gp_jit_compile = _sig('gp_jit_compile', C.c_int, [C.c_char_p, vp, C.c_int64, C.POINTER(C.c_int64)])
gp_fbs_create = _sig('gp_fbs_create', C.c_int, [C.POINTER(GpFbsDesc), C.c_int, C.POINTER(vp)])
gp_feeder_destroy = _sig('gp_feeder_destroy', C.c_int, [vp])
gp_fbs_solve_batch = _sig('gp_fbs_solve_batch', C.c_int,
                          [vp, C.c_int64, C.c_int64, vp, vp, vp, vp, vp, vp, C.c_double, C.c_int32, vp])
gp_fbs_solve_post_batch = _sig('gp_fbs_solve_post_batch', C.c_int,
                               [vp, C.c_int64, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp, C.c_double, C.c_int32, vp])
gp_fbs_post = _sig('gp_fbs_post', C.c_int,
                   [vp, C.c_int64, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp, vp])
gp_fbs_solve_batch_host = _sig('gp_fbs_solve_batch_host', C.c_int,
                               [vp, C.c_int64, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp,
                                C.c_double, C.c_int32])



class GpSolveExtras(C.Structure):
    """Optional per-scenario inputs of a solve (include/gigpower_b200.h: GpSolveExtras)."""
    _fields_ = [('gamma_dev', vp), ('warm_start', C.c_int32), ('n_peers', C.c_int32),
                ('peer_vmag', C.POINTER(vp)), ('peer_loss', C.POINTER(vp)), ('wpu_dev', vp),
                ('vv_ctl_dev', vp), ('vv_points', C.c_double * 4), ('vv_q', C.c_double * 2)]


GP_MAX_PEERS = 16
GP_PEER_HANDLE_BYTES = 64
gp_peer_alloc = _sig('gp_peer_alloc', C.c_int, [C.c_int, C.c_int64, C.POINTER(vp), vp])
gp_peer_free = _sig('gp_peer_free', C.c_int, [C.c_int, vp])
gp_peer_open = _sig('gp_peer_open', C.c_int, [C.c_int, vp, C.POINTER(vp)])
gp_peer_close = _sig('gp_peer_close', C.c_int, [C.c_int, vp])
gp_peer_push = _sig('gp_peer_push', C.c_int, [C.c_int, C.c_int32, C.POINTER(vp), vp, C.c_int64, vp])


gp_gamma_rows = _sig('gp_gamma_rows', C.c_int, [vp])
gp_fbs_solve_batch_ex = _sig('gp_fbs_solve_batch_ex', C.c_int,
                             [vp, C.c_int64, C.c_int64, vp, C.POINTER(GpSolveExtras), vp, vp, vp, vp, vp, vp, vp,
                              C.c_double, C.c_int32, vp])

gp_fbs_post_ex = _sig('gp_fbs_post_ex', C.c_int,
                      [vp, C.c_int64, C.c_int64, vp, C.POINTER(GpSolveExtras), vp, vp, vp, vp, vp, vp, vp, vp])
gp_der_rows = _sig('gp_der_rows', C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                             vp, vp, vp, vp, vp, vp, vp])
gp_der_rows_strided = _sig('gp_der_rows_strided', C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                                             C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp, vp])
gp_vmag_summary = _sig('gp_vmag_summary', C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int32, vp, C.c_double,
                                                     C.c_double, vp, vp, vp, vp])

gp_nr3_create = _sig('gp_nr3_create', C.c_int, [C.POINTER(GpNr3Desc), C.c_int, C.POINTER(vp)])
gp_nr3_workspace_bytes = _sig('gp_nr3_workspace_bytes', C.c_int64, [vp, C.c_int64])
gp_nr3_solve_batch = _sig('gp_nr3_solve_batch', C.c_int,
                          [vp, C.c_int64, C.c_int64, vp, vp, vp, vp, vp, vp, C.c_int64, C.c_int32, vp])
gp_nr3_solve_batch_ex = _sig('gp_nr3_solve_batch_ex', C.c_int,
                             [vp, C.c_int64, C.c_int64, vp, C.POINTER(GpSolveExtras), vp, vp, vp, vp, vp, C.c_int64,
                              C.c_int32, vp])
gp_nr3_map_output = _sig('gp_nr3_map_output', C.c_int,
                         [vp, C.c_int64, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp])

gp_nr3_map_output_ex = _sig('gp_nr3_map_output_ex', C.c_int,
                            [vp, C.c_int64, C.c_int64, vp, C.POINTER(GpSolveExtras), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp])

EXPORTS = ['gp_version', 'gp_last_error', 'gp_launch_count', 'gp_set_option', 'gp_fbs_active_kernel',
           'gp_fbs_onchip_warps', 'gp_measure_fp64_peak', 'gp_fbs_specialise', 'gp_fbs_jit_source', 'gp_fbs_jit_warps',
           'gp_jit_compile', 'gp_fbs_create', 'gp_feeder_destroy',
           'gp_fbs_solve_batch', 'gp_fbs_solve_post_batch', 'gp_fbs_post', 'gp_fbs_solve_batch_host', 'gp_der_rows', 'gp_vmag_summary', 'gp_nr3_create',
           'gp_nr3_workspace_bytes', 'gp_nr3_solve_batch', 'gp_nr3_map_output',
           'gp_gamma_rows', 'gp_fbs_solve_batch_ex', 'gp_nr3_solve_batch_ex', 'gp_der_rows_strided',
           'gp_measure_dmma', 'gp_fbs_schedule', 'gp_fbs_walk_events', 'gp_nr3_map_output_ex', 'gp_peer_alloc', 'gp_peer_free', 'gp_peer_open', 'gp_peer_close', 'gp_peer_push', 'gp_fbs_post_ex']


def make_extras(gamma=None, warm_start=False, peers=None, wpu=None, volt_var=None):
    """GpSolveExtras for a device tensor of regulator ratios ``[n_gamma_rows, ld]`` float64, a warm
    start and / or the destinations of the fused all-gather (``peers`` = (|V| pointers, loss pointers),
    two equally long lists of device addresses); returns None when nothing is asked for."""
    if gamma is None and not warm_start and not peers and wpu is None and volt_var is None:
        return None
    ex = GpSolveExtras()
    if volt_var is not None:        # (int32 device tensor [n_vrows], (V1, V2, V3, V4), minQ, maxQ)
        ctl, points, minQ, maxQ = volt_var
        ex._vv_keep = ctl
        ex.vv_ctl_dev = ctl.data_ptr()
        ex.vv_points = (C.c_double * 4)(*[float(x) for x in points])
        ex.vv_q = (C.c_double * 2)(float(minQ), float(maxQ))
    ex.gamma_dev = gamma.data_ptr() if gamma is not None else None
    ex.wpu_dev = wpu.data_ptr() if wpu is not None else None
    ex.warm_start = 1 if warm_start else 0
    if peers:
        vm, ls = peers
        if len(vm) != len(ls) or not 0 < len(vm) <= GP_MAX_PEERS:
            raise ValueError(f'need 1..{GP_MAX_PEERS} peer destinations')
        ex._keep = ((vp * len(vm))(*[int(x) for x in vm]), (vp * len(ls))(*[int(x) for x in ls]))
        ex.n_peers = len(vm)
        ex.peer_vmag = C.cast(ex._keep[0], C.POINTER(vp))
        ex.peer_loss = C.cast(ex._keep[1], C.POINTER(vp))
    return ex


def check(rc: int):
    if rc != 0:
        buf = C.create_string_buffer(512)
        gp_last_error(buf, 512)
        raise GpError(f"gigpower_b200 error {rc}: {buf.value.decode(errors='replace')}")


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def make_fbs_desc(t):
    """FbsTables -> (GpFbsDesc, keep-alive list)."""
    keep = {}
    d = GpFbsDesc()
    d.n_steps, d.n_vrows, d.n_irows, d.n_srows = t.n_steps, t.n_vrows, t.n_irows, t.n_srows
    d.n_child, d.n_lines = int(t.child_irow.shape[0]), t.n_lines
    for name in ('bus_vrow', 'par_vrow', 'edge_irow', 'load_srow', 'kind', 'child_begin',
                 'child_count', 'child_irow', 'line_irow', 'line_tx_vrow', 'line_rx_vrow'):
        keep[name] = _i32(getattr(t, name))
        setattr(d, name, keep[name].ctypes.data_as(i32p))
    for name in ('z', 'cappu', 'vr_gamma'):
        keep[name] = _f64(getattr(t, name))
        setattr(d, name, keep[name].ctypes.data_as(f64p))
    d.root_vrow = (C.c_int32 * 3)(*[int(x) for x in t.root_vrow])
    d.root_load_srow = (C.c_int32 * 3)(*[int(x) for x in t.root_load_srow])
    d.root_cappu = (C.c_double * 3)(*[float(x) for x in t.root_cappu])
    d.vref = (C.c_double * 6)(*[float(x) for x in t.vref])
    d.zip = (C.c_double * 3)(*[float(x) for x in t.zip])
    return d, keep


def make_nr3_desc(t):
    """Nr3Tables -> (GpNr3Desc, keep-alive dict)."""
    keep = {}
    d = GpNr3Desc()
    d.n_steps, d.n_vrows, d.n_irows, d.n_srows = t.n_steps, t.n_vrows, t.n_irows, t.n_srows
    d.n_child, d.n_lines = int(t.child_step.shape[0]), t.n_lines
    for name in ('bus_vrow', 'par_vrow', 'edge_irow', 'edge_orow', 'load_srow', 'kind', 'child_begin',
                 'child_count', 'child_step', 'child_irow', 'line_irow', 'line_tx_vrow', 'line_rx_vrow'):
        keep[name] = _i32(getattr(t, name))
        setattr(d, name, keep[name].ctypes.data_as(i32p))
    for name in ('z', 'cappu', 'vr_gamma'):
        keep[name] = _f64(getattr(t, name))
        setattr(d, name, keep[name].ctypes.data_as(f64p))
    d.root_vrow = (C.c_int32 * 3)(*[int(x) for x in t.root_vrow])
    d.root_load_srow = (C.c_int32 * 3)(*[int(x) for x in t.root_load_srow])
    d.root_cappu = (C.c_double * 3)(*[float(x) for x in t.root_cappu])
    d.vslack = (C.c_double * 6)(*[float(x) for x in t.vslack])
    d.zip = (C.c_double * 3)(*[float(x) for x in t.zip])
    d.cA = (C.c_double * 3)(*[float(x) for x in t.cA])
    d.cB = (C.c_double * 3)(*[float(x) for x in t.cB])
    d.f0_absent = float(t.f0_absent)
    return d, keep


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise GpError("gigpower_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch
