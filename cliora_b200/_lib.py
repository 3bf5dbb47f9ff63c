"""ctypes binding of libcliora_b200.so (the C ABI in include/cliora_b200.h).

The shared library is built in-tree by ``build()`` (nvcc, sm_100a only).  There is
no CPU fallback anywhere in this package: if the library is missing or a tensor is
not on a CUDA device the call raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, 'libcliora_b200.so')
SRC = os.path.join(_HERE, 'csrc', 'api.cu')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC']

_lock = threading.Lock()
_lib = None


class LevelPlan(Structure):
    _fields_ = [('fused', c_int32), ('splits', c_int32), ('cells', c_int32), ('column_slices', c_int32),
                ('slice_cols', c_int32), ('umma_n', c_int32), ('cells_per_tile', c_int32), ('tiles', c_int32),
                ('max_sentences_per_tile', c_int32), ('ring_bytes', c_int32), ('smem_bytes', c_int64)]


class Dims(Structure):
    _fields_ = [('B', c_int), ('n', c_int), ('D', c_int), ('R', c_int), ('share', c_int), ('flags', c_int)]


_W_FIELDS = ['W_leaf', 'b_leaf', 'W1', 'b1', 'W2', 'b2', 'Wb', 'root', 'oW1', 'ob1', 'oW2', 'ob2', 'oWb']


class Weights(Structure):
    _fields_ = [(k, c_void_p) for k in _W_FIELDS]


class WeightGrads(Structure):
    _fields_ = [(k, c_void_p) for k in _W_FIELDS]


_L_FIELDS = ['ws_floats', 'bws_floats', 'rows_in', 'rows_out', 'PI', 'Pin', 'Pout', 'q_in', 'nrm_in', 'nrm2_in',
             'att_in', 'nrm_out', 'leaf_t', 'Zin', 'Yin', 'Ein', 'Prin', 'Zout', 'Yout', 'Eout', 'Prout',
             'Wcat_in', 'Wcat_out', 'Gh_in', 'Gs_in', 'GP_in', 'Gh_out', 'Gs_out', 'GP_out', 'GA2', 'coef',
             'GE', 'GZ', 'splitk', 'gu', 'W2p', 'W2Tp', 'oW2p', 'oW2Tp', 'GPp', 'Hp', 'Mbin', 'Mbout', 'CSin', 'CSout',
             'GYp_in', 'GYp_out', 'GA', 'CM', 'db2acc', 'W2h', 'W2Th', 'oW2h', 'oW2Th']


class ProfileRow(Structure):
    _fields_ = [('name', ctypes.c_char * 48), ('launches', c_int64), ('ms', ctypes.c_double),
                ('flops', ctypes.c_double), ('bytes', ctypes.c_double)]


class Layout(Structure):
    _fields_ = [(k, c_int64) for k in _L_FIELDS]


def _sources():
    d = os.path.join(_HERE, 'csrc')
    return [os.path.join(d, f) for f in sorted(os.listdir(d))] + [os.path.join(_ROOT, 'include', 'cliora_b200.h')]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources() if os.path.exists(s))


def build_pytrees(force: bool = False) -> str:
    """Compile the CPython helper that turns backpointer tables into nested tuples (host glue, plain gcc)."""
    import sysconfig
    src = os.path.join(_HERE, 'csrc', 'pytrees.c')
    out = os.path.join(_HERE, '_pytrees' + (sysconfig.get_config_var('EXT_SUFFIX') or '.so'))
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    cmd = [os.environ.get('CC', 'gcc'), '-O2', '-shared', '-fPIC', '-I' + sysconfig.get_paths()['include'], src,
           '-o', out]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('gcc failed:\n' + ' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/api.cu for sm_100a into cliora_b200/libcliora_b200.so (nvcc cross-compiles without a GPU)."""
    build_pytrees(force)
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get('NVCC', 'nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB_PATH, SRC]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


def _declare(lib):
    vp, st = c_void_p, c_void_p
    lib.cliora_status_string.restype = c_char_p
    lib.cliora_status_string.argtypes = [c_int]
    lib.cliora_last_cuda_error.restype = c_char_p
    lib.cliora_abi_version.restype = c_int
    lib.cliora_launch_count.restype = c_int64
    lib.cliora_debug_ptr.argtypes = [c_int, c_void_p]
    lib.cliora_debug_ptr.restype = None
    lib.cliora_num_cells.restype = c_int64
    lib.cliora_num_cells.argtypes = [c_int]
    lib.cliora_level_offset.restype = c_int64
    lib.cliora_level_offset.argtypes = [c_int, c_int]
    lib.cliora_inside_index.argtypes = [c_int, c_int, POINTER(c_int64), POINTER(c_int64)]
    lib.cliora_outside_index.argtypes = [c_int, c_int, POINTER(c_int64), POINTER(c_int64)]
    lib.cliora_split_row_offset.restype = c_int64
    lib.cliora_split_row_offset.argtypes = [c_int, c_int, c_int, c_int]
    lib.cliora_chart_layout.argtypes = [POINTER(Dims), POINTER(Layout)]
    lib.cliora_level_plan_query.argtypes = [POINTER(Dims), c_int, c_int, c_int, POINTER(LevelPlan)]
    lib.cliora_inside_fwd.argtypes = [POINTER(Dims), POINTER(Weights), vp, vp, vp, vp, vp, vp, st]
    lib.cliora_outside_fwd.argtypes = [POINTER(Dims), POINTER(Weights), vp, vp, vp, vp, vp, st]
    lib.cliora_chart_bwd_begin.argtypes = [POINTER(Dims), vp, vp, vp, vp, vp, st]
    lib.cliora_outside_bwd.argtypes = [POINTER(Dims), POINTER(Weights), vp, vp, vp, vp, vp, vp,
                                       POINTER(WeightGrads), c_int, st]
    lib.cliora_inside_bwd.argtypes = [POINTER(Dims), POINTER(Weights), vp, vp, vp, vp, vp, vp, vp, vp, c_int, vp,
                                      vp, POINTER(WeightGrads), c_int, st]
    lib.cliora_atten_scores.argtypes = [c_int, c_int, c_int, c_int, vp, c_int64, vp, vp, st]
    lib.cliora_atten_max_fwd.argtypes = [c_int, c_int, c_int, c_int, vp, c_int64, vp, vp, vp, st]
    lib.cliora_atten_max_bwd.argtypes = [c_int, c_int, c_int, c_int, vp, c_int64, vp, vp, vp, vp, c_int64, vp, st]
    lib.cliora_contrastive_loss.argtypes = [c_int, c_int, c_int, vp, vp, vp, c_float, c_float, vp, vp, vp, vp, vp,
                                            st]
    lib.cliora_vg_loss.argtypes = [c_int, c_int, vp, c_float, vp, vp, vp, st]
    lib.cliora_cky.argtypes = [c_int, c_int, vp, vp, vp, st]
    lib.cliora_adam_table_bytes.restype = c_int64
    lib.cliora_adam_table_bytes.argtypes = [c_int]
    lib.cliora_adam_table_fill.argtypes = [c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                           POINTER(c_void_p), POINTER(c_int64), vp, POINTER(c_int64)]
    lib.cliora_adam_table_fill.restype = c_int
    lib.cliora_adam_step.argtypes = [vp, c_int, c_int64, c_float, c_float, c_float, c_float, c_float, vp, vp, st]
    lib.cliora_adam_step.restype = c_int
    lib.cliora_span_f1.argtypes = [c_int, c_int, c_int, vp, vp, vp, vp, st]
    lib.cliora_span_f1.restype = c_int
    lib.cliora_grounding_eval.argtypes = [c_int, c_int, c_int, c_int, vp, vp, vp, vp, c_float, vp, vp, vp, st]
    lib.cliora_grounding_eval.restype = c_int
    lib.cliora_gather_regions.argtypes = [c_int, c_int, c_int, c_int, vp, vp, vp, vp, vp, vp, vp, vp, st]
    lib.cliora_gather_regions.restype = c_int
    lib.cliora_tree_spans.argtypes = [c_int, c_int, vp, vp, vp, st]
    lib.cliora_tree_spans.restype = c_int
    lib.cliora_recon_ce_fwd.argtypes = [c_int, c_int, c_int, vp, vp, vp, vp, vp, st]
    lib.cliora_recon_ce_fwd.restype = c_int
    lib.cliora_recon_ce_bwd.argtypes = [c_int, c_int, c_int, vp, vp, vp, vp, vp, vp, vp, vp, st]
    lib.cliora_recon_ce_bwd.restype = c_int
    lib.cliora_linear.argtypes = [c_int, c_int, c_int, vp, vp, vp, c_int, vp, st]
    lib.cliora_matmul_nn.argtypes = [c_int, c_int, c_int, vp, vp, vp, c_int, st]
    lib.cliora_matmul_tn_scratch_floats.restype = c_int64
    lib.cliora_matmul_tn_scratch_floats.argtypes = [c_int, c_int, c_int]
    lib.cliora_matmul_tn.argtypes = [c_int, c_int, c_int, vp, vp, vp, c_int, vp, st]
    lib.cliora_split_tf32.argtypes = [vp, c_int64, vp, st]
    lib.cliora_split_tf32.restype = c_int
    lib.cliora_tc_linear.argtypes = [c_int, c_int, c_int, vp, vp, vp, c_int, vp, st]
    lib.cliora_tc_linear.restype = c_int
    lib.cliora_tc_matmul_tn_scratch_floats.restype = c_int64
    lib.cliora_tc_matmul_tn_scratch_floats.argtypes = [c_int, c_int, c_int]
    lib.cliora_tc_matmul_tn.argtypes = [c_int, c_int, c_int, vp, vp, vp, c_int, vp, st]
    lib.cliora_tc_matmul_tn.restype = c_int
    lib.cliora_tc_atten_max_fwd.argtypes = [c_int, c_int, c_int, c_int, vp, vp, vp, vp, st]
    lib.cliora_tc_atten_max_fwd.restype = c_int
    lib.cliora_debug_set.argtypes = [c_int, c_int]
    lib.cliora_debug_set.restype = None
    lib.cliora_profile_start.restype = None
    lib.cliora_profile_stop.restype = c_int
    lib.cliora_profile_stop.argtypes = [POINTER(ProfileRow), c_int]
    for name in ('cliora_inside_index', 'cliora_outside_index', 'cliora_chart_layout', 'cliora_inside_fwd',
                 'cliora_outside_fwd', 'cliora_chart_bwd_begin', 'cliora_outside_bwd', 'cliora_inside_bwd',
                 'cliora_atten_scores', 'cliora_atten_max_fwd', 'cliora_atten_max_bwd', 'cliora_contrastive_loss',
                 'cliora_vg_loss', 'cliora_cky', 'cliora_linear', 'cliora_matmul_nn', 'cliora_matmul_tn'):
        getattr(lib, name).restype = c_int


# every symbol include/cliora_b200.h declares (tests check the .so exports all of them)
EXPORTS = ['cliora_status_string', 'cliora_last_cuda_error', 'cliora_abi_version', 'cliora_num_cells',
           'cliora_level_offset', 'cliora_inside_index', 'cliora_outside_index', 'cliora_split_row_offset',
           'cliora_chart_layout', 'cliora_inside_fwd', 'cliora_outside_fwd', 'cliora_chart_bwd_begin',
           'cliora_outside_bwd', 'cliora_inside_bwd', 'cliora_atten_scores', 'cliora_atten_max_fwd',
           'cliora_atten_max_bwd', 'cliora_contrastive_loss', 'cliora_vg_loss', 'cliora_cky', 'cliora_linear',
           'cliora_matmul_nn', 'cliora_matmul_tn_scratch_floats', 'cliora_matmul_tn', 'cliora_launch_count',
           'cliora_profile_start', 'cliora_profile_stop', 'cliora_split_tf32', 'cliora_tc_linear', 'cliora_debug_set',
           'cliora_tc_matmul_tn_scratch_floats', 'cliora_tc_matmul_tn', 'cliora_tc_atten_max_fwd', 'cliora_recon_ce_fwd', 'cliora_recon_ce_bwd', 'cliora_tree_spans', 'cliora_span_f1', 'cliora_grounding_eval', 'cliora_gather_regions', 'cliora_adam_table_bytes',
           'cliora_adam_table_fill', 'cliora_adam_step', 'cliora_debug_ptr', 'cliora_level_plan_query']


def lib():
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        'cliora_b200: %s is missing. Build it first: python -c "import __graft_entry__ as g; '
                        'g.build()" (needs nvcc). There is no CPU fallback.' % LIB_PATH)
                handle = ctypes.CDLL(LIB_PATH)
                _declare(handle)
                _lib = handle
    return _lib


class ClioraError(RuntimeError):
    pass


def check(status: int, what: str):
    if status != 0:
        L = lib()
        msg = L.cliora_status_string(status).decode()
        if status == -3:
            msg += ': ' + L.cliora_last_cuda_error().decode()
        raise ClioraError('%s failed: %s' % (what, msg))


def ptr(t):
    """Device pointer of a contiguous fp32/uint8/int32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ClioraError('cliora_b200 kernels need CUDA tensors (got a %s tensor); there is no CPU path' % t.device)
    if not t.is_contiguous():
        raise ClioraError('cliora_b200 kernels need contiguous tensors')
    return t.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def layout(B, n, D, R, share) -> Layout:
    d = Dims(B, n, D, R, 1 if share else 0, 0)
    out = Layout()
    check(lib().cliora_chart_layout(ctypes.byref(d), ctypes.byref(out)), 'cliora_chart_layout')
    return out


def level_plan(B, n, D, R, level, outside=False, backward=False, share=True, flags=0) -> LevelPlan:
    """Tile plan of one chart level of the fused level kernels (include/cliora_b200.h: cliora_level_plan_query)."""
    d = Dims(B, n, D, R, 1 if share else 0, flags)
    out = LevelPlan()
    check(lib().cliora_level_plan_query(ctypes.byref(d), level, 1 if outside else 0, 1 if backward else 0,
                                        ctypes.byref(out)), 'cliora_level_plan_query')
    return out


def launch_count() -> int:
    return int(lib().cliora_launch_count())


def profile_start():
    lib().cliora_profile_start()


def profile_stop():
    """-> {kernel class: dict(launches, ms, flops, bytes)} summed since profile_start()."""
    rows = (ProfileRow * 64)()
    n = lib().cliora_profile_stop(rows, 64)
    return {rows[i].name.decode(): dict(launches=int(rows[i].launches), ms=rows[i].ms, flops=rows[i].flops,
                                        bytes=rows[i].bytes) for i in range(n)}
