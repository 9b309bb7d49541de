"""ctypes binding of libsimple_rf_b200.so (the C ABI declared in include/simple_rf_b200.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing or a call fails,
an exception is raised.  Tensors cross the boundary as raw device pointers + sizes + the current
CUDA stream handle.
"""
import ctypes
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint64, c_void_p
from pathlib import Path

import torch

import os

# SIMPLE_RF_B200_LIB selects an alternative build of the same ABI (kernel-tuning variants); default: the in-tree library
LIB_PATH = Path(os.environ.get('SIMPLE_RF_B200_LIB') or (Path(__file__).resolve().parent / 'libsimple_rf_b200.so'))
_P = c_void_p

_SIGNATURES = {
    'srf_last_error': (c_char_p, []),
    'srf_abi_version': (c_int, []),
    'srf_raygen': (c_int, [_P, c_int64, _P, _P, _P, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_int, c_int,
                           _P, _P, _P, _P, _P, _P]),
    'srf_stratified_z': (c_int, [_P, c_int, c_int64, _P, c_int, c_uint64, _P, _P]),
    'srf_box_march_z': (c_int, [_P, _P, c_int64, c_int, _P, c_float, c_float, c_float, _P, _P, _P]),
    'srf_sample_pdf_merge': (c_int, [_P, _P, _P, c_int64, c_uint64, c_int64, c_int, c_int, _P, _P, _P, _P, _P]),
    'srf_composite_fwd': (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, c_float,
                                  _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'srf_nerf_mlp_fwd': (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, _P, _P, _P, c_int, c_int, c_int, _P]),
    'srf_nerf_mlp_program_bytes': (c_int, []),
    'srf_mlp_set_pairing': (c_int, [c_int]),
    'srf_pack_alpha_bits': (c_int, [_P, c_int64, _P, _P]),
    'srf_compaction_blocks': (c_int, [c_int64]),
    'srf_tensorf_mask': (c_int, [_P, _P, _P, c_int64, c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    'srf_threshold_mask': (c_int, [_P, c_float, c_int64, _P, _P, _P]),
    'srf_patch_reprojection_masks': (c_int, [_P, _P, _P, _P, _P, c_int64, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, c_int,
                                             _P, _P, _P, _P, _P]),
    'srf_compact': (c_int, [_P, c_int64, _P, _P, _P, _P, _P]),
    'srf_vm_density_fwd': (c_int, [_P, _P, _P, c_int, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_int, c_float, _P, _P, _P]),
    'srf_vm_density_bwd': (c_int, [_P, _P, _P, c_int, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_int, c_float, _P, _P, _P, _P, _P]),
    'srf_vm_color_features_fwd': (c_int, [_P, _P, _P, c_int, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, _P]),
    'srf_alpha_corner_or_words': (c_int, [_P]),
    'srf_alpha_corner_or_bits': (c_int, [_P, _P, _P, _P]),
    'srf_tensorf_march': (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_float, c_float, c_float,
                                  _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'srf_tensorf_march_blocks': (c_int, [c_int64]),
    'srf_tensorf_march_compact': (c_int, [_P, c_int64, c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    'srf_ray_accumulate': (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, _P, _P]),
    'srf_vm_color_features_bwd': (c_int, [_P, _P, _P, c_int, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P]),
    'srf_cp_density_fwd': (c_int, [_P, _P, _P, c_int, _P, _P, c_int64, _P, _P, _P, c_int, _P, c_int, c_float, _P, _P, _P]),
    'srf_cp_density_bwd': (c_int, [_P, _P, _P, c_int, _P, _P, c_int64, _P, _P, _P, c_int, _P, c_int, c_float, _P, _P, _P, _P]),
    'srf_cp_color_features_fwd': (c_int, [_P, _P, _P, c_int, _P, _P, c_int64, _P, _P, _P, c_int, _P, _P, _P, c_int, _P]),
    'srf_cp_color_features_bwd': (c_int, [_P, _P, _P, c_int, _P, _P, c_int64, _P, _P, _P, c_int, _P, _P, c_int, _P, _P]),
    'srf_scatter_rows': (c_int, [_P, _P, c_int64, _P, c_int, _P, _P]),
    'srf_gather_rows': (c_int, [_P, _P, c_int64, _P, c_int, _P, _P]),
    'srf_mlp_rows_fwd': (c_int, [_P, _P, _P, _P, c_int, _P, c_int64, _P, _P, c_int, c_int, c_int, _P]),
    'srf_nerf_mlp_wgrad': (c_int, [_P, c_int, _P, c_int, _P, c_int, c_int64, _P, _P, _P]),
    'srf_wgrad_item_bytes': (c_int, []),
    'srf_nerf_mlp_dgrad': (c_int, [_P, _P, _P, _P, c_int, _P, _P, _P, _P, c_int64, _P, _P, c_int, _P, c_int, _P]),
    'srf_dgrad_program_bytes': (c_int, []),
    'srf_nerf_mlp_input_grad': (c_int, [_P, c_int, _P, _P, c_int, _P, _P, _P, _P, c_int64, _P, c_int, c_int, c_int, _P, _P, _P]),
    'srf_input_grad_source_bytes': (c_int, []),
    'srf_tv_loss': (c_int, [_P, _P, _P, c_int, c_float, _P, _P]),
    'srf_assemble_batch': (c_int, [_P, _P, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'srf_adam_step': (c_int, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_float, c_int64, _P]),
    'srf_frame_record_bytes': (c_int64, [c_int64, c_int]),
    'srf_frame_outputs': (c_int, [_P, _P, _P, _P, _P, c_int64, _P, _P]),
    'srf_adam_advance': (c_int, [_P, _P]),
    'srf_adam_step_capturable': (c_int, [_P, _P, _P, _P, c_int64, _P, c_float, c_float, c_float, c_float, _P, _P]),
    'srf_alpha_grid_words': (c_int, [_P]),
    'srf_alpha_grid_occupancy': (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_float, c_float, c_float, _P, _P]),
    'srf_alpha_grid_dilate': (c_int, [_P, _P, _P, _P, _P]),
    'srf_pack_alpha_bits_u8': (c_int, [_P, c_int64, _P, _P]),
    'srf_resample_plane': (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P]),
    'srf_composite_bwd': (c_int, [_P] * 17 + [c_int64, c_int, c_int, c_int, c_float, _P, _P, _P]),
}

_lib = None


class SimpleRFNativeError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise SimpleRFNativeError(
            f'{LIB_PATH} is missing: build it with `python -m simple_rf_b200.build` '
            '(nvcc, sm_100a). simple_rf_b200 has no CPU / PyTorch fallback.')
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def declared_symbols():
    return list(_SIGNATURES.keys())


_KEEP = []             # tensors whose raw pointers are in flight between ptr() and the end of the next call()


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor is kept alive until the next `call()` has returned:
    `ptr(f32c(x))` on a non-contiguous / non-fp32 `x` creates a temporary, and a temporary freed before the launch can be
    handed by the caching allocator to the NEXT temporary of the same size, aliasing two kernel arguments."""
    if t is None:
        return None
    _KEEP.append(t)
    return t.data_ptr()


def stream_handle():
    """Raw handle of torch's current stream on the current device (the private fast path: `torch.cuda.current_stream()` builds a
    Stream object per call, ~20 us — hundreds of launches per training iteration pay it)."""
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:
        return torch.cuda.current_stream().cuda_stream


LAUNCHES = {}          # entry point -> number of successful launches (bench.py reports the total)
TIMING = None          # when a list: (name, start_event, end_event, work) tuples appended per timed launch; `work` is the
                       # algorithmic work of the launch (FLOPs or bytes), either a float or (device count tensor, work per
                       # counted unit) when the unit count only exists on the device (resolve with resolve_work after a sync)


def resolve_work(work):
    if isinstance(work, tuple):
        count, per_unit = work
        return float(count.item()) * per_unit
    return float(work)


def call(name, *args, work=None):
    lib = load()
    timing = TIMING is not None and work is not None
    if timing:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    try:
        rc = getattr(lib, name)(*args)
    finally:
        _KEEP.clear()          # the launch is queued: stream order protects the buffers from here on
    if rc != 0:
        raise SimpleRFNativeError(lib.srf_last_error().decode())
    if timing:
        e1.record()
        TIMING.append((name, e0, e1, work))
    LAUNCHES[name] = LAUNCHES.get(name, 0) + 1


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise SimpleRFNativeError('simple_rf_b200 kernels take CUDA tensors only (no CPU fallback)')


def f32c(t):
    """Contiguous fp32 view/copy (None passes through)."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
