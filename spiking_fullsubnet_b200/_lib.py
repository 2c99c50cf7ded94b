"""ctypes binding of libgsn_b200.so (C ABI: include/gsn_b200.h).

The library is built in-tree (`spiking_fullsubnet_b200/csrc/libgsn_b200.so`, see `__graft_entry__.build`).
There is NO fallback: if the shared object is missing or a call fails, a Python exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libgsn_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "gsn_b200.h")

GSN_OK, GSN_EINVAL, GSN_ECUDA, GSN_ENOSUP = 0, 1, 2, 3
OPT_PDL, OPT_F32_MAX_CTAS = 1, 2
BACKEND_AUTO, BACKEND_SIMT, BACKEND_TCGEN05, BACKEND_TCGEN05_I8 = 0, 1, 2, 3
BACKENDS = {"auto": BACKEND_AUTO, "simt": BACKEND_SIMT, "tcgen05": BACKEND_TCGEN05, "tcgen05_i8": BACKEND_TCGEN05_I8}

_p, _i, _f, _i64, _sz = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_size_t

# name -> (restype, argtypes); must list every GSN_API symbol of include/gsn_b200.h
SIGNATURES = {
    "gsn_abi_version": (_i, []),
    "gsn_last_error": (C.c_char_p, []),
    "gsn_bind_device": (_i, [_i]),
    "gsn_set_option": (_i, [_i, _i]),
    "gsn_device_info": (_i, [C.POINTER(_i)] * 4),
    "gsn_compress_mag": (_i, [_p, _p, _i, _i, _i, _i, _f, _p]),
    "gsn_subband_features": (_i, [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _p, _p, _f, _p]),
    "gsn_subband_rowsums": (_i, [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _p]),
    "gsn_linear_f32": (_i, [_p, _p, _p, _p, _p, _i, _i64, _i, _i, _p]),
    "gsn_linear_spikes": (_i, [_p, _p, _p, _p, _p, _i, _i64, _i, _i, _i, _p]),
    "gsn_layer_recurrence_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "gsn_layer_recurrence": (_i, [_p] * 11 + [_i] * 6 + [_p, _p]),
    "gsn_layer_recurrence_bits": (_i, [_p] * 12 + [_i] * 6 + [_p, _p]),
    "gsn_pack_spikes": (_i, [_p, _p, _i64, _i, _p]),
    "gsn_recurrence_stream": (_i, [_p, _p, _p, _i, _p, _i] + [_p] * 9 + [_p, C.c_uint, _p, _p]
                              + [_p, _p, _i, _p, C.c_uint] + [_i] * 4 + [_p, _p]),
    "gsn_spike_image_bytes": (_sz, [_i] * 3),
    "gsn_compress_spec": (_i, [_p, _p, _i, _i, _i, _i, _f, _i, _p]),
    "gsn_deepfilter_spec": (_i, [_p, _p, _p, _p] + [_i] * 11 + [_p]),
    "gsn_spec_passthrough": (_i, [_p, _p, _p] + [_i] * 7 + [_p]),
    "gsn_overlap_add": (_i, [_p, _p, _p] + [_i] * 5 + [_p]),
    "gsn_frame_signal": (_i, [_p, _p, _p] + [_i] * 5 + [_p]),
    "gsn_stft_compress": (_i, [_p, _p, _p, _p] + [_i] * 6 + [_f, _p]),
    "gsn_irfft_frames": (_i, [_p, _p, _i, _i, _i, _p]),
    "gsn_deepfilter_irfft": (_i, [_p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _i, _i, _p]),
    "gsn_stream_preload": (_i, []),
    "gsn_xplanes_bytes": (_sz, [_i] * 4),
    "gsn_xplanes_stream": (_i, [_p, _i, _p, _i, _p, _p, _f, _p, _i, _p, _p, _i, _p, C.c_uint, _p, _p, C.c_uint] + [_i] * 8 + [_p]),
    "gsn_linear_spike_bits_stream": (_i, [_p] * 5 + [_i] * 6 + [_p, C.c_uint, _p, _p]),
    "gsn_pre_stream_supported": (_i, [_i, _i]),
    "gsn_pre_stream": (_i, [_p, _i, _p, _i, _p, _p, _f, _p, _p, _p, _p, C.c_uint, _p] + [_i] * 8 + [_p]),
    "gsn_recurrence_stream_tile": (_i, [_i] * 5),
    "gsn_recurrence_stream_ctas": (_i, [_i] * 5),
    "gsn_linear_spike_bits": (_i, [_p, _p, _p, _p, _p, _i, _i64, _i, _i, _i, _p]),
    "gsn_layer_train_workspace_bytes": (_sz, [_i, _i, _i]),
    "gsn_layer_train_forward": (_i, [_p] * 13 + [_i] * 5 + [_f, _f, _p, _p]),
    "gsn_layer_train_tc_supported": (_i, [_i, _i, _i]),
    "gsn_layer_train_tc_workspace_bytes": (_sz, [_i, _i]),
    "gsn_layer_train_forward_tc": (_i, [_p] * 13 + [_i] * 4 + [_f, _f, _i, _p, _p]),
    "gsn_layer_train_backward": (_i, [_p] * 13 + [_i] * 5 + [_f, _p, _p]),
    "gsn_layer_recurrence_pick_backend": (_i, [_i, _i, _i]),
    "gsn_layer_recurrence_tile": (_i, [_i, _i, _i, _i, _i]),
    "gsn_deepfilter_band": (_i, [_p] * 5 + [_i] * 9 + [_p]),
    "gsn_trace_set": (_i, [_p, _sz]),
}

_lib = None


class GsnError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GsnError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"or `make -C spiking_fullsubnet_b200/csrc`. spiking_fullsubnet_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    if lib.gsn_abi_version() != 1:
        raise GsnError(f"ABI version mismatch: library {lib.gsn_abi_version()} != binding 1")
    _lib = lib
    return lib


def load_probe():
    """The development-probe library (tools/tc_probe*.py, tools/tc_mma_timing.py): tcgen05 operand-encoding self test and
    MMA timing kernels.  Built next to the product library, never loaded by the package itself."""
    path = os.path.join(os.path.dirname(LIB_PATH), "libgsn_b200_probe.so")
    if not os.path.exists(path):
        raise GsnError(f"{path} not found: run `make -C spiking_fullsubnet_b200/csrc`")
    lib = C.CDLL(path)
    lib.gsn_last_error.restype = C.c_char_p
    lib.gsn_tc_selftest.restype = _i
    lib.gsn_tc_selftest.argtypes = [_p] * 4 + [_i] * 5 + [_p]
    return lib


_EXC = {GSN_EINVAL: ValueError, GSN_ECUDA: GsnError, GSN_ENOSUP: NotImplementedError}


def check(rc):
    if rc != GSN_OK:
        msg = load().gsn_last_error().decode("utf-8", "replace")
        raise _EXC.get(rc, GsnError)(msg)
