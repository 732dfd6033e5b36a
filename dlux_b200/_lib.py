"""ctypes binding of libdlux_b200.so (the C ABI in include/dlux_b200.h).

There is no CPU fallback: if the library is missing the import fails loudly and
tells the user how to build it (``python -m dlux_b200.build``)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdlux_b200.so")

PREC_3XTF32 = 0
PREC_FP32 = 1


class MftDesc(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("n_out", C.c_int32), ("batch", C.c_int32),
                ("inverse", C.c_int32), ("adjoint", C.c_int32), ("precision", C.c_int32),
                ("dft_period", C.c_int32), ("reserved", C.c_int32)]


class PolyPsfDesc(C.Structure):
    _fields_ = [("n_pupil", C.c_int32), ("n_psf", C.c_int32), ("n_wavels", C.c_int32),
                ("n_sources", C.c_int32), ("normalise", C.c_int32), ("precision", C.c_int32),
                ("save_field", C.c_int32), ("sparse", C.c_int32)]


class PolyPsfBatchDesc(C.Structure):
    _fields_ = [("n_pupil", C.c_int32), ("n_psf", C.c_int32), ("n_wavels", C.c_int32),
                ("n_batch", C.c_int32), ("n_basis", C.c_int32), ("normalise", C.c_int32),
                ("precision", C.c_int32), ("save_field", C.c_int32)]


class DluxError(RuntimeError):
    pass


_P = C.c_void_p
# name -> (restype, argtypes); must list every symbol include/dlux_b200.h declares
SIGNATURES = {
    "dlux_abi_version": (C.c_int, []),
    "dlux_error_string": (C.c_char_p, [C.c_int]),
    "dlux_last_cuda_error": (C.c_int, []),
    "dlux_launch_count": (C.c_uint64, []),
    "dlux_profile_enable": (C.c_int, [C.c_int]),
    "dlux_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    "dlux_tc_peak_probe": (C.c_int, [C.c_int32, C.c_int32, _P, C.POINTER(C.c_double), _P]),
    "dlux_mft_scratch_bytes": (C.c_size_t, [C.POINTER(MftDesc)]),
    "dlux_mft_c64": (C.c_int, [C.POINTER(MftDesc), _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "dlux_mft_coords": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "dlux_polypsf_scratch_bytes": (C.c_size_t, [C.POINTER(PolyPsfDesc)]),
    "dlux_polypsf_fwd": (C.c_int, [C.POINTER(PolyPsfDesc)] + [_P] * 10 + [_P, C.c_size_t, _P]),
    "dlux_polypsf_bwd": (C.c_int, [C.POINTER(PolyPsfDesc)] + [_P] * 17 + [_P, C.c_size_t, _P]),
    "dlux_polypsf_hvp": (C.c_int, [C.POINTER(PolyPsfDesc)] + [_P] * 13 + [_P, C.c_size_t, _P]),
    "dlux_polypsf_batch_scratch_bytes": (C.c_size_t, [C.POINTER(PolyPsfBatchDesc)]),
    "dlux_polypsf_batch_fwd": (C.c_int, [C.POINTER(PolyPsfBatchDesc)] + [_P] * 12 + [_P, C.c_size_t, _P]),
    "dlux_polypsf_batch_bwd": (C.c_int, [C.POINTER(PolyPsfBatchDesc)] + [_P] * 13 + [_P, C.c_size_t, _P]),
    "dlux_basis_eval": (C.c_int, [C.c_int32, C.c_int64, _P, _P, _P, _P, _P]),
    "dlux_basis_reduce": (C.c_int, [C.c_int32, C.c_int64, _P, _P, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"dlux_b200: native library not found at {LIB_PATH}. Build it with "
            "`python -m dlux_b200.build` (needs nvcc, targets sm_100a). There is no CPU or "
            "PyTorch fallback for the diffraction hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.dlux_abi_version() != 2:
        raise ImportError("dlux_b200: ABI version mismatch, rebuild the library")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        lib = load()
        msg = lib.dlux_error_string(rc).decode()
        raise DluxError(f"{what} failed: {msg} (code {rc}, cudaError {lib.dlux_last_cuda_error()})")


def launch_count() -> int:
    return int(load().dlux_launch_count())


def profile_enable(on: bool) -> None:
    load().dlux_profile_enable(int(on))


def profile_read():
    """(gemm_ms, gemm_launches, gemm_algorithmic_flops) since the last read."""
    ms, n, fl = C.c_double(), C.c_uint64(), C.c_double()
    load().dlux_profile_read(C.byref(ms), C.byref(n), C.byref(fl))
    return ms.value, int(n.value), fl.value


def tc_peak_probe(kind: int, n_batches: int, stream, sink=None) -> float:
    """Enqueue the MMA-only tensor-pipe probe (kind 0 tf32, 1 bf16, 2 the GEMM's mix) on `stream`;
    returns the real FLOPs the launch executes."""
    fl = C.c_double()
    check(load().dlux_tc_peak_probe(kind, n_batches, sink, C.byref(fl), stream), "dlux_tc_peak_probe")
    return fl.value
