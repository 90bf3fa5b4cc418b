"""ctypes binding of ``libcmhar_b200.so`` (the C ABI declared in ``include/cmhar_b200.h``).

The library holds every CUDA kernel of the hot path.  There is NO fallback: if the shared object
is missing or a call fails, a ``RuntimeError`` is raised -- the product path never routes around
the CUDA extension.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading
import weakref
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libcmhar_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "cmhar_b200.h")

FP32, BF16 = 0, 1
MAX_LAYERS = 8
MAX_SEQ = 16

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


class EncoderLayerParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "in_proj_weight", "in_proj_bias", "out_proj_weight", "out_proj_bias",
        "linear1_weight", "linear1_bias", "linear2_weight", "linear2_bias",
        "norm1_weight", "norm1_bias", "norm2_weight", "norm2_bias")]


class ImuEncoderParams(C.Structure):
    _fields_ = [("seq", C.c_int32), ("layers", C.c_int32),
                ("cls_token", C.c_void_p), ("pos_encoding", C.c_void_p),
                ("patch_weight", C.c_void_p), ("patch_bias", C.c_void_p),
                ("norm_weight", C.c_void_p), ("norm_bias", C.c_void_p),
                ("layer", EncoderLayerParams * MAX_LAYERS)]


class HeadParams(C.Structure):
    _fields_ = [("hidden1", C.c_int32), ("hidden2", C.c_int32), ("classes", C.c_int32)] + \
               [(n, C.c_void_p) for n in (
                   "w0", "b0", "bn0_weight", "bn0_bias", "bn0_mean", "bn0_var",
                   "w1", "b1", "bn1_weight", "bn1_bias", "bn1_mean", "bn1_var",
                   "w2", "b2")]


class ImuOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("cls", "cls_img", "tokens", "logits", "pred", "msp", "energy", "maha")]


class ConvLayerParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("weight", "bias", "bn_weight", "bn_bias", "bn_mean", "bn_var")]


class ConvEncoderParams(C.Structure):
    _fields_ = [("layer", ConvLayerParams * 3)]


_SIGNATURES = {
    "cmhar_abi_version": (C.c_int, []),
    "cmhar_last_error": (C.c_char_p, []),
    "cmhar_launch_count": (C.c_int64, []),
    "cmhar_imu_encoder_blob_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "cmhar_imu_encoder_pack": (C.c_int, [C.POINTER(ImuEncoderParams), C.c_void_p, C.c_void_p]),
    "cmhar_head_blob_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "cmhar_head_pack": (C.c_int, [C.POINTER(HeadParams), C.c_void_p, C.c_void_p]),
    "cmhar_maha_blob_bytes": (C.c_size_t, [C.c_int32]),
    "cmhar_maha_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "cmhar_imu_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int32, C.c_void_p]),
    "cmhar_imu_forward_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                       C.POINTER(ImuOutputs), C.c_int32, C.c_void_p]),
    "cmhar_head_kernel_kind": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "cmhar_fused_head_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cmhar_mlp2_forward_img": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "cmhar_similarity_img_work_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "cmhar_similarity_img": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.c_int32, C.c_int64, C.c_int64, C.c_int32,
                                       C.c_float, C.c_float, C.c_double, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_void_p]),
    "cmhar_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "cmhar_peer_free": (C.c_int, [C.c_void_p]),
    "cmhar_peer_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cmhar_peer_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "cmhar_peer_close": (C.c_int, [C.c_void_p]),
    "cmhar_peer_barrier": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_double,
                                     C.c_void_p, C.c_void_p]),
    "cmhar_video_pool_img": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "cmhar_video_pool_frames_img": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cmhar_operand_image_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "cmhar_linear_forward_img": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "cmhar_debug_cta_trace": (C.c_int, [C.c_void_p, C.c_int64]),
    "cmhar_debug_set_option": (C.c_int, [C.c_char_p, C.c_int32]),
    "cmhar_blob_release": (C.c_int, [C.c_void_p]),
    "cmhar_debug_imu_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "cmhar_head_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "cmhar_logit_scores": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "cmhar_linear_blob_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "cmhar_linear_pack": (C.c_int, [C.c_void_p] * 6 + [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "cmhar_linear_work_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "cmhar_linear_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p]),
    "cmhar_concat_linear_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32,
                                              C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p]),
    "cmhar_l2_normalize": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "cmhar_xattn_blob_bytes": (C.c_size_t, [C.c_int32]),
    "cmhar_xattn_pack": (C.c_int, [C.c_void_p] * 8 + [C.c_int32, C.c_void_p, C.c_void_p]),
    "cmhar_xattn_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                      C.c_void_p, C.c_void_p]),
    "cmhar_cross_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "cmhar_residual_ln_pool": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float,
                                         C.c_void_p, C.c_void_p]),
    "cmhar_conv_encoder_blob_bytes": (C.c_size_t, []),
    "cmhar_conv_encoder_pack": (C.c_int, [C.POINTER(ConvEncoderParams), C.c_void_p, C.c_void_p]),
    "cmhar_conv_encoder_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "cmhar_conv_encoder_forward_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "cmhar_frames_normalize": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int32, C.c_void_p, C.c_void_p]),
    "cmhar_video_pool_nhwc": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cmhar_video_pool": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_void_p, C.c_void_p]),
    "cmhar_video_pool_coresident": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_void_p, C.c_void_p]),
    "cmhar_similarity_work_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int32]),
    "cmhar_similarity": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int64,
                                   C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_float, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "cmhar_maha_fit64_doubles": (C.c_size_t, [C.c_int32]),
    "cmhar_maha_finalize": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cmhar_maha_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_int32, C.c_void_p]),
    "cmhar_maha_score": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]),
    "cmhar_near_tie_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cmhar_score_key_range": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "cmhar_score_histogram": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint32, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p]),
}

EXPORTED = tuple(sorted(_SIGNATURES))

_lib: Optional[C.CDLL] = None
_lock = threading.Lock()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile ``csrc/cmhar_b200.cu`` (unity build) for sm_100a into ``lib/libcmhar_b200.so``."""
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = sources() + [HEADER]
    if not force and os.path.exists(LIB_PATH) and \
            os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + ["-o", LIB_PATH, os.path.join(CSRC, "cmhar_b200.cu")]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


def lib() -> C.CDLL:
    """Load the shared object (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: the CUDA extension was not built. Run "
                        "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                        "There is no CPU or PyTorch fallback for the inference hot path.")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype, fn.argtypes = res, args
                _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"cmhar_b200 error {rc}: {lib().cmhar_last_error().decode()}")


def launch_count() -> int:
    return int(lib().cmhar_launch_count())


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what}: the eval-mode inference path of cmhar_b200 runs only on CUDA (sm_100a) tensors; "
            f"got a tensor on '{t.device}'. There is no CPU fallback.")


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 contiguous view/copy (no-op for the usual case)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


UNSUPPORTED = -3


def ptr_array(ptrs):
    """HOST array of device pointers (``void* const*`` arguments)."""
    return (C.c_void_p * len(ptrs))(*[int(p) for p in ptrs])


def alloc_blob(nbytes: int, device) -> torch.Tensor:
    """Caller-owned, 1 KiB-aligned device byte buffer (torch's allocator aligns to >= 512 B; we
    over-allocate and slice to guarantee 1024)."""
    if nbytes <= 0:
        raise RuntimeError("cmhar_b200: unsupported dimensions for this build (blob size 0)")
    raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    off = (-raw.data_ptr()) % 1024
    blob = raw[off:off + nbytes]
    # the library's host-side registry is keyed by the blob pointer: forget the entry when the buffer dies, so that memory the
    # allocator hands out again at the same address is never taken for a packed blob
    weakref.finalize(blob, _release_blob, blob.data_ptr()).atexit = False
    return blob


def enable_dev_env(on: bool = True) -> None:
    """Development tools only: let the library read its CMHAR_* A/B switches from the environment (ignored by default)."""
    check(lib().cmhar_debug_set_option(b"dev_env", int(on)))


def _release_blob(ptr: int) -> None:
    try:
        if _lib is not None:
            _lib.cmhar_blob_release(ptr)
    except Exception:
        pass
