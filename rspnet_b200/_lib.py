"""ctypes binding of ``librspnet_b200.so`` (C ABI: include/rspnet_b200.h).

There is deliberately no fallback: if the library is missing or the device is not sm_100 every entry point
raises.  PyTorch is only used for device memory and streams.
"""
import ctypes as C
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "librspnet_b200.so"

c_i32, c_i64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class ConvDesc(C.Structure):
    """Mirror of ``rsp_conv3d_desc``."""
    _fields_ = [(n, c_i32) for n in
                ("N", "Ti", "Hi", "Wi", "Ci", "Co", "kt", "kh", "kw", "st", "sh", "sw", "pt", "ph", "pw")]

    def out_dims(self):
        return ((self.Ti + 2 * self.pt - self.kt) // self.st + 1,
                (self.Hi + 2 * self.ph - self.kh) // self.sh + 1,
                (self.Wi + 2 * self.pw - self.kw) // self.sw + 1)


class PoolDesc(C.Structure):
    """Mirror of ``rsp_pool3d_desc``."""
    _fields_ = [(n, c_i32) for n in
                ("N", "Ti", "Hi", "Wi", "C", "kt", "kh", "kw", "st", "sh", "sw", "pt", "ph", "pw")]

    def out_dims(self):
        return ((self.Ti + 2 * self.pt - self.kt) // self.st + 1,
                (self.Hi + 2 * self.ph - self.kh) // self.sh + 1,
                (self.Wi + 2 * self.pw - self.kw) // self.sw + 1)


# name -> (restype, argtypes); kept in the order of include/rspnet_b200.h
_P = c_vp
SIGNATURES = {
    "rsp_abi_version": (c_i32, []),
    "rsp_init": (c_i32, []),
    "rsp_last_error": (C.c_char_p, []),
    "rsp_conv3d_kpad": (c_i32, [C.POINTER(ConvDesc), c_i32]),
    "rsp_conv3d_packed_elems": (c_i64, [C.POINTER(ConvDesc), c_i32]),
    "rsp_conv3d_pack_weight": (c_i32, [C.POINTER(ConvDesc), c_i32, c_i32, _P, _P, c_i32, _P]),
    "rsp_conv3d_pack_weights": (c_i32, [c_i32, C.POINTER(ConvDesc), C.POINTER(c_i32), C.POINTER(c_i32),
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _P]),
    "rsp_conv3d_workspace_bytes": (c_i64, [C.POINTER(ConvDesc), c_i32]),
    "rsp_conv3d_fprop": (c_i32, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P]),
    "rsp_conv3d_dgrad": (c_i32, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P]),
    "rsp_conv3d_wgrad": (c_i32, [C.POINTER(ConvDesc), c_i32, c_i32, _P, _P, _P, _P, c_i32, _P]),
    "rsp_bn_stats": (c_i32, [_P, c_i64, c_i32, _P, _P, _P]),
    "rsp_bn_finalize": (c_i32, [_P, _P, c_i32, c_i64, _P, _P, c_f32, c_f32, _P, _P, _P, _P, _P, _P, c_i32, c_i32, _P]),
    "rsp_bn_act_fwd": (c_i32, [_P, _P, _P, _P, c_i32, _P, c_i64, c_i32, _P]),
    "rsp_bn_finalize_act_fwd": (c_i32, [_P, _P, _P, _P, c_i64, _P, _P, c_f32, c_f32, _P, _P, _P, _P, c_i32, _P, c_i64,
                                        c_i32, c_i32, _P]),
    "rsp_bn_act_bwd_reduce": (c_i32, [_P, _P, _P, _P, _P, c_i32, _P, _P, c_i64, c_i32, _P]),
    "rsp_bn_act_bwd_apply": (c_i32, [_P, _P, _P, _P, _P, _P, _P, _P, c_i32, _P, _P, c_i64, c_i32, c_i32, _P]),
    "rsp_maxpool3d_fwd": (c_i32, [C.POINTER(PoolDesc), _P, _P, _P, _P]),
    "rsp_maxpool3d_bwd": (c_i32, [C.POINTER(PoolDesc), _P, _P, _P, _P]),
    "rsp_bn_relu_maxpool_supported": (c_i32, [C.POINTER(PoolDesc)]),
    "rsp_bn_relu_maxpool_fwd": (c_i32, [C.POINTER(PoolDesc), _P, _P, _P, _P, _P, _P, _P]),
    "rsp_bn_relu_maxpool_bwd_sums": (c_i32, [C.POINTER(PoolDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "rsp_bn_relu_maxpool_bwd_dx": (c_i32, [C.POINTER(PoolDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P, c_i32, _P, _P, _P,
                                           _P]),
    "rsp_head_fwd": (c_i32, [_P, c_i32, c_i32, c_i32, c_i32, c_i32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "rsp_head_bwd": (c_i32, [_P, _P, _P, _P, _P, c_i32, c_i32, c_i32, c_i32, c_i32, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "rsp_gate_fwd": (c_i32, [_P, c_i32, c_i32, c_i32, c_i32, _P, _P, _P, _P, _P, _P, _P]),
    "rsp_gate_bwd": (c_i32, [_P, _P, c_i32, c_i32, c_i32, c_i32, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "rsp_copy_channels": (c_i32, [_P, c_i32, c_i32, _P, c_i32, c_i32, c_i32, c_i64, _P]),
    "rsp_ncdhw_to_ndhwc_bf16": (c_i32, [_P, _P, c_i32, c_i32, c_i32, c_i64, _P]),
    "rsp_ndhwc_bf16_to_ncdhw": (c_i32, [_P, _P, c_i32, c_i32, c_i32, c_i64, _P]),
    "rsp_clip_sample": (c_i32, [_P, _P, _P, _P, C.POINTER(c_f32), C.POINTER(c_f32), c_i32, c_i32, c_i32, c_i32, c_i32,
                                c_i32, _P, _P]),
    "rsp_clip_sample_jitter": (c_i32, [_P, _P, _P, _P, _P, _P, C.POINTER(c_f32), C.POINTER(c_f32), c_i32, c_i32, c_i32,
                                       c_i32, c_i32, c_i32, _P, _P]),
    "rsp_ema_update": (c_i32, [_P, _P, c_i64, c_f32, c_f32, _P]),
    "rsp_sgd_step": (c_i32, [_P, _P, _P, c_i64, c_f32, c_f32, c_f32, c_f32, c_i32, _P]),
    "rsp_speed_gather": (c_i32, [_P, _P, _P, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, _P, _P, _P, _P]),
    "rsp_gather_rows": (c_i32, [_P, _P, _P, c_i64, c_i64, _P]),
    "rsp_peer_alloc": (c_i32, [c_i64, C.POINTER(C.c_void_p)]),
    "rsp_peer_free": (c_i32, [_P]),
    "rsp_peer_export": (c_i32, [_P, C.c_char_p]),
    "rsp_peer_open": (c_i32, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "rsp_peer_close": (c_i32, [_P]),
    "rsp_gather_rows_peer": (c_i32, [_P, _P, _P, c_i64, c_i32, c_i64, _P]),
    "rsp_invert_permutation": (c_i32, [_P, _P, c_i32, _P]),
    "rsp_queue_enqueue": (c_i32, [_P, _P, _P, c_i32, c_i32, c_i32, _P]),
    "rsp_moco_logits_workspace": (c_i64, [c_i32, c_i32]),
    "rsp_moco_logits_fwd": (c_i32, [_P] * 7 + [c_i32, c_i32, c_i32, c_f32] + [_P] * 9 + [_P]),
    "rsp_moco_logits_fwd_ranked": (c_i32, [_P] * 7 + [c_i32, c_i32, c_i32, c_f32] + [_P] * 10 + [_P]),
    "rsp_metrics_update": (c_i32, [_P, _P, _P, _P, c_i32, _P, _P]),
    "rsp_moco_logits_bwd": (c_i32, [_P] * 7 + [c_i32, c_i32, c_i32, c_f32] + [_P] * 14 + [_P]),
    "rsp_moco_loss_fwd": (c_i32, [_P] * 6 + [c_i32, c_f32, c_f32, c_f32, _P, _P]),
    "rsp_moco_loss_bwd": (c_i32, [_P, _P, c_i32, c_f32, c_f32, c_f32] + [_P] * 7 + [_P]),
    "rsp_ce0_fwd": (c_i32, [_P, c_i32, c_i32, _P, _P]),
    "rsp_ce0_bwd": (c_i32, [_P, _P, c_i32, c_i32, _P, _P, _P]),
}

_lib = None
_initialised = False


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """dlopen the library and bind every symbol the header declares (no device needed)."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f"rspnet_b200: {_LIB_PATH} is missing. Build it with `python -m rspnet_b200.build` "
                "(or __graft_entry__.build()). There is no fallback path.")
        lib = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if lib.rsp_abi_version() != 2:
            raise RuntimeError("rspnet_b200: ABI version mismatch between python binding and library")
        _lib = lib
    return _lib


def _ensure_device():
    global _initialised
    if not _initialised:
        if not torch.cuda.is_available():
            raise RuntimeError("rspnet_b200: no CUDA device; this package has no CPU path")
        rc = load().rsp_init()
        if rc != 0:
            raise RuntimeError(load().rsp_last_error().decode())
        _initialised = True


_device_index = None


def stream_ptr() -> int:
    """Raw cudaStream_t of torch's current stream (one process drives one GPU: the device index is read once).
    torch.cuda.current_stream() builds a Stream object through several Python layers (~15 us); this is ~0.3 us."""
    global _device_index
    if _device_index is None:
        _device_index = torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(_device_index)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda, "rspnet_b200 kernels take CUDA tensors only"
    return t.data_ptr()


_fn_cache = {}


def call(name: str, *args):
    """Invoke a C-ABI entry point on the current stream; raises RuntimeError on a non-zero return."""
    fn = _fn_cache.get(name)
    if fn is None:
        _ensure_device()
        fn = _fn_cache[name] = getattr(load(), name)
    rc = fn(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {load().rsp_last_error().decode()}")
