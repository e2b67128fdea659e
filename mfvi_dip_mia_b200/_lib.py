"""ctypes binding of libmfvidip.so (include/mfvi_dip.h) — the only way the Python host reaches the GPU kernels.

There is no CPU or PyTorch fallback: if the shared library is missing the import fails loudly, and every
call with a non-zero return code raises `MfviError` carrying `mfvi_last_error()`.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmfvidip.so")

ABI_VERSION = 1
MATH_FP32 = 0
MATH_TF32 = 1
STREAM_WEIGHTS = 0
STREAM_INPUT_JITTER = 1


class MfviError(RuntimeError):
    pass


class PhiloxKey(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("step", C.c_uint32), ("sample0", C.c_uint32), ("step_dev", C.c_void_p)]


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sstride", C.c_longlong), ("hstride", C.c_int), ("wstride", C.c_int)]


class ConvDesc(C.Structure):
    _fields_ = [("S", C.c_int), ("Cin", C.c_int), ("Cout", C.c_int), ("KH", C.c_int), ("KW", C.c_int),
                ("stride", C.c_int), ("Hin", C.c_int), ("Win", C.c_int), ("Hout", C.c_int), ("Wout", C.c_int),
                ("math", C.c_int)]


class PlanInfo(C.Structure):
    """MfviPlanInfo: answer of mfvi_conv2d_plan (host-only query, works without a GPU)."""
    _fields_ = [("family", C.c_char * 24), ("grid", C.c_uint * 3), ("block", C.c_uint), ("smem_bytes", C.c_ulonglong),
                ("launches", C.c_int), ("detail", C.c_char * 200)]


_P = C.c_void_p
_LL = C.c_longlong
_I = C.c_int
_F = C.c_float
_D = C.c_double
_SZ = C.c_size_t
_U32 = C.c_uint32
_CD = C.POINTER(ConvDesc)

# name -> argtypes (the trailing stream argument is appended automatically by `call`)
_SIGS = {
    "mfvi_philox_raw_fill": [_P, _SZ, PhiloxKey, _U32],
    "mfvi_philox_normal_fill": [_P, _SZ, PhiloxKey, _U32],
    "mfvi_sample_weights": [_P, _P, _SZ, _I, _P, _LL, PhiloxKey, _P, _LL],
    "mfvi_conv2d_fwd": [_CD, View, _P, _P, _LL, View, _P],
    "mfvi_conv2d_dgrad": [_CD, View, _P, _LL, View, _I],
    "mfvi_conv2d_wgrad": [_CD, View, View, _P, _P, _LL],
    "mfvi_kl_reparam_fwd_bwd": [_P, _P, _SZ, _F, _D, _I, _F, _P, _P, _LL, _I, _P, _LL, PhiloxKey, _F, _P, _P, _P, _I],
    "mfvi_bn_act_pad_fwd": [View, _I, _I, _I, _I, _P, _P, _P, _I, _I, View],
    "mfvi_cat_up_fwd": [View, _I, _P, _P, _P, View, _I, _P, _P, _P, _I, _I, _I, _I, View, _P],
    "mfvi_pad_act_bwd": [View, _I, _I, _I, _I, _I, View, _P, _P, _P, _I, View, _P],
    "mfvi_bn_bwd_apply": [View, View, _I, _I, _I, _I, _P, _P, _P, View, _P, _P],
    "mfvi_cat_up_bwd": [View, _I, _I, _I, _I, View, _I, _P, _P, _P, View, _P, View, _I, _P, _P, _P, View, _P, _I],
    "mfvi_bn_running_update": [_P, _P, _P, _P, _P, _I, _I, _F, _P, _P],
    "mfvi_gauss_nll_fwd_bwd": [_I, View, _I, _I, _I, _I, _I, _P, _P, _P, View],
    "mfvi_mse_fwd_bwd": [_P, _LL, _P, _SZ, _I, _P, _P],
    "mfvi_radon_fwd": [View, _I, _I, _I, _I, _P, _I, _P],
    "mfvi_radon_bwd": [_P, _I, _I, _I, _I, _P, _I, View],
    "mfvi_input_jitter_pad": [_P, _P, _I, _I, _I, _F, _I, PhiloxKey, View],
    "mfvi_adamw_step": [_P, _P, _P, _P, _SZ, _F, _F, _F, _F, _F, _I, _P, _P],
    "mfvi_mega_run": [_P, _I, _I, _P],
    "mfvi_conv2d_fwd_simt": [_CD, View, _P, _P, _LL, View, _P],
    "mfvi_conv2d_dgrad_simt": [_CD, View, _P, _LL, View, _I],
    "mfvi_conv2d_wgrad_simt": [_CD, View, View, _P, _P, _LL],
    "mfvi_conv2d_fwd_mma": [_CD, View, _P, _P, _LL, View, _P],
    "mfvi_conv2d_dgrad_mma": [_CD, View, _P, _LL, View, _I],
    "mfvi_conv2d_wgrad_mma": [_CD, View, View, _P, _P, _LL],
    "mfvi_bn_param_grads": [_P, _P, _P, _P, _I, _I],
    "mfvi_counter_add": [_P, _U32],
    "mfvi_counter_add_if_finite": [_P, _U32, _P],
    "mfvi_loss_flag": [_P, _P, _F, _P],
    "mfvi_bookkeep_step": [View, _I, _I, _I, _F, _P, _P, _P, _P, _P, _I, _P, _I, _P],
    "mfvi_bookkeep_step_ex": [View, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P, _I, _P, _I, _P],
    "mfvi_ssim": [_P, _P, _I, _I, _I, _P],
    "mfvi_ring_uncertainty": [_P, _P, _I, _I, _I, _P, _P, _P, _P],
    "mfvi_softplus_sq_fwd": [_P, _SZ, _P],
    "mfvi_softplus_sq_bwd": [_P, _P, _SZ, _P, _I],
    "mfvi_square_fwd": [_P, _SZ, _P],
    "mfvi_square_bwd": [_P, _P, _SZ, _P, _I],
    "mfvi_lrt_noise_fwd": [_P, _P, _P, _SZ, _P],
    "mfvi_lrt_noise_bwd": [_P, _P, _P, _SZ, _P],
    "mfvi_fill_f32": [_P, _SZ, _F],
    "mfvi_nchw_to_nhwc": [_P, _P, _I, _I, _I, _I],
    "mfvi_nhwc_to_nchw": [_P, _P, _I, _I, _I, _I],
}

EXPORTS = ["mfvi_abi_version", "mfvi_last_error", "mfvi_conv2d_plan", "mfvi_mega_begin", "mfvi_mega_mark_nosync",
           "mfvi_mega_stage_bytes", "mfvi_mega_end"] + list(_SIGS)


def _load():
    if not os.path.exists(LIB_PATH):
        raise MfviError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
            f"`bash mfvi_dip_mia_b200/csrc/build.sh`. There is no CPU/PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    lib.mfvi_abi_version.restype = C.c_int
    lib.mfvi_last_error.restype = C.c_char_p
    got = lib.mfvi_abi_version()
    if got != ABI_VERSION:
        raise MfviError(f"libmfvidip ABI version {got}, host expects {ABI_VERSION}: rebuild the library")
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = list(args) + [_P]
        fn.restype = C.c_int
    lib.mfvi_conv2d_plan.argtypes = [_CD, _I, View, View, _LL, _I, _I, C.POINTER(PlanInfo)]   # no stream: nothing is launched
    lib.mfvi_conv2d_plan.restype = C.c_int
    # program recording of the persistent multi-stage kernel (host-only calls, no stream)
    lib.mfvi_mega_begin.argtypes, lib.mfvi_mega_begin.restype = [], C.c_int
    lib.mfvi_mega_mark_nosync.argtypes, lib.mfvi_mega_mark_nosync.restype = [], C.c_int
    lib.mfvi_mega_stage_bytes.argtypes, lib.mfvi_mega_stage_bytes.restype = [], C.c_size_t
    lib.mfvi_mega_end.argtypes = [_P, C.c_size_t, C.POINTER(C.c_int)]
    lib.mfvi_mega_end.restype = C.c_int
    return lib


lib = _load()
launch_count = 0   # number of kernel-launching C-ABI calls made by this process (bench.py reports it)


timeline = None     # when a list: every call appends (name, start_event, end_event, meta) — bench.py's per-kernel timing


def call(name: str, *args, stream=None, meta=None):
    """Invoke `name(*args, stream)`; raises MfviError on a non-zero return code."""
    global launch_count
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    if timeline is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args, stream)
        e1.record()
        timeline.append((name, e0, e1, meta))
    else:
        rc = getattr(lib, name)(*args, stream)
    launch_count += 1
    if rc != 0:
        raise MfviError(f"{name} failed (rc={rc}): {lib.mfvi_last_error().decode()}")


PASS_FWD, PASS_DGRAD, PASS_WGRAD = 0, 1, 2


def conv_plan(desc: ConvDesc, pass_: int, a: View, b: View, w_sstride: int, accumulate: int = 0, with_bias: bool = True) -> dict:
    """Which kernel family mfvi_conv2d_{fwd,dgrad,wgrad} would run for this geometry and these views, with its launch
    geometry and tile plan — a host-only query (no device is touched; usable on a machine without a GPU)."""
    info = PlanInfo()
    rc = lib.mfvi_conv2d_plan(C.byref(desc), pass_, a, b, w_sstride, accumulate, 1 if with_bias else 0, C.byref(info))
    if rc != 0:
        raise MfviError(f"mfvi_conv2d_plan failed (rc={rc}): {lib.mfvi_last_error().decode()}")
    plan = {}
    for kv in info.detail.decode().split():
        k, _, v = kv.partition("=")
        plan[k] = int(v) if v.lstrip("-").isdigit() else v
    return {"family": info.family.decode(), "grid": tuple(info.grid), "block": info.block, "smem_bytes": info.smem_bytes,
            "launches": info.launches, "plan": plan}


def record_program(ops, device) -> dict:
    """Record an op list [(name, args, meta), ...] as ONE program of the persistent multi-stage kernel (include/mfvi_dip.h,
    mfvi_mega_*): every op is passed to its usual entry point, which appends a stage instead of launching.  An op whose meta
    has "indep" runs without a barrier after the op before it.  Returns the program (device buffer, stage count, op names)."""
    global launch_count
    n = len(ops)
    prog = torch.empty(max(n, 1) * int(lib.mfvi_mega_stage_bytes()), dtype=torch.uint8, device=device)
    if lib.mfvi_mega_begin() != 0:
        raise MfviError(f"mfvi_mega_begin failed: {lib.mfvi_last_error().decode()}")
    n_st = C.c_int(0)
    try:
        before = launch_count
        for name, args, meta in ops:
            if meta.get("indep"):
                lib.mfvi_mega_mark_nosync()
            call(name, *args, stream=0)
        launch_count = before                                  # nothing was launched
    finally:
        rc = lib.mfvi_mega_end(prog.data_ptr(), prog.numel(), C.byref(n_st))
    if rc != 0:
        raise MfviError(f"mfvi_mega_end failed (rc={rc}): {lib.mfvi_last_error().decode()}")
    prof = torch.zeros(n_st.value + 1, dtype=torch.int64, device=device) if os.environ.get("MFVI_MEGA_PROFILE") else None
    return {"prog": prog, "n_stages": n_st.value, "prof": prof,
            "names": [op[0] + (":" + op[2].get("layer", "") if op[2].get("layer") else "") for op in ops]}


def on_device(fn):
    """Method decorator: run with the object's CUDA device current.  The library launches on the CURRENT device and `call`
    takes the current stream, so an engine living on cuda:k (trial fan-out: one runner per device; bayesian_optimization.py:3762)
    must make that device current around every call — the reference gets this for free from ATen's device guards."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        dev = getattr(self, "device", None)
        if dev is None or torch.device(dev).type != "cuda":
            return fn(self, *a, **k)
        with torch.cuda.device(dev):
            return fn(self, *a, **k)
    return wrapper


def device_guarded(cls):
    """Class decorator for torch.autograd.Function: forward / backward run with the device of their first CUDA tensor current."""
    import functools

    def guard(f):
        @functools.wraps(f)
        def wrapper(ctx, *a, **k):
            t = next((x for x in a if torch.is_tensor(x) and x.is_cuda), None)
            if t is None:
                return f(ctx, *a, **k)
            with torch.cuda.device(t.device):
                return f(ctx, *a, **k)
        return wrapper
    cls.forward = staticmethod(guard(cls.forward))
    cls.backward = staticmethod(guard(cls.backward))
    return cls


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise MfviError(f"{what}: tensor is on {t.device}; this path runs only on a CUDA device (sm_100a) — "
                        "there is no CPU fallback")


def ptr(t):
    """Device pointer of a tensor (or NULL for None)."""
    return None if t is None else t.data_ptr()


def view(t: torch.Tensor, broadcast: bool = False) -> View:
    """MfviView of an NHWC fp32 tensor (S,H,W,C) (channel stride must be 1; other strides arbitrary, in elements)."""
    assert t.dim() == 4 and t.dtype == torch.float32 and (t.stride(3) == 1 or t.shape[3] == 1), (t.shape, t.stride())
    ss = 0 if (broadcast or t.shape[0] == 1) else t.stride(0)
    return View(t.data_ptr(), ss, t.stride(1), t.stride(2))


def key(seed: int, step: int = 0, sample0: int = 0, step_dev: torch.Tensor | None = None) -> PhiloxKey:
    return PhiloxKey(seed & 0xFFFFFFFFFFFFFFFF, step & 0xFFFFFFFF, sample0, ptr(step_dev))
