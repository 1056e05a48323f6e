"""ctypes binding of libspdm.so — the C ABI declared in include/spdm.h.

This is the only door between the Python mirror of the reference's module surface and the CUDA
kernels.  It fails loudly when the library or a CUDA device is missing: there is no CPU fallback.
"""
import ctypes
import os

from . import _build

_c = ctypes
_lib = None


class SpdmConfig(_c.Structure):
    _fields_ = [(n, _c.c_int32) for n in (
        "variant", "precision", "batch_max", "rows", "dim", "obs_horizon", "cond_dim", "inpaint_rows",
        "time_dim", "device", "graph_steps", "flags")]


VARIANT_ATTENTION, VARIANT_NO_ATTENTION, VARIANT_SIMPLE_UNET = 0, 1, 2
PRECISION_FP32, PRECISION_BF16, PRECISION_TF32 = 0, 1, 2
SCHED_DDPM, SCHED_DDIM = 0, 1
FLAG_SCHEDULER_ONLY = 1
FLAG_ENCODER_RESNET18 = 2
PROFILE_CLASSES = ("conv3x3", "gemm1x1", "gn_apply", "gn_stats", "resample", "layernorm", "sdpa", "io_conv", "step", "conv3x3_gn")

_P = _c.c_void_p


IMAGES_U8_HWC, IMAGES_F32_CHW = 0, 1


class SpdmDataStats(_c.Structure):  # include/spdm.h: spdm_data_stats
    _fields_ = [("pos_min", _c.c_float), ("pos_max", _c.c_float), ("vel_min", _c.c_float * 2), ("vel_max", _c.c_float * 2),
                ("act_min", _c.c_float * 3), ("act_max", _c.c_float * 3)]


_PROTOTYPES = {
    "spdm_plan_create": (_c.c_int, [_c.POINTER(_P), _c.POINTER(SpdmConfig)]),
    "spdm_plan_destroy": (_c.c_int, [_P]),
    "spdm_plan_load_weight": (_c.c_int, [_P, _c.c_char_p, _P, _c.POINTER(_c.c_int64), _c.c_int32, _P]),
    "spdm_plan_missing_weights": (_c.c_int, [_P]),
    "spdm_plan_set_schedule": (_c.c_int, [_P, _c.c_int32, _c.c_int32, _P, _P, _P]),
    "spdm_encode_images": (_c.c_int, [_P, _P, _P, _c.c_int32, _P]),
    "spdm_encode_cond": (_c.c_int, [_P, _P, _P, _P, _P, _c.c_int32, _P]),
    "spdm_encode_cond_u8": (_c.c_int, [_P, _P, _P, _P, _P, _c.c_int32, _P]),
    "spdm_set_cond": (_c.c_int, [_P, _P, _c.c_int32, _P]),
    "spdm_get_cond": (_c.c_int, [_P, _P, _c.c_int32, _P]),
    "spdm_unet_forward": (_c.c_int, [_P, _P, _P, _c.c_int32, _P, _c.c_int32, _P, _c.c_int32, _P]),
    "spdm_debug_forward": (_c.c_int64, [_P, _P, _P, _c.c_int32, _P, _c.c_int32, _P, _c.c_int32, _c.c_char_p, _P, _P]),
    "spdm_step": (_c.c_int, [_P, _P, _P, _P, _P, _P, _c.c_int32, _c.c_int32, _P]),
    "spdm_sample": (_c.c_int, [_P, _P, _P, _P, _P, _P, _c.c_uint64, _c.c_int32, _P]),
    "spdm_add_noise": (_c.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _c.c_int32, _P]),
    "spdm_profile_step": (_c.c_int, [_P, _c.c_int32, _c.c_int32, _c.POINTER(_c.c_double), _P]),
    "spdm_microbench_conv": (_c.c_int, [_c.c_int32] * 8 + [_c.POINTER(_c.c_float)]),
    "spdm_train_enable": (_c.c_int, [_P]),
    "spdm_train_bind": (_c.c_int, [_P, _c.c_char_p, _c.c_int64, _c.POINTER(_c.c_int64), _c.c_int32]),
    "spdm_train_set_buffers": (_c.c_int, [_P, _P, _P, _c.c_int64]),
    "spdm_train_sync_weights": (_c.c_int, [_P, _P]),
    "spdm_train_fwd_bwd": (_c.c_int, [_P] + [_P] * 11 + [_c.c_int32, _P]),
    "spdm_train_set_image_stride": (_c.c_int, [_P, _c.c_int64]),
    "spdm_train_set_valid": (_c.c_int, [_P, _c.c_int32]),
    "spdm_train_wait_phase": (_c.c_int, [_P, _c.c_int32, _P]),
    "spdm_adam_step": (_c.c_int, [_P, _P, _P, _P, _c.c_int64, _c.c_float, _c.c_float, _c.c_float, _c.c_float, _c.c_int32,
                                  _c.c_float, _c.c_float, _P, _P]),
    "spdm_gather_windows": (_c.c_int, [_P, _c.c_int32, _c.c_int32, _c.c_int32, _P, _P, _P, _P, _c.c_int32, _c.c_int32, _c.c_int32,
                                       _c.c_int32, _P, _P, _P, _P, _P, _P, _P]),
    "spdm_unnormalize_position": (_c.c_int, [_P, _P, _c.c_float, _c.c_float, _c.c_int64, _c.c_int32, _P, _P]),
    "spdm_data_launch_count": (_c.c_int64, []),
    "spdm_plan_launch_count": (_c.c_int64, [_P]),
    "spdm_plan_workspace_bytes": (_c.c_int64, [_P]),
    "spdm_plan_batch_multiple": (_c.c_int32, [_P]),
    "spdm_last_error": (_c.c_char_p, []),
    "spdm_version": (_c.c_char_p, []),
}
EXPORTED_SYMBOLS = tuple(_PROTOTYPES)


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """dlopen libspdm.so (building it first if the sources are newer) and set the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build_if_missing and _build.is_stale():
        try:
            _build.build()
        except Exception as e:  # a stale-but-present library is still usable (e.g. no nvcc on the box)
            if not os.path.exists(path):
                raise RuntimeError(
                    "libspdm.so is missing and could not be built (%s). The spdm CUDA path has no CPU "
                    "fallback; build it with `python -c 'import __graft_entry__ as g; g.build()'`." % e)
    if not os.path.exists(path):
        raise RuntimeError("libspdm.so not found at %s — the spdm CUDA path has no CPU fallback" % path)
    lib = _c.CDLL(path)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class SpdmError(RuntimeError):
    pass


def check(rc):
    if rc is not None and rc < 0:
        raise SpdmError(load().spdm_last_error().decode("utf-8", "replace"))
    return rc
