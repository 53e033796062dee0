"""ctypes binding of libb200lidar.so (C-ABI declared in include/b200lidar.h).

The product path has NO fallback: if the shared library is missing, or the device is not sm_100,
loading raises.  PyTorch is only used by callers for device memory / streams; every pointer passed
here is a raw device address (tensor.data_ptr()).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libb200lidar.so")

P = C.c_void_p
I = C.c_int
F = C.c_float
SZ = C.c_size_t
D = C.c_double

# name -> (restype, argtypes); must mirror include/b200lidar.h exactly (tests/test_abi.py checks it)
PROTOTYPES = {
    "b200_version": (I, []),
    "b200_device_check": (I, [I]),
    "b200_last_error": (C.c_char_p, []),
    "b200_conv_tc": (I, [P, P, P, P, F, F, P, P, I, I, I, I, I, I, I, I, I, I, P]),
    "b200_conv_tc_splitk": (I, [P, P, P, P, F, F, P, P, P, I, I, I, I, I, I, I, I, I, I, I, P]),
    "b200_conv_gn_tc": (I, [P, I, P, I, P, P, P, P, P, I, I, F, I, P, P, P, F, F, P, P, I, I, I, I, I, I, I, I, I, P]),
    "b200_conv_set_debug": (I, [P]),
    "b200_conv_set_ablate": (I, [I]),
    "b200_packed_weight_elems": (SZ, [I, I, I, I]),
    "b200_pack_conv_weight": (I, [P, P, I, I, I, I, I, I, F, P]),
    "b200_conv_merged": (I, [I, I, I]),
    "b200_pack_conv_weight_plain": (I, [P, P, I, I, I, I, F, P]),
    "b200_conv_ffma": (I, [P, P, P, P, F, F, P, P, I, I, I, I, I, I, I, I, P]),
    "b200_gn_act_f16": (I, [P, I, P, I, P, P, P, P, P, I, I, F, I, P, P, I, I, I, I, P]),
    "b200_gn_act_f32": (I, [P, P, P, P, I, F, I, P, I, I, I, P]),
    "b200_attn_set_debug": (I, [P]),
    "b200_flash_attention_workspace": (SZ, [I, I, I, I, I, I]),
    "b200_flash_attention": (I, [P, I, P, I, I, I, I, I, F, P, P]),
    "b200_flash_attention_oa": (I, [P, P, P, P, P, P, I, I, I, I, I, I, I, F, P, P]),
    "b200_channel_stats": (I, [P, P, I, I, I, P]),
    "b200_fir_resample": (I, [P, P, P, I, I, I, I, I, I, P]),
    "b200_fir_up_operand": (I, [P, P, I, I, I, I, I, I, P]),
    "b200_time_embed": (I, [P, P, P, P, P, P, P, P, P, P, I, I, I, I, P]),
    "b200_in_conv": (I, [P, P, P, I, P, P, I, I, I, I, I, I, P]),
    "b200_conv_direct_f32": (I, [P, P, P, P, I, I, I, I, I, I, I, P]),
    "b200_out_conv": (I, [P, I, P, P, P, I, I, I, I, I, I, P]),
    "b200_sampler_update": (I, [P, P, P, P, P, I, I, I, I, F, P]),
    "b200_sampler_coefficients": (I, [P, P, I, F, F, F, F, F, P, P, I, P]),
    "b200_range_project": (I, [P, P, P, P, P, I, I, I, I, F, F, F, F, P]),
    "b200_range_project_f64": (I, [P, P, P, P, P, I, I, I, I, F, F, F, F, P]),
    "b200_boxes_to_mask_workspace": (SZ, [I, I]),
    "b200_boxes_to_mask": (I, [P, I, I, I, I, I, F, F, P, P, P, P, P]),
    "b200_points_in_boxes": (I, [P, P, P, I, I, P]),
    "b200_points_in_boxes_first": (I, [P, P, P, I, I, I, P]),
    "b200_voxel_index": (I, [P, P, P, I, I, I, I, I, P]),
    "b200_depth_to_xyz": (I, [P, P, P, P, I, I, I, F, F, P]),
    "b200_pcd2range": (I, [P, P, P, P, P, P, I, I, I, I, F, F, F, F, F, P]),
    "b200_range2xyz": (I, [P, P, I, I, I, F, F, F, F, F, I, P]),
    "b200_quantize_coords": (I, [P, I, I, I, I, D, D, D, I, P, P, P]),
    "b200_sparse_quantize_workspace": (SZ, [P, I, I]),
    "b200_ravel_hash": (I, [P, I, I, P, P, P]),
    "b200_sparse_quantize": (I, [P, I, I, P, P, SZ, P, P, P, P, P]),
    "b200_bev_occupancy_sum": (I, [P, P, I, I, I, F, F, F, F, F, I, I, I, I, P, P, P]),
    "b200_voxel_occupancy": (I, [P, I, I, P, F, P, P, P, P]),
}

_NO_STATUS = {"b200_conv_merged", "b200_version", "b200_device_check", "b200_last_error", "b200_packed_weight_elems",
              "b200_sparse_quantize_workspace", "b200_boxes_to_mask_workspace", "b200_flash_attention_workspace"}


class B200LidarError(RuntimeError):
    pass


class Lib:
    """Thin checked wrapper: ``lib.conv_tc(...)`` calls ``b200_conv_tc`` and raises on a non-zero status."""

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise B200LidarError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C lidarcrafter_b200/csrc`).  There is no CPU / PyTorch fallback.")
        self.path = path
        self.cdll = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(self.cdll, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        self.n_launches = 0

    def __getattr__(self, short: str):
        name = "b200_" + short
        if name not in PROTOTYPES:
            raise AttributeError(short)
        fn = getattr(self.cdll, name)
        if name in _NO_STATUS:
            return fn

        def call(*args):
            rc = fn(*args)
            if rc != 0:
                msg = self.cdll.b200_last_error()
                raise B200LidarError(f"{name} failed ({rc}): {msg.decode() if msg else ''}")
            self.n_launches += 1
            return rc

        self.__dict__[short] = call
        return call


_LIB = None
_TEST_LIB = None


def get_lib():
    global _LIB
    if _TEST_LIB is not None:
        return _TEST_LIB
    if _LIB is None:
        _LIB = Lib()
    return _LIB


def set_test_lib(lib) -> None:
    """tests/ only: inject an emulator of the C-ABI (tests/abi_emulator.py) so the HOST logic (plans,
    weight packing order, buffer wiring) can be checked on a machine without a GPU.  Never set by the
    product path; with it unset a missing .so / non-sm_100 device raises."""
    global _TEST_LIB
    _TEST_LIB = lib


def compute_device() -> str:
    """"cuda" on the product path; "cpu" only while tests/ have injected the C-ABI emulator (set_test_lib)."""
    return "cuda" if _TEST_LIB is None else "cpu"


def current_stream(device) -> int:
    import torch
    if _TEST_LIB is not None and device.type != "cuda":
        return 0
    return torch.cuda.current_stream(device).cuda_stream


def require_b200(device_index: int = 0) -> int:
    """Returns the SM count; raises unless a compute-capability-10.x device is present."""
    if _TEST_LIB is not None:
        return 148
    lib = get_lib()
    n = lib.cdll.b200_device_check(device_index)
    if n <= 0:
        msg = lib.cdll.b200_last_error()
        raise B200LidarError(f"no usable B200 device: {msg.decode() if msg else n}")
    return n
