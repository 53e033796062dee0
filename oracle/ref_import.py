"""Import the UNMODIFIED reference (worldbench/lidarcrafter) leaf modules from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used in this container to (a) validate the oracle restatement and
(b) generate the golden fixtures under tests/golden/.  /root/reference does not exist on the GPU
box, so nothing that runs there may import this module (it raises if the tree is absent).

Recipe: the package __init__ files of the reference pull in un-installed
dependencies (timm, clip, generated version.py, pydantic mutable defaults), so we pre-seed
sys.modules with empty stub packages whose __path__ points at the real directories and import only
leaf modules.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("LIDARCRAFTER_REF", "/root/reference")

_STUBS = [
    ("lidargen", "lidargen"),
    ("lidargen.models", "lidargen/models"),
    ("lidargen.models.unets", "lidargen/models/unets"),
    ("lidargen.utils", "lidargen/utils"),
    ("lidargen.dataset", "lidargen/dataset"),
    ("lidargen.dataset.transforms_3d", "lidargen/dataset/transforms_3d"),
]


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "lidargen"))


def setup() -> None:
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for name, rel in _STUBS:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF_ROOT, rel)]
            sys.modules[name] = m


def efficient_unet():
    setup()
    from lidargen.models.unets import efficient_unet as m
    return m


def layout_unet_v1():
    setup()
    from lidargen.models.unets import layout_unet_v1 as m
    return m


def layout_encoder():
    setup()
    from lidargen.models.unets import layout_encoder as m
    return m


def unet_ops():
    setup()
    from lidargen.models.unets import ops as m
    return m


def continuous_time():
    setup()
    from lidargen.models.diffusion import continuous_time as m
    return m


def continuous_time_cond():
    setup()
    from lidargen.models.diffusion import continuous_time_cond as m
    return m


def lidar():
    setup()
    from lidargen.utils import lidar as m
    return m


def transforms_common():
    setup()
    from lidargen.dataset.transforms_3d import common as m
    return m


def metric_utils():
    """lidargen/metrics/metric_utils.py (the package __init__ imports cleanly: torchsparse / chamfer / emd are optional
    there and only print a hint when missing)."""
    setup()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        from lidargen.metrics import metric_utils as m
    return m
