"""coivo_b200 -- B200-native photometric-loss hot path of ColVO (HNUicda/CoIVO).

Only the path BASELINE.json's north_star names lives here: the fused CUDA kernels and their
C ABI (`csrc/`, `include/colvo.h`) and the host-side operator that mirrors the interface a
PyTorch training loop calls (`photometric_loss`, `consistency`).  CUDA-only by design.
"""
from .loss import photometric_loss, HostStepper, pack_images, unpack_images  # noqa: F401
from .consistency import consistency  # noqa: F401
from .graph import GraphedStep  # noqa: F401
from .frontend import disp_to_depth, photometric_loss_raw, poses_from_parameters  # noqa: F401
from . import dist, synthetic  # noqa: F401

__all__ = ["photometric_loss", "consistency", "HostStepper", "GraphedStep", "pack_images", "unpack_images", "poses_from_parameters",
           "disp_to_depth",
           "photometric_loss_raw", "dist", "synthetic"]
