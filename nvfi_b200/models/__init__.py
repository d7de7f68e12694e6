"""Host-side mirror of the reference's ``models`` package (models/__init__.py:1-6).

Same public names, constructor signatures, attributes and ``state_dict`` keys as the
reference, so ``train_nvfi.py`` / ``test_transfer_vel.py`` can import this package in
place of theirs.  The hot path of every class runs in the CUDA library
(``nvfi_b200/csrc`` through ``nvfi_b200.engine``); these modules only hold parameters,
draw the host-side random numbers in the reference's order, and wire autograd.
"""
from .rays import Ray, Camera, BatchedRays
from .render import Renderer
from .wrapper import NVFi
from .field import TensorVMKeyframeTimeKplane, AlphaGridMask
from .velocity import N_to_reso, VelocityAABB, VelocityAABBSur, VelBasis, PositionEncoder
from .masks import MaskField

__all__ = ["Ray", "Camera", "BatchedRays", "Renderer", "NVFi", "AlphaGridMask", "N_to_reso",
           "VelocityAABB", "VelocityAABBSur", "VelBasis", "PositionEncoder", "MaskField",
           "TensorVMKeyframeTimeKplane"]
