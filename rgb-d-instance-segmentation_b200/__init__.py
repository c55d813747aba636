"""rgbd_b200 -- B200-native (sm_100a) DGGM / E-DSAM depth-guidance hot path.

Drop-in ``nn.Module`` mirrors of the reference's ``DepthGradientInjectionResidual``,
``DSAModule`` and ``EnhancedDepthImageRatioPredictor`` (mask2former/utils/custom_model.py)
whose forwards dispatch through the C-ABI library ``csrc/librgbd_b200.so`` (declared in
``include/rgbd_b200.h``) into hand-written CUDA kernels.  There is no CPU fallback: ops raise
if the library is missing or the tensors are not on an sm_100 device.
"""
__version__ = "0.1.0"

from . import synthetic  # noqa: F401  (numpy only)
