"""Test-infrastructure alias: the deterministic weight factory lives in the product package
(``rgbd_b200.synthetic_weights``) so that ``bench.py``'s own arm never imports ``oracle``; the oracle and the golden
generators keep using it under this name."""
import rgbd_b200  # noqa: F401  (root shim)
from rgbd_b200.synthetic_weights import *  # noqa: F401,F403
from rgbd_b200.synthetic_weights import (bn_weights, conv_weights, dggm_weights, dsam_weights, guidance_weights,  # noqa: F401
                                         guidance_weights_feature_ratio, linear_weights, ratio_feat_weights, ratio_weights)
