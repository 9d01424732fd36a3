"""multimodal_tta_b200 -- B200-native (sm_100a) test-time-adaptation step for the 3-D
segmentation UNet of zhm1205/Multimodal_TTA, behind the reference's registry surface.

Importing the package registers (reference plugin surface, /root/reference/src/registry.py):
    model                ``unet_b200``     (drop-in for ``unet``)
    model                ``unet_multimodal_deepfusion_b200`` / ``unet_multimodal_midfusion_b200``
                                           (drop-in for ``unet_multimodal_deepfusion`` / ``..._midfusion``)
    evaluation strategy  ``tta_seg_eval``  (drop-in beside ``seg_eval``)
    plugin               ``tent_b200``     (the TENT method object)
The CUDA library (libtta_b200.so) is loaded lazily on first use and there is no fallback.
"""
from .config import DictConfig, compose_yaml, create, get_config, require_config  # noqa: F401
from .registry import (EVALUATION_STRATEGIES, MODELS, PLUGINS, get_evaluation_strategy,  # noqa: F401
                       get_model, get_plugin, register_evaluation_strategy, register_model,
                       register_plugin)
from .unet_b200 import B200Model, UNetB200  # noqa: F401
from .multimodal_b200 import MultimodalUNetB200  # noqa: F401
from .tent import TentB200  # noqa: F401
from .sliding_window import SlidingWindowTTA  # noqa: F401
from .evaluation import TTASegmentationEvaluationStrategy  # noqa: F401
from .intensity import IntensityPolicy  # noqa: F401

__all__ = ["UNetB200", "MultimodalUNetB200", "B200Model", "TentB200", "SlidingWindowTTA", "TTASegmentationEvaluationStrategy", "IntensityPolicy",
           "get_model", "get_plugin", "get_evaluation_strategy"]
