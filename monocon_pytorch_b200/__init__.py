"""monocon_pytorch_b200 -- B200-native (sm_100a) MonoCon forward + decode behind the reference's nn.Module API.

(The directory is named with an underscore so that it is importable; it is the package the task text
calls ``monocon-pytorch_b200``.)
"""
from .detector import MonoConDetector, default_head_config, default_test_config   # noqa: F401
from .engine import Engine, EngineError, PRED_NAMES, PRED_CHANNELS, conv2d, inverse_viewpad, load_library  # noqa: F401

from .train_ops import TargetGenerator, get_losses, ClipAdamW, ResidentClipAdamW, LOSS_NAMES   # noqa: F401

__all__ = ['TargetGenerator', 'get_losses', 'ClipAdamW', 'ResidentClipAdamW', 'LOSS_NAMES', 'MonoConDetector', 'Engine', 'EngineError', 'PRED_NAMES', 'PRED_CHANNELS', 'conv2d', 'inverse_viewpad',
           'load_library', 'default_head_config', 'default_test_config']
