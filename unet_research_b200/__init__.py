"""unet_research_b200 -- B200-native (sm_100a) implementation of the U-Net / MC-DropBlock hot path of
JohnDLee/Unet-Research.  Public surface mirrors the reference modules:

    from unet_research_b200 import UNet, DropBlock2D, Dropblock2d_ichan, LinearScheduler   # utils_unet / utils_modules
    from unet_research_b200 import DropBlockEval, RotationEval, set_dropblock_on   # uncertainty scripts
    from unet_research_b200 import BaseUNetTraining                             # utils_training
    from unet_research_b200 import FusedSGD                                     # training.py:32 SGD + Lightning's clip

All compute goes through the C-ABI library csrc/libb2u.so (include/b2u.h); there is no CPU fallback.
"""
from .modules import DropBlock2D, Dropblock2d_ichan, LinearScheduler
from .unet import UNet
from .uncertainty import DropBlockEval, RotationEval, set_dropblock_on, shard_range
from .training import BaseUNetTraining
from .optim import FusedSGD
from .metrics import get_accuracy_metrics
from .resize import square_pad_resize

__all__ = ["UNet", "DropBlock2D", "Dropblock2d_ichan", "LinearScheduler", "DropBlockEval", "RotationEval", "set_dropblock_on",
           "shard_range", "BaseUNetTraining", "FusedSGD", "get_accuracy_metrics", "square_pad_resize"]
