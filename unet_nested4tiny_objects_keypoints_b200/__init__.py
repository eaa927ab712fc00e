"""B200-native UNet_Nested (UNet++) hot path: hand-written sm_100a kernels (libunpp.so, C ABI in
include/unpp.h) behind the reference's ``models.UNet_Nested`` interface.

``from unet_nested4tiny_objects_keypoints_b200 import models`` mirrors the reference's ``models``
package for the hot path (trainer/trainer.py:337 selects ``getattr(models, 'UNet_Nested')()``).
"""
from . import models  # noqa: F401
from .models import UNet_Nested  # noqa: F401

__all__ = ["models", "UNet_Nested"]
