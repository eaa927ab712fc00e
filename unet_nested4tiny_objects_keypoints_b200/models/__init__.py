"""Mirror of the reference's ``models`` package surface for the hot path: the trainer selects a
model with ``getattr(models, args.model_name)()`` (reference trainer/trainer.py:337,340)."""
from .unet import UNet_Nested, unetConv2, unetUp, init_weights, weights_init_kaiming, count_param  # noqa: F401
