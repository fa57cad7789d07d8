"""`models.unet_cond_base` of the reference, served by controlnet-pytorch_b200/models/unet_cond_base.py."""
from _cnb200_bootstrap import reexport

reexport(globals(), "models.unet_cond_base")
