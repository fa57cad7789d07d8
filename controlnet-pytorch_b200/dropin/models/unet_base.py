"""`models.unet_base` of the reference, served by controlnet-pytorch_b200/models/unet_base.py."""
from _cnb200_bootstrap import reexport

reexport(globals(), "models.unet_base")
