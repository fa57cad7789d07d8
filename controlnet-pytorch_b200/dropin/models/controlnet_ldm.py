"""`models.controlnet_ldm` of the reference, served by controlnet-pytorch_b200/models/controlnet_ldm.py."""
from _cnb200_bootstrap import reexport

reexport(globals(), "models.controlnet_ldm")
