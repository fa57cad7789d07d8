"""`models.controlnet` of the reference, served by controlnet-pytorch_b200/models/controlnet.py."""
from _cnb200_bootstrap import reexport

reexport(globals(), "models.controlnet")
