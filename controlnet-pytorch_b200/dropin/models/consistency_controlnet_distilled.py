"""`models.consistency_controlnet_distilled` of the reference, served by controlnet-pytorch_b200/models/consistency_controlnet_distilled.py."""
from _cnb200_bootstrap import reexport

reexport(globals(), "models.consistency_controlnet_distilled")
