"""`models.distribution_matching_controlnet` of the reference, served by controlnet-pytorch_b200/models/distribution_matching_controlnet.py."""
from _cnb200_bootstrap import reexport

reexport(globals(), "models.distribution_matching_controlnet")
