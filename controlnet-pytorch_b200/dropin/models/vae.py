"""`models.vae` of the reference, served by controlnet-pytorch_b200/models/vae.py."""
from _cnb200_bootstrap import reexport

reexport(globals(), "models.vae")
