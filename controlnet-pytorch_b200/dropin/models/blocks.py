"""`models.blocks` of the reference, served by controlnet-pytorch_b200/models/blocks.py."""
from _cnb200_bootstrap import reexport

reexport(globals(), "models.blocks")
