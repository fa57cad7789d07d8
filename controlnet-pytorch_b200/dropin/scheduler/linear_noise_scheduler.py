"""`scheduler.linear_noise_scheduler` of the reference, served by controlnet-pytorch_b200/scheduler/."""
from _cnb200_bootstrap import reexport

reexport(globals(), "scheduler.linear_noise_scheduler")
