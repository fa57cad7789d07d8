"""B200-native ControlNet denoising hot path (drop-in for henriChevreux/ControlNet-PyTorch's
models.* / scheduler.* modules on that path).  See DESIGN.md and INTEGRATION.md.

The directory name carries a hyphen, so import it with
    importlib.import_module("controlnet-pytorch_b200")
or put `dropin/` (shim packages `models` / `scheduler` that re-export the modules here) first on sys.path and use
the reference's own import lines (`from models.controlnet import ControlNet`, `from scheduler.linear_noise_scheduler
import ...`); see INTEGRATION.md section 1.
"""
__all__ = ["models", "scheduler", "utils", "ops", "runtime"]
