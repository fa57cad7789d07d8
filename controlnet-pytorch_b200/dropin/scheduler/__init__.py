"""Shim for the reference's `scheduler` package: `linear_noise_scheduler` is the B200 drop-in; other modules
(`consistency_scheduler`) fall through to the reference checkout on sys.path."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
