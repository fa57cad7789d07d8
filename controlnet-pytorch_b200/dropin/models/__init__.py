"""Shim for the reference's `models` package: the hot-path modules next to this file re-export the B200 drop-ins;
every other `models.*` module (discriminator, lpips, ...) falls through to the reference checkout on sys.path."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
