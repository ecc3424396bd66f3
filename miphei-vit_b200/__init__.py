"""miphei-vit_b200 — B200-native (sm_100a) implementation of the MIPHEI-ViT generator hot path.

The directory name carries a hyphen (it mirrors the reference repository's name); import it as
`miphei_vit_b200` (alias package at the repo root) or through importlib.
"""
from . import lib  # noqa: F401

__all__ = ["lib"]
