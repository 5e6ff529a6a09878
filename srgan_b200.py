"""Importable alias of the package directory (its name contains hyphens): ``import srgan_b200``."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("super_resolution-image-reconstructer-multi_generator_gan_b200")
sys.modules[__name__] = _pkg
