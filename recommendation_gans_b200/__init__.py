"""B200-native implicit matrix factorisation behind Spotlight's API.

Importing this package puts the drop-in modules (`implicit`, `spotlight.*`, `utils.*` -- the names
the reference's `mf_spotlight.py` imports) on `sys.path`, so

    import recommendation_gans_b200          # noqa
    from implicit import ImplicitFactorizationModel
    from spotlight.factorization.representations import BilinearNet

resolve to the CUDA-backed implementations in `recommendation_gans_b200/dropin/`.
"""
import os
import sys

__version__ = '0.1.0'

DROPIN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'dropin')


def install_dropin():
    """Make `implicit`, `spotlight`, `utils` importable (idempotent)."""
    if DROPIN_PATH not in sys.path:
        sys.path.insert(0, DROPIN_PATH)
    return DROPIN_PATH


install_dropin()
