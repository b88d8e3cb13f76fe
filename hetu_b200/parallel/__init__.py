from .symm import SymmetricBuffer, symm_available  # noqa: F401
