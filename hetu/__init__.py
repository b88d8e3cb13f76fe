"""`import hetu` compatibility alias: the reference's package name resolves to hetu_b200 (same API surface).  Every
sub-module path works too -- `hetu.nn`, `hetu.engine.trainer`, `hetu.models.gpt.generate_gpt_4d_config`,
`python -m hetu.rpc.pssh_start_config` ... -- through an import hook that maps `hetu.X` onto the already-imported (or
importable) `hetu_b200.X` module object, so there is exactly one copy of every module."""
import importlib
import importlib.abc
import importlib.util
import sys

import hetu_b200 as _impl
from hetu_b200 import *  # noqa: F401,F403


from hetu_b200._refpaths import AliasLoader as _AliasLoader


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith("hetu."):
            return None
        real_name = "hetu_b200." + fullname[len("hetu."):]
        try:
            real = importlib.import_module(real_name)
        except ModuleNotFoundError as e:
            if e.name and real_name.startswith(e.name):
                return None
            raise
        spec = importlib.util.spec_from_loader(fullname, _AliasLoader(real), origin=getattr(real.__spec__, "origin", None),
                                               is_package=hasattr(real, "__path__"))
        return spec


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
__path__ = list(_impl.__path__)          # lets `python -m hetu.x.y` locate sub-modules before the hook is consulted


def __getattr__(name):
    return getattr(_impl, name)


__version__ = _impl.__version__
