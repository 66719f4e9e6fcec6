"""Import shim: the package directory is ``deepsc-gan_b200/`` (hyphen, not importable by name), so this
module registers it in ``sys.modules`` as the package ``deepsc_gan_b200``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "deepsc-gan_b200")
_spec = importlib.util.spec_from_file_location("deepsc_gan_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["deepsc_gan_b200"] = _mod
_spec.loader.exec_module(_mod)
