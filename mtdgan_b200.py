"""Import alias: `import mtdgan_b200` resolves to the package directory ./mtd-gan_b200/ (the directory
name the build layout prescribes is not a valid Python identifier, so this module lends it one)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "mtd-gan_b200")]
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__

from mtdgan_b200._ext import lib_path, is_built, require_cuda_extension  # noqa: E402,F401


def invalidate_weight_caches():
    """Drop every packed / TF32-split weight copy (needed after in-place writes through `param.data`, which do not
    move the parameter's version counter)."""
    from mtdgan_b200 import ops
    ops.clear_pack_cache()
