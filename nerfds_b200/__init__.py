"""nerfds_b200: B200-native NeRF-DS volumetric ray-marching path.

Host mirror of the reference's interface for this path (models.NerfModel /
construct_nerf, evaluation.render_image, utils.shard/unshard) on top of the
C-ABI CUDA library ``lib/libnerfds_b200.so`` (include/nerfds_b200.h).
Importing the config / params helpers needs no GPU; constructing a model does,
and fails loudly without the CUDA library -- there is no CPU fallback.
"""
from .config import NerfDSConfig, nerf_ds_config, tiny_config, from_gin_bindings  # noqa: F401
from .params import init_params, flatten_params, unflatten_params, param_count  # noqa: F401

__all__ = ['NerfDSConfig', 'nerf_ds_config', 'tiny_config', 'from_gin_bindings', 'init_params',
           'flatten_params', 'unflatten_params', 'param_count']
