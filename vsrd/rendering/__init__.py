from .renderers import *  # noqa: F401,F403
from .renderers import UnionField, UnsupportedFieldError, hierarchical_volumetric_rendering, match_union_field  # noqa: F401
from .samplers import inverse_transform_sampler, quadrature_sampler  # noqa: F401
from .utils import ray_casting  # noqa: F401
from . import sdfs  # noqa: F401
