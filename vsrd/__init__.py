"""Drop-in `vsrd` package: the reference's Python API surface for the per-frame optimisation loop
(scripts/main.py), backed by the vsrd_b200 sm_100a kernels.  Only what the hot path touches is
provided: rendering, models, operations, losses, utils, configuration, distributed, the
logging-time drawing helpers and a synthetic dataset with the reference's per-view dict contract; the on-disk dataset
readers and preprocessing transforms are out of scope."""
from . import configuration  # noqa: F401
from . import datasets  # noqa: F401
from . import distributed  # noqa: F401
from . import losses  # noqa: F401
from . import models  # noqa: F401
from . import operations  # noqa: F401
from . import rendering  # noqa: F401
from . import utils  # noqa: F401
from . import visualization  # noqa: F401
