"""`vsrd.datasets` — the dataset surface scripts/main.py touches (`config.datasets.train.function`, `len()`, indexing
through the DistributedDataLoader, and `get_root_dirname`, main.py:83,125).

The KITTI-360 / KITTI-raw readers of the reference (datasets/kitti_360_dataset.py; skimage, pycocotools, files on disk)
are out of scope.  `SyntheticKITTI360Dataset` produces frames with the same per-view dict contract
 from the seeded synthetic scene generator, so the unmodified script runs end to end."""
from .synthetic_dataset import SyntheticKITTI360Dataset  # noqa: F401
