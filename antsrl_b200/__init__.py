"""antsrl_b200 -- B200-native (sm_100a CUDA) implementation of the AntsRL environment step loop
(RLApi.observation / RLApi.step / Environment.update of SelennLamson/AntsRL) over batches of independent
environments.  `BatchedAnts` is the batched host object over the C ABI (include/antsrl_b200.h);
`antsrl_b200.dropin` mirrors the reference's `environment` / `generator` packages for existing agents."""
from ._cabi import AntsError, LIB_PATH, load_library          # noqa: F401
from .batch import BatchedAnts, make_config, DEFAULT_MASK, KERNEL_FAMILIES   # noqa: F401

__version__ = "0.1.0"


def dropin_path():
    """Directory to put on sys.path in place of the reference checkout: it provides `environment`, `generator` and
    `utils` with the reference's names and signatures, backed by the CUDA step loop."""
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")
