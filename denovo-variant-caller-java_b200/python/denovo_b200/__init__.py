"""denovo_b200 — B200-native trio de novo hot path (Python host binding over the C ABI)."""
from ._lib import (CALL_DT, PCALL_DT, SITE_DT, VCALL_DT, DenovoError, build, load,  # noqa: F401
                   DNV_E_BADBASE, DNV_E_CAPACITY, DNV_E_CUDA, DNV_E_INVALID, DNV_E_UNSORTED, EXPORTS, LIB_PATH)
from .engine import ALLELES, GENOTYPES, DeviceGroup, Gather, TrioEngine, poisson_thresholds  # noqa: F401
from . import sharding  # noqa: F401,E402
