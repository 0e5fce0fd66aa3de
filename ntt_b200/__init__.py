"""ntt_b200 -- B200-native (sm_100a) 256-bit NTT.  The compute path is libntt256.so
(ntt_b200/csrc, C ABI in include/ntt256.h); this package holds the thin host-side mirror."""
from .lib import lib, Ntt256Error, plan_describe, FIELD_BN254_FR, FIELD_BLS12_381_FR  # noqa: F401
