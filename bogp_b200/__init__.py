"""bogp_b200 -- B200-native (sm_100a) hot path of NeptuneProjects/BOGP.

Boundary A (matched-field objective): :mod:`bogp_b200.mfp`  -- ``MatchedFieldProcessor``, ``ModeRunner``, ``beamformer``
Boundary B (GP posterior + acquisition): :mod:`bogp_b200.gp`, :mod:`bogp_b200.acquisition`, :mod:`bogp_b200.optim`
Sampling strategies and BAxUS: :mod:`bogp_b200.generation` (``MaxPosteriorSampling``), :mod:`bogp_b200.baxus`
Callers of the path: :mod:`bogp_b200.oao`, :mod:`bogp_b200.strategy`, :mod:`bogp_b200.sweep`, :mod:`bogp_b200.shim`
Data contract and synthetic inputs: :mod:`bogp_b200.modes`; KRAKEN ``.mod`` files -> ``ModeSet``: :mod:`bogp_b200.kraken_mod`
Multi-GPU sharding and the peer-memory final reduction: :mod:`bogp_b200.sharding`
C ABI underneath: ``include/bogp.h`` / ``bogp_b200/libbogp_sm100.so`` (``python -m bogp_b200.build``)
"""
from .modes import ModeSet, pekeris_modes, synthetic_covariance  # noqa: F401

__version__ = "0.1.0"
