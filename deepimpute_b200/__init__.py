"""deepimpute_b200 -- B200-native hot path of DeepImpute behind the reference's MultiNet / deepImpute API.

Public surface mirrors the reference package (`/root/reference/deepimpute/__init__.py`, `multinet.py:65`,
`deepImpute.py:6`): ``MultiNet(...).fit(df).predict(df)`` and ``deepImpute(**kwargs)``.  The neural-network
engine underneath (``engine.Engine``) is hand-written sm_100a CUDA reached through a C-ABI shared library
(``include/deepimpute_b200.h``); there is no CPU fallback.
"""
from .multinet import MultiNet, get_distance_matrix, inspect_data, wMSE  # noqa: F401
from .deepImpute import deepImpute  # noqa: F401

__version__ = "0.1.0"
