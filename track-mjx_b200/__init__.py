"""B200-native batched `reset` / `step` of track-mjx's rodent motion-tracking environment.

Only the hot path is here: the model-constant compiler, the task configuration, the
reference-clip container, the ctypes binding of the C-ABI library (`include/tmjx.h`) and the Brax-style
`Env` mirror (`env.py`).  The CUDA library is mandatory on the compute path: there is no CPU fallback.
"""

__version__ = "0.1.0"
