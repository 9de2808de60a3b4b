"""`gym.spaces` compatibility: use gym's Space/Box when gym is installed (the reference type-checks
`isinstance(space, gym.spaces.Space)`, agents/algorithms/rl/ppo/ppo.py:34-39), else minimal stand-ins
with the same attributes (`shape`, `low`, `high`)."""
import numpy as np

try:  # pragma: no cover - depends on the environment
    from gym.spaces import Box, Space  # type: ignore
except Exception:  # gym is not in this image
    class Space:
        def __init__(self, shape=None, dtype=None):
            self.shape = None if shape is None else tuple(shape)
            self.dtype = dtype

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.shape(low)
            super().__init__(shape, dtype)
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape)
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape)
