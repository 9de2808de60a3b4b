"""Frame providers: the stand-in for Isaac Gym's `simulate` + `refresh_*_tensor` (out of scope: PhysX).

A provider owns state frames in Isaac Gym's exact tensor layout and hands the
task the frame that `gym.refresh_*` would have exposed.  It also plays the simulator's role as the sink
of the task's side effects (`set_dof_actuation_force_tensor`, `set_*_state_tensor_indexed`,
`apply_rigid_body_force_tensors`): the tensors are kept (device-side, no host sync) so tests can
compare them with the reference's.

  ReplayProvider      frames resident in HBM, `[F, rows, cols]` per tensor; zero-copy views per step
  HostReplayProvider  frames in pinned host memory; every step copies one frame host->device on the
                      current stream (the end-to-end path bench.py times)
"""
from typing import Dict, Optional

import torch


class FrameProvider:
    keys = ("root", "dof", "sensor")

    def __init__(self):
        self.cursor = -1
        self.sink: Dict[str, object] = {}

    # --- simulator side ---------------------------------------------------------------------
    def simulate(self):
        self.cursor += 1

    def frame(self) -> Dict[str, Optional[torch.Tensor]]:
        raise NotImplementedError

    @property
    def num_frames(self) -> int:
        raise NotImplementedError

    # --- sinks (what the reference passes to gym.set_* / gym.apply_*) ------------------------
    def set_dof_actuation_force_tensor(self, forces):
        self.sink["dof_forces"] = forces

    def apply_rigid_body_force_tensors(self, forces):
        self.sink["body_forces"] = forces

    def set_actor_root_state_tensor_indexed(self, states, indices, count):
        self.sink["root_indexed"] = (states, indices, count)

    def set_dof_state_tensor_indexed(self, states, indices, count):
        self.sink["dof_indexed"] = (states, indices, count)


class ReplayProvider(FrameProvider):
    """Frames already in device memory: dict of `[F, ...]` tensors (keys root / dof / sensor)."""

    def __init__(self, frames: Dict[str, torch.Tensor], device=None, loop=True):
        super().__init__()
        self.frames = {k: (v if device is None else v.to(device)).contiguous() for k, v in frames.items()
                       if k in self.keys and v is not None}
        self.loop = loop
        self._F = next(iter(self.frames.values())).shape[0]

    @property
    def num_frames(self):
        return self._F

    def frame(self):
        if self.cursor < 0:
            raise RuntimeError("frame() before the first simulate()")
        i = self.cursor % self._F if self.loop else self.cursor
        if i >= self._F:
            raise IndexError("replay exhausted (%d frames)" % self._F)
        return {k: v[i] for k, v in self.frames.items()}

    def window(self, start, length):
        """`length` consecutive frames starting at `start` as [T, ...] views (horizon-batched launches)."""
        if start + length > self._F:
            raise IndexError("window [%d,%d) outside the %d replayed frames" % (start, start + length, self._F))
        return {k: v[start:start + length] for k, v in self.frames.items()}


class HostReplayProvider(FrameProvider):
    """Frames in pinned host memory; `frame()` uploads the current one (async on the current stream)
    into one of two device staging sets, so a frame stays valid while the next one is uploaded."""

    def __init__(self, frames: Dict[str, torch.Tensor], device, loop=True):
        super().__init__()
        self.host = {k: v.contiguous().pin_memory() for k, v in frames.items() if k in self.keys and v is not None}
        self.device = device
        self.loop = loop
        self._F = next(iter(self.host.values())).shape[0]
        self.stage = [{k: torch.empty_like(v[0], device=device) for k, v in self.host.items()} for _ in range(2)]
        self.h2d_bytes_per_frame = sum(v[0].numel() * v.element_size() for v in self.host.values())

    @property
    def num_frames(self):
        return self._F

    def frame(self):
        i = self.cursor % self._F if self.loop else self.cursor
        st = self.stage[self.cursor & 1]
        for k, v in self.host.items():
            st[k].copy_(v[i], non_blocking=True)
        return st
