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
        self._views = [None] * self._F       # per-frame dicts of views, built on first use (the per-step path is host-bound)

    @property
    def num_frames(self):
        return self._F

    def frame(self):
        if self.cursor < 0:
            raise RuntimeError("frame() before the first simulate()")
        i = self.cursor % self._F if self.loop else self.cursor
        if i >= self._F:
            raise IndexError("replay exhausted (%d frames)" % self._F)
        fr = self._views[i]
        if fr is None:
            fr = self._views[i] = {k: v[i] for k, v in self.frames.items()}
        return fr

    def window(self, start, length):
        """`length` consecutive frames starting at `start` as [T, ...] views (horizon-batched launches)."""
        if start + length > self._F:
            raise IndexError("window [%d,%d) outside the %d replayed frames" % (start, start + length, self._F))
        return {k: v[start:start + length] for k, v in self.frames.items()}


class HostReplayProvider(FrameProvider):
    """Frames in pinned host memory.  Frame t+1 is uploaded on a dedicated copy stream while the kernels of step t
    run (two device staging sets), so the PCIe transfer overlaps compute; `frame()` makes the compute stream wait
    for the upload of the current frame only."""

    def __init__(self, frames: Dict[str, torch.Tensor], device, loop=True, extra_keys=(), packed=True):
        super().__init__()
        keys = tuple(self.keys) + tuple(extra_keys)
        src = {k: v.contiguous() for k, v in frames.items() if k in keys and v is not None}
        self.device = device
        self.loop = loop
        self._F = next(iter(src.values())).shape[0]
        self.packed = packed and all(v.dtype == torch.float32 and (v[0].numel() * 4) % 16 == 0 for v in src.values())
        if self.packed:
            # one pinned row per frame [root | dof | sensor | extras] and one device staging row per set: a frame is ONE
            # host->device copy (one DMA descriptor instead of one per tensor); the task sees views into the staging row
            sizes = {k: v[0].numel() for k, v in src.items()}
            total = sum(sizes.values())
            self._packed_host = torch.empty(self._F, total, dtype=torch.float32).pin_memory()
            self._packed_stage = [torch.empty(total, dtype=torch.float32, device=device) for _ in range(2)]
            self.host, off = {}, 0
            self.stage = [{}, {}]
            for k, v in src.items():
                n = sizes[k]
                self._packed_host[:, off:off + n].copy_(v.reshape(self._F, n))
                self.host[k] = self._packed_host[:, off:off + n].view(v.shape)
                for s_ in range(2):
                    self.stage[s_][k] = self._packed_stage[s_][off:off + n].view(v.shape[1:])
                off += n
        else:
            self.host = {k: v.pin_memory() for k, v in src.items()}
            self.stage = [{k: torch.empty_like(v[0], device=device) for k, v in self.host.items()} for _ in range(2)]
        self.h2d_bytes_per_frame = sum(v[0].numel() * v.element_size() for v in self.host.values())
        self.copy_stream = torch.cuda.Stream(device=device)
        self._ready = [torch.cuda.Event(), torch.cuda.Event()]     # upload of the set finished (copy stream)
        self._free = [torch.cuda.Event(), torch.cuda.Event()]      # kernels that read the set were enqueued (compute stream)
        self._prefetched = -1

    @property
    def num_frames(self):
        return self._F

    def _upload(self, cursor):
        st = self.stage[cursor & 1]
        i = cursor % self._F if self.loop else cursor
        if i >= self._F:
            return
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._free[cursor & 1])   # the previous user of this staging set is done
            if self.packed:
                self._packed_stage[cursor & 1].copy_(self._packed_host[i], non_blocking=True)
            else:
                for k, v in self.host.items():
                    st[k].copy_(v[i], non_blocking=True)
            self._ready[cursor & 1].record(self.copy_stream)
        self._prefetched = cursor

    def simulate(self):
        self.cursor += 1
        if self._prefetched < self.cursor:
            self._upload(self.cursor)

    def frame(self):
        cur = torch.cuda.current_stream()
        cur.wait_event(self._ready[self.cursor & 1])
        st = self.stage[self.cursor & 1]
        # the other staging set was last read by the previous step, whose kernels are already enqueued: free it and
        # start uploading the next frame into it now
        self._free[(self.cursor + 1) & 1].record(cur)
        self._upload(self.cursor + 1)
        return st
