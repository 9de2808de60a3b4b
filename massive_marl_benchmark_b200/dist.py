"""Env-sharded data parallelism (one process per GPU, torch.distributed over NCCL/NVLink).

The reference has no multi-GPU code.  The hot path shards naturally: every op of
the task pipeline is row-wise over environments, so GPU g of G owns envs [g*N/G, (g+1)*N/G) with its own
frame provider, buffers and rollout storage, and NO data-path collective is needed for obs / reward /
reset / GAE.  Only two reductions cross environments:

  1. advantage normalisation: (count, sum, sumsq) - three fp64 per storage (per agent for MARL) -
     `all_reduce_stats` between mmb_gae_* and mmb_adv_normalize;
  2. the PPO/MAPPO gradient all-reduce after backward - `all_reduce_grads` (sum / world).

The first has two implementations.  `StatsExchange` is the product path: the normalise kernel's first block
stores the shard's three doubles + a sequence flag into every rank's mailbox with NVLink peer stores, then all
blocks wait on their own mailbox and normalise - exchange and compute in ONE launch, no collective kernel, no
host involvement, CUDA-graph replayable.  `all_reduce_stats` (a plain NCCL all-reduce between
the two launches) is kept as the baseline it is measured against.  The second belongs to autograd (out of
scope for the kernels) and is a bucketed NCCL all-reduce.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun); returns
    (rank, world, local_rank).  World size 1 without the env vars -> no process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def bind_to_gpu_numa(local_rank: int) -> bool:
    """Pin this process to the CPU cores NVML reports as local to GPU `local_rank`, so pinned host buffers allocated
    afterwards are first-touched on the GPU's NUMA node (the host->device path of the end-to-end loop then does not
    cross the socket interconnect).  Returns False (and changes nothing) when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def shard_range(num_envs: int, rank: int, world: int):
    """Contiguous env shard of `rank`: [lo, hi).  Remainder envs go to the lowest ranks."""
    base, rem = divmod(num_envs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_stats(stats: torch.Tensor, group=None):
    """Sum the (count, sum, sumsq) fp64 triples over the env shards, in place."""
    if dist.is_initialized() and dist.get_world_size(group if group is not True else None) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=None if group is True else group)
    return stats


def all_reduce_grads(parameters, group=None, bucket_bytes=64 << 20):
    """Average gradients over the shards: flattened into a few large buckets (sized for launch latency and
    overlap, not link count - NVSwitch gives every peer full bandwidth)."""
    if not dist.is_initialized():
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    grads = [p.grad for p in parameters if p.grad is not None]
    bucket, size = [], 0
    def flush():
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
            bucket, size = [], 0
    flush()


def _avg_op(group=None):
    """ReduceOp.AVG where the backend has it (NCCL: the division happens inside the collective), else None (SUM + div_)."""
    try:
        return dist.ReduceOp.AVG if dist.get_backend(group) == "nccl" else None
    except Exception:
        return None


class FlatGradReducer:
    """Asynchronous all-reduce (average) of slices of ONE flat gradient buffer - `GroupedAdam.flat_grads` - so that the
    collective of one network overlaps the forward / backward of the next:

        red = FlatGradReducer(opt.flat_grads)
        for agent in team:                      # per-agent loop of the update
            loss(agent).backward(); opt.collect_grads(groups_of(agent))
            red.reduce_async(*slice_of(agent))  # NCCL runs on its own stream, behind this agent's gradient writes only
        red.wait()                              # before opt.step()

    No flatten / unflatten copies: the slice IS the communication buffer.  `bytes_reduced` / `calls` feed the bench's bus
    bandwidth figure.  Identity without a process group (world 1)."""

    def __init__(self, flat, group=None):
        self.flat, self.group = flat, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._works = []
        self.bytes_reduced, self.calls = 0, 0

    def reduce_async(self, lo, hi):
        if self.world == 1 or hi <= lo:
            return
        view = self.flat[lo:hi]
        avg = _avg_op(self.group)
        work = dist.all_reduce(view, op=avg if avg is not None else dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._works.append((work, view, avg is None))
        self.bytes_reduced += view.numel() * view.element_size()
        self.calls += 1

    def wait(self):
        for work, view, need_div in self._works:
            work.wait()
            if need_div:
                view.div_(self.world)
        self._works = []


class OverlappedGradAllReduce:
    """Bucketed gradient all-reduce (average) overlapped with backward, for one network trained data-parallel over env
    shards (the PPO `ActorCritic`, 16 MB of fp32 gradients): a `post_accumulate_grad_hook` on every parameter counts the
    bucket down; the moment the last gradient of a bucket exists the bucket is packed (one multi-tensor copy into its
    slice of a flat buffer) and its all-reduce is issued asynchronously, while autograd keeps producing the gradients of
    the earlier layers.  `finish()` (after `backward()`) waits for the collectives and points every `.grad` at its reduced
    slice - no copy back.  Buckets are filled in reverse parameter order (the order backward produces gradients) and
    sized for launch latency and overlap, not link count (NVSwitch gives every peer full bandwidth).
    Identity without a process group."""

    def __init__(self, parameters, group=None, bucket_bytes=4 << 20):
        self.params = [p for p in parameters if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bytes_reduced, self.calls = 0, 0
        self._handles, self._works = [], []
        if self.world == 1 or not self.params:
            self.buckets = []
            return
        dev, dtype = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, device=dev, dtype=dtype)
        self.buckets, cur, cur_bytes, off = [], [], 0, 0
        for p in reversed(self.params):
            cur.append((p, off, p.numel()))
            off += p.numel()
            cur_bytes += p.numel() * p.element_size()
            if cur_bytes >= bucket_bytes:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(cur)
        self._pending = [len(b) for b in self.buckets]
        for bi, b in enumerate(self.buckets):
            for p, _, _ in b:
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))

    def _make_hook(self, bi):
        def hook(_param):
            self._pending[bi] -= 1
            if self._pending[bi] == 0:
                self._launch(bi)
        return hook

    @torch.no_grad()
    def _launch(self, bi):
        b = self.buckets[bi]
        lo, hi = b[0][1], b[-1][1] + b[-1][2]
        views = [self.flat[o:o + n].view(p.shape) for p, o, n in b]
        torch._foreach_copy_(views, [p.grad for p, _, _ in b])
        avg = _avg_op(self.group)
        work = dist.all_reduce(self.flat[lo:hi], op=avg if avg is not None else dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._works.append((work, bi, views, avg is None))
        self.bytes_reduced += (hi - lo) * self.flat.element_size()
        self.calls += 1

    @torch.no_grad()
    def finish(self):
        """Call after `backward()`: buckets whose hooks did not all fire (parameters without a gradient this step) are
        flushed with zeros for the missing ones, every collective is awaited, `.grad` = the averaged slice."""
        if self.world == 1:
            return
        for bi, n in enumerate(self._pending):
            if n != 0 and n != len(self.buckets[bi]):       # partially produced bucket: reduce what exists (others: zeros)
                for p, o, m in self.buckets[bi]:
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                self._launch(bi)
        for work, bi, views, need_div in self._works:
            work.wait()
            if need_div:
                lo, hi = self.buckets[bi][0][1], self.buckets[bi][-1][1] + self.buckets[bi][-1][2]
                self.flat[lo:hi].div_(self.world)
            for (p, _, _), v in zip(self.buckets[bi], views):
                p.grad = v
        self._works = []
        self._pending = [len(b) for b in self.buckets]

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []


def mean_over_ranks(t: torch.Tensor, group=None) -> torch.Tensor:
    """Average of a (scalar) tensor over the shards - e.g. the KL estimate that drives the adaptive step size, which every
    rank must see identically.  Returns a new tensor; identity without a process group."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    out = t.detach().clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out / dist.get_world_size(group)


def max_over_ranks(value: float, device) -> float:
    """Device-timed durations are reported as the max over ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


class StatsExchange:
    """Per-process endpoint of the peer-memory statistics exchange (include/mmb.h, `mmb_xchg`).

    `StatsExchange(group)` allocates this rank's mailbox, trades the 64-byte CUDA IPC handles through
    torch.distributed and maps every peer's mailbox.  Attach it to any number of `RolloutStorage`s
    (`storage.stats_exchange = xchg`); exchanges are numbered by device-side counters in issue order, so all ranks
    must issue the same sequence of normalize_advantages calls on storages attached to it.

    `StatsExchange.local(world, slots)` builds `world` endpoints inside ONE process on one device (mailboxes are
    ordinary local allocations, no IPC) - used by the single-GPU tests of the protocol."""

    def __init__(self, group=None, slots=4, device=None, timeout_ms=0):
        from . import _lib as L
        import ctypes as C
        lib = L.lib()
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > L.MAX_RANKS:
            raise L.MmbError("StatsExchange supports up to %d ranks (one NVLink domain); got %d" % (L.MAX_RANKS, self.world))
        self.slots = slots
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        nbytes = lib.mmb_xchg_mailbox_bytes(self.world, slots)
        if nbytes <= 0:
            raise L.MmbError("bad exchange geometry world=%d slots=%d" % (self.world, slots))
        ptr, handle = C.c_void_p(), (C.c_uint8 * 64)()
        with torch.cuda.device(self.device):
            L.check(lib.mmb_xchg_alloc(nbytes, C.byref(ptr), handle), "mmb_xchg_alloc")
            self._local = ptr.value
            self._opened = []
            handles = [bytes(handle)]
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(handle), group=group)
            self.state = torch.zeros(8, dtype=torch.int64, device=self.device)
            self.desc = L.Xchg()
            self.desc.world, self.desc.rank, self.desc.slots = self.world, self.rank, slots
            self.desc.timeout_ms = int(timeout_ms)     # 0: library default (10 s); < 0: wait for ever (what NCCL would do)
            self.desc.state = self.state.data_ptr()
            for r in range(self.world):
                if r == self.rank:
                    self.desc.mailbox[r] = self._local
                else:
                    peer, hb = C.c_void_p(), (C.c_uint8 * 64).from_buffer_copy(handles[r])
                    L.check(lib.mmb_xchg_open(hb, C.byref(peer)), "mmb_xchg_open (rank %d)" % r)
                    self._opened.append(peer.value)
                    self.desc.mailbox[r] = peer.value
            torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=group)      # nobody publishes before every mailbox is mapped

    @classmethod
    def local(cls, world, slots=4, device="cuda", timeout_ms=0):
        """`world` endpoints in this process (protocol tests): returns a list of StatsExchange-like objects."""
        from . import _lib as L
        nwords = int(L.lib().mmb_xchg_mailbox_bytes(world, slots)) // 8
        boxes = [torch.zeros(nwords, dtype=torch.int64, device=device) for _ in range(world)]
        out = []
        for r in range(world):
            x = cls.__new__(cls)
            x.world, x.rank, x.slots, x.device = world, r, slots, torch.device(device)
            x._local, x._opened, x._boxes = None, [], boxes
            x.state = torch.zeros(8, dtype=torch.int64, device=device)
            x.desc = L.Xchg()
            x.desc.world, x.desc.rank, x.desc.slots, x.desc.state = world, r, slots, x.state.data_ptr()
            x.desc.timeout_ms = int(timeout_ms)
            for k in range(world):
                x.desc.mailbox[k] = boxes[k].data_ptr()
            out.append(x)
        return out

    @property
    def errors(self) -> int:
        """Exchanges whose flags timed out or whose slot was overrun (host sync)."""
        return int(self.state[3].item())

    def close(self):
        from . import _lib as L
        lib = L.lib()
        for p in self._opened:
            lib.mmb_xchg_close(p)
        self._opened = []
        if self._local:
            lib.mmb_xchg_free(self._local)
            self._local = None
