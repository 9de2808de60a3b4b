"""`GroupedAdam`: clip_grad_norm_ + Adam for many networks in two launches (`mmb_grad_sumsq_group`, `mmb_adam_group`).

The reference builds one `torch.optim.Adam` per network - 2 x num_agents of them for a MARL team
(agents/algorithms/marl/mappo_policy.py:32-37, ippo_policy.py:39-45, happo_policy.py) - and steps them one after the other
in the per-agent loop (runner.py:266-317 -> mappo_trainer.py:143-170), each step preceded by `nn.utils.clip_grad_norm_` over
that network: about 300 small launches per update for TenAnt's 20 networks.  Here all parameters, gradients and moments of
all networks are slices of four flat fp32 buffers:

    opt = GroupedAdam([dict(params=policy.actor.parameters(), lr=5e-4, eps=1e-5, weight_decay=0.0, max_grad_norm=10.0), ...])
    ...backward()...
    opt.collect_grads()   # the gradients autograd just produced -> flat buffer: one multi-tensor copy
    opt.step()            # per-group gradient norm + clipped Adam update of every group: 2 launches
    opt.zero_grad()       # one memset of the flat gradient buffer

* the modules' parameters are re-pointed at views of `opt.flat_params` (values preserved), their `.grad` at views of
  `opt.flat_grads`, so the reference's modules, `state_dict()` / checkpoints and autograd keep working unchanged;
* `opt.flat_grads` is contiguous per group and overall: it is the NCCL all-reduce buffer of the env-sharded runs
  (`dist.GradBuckets`) - no flatten / unflatten copies;
* arithmetic = `torch.optim.Adam` (amsgrad off) after `clip_grad_norm_(params, max_grad_norm)`, see csrc/adam.cu; `state_dict()`
  / `load_state_dict()` carry the step counts and both moment buffers.

There is no CPU path (the parameters must live on a CUDA device).
"""
import math

import torch

from . import _lib as L


class GroupedAdam:
    def __init__(self, groups, betas=(0.9, 0.999)):
        L.lib()
        self.groups = []
        for g in groups:
            params = [p for p in g["params"]]
            if not params:
                raise ValueError("empty parameter group")
            self.groups.append(dict(params=params, lr=float(g.get("lr", 1e-3)), eps=float(g.get("eps", 1e-8)),
                                    weight_decay=float(g.get("weight_decay", 0.0)),
                                    max_grad_norm=float(g.get("max_grad_norm", 0.0) or 0.0), step=0))
        if not 1 <= len(self.groups) <= L.ADAM_MAX_GROUPS:
            raise L.MmbError("GroupedAdam takes 1..%d groups" % L.ADAM_MAX_GROUPS)
        dev = self.groups[0]["params"][0].device
        if dev.type != "cuda":
            raise L.MmbError("GroupedAdam needs CUDA parameters (there is no CPU path); got %s" % dev)
        self.device, self.betas = dev, (float(betas[0]), float(betas[1]))
        # layout: parameters in group order, every group padded to a multiple of 4 elements (128-bit accesses never straddle)
        off, self._starts, self._slots = 0, [], []
        for g in self.groups:
            self._starts.append(off)
            for p in g["params"]:
                if p.device != dev or p.dtype != torch.float32:
                    raise L.MmbError("GroupedAdam: all parameters must be fp32 on %s" % dev)
                self._slots.append((p, off, p.numel()))
                off += p.numel()
            off = (off + 3) // 4 * 4
        self.total = off
        self.flat_params = torch.zeros(off, device=dev)
        self.flat_grads = torch.zeros(off, device=dev)
        self.exp_avg = torch.zeros(off, device=dev)
        self.exp_avg_sq = torch.zeros(off, device=dev)
        self._sumsq = torch.zeros(len(self.groups), device=dev, dtype=torch.float64)
        self._group_slots, k = [], 0
        for g in self.groups:
            self._group_slots.append(self._slots[k:k + len(g["params"])])
            k += len(g["params"])
        with torch.no_grad():
            for p, o, n in self._slots:
                self.flat_params[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_params[o:o + n].view(p.shape)           # the module now reads / writes the flat buffer
        self._p = None

    # ---- gradients ---------------------------------------------------------------------------------------------------
    def attach_grads(self):
        """Points every parameter's `.grad` at its slice of the flat gradient buffer, so autograd accumulates there (one
        in-place add per parameter and backward).  The alternative, cheaper in launches, is `zero_grad()` + `collect_grads()`:
        autograd then hands over its freshly produced gradient tensors and ONE multi-tensor copy moves them."""
        for p, o, n in self._slots:
            p.grad = self.flat_grads[o:o + n].view(p.shape)

    def zero_grad(self, set_to_none=True):
        """Clears the flat gradient buffer (one memset).  set_to_none=True (default, as torch's optimizers): the parameters'
        `.grad` are dropped, the next backward produces fresh tensors and `collect_grads()` gathers them; False: `.grad` stay
        attached to the flat buffer."""
        self.flat_grads.zero_()
        if set_to_none:
            for p, _, _ in self._slots:
                p.grad = None
        else:
            self.attach_grads()

    @torch.no_grad()
    def collect_grads(self, groups=None):
        """Moves the gradients autograd produced since `zero_grad()` into the flat buffer - one multi-tensor copy for all
        listed groups (default: all) - and attaches `.grad` to the flat slices.  Parameters without a gradient keep zeros."""
        want = None if groups is None else set(groups)
        dsts, srcs = [], []
        for gi, (lo, hi) in enumerate(self.group_slice(g) for g in range(len(self.groups))):
            if want is not None and gi not in want:
                continue
            for p, o, n in self._group_slots[gi]:
                view = self.flat_grads[o:o + n].view(p.shape)
                if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
                    dsts.append(view)
                    srcs.append(p.grad)
                p.grad = view
        if dsts:
            torch._foreach_copy_(dsts, srcs)

    def group_slice(self, g):
        """(start, end) of group g in the flat buffers."""
        return self._starts[g], (self._starts[g + 1] if g + 1 < len(self.groups) else self.total)

    def grad_norms(self):
        """Gradient 2-norms per group as of the last `step()` (device tensor, fp64; what clip_grad_norm_ returns)."""
        return self._sumsq.sqrt()

    # ---- the step ----------------------------------------------------------------------------------------------------
    def _params(self):
        p = self._p
        if p is None:
            p = self._p = L.AdamParams()
            p.num_groups, p.total = len(self.groups), self.total
            for i, s in enumerate(self._starts):
                p.group_start[i] = s
            p.params, p.grads = self.flat_params.data_ptr(), self.flat_grads.data_ptr()
            p.exp_avg, p.exp_avg_sq, p.sumsq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self._sumsq.data_ptr()
            p.one_minus_beta1, p.beta2, p.one_minus_beta2 = 1.0 - self.betas[0], self.betas[1], 1.0 - self.betas[1]
        return p

    def step(self):
        """One clipped Adam update of every group from the flat gradient buffer (call `collect_grads()` first unless the
        gradients are attached)."""
        p = self._params()
        b1, b2 = self.betas
        for i, g in enumerate(self.groups):
            g["step"] += 1
            bc1 = 1.0 - b1 ** g["step"]                                  # torch/optim/adam.py: Python doubles
            bc2 = 1.0 - b2 ** g["step"]
            p.step_size[i] = g["lr"] / bc1
            p.bc2_sqrt[i] = math.sqrt(bc2)
            p.eps[i], p.weight_decay[i], p.max_grad_norm[i] = g["eps"], g["weight_decay"], g["max_grad_norm"]
        self._sumsq.zero_()
        st = L.stream_ptr()
        L.check(L.lib().mmb_grad_sumsq_group(p, st), "mmb_grad_sumsq_group")
        L.check(L.lib().mmb_adam_group(p, st), "mmb_adam_group")
        # the kernel wrote the parameters behind autograd's back: bump their version counters, so that everything that
        # watches them (mlp.FusedMLP's bf16 copies, autograd's saved-tensor checks) sees an in-place update
        for p_, _, _ in self._slots:
            torch.autograd.graph.increment_version(p_)

    # ---- checkpointing -----------------------------------------------------------------------------------------------
    def state_dict(self):
        return {"steps": [g["step"] for g in self.groups], "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "hyper": [{k: g[k] for k in ("lr", "eps", "weight_decay", "max_grad_norm")} for g in self.groups]}

    def load_state_dict(self, sd):
        if len(sd["steps"]) != len(self.groups) or sd["exp_avg"].numel() != self.total:
            raise ValueError("GroupedAdam.load_state_dict: layout mismatch")
        for g, s, h in zip(self.groups, sd["steps"], sd["hyper"]):
            g["step"] = int(s)
            g.update(h)
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])

    @classmethod
    def for_marl_policies(cls, policies, config):
        """One optimiser for a team: groups 2a / 2a+1 = agent a's actor / critic with the reference's hyper-parameters
        (mappo_policy.py:24-37: lr, critic_lr, opti_eps, weight_decay; mappo_trainer.py:41: max_grad_norm when
        use_max_grad_norm)."""
        clip = float(config["max_grad_norm"]) if config.get("use_max_grad_norm", True) else 0.0
        groups = []
        for pol in policies:
            groups.append(dict(params=list(pol.actor.parameters()), lr=config["lr"], eps=config["opti_eps"],
                               weight_decay=config["weight_decay"], max_grad_norm=clip))
            groups.append(dict(params=list(pol.critic.parameters()), lr=config["critic_lr"], eps=config["opti_eps"],
                               weight_decay=config["weight_decay"], max_grad_norm=clip))
        return cls(groups)
