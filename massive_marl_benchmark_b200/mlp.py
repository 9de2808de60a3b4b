"""Actor-critic MLP forward on the tcgen05 kernels (`mmb_mlp_layer`, `mmb_ln_cast`).

Covers the forward passes on the rollout path:

  PPO   `ActorCritic.actor / .critic`  agents/algorithms/rl/ppo/module.py:25-55
        Linear(obs,1024) ELU Linear(1024,1024) ELU Linear(1024,512) ELU Linear(512,act|1)
  MARL  `Actor.base + act.action_out.fc_mean`, `Critic.base + v_out`
        agents/algorithms/marl/actor_critic.py:42-69,149-168, agents/algorithms/utils/mlp.py:6-65
        LayerNorm(in) [Linear ELU LayerNorm] x3 Linear(512, 8|1)

Weights are taken from the reference's own modules / checkpoints (`from_sequential`, `from_marl_state_dict`), cast
to bf16 and zero-padded to the kernel's tile multiples, and kept in step with the live parameters: every forward compares
the source tensors' version counters with those of the last cast and re-casts in place after an `optimizer.step()` /
`load_state_dict` (`refresh_from` re-binds to another module / state dict of the same architecture).  On top of the means / values: `gaussian_act` (sampling and
log-probs in one launch, incl. PPO's sigma^2 scale_tril quirk, module.py:76-77), `PPOActorCriticForward` (actor and critic
as one grouped launch per layer), `GroupedMLP` / `MarlTeamForward` (all agents' networks in one launch per layer).
Numerics: bf16 operands, fp32 accumulation and fp32 bias / ELU / LayerNorm; the reference is fp32 SGEMM, so
outputs agree to bf16 operand precision (tests/test_gpu_mlp.py states the tolerances).
"""
import ctypes as C
import os
from typing import List, Optional

import torch

from . import _lib as L


def _round_up(x, m):
    return (x + m - 1) // m * m


def _ln_parts(ln):
    """(weight, bias, eps) of a LayerNorm given as an nn.LayerNorm or as that tuple (live tensors, not copies)."""
    if ln is None:
        return None
    if isinstance(ln, (tuple, list)):
        return ln[0], ln[1], float(ln[2])
    return ln.weight, ln.bias, float(ln.eps)


class _Layer:
    """One Linear [+ELU [+LayerNorm]]: the kernel-side copies (bf16 weight zero-padded to the tile multiples, fp32 bias /
    gamma / beta) and the LIVE source tensors they were cast from.  `stale()` compares the sources' version counters
    with those seen at the last cast (an in-place optimizer step or load_state_dict bumps them); `refresh()` re-casts in
    place into the existing buffers, so cached activation buffers / CUDA graphs built on them stay valid.
    `n_out` > weight.shape[0] zero-pads the output width (the critic's 1-wide head run next to the actor's)."""

    def __init__(self, weight, bias, act: bool, ln=None, device="cuda", n_out=None, fold=None, stats_out=False):
        n_src, K = weight.shape
        N = n_src if n_out is None else int(n_out)
        self.N, self.K, self.n_src = N, K, n_src
        self.Kpad = _round_up(K, 64)
        self.act = act
        self.ln = _ln_parts(ln)
        # Deferred LayerNorm (include/mmb.h, `ln_in_stats`): `fold` = (gamma, beta, eps) of the LayerNorm that precedes this
        # layer's Linear - gamma is multiplied into the weight columns, W beta into the bias, c[n] = sum_k of the folded bf16
        # weights; `stats_out`: this layer's epilogue writes the partial row sums its successor normalises with.
        self.fold = _ln_parts(fold)
        self.stats_out = bool(stats_out)
        if self.fold is not None:
            self.c = torch.zeros(N, dtype=torch.float32, device=device)
        if self.ln is not None:                  # bias + ELU + LayerNorm over the whole row: one CTA owns full rows
            if N not in (32, 64, 96, 128, 160, 192, 224, 256, 512):
                raise L.MmbError("LayerNorm epilogue needs N <= 256 (multiple of 32) or N == 512, got %d" % N)
            self.n_tile, self.Npad, self.epilogue = N, N, 2
        else:                                    # n_tile is chosen per call from M (see FusedMLP._n_tile)
            self.Npad = _round_up(N, 256) if N >= 256 else _round_up(N, 32)
            self.n_tile = None
            self.epilogue = 1 if act else 0
        self.w = torch.zeros(self.Npad, self.Kpad, dtype=torch.bfloat16, device=device)
        self.bias = torch.zeros(N, dtype=torch.float32, device=device)
        if self.ln is not None:
            self.gamma = torch.zeros(N, dtype=torch.float32, device=device)
            self.beta = torch.zeros(N, dtype=torch.float32, device=device)
            self.eps = self.ln[2]
        self._src, self._ver = None, None
        self.refresh(weight, bias, ln)

    def _sources(self):
        return [t for t in (self._src[0], self._src[1]) + ((self.ln[0], self.ln[1]) if self.ln is not None else ()) +
                ((self.fold[0], self.fold[1]) if self.fold is not None else ())]

    def stale(self):
        return tuple(t._version for t in self._sources()) != self._ver

    @torch.no_grad()
    def refresh(self, weight=None, bias=None, ln=None, fold=None):
        """Re-cast from the (given or remembered) source tensors into the existing kernel-side buffers."""
        if weight is not None:
            if tuple(weight.shape) != (self.n_src, self.K):
                raise L.MmbError("refresh: weight shape %r does not match the layer (%d, %d)" % (tuple(weight.shape), self.n_src, self.K))
            self._src = (weight, bias)
            if self.ln is not None and ln is not None:
                self.ln = _ln_parts(ln)
            if self.fold is not None and fold is not None:
                self.fold = _ln_parts(fold)
        w, b = self._src
        if self.fold is not None:
            dev = self.w.device
            wf32 = w.detach().to(dev, torch.float32)
            gamma, beta = self.fold[0].detach().to(dev, torch.float32), self.fold[1].detach().to(dev, torch.float32)
            self.w[:self.n_src, :self.K].copy_(wf32 * gamma[None, :])          # W gamma -> bf16
            self.c[:self.n_src].copy_(self.w[:self.n_src, :self.K].float().sum(dim=1))   # of the weights AS STORED
            self.bias[:self.n_src].copy_(b.detach().to(dev, torch.float32) + wf32 @ beta)
        else:
            self.w[:self.n_src, :self.K].copy_(w.detach(), non_blocking=True)       # fp32 -> bf16 in the copy kernel
            self.bias[:self.n_src].copy_(b.detach(), non_blocking=True)
        if self.ln is not None:
            self.gamma.copy_(self.ln[0].detach(), non_blocking=True)
            self.beta.copy_(self.ln[1].detach(), non_blocking=True)
        self._ver = tuple(t._version for t in self._sources())


_CHAIN_ENABLED = os.environ.get("MMB_MLP_CHAIN", "1") != "0"
_DEFER_LN = os.environ.get("MMB_MLP_DEFER_LN", "1") != "0"     # MARL trunks: LayerNorms applied by the consuming layer


def _chain_launch(owner, arr, num_layers, count, stream, x_fp32=None):
    """The whole chain in ONE launch (`mmb_mlp_chain`: a cluster of 4 CTAs per 128-row block walks all layers) when the
    geometry allows it - plain Linear-ELU chains with hidden widths that are multiples of 256 (the PPO networks) - else False
    and the caller launches layer by layer.  The verdict of the library is remembered per object."""
    if not _CHAIN_ENABLED or owner.__dict__.get("_chain_ok") is False or num_layers < 2:
        return False
    rc = L.lib().mmb_mlp_chain(arr, num_layers, count, x_fp32, stream)
    if rc == 0:
        owner._chain_ok = True
        return True
    if rc != -4:                      # MMB_EUNSUPPORTED: geometry outside the fused kernel; anything else is an error
        L.check(rc, "mmb_mlp_chain")
    owner._chain_ok = False
    return False


class FusedMLP:
    """A chain of Linear [+ELU [+LayerNorm]] layers, optionally preceded by a LayerNorm of the input."""

    def __init__(self, layers: List[_Layer], in_ln=None, device="cuda"):
        L.lib()
        self.layers = layers
        self.device = torch.device(device)
        self.in_dim = layers[0].K
        self.out_dim = layers[-1].N
        self.in_ln = _ln_parts(in_ln)
        if self.in_ln is not None:
            self.in_gamma = torch.zeros(self.in_dim, dtype=torch.float32, device=device)
            self.in_beta = torch.zeros(self.in_dim, dtype=torch.float32, device=device)
            self.in_eps = self.in_ln[2]
        self._in_ver = None
        self._bufs, self._stats = {}, {}
        self.auto_refresh = True      # forward() re-casts the weights when their source tensors have changed
        self._refresh_in_ln()

    # -- keeping the kernel-side copies in step with the live parameters ----------------------------------
    @torch.no_grad()
    def _refresh_in_ln(self):
        if self.in_ln is not None:
            self.in_gamma.copy_(self.in_ln[0].detach(), non_blocking=True)
            self.in_beta.copy_(self.in_ln[1].detach(), non_blocking=True)
            self._in_ver = (self.in_ln[0]._version, self.in_ln[1]._version)

    def _flat_sources(self):
        """Every live tensor the kernel-side copies were cast from, in a fixed order (rebuilt after a re-binding refresh)."""
        src = self.__dict__.get("_flat_src")
        if src is None:
            src = list(self.in_ln[:2]) if self.in_ln is not None else []
            for l in self.layers:
                src.extend(l._sources())
            self._flat_src = src
            self._flat_ver = None
        return src

    def stale(self):
        """True when a source parameter was modified in place since the last cast (optimizer.step(), load_state_dict)."""
        ver = tuple([t._version for t in self._flat_sources()])
        if ver == self._flat_ver:
            return False
        if self.in_ln is not None and (self.in_ln[0]._version, self.in_ln[1]._version) != self._in_ver:
            return True
        if any(l.stale() for l in self.layers):
            return True
        self._flat_ver = ver          # (the per-layer records agree: remember the flat tuple for the one-compare fast path)
        return False

    def refresh(self):
        """Re-cast every layer from its source tensors, in place (buffers, pointers and captured graphs stay valid)."""
        self._refresh_in_ln()
        for l in self.layers:
            l.refresh()
        self._flat_src = None

    def refresh_from_sequential(self, seq):
        """Re-bind to (and re-cast from) another nn.Sequential of the same architecture, e.g. after the module object was
        replaced; in place."""
        lins = [m for m in seq if isinstance(m, torch.nn.Linear)]
        if len(lins) != len(self.layers):
            raise L.MmbError("refresh_from_sequential: %d Linear layers, expected %d" % (len(lins), len(self.layers)))
        for l, lin in zip(self.layers, lins):
            l.refresh(lin.weight, lin.bias)
        self._flat_src = None

    def refresh_from_marl_state_dict(self, sd, head="act.action_out.fc_mean"):
        """Re-bind to (and re-cast from) another Actor / Critic state dict of the same architecture; in place."""
        if self.in_ln is not None:
            self.in_ln = (sd["base.feature_norm.weight"], sd["base.feature_norm.bias"], self.in_eps)
            self._refresh_in_ln()
        names = ["base.mlp.fc1"] + ["base.mlp.fc2.%d" % i for i in range(len(self.layers) - 2)]
        self._flat_src = None
        if any(l.fold is not None for l in self.layers):         # deferred LayerNorm: layer i folds the LayerNorm of layer i - 1
            prev = None
            for l, n in zip(self.layers[:-1], names):
                l.refresh(sd[n + ".0.weight"], sd[n + ".0.bias"], fold=prev)
                prev = (sd[n + ".2.weight"], sd[n + ".2.bias"], 1e-5)
            self.layers[-1].refresh(sd[head + ".weight"], sd[head + ".bias"], fold=prev)
            return
        for l, n in zip(self.layers[:-1], names):
            l.refresh(sd[n + ".0.weight"], sd[n + ".0.bias"], (sd[n + ".2.weight"], sd[n + ".2.bias"], l.eps))
        self.layers[-1].refresh(sd[head + ".weight"], sd[head + ".bias"])

    # -- constructors from the reference's modules ------------------------------------------------------
    @classmethod
    def from_sequential(cls, seq, device="cuda"):
        """PPO: nn.Sequential(Linear, ELU, Linear, ELU, ..., Linear) (module.py:25-49)."""
        mods = list(seq)
        layers = []
        i = 0
        while i < len(mods):
            lin = mods[i]
            if not isinstance(lin, torch.nn.Linear):
                raise L.MmbError("unsupported module in sequential: %r" % (lin,))
            act = i + 1 < len(mods) and isinstance(mods[i + 1], torch.nn.ELU)
            if i + 1 < len(mods) and not act:
                raise L.MmbError("only ELU activations are on the reference path (cfg/ppo/config.yaml:9): %r" % (mods[i + 1],))
            layers.append(_Layer(lin.weight, lin.bias, act, None, device))
            i += 2 if act else 1
        return cls(layers, None, device)

    @classmethod
    def from_marl_state_dict(cls, sd, head="act.action_out.fc_mean", device="cuda"):
        """MARL Actor / Critic checkpoint (runner.py:319-339 format): base.feature_norm, base.mlp.fc1, base.mlp.fc2.*,
        then `head` ('act.action_out.fc_mean' for the actor mean, 'v_out' for the critic value)."""
        def ln_of(prefix):      # live tensors (a module's state_dict() shares storage and version counters with its parameters)
            return sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5
        in_ln = ln_of("base.feature_norm") if "base.feature_norm.weight" in sd else None
        names = ["base.mlp.fc1"]
        while "base.mlp.fc2.%d.0.weight" % (len(names) - 1) in sd:
            names.append("base.mlp.fc2.%d" % (len(names) - 1))
        if _DEFER_LN and all(sd[n + ".0.weight"].shape[0] % 256 == 0 for n in names):
            # every hidden LayerNorm is applied by the layer that consumes it (deferred LayerNorm): free tile widths, one
            # pass over the accumulator, and no epilogue that needs all of tensor memory
            layers, prev = [], None
            for n in names:
                layers.append(_Layer(sd[n + ".0.weight"], sd[n + ".0.bias"], True, None, device, fold=prev, stats_out=True))
                prev = ln_of(n + ".2")
            layers.append(_Layer(sd[head + ".weight"], sd[head + ".bias"], False, None, device, fold=prev))
            return cls(layers, in_ln, device)
        layers = [_Layer(sd[n + ".0.weight"], sd[n + ".0.bias"], True, ln_of(n + ".2"), device) for n in names]
        layers.append(_Layer(sd[head + ".weight"], sd[head + ".bias"], False, None, device))
        return cls(layers, in_ln, device)

    # -- forward ---------------------------------------------------------------------------------------------
    def _buffers(self, M):
        if M not in self._bufs:
            Mpad = _round_up(M, 128)
            acts = [torch.zeros(Mpad, self.layers[0].Kpad, dtype=torch.bfloat16, device=self.device)]
            for l in self.layers[:-1]:
                acts.append(torch.zeros(Mpad, _round_up(l.N, 64), dtype=torch.bfloat16, device=self.device))
            self._bufs[M] = (Mpad, acts)
            # deferred LayerNorm: per producing layer, [Mpad][Npad / 64] (sum, sum of squares) partials
            self._stats[M] = [torch.zeros(Mpad, l.Npad // 64, 2, dtype=torch.float32, device=self.device) if l.stats_out else None
                              for l in self.layers]
        return self._bufs[M]

    def _fill_deferred(self, p, i, stats, parts_of):
        """The deferred-LayerNorm fields of layer i's launch parameters (stats: per layer, this network's partials)."""
        l = self.layers[i]
        if l.stats_out:
            p.ln_out_stats = L.ptr(stats[i])
        if l.fold is not None:
            p.ln_in_stats, p.ln_in_parts, p.ln_in_n = L.ptr(stats[i - 1]), parts_of(i - 1), self.layers[i - 1].N
            p.ln_in_eps, p.ln_c = l.fold[2], L.ptr(l.c)

    @staticmethod
    def _n_tile(l, Mpad, target_ctas=118):
        """Output-tile width: the widest of 256/128/64/32 that divides Npad and still gives about one CTA per SM
        (148 SMs; wide tiles reuse the A tile more, but a 64-CTA grid leaves half the GPU idle)."""
        if l.n_tile is not None:
            return l.n_tile
        if l.stats_out:                 # deferred LayerNorm: the producer's partial sums are per 64 columns of a 256-wide tile
            return 256
        cands = [t for t in (256, 128, 64, 32) if l.Npad % t == 0]
        for t in cands:
            if (Mpad // 128) * (l.Npad // t) >= target_ctas:
                return t
        return cands[-1]

    def forward_tf32(self, x, out=None):
        """The same forward with tf32 tensor-core operands (`mmb_mlp_chain`, operand_type 1): the fp32 observations, the LIVE
        fp32 nn.Linear weights and fp32 hidden activations go to the tensor core as they are (it reads 19 of the 32 bits) -
        no cast launch, no kernel-side weight copies, nothing to refresh after an optimiser step.  About half the rate of
        the bf16 path and ~20x closer to the reference's fp32 SGEMM (tests/test_gpu_mlp.py::test_tf32_chain).  Geometries
        the single-launch kernel does not take raise (there is no per-layer tf32 kernel)."""
        if x.device.type != "cuda":
            raise L.MmbError("FusedMLP runs on CUDA tensors only")
        if self.in_ln is not None or any(l.epilogue == 2 or l.fold is not None or l.stats_out for l in self.layers):
            raise L.MmbError("forward_tf32: Linear-ELU chains only")
        M = x.shape[0]
        x = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous()
        Mpad = _round_up(M, 128)
        bufs = self.__dict__.setdefault("_bufs32", {})
        if M not in bufs:
            bufs[M] = [torch.zeros(Mpad, l.N, dtype=torch.float32, device=self.device) for l in self.layers[:-1]]
        acts = bufs[M]
        if out is None:
            out = torch.empty(M, self.out_dim, dtype=torch.float32, device=self.device)
        nl = len(self.layers)
        arr = (L.MlpLayerParams * nl)()
        keep = []
        for i, l in enumerate(self.layers):
            w, b = l._src
            if not (w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and b.is_cuda and b.dtype == torch.float32):
                raise L.MmbError("forward_tf32 reads the live fp32 parameters: they must be contiguous CUDA tensors")
            if l.n_src != l.N:
                raise L.MmbError("forward_tf32: zero-padded heads are a bf16-path feature")
            p = arr[i]
            src = x if i == 0 else acts[i - 1]
            p.M, p.N, p.K, p.Mpad, p.Kpad, p.Npad, p.n_tile, p.epilogue = M, l.N, l.K, Mpad, src.stride(0), l.N, 0, l.epilogue
            p.x, p.w, p.bias = src.data_ptr(), w.data_ptr(), b.data_ptr()
            p.overlap_prev, p.operand_type = 0, 1
            dst = out if i == nl - 1 else acts[i]
            p.y, p.y_stride = dst.data_ptr(), dst.stride(0)
            keep += [w, b]
        L.check(L.lib().mmb_mlp_chain(arr, nl, 1, None, L.stream_ptr()), "mmb_mlp_chain (tf32)")
        return out

    def forward(self, x, out=None):
        """x fp32 [M, in_dim] on the device -> fp32 [M, out_dim]."""
        if x.device.type != "cuda":
            raise L.MmbError("FusedMLP runs on CUDA tensors only")
        if self.auto_refresh and self.stale():
            self.refresh()
        M = x.shape[0]
        x = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous()
        Mpad, acts = self._buffers(M)
        lib, st = L.lib(), L.stream_ptr()
        l0 = self.layers[0]
        use_ln = self.in_ln is not None
        if out is None:
            out = torch.empty(M, self.out_dim, dtype=torch.float32, device=self.device)
        nl = len(self.layers)
        # steady state of a rollout loop: same batch, same buffers -> the filled parameter array of the last such call
        # (keyed by the batch size only: the input and output pointers - a fresh observation tensor and a fresh result per step
        # in the reference's loops - are patched into the cached array)
        key = M
        hit = self.__dict__.setdefault("_chain_cache", {}).get(key)
        if hit is not None and _CHAIN_ENABLED and x.data_ptr() % 16 == 0:
            hit[0][nl - 1].y, hit[0][nl - 1].y_stride = out.data_ptr(), out.stride(0)
            hit[1][0] = x.data_ptr()
            L.check(lib.mmb_mlp_chain(hit[0], nl, 1, hit[1], st), "mmb_mlp_chain")
            return out
        arr = (L.MlpLayerParams * nl)()
        for i, l in enumerate(self.layers):
            p = arr[i]
            p.M, p.N, p.K, p.Mpad, p.Kpad, p.Npad, p.n_tile, p.epilogue = (M, l.N, l.K, Mpad, l.Kpad, l.Npad,
                                                                                 self._n_tile(l, Mpad), l.epilogue)
            p.x, p.w, p.bias = L.ptr(acts[i]), L.ptr(l.w), L.ptr(l.bias)
            if l.epilogue == 2:
                p.ln_gamma, p.ln_beta, p.ln_eps = L.ptr(l.gamma), L.ptr(l.beta), l.eps
            p.overlap_prev = 1        # the predecessor in the stream is ln_cast / the previous layer: never writes weights
            self._fill_deferred(p, i, self._stats[M], lambda j: self._stats[M][j].shape[1])
            if i == nl - 1:
                p.y, p.y_stride = L.ptr(out), out.stride(0)
            else:
                p.y, p.y_stride = L.ptr(acts[i + 1]), acts[i + 1].stride(0)
        # single launch with the input cast folded in (no LayerNorm in front, 16-byte addressable fp32 rows) ...
        folded = False
        if not use_ln and l0.K % 4 == 0 and x.data_ptr() % 16 == 0:
            folded = True
            xp = (C.c_void_p * 1)(x.data_ptr())
            if _chain_launch(self, arr, nl, 1, st, xp):
                if len(self._chain_cache) >= 8:
                    self._chain_cache.clear()
                self._chain_cache[key] = (arr, xp)
                return out
        L.check(lib.mmb_ln_cast(L.ptr(x), M, Mpad, l0.K, l0.Kpad, L.ptr(self.in_gamma) if use_ln else None,
                                L.ptr(self.in_beta) if use_ln else None, self.in_eps if use_ln else 0.0, int(use_ln),
                                L.ptr(acts[0]), st), "mmb_ln_cast")
        # ... or behind the cast / LayerNorm launch, or layer by layer
        if not folded and _chain_launch(self, arr, nl, 1, st):
            return out
        for i in range(nl):
            L.check(lib.mmb_mlp_layer(arr[i], st), "mmb_mlp_layer")
        return out

    __call__ = forward


class GroupedMLP:
    """Several `FusedMLP`s of identical architecture run as ONE launch per layer (grid z = network): the per-agent
    actors (or critics) of the MARL policies, which the reference evaluates one after the other (runner.py:205-217:
    2 x num_agents forwards of 4 small GEMMs + LayerNorms each, per env step).  `forward(xs)` takes the per-network
    inputs (list of [M, in] tensors, or one [G, M, in] tensor) and returns a [G, M, out] tensor."""

    def __init__(self, mlps):
        if not 1 <= len(mlps) <= L.MAX_GROUP:
            raise L.MmbError("GroupedMLP takes 1..%d networks" % L.MAX_GROUP)
        a = mlps[0]
        for m in mlps[1:]:
            if len(m.layers) != len(a.layers) or (m.in_ln is None) != (a.in_ln is None) or any(
                    (x.N, x.K, x.epilogue, x.fold is None, x.stats_out) != (y.N, y.K, y.epilogue, y.fold is None, y.stats_out)
                    for x, y in zip(m.layers, a.layers)):
                raise L.MmbError("GroupedMLP needs networks of identical architecture")
        self.mlps, self.G, self.device = list(mlps), len(mlps), a.device
        self.in_dim, self.out_dim = a.in_dim, a.out_dim
        self._bufs, self._stats = {}, {}
        self.auto_refresh = True

    def stale(self):
        return any(m.stale() for m in self.mlps)

    def refresh(self):
        for m in self.mlps:
            m.refresh()

    def _buffers(self, M):
        if M not in self._bufs:
            a = self.mlps[0]
            Mpad = _round_up(M, 128)
            acts = [torch.zeros(self.G, Mpad, a.layers[0].Kpad, dtype=torch.bfloat16, device=self.device)]
            for l in a.layers[:-1]:
                acts.append(torch.zeros(self.G, Mpad, _round_up(l.N, 64), dtype=torch.bfloat16, device=self.device))
            self._bufs[M] = (Mpad, acts)
            self._stats[M] = [torch.zeros(self.G, Mpad, l.Npad // 64, 2, dtype=torch.float32, device=self.device) if l.stats_out else None
                              for l in a.layers]
        return self._bufs[M]

    def forward(self, xs, out=None):
        import ctypes as C
        if self.auto_refresh and self.stale():
            self.refresh()
        if isinstance(xs, torch.Tensor):
            xs = [xs[g] for g in range(self.G)]
        shared_input = all(x is xs[0] for x in xs)           # e.g. actor and critic of one policy on the same observations
        xs = [x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous() for x in ([xs[0]] if shared_input else xs)]
        M = xs[0].shape[0]
        Mpad, acts = self._buffers(M)
        a0 = self.mlps[0]
        lib, st, G = L.lib(), L.stream_ptr(), self.G
        l0 = a0.layers[0]
        use_ln = a0.in_ln is not None
        shared_input = shared_input and not use_ln           # per-network input LayerNorms make the operands differ

        def cast_inputs():
            if shared_input:                                 # one cast serves every network: slot 0 of the first operand buffer
                L.check(lib.mmb_ln_cast(xs[0].data_ptr(), M, Mpad, l0.K, l0.Kpad, None, None, 0.0, 0, acts[0][0].data_ptr(), st), "mmb_ln_cast")
                return
            cc = self.__dict__.setdefault("_cast_cache", {})
            args = cc.get(M)
            if args is None:                                 # the static pointer arrays are built once per batch size
                vp = C.c_void_p * G
                args = cc[M] = (vp(), vp(*[m.in_gamma.data_ptr() for m in self.mlps]) if use_ln else None,
                                vp(*[m.in_beta.data_ptr() for m in self.mlps]) if use_ln else None, vp(*[acts[0][g].data_ptr() for g in range(G)]))
            xl = xs if len(xs) == G else xs * G
            for g in range(G):
                args[0][g] = xl[g].data_ptr()
            L.check(lib.mmb_ln_cast_group(args[0], G, M, Mpad, l0.K, l0.Kpad, args[1], args[2],
                                          a0.in_eps if use_ln else 0.0, int(use_ln), args[3], st), "mmb_ln_cast_group")
        if out is None:
            out = torch.empty(G, M, self.out_dim, dtype=torch.float32, device=self.device)
        nl = len(a0.layers)

        def fill(p, g, i):
            l = self.mlps[g].layers[i]
            p.M, p.N, p.K, p.Mpad, p.Kpad, p.Npad, p.n_tile, p.epilogue = (M, l.N, l.K, Mpad, l.Kpad, l.Npad,
                                                                             FusedMLP._n_tile(l, Mpad * G), l.epilogue)
            p.x, p.w, p.bias = acts[i][0 if (i == 0 and shared_input) else g].data_ptr(), l.w.data_ptr(), l.bias.data_ptr()
            if l.epilogue == 2:
                p.ln_gamma, p.ln_beta, p.ln_eps = l.gamma.data_ptr(), l.beta.data_ptr(), l.eps
            p.overlap_prev = 1
            st_g = [None if t is None else t[g] for t in self._stats[M]]
            self.mlps[g]._fill_deferred(p, i, st_g, lambda j: self._stats[M][j].shape[2])
            if i == nl - 1:
                p.y, p.y_stride = out[g].data_ptr(), out.stride(1)
            else:
                p.y, p.y_stride = acts[i + 1][g].data_ptr(), acts[i + 1].stride(1)

        casted = False
        # caches keyed by (batch size, shared input or not): the input / output pointers are patched in per call
        key = (M, len(xs))
        hit = self.__dict__.setdefault("_chain_cache", {}).get(key)
        if hit is not None and _CHAIN_ENABLED and all(x.data_ptr() % 16 == 0 for x in xs):
            xl = xs if len(xs) == G else xs * G
            for g in range(G):
                hit[1][g] = xl[g].data_ptr()
                hit[0][g * nl + nl - 1].y, hit[0][g * nl + nl - 1].y_stride = out[g].data_ptr(), out.stride(1)
            L.check(lib.mmb_mlp_chain(hit[0], nl, G, hit[1], st), "mmb_mlp_chain")
            return out
        if G <= 2 and self.__dict__.get("_chain_ok") is not False:       # network-major array: [G][layers]
            net_major = (L.MlpLayerParams * (G * nl))()
            for g in range(G):
                for i in range(nl):
                    fill(net_major[g * nl + i], g, i)
            xl = xs if len(xs) == G else xs * G
            if not use_ln and l0.K % 4 == 0 and all(x.data_ptr() % 16 == 0 for x in xl):
                xp = (C.c_void_p * G)(*[x.data_ptr() for x in xl])
                if _chain_launch(self, net_major, nl, G, st, xp):   # cast folded in
                    if len(self._chain_cache) >= 8:
                        self._chain_cache.clear()
                    self._chain_cache[key] = (net_major, xp)
                    return out
            else:
                cast_inputs()
                casted = True
                if _chain_launch(self, net_major, nl, G, st):
                    return out
        if not casted:
            cast_inputs()
        # steady state of a rollout loop (same batch, same buffers): the filled per-layer parameter arrays of the last such call
        lcache = self.__dict__.setdefault("_layer_cache", {})
        arrs = lcache.get(key)
        if arrs is None:
            arrs = []
            for i in range(nl):
                arr = (L.MlpLayerParams * G)()
                for g in range(G):
                    fill(arr[g], g, i)
                arrs.append(arr)
            if len(lcache) >= 8:
                lcache.clear()
            lcache[key] = arrs
        else:
            for g in range(G):
                arrs[-1][g].y, arrs[-1][g].y_stride = out[g].data_ptr(), out.stride(1)
        for arr in arrs:
            L.check(lib.mmb_mlp_layer_group(arr, G, st), "mmb_mlp_layer_group")
        return out

    __call__ = forward


def gaussian_act(mean, std, seed=0, step=0, noise=None, deterministic=False, per_dim=False, std_group_rows=0, sigma_src=None,
                 step_counter=None, out=None):
    """actions = mean + z * std and log-probs in one launch (`mmb_gaussian_act`).  mean [M, A] (row stride allowed), std [A] -
    or [groups, A] with std_group_rows = rows per group (a team's agent-major means with per-agent std rows).
    Returns (actions [M, A], logp): logp [M] summed over the action dims, or [M, A] with per_dim=True.  With `sigma_src`
    (same shape as std, fp32) a third output [M, A] holds that row broadcast over the rows - the `log_std.repeat(N, 1)` PPO's
    `act()` returns (module.py:87) - written by the same launch.  `step_counter`: int64 device tensor of `_lib.ACT_COUNTER_WORDS` words {step, 0, ...}; the launch
    takes the Philox step from it and advances it (a captured CUDA graph then draws fresh numbers on every replay).
    `out` = (actions, logp, sigma or None): contiguous fp32 tensors to write into (rollout-storage slots) instead of fresh ones."""
    M, A = mean.shape
    if mean.stride(1) != 1:
        mean = mean.contiguous()
    std = std.reshape(-1).float().contiguous()
    if std.numel() != (A if not std_group_rows else A * ((M + std_group_rows - 1) // std_group_rows)):
        raise L.MmbError("gaussian_act: std has %d elements for %d rows x %d dims (std_group_rows %d)" % (std.numel(), M, A, std_group_rows))
    if out is not None:
        actions, logp = out[0], out[1]
        want = {"actions": (actions, M * A), "logp": (logp, M * A if per_dim else M)}
        if sigma_src is not None:
            want["sigma"] = (out[2], M * A)
        for k, (t_, n_) in want.items():
            if t_.dtype != torch.float32 or not t_.is_contiguous() or t_.numel() != n_ or t_.device != mean.device:
                raise L.MmbError("gaussian_act: out %s must be a contiguous fp32 tensor of %d elements on the means' device" % (k, n_))
    else:
        actions = torch.empty(M, A, dtype=torch.float32, device=mean.device)
        logp = torch.empty((M, A) if per_dim else (M,), dtype=torch.float32, device=mean.device)
    p = L.GaussianActParams()
    p.num_rows, p.act_dim, p.deterministic, p.std_group_rows = M, A, int(bool(deterministic)), int(std_group_rows)
    p.mean, p.mean_stride, p.std = mean.data_ptr(), mean.stride(0), std.data_ptr()
    if noise is not None:
        noise = noise.float().contiguous()
        p.noise = noise.data_ptr()
    p.seed, p.step = int(seed) & 0xFFFFFFFFFFFFFFFF, int(step) & 0xFFFFFFFFFFFFFFFF
    if step_counter is not None:
        if (step_counter.dtype != torch.int64 or step_counter.numel() < L.ACT_COUNTER_WORDS or step_counter.device != mean.device or
                not step_counter.is_contiguous()):
            raise L.MmbError("gaussian_act: step_counter must be a contiguous int64 tensor of %d words {step, 0, ...} on the means' "
                             "device" % L.ACT_COUNTER_WORDS)
        p.step_counter = step_counter.data_ptr()
    p.actions = actions.data_ptr()
    if per_dim:
        p.logp_per_dim = logp.data_ptr()
    else:
        p.logp_sum = logp.data_ptr()
    sigma = None
    if sigma_src is not None:
        sigma_src = sigma_src.detach().reshape(-1)
        if sigma_src.dtype != torch.float32 or not sigma_src.is_contiguous() or sigma_src.device != mean.device:
            sigma_src = sigma_src.to(mean.device, torch.float32).contiguous()
        if sigma_src.numel() != std.numel():
            raise L.MmbError("gaussian_act: sigma_src has %d elements, std %d" % (sigma_src.numel(), std.numel()))
        sigma = out[2] if out is not None else torch.empty(M, A, dtype=torch.float32, device=mean.device)
        p.sigma_src, p.sigma_out = sigma_src.data_ptr(), sigma.data_ptr()
    L.check(L.lib().mmb_gaussian_act(p, L.stream_ptr()), "mmb_gaussian_act")
    return (actions, logp) if sigma is None else (actions, logp, sigma)


class PPOActorCriticForward:
    """Rollout-time interface of the reference's PPO `ActorCritic` (module.py:73-91): `act(observations, states)
    -> (actions, log_prob (N,), value (N,1), mean, log_std.repeat(N,1))` and `act_inference(observations) -> mean`,
    with both MLPs on the tcgen05 kernels.  `evaluate` (needs autograd) stays on the reference module.
    The reference's quirk is kept: `MultivariateNormal(mean, scale_tril=diag(exp(log_std)^2))`, i.e. the effective
    standard deviation is sigma^2 (module.py:76-77)."""

    def __init__(self, actor_critic, device="cuda"):
        self.asymmetric = bool(getattr(actor_critic, "asymmetric", False))
        self.actor = FusedMLP.from_sequential(actor_critic.actor, device)
        self.critic = FusedMLP.from_sequential(actor_critic.critic, device)
        self.log_std = actor_critic.log_std.detach().to(device)
        # actor and critic have the same hidden architecture (module.py:25-55); with the critic's 1-wide head zero-padded to
        # the actor's width the two networks run as ONE grouped launch per layer (grid z = network)
        self._pair = None
        a, c = self.actor, self.critic
        same_hidden = (not self.asymmetric and len(a.layers) == len(c.layers) and a.in_dim == c.in_dim and
                       all((x.N, x.K) == (y.N, y.K) for x, y in zip(a.layers[:-1], c.layers[:-1])) and
                       a.layers[-1].K == c.layers[-1].K and c.layers[-1].N <= a.layers[-1].N)
        if same_hidden:
            la, lc = a.layers[-1], c.layers[-1]
            head = [m for m in actor_critic.critic if isinstance(m, torch.nn.Linear)][-1]
            padded = FusedMLP(c.layers[:-1] + [_Layer(head.weight, head.bias, False, None, device, n_out=la.N)], None, device)
            self._pair = GroupedMLP([a, padded])
            self._value_cols = lc.N

    def refresh_from(self, actor_critic=None):
        """Re-cast the kernel-side weights from the module's CURRENT parameters, in place.  Not needed after an in-place
        `optimizer.step()` / `load_state_dict` on the module this object was built from (every forward checks the
        parameters' version counters and refreshes by itself); pass another `ActorCritic` of the same architecture to
        re-bind to it (e.g. a freshly constructed module after a restart)."""
        if actor_critic is None:
            for m in (self.actor, self.critic) + ((self._pair.mlps[1],) if self._pair is not None else ()):
                m.refresh()
            return
        self.actor.refresh_from_sequential(actor_critic.actor)
        self.critic.refresh_from_sequential(actor_critic.critic)
        if self._pair is not None:
            self._pair.mlps[1].refresh_from_sequential(actor_critic.critic)
        self.log_std = actor_critic.log_std.detach().to(self.actor.device)

    def _mean_value(self, observations, states):
        if self._pair is not None:
            out = self._pair([observations, observations])
            return out[0], out[1][:, :self._value_cols]
        return self.actor(observations), self.critic(states if self.asymmetric else observations)

    @torch.no_grad()
    def act(self, observations, states=None, noise=None, out=None):
        """`noise` [N, act]: standard normal draws to use (parity tests); default: in-kernel Philox keyed by (seed, call #).
        `out` = (actions [N, act], log_prob [N] or [N, 1], sigma [N, act]): write these three into the given tensors (slots of a
        `RolloutStorage`) instead of fresh ones."""
        mean, value = self._mean_value(observations, states)
        # MultivariateNormal with scale_tril = diag(exp(log_std)^2): sample + log_prob (sum over dims) in one launch
        self._calls = getattr(self, "_calls", 0) + 1
        self._refresh_scale()
        actions, log_prob, sigma = gaussian_act(mean, self._scale, seed=getattr(self, "seed", 0), step=self._calls, noise=noise,
                                                sigma_src=self.log_std, step_counter=self.__dict__.get("_step_counter_dev"), out=out)
        return actions, log_prob, value, mean, sigma      # sigma = log_std.repeat(N, 1), written by the same launch

    def _refresh_scale(self):
        """exp(log_std)^2, recomputed IN PLACE only when log_std has changed (its version counter: an optimiser step bumps
        it), so a CUDA graph that captured `act()` keeps reading the right tensor."""
        if self.__dict__.get("_scale_ver") != self.log_std._version:
            e = self.log_std.detach().float().exp()
            if self.__dict__.get("_scale") is None:
                self._scale = (e * e).contiguous()
            else:
                torch.mul(e, e, out=self._scale)
            self._scale_ver = self.log_std._version

    def use_device_step_counter(self):
        """From now on the sampling's Philox step lives on the device (`mmb_gaussian_act` reads and advances it,
        include/mmb.h `step_counter`): a captured CUDA graph of `act()` draws fresh numbers on every replay.  The sequence
        continues where the host-side call counter stood."""
        if self.__dict__.get("_step_counter_dev") is None:
            self._step_counter_dev = torch.zeros(L.ACT_COUNTER_WORDS, dtype=torch.int64, device=self.actor.device)
            self._step_counter_dev[0] = getattr(self, "_calls", 0) + 1
        return self._step_counter_dev

    def sync_parameters(self):
        """What every eager forward does by itself, for callers that replay a captured graph of `act()`: re-cast the
        kernel-side weights in place if the module's parameters have changed, and the sampling scale likewise."""
        for m in (self.actor, self.critic) + ((self._pair,) if self._pair is not None else ()):
            if m.stale():
                m.refresh()
        self._refresh_scale()

    @torch.no_grad()
    def act_inference(self, observations):
        return self.actor(observations)


class MarlPolicyForward:
    """Rollout-time forward of one MARL agent (mappo_policy.get_actions path: actor_critic.py:42-69,149-168):
    `get_actions(share_obs, obs, deterministic=False) -> (values (N,1), actions (N,A), action_log_probs (N,A))`.
    DiagGaussian semantics of agents/algorithms/utils/distributions.py:94-117: std = sigmoid(log_std / std_x_coef) *
    std_y_coef, per-dimension log-probs (not summed)."""

    def __init__(self, actor_sd, critic_sd, std_x_coef=1.0, std_y_coef=0.5, device="cuda"):
        self.actor = FusedMLP.from_marl_state_dict(actor_sd, "act.action_out.fc_mean", device)
        self.critic = FusedMLP.from_marl_state_dict(critic_sd, "v_out", device)
        self._std_coef = (std_x_coef, std_y_coef)
        self._log_std = actor_sd["act.action_out.log_std"]       # live: the std below is recomputed when it changes
        self._set_std()

    def _set_std(self):
        ls = self._log_std.detach().to(self.actor.device)
        self.std = torch.sigmoid(ls / self._std_coef[0]) * self._std_coef[1]
        self._std_ver = self._log_std._version

    def refresh_from(self, actor_sd=None, critic_sd=None):
        """As `PPOActorCriticForward.refresh_from`: in-place updates of the tensors the state dicts came from are picked up
        automatically; pass new state dicts (same architecture) to re-bind."""
        if actor_sd is not None:
            self.actor.refresh_from_marl_state_dict(actor_sd, "act.action_out.fc_mean")
            self._log_std = actor_sd["act.action_out.log_std"]
        else:
            self.actor.refresh()
        if critic_sd is not None:
            self.critic.refresh_from_marl_state_dict(critic_sd, "v_out")
        else:
            self.critic.refresh()
        self._set_std()

    @torch.no_grad()
    def get_actions(self, share_obs, obs, deterministic=False, noise=None):
        if self._log_std._version != self._std_ver:
            self._set_std()
        mean = self.actor(obs)
        self._calls = getattr(self, "_calls", 0) + 1
        actions, logp = gaussian_act(mean, self.std, seed=getattr(self, "seed", 0), step=self._calls, noise=noise,
                                     deterministic=deterministic, per_dim=True)
        return self.critic(share_obs), actions, logp

    @torch.no_grad()
    def get_values(self, share_obs):
        return self.critic(share_obs)


class MarlTeamForward:
    """All agents' actors and critics at once (the `collect` loop of agents/algorithms/marl/runner.py:197-227): two
    grouped forwards per env step instead of 2 x num_agents.  `get_actions(share_obs [N, S] or [A, N, S], obs [A, N, O]
    or [N, A, O], deterministic=False) -> (values [A, N, 1], actions [A, N, act], action_log_probs [A, N, act])` with the
    DiagGaussian semantics of `MarlPolicyForward`."""

    def __init__(self, actor_sds, critic_sds, std_x_coef=1.0, std_y_coef=0.5, device="cuda"):
        self.actors = GroupedMLP([FusedMLP.from_marl_state_dict(sd, "act.action_out.fc_mean", device) for sd in actor_sds])
        self.critics = GroupedMLP([FusedMLP.from_marl_state_dict(sd, "v_out", device) for sd in critic_sds])
        self._std_coef = (std_x_coef, std_y_coef)
        self._log_stds = [sd["act.action_out.log_std"] for sd in actor_sds]                                  # live
        self.A = len(actor_sds)
        self._set_std()

    def _set_std(self):
        log_std = torch.stack([t.detach().to(self.actors.device) for t in self._log_stds])                  # [A, act]
        self.std = (torch.sigmoid(log_std / self._std_coef[0]) * self._std_coef[1]).unsqueeze(1)             # [A, 1, act]
        self._std_ver = tuple(t._version for t in self._log_stds)

    def refresh_from(self, actor_sds=None, critic_sds=None):
        """In-place updates of the source tensors are picked up automatically at the next forward; pass new lists of state
        dicts (same architectures) to re-bind."""
        if actor_sds is not None:
            for m, sd in zip(self.actors.mlps, actor_sds):
                m.refresh_from_marl_state_dict(sd, "act.action_out.fc_mean")
            self._log_stds = [sd["act.action_out.log_std"] for sd in actor_sds]
        else:
            self.actors.refresh()
        if critic_sds is not None:
            for m, sd in zip(self.critics.mlps, critic_sds):
                m.refresh_from_marl_state_dict(sd, "v_out")
        else:
            self.critics.refresh()
        self._set_std()

    def _agent_major(self, x, width):
        if isinstance(x, (list, tuple)):                  # already one [N, .] tensor per agent (buffer slots)
            return list(x)
        if x.dim() == 2:                                  # one shared row per env: the same input for every critic
            return [x] * self.A
        if x.shape[0] != self.A:                          # (N, A, .) as MultiVecTaskPython returns it
            x = x.transpose(0, 1)
        return [x[a] for a in range(self.A)]

    @torch.no_grad()
    def get_actions(self, share_obs, obs, deterministic=False):
        if tuple(t._version for t in self._log_stds) != self._std_ver:
            self._set_std()
        mean = self.actors(self._agent_major(obs, self.actors.in_dim))                     # [A, N, act], contiguous
        # the whole team's sampling + per-dimension log-probs in ONE launch (a dozen elementwise torch launches before): the
        # agent-major means as [A * N] rows, agent a's std row for rows a * N .. a * N + N - 1
        A_, N_, D_ = mean.shape
        self._calls = getattr(self, "_calls", 0) + 1
        actions, logp = gaussian_act(mean.view(A_ * N_, D_), self.std, seed=getattr(self, "seed", 0), step=self._calls,
                                     deterministic=deterministic, per_dim=True, std_group_rows=N_)
        return self.critics(self._agent_major(share_obs, self.critics.in_dim)), actions.view(A_, N_, D_), logp.view(A_, N_, D_)

    @torch.no_grad()
    def get_values(self, share_obs):
        return self.critics(self._agent_major(share_obs, self.critics.in_dim))
