"""GPU-box probe (round 1): facts about torch-CUDA eager numerics the kernels must copy.

Answers SURVEY.md section 7 open questions 3/4: reduction association for sum(-1) over 3/8/80
elements, norm over 3, the 1x3 . 3x1 bmm, torch.cross, remainder, scalar division; and whether
libdevice (nvcc 12.9) sinf/cosf/atanf/asinf/atan2f/fmodf match torch's CUDA kernels bit for bit.
Writes gpurun_out/probe_torch_cuda.json.  Not part of the product.
"""
import ctypes
import json
import os
import time

import numpy as np
import torch

out = {}
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(7)
N = 1 << 20


def frac_equal(a, b):
    return float((a.view(torch.int32) == b.view(torch.int32)).float().mean().item())


# ---- sum(-1) over 3 elements ------------------------------------------------------------
x = torch.randn(N, 3, generator=g).to(dev)
s = x.sum(-1)
c = {
    "(a0+a1)+a2": (x[:, 0] + x[:, 1]) + x[:, 2],
    "(a0+a2)+a1": (x[:, 0] + x[:, 2]) + x[:, 1],
    "a0+(a1+a2)": x[:, 0] + (x[:, 1] + x[:, 2]),
}
out["sum3_contig"] = {k: frac_equal(s, v) for k, v in c.items()}
# same, as in compute_ingenuity_reward: square(goal - obs[:, :3]).sum(-1) with obs (N,13)
obs = torch.randn(N, 13, generator=g).to(dev)
goal = torch.tensor([4.0, 2.0, 1.0], device=dev).repeat(N, 1)
sq = torch.square(goal - obs[:, :3])
s = sq.sum(-1)
c = {
    "(a0+a1)+a2": (sq[:, 0] + sq[:, 1]) + sq[:, 2],
    "(a0+a2)+a1": (sq[:, 0] + sq[:, 2]) + sq[:, 1],
    "a0+(a1+a2)": sq[:, 0] + (sq[:, 1] + sq[:, 2]),
}
out["sum3_ingenuity_expr"] = {k: frac_equal(s, v) for k, v in c.items()}
out["sum3_small_N64"] = {}
for n_small in (4, 64, 1000, 4096):
    xs = sq[:n_small].contiguous()
    ss = xs.sum(-1)
    out["sum3_small_N64"][str(n_small)] = {
        "(a0+a1)+a2": frac_equal(ss, (xs[:, 0] + xs[:, 1]) + xs[:, 2]),
        "(a0+a2)+a1": frac_equal(ss, (xs[:, 0] + xs[:, 2]) + xs[:, 1]),
    }

# ---- sum(-1) over 8 (abs(a*b)) ---------------------------------------------------------
a = torch.rand(N, 80, generator=g).to(dev) * 2 - 1
o = torch.randn(N, 38, generator=g).to(dev)
e = torch.abs(a[:, 0:8] * o[:, 22:30])
s = torch.sum(e, dim=-1)
seq = e[:, 0]
for j in range(1, 8):
    seq = seq + e[:, j]
pair = ((e[:, 0] + e[:, 1]) + (e[:, 2] + e[:, 3])) + ((e[:, 4] + e[:, 5]) + (e[:, 6] + e[:, 7]))
# candidates: 4 accumulators (vt0=4) combined: (e0+e4),(e1+e5),(e2+e6),(e3+e7) then ((v0+v1)+v2)+v3
v = [e[:, i] + e[:, i + 4] for i in range(4)]
acc4 = ((v[0] + v[1]) + v[2]) + v[3]
# block.x = 8 lanes then shuffle tree: offsets 4,2,1
t = [e[:, i] + e[:, i + 4] for i in range(4)]
t2 = [t[0] + t[2], t[1] + t[3]]
tree = t2[0] + t2[1]
# block.x=4 threads each strided 2 elems (x, x+4) then tree
out["sum8_abs"] = {"sequential": frac_equal(s, seq), "pairwise": frac_equal(s, pair),
                   "acc4": frac_equal(s, acc4), "shfl_tree8": frac_equal(s, tree)}
out["sum8_abs_maxrel"] = float(((s - seq).abs() / seq.abs().clamp(min=1e-30)).max().item())

# ---- sum over 80 of a**2 -------------------------------------------------------------------
p = a ** 2
s = torch.sum(p, dim=-1)
seq = p[:, 0]
for j in range(1, 80):
    seq = seq + p[:, j]
out["sum80_sq"] = {"sequential": frac_equal(s, seq),
                   "maxrel_vs_seq": float(((s - seq).abs() / seq).max().item()),
                   "pow2_is_mul": frac_equal(p, a * a)}
s64 = torch.sum(p.double(), dim=-1)
out["sum80_sq"]["maxrel_vs_f64"] = float(((s.double() - s64).abs() / s64).max().item())

# ---- norm over 3 ---------------------------------------------------------------------------
x = torch.randn(N, 3, generator=g).to(dev)
nr = x.norm(p=2, dim=-1)
c1 = torch.sqrt((x[:, 0] * x[:, 0] + x[:, 1] * x[:, 1]) + x[:, 2] * x[:, 2])
x64 = x.double()
c2 = torch.sqrt((x64 ** 2).sum(-1)).float()
out["norm3"] = {"sqrt((x0^2+x1^2)+x2^2)": frac_equal(nr, c1), "f64_rounded": frac_equal(nr, c2)}
# z = 0 case (to_target has z zeroed)
x2 = x.clone(); x2[:, 2] = 0
nr = x2.norm(p=2, dim=-1)
out["norm3_z0"] = {"sqrt(x0^2+x1^2)": frac_equal(nr, torch.sqrt(x2[:, 0] * x2[:, 0] + x2[:, 1] * x2[:, 1])),
                   "f64_rounded": frac_equal(nr, torch.sqrt((x2.double() ** 2).sum(-1)).float()),
                   "fma(x1,x1,x0*x0)": frac_equal(nr, torch.sqrt(torch.addcmul(x2[:, 0] * x2[:, 0], x2[:, 1], x2[:, 1])))}

# ---- bmm 1x3 . 3x1 ---------------------------------------------------------------------------
u = torch.randn(N, 3, generator=g).to(dev)
w = torch.randn(N, 3, generator=g).to(dev)
d = torch.bmm(u.view(N, 1, 3), w.view(N, 3, 1)).view(N)
plain = (u[:, 0] * w[:, 0] + u[:, 1] * w[:, 1]) + u[:, 2] * w[:, 2]
fma_chain = torch.addcmul(torch.addcmul(u[:, 0] * w[:, 0], u[:, 1], w[:, 1]), u[:, 2], w[:, 2])
d64 = (u.double() * w.double()).sum(-1).float()
fma_rev = torch.addcmul(torch.addcmul(u[:, 2] * w[:, 2], u[:, 1], w[:, 1]), u[:, 0], w[:, 0])
out["bmm3"] = {"plain_l2r": frac_equal(d, plain), "fma_chain_l2r": frac_equal(d, fma_chain),
               "fma_chain_r2l": frac_equal(d, fma_rev), "f64_rounded": frac_equal(d, d64),
               "tf32_allowed": bool(torch.backends.cuda.matmul.allow_tf32)}
out["bmm3_small"] = {}
for n_small in (64, 4096, 40960):
    ds = torch.bmm(u[:n_small].reshape(n_small, 1, 3), w[:n_small].reshape(n_small, 3, 1)).view(n_small)
    out["bmm3_small"][str(n_small)] = {"plain_l2r": frac_equal(ds, plain[:n_small]),
                                       "fma_chain_l2r": frac_equal(ds, fma_chain[:n_small])}

# ---- cross ---------------------------------------------------------------------------------
cr = torch.cross(u, w, dim=-1)
p0 = u[:, 1] * w[:, 2] - u[:, 2] * w[:, 1]
f0 = torch.addcmul(-(u[:, 2] * w[:, 1]), u[:, 1], w[:, 2])
f0b = torch.addcmul(u[:, 1] * w[:, 2], -u[:, 2], w[:, 1])
out["cross_x"] = {"plain": frac_equal(cr[:, 0].contiguous(), p0), "fma(a1,b2,-(a2*b1))": frac_equal(cr[:, 0].contiguous(), f0),
                  "fma(-a2,b1,a1*b2)": frac_equal(cr[:, 0].contiguous(), f0b)}

# ---- remainder, scalar division ----------------------------------------------------------------
ang = (torch.rand(N, generator=g).to(dev) * 2 - 1) * 3.1415927
r = ang % (2 * np.pi)
two_pi = torch.tensor(2 * np.pi, dtype=torch.float32, device=dev)
fm = torch.fmod(ang, two_pi)
cand = torch.where((fm != 0) & (fm < 0), fm + two_pi, fm)
out["remainder_2pi"] = {"fmod_then_add": frac_equal(r, cand)}
xx = torch.randn(N, generator=g).to(dev)
dt = 0.0166
q = xx / dt
out["div_scalar"] = {"true_div": frac_equal(q, xx / torch.tensor(dt, dtype=torch.float32, device=dev)),
                     "mul_recip_f32": frac_equal(q, xx * (torch.tensor(1.0, dtype=torch.float32, device=dev) / torch.tensor(dt, dtype=torch.float32, device=dev)))}

# ---- libdevice vs torch transcendental kernels ---------------------------------------------------
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libprobe.so"))
lib.run_unary.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_void_p]
lib.run_binary.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int, ctypes.c_void_p]
stream = torch.cuda.current_stream().cuda_stream
xs = (torch.rand(N, generator=g).to(dev) * 2 - 1)
res = {}
for name, scale in (("unit", 1.0), ("pi", 3.1415927), ("wide", 50.0)):
    xi = (xs * scale).contiguous()
    o = [torch.empty_like(xi) for _ in range(4)]
    rc = lib.run_unary(xi.data_ptr(), o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(), N, stream)
    torch.cuda.synchronize()
    res[name] = {"rc": rc, "sin": frac_equal(o[0], torch.sin(xi)), "cos": frac_equal(o[1], torch.cos(xi)),
                 "atan": frac_equal(o[2], torch.atan(xi))}
    if name == "unit":
        res[name]["asin"] = frac_equal(o[3], torch.asin(xi))
aa = torch.randn(N, generator=g).to(dev)
bb = torch.randn(N, generator=g).to(dev)
o = [torch.empty_like(aa) for _ in range(4)]
rc = lib.run_binary(aa.data_ptr(), bb.data_ptr(), o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(), N, stream)
torch.cuda.synchronize()
res["binary"] = {"rc": rc, "atan2": frac_equal(o[0], torch.atan2(aa, bb)), "fmod": frac_equal(o[1], torch.fmod(aa, bb)),
                 "div_rn": frac_equal(o[2], aa / bb), "sqrt_rn": frac_equal(o[3], torch.sqrt(aa.abs()))}
out["libdevice_vs_torch"] = res

# ---- launch latency / sync costs ------------------------------------------------------------------
z = torch.zeros(1024, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000):
    z.add_(1.0)
torch.cuda.synchronize()
out["torch_tiny_kernel_us"] = (time.perf_counter() - t0) / 2000 * 1e6
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(2000):
    z.add_(1.0)
ev1.record(); torch.cuda.synchronize()
out["torch_tiny_kernel_dev_us"] = ev0.elapsed_time(ev1) / 2000 * 1e3
out["host"] = {"cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads()}
out["gpu"] = torch.cuda.get_device_name(0)
out["nccl"] = ".".join(map(str, torch.cuda.nccl.version()))

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/probe_torch_cuda.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
