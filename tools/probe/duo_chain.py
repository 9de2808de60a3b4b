"""PPO actor + critic on the same observations: the dual-network chain (mlp_chain_duo_kernel, default) against the
side-by-side launch (MMB_MLP_DUO=0: grid z = network) - one process per mode (the switch is read once), graph-replayed."""
import json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, root)
if len(sys.argv) > 1:
    import ctypes as C
    import torch
    from massive_marl_benchmark_b200 import _lib as L
    from massive_marl_benchmark_b200 import mlp as mm
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(0)

    def net(i, hs, o):
        layers, d = [], i
        for h in hs:
            layers += [torch.nn.Linear(d, h), torch.nn.ELU()]
            d = h
        layers.append(torch.nn.Linear(d, o))
        return torch.nn.Sequential(*layers).to(dev)

    res = {"mode": sys.argv[1]}
    for M in (2048, 4096, 8192, 16384):
        actor, critic = net(388, [1024, 1024, 512], 80), net(388, [1024, 1024, 512], 80)
        x = torch.clamp(torch.randn(M, 388, device=dev) * 2, -5, 5)
        fa, fc = mm.FusedMLP.from_sequential(actor, dev), mm.FusedMLP.from_sequential(critic, dev)
        pair = mm.GroupedMLP([fa, fc])
        out = torch.empty(2, M, 80, device=dev)
        for _ in range(5):
            pair([x, x], out=out)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            pair([x, x], out=out)
            with torch.cuda.graph(gr, stream=side):
                for _ in range(20):
                    pair([x, x], out=out)
        torch.cuda.synchronize()
        for _ in range(3):
            gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        R = 20
        for _ in range(R):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / (R * 20) * 1e3
        flops = 2 * 2 * M * (388 * 1024 + 1024 * 1024 + 1024 * 512 + 512 * 80)
        res["M%d" % M] = {"us": round(us, 2), "tflops": round(flops / us / 1e6, 1)}
    st = (C.c_uint32 * 4)()
    L.lib().mmb_mlp_debug_status(st)
    res["debug_status"] = [hex(v) for v in st]
    print(json.dumps(res))
else:
    out = []
    modes = [("duo", {}), ("side_by_side", {"MMB_MLP_DUO": "0"})]
    for d in os.environ.get("DUO_DBG_MODES", "").split(","):
        if d:
            modes.append(("duo_dbg" + d, {"MMB_DUO_DBG": d}))
    for mode, env in modes:
        r = subprocess.run([sys.executable, __file__, mode], capture_output=True, text=True, timeout=300, env=dict(os.environ, **env))
        line = (r.stdout.strip().splitlines() or ["{}"])[-1]
        try:
            out.append(json.loads(line))
        except Exception:
            out.append({"mode": mode, "error": (r.stderr.strip().splitlines() or ["?"])[-1][:300]})
        print(out[-1], flush=True)
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(root, "gpurun_out", "duo_chain.json"), "w"), indent=1)
