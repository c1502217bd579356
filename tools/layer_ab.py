"""Per-layer time of one forward under two option sets, same process (python tools/layer_ab.py)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nind_denoise_b200 as nb

dev = torch.device("cuda:0")
for cs, batch in ((248, 168), (504, 39)):
    x = torch.rand((batch, 3, cs, cs), device=dev)
    tabs = {}
    for name, opts in (("deep256", {"n_tile_deep": 256}), ("deep128", {"n_tile_deep": 128})):
        torch.manual_seed(0)
        m = nb.UtNet().to(dev).eval()
        for k, v in opts.items():
            m.set_option(k, v)
        for _ in range(3):
            m(x)
        acc = {}
        for rep in range(5):
            for n, ms, fl in m.layer_times(x):
                acc.setdefault(n, []).append(ms)
        tabs[name] = {n: sorted(v)[len(v) // 2] for n, v in acc.items()}
        del m
        torch.cuda.empty_cache()
    print(f"cs {cs} batch {batch}")
    tot = [0, 0, 0]
    for n in tabs["deep256"]:
        a, b = tabs["deep256"][n], tabs["deep128"][n]
        tot[0] += a; tot[1] += b; tot[2] += min(a, b)
        print(f"  {n:18s} 256: {a:7.3f}  128: {b:7.3f}  {'<-- 128' if b < a * 0.99 else ''}")
    print(f"  total 256 {tot[0]:.3f}  128 {tot[1]:.3f}  best-of {tot[2]:.3f}")
