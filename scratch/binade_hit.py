"""How often does a (batch x strip) window of the fused kernel hold values of ONE f32 binade (same sign and exponent)?"""
import sys; sys.path.insert(0, ".")
import torch, torch.nn.functional as F
from fujishadergpu_b200 import kernels as k
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for seed in (20261017 + 2, 11):
    d = k.synth_dem((S, S), seed=seed)
    e = ((d.view(torch.int32) >> 23) & 0x1ff).to(torch.float32)[None, None]
    print("seed", seed, "z range", float(d.min()), float(d.max()))
    for (kh, kw) in ((37, 268), (49, 280), (97, 328), (32, 264), (32, 328), (16, 136), (8, 72)):
        mx = F.max_pool2d(e, (kh, kw), stride=(32, 264))
        mn = -F.max_pool2d(-e, (kh, kw), stride=(32, 264))
        hit = (mx == mn).float().mean().item()
        two = ((mx - mn) <= 1).float().mean().item()
        print(f"  window {kh}x{kw}: single binade {hit:.3f}, <=2 adjacent binades {two:.3f}")
