"""Time K2 standalone (oodb200_vec_score_f32) on fit-sized inputs; OODB200_VEC_FAST=0 selects the one-row-per-warp kernel.
Usage: python scripts/time_vec_score.py [n_rows] ; prints ms per launch and % of HBM peak (4*D bytes per row)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ood_in_object_detection_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
dev = torch.device("cuda", 0)
out = {"fast": os.environ.get("OODB200_VEC_FAST", "1"), "rows": n}
for dim, k, nseg in ((576, 16, 20), (640, 64, 20), (128, 10, 20), (32, 10, 20)):
    g = torch.Generator(device=dev); g.manual_seed(1)
    x = torch.randn((n, dim), device=dev, generator=g)
    cent = torch.randn((nseg * k, dim), device=dev, generator=g)
    unit = cent / cent.norm(dim=1, keepdim=True)
    off = [i * (n // nseg) for i in range(nseg)] + [n]
    crow = [i * k for i in range(nseg)]
    for name, slot in (("l1", 0), ("l2", 1), ("cosine", 2)):
        for _ in range(2):
            d, a = ops.vec_score(x, off, cent, unit, crow, [k] * nseg, 1 << slot, normalize=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            d, a = ops.vec_score(x, off, cent, unit, crow, [k] * nseg, 1 << slot, normalize=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        # float64 check on a sample of rows
        idx = torch.randint(0, n, (2048,), device=dev, generator=g)
        xs = x[idx].double(); xs = xs / xs.norm(dim=1, keepdim=True)
        seg = torch.clamp(idx // (n // nseg), max=nseg - 1)
        c = cent.double().view(nseg, k, dim)[seg]
        if name == "l1": ref = (xs[:, None, :] - c).abs().sum(-1)
        elif name == "l2": ref = (xs[:, None, :] - c).pow(2).sum(-1).sqrt()
        else: ref = 1 - (xs[:, None, :] * (c / c.norm(dim=-1, keepdim=True))).sum(-1)
        rd, ra = ref.min(1)
        err = float(((d[slot][idx].double() - rd).abs() / rd.abs().clamp_min(1e-6)).max())
        agree = float((a[slot][idx].long() == ra).float().mean())
        out[f"D{dim}_K{k}_{name}"] = {"ms": round(ms, 3), "GBps": round(4 * dim * n / ms / 1e6, 1), "max_rel_err": err, "argmin_agree": agree}
    del x
print(json.dumps(out))
