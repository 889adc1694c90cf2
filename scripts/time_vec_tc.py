import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ood_in_object_detection_b200 import ops
dev = torch.device("cuda:0")
n_seg, per, dim, k = 20, 200000, 576, int(sys.argv[1]) if len(sys.argv) > 1 else 16
x = torch.randn(n_seg * per, dim, device=dev); x /= x.norm(dim=1, keepdim=True)
cent = torch.randn(n_seg * k, dim, device=dev) * 0.05
off = [g * per for g in range(n_seg + 1)]
crow = [g * k for g in range(n_seg)]
def t(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b)/reps:.3f} ms")
t("vec_score fp32 l2", lambda: ops.vec_score(x, off, cent, None, crow, [k] * n_seg, 0b010, normalize=False))
t("vec_score_tc l2  ", lambda: ops.vec_score_tc(x, off, cent, crow, [k] * n_seg, "l2"))
