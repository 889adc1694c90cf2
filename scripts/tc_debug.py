import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ood_in_object_detection_b200 import kmeans
dev = torch.device("cuda:0")
be = kmeans.CudaBackend(dev)
n_seg, per, dim, k = 20, 200000, 576, 16
x = torch.randn(n_seg * per, dim, device=dev) * 0.05
sizes = [per] * n_seg
cent = torch.randn(n_seg, k, dim, device=dev) * 0.05
table, _, _ = kmeans.build_blocks(sizes, 1, 0, dev)
seg_k = torch.full((n_seg,), k, dtype=torch.int32, device=dev)
labels = torch.full((x.shape[0],), -1, dtype=torch.int32, device=dev)
chg = torch.zeros(n_seg, dtype=torch.int32, device=dev)
import itertools
for (xs, ls), dbg in itertools.product(((5, 2),), (0, 3, 4)):
    os.environ["OODB200_TC_DEBUG"] = str(dbg)
    os.environ["OODB200_TC_XS"], os.environ["OODB200_TC_LS"] = str(xs), str(ls)
    be.step(x, k, seg_k, cent, table, None, labels, chg, 1); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        be.step(x, k, seg_k, cent, table, None, labels, chg, 1)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"xs={xs} ls={ls} debug={dbg:2d} (1 nosplit, 2 nomma, 4 noacc): {ms:.3f} ms  {x.numel() * 4 / ms / 1e6:.0f} GB/s")
