import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ood_in_object_detection_b200 import kmeans
dev = torch.device("cuda:0")
be = kmeans.CudaBackend(dev)
n_seg, per, dim, k = 20, int(sys.argv[1]) if len(sys.argv) > 1 else 200000, 576, 16
upd = int(sys.argv[2]) if len(sys.argv) > 2 else 1
x = torch.randn(n_seg * per, dim, device=dev) * 0.05
sizes = [per] * n_seg
cent = torch.randn(n_seg, k, dim, device=dev) * 0.05
table, _, _ = kmeans.build_blocks(sizes, 1, 0, dev)
seg_k = torch.full((n_seg,), k, dtype=torch.int32, device=dev)
labels = torch.full((x.shape[0],), -1, dtype=torch.int32, device=dev)
chg = torch.zeros(n_seg, dtype=torch.int32, device=dev)
for _ in range(2):
    be.step(x, k, seg_k, cent, table, None, labels, chg, upd)
torch.cuda.synchronize()
print("ok")
