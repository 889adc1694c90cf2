"""tcgen05 Lloyd step vs the FP32 step kernel: labels, counts, block partial sums; then timing at the C3 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ood_in_object_detection_b200 import kmeans, synth

dev = torch.device("cuda:0")
be_tc, be_fp = kmeans.CudaBackend(dev), kmeans.CudaBackend(dev)
be_tc.tensor_core, be_fp.tensor_core = True, False

def run(be, x, sizes, k, cent, update=1):
    table, shard, local_off = kmeans.build_blocks(sizes, 1, 0, dev)
    n_seg = len(sizes)
    seg_k = torch.tensor([min(k, n) for n in sizes], dtype=torch.int32, device=dev)
    labels = torch.full((x.shape[0],), -1, dtype=torch.int32, device=dev)
    chg = torch.zeros(n_seg, dtype=torch.int32, device=dev)
    ps, pc = be.step(x, k, seg_k, cent, table, None, labels, chg, update)
    torch.cuda.synchronize()
    return labels, chg, (ps.clone() if ps is not None else None), (pc.clone() if pc is not None else None), table

ok = True
for dim, k, spec in ((576, 16, ((1, 5000), (2, 777), (3, 33), (4, 0), (5, 1500))), (128, 5, ((6, 2000), (7, 1029))),
                     (256, 16, ((8, 4100),)), (640, 12, ((9, 1300), (10, 64)))):
    segs = [synth.blob_vectors(s, n, dim, k, 3.0)[0] if n else np.zeros((0, dim), np.float32) for s, n in spec]
    sizes = [len(s) for s in segs]
    x = torch.from_numpy(np.concatenate(segs)).to(dev)
    g = torch.Generator(device="cpu").manual_seed(0)
    cent = torch.stack([torch.from_numpy(s[torch.randint(0, max(len(s), 1), (k,), generator=g).numpy()] if len(s) else np.zeros((k, dim), np.float32)) for s in segs]).to(dev)
    cent = cent + 0.01 * torch.randn(cent.shape, device=dev)
    la, ca, psa, pca, table = run(be_fp, x, sizes, k, cent)
    lb, cb, psb, pcb, _ = run(be_tc, x, sizes, k, cent)
    same = (la == lb).float().mean().item()
    print(f"D={dim} K={k}: labels equal {same:.6f}, changed fp {ca.tolist()} tc {cb.tolist()}")
    if same < 1.0:
        # a differing label must be a near tie in float32: compare distances in float64
        idx = torch.nonzero(la != lb).flatten()[:5]
        off = np.concatenate([[0], np.cumsum(sizes)])
        for i in idx.tolist():
            gseg = int(np.searchsorted(off, i, side="right") - 1)
            c64 = cent[gseg].double(); x64 = x[i].double()
            d = ((c64 - x64) ** 2).sum(1)
            print("   row", i, "fp", int(la[i]), "tc", int(lb[i]), "d_fp", float(d[la[i]]), "d_tc", float(d[lb[i]]))
        ok = ok and same > 0.999
    eq = la == lb
    if eq.all():
        print("   counts equal", torch.equal(pca, pcb), " psums max abs diff", (psa - psb).abs().max().item(),
              " bit-equal", torch.equal(psa, psb))
        ok = ok and torch.equal(pca, pcb) and torch.equal(psa, psb)
    l0, _, _, _, _ = run(be_tc, x, sizes, k, cent, update=0)
    print("   update=0 labels equal", torch.equal(l0, lb))
    ok = ok and torch.equal(l0, lb)
print("CHECK", "PASS" if ok else "FAIL")

# timing at the C3 shape
n_seg, per, dim, k = 20, int(sys.argv[1]) if len(sys.argv) > 1 else 200000, 576, 16
x = torch.randn(n_seg * per, dim, device=dev) * 0.05
sizes = [per] * n_seg
cent = torch.randn(n_seg, k, dim, device=dev) * 0.05
table, _, _ = kmeans.build_blocks(sizes, 1, 0, dev)
seg_k = torch.full((n_seg,), k, dtype=torch.int32, device=dev)
labels = torch.full((x.shape[0],), -1, dtype=torch.int32, device=dev)
chg = torch.zeros(n_seg, dtype=torch.int32, device=dev)
for name, be in (("tcgen05", be_tc), ("fp32", be_fp)):
    for upd in (1, 0):
        be.step(x, k, seg_k, cent, table, None, labels, chg, upd); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            be.step(x, k, seg_k, cent, table, None, labels, chg, upd)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"{name} update={upd}: {ms:.3f} ms  {x.numel() * 4 / ms / 1e6:.0f} GB/s")
