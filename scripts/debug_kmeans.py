import numpy as np, torch, sys
sys.path.insert(0, '.')
from ood_in_object_detection_b200 import kmeans, ops
from oracle import kmeans as ok, distance
from tests.helpers import unpack_nested
g = np.load('tests/golden/golden_kmeans.npz')
acts = unpack_nested(g, "fit_acts", 3)
a = acts[1][2].reshape(len(acts[1][2]), -1)
xn = distance.normalize_rows(a)
X = np.array(xn, dtype=np.float32, copy=True)
tol_abs = np.mean(np.var(X, axis=0)) * 1e-4
mean = X.mean(axis=0); X -= mean
centers, _ = ok.kmeans_plusplus(X, 5, np.random.RandomState(10))
print("tol_abs", tol_abs)
labels_old = None
for i in range(6):
    labels, cnew, shift = ok.lloyd_iter(X, centers)
    r = kmeans.kmeans_fit_predict_single(torch.from_numpy(xn).cuda(), [len(a)], 5, max_iter=i, tol=0.0) if i > 0 else None
    if r is not None:
        gl = r.labels.cpu().numpy()
        print(f"E-step {i}: oracle-vs-gpu label mismatches {int((gl != labels).sum())}; gpu centres vs oracle centres max diff {np.abs(r.centers[0].cpu().numpy() - (centers + mean)).max():.3e}")
    print(f"iter {i+1}: oracle changed {None if labels_old is None else int((labels != labels_old).sum())} shift^2 {(shift**2).sum():.4e}")
    centers = cnew; labels_old = labels
r = kmeans.kmeans_fit_predict_single(torch.from_numpy(xn).cuda(), [len(a)], 5)
print(r.n_iter, r.strict, r.seconds)
r0 = kmeans.kmeans_fit_predict_single(torch.from_numpy(xn).cuda(), [len(a)], 5, max_iter=0, tol=0.0)
c0, _ = ok.kmeans_plusplus(X, 5, np.random.RandomState(10))
gc = r0.centers[0].cpu().numpy() - mean
print("init centres diff per centre", np.abs(gc - c0).max(axis=1))
# which rows
for j in range(5):
    d = np.abs(X - gc[j]).max(axis=1); print("gpu seed", j, "row", int(d.argmin()), d.min(), "oracle row", int(np.abs(X - c0[j]).max(axis=1).argmin()))
l0, _, _ = ok.lloyd_iter(X, c0, update_centers=False)
print("E-step 0 mismatches", int((r0.labels.cpu().numpy() != l0).sum()))
