"""CPU oracle for the OoD-scoring hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's OoD-scoring
path (`ood_utils.py`, `cluster_utils.py`,
`ultralytics/models/yolo/detect/predict.py:13-90` under /root/reference) and of
the third-party arithmetic that path calls (torchvision 0.26 `roi_align`,
scikit-learn 1.9 `normalize` / `pairwise_distances` / `KMeans`, scipy 1.18
`cdist`, numpy 2.3 `percentile(method='lower')`).

Who may import it: `tests/`, `__graft_entry__.smoke()`, and the `cpu_baseline`
/ `--impl reference` legs of `bench.py`.  Nothing under
`ood_in_object_detection_b200/` imports it; the product path fails loudly when
the CUDA library is missing instead of falling back to anything here.

Parity pin: the reference ships no test, golden vector or fixture for this path
(SURVEY.md §4, §8c), so the oracle is pinned against the *reference itself*:
`tests/golden/make_golden.py` imports the unmodified reference from
/root/reference through `oracle.ref_shim`, runs its own functions on seeded
synthetic inputs and freezes inputs+outputs under `tests/golden/*.npz`.
`tests/test_oracle_vs_golden.py` checks every oracle function against those
files (and against the live reference when /root/reference is present).

Modules
    ref_shim      import shim for the unmodified reference (build container only)
    roi_align     torchvision RoIAlign 1x1 adaptive + the per-stride extractor
    distance      sklearn normalize / pairwise l1,l2,cosine / min
    logits        MSP / Energy / ODIN / Sigmoid scorers, INDness maps
    decide        decide-on-results loops (incl. quirks Q1, Q2, Q4), fusion rules
    fit           cluster generation ('one', 'all', KMeans_<k>), scores, thresholds
    kmeans        sklearn KMeans (k-means++ init + Lloyd) restated in numpy
    silhouette    sklearn silhouette / Calinski-Harabasz scores and the reference's search of the number of clusters
    eul           ranking of unknown-object proposals against the clusters of every class
    cpu_path      loop-for-loop port of the reference's per-box CPU path (timing)
"""
