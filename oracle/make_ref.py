"""Recipe for oracle/_ref: the reference's own implementation of the path, BYTE-COMPILED from the sources where they lie
under /root/reference (no source file is copied; oracle/_ref is git-ignored build output, like a compiled .so would be).

TEST / MEASUREMENT INFRASTRUCTURE: `__graft_entry__.build()` runs it in the build container; the GPU box, which has no
/root/reference, imports the compiled modules through oracle/ref_shim.py for
  * `bench.py --impl reference` and the `cpu_baseline` leg (`kind: "reference"`: the reference's own
    `compute_ood_decision_on_results` / sklearn call timed on the box's host cores), and
  * tests/test_gpu_pipeline.py (the reference's unmodified evaluation driver run against this repo's classes).
Nothing under ood_in_object_detection_b200/ imports it.

    python oracle/make_ref.py            # -> oracle/_ref/**.pyb (+ the yaml defaults ultralytics reads at import time)

The compiled modules carry the suffix .pyb (the content is what py_compile writes into a .pyc): snapshots of the repo skip
*.pyc, so oracle/ref_shim.py registers a finder that loads <module>.pyb through importlib's SourcelessFileLoader.
"""
from __future__ import annotations

import os
import py_compile
import shutil
import sys

SRC = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
TOP_LEVEL = ("ood_utils.py", "cluster_utils.py", "custom_hyperparams.py", "constants.py", "ood_evaluation.py", "data_utils.py",
             "unknown_localization_utils.py", "visualization_utils.py", "log.py")
PACKAGES = ("ultralytics", "datasets_utils", "custom_datasets")
DATA_SUFFIXES = (".yaml",)                             # ultralytics/cfg/default.yaml etc. are read when the package is imported


def build(verbose: bool = False) -> int:
    if not os.path.isfile(os.path.join(SRC, "ood_utils.py")):
        return 0                                       # not in the build container: keep whatever travelled with the snapshot
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    n = 0
    jobs = [(os.path.join(SRC, f), f) for f in TOP_LEVEL if os.path.isfile(os.path.join(SRC, f))]
    for pkg in PACKAGES:
        for root, _, files in os.walk(os.path.join(SRC, pkg)):
            for f in files:
                rel = os.path.relpath(os.path.join(root, f), SRC)
                if f.endswith(".py") or f.endswith(DATA_SUFFIXES):
                    jobs.append((os.path.join(root, f), rel))
    for src, rel in jobs:
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if rel.endswith(".py"):
            try:                                       # sourceless import: <module>.pyb where the .py would be
                py_compile.compile(src, cfile=dst + "b", dfile=rel, doraise=True)
                n += 1
            except py_compile.PyCompileError as e:     # a reference file that does not even parse is not on the path
                if verbose:
                    print("skipped", rel, str(e).splitlines()[0])
        else:
            shutil.copyfile(src, dst)
    with open(os.path.join(OUT, "BUILD_INFO"), "w") as f:
        f.write(f"byte-compiled from {SRC} with python {sys.version.split()[0]}: {n} modules\n")
    return n


if __name__ == "__main__":
    print(f"oracle/_ref: {build(verbose=True)} modules compiled")
