"""Producer side of the hand-off: ood_in_object_detection_b200.postprocess.postprocess against
the reference's `DetectionPredictor.postprocess` (ultralytics/models/yolo/detect/predict.py:117-363) -- frozen outputs
(tests/golden/golden_postprocess.npz, every extraction mode, sigmoid and raw-logit heads, an image without detections) and,
where the reference's code is present (oracle/_ref on the GPU box), the reference itself on the same CUDA tensors.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import fake_predictor, postprocess_inputs  # noqa: E402

pytestmark = pytest.mark.gpu
CASES = {"fs": ("ftmaps_and_strides", False, 0.25), "fs_hi": ("ftmaps_and_strides", False, 0.97),
         "pos": ("ftmaps_and_strides_exact_pos", False, 0.25), "all": ("all_ftmaps", False, 0.25),
         "roi": ("roi_aligned_ftmaps", False, 0.25), "lg": ("logits", False, 0.25), "lg_raw": ("logits", True, 0.25),
         "lg_hi": ("logits", True, 0.97)}


@pytest.fixture(scope="module")
def inputs():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    pred, logits, maps = postprocess_inputs()
    dev = torch.device("cuda", 0)
    p = torch.from_numpy(pred).to(dev)
    raw = torch.cat([p[:, :4], torch.from_numpy(logits).to(dev)], 1)
    return dict(p=p, raw=raw, maps=[torch.from_numpy(m).to(dev) for m in maps], img=torch.zeros((4, 3, 320, 320), device=dev))


def _run(fn, inputs, mode, before, conf):
    head = (inputs["raw"] if before else inputs["p"]).clone()
    return fn(fake_predictor(mode, before, conf, device="cuda:0"), ((head,), None if mode == "logits" else inputs["maps"]),
              inputs["img"], inputs["img"])


def _check_against(res, tag, mode, want, inputs):
    """want: dict with n, boxes, extra... in the golden's layout."""
    n = np.array([len(r.boxes) for r in res])
    assert np.array_equal(n, want[f"{tag}_n"]), (tag, n)
    boxes = np.concatenate([r.boxes.data.cpu().numpy().reshape(-1, 6) for r in res])
    np.testing.assert_allclose(boxes[:, :4], want[f"{tag}_boxes"][:, :4], rtol=0, atol=1e-4)      # xywh -> xyxy in float32
    np.testing.assert_allclose(boxes[:, 4], want[f"{tag}_boxes"][:, 4], rtol=0, atol=1e-7)
    assert np.array_equal(boxes[:, 5], want[f"{tag}_boxes"][:, 5])
    assert tuple(res[0].orig_img.shape) == tuple(want[f"{tag}_shape"])
    assert [r.path for r in res] == [f"im{i}.jpg" for i in range(4)]
    if mode == "logits":
        extra = np.concatenate([r.extra_item.cpu().numpy().reshape(len(r.boxes), 20) for r in res])
        assert np.array_equal(extra, want[f"{tag}_extra"])                                         # gathered rows: bit-exact
    elif mode in ("ftmaps_and_strides", "ftmaps_and_strides_exact_pos"):
        extra = np.concatenate([r.extra_item[1].cpu().numpy().reshape(-1) for r in res]).astype(np.float64)
        assert np.array_equal(extra, want[f"{tag}_extra"])
        for i, r in enumerate(res):                                                                # views of the batched maps: no copy
            for s in range(3):
                assert r.extra_item[0][s].data_ptr() == inputs["maps"][s][i].data_ptr() and r.extra_item[0][s].shape == inputs["maps"][s][i].shape
    elif mode == "roi_aligned_ftmaps":
        for s in range(3):
            assert np.array_equal(np.array([len(r.extra_item[s][0]) for r in res]), want[f"{tag}_cnt{s}"])
            idx = np.concatenate([r.extra_item[s][0].cpu().numpy().reshape(-1).astype(np.int64) for r in res])
            assert np.array_equal(idx, want[f"{tag}_idx{s}"])
            feat = np.concatenate([r.extra_item[s][1].cpu().numpy().reshape(len(r.extra_item[s][0]), -1) for r in res])
            np.testing.assert_allclose(feat, want[f"{tag}_feat{s}"], rtol=2e-5, atol=2e-6)          # RoIAlign of boxes equal to 1e-4 px
    else:
        for i, r in enumerate(res):
            for s in range(3):
                assert r.extra_item[s].data_ptr() == inputs["maps"][s][i].data_ptr()


@pytest.mark.parametrize("tag", sorted(CASES))
def test_postprocess_matches_reference_golden(inputs, tag):
    from ood_in_object_detection_b200.postprocess import postprocess
    golden = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_postprocess.npz"))
    mode, before, conf = CASES[tag]
    res = _run(postprocess, inputs, mode, before, conf)
    _check_against(res, tag, mode, golden, inputs)
    if tag in ("fs_hi", "lg_hi"):
        assert 0 in [len(r.boxes) for r in res]                                                    # an image without detections


def test_postprocess_equals_the_reference_run_on_the_device(inputs):
    """The reference's own method on the same CUDA tensors (its code: /root/reference here, oracle/_ref on the GPU box)."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("neither /root/reference nor the compiled build oracle/_ref is present")
    ref_shim.load()
    from ultralytics.models.yolo.detect.predict import DetectionPredictor
    from ood_in_object_detection_b200.postprocess import postprocess
    for tag in ("fs", "lg_raw", "pos"):
        mode, before, conf = CASES[tag]
        ref = _run(DetectionPredictor.postprocess, inputs, mode, before, conf)
        want = {f"{tag}_n": np.array([len(r.boxes) for r in ref]),
                f"{tag}_boxes": np.concatenate([r.boxes.data.cpu().numpy().reshape(-1, 6) for r in ref]),
                f"{tag}_shape": np.array(ref[0].orig_img.shape)}
        if mode == "logits":
            want[f"{tag}_extra"] = np.concatenate([r.extra_item.cpu().numpy().reshape(len(r.boxes), 20) for r in ref])
        else:
            want[f"{tag}_extra"] = np.concatenate([r.extra_item[1].cpu().numpy().reshape(-1) for r in ref]).astype(np.float64)
        _check_against(_run(postprocess, inputs, mode, before, conf), tag, mode, want, inputs)


def test_postprocess_rejects_what_it_does_not_serve(inputs):
    from ood_in_object_detection_b200.postprocess import postprocess
    fp = fake_predictor("ftmaps_and_strides", False, 0.25, device="cuda:0")
    with pytest.raises(NotImplementedError):
        postprocess(fp, ((inputs["p"],), inputs["maps"]), inputs["img"], [np.zeros((320, 320, 3), np.uint8)] * 4)
    fp.args.classes = [1]
    with pytest.raises(NotImplementedError):
        postprocess(fp, ((inputs["p"],), inputs["maps"]), inputs["img"], inputs["img"])
    fa = fake_predictor("ftmaps_and_strides", False, 0.25, device="cuda:0")
    fa.args.agnostic_nms = True                                                                # served: fewer boxes survive
    res_a = postprocess(fa, ((inputs["p"].clone(),), inputs["maps"]), inputs["img"], inputs["img"])
    res_c = postprocess(fake_predictor("ftmaps_and_strides", False, 0.25, device="cuda:0"), ((inputs["p"].clone(),), inputs["maps"]),
                        inputs["img"], inputs["img"])
    assert sum(len(r.boxes) for r in res_a) < sum(len(r.boxes) for r in res_c)
    with pytest.raises(RuntimeError):
        postprocess(fake_predictor("all_ftmaps", False, 0.25), ((inputs["p"].cpu(),), [m.cpu() for m in inputs["maps"]]),
                    inputs["img"].cpu(), inputs["img"].cpu())
