"""csrc/nms.cu (NMS with the OoD payload, one CTA per image) against the golden vectors of the reference's
`non_max_suppression_old` (ultralytics/utils/ops.py:348-530) and against the oracle on ragged random batches."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from tests.helpers import nms_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nms():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from ood_in_object_detection_b200 import nms
    return nms


def test_nms_matches_reference_golden(nms, golden):
    g = golden("golden_nms.npz")
    pred, logits, strides = nms_inputs(int(g["seed"]))
    tp, tl, ts = torch.from_numpy(pred).cuda(), torch.from_numpy(logits).cuda(), torch.from_numpy(strides).cuda()
    for tag in ("a", "b", "agn"):                                                             # agn: class-agnostic suppression
        conf, iou, max_det = g[f"{tag}_cfg"]
        out, extra, st = nms.non_max_suppression(tp, conf, iou, max_det=int(max_det), extra_item=tl, strides=ts, agnostic=tag == "agn")
        assert [len(o) for o in out] == g[f"{tag}_n"].tolist()
        assert np.array_equal(torch.cat(out).cpu().numpy(), g[f"{tag}_det"])                 # boxes, confidence, class: bit for bit
        assert np.array_equal(torch.cat([e.reshape(len(o), -1) for e, o in zip(extra, out)]).cpu().numpy(), g[f"{tag}_extra"])
        assert np.array_equal(torch.cat(st).cpu().numpy(), g[f"{tag}_strides"])
    only = nms.non_max_suppression(tp, 0.25, 0.45)                                            # no payload: just the list
    assert isinstance(only, list) and [len(o) for o in only] == g["a_n"].tolist()
    with pytest.raises(NotImplementedError):
        nms.non_max_suppression(tp, 0.25, 0.45, classes=[1, 2])


def test_nms_matches_oracle_on_ragged_batches(nms):
    from oracle import nms as onms
    for seed, (conf, iou, max_det) in ((5, (0.3, 0.5, 300)), (6, (0.05, 0.6, 100)), (7, (0.999999, 0.45, 300))):
        pred, logits, strides = nms_inputs(seed, bs=6, nc=7, img=320)
        ref_out, ref_ex, ref_st = onms.non_max_suppression(pred, conf, iou, max_det, extra_item=logits, strides=strides)
        out, ex, st = nms.non_max_suppression(torch.from_numpy(pred).cuda(), conf, iou, max_det=max_det,
                                              extra_item=torch.from_numpy(logits).cuda(), strides=torch.from_numpy(strides).cuda())
        for i in range(len(ref_out)):
            assert np.array_equal(out[i].cpu().numpy().reshape(-1, 6), ref_out[i]), (seed, i)
            if len(ref_out[i]):
                assert np.array_equal(ex[i].cpu().numpy(), ref_ex[i]) and np.array_equal(st[i].cpu().numpy(), ref_st[i])
            else:
                assert ex[i].numel() == 0 and st[i].numel() == 0
