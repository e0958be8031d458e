"""Post-path focal estimate (first step): oracle vs golden values from the REAL reference
function (CPU), CUDA path vs both (GPU).  Floating point: sums over ~2e5 pixels in a different order -> 1e-4 relative."""
import importlib.util
import json
import os

import pytest
import torch

from conftest import GOLDEN, ROOT

TOL = 1e-4


def _pointmap():
    spec = importlib.util.spec_from_file_location("make_golden_focal", os.path.join(ROOT, "tools", "make_golden_focal.py"))
    src = open(spec.origin).read().replace('from dust3r.post_process import estimate_focal_knowing_depth  # noqa: E402', '')
    ns = {"__name__": "golden_helper", "__file__": spec.origin}
    exec(compile(src, spec.origin, "exec"), ns)
    return ns["pointmap"]


def _cases():
    return json.load(open(os.path.join(GOLDEN, "focal.json")))["cases"]


def test_oracle_matches_reference_golden():
    from oracle.postprocess_oracle import focal_weiszfeld
    pm = _pointmap()
    for c in _cases():
        pts = pm(c["seed"], c["B"], c["H"], c["W"], c["f_true"])
        f = focal_weiszfeld(pts, (c["W"] / 2, c["H"] / 2))
        for a, b in zip(f.tolist(), c["focal"]):
            assert abs(a - b) <= 2e-6 * abs(b), (a, b)
            assert abs(a - c["f_true"]) < 0.05 * c["f_true"]      # and it does recover the synthetic camera


@pytest.mark.gpu
def test_cuda_focal_matches_reference():
    from oracle.postprocess_oracle import focal_weiszfeld
    from spann3r_b200.postprocess import estimate_focal_knowing_depth
    pm = _pointmap()
    for c in _cases():
        pts = pm(c["seed"], c["B"], c["H"], c["W"], c["f_true"])
        pp = torch.tensor((c["W"] / 2, c["H"] / 2))
        f = estimate_focal_knowing_depth(pts.cuda(), pp, focal_mode="weiszfeld").cpu()
        ref = focal_weiszfeld(pts, pp)
        for a, b, g in zip(f.tolist(), ref.tolist(), c["focal"]):
            assert abs(a - g) <= TOL * abs(g), (a, g)
            assert abs(a - b) <= TOL * abs(b), (a, b)
    # clipping (min_focal / max_focal in units of the 60-degree base focal, post_process.py:55-56)
    c = _cases()[0]
    pts = pm(c["seed"], c["B"], c["H"], c["W"], c["f_true"])
    f = estimate_focal_knowing_depth(pts.cuda(), (c["W"] / 2, c["H"] / 2), min_focal=1.0, max_focal=1.0).cpu()
    assert abs(float(f[0]) - max(c["H"], c["W"]) / 1.1547005383792515) < 1e-2
