"""Input adapter: bit-exact parity with the reference's CPU preprocessing.

CPU: the oracle restatement (oracle/input_adapter_oracle.py) against Pillow itself, against the committed golden vectors
made from the REAL reference (tools/make_golden_adapter.py), and the product's host logic (geometry + coefficient
tables) against the oracle's.  GPU: the CUDA path against the oracle and the golden vectors, bit for bit."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN


def _cases():
    return json.load(open(os.path.join(GOLDEN, "input_adapter.json")))["cases"]


def _image(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("h,w,ow,oh", [(300, 400, 128, 96), (1080, 1920, 682, 384), (37, 53, 20, 11), (384, 512, 512, 384),
                                       (700, 500, 384, 537), (100, 120, 200, 150)])
def test_oracle_resize_is_pillow(h, w, ow, oh):
    import PIL.Image
    from oracle import input_adapter_oracle as O
    img = _image(11, h, w)
    ours = O.pil_resize_lanczos(img, ow, oh)
    ref = np.asarray(PIL.Image.fromarray(img).resize((ow, oh), resample=PIL.Image.Resampling.LANCZOS))
    assert np.array_equal(ours, ref)


def test_oracle_matches_reference_golden():
    from oracle import input_adapter_oracle as O
    for c in _cases():
        if c["h"] * c["w"] > 2_500_000:      # the 4K case takes ~10 s in numpy: covered on the GPU
            continue
        x = O.preprocess_frame(_image(c["seed"], c["h"], c["w"]), tuple(c["resolution"]), bool(c["square_flip"]))
        assert list(x.shape) == c["shape"]
        assert hashlib.sha256(x.tobytes()).hexdigest() == c["sha256"], c
        for ch, y, xx, v in c["probes"]:
            assert float(x[ch, y, xx]) == v


def test_host_logic_matches_oracle():
    """The product's geometry and Pillow coefficient tables (spann3r_b200/preprocess.py) equal the oracle's."""
    from oracle import input_adapter_oracle as O
    from spann3r_b200 import preprocess as P
    for c in _cases():
        a = P.plan_frame(c["h"], c["w"], tuple(c["resolution"]), bool(c["square_flip"]))
        b = O.plan_frame(c["h"], c["w"], tuple(c["resolution"]), bool(c["square_flip"]))
        assert a == b
    for n_in, n_out in [(1920, 682), (1080, 384), (640, 512), (53, 20), (512, 512), (120, 200), (3840, 597)]:
        b1, k1, s1 = P.lanczos_coeffs(n_in, n_out)
        b2, k2, s2 = O.precompute_coeffs(n_in, 0.0, float(n_in), n_out)
        assert s1 == s2 and np.array_equal(b1, b2) and np.array_equal(k1, k2)


def test_adapter_needs_gpu():
    from spann3r_b200 import preprocess as P
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(Exception):
        P.FrameAdapter()


@pytest.mark.gpu
def test_cuda_adapter_bit_exact():
    from oracle import input_adapter_oracle as O
    from spann3r_b200 import preprocess as P
    adapters = {}
    for c in _cases():
        res = tuple(c["resolution"])
        ad = adapters.setdefault(res, P.FrameAdapter(res))
        img = _image(c["seed"], c["h"], c["w"])
        x = ad(img, square_flip=bool(c["square_flip"]))
        torch.cuda.synchronize()
        x = x[0].cpu().numpy()
        assert list(x.shape) == c["shape"]
        assert hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest() == c["sha256"], c   # == the real reference
        if c["h"] * c["w"] <= 2_500_000:
            assert np.array_equal(x, O.preprocess_frame(img, res, bool(c["square_flip"])))
    # odd sizes / tiny / up-scaling inputs not in the golden set, against the oracle
    ad = P.FrameAdapter((512, 384))
    for seed, (h, w) in enumerate([(301, 399), (97, 131), (2000, 1500), (385, 513)]):
        img = _image(100 + seed, h, w)
        x = ad(img)[0].cpu().numpy()
        assert np.array_equal(x, O.preprocess_frame(img, (512, 384)))


@pytest.mark.gpu
def test_load_frames_feed_the_model_shapes():
    from spann3r_b200 import preprocess as P
    views = P.load_frames([_image(1, 480, 640), _image(2, 480, 640)], (224, 224))
    assert views[0]["img"].shape == (1, 3, 224, 224) and views[0]["img"].is_cuda
    assert float(views[0]["img"].min()) >= -1.0 and float(views[0]["img"].max()) <= 1.0
