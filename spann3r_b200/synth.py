"""Deterministic synthetic checkpoints in the reference's state-dict layout.

There is no network and no trained checkpoint, so parity and the benchmark run on
random-init weights (BASELINE.json: "random-init ViT-L/Base-dec").  The weights must be
bit-identical here (where the real reference runs on CPU and the golden vectors are made)
and on the GPU box (where /root/reference does not exist), so every tensor is drawn from
its own `torch.Generator` seeded by (seed, crc32(key)): the result depends only on the key
name, the shape and the torch version, not on construction order.

The key/shape inventory is `spann3r_b200/state_dict_spec.json` (package data), dumped from the real
reference model by `tools/make_golden.py` (1101 keys for Spann3R).

Init rules follow what the reference constructors do (so activations are conditioned like
the reference's own random init, which SURVEY.md §8d probed finite for 10 frames):
  * nn.Linear weights        xavier-uniform       croco/models/croco.py:111-127
  * conv / conv-transpose    U(+-1/sqrt(fan_in))  (torch default kaiming_uniform(a=sqrt(5)))
  * biases                   small non-zero noise (the reference zeros Linear biases; we do
                             not, so that a dropped bias add cannot hide)
  * LayerNorm gains          1 + 0.1*U(-1,1)      (same reason)
"sharpen=True" multiplies norm_q.weight by 8: with raw random-init
weights the memory-read attention is near-uniform and the eval-mode 5e-4 threshold
(spann3r/model.py:170-172) zeroes whole rows once the bank is large.
"""
from __future__ import annotations

import json
import math
import os
import re
import zlib

import torch

_SPEC_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "state_dict_spec.json")


def load_spec(path: str = _SPEC_PATH) -> dict:
    with open(path) as f:
        return json.load(f)


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFFFFFFFFFF)
    return g


def _uniform(shape, bound, g):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound


def _is_norm_key(key: str) -> bool:
    leaf = key.split(".")[-2]
    return leaf.startswith("norm") or leaf.endswith("_norm") or leaf in ("norm_q", "norm_k", "norm_v")


_ALIAS = re.compile(r"scratch\.layer([1-4])_rn\.")


def canonical_key(key: str) -> str:
    """The DPT head registers each `scratch.layerK_rn` conv twice (also as `scratch.layer_rn.K-1`,
    croco/models/dpt_block.py:59-65); both state-dict keys must carry the same tensor."""
    return _ALIAS.sub(lambda m: f"scratch.layer_rn.{int(m.group(1)) - 1}.", key)


def synth_tensor(key: str, shape, seed: int = 0) -> torch.Tensor:
    key = canonical_key(key)
    g = _gen(seed, key)
    shape = tuple(shape)
    if key.endswith("mask_token"):
        return torch.randn(shape, generator=g, dtype=torch.float32) * 0.02
    if key.endswith(".weight"):
        if len(shape) == 1:
            if _is_norm_key(key):
                return 1.0 + _uniform(shape, 0.1, g)
            raise ValueError(f"unexpected 1-D weight {key}")
        if len(shape) == 2:  # nn.Linear [out, in]
            fan_out, fan_in = shape
            return _uniform(shape, math.sqrt(6.0 / (fan_in + fan_out)), g)
        if len(shape) == 4:
            if key.endswith("patch_embed.proj.weight"):
                # PatchEmbed._init_weights: xavier on w.view(out, -1)   blocks.py:237-239
                fan_out, fan_in = shape[0], shape[1] * shape[2] * shape[3]
                return _uniform(shape, math.sqrt(6.0 / (fan_in + fan_out)), g)
            # Conv2d [out,in,kh,kw] / ConvTranspose2d [in,out,kh,kw]: torch's fan_in = shape[1]*kh*kw
            fan_in = shape[1] * shape[2] * shape[3]
            return _uniform(shape, 1.0 / math.sqrt(fan_in), g)
    if key.endswith(".bias"):
        return _uniform(shape, 0.02, g)
    raise ValueError(f"no init rule for {key} {shape}")


def make_state_dict(spec: dict | None = None, seed: int = 0, sharpen: bool = False,
                    prefix: str | None = None) -> dict:
    """Return {key: fp32 CPU tensor} for every key in the spec (optionally only keys under `prefix`,
    with the prefix stripped -- used to build the DUSt3R-layout checkpoint the ctor consumes)."""
    spec = spec or load_spec()
    out = {}
    for key, shape in spec["spann3r"].items():
        if prefix is not None:
            if not key.startswith(prefix):
                continue
            name = key[len(prefix):]
        else:
            name = key
        t = synth_tensor(key, shape, seed)
        if sharpen and key == "norm_q.weight":
            t = t * 8.0
        out[name] = t
    return out


DUST3R_ARGS = ("AsymmetricCroCo3DStereo(pos_embed='RoPE100', patch_embed_cls='ManyAR_PatchEmbed', "
               "img_size=(512, 512), head_type='dpt', output_mode='pts3d', depth_mode=('exp', -inf, inf), "
               "conf_mode=('exp', 1, inf), enc_embed_dim=1024, enc_depth=24, enc_num_heads=16, "
               "dec_embed_dim=768, dec_depth=12, dec_num_heads=12)")


def make_frames(n_frames: int, height: int, width: int, batch: int = 1, seed0: int = 1):
    """Synthetic frames as SURVEY.md §8d: img_i = rand(B,3,H,W)*2-1 with seed seed0+i."""
    frames = []
    for i in range(n_frames):
        g = torch.Generator(device="cpu")
        g.manual_seed(seed0 + i)
        img = torch.rand((batch, 3, height, width), generator=g, dtype=torch.float32) * 2.0 - 1.0
        frames.append({"img": img})
    return frames


def make_pointmap_case(height: int, width: int, focal: float, rvec, tvec, noise: float, outlier_frac: float, seed: int):
    """A synthetic world-frame pointmap for the post-path geometry tests / bench (numpy): a smooth depth surface seen by a
    pinhole camera (focal, principal point at the image centre) whose pose is x_cam = R(rvec) x_world + tvec, with
    Gaussian noise on the points and a fraction of gross outliers.  Returns (pts3d [H, W, 3] float32, K [3, 3] float64)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    u, v = np.meshgrid(np.arange(width), np.arange(height))
    K = np.array([[focal, 0, width / 2], [0, focal, height / 2], [0, 0, 1]], np.float64)
    d = 2 + 0.5 * np.sin(u / 50.0) + 0.3 * np.cos(v / 40.0)
    xc = np.stack(((u - width / 2) / focal * d, (v - height / 2) / focal * d, d), -1).reshape(-1, 3)
    r = np.asarray(rvec, np.float64)
    th = float(np.linalg.norm(r))
    if th < 1e-12:
        R = np.eye(3)
    else:
        kx, ky, kz = r / th
        Kx = np.array([[0, -kz, ky], [kz, 0, -kx], [-ky, kx, 0]])
        R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    xw = (xc - np.asarray(tvec, np.float64)) @ R          # x_cam = R x_world + t  =>  x_world = R^T (x_cam - t)
    xw += rng.normal(0, noise, xw.shape)
    n_out = int(outlier_frac * len(xw))
    if n_out:
        idx = rng.choice(len(xw), n_out, replace=False)
        xw[idx] += rng.normal(0, 0.5, (n_out, 3))
    return xw.astype(np.float32).reshape(height, width, 3), K


# (height, width, focal, rvec, tvec, noise, outlier_frac, seed): the cases of tests/golden/pnp.json
PNP_CASES = [
    (384, 512, 400.0, (0.1, -0.2, 0.05), (0.3, -0.1, 0.2), 0.002, 0.0, 1),
    (384, 512, 450.0, (0.3, 0.2, -0.1), (-0.5, 0.2, 0.4), 0.005, 0.2, 2),
    (224, 224, 250.0, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0), 0.001, 0.1, 3),
    (384, 512, 420.0, (-0.2, 0.4, 0.3), (0.1, 0.6, -0.3), 0.003, 0.3, 4),
    (384, 512, 380.0, (0.05, 1.2, -0.4), (1.0, -0.3, 0.8), 0.004, 0.45, 5),
]
