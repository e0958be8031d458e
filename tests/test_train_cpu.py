"""CPU tests of the training-mode host logic (spann3r_b200/train.py, _recompute.py).

`_recompute.py` is what the BACKWARD pass of training mode differentiates (PyTorch recompute of each engine stage); the
forward always runs the CUDA library.  Here the restatements are pinned, stage by stage, to the oracle (itself pinned to the
real reference), and the autograd plumbing of `train._Stage` is checked with a stand-in for the native call."""
import numpy as np
import pytest
import torch

from conftest import get_state_dict, rel_l2
from oracle import spann3r_oracle as orc
from spann3r_b200 import _recompute as R
from spann3r_b200 import synth, train

H, W = 64, 96


@pytest.fixture(scope="module")
def sd():
    return get_state_dict(True)


def test_recompute_stages_equal_the_oracle(sd):
    torch.manual_seed(0)
    fr = synth.make_frames(2, H, W)
    img = torch.cat([f["img"] for f in fr])
    with torch.no_grad():
        feats = R.encode(sd, img)
        ref, pos = orc.encode_image(sd, img)
        assert rel_l2(feats, ref) < 1e-5
        f1, f2 = ref[:1], ref[1:]
        fuse = f1 + 0.1 * torch.randn_like(f1)
        k1, k2, pts, conf = R.step(sd, fuse, f1, f2, H, W)
        d1, d2 = orc.decoder(sd, fuse, pos[:1], f2, pos[1:])
        assert rel_l2(k1, orc.key_head(sd, 1, f1, d1[-1])) < 1e-5 and rel_l2(k2, orc.key_head(sd, 2, f2, d2[-1])) < 1e-5
        r1 = orc.dpt_head(sd, "dust3r.downstream_head1", d1, H, W)
        r2 = orc.dpt_head(sd, "dust3r.downstream_head2", d2, H, W)
        assert rel_l2(pts[0], r1["pts3d"]) < 1e-5 and rel_l2(conf[1], r2["conf"]) < 1e-5 and rel_l2(pts[1], r2["pts3d"]) < 1e-5
        for rope in (False, True):
            v = R.value(sd, r1["pts3d"], k1, rope)
            assert rel_l2(v, orc.encode_cur_value(sd, r1["pts3d"], mem_pos_enc=rope) + k1) < 1e-5
        # training-mode read: attn_thresh = 0 (no cut, no renormalisation), optional dropout keep-scale
        om = orc.SpatialMemory(sd, attn_thresh=0)
        g = torch.Generator().manual_seed(3)
        ks = [torch.randn(1, 24, 1024, generator=g) for _ in range(2)]
        vs = [torch.randn(1, 24, 1024, generator=g) for _ in range(2)]
        for k, v in zip(ks, vs):
            om.add_mem(k, v)
        q = torch.randn(1, 24, 1024, generator=g)
        assert rel_l2(R.memory_read(sd, q, torch.cat(ks, 1), torch.cat(vs, 1)), om.memory_read(q)) < 1e-5


def test_stage_function_backward_is_autograd_of_the_recompute():
    """train._Stage: forward = the 'native' callable (no graph), backward = autograd of the torch restatement -> the
    gradients must equal plain autograd through the restatement (activations AND parameters, unused ones None)."""
    torch.manual_seed(1)
    names = ["w", "b", "unused"]
    params = [torch.randn(5, 7, requires_grad=True), torch.randn(5, requires_grad=True), torch.randn(3, requires_grad=True)]
    x = torch.randn(4, 7, requires_grad=True)

    def torch_fn(P, a):
        y = torch.tanh(a @ P["w"].t() + P["b"])
        return y, y.sum(dim=1)

    calls = []

    def native(a):
        calls.append(a.requires_grad)
        return tuple(t.detach() for t in torch_fn(dict(zip(names, params)), a))

    y, s = train._apply(native, torch_fn, names, params, x)
    assert calls == [True] or calls == [False]
    loss = (y * torch.arange(5.0)).sum() + (s ** 2).sum()
    loss.backward()
    got = [x.grad.clone()] + [p.grad.clone() if p.grad is not None else None for p in params]
    x.grad = None
    for p in params:
        p.grad = None
    y2, s2 = torch_fn(dict(zip(names, params)), x)
    ((y2 * torch.arange(5.0)).sum() + (s2 ** 2).sum()).backward()
    assert torch.allclose(got[0], x.grad, atol=1e-6)
    assert torch.allclose(got[1], params[0].grad, atol=1e-6) and torch.allclose(got[2], params[1].grad, atol=1e-6)
    assert got[3] is None and params[2].grad is None


def test_stage_parameter_partition_covers_every_key_once(spec):
    """Every parameter belongs to exactly one stage Function (else it would get no / a double gradient)."""
    from spann3r_b200 import Spann3R
    m = Spann3R(dus3r_name=None)
    seen = {}
    for stage in ("encode", "memread", "step", "value"):
        names, params = train.stage_params(m, stage)
        assert len(names) == len(set(names)) and len(names) == len(params)
        for n in names:
            assert n not in seen, (n, stage, seen.get(n))
            seen[n] = stage
    every = [n for n, _ in m.named_parameters(remove_duplicate=False)]
    missing = sorted(set(every) - set(seen))
    assert missing == ["dust3r.mask_token"], missing        # unused by the forward path (dust3r/model.py: masking is off)
    assert all(p.requires_grad for p in m.parameters())     # trainable by default, like the reference's modules


def philox_keep_scale_numpy(seed: int, n: int, p: float) -> np.ndarray:
    """Host restatement of csrc/memory.cu:dropout_scale (Philox4x32-10, counter = idx / 4, key = seed)."""
    idx = np.arange(n, dtype=np.uint64)
    c = idx >> np.uint64(2)
    ctr = [(c & np.uint64(0xFFFFFFFF)).astype(np.uint64), (c >> np.uint64(32)).astype(np.uint64),
           np.zeros(n, np.uint64), np.zeros(n, np.uint64)]
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * ctr[0], M1 * ctr[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        ctr = [hi1 ^ ctr[1] ^ k0, lo1, hi0 ^ ctr[3] ^ k1, lo0]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & MASK, (k1 + np.uint64(0xBB67AE85)) & MASK
    words = np.stack(ctr, axis=1)[np.arange(n), (idx & np.uint64(3)).astype(np.int64)]
    u = (words >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return np.where(u >= np.float32(p), np.float32(1.0 / (1.0 - p)), np.float32(0.0)).astype(np.float32)


def test_philox_host_restatement_statistics():
    ks = philox_keep_scale_numpy(1234567890123, 200000, 0.15)
    keep = (ks > 0).mean()
    assert abs(keep - 0.85) < 0.004 and np.allclose(ks[ks > 0], 1 / 0.85)
    assert not np.array_equal(ks, philox_keep_scale_numpy(1234567890124, 200000, 0.15))
    # known answer of Philox4x32-10 (Random123 kat_vectors: counter 0, key 0)
    z = np.zeros(1, np.uint64)
    ctr, k0, k1 = [z.copy(), z.copy(), z.copy(), z.copy()], np.uint64(0), np.uint64(0)
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * ctr[0], M1 * ctr[2]
        ctr = [(p1 >> np.uint64(32)) ^ ctr[1] ^ k0, p1 & MASK, (p0 >> np.uint64(32)) ^ ctr[3] ^ k1, p0 & MASK]
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & MASK, (k1 + np.uint64(0xBB67AE85)) & MASK
    assert [int(v[0]) for v in ctr] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
