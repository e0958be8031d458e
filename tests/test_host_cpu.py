"""CPU-side checks: state-dict layout of the drop-in module, C-ABI symbol table, host logic."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def test_state_dict_keys_match_reference(spec):
    from spann3r_b200 import Spann3R
    m = Spann3R(dus3r_name=None)
    sd = m.state_dict()
    assert list(sd.keys()) == list(spec["spann3r"].keys())          # same keys, same order as the reference
    for k, shape in spec["spann3r"].items():
        assert list(sd[k].shape) == shape, k
    # aliased DPT convs share storage, like the reference (dpt_block.py:59-65)
    a = sd["dust3r.downstream_head1.dpt.scratch.layer1_rn.weight"]
    b = sd["dust3r.downstream_head1.dpt.scratch.layer_rn.0.weight"]
    assert a.data_ptr() == b.data_ptr()


def test_strict_load_and_roundtrip(spec):
    from spann3r_b200 import Spann3R, synth
    m = Spann3R(dus3r_name=None)
    sd = synth.make_state_dict(spec)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    out = m.state_dict()
    for k in ("dust3r.enc_blocks.3.attn.qkv.weight", "norm_q.bias", "attn_head_2.2.weight"):
        assert torch.equal(out[k], sd[k])


def test_dust3r_checkpoint_ctor(tmp_path, spec):
    """Constructor path of the reference: Spann3R(dus3r_name=<DUSt3R ckpt file>) (spann3r/model.py:222)."""
    import argparse
    from spann3r_b200 import Spann3R, synth
    sd = synth.make_state_dict(spec, prefix="dust3r.")
    sd = {k: v for k, v in sd.items() if not k.startswith("dec_blocks2")}   # released ckpts lack dec_blocks2
    path = tmp_path / "fake_dust3r.pth"
    torch.save({"args": argparse.Namespace(model=synth.DUST3R_ARGS), "model": sd}, path)
    m = Spann3R(dus3r_name=str(path))
    assert torch.equal(m.dust3r.dec_blocks2._modules["0"].attn.qkv.weight, m.dust3r.dec_blocks._modules["0"].attn.qkv.weight)
    assert torch.equal(m.pos_patch_embed.proj.weight, m.dust3r.patch_embed.proj.weight)


def test_no_cpu_fallback(spec):
    from spann3r_b200 import Spann3R, synth
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = Spann3R(dus3r_name=None).eval()
    with pytest.raises(Exception):
        m(synth.make_frames(2, 64, 64))


def test_library_exports_every_declared_symbol():
    from spann3r_b200 import _lib, engine  # noqa: F401  (engine registers the model-level prototypes)
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "spann3r_b200.h")).read()
    declared = set(re.findall(r"\b(s3r_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 24
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/spann3r_b200.h but not exported"
    assert declared == set(_lib.declared_symbols()), declared ^ set(_lib.declared_symbols())
    assert L.s3r_version() == 100


def test_rope_table_matches_reference_fallback():
    """engine.rope_cs_table == cos/sin of croco/models/pos_embed.py:120-129 (through the pinned oracle)."""
    from oracle.spann3r_oracle import rope_tables
    from spann3r_b200.engine import rope_cs_table
    cs = rope_cs_table(64)
    cos, sin = rope_tables(32, 64)
    assert torch.equal(cs[..., 0], cos[:, :16]) and torch.equal(cs[..., 1], sin[:, :16])


def test_ctypes_mirrors_match_the_compiled_structs():
    """The ctypes mirrors of s3r_gemm_desc / s3r_model_w / s3r_bank have the sizes the library was compiled with."""
    from spann3r_b200 import _lib, engine
    L = _lib.lib()
    assert L.s3r_abi_sizeof(0) == ctypes.sizeof(_lib.GemmDesc)
    assert L.s3r_abi_sizeof(1) == ctypes.sizeof(engine.ModelW)
    assert L.s3r_abi_sizeof(2) == ctypes.sizeof(engine.Bank)


def test_layernorm_fold_algebra():
    """engine.fold_layernorm + the epilogue formula of the folded LayerNorm (gemm_epilogue.cuh: per-row statistics
    from 32-column chunk sums, rstd * (acc - mean * colsum) + bias') reproduce Linear(LayerNorm(x)) of
    croco/models/blocks.py:127-130 -- restated here in fp64 / fp32 on the CPU."""
    from spann3r_b200.engine import fold_layernorm
    g = torch.Generator().manual_seed(0)
    C_, N_, R_ = 768, 96, 40
    x = torch.randn(R_, C_, generator=g) * 3 + 0.7          # non-zero mean: exercises the mean * colsum term
    w = torch.randn(N_, C_, generator=g) * C_ ** -0.5
    b = torch.randn(N_, generator=g)
    gamma = 1 + 0.2 * torch.randn(C_, generator=g)
    beta = 0.1 * torch.randn(C_, generator=g)
    ref = torch.nn.functional.linear(torch.nn.functional.layer_norm(x.double(), (C_,), gamma.double(), beta.double(), 1e-6),
                                     w.double(), b.double())
    wf, bf = fold_layernorm(w, b, gamma, beta)
    # producer epilogue: (sum, sum of squares) per 32-column chunk, fp32
    ch = x.view(R_, C_ // 32, 32)
    s1, s2 = ch.sum(-1), (ch * ch).sum(-1)
    mean = s1.sum(-1) / C_
    var = (s2.sum(-1) / C_ - mean * mean).clamp_min(0)
    rstd = torch.rsqrt(var + 1e-6)
    acc = x @ wf.t()                                        # the tensor-core GEMM on the raw stream (fp32 here)
    cs = wf.sum(dim=1)
    out = rstd[:, None] * acc - (rstd * mean)[:, None] * cs[None, :] + bf[None, :]
    err = float((out.double() - ref).norm() / ref.norm())
    assert err < 2e-6, err
