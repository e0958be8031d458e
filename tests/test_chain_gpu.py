"""The persistent GEMM-chain kernel (csrc/gemm_chain.cu): a block's dependent GEMMs (proj -> fc1 -> fc2 -> next qkv; proj -> q;
cproj -> fc1 -> fc2 -> next qkv') as ONE launch with per-row-block dependency counters must give what the same GEMMs give
as separate launches -- on every stage that uses it (encoder at 1 and several images, twin decoder, value encoder), and
again on replay (the counters are reset by the kernel itself)."""
import pytest
import torch

from conftest import get_state_dict, rel_l2

pytestmark = pytest.mark.gpu


def _model(chain: int):
    from spann3r_b200 import Spann3R, _lib
    _lib.check(_lib.lib().s3r_set_option(b"chain", chain), "s3r_set_option")
    m = Spann3R(dus3r_name=None)
    m.load_state_dict(get_state_dict(True), strict=True)
    return m.cuda().eval()


@pytest.fixture(scope="module")
def pair():
    from spann3r_b200 import _lib
    try:
        yield _model(0), _model(1)     # plans are built lazily: every first call below sets the option again
    finally:
        _lib.lib().s3r_set_option(b"chain", 0)      # the library default (the chain is off: profiles/r2_chain.md)


@pytest.mark.parametrize("H,W,nimg", [(224, 224, 2), (384, 512, 2), (384, 512, 10), (224, 224, 5)])
def test_chain_equals_separate_launches_per_stage(pair, H, W, nimg):
    from spann3r_b200 import _lib, synth
    m0, m1 = pair
    L = _lib.lib()
    img = torch.cat([f["img"] for f in synth.make_frames(nimg, H, W)]).cuda()
    outs = []
    for chain, m in ((0, m0), (1, m1)):
        L.s3r_set_option(b"chain", chain)          # read when this engine builds its plans (first call of each stage)
        eng = m._engine_for(1, H, W, n_frames=nimg)
        res = {}
        for rep in range(2):                         # second pass = replay of the cached plans / self-reset counters
            feats = eng.encode(img)
            f1, f2 = feats[:1].contiguous(), feats[1:2].contiguous()
            dec = eng.decode(f1, f2, want_all=True)
            k1, k2 = eng.keyheads(f1, f2)
            pts, conf = eng.heads()
            val = eng.value(pts[0].contiguous(), k1)
            torch.cuda.synchronize()
            res[rep] = dict(feats=feats.clone(), dec=dec.clone(), k1=k1.clone(), pts=pts.clone(), conf=conf.clone(), val=val.clone())
        for k in res[0]:
            assert torch.equal(res[0][k], res[1][k]), ("replay differs", chain, k)
        outs.append(res[0])
    worst = max(rel_l2(outs[1][k].cpu(), outs[0][k].cpu()) for k in outs[0])
    print("chain vs separate launches, worst rel-L2 over stages: %.2e" % worst)
    for k in outs[0]:
        assert torch.isfinite(outs[1][k]).all(), k
        assert rel_l2(outs[1][k].cpu(), outs[0][k].cpu()) < 2e-5, k    # same arithmetic; pair tiles vs 1-CTA tiles may reorder fp32 sums


def test_chain_sequence_end_to_end_and_launch_count(pair):
    from spann3r_b200 import _lib, synth
    m0, m1 = pair
    L = _lib.lib()
    frames = synth.make_frames(4, 224, 224)
    L.s3r_set_option(b"chain", 0)
    p0, _ = m0(frames)
    e0 = m0._engine_for(1, 224, 224, n_frames=4)
    e0.take_launches()
    p0, _ = m0(frames)
    n0 = e0.take_launches()
    p0 = [{k: v.clone() for k, v in p.items()} for p in p0]
    L.s3r_set_option(b"chain", 1)
    p1, _ = m1(frames)
    e1 = m1._engine_for(1, 224, 224, n_frames=4)
    e1.take_launches()
    p1, _ = m1(frames)
    n1 = e1.take_launches()
    print("launches per 4-frame sequence: separate", n0, "chained", n1)
    for a, b in zip(p0, p1):
        for k in a:
            assert rel_l2(b[k].cpu(), a[k].cpu()) < 5e-5, k
    assert n1 < n0      # the 4 x 196-token encoder call has 7 row tiles (odd: no CTA pairs) and stays unchained; the steps chain
