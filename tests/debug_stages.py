#!/usr/bin/env python
"""Print per-stage relative errors of the CUDA engine vs the oracle (GPU, strict fp32). Debug aid (parity checker: lives under tests/ because it imports the oracle)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import spann3r_oracle as orc  # noqa: E402
from spann3r_b200 import Spann3R, synth  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (224, 224)
    sdc = synth.make_state_dict(sharpen=True)
    m = Spann3R(dus3r_name=None)
    m.load_state_dict(sdc, strict=True)
    m = m.cuda().eval()
    sd = {k: v.cuda() for k, v in sdc.items()}
    frames = synth.make_frames(2, H, W)
    img = torch.cat([f["img"] for f in frames]).cuda()
    t0 = time.time()
    eng = m._engine_for(1, H, W)
    torch.cuda.synchronize()
    print("pack+engine %.1fs" % (time.time() - t0))
    # encoder, block by block
    x, pos = orc.patch_embed(sd, "dust3r.patch_embed", img)
    feats = eng.encode(img)
    ref, _ = orc.encode_image(sd, img)
    print("encode", rel(feats, ref))
    f1, f2 = ref[:1].contiguous(), ref[1:].contiguous()
    dec_all = eng.decode(f1, f2, want_all=True)
    r1, r2 = orc.decoder(sd, f1, pos[:1], f2, pos[1:])
    for l in range(12):
        print("dec", l, rel(dec_all[l, 0], r1[l + 1]), rel(dec_all[l, 1], r2[l + 1]))
    k1, k2 = eng.keyheads(f1, f2)
    print("key", rel(k1, orc.key_head(sd, 1, f1, r1[-1])), rel(k2, orc.key_head(sd, 2, f2, r2[-1])))
    pts, conf = eng.heads()
    o1 = orc.dpt_head(sd, "dust3r.downstream_head1", r1, H, W)
    o2 = orc.dpt_head(sd, "dust3r.downstream_head2", r2, H, W)
    print("dpt", rel(pts[0], o1["pts3d"]), rel(conf[0], o1["conf"]), rel(pts[1], o2["pts3d"]), rel(conf[1], o2["conf"]))
    rk1 = orc.key_head(sd, 1, f1, r1[-1])
    v = eng.value(o1["pts3d"].contiguous(), rk1.contiguous())
    print("value", rel(v, orc.encode_cur_value(sd, o1["pts3d"]) + rk1))


if __name__ == "__main__":
    main()
