"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product; nothing under spann3r_b200/ imports it.

A plain-PyTorch fp32 restatement of the reference's per-frame forward path
(HengyiWang/spann3r @ f89d6a23: `Spann3R.forward`, spann3r/model.py:473-539, and everything it
calls), written as functions over a state dict with the reference's key names.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it,
and only as the checker / CPU baseline.

Parity pin: the reference has no tests or golden vectors, so this restatement is
pinned against outputs of the REAL reference executed in the authoring container
(`tools/make_golden.py` -> `tests/golden/*.npz`); `tests/test_oracle_vs_golden.py` holds it to
<= 2e-5 relative L2 (fp32 reassociation noise) on every stored tensor.  Each function cites the
reference file:line it follows (paths relative to the reference root).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------------
def rope_tables(D: int, max_pos: int, base: float = 100.0, device="cpu"):
    """cos/sin tables of the PyTorch RoPE2D fallback.  croco/models/pos_embed.py:120-129 (D = head_dim/2)."""
    inv_freq = 1.0 / (base ** (torch.arange(0, D, 2, device=device).float() / D))
    t = torch.arange(max_pos, device=device, dtype=inv_freq.dtype)
    freqs = torch.einsum("i,j->ij", t, inv_freq)
    freqs = torch.cat((freqs, freqs), dim=-1)
    return freqs.cos(), freqs.sin()


# bench.py's eager-GPU baseline leg may install a fused RoPE here (signature of rope2d) to time the reference "with a
# working curope" (croco/models/pos_embed.py:106-111 prefers cuRoPE2D when the extension imports); None = the PyTorch
# fallback the reference uses when it does not.  Never set by tests: the checker always runs the exact fallback.
ROPE_OVERRIDE = None


def rope2d(tokens: torch.Tensor, positions: torch.Tensor, base: float = 100.0):
    """tokens [B,H,N,dh], positions [B,N,2] (y,x).  croco/models/pos_embed.py:131-159."""
    if ROPE_OVERRIDE is not None and tokens.is_cuda:
        return ROPE_OVERRIDE(tokens, positions, base)
    D = tokens.size(3) // 2
    cos, sin = rope_tables(D, int(positions.max()) + 1, base, tokens.device)

    def rotate_half(x):
        x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
        return torch.cat((-x2, x1), dim=-1)

    def apply1d(tok, pos1d):
        c = F.embedding(pos1d, cos)[:, None, :, :]
        s = F.embedding(pos1d, sin)[:, None, :, :]
        return (tok * c) + (rotate_half(tok) * s)

    y, x = tokens.chunk(2, dim=-1)
    y = apply1d(y, positions[:, :, 0])
    x = apply1d(x, positions[:, :, 1])
    return torch.cat((y, x), dim=-1)


def linear(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def layernorm(sd, name, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def mlp(sd, name, x):
    """croco/models/blocks.py:73-79 (GELU = exact erf form)."""
    return linear(sd, name + ".fc2", F.gelu(linear(sd, name + ".fc1", x)))


def attention(sd, name, x, xpos, num_heads, use_rope=True):
    """croco/models/blocks.py:94-112."""
    B, N, C = x.shape
    qkv = linear(sd, name + ".qkv", x).reshape(B, N, 3, num_heads, C // num_heads).transpose(1, 3)
    q, k, v = [qkv[:, :, i] for i in range(3)]
    if use_rope:
        q = rope2d(q, xpos)
        k = rope2d(k, xpos)
    attn = (q @ k.transpose(-2, -1)) * ((C // num_heads) ** -0.5)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return linear(sd, name + ".proj", x)


def cross_attention(sd, name, query, key, value, qpos, kpos, num_heads):
    """croco/models/blocks.py:149-169."""
    B, Nq, C = query.shape
    Nk = key.shape[1]
    dh = C // num_heads
    q = linear(sd, name + ".projq", query).reshape(B, Nq, num_heads, dh).permute(0, 2, 1, 3)
    k = linear(sd, name + ".projk", key).reshape(B, Nk, num_heads, dh).permute(0, 2, 1, 3)
    v = linear(sd, name + ".projv", value).reshape(B, Nk, num_heads, dh).permute(0, 2, 1, 3)
    q = rope2d(q, qpos)
    k = rope2d(k, kpos)
    attn = (q @ k.transpose(-2, -1)) * (dh ** -0.5)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, Nq, C)
    return linear(sd, name + ".proj", x)


def block(sd, name, x, xpos, num_heads, use_rope=True, eps=1e-6):
    """croco/models/blocks.py:127-130."""
    x = x + attention(sd, name + ".attn", layernorm(sd, name + ".norm1", x, eps), xpos, num_heads, use_rope)
    x = x + mlp(sd, name + ".mlp", layernorm(sd, name + ".norm2", x, eps))
    return x


def decoder_block(sd, name, x, y, xpos, ypos, num_heads, eps=1e-6):
    """croco/models/blocks.py:186-191."""
    x = x + attention(sd, name + ".attn", layernorm(sd, name + ".norm1", x, eps), xpos, num_heads)
    y_ = layernorm(sd, name + ".norm_y", y, eps)
    x = x + cross_attention(sd, name + ".cross_attn", layernorm(sd, name + ".norm2", x, eps), y_, y_, xpos, ypos,
                            num_heads)
    x = x + mlp(sd, name + ".mlp", layernorm(sd, name + ".norm3", x, eps))
    return x


def patch_embed(sd, name, img):
    """dust3r/patch_embed.py:19-29 + croco/models/blocks.py:195-207 (positions = cartesian_prod(y, x))."""
    x = F.conv2d(img, sd[name + ".proj.weight"], sd[name + ".proj.bias"], stride=16)
    B, _, gh, gw = x.shape
    pos = torch.cartesian_prod(torch.arange(gh, device=img.device), torch.arange(gw, device=img.device))
    pos = pos.view(1, gh * gw, 2).expand(B, -1, 2).clone()
    return x.flatten(2).transpose(1, 2), pos


# ------------------------------------------------------------------------------------------------
# DUSt3R stages
# ------------------------------------------------------------------------------------------------
ENC_DEPTH, ENC_HEADS, DEC_DEPTH, DEC_HEADS = 24, 16, 12, 12


def encode_image(sd, img):
    """dust3r/model.py:131-154."""
    x, pos = patch_embed(sd, "dust3r.patch_embed", img)
    for i in range(ENC_DEPTH):
        x = block(sd, f"dust3r.enc_blocks.{i}", x, pos, ENC_HEADS)
    return layernorm(sd, "dust3r.enc_norm", x, 1e-6), pos


def decoder(sd, f1, pos1, f2, pos2):
    """dust3r/model.py:186-205.  Returns (dec1, dec2): 13 tensors each."""
    out = [(f1, f2)]
    f1 = linear(sd, "dust3r.decoder_embed", f1)
    f2 = linear(sd, "dust3r.decoder_embed", f2)
    out.append((f1, f2))
    for i in range(DEC_DEPTH):
        a, b = out[-1]
        n1 = decoder_block(sd, f"dust3r.dec_blocks.{i}", a, b, pos1, pos2, DEC_HEADS)
        n2 = decoder_block(sd, f"dust3r.dec_blocks2.{i}", b, a, pos2, pos1, DEC_HEADS)
        out.append((n1, n2))
    del out[1]
    out[-1] = tuple(layernorm(sd, "dust3r.dec_norm", t, 1e-6) for t in out[-1])
    return list(zip(*out))


def _rcu(sd, name, x):
    """ResidualConvUnit_custom, croco/models/dpt_block.py:121-142 (bn=False, ReLU not in place)."""
    out = F.relu(x)
    out = F.conv2d(out, sd[name + ".conv1.weight"], sd[name + ".conv1.bias"], padding=1)
    out = F.relu(out)
    out = F.conv2d(out, sd[name + ".conv2.weight"], sd[name + ".conv2.bias"], padding=1)
    return out + x


def _fusion(sd, name, *xs):
    """FeatureFusionBlock_custom, croco/models/dpt_block.py:189-218 (width_ratio=1, align_corners=True)."""
    output = xs[0]
    if len(xs) == 2:
        output = output + _rcu(sd, name + ".resConfUnit1", xs[1])
    output = _rcu(sd, name + ".resConfUnit2", output)
    output = F.interpolate(output, scale_factor=2, mode="bilinear", align_corners=True)
    return F.conv2d(output, sd[name + ".out_conv.weight"], sd[name + ".out_conv.bias"])


def dpt_head(sd, name, dec, H, W, return_raw=False):
    """DPTOutputAdapter_fix.forward, dust3r/heads/dpt_head.py:34-65 (hooks [0,6,9,12]) + postprocess.py:10-58."""
    p = name + ".dpt"
    nh, nw = H // 16, W // 16
    layers = [dec[h] for h in (0, 6, 9, 12)]
    layers = [t.reshape(t.shape[0], nh, nw, t.shape[-1]).permute(0, 3, 1, 2) for t in layers]
    ap = p + ".act_postprocess"
    l0 = F.conv2d(layers[0], sd[ap + ".0.0.weight"], sd[ap + ".0.0.bias"])
    l0 = F.conv_transpose2d(l0, sd[ap + ".0.1.weight"], sd[ap + ".0.1.bias"], stride=4)
    l1 = F.conv2d(layers[1], sd[ap + ".1.0.weight"], sd[ap + ".1.0.bias"])
    l1 = F.conv_transpose2d(l1, sd[ap + ".1.1.weight"], sd[ap + ".1.1.bias"], stride=2)
    l2 = F.conv2d(layers[2], sd[ap + ".2.0.weight"], sd[ap + ".2.0.bias"])
    l3 = F.conv2d(layers[3], sd[ap + ".3.0.weight"], sd[ap + ".3.0.bias"])
    l3 = F.conv2d(l3, sd[ap + ".3.1.weight"], sd[ap + ".3.1.bias"], stride=2, padding=1)
    ls = [l0, l1, l2, l3]
    ls = [F.conv2d(t, sd[p + f".scratch.layer_rn.{i}.weight"], None, padding=1) for i, t in enumerate(ls)]
    path4 = _fusion(sd, p + ".scratch.refinenet4", ls[3])[:, :, : ls[2].shape[2], : ls[2].shape[3]]
    path3 = _fusion(sd, p + ".scratch.refinenet3", path4, ls[2])
    path2 = _fusion(sd, p + ".scratch.refinenet2", path3, ls[1])
    path1 = _fusion(sd, p + ".scratch.refinenet1", path2, ls[0])
    out = F.conv2d(path1, sd[p + ".head.0.weight"], sd[p + ".head.0.bias"], padding=1)
    out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
    out = F.conv2d(out, sd[p + ".head.2.weight"], sd[p + ".head.2.bias"], padding=1)
    out = F.relu(out)
    out = F.conv2d(out, sd[p + ".head.4.weight"], sd[p + ".head.4.bias"])
    if return_raw:
        return out
    return postprocess(out)


def downstream_head(sd, name, dec, H, W):
    """AsymmetricCroCo3DStereo.head{1,2} = transpose_to_landscape(head, activate=landscape_only=True)
    (dust3r/model.py:128-129, spann3r/model.py:222; wrapper_yes, dust3r/utils/misc.py:66-94).  With the
    PatchEmbedDust3R that load_model substitutes (dust3r/model.py:33) the tokens keep the image's own (H/16, W/16)
    grid, so for a portrait frame the wrapper calls head(decout, (max, min)) = the true (H, W) and then swaps axes
    1 and 2 of every output: callers -- and the value encoder -- see a [B, W, H, ...] landscape map."""
    res = dpt_head(sd, name, dec, H, W)
    if H > W:
        res = {k: v.swapaxes(1, 2) for k, v in res.items()}
    return res


def postprocess(out):
    """dust3r/heads/postprocess.py:10-58 with depth_mode=('exp',-inf,inf), conf_mode=('exp',1,inf)."""
    fmap = out.permute(0, 2, 3, 1)
    xyz = fmap[..., 0:3]
    d = xyz.norm(dim=-1, keepdim=True)
    xyz = xyz / d.clip(min=1e-8)
    pts3d = xyz * torch.expm1(d)
    conf = 1 + fmap[..., 3].exp().clip(max=float("inf"))
    return {"pts3d": pts3d, "conf": conf}


# ------------------------------------------------------------------------------------------------
# Spann3R: spatial memory + frame loop
# ------------------------------------------------------------------------------------------------
class SpatialMemory:
    """spann3r/model.py:11-210 (eval-mode semantics: no dropout; attn_thresh as given)."""

    def __init__(self, sd, long_mem_size=4000, work_mem_size=5, attn_thresh=5e-4, sim_thresh=0.95):
        self.sd = sd
        self.attn_thresh = attn_thresh
        self.long_mem_size = long_mem_size
        self.work_mem_size = work_mem_size
        self.top_k = long_mem_size
        self.sim_thresh = sim_thresh
        self.num_patches = None
        self.mem_k = self.mem_v = self.mem_count = self.mem_attn = None
        self.lm = 0
        self.wm = 0

    def add_mem(self, feat_k, feat_v):  # :80-95
        if self.num_patches is None:
            self.num_patches = feat_k.shape[1]
        z = torch.zeros_like(feat_k[:, :, :1])
        if self.mem_count is None:
            self.mem_count, self.mem_attn = z, z.clone()
            self.mem_k, self.mem_v = feat_k, feat_v
        else:
            self.mem_count = torch.cat((self.mem_count + 1, z), dim=1)
            self.mem_attn = torch.cat((self.mem_attn, z.clone()), dim=1)
            self.mem_k = torch.cat((self.mem_k, feat_k), dim=1)
            self.mem_v = torch.cat((self.mem_v, feat_v), dim=1)

    def check_sim(self, feat_k, thresh):  # :97-118
        if self.mem_k is None or thresh == 1.0:
            return False
        wmem = self.wm * self.num_patches
        wm = self.mem_k[:, -wmem:].reshape(self.mem_k.shape[0], -1, self.num_patches, self.mem_k.shape[-1])
        corr = torch.einsum("bpc,btpc->btp", F.normalize(feat_k, p=2, dim=-1), F.normalize(wm, p=2, dim=-1))
        return bool(corr.mean(dim=-1).max() > thresh)

    def add_mem_check(self, feat_k, feat_v):  # :120-143
        if self.num_patches is None:
            self.num_patches = feat_k.shape[1]
        if self.check_sim(feat_k, self.sim_thresh):
            return
        self.add_mem(feat_k, feat_v)
        self.wm += 1
        if self.wm > self.work_mem_size:
            self.wm -= 1
            if self.long_mem_size == 0:
                P = self.num_patches
                self.mem_k, self.mem_v = self.mem_k[:, P:], self.mem_v[:, P:]
                self.mem_count, self.mem_attn = self.mem_count[:, P:], self.mem_attn[:, P:]
            else:
                self.lm += self.num_patches
        if self.lm > self.long_mem_size:
            self.memory_prune()
            self.lm = self.top_k - self.wm * self.num_patches

    def memory_read(self, feat, res=True):  # :145-183
        sd = self.sd
        q = layernorm(sd, "norm_q", feat, 1e-5)
        k = layernorm(sd, "norm_k", self.mem_k, 1e-5)
        affinity = torch.einsum("bpc,bxc->bpx", q, k) / math.sqrt(feat.shape[-1])
        attn = torch.softmax(affinity, dim=-1)
        if self.attn_thresh > 0:
            attn = torch.where(attn < self.attn_thresh, torch.zeros_like(attn), attn)
            attn = attn / attn.sum(dim=-1, keepdim=True)
        out = torch.einsum("bpx,bxc->bpc", attn, layernorm(sd, "norm_v", self.mem_v, 1e-5))
        if res:
            out = out + feat
        self.mem_attn = self.mem_attn + attn.sum(dim=-2)[..., None]
        return out

    def memory_prune(self):  # :185-210
        weights = self.mem_attn / self.mem_count
        weights[self.mem_count < self.work_mem_size + 5] = 1e8
        _, idx = torch.topk(weights, self.top_k, dim=1)
        idx_e = idx.expand(-1, -1, self.mem_k.size(-1))
        self.mem_k = torch.gather(self.mem_k, -2, idx_e)
        self.mem_v = torch.gather(self.mem_v, -2, idx_e)
        self.mem_attn = torch.gather(self.mem_attn, -2, idx)
        self.mem_count = torch.gather(self.mem_count, -2, idx)


def key_head(sd, num, feat, dec_last):
    """spann3r/model.py:299-303, 250-261 (Linear 1792->1792, GELU, Linear 1792->1024)."""
    x = torch.cat((feat, dec_last), dim=-1)
    x = F.gelu(linear(sd, f"attn_head_{num}.0", x))
    return linear(sd, f"attn_head_{num}.2", x)


def encode_cur_value(sd, pts3d, mem_pos_enc=False):
    """spann3r/model.py:305-320 (use_feat=False).  mem_pos_enc (ctor flag, :228-235) puts dust3r's RoPE into the
    value encoder's blocks, with the positions pos_patch_embed returns for the pointmap it is given."""
    x, pos = patch_embed(sd, "pos_patch_embed", pts3d.permute(0, 3, 1, 2))
    for i in range(6):
        x = block(sd, f"value_encoder.{i}", x, pos, 16, use_rope=mem_pos_enc)
    x = layernorm(sd, "value_norm", x, 1e-6)
    return linear(sd, "value_out", x)


@torch.no_grad()
def forward(sd, frames, return_memory=False, trace=None, mem_pos_enc=False, **mem_kw):
    """Spann3R.forward in eval mode, spann3r/model.py:473-539.  frames: list of {'img': [B,3,H,W]}."""
    sp_mem = SpatialMemory(sd, **mem_kw)
    feat1 = feat2 = pos1 = pos2 = None
    feat_k2 = None
    preds, preds_all = None, []
    H, W = frames[0]["img"].shape[-2:]
    for i in range(len(frames) - 1):
        if feat1 is None:  # encode_image_pairs :272-287
            out, pos = encode_image(sd, torch.cat((frames[i]["img"], frames[i + 1]["img"]), dim=0))
            feat1, feat2 = out.chunk(2, dim=0)
            pos1, pos2 = pos.chunk(2, dim=0)
        else:  # :294-295
            feat1, pos1 = feat2, pos2
            feat2, pos2 = encode_image(sd, frames[i + 1]["img"])
        feat_fuse = sp_mem.memory_read(feat_k2, res=True) if feat_k2 is not None else feat1
        dec1, dec2 = decoder(sd, feat_fuse, pos1, feat2, pos2)
        feat_k1 = key_head(sd, 1, feat1, dec1[-1])
        feat_k2 = key_head(sd, 2, feat2, dec2[-1])
        res1 = downstream_head(sd, "dust3r.downstream_head1", dec1, H, W)
        res2 = downstream_head(sd, "dust3r.downstream_head2", dec2, H, W)
        cur_v = encode_cur_value(sd, res1["pts3d"], mem_pos_enc)
        sp_mem.add_mem_check(feat_k1, cur_v + feat_k1)
        if trace is not None:
            trace.append(dict(feat1=feat1, feat2=feat2, feat_fuse=feat_fuse, dec1=dec1, dec2=dec2, feat_k1=feat_k1,
                              feat_k2=feat_k2, cur_v=cur_v))
        res2["pts3d_in_other_view"] = res2.pop("pts3d")
        if preds is None:
            preds = [res1]
            preds_all = [(res1, res2)]
        else:
            res1["pts3d_in_other_view"] = res1.pop("pts3d")
            preds.append(res1)
            preds_all.append((res1, res2))
    preds.append(res2)
    if return_memory:
        return preds, preds_all, sp_mem
    return preds, preds_all


# ------------------------------------------------------------------------------------------------
# offline mode: pairwise DUSt3R forward + best-view-first reconstruction
# ------------------------------------------------------------------------------------------------
@torch.no_grad()
def dust3r_forward(sd, view1, view2):
    """AsymmetricCroCo3DStereo.forward, dust3r/model.py:213-225 (the symmetrized-batch shortcut of :156-184 only
    skips redundant encoder work; the result per pair is the same)."""
    img1, img2 = view1["img"], view2["img"]
    B, _, H, W = img1.shape
    feats, pos = encode_image(sd, torch.cat((img1, img2), dim=0))
    dec1, dec2 = decoder(sd, feats[:B], pos[:B], feats[B:], pos[B:])
    res1 = downstream_head(sd, "dust3r.downstream_head1", dec1, H, W)
    res2 = downstream_head(sd, "dust3r.downstream_head2", dec2, H, W)
    res2["pts3d_in_other_view"] = res2.pop("pts3d")
    return res1, res2


def conf_score(conf):
    """mean of (conf-1)/conf, spann3r/model.py:346-352, 372-381."""
    return ((conf - 1) / conf).mean()


def find_initial_pair(graph, n_frames):
    """spann3r/model.py:333-357."""
    view1, view2, pred1, pred2 = graph["view1"], graph["view2"], graph["pred1"], graph["pred2"]
    conf_matrix = torch.zeros(n_frames, n_frames)
    for i in range(len(view1["idx"])):
        conf_matrix[view1["idx"][i], view2["idx"][i]] = conf_score(pred1["conf"][i]) + conf_score(pred2["conf"][i])
    flat = int(conf_matrix.argmax())
    return flat // n_frames, flat % n_frames


@torch.no_grad()
def offline_reconstruction(sd, frames, graph, **mem_kw):
    """Spann3R.offline_reconstruction, spann3r/model.py:394-471 (+ find_next_best_view :359-392), eval mode."""
    n_frames = len(frames)
    idx_todo = list(range(n_frames))
    H, W = frames[0]["img"].shape[-2:]
    sp_mem = SpatialMemory(sd, **mem_kw)
    p0, p1 = find_initial_pair(graph, n_frames)
    idx_used = [p0, p1]
    idx_todo.remove(p0)
    idx_todo.remove(p1)
    out, pos = encode_image(sd, torch.cat((frames[p0]["img"], frames[p1]["img"]), dim=0))
    feat1, feat2 = out.chunk(2, dim=0)
    pos1, pos2 = pos.chunk(2, dim=0)
    dec1, dec2 = decoder(sd, feat1, pos1, feat2, pos2)
    res1 = downstream_head(sd, "dust3r.downstream_head1", dec1, H, W)
    res2 = downstream_head(sd, "dust3r.downstream_head2", dec2, H, W)
    feat_k2, preds, preds_all = None, None, []
    while True:
        if feat_k2 is not None:
            feat1, pos1 = feat2, pos2
            feat_fuse = sp_mem.memory_read(feat_k2, res=True)
            best_conf, best = 0.0, None
            for i in idx_todo:   # find_next_best_view
                f2, ps2 = encode_image(sd, frames[i]["img"])
                d1, d2 = decoder(sd, feat_fuse, pos1, f2, ps2)
                r1 = downstream_head(sd, "dust3r.downstream_head1", d1, H, W)
                r2 = downstream_head(sd, "dust3r.downstream_head2", d2, H, W)
                total = conf_score(r1["conf"]) + conf_score(r2["conf"])
                if total > best_conf:
                    best_conf, best = total, (i, d1, d2, r1, r2, f2, ps2)
            id_n, dec1, dec2, res1, res2, feat2, pos2 = best
            idx_todo.remove(id_n)
            idx_used.append(id_n)
        feat_k1 = key_head(sd, 1, feat1, dec1[-1])
        feat_k2 = key_head(sd, 2, feat2, dec2[-1])
        cur_v = encode_cur_value(sd, res1["pts3d"])
        sp_mem.add_mem_check(feat_k1, cur_v + feat_k1)
        res2["pts3d_in_other_view"] = res2.pop("pts3d")
        if preds is None:
            preds = [res1]
            preds_all = [(res1, res2)]
        else:
            res1["pts3d_in_other_view"] = res1.pop("pts3d")
            preds.append(res1)
            preds_all.append((res1, res2))
        if len(idx_todo) == 0:
            break
    preds.append(res2)
    return preds, preds_all, idx_used
