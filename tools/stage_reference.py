#!/usr/bin/env python
"""Stage the UNMODIFIED reference's forward path under the git-ignored `baseline/_ref/` so that it travels to the GPU
box (`/root/reference` does not exist there) and `bench.py` can time the reference ITSELF -- `spann3r.model.Spann3R`
(`/root/reference/spann3r/model.py:473-539`) -- as the CPU arm (`--impl reference`) and as the PyTorch-eager peer on
the same B200.

    python tools/stage_reference.py              # copy the ~20 pure-Python files of the path (byte-identical)
    python tools/stage_reference.py --curope     # also build the reference's own CUDA RoPE extension for sm_100

Nothing is copied into tracked paths: `baseline/_ref/` and `baseline/_ref_curope/` are listed in .gitignore (and NOT
in .gpurunignore).  `__graft_entry__.build()` calls this when /root/reference is present; on the GPU box only the staged
copy exists.

--curope: the reference's extension does not compile against this torch as shipped (`kernels.cu:101` uses the removed
`Tensor::type()` dispatch argument).  To give the reference its best shot a COPY of the four curope
source files is made under `baseline/_ref_curope/`, that one token is replaced (`tokens.type()` ->
`tokens.scalar_type()`), and the extension is built there for sm_100 with the reference's own flags
(`-O3 --use_fast_math`, setup.py:24-27).  The staged reference picks it up when `baseline/_ref_curope` is on sys.path
(`import curope`, curope2d.py:6-9); otherwise it falls back to its PyTorch RoPE2D exactly as /root/reference does here.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(REPO, "baseline", "_ref")
DST_CUROPE = os.path.join(REPO, "baseline", "_ref_curope")

# the import closure of spann3r.model (forward path) + what the training step of config 5 needs on top of it
FILES = [
    "spann3r/model.py",
    "spann3r/loss.py",
    "dust3r/__init__.py", "dust3r/model.py", "dust3r/patch_embed.py", "dust3r/losses.py", "dust3r/inference.py",
    "dust3r/heads/__init__.py", "dust3r/heads/dpt_head.py", "dust3r/heads/linear_head.py", "dust3r/heads/postprocess.py",
    "dust3r/utils/__init__.py", "dust3r/utils/misc.py", "dust3r/utils/path_to_croco.py", "dust3r/utils/geometry.py",
    "dust3r/utils/device.py",
    "croco/models/blocks.py", "croco/models/croco.py", "croco/models/pos_embed.py", "croco/models/dpt_block.py",
    "croco/models/masking.py",
    "croco/models/curope/__init__.py", "croco/models/curope/curope2d.py",
]
CUROPE_SRC = ["curope.cpp", "kernels.cu"]


def stage(verbose=True) -> dict:
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found: the staged copy under baseline/_ref is all there is on this machine")
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            if verbose:
                print("skip (absent in the reference):", rel)
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()[:16]
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1)
    if verbose:
        print(f"staged {len(manifest)} files under {DST}")
    return manifest


def build_curope(verbose=True) -> str:
    os.makedirs(DST_CUROPE, exist_ok=True)
    for name in CUROPE_SRC:
        s = open(os.path.join(REF, "croco/models/curope", name)).read()
        if name == "kernels.cu":
            assert s.count("tokens.type()") == 1
            s = s.replace("tokens.type()", "tokens.scalar_type()")   # the ONE token (file docstring)
        open(os.path.join(DST_CUROPE, name), "w").write(s)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils.cpp_extension import load
    load(name="curope", sources=[os.path.join(DST_CUROPE, n) for n in CUROPE_SRC], build_directory=DST_CUROPE,
         extra_cflags=["-O3"], extra_cuda_cflags=["-O3", "--use_fast_math"], verbose=verbose, is_python_module=False)
    so = os.path.join(DST_CUROPE, "curope.so")
    if verbose:
        print("built", so)
    return so


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--curope", action="store_true")
    a = ap.parse_args()
    stage()
    if a.curope:
        build_curope()
    sys.exit(0)
