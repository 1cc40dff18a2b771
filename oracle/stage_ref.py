"""Stage the UNMODIFIED reference sources of the hot path under oracle/_ref/ (TEST / BENCH INFRASTRUCTURE ONLY).

The reference (BoifZ/VDN-NeRF) is pure Python: its hot path is three files, dpt_models/{embedder,fields,renderer}.py,
which import nothing but torch / numpy - plus `mcubes` and `icecream.ic` at the top of renderer.py, neither used by any
rendering arithmetic.  /root/reference does not exist on the GPU box, so this recipe copies the
three files byte for byte into oracle/_ref/dpt_models/ and writes the two import stubs next to them; oracle/_ref/ is
git-ignored (the reference's sources never enter the history) but travels to the GPU box with the snapshot, like the
built .so.  `bench.py --impl reference` and the `reference_cuda_eager` leg import the reference from there.

    python -m oracle.stage_ref            # run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref")
FILES = ("embedder.py", "fields.py", "renderer.py")


def stage(ref: str = REF, dst: str = DST) -> bool:
    """Copy the reference files; returns False when the reference checkout is not available (GPU box)."""
    src_dir = os.path.join(ref, "dpt_models")
    if not all(os.path.exists(os.path.join(src_dir, f)) for f in FILES):
        return False
    out = os.path.join(dst, "dpt_models")
    os.makedirs(out, exist_ok=True)
    sums = []
    for f in FILES:
        shutil.copyfile(os.path.join(src_dir, f), os.path.join(out, f))
        sums.append("%s  dpt_models/%s" % (hashlib.sha256(open(os.path.join(out, f), "rb").read()).hexdigest(), f))
    open(os.path.join(out, "__init__.py"), "w").close()
    with open(os.path.join(dst, "SHA256SUMS"), "w") as fh:
        fh.write("\n".join(sums) + "\n")
    return True


def available(dst: str = DST) -> bool:
    return all(os.path.exists(os.path.join(dst, "dpt_models", f)) for f in FILES)


def import_reference(dst: str = DST):
    """(fields, renderer, embedder) modules of the staged reference, with mcubes / icecream stubbed."""
    if not available(dst):
        raise ImportError("oracle/_ref is not staged; run `python -m oracle.stage_ref` where /root/reference exists")
    sys.modules.setdefault("mcubes", types.ModuleType("mcubes"))
    if "icecream" not in sys.modules:
        ic = types.ModuleType("icecream")
        ic.ic = lambda *a, **k: None
        sys.modules["icecream"] = ic
    if dst not in sys.path:
        sys.path.insert(0, dst)
    import dpt_models.embedder as re_
    import dpt_models.fields as rf
    import dpt_models.renderer as rr
    return rf, rr, re_


if __name__ == "__main__":
    print("staged" if stage() else "reference checkout not found; nothing staged")
