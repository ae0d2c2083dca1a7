"""TEST / MEASUREMENT INFRASTRUCTURE ONLY - recipe that stages the UNMODIFIED reference modules of the hot path.

    python oracle/build_ref.py        # needs /root/reference (the build container); writes oracle/_ref/

The reference is pure Python: its implementation of the path is `Model.py` (UNet, Model.py:95-169) and `loss.py`
(calc_loss, loss.py:442-516), which import cleanly with this image's torch / scipy / cv2. /root/reference does not
exist on the GPU box, so this recipe copies those two files BYTE FOR BYTE into the git-ignored `oracle/_ref/` (it
travels with the gpurun snapshot like a built .so; nothing of it enters the history) together with a manifest of
their sha256. `oracle/ref_loader.py` imports them under private module names. Used by bench.py's CPU legs
(`cpu_baseline.kind = "reference"`), by scripts/ref_gpu_yardstick.py (the cuDNN context number) and by tests as the
checker. Never imported by the product package.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("B200UNET_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["Model.py", "loss.py", "Trainer.py"]  # Trainer.py: driven (unchanged) by the integration test only


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def build(verbose=False):
    """Stage the reference files. Returns the destination, or None when the reference is not mounted."""
    if not all(os.path.exists(os.path.join(REF, f)) for f in FILES):
        return DST if os.path.exists(os.path.join(DST, "MANIFEST.json")) else None
    os.makedirs(DST, exist_ok=True)
    manifest = {"source": REF, "files": {}}
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(DST, f)
        shutil.copyfile(src, dst)
        manifest["files"][f] = sha256(dst)
        assert manifest["files"][f] == sha256(src)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    if verbose:
        print(json.dumps(manifest, indent=1))
    return DST


if __name__ == "__main__":
    out = build(verbose=True)
    if out is None:
        print(f"{REF} is not mounted and oracle/_ref/ has not been staged", file=sys.stderr)
        sys.exit(1)
