"""Generate tests/golden/ref_edge.pt from the UNMODIFIED reference (build container only).

    python oracle/make_golden_edge.py

test_mc3serousv5.py cannot be imported (it np.load()s a hard-coded path at import), so its
`preprocess` function is taken out of the file with `ast` and executed as it stands (numpy, scipy zoom and torch in
its globals; Tensor.cuda is made the identity because this container has no GPU). Inputs are seeded uint8 images of
the kinds cv2.imread returns (BGR colour, grey), including one with a constant channel (std = 0).
"""
import ast
import os
import sys

sys.dont_write_bytecode = True
import numpy as np
import torch
from scipy.ndimage import zoom

REF = os.environ.get("B200UNET_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_edge.pt")


def reference_preprocess():
    src = open(os.path.join(REF, "test_mc3serousv5.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "preprocess")
    ns = {"np": np, "torch": torch, "zoom": zoom}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "test_mc3serousv5.py", "exec"), ns)
    return ns["preprocess"]


def reference_preprocess_crop():
    """`preprocessCrop` of test.py:91-126 (pad to a multiple of crop_size with 255, z-normalise the padded image)."""
    src = open(os.path.join(REF, "test.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "preprocessCrop")
    ns = {"np": np, "torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "test.py", "exec"), ns)
    return ns["preprocessCrop"]


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self  # no GPU here; the function ends in .cuda()
    pre = reference_preprocess()
    rng = np.random.default_rng(2024)
    cases = {}

    def smooth(h, w, c):
        base = rng.normal(size=(h // 4 + 1, w // 4 + 1, c)).repeat(4, 0).repeat(4, 1)[:h, :w]
        img = 128 + 50 * base + rng.normal(scale=12, size=(h, w, c)) + np.array([20, -10, 5, 0][:c])
        return np.clip(img, 0, 255).astype(np.uint8)

    cases["bgr_64x96"] = smooth(64, 96, 3)
    cases["bgr_50x70_ragged"] = smooth(50, 70, 3)          # H*W not a multiple of 16: scalar kernel path
    cases["grey_48x48"] = smooth(48, 48, 1)[:, :, 0]
    full = rng.integers(0, 256, size=(32, 32, 3), dtype=np.uint8)
    cases["bgr_uniform_noise"] = full
    const = smooth(32, 48, 3)
    const[:, :, 1] = 77                                    # std = 0 -> nan in that channel, like the reference
    cases["bgr_constant_channel"] = const
    out = {}
    with np.errstate(all="ignore"):
        for k, img in cases.items():
            out[k] = dict(img=torch.from_numpy(img.copy()), out=pre(img, img.shape[:2]).clone())
    # tiled inference front end (test.py:91-126): image, crop size -> padded z-normalised tensor (+ padded label)
    crop = reference_preprocess_crop()
    for k, (img, cs) in {"crop_bgr_50x70_c32": (cases["bgr_50x70_ragged"], 32), "crop_bgr_64x96_c32": (cases["bgr_64x96"], 32),
                         "crop_bgr_50x70_c48": (cases["bgr_50x70_ragged"], 48)}.items():
        lab = (rng.random(img.shape[:2]) > 0.5).astype(np.uint8)
        x, lab_p, _ = crop(img, lab, lab.copy(), cs)
        out[k] = dict(img=torch.from_numpy(img.copy()), crop=cs, out=x.clone(), label_shape=tuple(lab_p.shape))
    torch.save(out, OUT)
    print(OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
