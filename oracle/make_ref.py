"""Recipe for oracle/_ref: a verbatim, git-ignored copy of the reference's own Python sources (``/root/reference/src``)
so that the UNMODIFIED reference train step (src/train.py:175-203) can be timed as the CPU arm on the GPU box, where
/root/reference does not exist (bench.py --impl reference, bench.py's cpu_baseline leg).

TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under oracle/ is imported by the product package; oracle/_ref/ is listed
in .gitignore (never committed) but travels with the gpurun snapshot like the built .so does.

    python oracle/make_ref.py            # run in the build container; __graft_entry__.build() calls it when possible

Nothing is edited: files are copied byte for byte.  The two imports the reference makes that are absent in this image
(skimage, matplotlib -- used only by its plotting / PSNR helpers) are satisfied at import time by empty
stand-in modules registered by ``import_reference()`` below; ``torch.cuda.empty_cache`` is a no-op on CPU.
"""
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
DST = os.path.join(HERE, "_ref")


def make(verbose: bool = True) -> bool:
    if not os.path.isdir(REF_SRC):
        return os.path.isdir(os.path.join(DST, "src"))
    os.makedirs(os.path.join(DST, "src"), exist_ok=True)
    n = 0
    for name in sorted(os.listdir(REF_SRC)):
        if name.endswith(".py"):
            shutil.copyfile(os.path.join(REF_SRC, name), os.path.join(DST, "src", name))
            n += 1
    if verbose:
        print(f"oracle/_ref: {n} reference source files copied from {REF_SRC}")
    return True


def available() -> bool:
    return os.path.isfile(os.path.join(DST, "src", "train.py")) and os.path.isfile(os.path.join(DST, "src", "models.py"))


def import_reference():
    """(models, train, utils) of the unmodified reference, imported from oracle/_ref (or /root/reference when present)."""
    import torch
    for name in ("skimage", "skimage.metrics", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage.metrics"].structural_similarity = None
    sys.modules["skimage.metrics"].peak_signal_noise_ratio = None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    root = DST if available() else "/root/reference"
    if root not in sys.path:
        sys.path.insert(0, root)
    if not torch.cuda.is_available():
        torch.cuda.empty_cache = lambda: None
    import src.models as models
    import src.train as train
    import src.utils as utils
    return models, train, utils


if __name__ == "__main__":
    ok = make()
    sys.exit(0 if ok else 1)
