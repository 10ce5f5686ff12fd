"""ORACLE SUPPORT (test infrastructure): byte-compile the UNMODIFIED reference modules of the hot path into
`oracle/_ref/` so that the reference's OWN classes (UNet, DropBlock2D, DropBlockEval, RotationEval, UNetTraining) can be
imported on the GPU box, where /root/reference does not exist.

Recipe (run by `__graft_entry__.build()` whenever /root/reference is present): `py_compile` of the files listed below,
read where they lie under /root/reference, outputs ONLY under `oracle/_ref/unet_code/...` as CPython byte-code files
(`<module>.bytecode` = the .pyc format; loaded by `oracle/ref_shims.py` with `SourcelessFileLoader`; the `.pyc` suffix is
avoided because snapshot tools commonly drop it).  No reference SOURCE is copied: `oracle/_ref/` holds compiled
artefacts, is git-ignored, and travels to the GPU box like the repo's own built `.so`.  The bytecode is tied to the
interpreter that built it (same image here and on the box: CPython 3.12); a mismatch makes the import fail loudly.
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CODE = "/root/reference/Unet_research/unet_code"
OUT_CODE = os.path.join(HERE, "_ref", "unet_code")
EXT = ".bytecode"

# the files on the hot path plus the helper modules they import at module level
FILES = [
    "utils/utils_unet.py", "utils/utils_modules.py", "utils/utils_training.py", "utils/utils_general.py",
    "utils/utils_dataset.py", "utils/utils_metrics.py",
    "uncertainty_tests/Dropblock_Uncertainty.py", "uncertainty_tests/Rotational_Uncertainty.py",
    "base_model_tests/training.py",
]


def ref_available() -> bool:
    return os.path.isdir(REF_CODE)


def built() -> bool:
    return all(os.path.exists(os.path.join(OUT_CODE, f[:-3] + EXT)) for f in FILES)


def build_ref(force: bool = False) -> bool:
    """Returns True when oracle/_ref is complete afterwards."""
    if not ref_available():
        return built()
    for rel in FILES:
        src = os.path.join(REF_CODE, rel)
        dst = os.path.join(OUT_CODE, rel[:-3] + EXT)
        if not force and os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: the path recorded in tracebacks points at the original location, not at a file in this repo
        py_compile.compile(src, cfile=dst, dfile=src, doraise=True, optimize=0,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(HERE, "_ref", "PYTHON_VERSION"), "w") as f:
        f.write("%d.%d\n" % sys.version_info[:2])
    return built()


if __name__ == "__main__":
    ok = build_ref(force="--force" in sys.argv)
    print("oracle/_ref", "complete" if ok else "NOT built (reference tree absent)")
