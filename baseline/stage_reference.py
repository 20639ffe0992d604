#!/usr/bin/env python
"""Stage the UNMODIFIED reference package under baseline/_ref/ (git-ignored; travels to the GPU box with the tree).

The contract's `pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref
/root/reference` fails in this image (build backend `hatchling` is not installed, no network), so the package
directory -- pure Python plus its data files -- is copied as it lies: src/silero_vad/{__init__,model,utils_vad}.py and
data/silero_vad.jit (the default model of load_silero_vad(), /root/reference/src/silero_vad/model.py:17,34).
Nothing under baseline/_ref/ is tracked or edited.  bench.py --impl reference and the GPU arm's `cpu_baseline`
import it from there at run time (never from /root/reference, which does not exist on the GPU box).
"""
import shutil
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
SRC = Path("/root/reference/src/silero_vad")
DST = REPO / "baseline" / "_ref" / "silero_vad"


def main():
    if not SRC.exists():
        print("reference tree not present; nothing staged")
        return 0 if DST.exists() else 1
    if DST.exists():
        shutil.rmtree(DST)
    (DST / "data").mkdir(parents=True)
    for f in ("__init__.py", "model.py", "utils_vad.py"):
        shutil.copy2(SRC / f, DST / f)
    shutil.copy2(SRC / "data" / "__init__.py", DST / "data" / "__init__.py")
    shutil.copy2(SRC / "data" / "silero_vad.jit", DST / "data" / "silero_vad.jit")
    print("staged", DST, sum(p.stat().st_size for p in DST.rglob("*") if p.is_file()), "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
