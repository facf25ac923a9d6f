"""Stage the reference's PYTHON hot-path files under the git-ignored baseline/_ref/ so that they travel to the
GPU box with the working-tree snapshot (like the compiled oracle/_ref natives): there the reference's own
`UniformProjection`, `frnn.frnn_grid_points` and `EllipticalRasterizer` run on the B200 on top of its own
recompiled CUDA extensions -- the reference-GPU arm of bench.py (`ref_cuda`) and of the drop-in tests.

TEST / BENCH INFRASTRUCTURE ONLY.  Nothing is copied into tracked paths; baseline/_ref/ is listed in .gitignore.

    python -m oracle.stage_ref
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.environ.get("ISO_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def stage():
    if not os.path.isdir(os.path.join(SRC, "DSS")):
        print("reference tree not present; keeping baseline/_ref as is")
        return None
    n = 0
    for sub in ("DSS", os.path.join("external", "FRNN", "frnn")):
        for d, _, files in os.walk(os.path.join(SRC, sub)):
            if "csrc" in d.split(os.sep):
                continue
            for f in files:
                if f.endswith(".py"):
                    rel = os.path.relpath(os.path.join(d, f), SRC)
                    out = os.path.join(DST, rel)
                    os.makedirs(os.path.dirname(out), exist_ok=True)
                    shutil.copyfile(os.path.join(d, f), out)
                    n += 1
    print("staged %d reference .py files under %s" % (n, DST))
    return DST


if __name__ == "__main__":
    stage()
