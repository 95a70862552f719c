"""Stage the reference's Python sources for ONE purpose: letting the GPU box run the reference's unmodified train
loop against the hot path (tests/test_gpu_training_boundary.py) and time the reference's stock eager GPU path
(bench.py --extras gpu_reference).  /root/reference does not travel with `gpurun`; baseline/_ref/ does (it is
git-ignored, so the history stays free of reference sources, and not gpurun-ignored).

    python tools/stage_reference.py          # copies /root/reference/AGQA/src/**/*.py -> baseline/_ref/AGQA/src/

Nothing in the product path, smoke() or the default bench line reads the copy; tests that need it skip without it.
"""
import os
import shutil
import sys

SRC = "/root/reference/AGQA"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "AGQA")


def main():
    if not os.path.isdir(os.path.join(SRC, "src")):
        print("no reference at %s: nothing staged" % SRC)
        return 1
    n = 0
    for root, _, files in os.walk(os.path.join(SRC, "src")):
        for f in files:
            if not f.endswith(".py"):
                continue
            src = os.path.join(root, f)
            dst = os.path.join(DST, os.path.relpath(src, SRC))
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            n += 1
    print("staged %d files under %s" % (n, DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
