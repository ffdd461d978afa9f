"""Copies the reference's UNet package (~3k lines of Python, read-only, unmodified) into the git-ignored
oracle/_ref/ so that `bench.py --impl reference` can time the reference's OWN code on the GPU box's host cores
(gpurun ships /root/repo only).  Never committed; never imported by the product."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def stage(src_root: str = "/root/reference") -> str:
    src = os.path.join(src_root, "avgen", "models", "unets")
    dst = os.path.join(HERE, "_ref", "avgen", "models", "unets")
    if not os.path.isdir(src):
        raise SystemExit(f"{src} not found")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__"))
    return dst


if __name__ == "__main__":
    print(stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
