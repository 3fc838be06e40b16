"""TEST INFRASTRUCTURE — makes the reference's own hot-path files available on the GPU box.

The reference is pure Python: there is nothing to compile.  When /root/reference is mounted
(authoring container) this copies its four GPSLayer hot-path files *verbatim* into the git-ignored
directory oracle/_ref/ (listed in .gitignore, not in .gpurunignore, so it travels with gpurun but
never enters history).  oracle/ref_shim.py then runs them unmodified on the box's host cores as
the `--impl reference` / cpu_baseline "reference" arm of bench.py.  Nothing is copied into tracked
files.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/graphgps/layer"
FILES = ["performer_layer.py", "gatedgcn_layer.py", "gine_conv_layer.py", "gps_layer.py"]


def build_ref() -> bool:
    if not all(os.path.isfile(os.path.join(SRC, f)) for f in FILES):
        return os.path.isdir(os.path.join(HERE, "_ref"))
    dst = os.path.join(HERE, "_ref")
    os.makedirs(dst, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(dst, f))
    return True


if __name__ == "__main__":
    print("oracle/_ref ready:", build_ref())
