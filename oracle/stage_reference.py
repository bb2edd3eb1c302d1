"""Stage the reference's own Python sources into the git-ignored `oracle/_ref/` (TEST INFRASTRUCTURE).

    python oracle/stage_reference.py            # in the build container, where /root/reference exists

The reference (hcnoh/object-detection-collection-pytorch) is pure Python: there is nothing to compile, and
`/root/reference` does not exist on the GPU box.  `oracle/_ref/` is listed in .gitignore (so no reference source
ever enters the history) but not in .gpurunignore (so the staged copy travels to the GPU box with the snapshot,
like the built .so files).  It lets the GPU-box legs run the UNMODIFIED reference:
  * `bench.py --impl reference`  times the reference's get_loss + backward + per-image nms on the host cores
    (`cpu_baseline.kind: "reference"`; without the staged copy the oracle's port of the same op sequence is timed);
  * `tests/test_gpu_reference_model.py` runs the reference's real YOLOv2 (Darknet19 backbone) through one
    run_one_epoch batch with and without the drop-in patch.
Only `oracle/refharness.py`, `tests/` and `bench.py`'s reference legs import from it; the product never does.
`__graft_entry__.build()` calls this when /root/reference is present.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["config.py", "LICENSE", "models/utils.py", "models/yolov1.py", "models/yolov2.py",
         "models/backbones/darknet19.py", "models/backbones/darknet53.py", "models/backbones/googlenet.py"]


def stage(src="/root/reference", dest=DEST, quiet=False):
    """Copy the files of the path (and the model code around it) verbatim; returns the list staged."""
    if not os.path.isdir(src):
        return []
    done = []
    for rel in FILES:
        a, b = os.path.join(src, rel), os.path.join(dest, rel)
        if not os.path.exists(a):
            continue
        os.makedirs(os.path.dirname(b), exist_ok=True)
        shutil.copyfile(a, b)
        done.append(rel)
    with open(os.path.join(dest, "STAGED_FROM"), "w") as f:
        f.write("verbatim copy of %s (staged by oracle/stage_reference.py; git-ignored test infrastructure)\n" % src)
    if not quiet:
        print("staged %d reference files into %s" % (len(done), dest))
    return done


if __name__ == "__main__":
    sys.exit(0 if stage(*(sys.argv[1:2] or ["/root/reference"])) else 1)
