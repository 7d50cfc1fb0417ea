"""TEST INFRASTRUCTURE ONLY -- stage the UNMODIFIED reference package for the GPU box.

``/root/reference`` exists only in the build container.  ``stage()`` (called by
``__graft_entry__.build()``) copies the reference's Python package ``ReTiSAR/*.py`` -- byte for
byte, nothing edited -- and its LICENSE into ``oracle/_ref/`` so that the GPU box can

  * time the reference's own ``AdjustableShConvolver.filter_block`` as the CPU arm of
    ``bench.py --impl reference`` (``cpu_baseline.kind = "reference"``), and
  * run the drop-in test that drives the reference's ``JackRenderer`` / ``JackRendererBenchmark``
    with this repository's convolver classes swapped in (``tests/test_gpu_dropin.py``).

``oracle/_ref/`` is git-ignored (the reference's sources never enter this repository's history)
but NOT gpurun-ignored, so it travels with the snapshot like the built ``.so``.  The package is
imported through ``oracle/ref_harness.py`` (third-party modules absent here are mocked there).
Nothing under ``retisar_b200/`` imports it.
"""
import hashlib
import os
import shutil

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def staged_root():
    """Directory holding a staged copy (``<root>/ReTiSAR``), or None."""
    return DST if os.path.isfile(os.path.join(DST, "ReTiSAR", "_convolver.py")) else None


def stage(verbose=False):
    """Copy ``/root/reference/ReTiSAR/*.py`` into ``oracle/_ref/ReTiSAR`` when the reference is
    present; keeps an existing staged copy otherwise.  Returns the staged root or None."""
    src_pkg = os.path.join(SRC, "ReTiSAR")
    if not os.path.isdir(src_pkg):
        return staged_root()
    dst_pkg = os.path.join(DST, "ReTiSAR")
    os.makedirs(dst_pkg, exist_ok=True)
    digest = hashlib.sha256()
    for name in sorted(os.listdir(src_pkg)):
        if not name.endswith(".py"):
            continue
        shutil.copyfile(os.path.join(src_pkg, name), os.path.join(dst_pkg, name))
        with open(os.path.join(src_pkg, name), "rb") as f:
            digest.update(name.encode() + b"\0" + f.read())
    for name in ("LICENSE", "AUTHORS"):
        if os.path.isfile(os.path.join(SRC, name)):
            shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write(f"{SRC}/ReTiSAR/*.py, unmodified; sha256 over (name, bytes) = {digest.hexdigest()}\n")
    if verbose:
        print(f"staged the reference package into {DST}")
    return DST


if __name__ == "__main__":
    stage(verbose=True)
