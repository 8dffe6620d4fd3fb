"""TEST INFRASTRUCTURE ONLY — stages the UNMODIFIED reference package into `oracle/_ref/`.

`/root/reference` exists only in the build container; the GPU box gets a snapshot of this repository. So that
`bench.py --impl reference` and the `cpu_baseline` leg can time the reference's OWN code there (kind "reference"),
`__graft_entry__.build()` calls `stage()` here: an offline `pip install --no-deps --target oracle/_ref` of the
reference tree (from a temporary copy, `/root/reference` is read-only). `oracle/_ref/` is git-ignored (no reference
source ever enters the history) but not gpurun-ignored, so it travels with the snapshot like the built `.so`.
Nothing under `torch-fem_b200/` may import it.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("TFEM_REFERENCE_ROOT", "/root/reference")


def staged() -> bool:
    return os.path.isfile(os.path.join(DEST, "torchfem", "__init__.py"))


def stage(force: bool = False) -> str:
    """Returns 'staged', 'present' or 'unavailable: <why>'."""
    if staged() and not force:
        return "present"
    if not os.path.isdir(os.path.join(SOURCE, "src", "torchfem")):
        return f"unavailable: {SOURCE} not found (GPU box: uses the prebuilt oracle/_ref of the snapshot)"
    tmp = tempfile.mkdtemp(prefix="tfem_ref_")
    try:
        copy = os.path.join(tmp, "reference")
        shutil.copytree(SOURCE, copy, ignore=shutil.ignore_patterns(".git", "docs", "examples", "benchmarks"))
        shutil.rmtree(DEST, ignore_errors=True)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", DEST, copy]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or not staged():
            # the wheel build failed: the package is pure Python, take the package directory as it is
            shutil.rmtree(DEST, ignore_errors=True)
            shutil.copytree(os.path.join(SOURCE, "src", "torchfem"), os.path.join(DEST, "torchfem"))
        return "staged"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
