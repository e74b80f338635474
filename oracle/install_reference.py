"""Recipe for ``baseline/_ref``: the UNMODIFIED reference package, so that it travels to the GPU box.

    python oracle/install_reference.py            (build container only: needs /root/reference)

``/root/reference`` does not exist on the GPU box; ``baseline/_ref`` is git-ignored (no reference source enters the
history) but NOT gpurun-ignored, so what this script puts there ships with every ``gpurun`` snapshot.  It first tries
the contract's offline pip install; the reference's build backend is poetry-core, which is not in the wheelhouse, so
pip fails here ("No module named 'poetry'") and the script falls back to what a wheel of a pure-Python package
amounts to: a verbatim copy of the ``beetroots/`` package directory (0.9 MB, no compiled code).  Nothing is edited;
``oracle/refshim.install()`` puts the three import-time shims of SURVEY.md App. A around it at run time.

Test / benchmark infrastructure only: ``bench.py --impl reference`` times it, ``tests/test_gpu_reference.py`` drives
``B200Sampler`` with its ``Posterior`` and ``MySaver``.  The product path never imports it.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def install(verbose: bool = True) -> str:
    if not os.path.isdir(os.path.join(SRC, "beetroots")):
        raise RuntimeError("/root/reference is only present in the build container")
    if os.path.isdir(os.path.join(DST, "beetroots")):
        return "present"
    os.makedirs(DST, exist_ok=True)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
           "/opt/wheelhouse", "--target", DST, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode == 0 and os.path.isdir(os.path.join(DST, "beetroots")):
        how = "pip"
    else:
        if verbose:
            last = (r.stderr.strip().splitlines() or ["?"])[-1]
            print(f"pip install failed ({last}); copying the pure-Python package tree instead")
        shutil.copytree(os.path.join(SRC, "beetroots"), os.path.join(DST, "beetroots"),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        how = "copy"
    with open(os.path.join(DST, "INSTALL_NOTE.txt"), "w") as f:
        f.write(f"unmodified copy of /root/reference/beetroots made by oracle/install_reference.py ({how}); git-ignored\n")
    return how


if __name__ == "__main__":
    print("baseline/_ref:", install())
