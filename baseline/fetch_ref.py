"""Install the UNMODIFIED reference (PFLeget/cosmogp) into baseline/_ref (git-ignored, shipped to the GPU box).

    python baseline/fetch_ref.py

The contract's offline install: pip needs to write build/ and egg-info into the source tree and /root/reference is
read-only, so the install runs from a copy under /tmp; dependency resolution is skipped (--no-deps: numpy / scipy
are in the image, matplotlib -- plotting only -- is not).  The sources land verbatim in baseline/_ref/cosmogp; they
are Python 2 (three `print` statements), which oracle/ref_loader.py rewrites IN MEMORY when it imports them.
Nothing under baseline/_ref is tracked by git and the product never imports it: it is the CPU arm of bench.py
(--impl reference, cpu_baseline) and nothing else.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("COSMOGP_REFERENCE_SOURCE", "/root/reference")


def installed():
    return os.path.isfile(os.path.join(TARGET, "cosmogp", "Gaussian_process.py"))


def fetch(force=False):
    if installed() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(SOURCE, "cosmogp")):
        raise RuntimeError("reference tree not found at %s" % SOURCE)
    tmp = tempfile.mkdtemp(prefix="cosmogp_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(SOURCE, src)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        subprocess.check_call([sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation",
                               "--no-deps", "--no-compile", "--find-links", "/opt/wheelhouse", "--target", TARGET, src])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    if not installed():
        raise RuntimeError("pip reported success but %s/cosmogp is missing" % TARGET)
    return TARGET


if __name__ == "__main__":
    print(fetch(force="--force" in sys.argv))
