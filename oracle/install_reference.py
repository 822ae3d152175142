"""TEST / BASELINE INFRASTRUCTURE ONLY -- copies the UNMODIFIED reference tree into baseline/_ref/.

The reference (crockwell/rel_pose @35d1352) is pure Python with no setup.py / pyproject.toml, so the
`pip install --target baseline/_ref /root/reference` of the bench contract cannot work (pip: "neither
'setup.py' nor 'pyproject.toml' found"); a byte-for-byte copy is the equivalent install.  baseline/_ref
is git-ignored (never in history) but NOT gpurun-ignored: it travels to the GPU box, where /root/reference
does not exist, so that `bench.py --impl reference`, the `gpu_eager_baseline` leg and the config-1
`demo.py` run can execute the real reference there (oracle/ref_loader.py finds it).

    python oracle/install_reference.py            # no-op when /root/reference is absent
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("RELPOSE_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def install(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "src")):
        if verbose:
            print(f"install_reference: {SRC} not present, nothing to do")
        return False
    if os.path.isdir(DST):
        cmp = filecmp.dircmp(SRC, DST)
        if not (cmp.left_only or cmp.diff_files or cmp.funny_files) and os.path.isfile(os.path.join(DST, "src", "model.py")):
            return True
        shutil.rmtree(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for dp, _, fs in os.walk(DST):          # the source tree is read-only; make the copy removable
        os.chmod(dp, 0o755)
        for f in fs:
            os.chmod(os.path.join(dp, f), 0o644)
    if verbose:
        print(f"install_reference: copied {SRC} -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() or True else 1)
