"""`python -m rel_pose_b200.run <reference script> [its arguments]` -- runs a reference script unchanged
with rel_pose_b200 registered as `src.model` / `lietorch` (see dropin.py).  The script's own directory is put
first on sys.path exactly as `python script.py` would."""
import os
import runpy
import sys


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = os.path.abspath(sys.argv[1])
    sys.argv = [script] + sys.argv[2:]
    sys.path.insert(0, os.path.dirname(script))
    from . import dropin
    dropin.install()
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
