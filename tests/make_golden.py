"""Regenerate tests/golden/*.npz from the unmodified reference
(oracle/_ref/ref_dump; build it with `make -C oracle`).  Needs /root/reference."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cases  # noqa: E402

if __name__ == "__main__":
    # `make_golden.py NAME ...` regenerates the named cases, no argument all of them
    for name in (sys.argv[1:] or cases.CASES):
        p = cases.write_golden(name)
        print(p, os.path.getsize(p))
