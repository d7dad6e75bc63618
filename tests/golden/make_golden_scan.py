#!/usr/bin/env python
"""Golden stdout of the reference's lookup scanners (triobin / trioeval / chkerr / sexchr) from the UNMODIFIED
reference binary (oracle/_ref/yak, built from /root/reference by oracle/Makefile) on the seeded inputs of
tests/scan_inputs.py.  Run in the build container:   python tests/golden/make_golden_scan.py"""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
import scan_inputs as S  # noqa: E402


def main():
    oracle_lib.build()
    tmp = tempfile.mkdtemp()
    paths = S.write_all(tmp)
    for y, (fa, k) in S.COUNTS.items():
        paths[y] = os.path.join(tmp, y)
        subprocess.run([oracle_lib.REF_YAK, "count", f"-k{k}", "-p10", "-t4", "-o", paths[y], paths[fa]], check=True, capture_output=True)
    for gold, cmd in S.CASES:
        argv = [oracle_lib.REF_YAK] + S.argv(cmd, paths)
        r = subprocess.run(argv, check=True, capture_output=True)
        open(os.path.join(HERE, gold), "wb").write(r.stdout)
        print(gold, len(r.stdout), "bytes,", r.stdout.count(b"\n"), "lines")


if __name__ == "__main__":
    main()
