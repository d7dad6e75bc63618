"""The CLI's host code (yak_b200/cli/main.c, inspect_logic.c) run on the CPU: linked, in a throw-away executable, against
tests/mock_yak_api.c - the yak.h entry points answered by the oracle - and compared with the UNMODIFIED reference binary on the
same command lines.  What is under test is the command flow: options and their defaults, the two-pass protocol of `count -b`,
cntasm's shrink / setcnt / merge schedule, subtract / isec / recount, one- and two-file inspect.  (The product never links the
oracle: the real CLI binds libyakb200.so, and tests/test_setops.py / test_gpu_golden.py run that one on a GPU.)"""
import os
import subprocess

import numpy as np
import pytest

import golden_util as G
import oracle_lib as O
import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.exists(O.REF_YAK), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def exe():
    O.build()
    out = os.path.join(util.TMP, "yakb_cli_mock")
    cli = os.path.join(ROOT, "yak_b200", "cli")
    subprocess.run(["gcc", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), "-o", out,
                    os.path.join(cli, "main.c"), os.path.join(cli, "qv_solve.c"), os.path.join(cli, "inspect_logic.c"),
                    os.path.join(cli, "scan.c"), os.path.join(cli, "scan_logic.c"),
                    os.path.join(ROOT, "tests", "mock_yak_api.c"), "-L" + O.ORACLE_DIR, "-loracle", "-Wl,-rpath," + O.ORACLE_DIR,
                    # the record reader of the scanners is the library's own host code (no device involved); every yak.h symbol
                    # resolves to the mock, which the linker sees first
                    "-L" + os.path.join(ROOT, "yak_b200", "lib"), "-lyakb200", "-Wl,-rpath," + os.path.join(ROOT, "yak_b200", "lib"), "-lm", "-lz"], check=True)
    return out


def _both(exe, args, out_flag=True):
    """run the command on both binaries; returns (ours, reference) as (returncode, stdout, output file bytes)"""
    res = []
    for b, tag in ((exe, "m"), (O.REF_YAK, "r")):
        o = os.path.join(util.TMP, f"yakb_flow_{tag}.yak")
        if os.path.exists(o):
            os.unlink(o)
        r = subprocess.run([b, args[0]] + (["-o", o] if out_flag else []) + args[1:], capture_output=True)
        res.append((r.returncode, r.stdout, open(o, "rb").read() if os.path.exists(o) else None))
    return res


def test_count_command_lines(exe):
    fa, fq, fc = G.input_path("reads_a"), G.input_path("reads_q"), G.input_path("reads_c")
    for args in (["count", fa], ["count", "-k27", "-p11", "-t3", fa], ["count", "-b22", fq], ["count", "-b", "23", "-H", "3", "-K", "2m", fq, fc],
                 ["count", "-p12", "-b12", fa],            # -b not above -p: two passes without a filter (main.c:54)
                 ["count", "-k", "33", fa], ["count", fq, "-k21", "-b21"]):   # options behind the file name
        m, r = _both(exe, args)
        assert m[0] == r[0] == 0 and m[2] == r[2] and len(r[2]) > 10000, args
    for args in (["count", "-p9", fa], ["count", "-k64", fa], ["count"]):
        m, r = _both(exe, args)
        assert m[0] == r[0] == 1 and m[2] is None and r[2] is None, args


def test_cntasm_command_lines(exe):
    import test_setops as S
    fns = S._cntasm_inputs(4)
    prev = os.path.join(util.TMP, "yakb_flow_prev.yak")
    subprocess.run([O.REF_YAK, "cntasm", "-p10", "-o", prev] + fns[:2], check=True, capture_output=True)
    for args in (["-p10"] + fns[:3], ["-p10", "-c1", "-x3", "-e1", "-s2"] + fns, ["-k27", "-p11", "-c1", "-x2", "-r", "-s1"] + fns[:3],
                 ["-p10", "-c2", "-x1023", "-e2", "-s1", "-i", prev] + fns, ["-p10", "-e", "3", "-s", "3", "-t", "2", "-K", "1m"] + fns,
                 ["-p10", "-i", "/nonexistent.yak"] + fns[:2]):          # main.c:142: a warning, then on without it
        m, r = _both(exe, ["cntasm"] + args)
        assert m[0] == r[0] == 0 and m[2] == r[2] and len(r[2]) > 16 + 8 * 1024, args
    for args in (["-p9"] + fns[:1], ["-k32"] + fns[:1], []):
        m, r = _both(exe, ["cntasm"] + args)
        assert m[0] == r[0] == 1, args


def test_setop_and_inspect_command_lines(exe):
    import test_setops as S
    ya, yb, yc = S._yak_files()
    reads = G.input_path("reads_q")
    for args in (["subtract", ya, yb], ["isec", ya, yb, yc], ["isec", "-t", "2", yc, ya], ["recount", yc, reads]):
        m, r = _both(exe, args)
        assert m[0] == r[0] == 0 and m[2] == r[2] and len(r[2]) > 16 + 8 * 1024, args
    for args in (["inspect", ya], ["inspect", ya, yb], ["inspect", "-m7", yb, ya], ["inspect", "-m", "40", yc, ya], ["inspect", ya, ya]):
        m, r = _both(exe, args, out_flag=False)
        assert m[0] == r[0] == 0 and m[1] == r[1] and r[1].count(b"\n") > 3, args


def test_scanner_command_lines(exe):
    """triobin / trioeval / chkerr / sexchr of cli/scan.c - options, table loads, batching, what follows truncated records -
    against the reference binary's committed stdout (tests/golden/scan_*.txt)"""
    import scan_inputs as S
    paths = S.write_all(util.TMP)
    for y, (fa, k) in S.COUNTS.items():
        paths[y] = os.path.join(util.TMP, "yakb_flow_" + y)
        h, _ = O.count_file(paths[fa], k=k, pre=10, bf_shift=0)
        assert O.lib().yo_ch_dump(h, paths[y].encode()) == 0
        O.lib().yo_ch_destroy(h)
    for gold, cmd in S.CASES:
        r = subprocess.run([exe] + S.argv(cmd, paths), capture_output=True)
        assert r.returncode == 0, (gold, r.stderr.decode()[-500:])
        assert r.stdout == open(os.path.join(ROOT, "tests", "golden", gold), "rb").read(), gold


def test_qv_command_lines(exe):
    """`qv` without -p / -E: option parsing, the hist of the table, the solver and the CT / FR / ER / CV / QV lines"""
    import scan_inputs as S
    paths = S.write_all(util.TMP)
    y = os.path.join(util.TMP, "yakb_flow_qv.yak")
    h, _ = O.count_file(G.input_path("reads_q"), k=31, pre=10, bf_shift=0)
    assert O.lib().yo_ch_dump(h, y.encode()) == 0
    O.lib().yo_ch_destroy(h)
    asm = G.input_path("reads_a")
    for args in (["qv", y, asm], ["qv", "-l", "100", "-f", "0.3", "-e", "0.0001", "-t2", y, asm], ["qv", "-K", "20k", y, asm], ["qv", y, paths["child.fa"]]):
        m, r = _both(exe, args, out_flag=False)
        assert m[0] == r[0] == 0 and m[1] == r[1] and r[1].count(b"\n") > 1000, args
    m, r = _both(exe, ["qv", y], out_flag=False)
    assert m[0] == r[0] == 1
