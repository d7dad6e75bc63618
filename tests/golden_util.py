"""Regenerate golden inputs (seeded) and look up the committed digests."""
import json
import os

from yak_b200 import synth
import util

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLD = json.load(open(os.path.join(HERE, "golden.json")))


def input_path(name):
    sg, G, sr, n, fq = GOLD["inputs"][name]["params"]
    p = os.path.join(util.TMP, f"yakb_gold_{name}" + (".fq" if fq else ".fa"))
    if not os.path.exists(p) or os.path.getsize(p) != GOLD["inputs"][name]["bytes"]:
        with open(p, "wb") as f:
            f.write(synth.reads_file_bytes(sg, G, sr, n, fastq=fq))
    return p


def case_id(c):
    return f"{c['input']}-k{c['k']}-p{c['pre']}-b{c['bf_shift']}"
