"""Config 2 at (a tenth of) its size, once, with the reference beside it (BASELINE.md plan steps 3-4): a 9 Gbp prefix of the cfg2
read stream - 60 M reads of 150 bp from the 3 Gbp genome - from a FASTQ FILE, `count -k31 -p12 -b34 -o`:
the .yak of yak-b200 against the .yak of the UNMODIFIED reference (oracle/_ref/yak) on the same file (sha256), both wall times.

  python tools/validate_cfg2.py [reads=60000000] [bf=34] [gpus=1] > profiles/r02_cfg2_9gbp.log

The input is written by oracle/_bin/synthgen (the CPU twin of the device generator, same stream as bench.py).  Everything
lives in /dev/shm (input 307 B per read, two outputs of ~8 B per distinct k-mer)."""
import hashlib
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 60_000_000
bf = int(sys.argv[2]) if len(sys.argv) > 2 else 34
gpus = int(sys.argv[3]) if len(sys.argv) > 3 else 1
K, PRE, G = 31, 12, 3_000_000_000
threads = os.cpu_count() or 1
fn = "/dev/shm/yakb_cfg2.fq"


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


t0 = time.time()
r = subprocess.run([os.path.join(ROOT, "oracle", "_bin", "synthgen"), "20260925", str(G), "7", "0", str(n_reads), "150", "0.005", "1", "2",
                    str(K), fn, str(threads)], check=True, capture_output=True, text=True)
n_ev = int(r.stdout.strip())
print(f"input: {n_reads} reads, {n_reads * 150 / 1e9:.2f} Gbp, {os.path.getsize(fn) / 1e9:.2f} GB of FASTQ, {n_ev} k-mer events per pass "
      f"(generated in {time.time() - t0:.1f} s); host: {threads} threads", flush=True)
res = {}
for tag, cmd, env in (
        ("b200", [os.path.join(ROOT, "yak_b200", "bin", "yak-b200"), "count", f"-k{K}", f"-p{PRE}", f"-b{bf}", "-o", "/dev/shm/yakb_cfg2_b200.yak"] +
         (["-g", str(gpus)] if gpus > 1 else []) + [fn], dict(os.environ, YAKB_TIMING="1")),
        ("reference", [os.path.join(ROOT, "oracle", "_ref", "yak"), "count", f"-k{K}", f"-p{PRE}", f"-b{bf}", f"-t{threads}", "-o", "/dev/shm/yakb_cfg2_ref.yak", fn], dict(os.environ))):
    t0 = time.time()
    p = subprocess.run(cmd, capture_output=True, text=True, env=env)
    dt = time.time() - t0
    out = cmd[cmd.index("-o") + 1]
    if p.returncode != 0:
        print(f"{tag}: FAILED rc={p.returncode}\n{p.stderr[-2000:]}", flush=True)
        sys.exit(1)
    res[tag] = (dt, os.path.getsize(out), sha(out))
    tail = [ln for ln in p.stderr.splitlines() if ln.startswith("[T::") and "batch " not in ln or "Real time" in ln or "distinct k-mers after" in ln]
    print(f"{tag}: {' '.join(cmd)}\n  wall {dt:.2f} s = {2 * n_ev / dt / 1e9:.3f} G input events/s over both passes; output {res[tag][1]} bytes, sha256 {res[tag][2]}", flush=True)
    for ln in tail[-14:]:
        print("    " + ln, flush=True)
same = res["b200"][2] == res["reference"][2]
print(f"identical .yak bytes: {same}; speed-up of the whole job: {res['reference'][0] / res['b200'][0]:.1f}x ({gpus} GPU(s) vs {threads} host threads)", flush=True)
for f in (fn, "/dev/shm/yakb_cfg2_b200.yak", "/dev/shm/yakb_cfg2_ref.yak"):
    try:
        os.unlink(f)
    except OSError:
        pass
sys.exit(0 if same else 1)
