"""`swimm -S search -m 3 -x N` on all N GPUs of the box against the unmodified reference CLI (`-m 0 -v 32 -c <cores>`) on
a cfg3-shaped database (Environmental-NR-like: ~217 residues per sequence), written directly in the preprocessed format
both binaries read.  Compares the printed hit lists and reports both search times.
usage: python tools/cli_multi_gpu.py [scale of the 6M-sequence database] [gpus] [top]"""
import os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from swimm_b200 import synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
gpus = sys.argv[2] if len(sys.argv) > 2 else "0"
top = sys.argv[3] if len(sys.argv) > 3 else "10"
REF = os.path.join(ROOT, "oracle", "_ref", "swimm")
OURS = os.path.join(ROOT, "swimm_b200", "swimm")
cores = str(os.cpu_count() or 4)


def hits(stdout):
    out, cur = [], None
    for line in stdout.split("\n"):
        if line.startswith("Query description:"):
            cur = []
            out.append(cur)
        m = re.match(r"^(-?\d+)\t.*syn\|(\d+)\|", line)
        if m and cur is not None:
            cur.append((int(m.group(1)), int(m.group(2))))
    return out


def field(stdout, name):
    for line in stdout.split("\n"):
        if line.startswith(name):
            return line.split("\t")[-1].strip()
    return "?"


q = synth.make_queries(np.random.default_rng(7), synth.QUERY_LENGTHS)
n = int(6_000_000 * scale) // 16 * 16
t0 = time.time()
dl, dc = synth.sorted_db(3000, n, mu=5.2, sigma=0.6, queries=q)
with tempfile.TemporaryDirectory() as tmp:
    prefix = os.path.join(tmp, "db")
    with open(prefix + ".info", "w") as f:
        f.write("%d %d %d" % (n, len(dc), 24))
    with open(prefix + ".seq", "wb") as f:
        f.write(np.asarray(dl, "<u2").tobytes())
        f.write(np.asarray(dc, np.int8).tobytes())
    with open(prefix + ".desc", "w") as f:
        f.write("".join(">syn|%09d| s\n" % i for i in range(n)))
    qf = os.path.join(tmp, "q.fasta")
    synth.write_fasta(qf, q)
    print("database: %d sequences, %d residues, %d queries (%d residues); files written in %.1f s"
          % (n, len(dc), q.n, int(q.lengths.sum()), time.time() - t0), flush=True)
    t0 = time.time()
    ours = subprocess.run([OURS, "-S", "search", "-q", qf, "-d", prefix, "-m", "3", "-x", gpus, "-r", top, "--verbose"],
                          check=True, capture_output=True)
    t_ours = time.time() - t0
    print(ours.stderr.decode("latin-1").strip())
    ours = ours.stdout.decode("latin-1")
    print("this repo: Search time %s, %s, wall %.1f s incl. database load (%s)"
          % (field(ours, "Search time:"), field(ours, "Search speed:"), t_ours, field(ours, "Execution mode:")), flush=True)
    t0 = time.time()
    ref = subprocess.run([REF, "-S", "search", "-q", qf, "-d", prefix, "-m", "0", "-v", "32", "-c", cores, "-r", top],
                         check=True, capture_output=True).stdout.decode("latin-1")
    t_ref = time.time() - t0
    print("reference: Search time %s, %s, wall %.1f s (%s threads)" % (field(ref, "Search time:"), field(ref, "Search speed:"), t_ref, cores))
    h_ref, h_ours = hits(ref), hits(ours)
    ok = h_ref == h_ours and len(h_ref) == q.n
    print("best score %d; hit lists identical for %d queries x top %s: %s" % (max(h[0][0] for h in h_ref if h), len(h_ref), top, ok))
sys.exit(0 if ok else 1)
