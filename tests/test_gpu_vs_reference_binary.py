"""End-to-end parity with the UNMODIFIED reference binary at BASELINE.json's cfg1 size, on the GPU box:
the same FASTA files go through `oracle/_ref/swimm -S preprocess/search -m 0 -v 32` (the reference's own CPU
search) and through this repo's `swimm -S preprocess/search -m 3`; the printed hit lists must be identical.
oracle/_ref/swimm is the prebuilt checker binary (it travels with the repo snapshot); nothing here reads
/root/reference."""
import os
import re
import subprocess

import numpy as np
import pytest

from swimm_b200 import synth
from tests.helpers import ROOT

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "oracle", "_ref", "swimm")
OURS = os.path.join(ROOT, "swimm_b200", "swimm")


def _hits(stdout):
    """[(query title, [(score, db title), ...]), ...] from a search report (reference swimm.c:154-160)."""
    out, cur = [], None
    for line in stdout.split("\n"):
        if line.startswith("Query description:"):
            cur = []
            out.append((line.split("\t")[-1].strip(), cur))
        m = re.match(r"^(-?\d+)\t(.*)$", line)
        if m and cur is not None:
            cur.append((int(m.group(1)), re.sub(r"[^\x20-\x7e]", "", m.group(2)).strip()))
    return out


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/swimm not built (make -C oracle ref)")
@pytest.mark.parametrize("matrix,go,ge", [("blosum62", 10, 2), ("pam250", 8, 1)])
def test_cfg1_hit_lists_equal_the_reference_binary(tmp_path, matrix, go, ge):
    rng = np.random.default_rng(42)
    q = synth.make_queries(rng, [144, 464, 1000])            # ascending: reference sequences.c:276 vs :344
    db = synth.make_db(42, 100_000, queries=q, plant_fraction=0.002)
    # exact duplicates -> ties between different database indices inside the printed top-r
    for a, b in [(10, 20), (300, 301), (5000, 70000)]:
        k = min(db.lengths[a], db.lengths[b])
        db.residues[db.offsets[b]:db.offsets[b] + k] = db.residues[db.offsets[a]:db.offsets[a] + k]
    dbf, qf = str(tmp_path / "db.fasta"), str(tmp_path / "q.fasta")
    synth.write_fasta(dbf, db)
    synth.write_fasta(qf, q)
    cores = str(os.cpu_count() or 4)
    subprocess.run([REF, "-S", "preprocess", "-i", dbf, "-o", str(tmp_path / "ref"), "-c", cores], check=True,
                   stdout=subprocess.DEVNULL)
    subprocess.run([OURS, "-S", "preprocess", "-i", dbf, "-o", str(tmp_path / "ours")], check=True,
                   stdout=subprocess.DEVNULL)
    for ext in ("info", "seq"):
        assert open(str(tmp_path / "ref") + "." + ext, "rb").read() == open(str(tmp_path / "ours") + "." + ext, "rb").read()
    top = "200"
    ref = subprocess.run([REF, "-S", "search", "-q", qf, "-d", str(tmp_path / "ref"), "-m", "0", "-v", "32", "-c", cores,
                          "-r", top, "-s", matrix, "-g", str(go), "-e", str(ge)], check=True, capture_output=True)
    ours = subprocess.run([OURS, "-S", "search", "-q", qf, "-d", str(tmp_path / "ours"), "-m", "3", "-r", top,
                           "-s", matrix, "-g", str(go), "-e", str(ge)], check=True, capture_output=True)
    h_ref = _hits(ref.stdout.decode("latin-1"))
    h_ours = _hits(ours.stdout.decode("latin-1"))
    assert len(h_ref) == 3 and all(len(h) == 200 for _, h in h_ref)
    assert h_ours == h_ref
