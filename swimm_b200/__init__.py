"""swimm_b200 -- B200-native Smith-Waterman protein database search behind SWIMM's `search` interface.

    csrc/   hand-written sm_100a CUDA kernels + the C ABI (libswimm_cuda.so, include/swimm_gpu.h)
    host/   the plain-C host (swimm CLI, preprocess, FASTA / database I/O, matrices; libswimm_host.so)
    gpu.py  ctypes view of the C ABI          host.py  ctypes view of the host data layer
    synth.py seeded synthetic workloads (BASELINE.json configurations)
"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.dirname(os.path.abspath(__file__))


def build(verbose: bool = False) -> None:
    """Compile the CUDA library (nvcc, sm_100a), the C host and the CLI in-tree."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.run(["make", "-C", os.path.join(PKG, "csrc"), "-j", str(os.cpu_count() or 4)], check=True, stdout=out)
    subprocess.run(["make", "-C", os.path.join(PKG, "host")], check=True, stdout=out)
