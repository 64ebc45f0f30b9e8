"""CPU emulation of the position-parallel stage-1 kernel (kmtricks_b200/csrc/s1_v5.cuh): the phase bodies
that the CUDA kernel s1_superk_v5 runs are compiled for the host and driven CTA by CTA, phase by phase, by
tests/emul/s1v5_emul.cpp; the (partition, canonical k-mer) multiset decoded from the emitted super-k-mer
records must equal the oracle's stage 1 (orc_s1_seq: Model.hpp:725-765,857-884,1254-1287,
Sequence2SuperKmer.hpp:137-147, fill_partitions.hpp:59-63) on reads with invalid characters, lower case,
lengths around k, low-complexity stretches and every text alignment."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    d = tmp_path_factory.mktemp("s1v5")
    exe = str(d / "s1v5_emul")
    obj = str(d / "orc.o")
    subprocess.run(["gcc", "-O1", "-c", "-o", obj, os.path.join(ROOT, "oracle", "kmx_oracle.c")], check=True)
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-Wall", "-o", exe,
                    os.path.join(ROOT, "tests", "emul", "s1v5_emul.cpp"), obj], check=True)
    return exe


CASES = [
    # k, m, P, reads, maxlen, seed, reads per CTA
    (31, 10, 64, 400, 150, 1, 32),
    (31, 10, 64, 300, 151, 2, 64),
    (21, 8, 16, 300, 101, 3, 32),
    (63, 10, 256, 300, 150, 4, 32),
    (63, 12, 64, 200, 250, 5, 32),
    (31, 10, 4, 100, 40, 6, 32),
    (32, 10, 8, 200, 300, 7, 16),
    (33, 4, 8, 200, 100, 8, 32),
    (15, 12, 8, 100, 80, 9, 32),
    (8, 4, 4, 100, 50, 10, 128),
    (31, 10, 64, 300, 31, 14, 32),
    (63, 4, 64, 300, 200, 15, 32),
    (15, 12, 8, 300, 80, 16, 128),    # > 1024 events per CTA: several flush rounds
]


@pytest.mark.parametrize("k,m,P,reads,maxlen,seed,R", CASES)
def test_s1v5_phases_match_oracle(emul, k, m, P, reads, maxlen, seed, R):
    r = subprocess.run([emul] + [str(v) for v in (k, m, P, reads, maxlen, seed, R)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert r.stdout.startswith("OK")
