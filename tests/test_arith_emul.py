"""Device arithmetic of stage 2 on the CPU: kmtricks_b200/csrc/common.cuh and records.cuh (XXH64 8/16-byte paths of
KmXXHash, sorting_count.hpp:346-363 / xxhash.h:3454-3673; exact `% W` by multiplication; reverse complement,
LargeInt1.pri:137-158, LargeInt2.pri:170-198; k-mer extraction from the super-k-mer records) are compiled UNMODIFIED
for the host through a stand-in <cuda_runtime.h> (tests/emul/shim) and compared with the oracle and plain integer
arithmetic; the records are built by stage 1's own record builder, so the bucket format's writer and reader are
checked against each other."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stage2_device_arithmetic_matches_oracle(tmp_path):
    exe = str(tmp_path / "arith_emul")
    obj = str(tmp_path / "orc.o")
    subprocess.run(["gcc", "-O1", "-c", "-o", obj, os.path.join(ROOT, "oracle", "kmx_oracle.c")], check=True)
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-Wall", "-Wno-unused-variable", "-Wno-unused-function",
                    "-I" + os.path.join(ROOT, "tests", "emul", "shim"), "-o", exe,
                    os.path.join(ROOT, "tests", "emul", "arith_emul.cpp"), obj], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "OK", r.stderr + r.stdout
