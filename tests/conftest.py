import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def synth_samples():
    """4 small samples of strict 4-line FASTQ, both strands, SNPs + errors."""
    from kmtricks_b200 import synth
    return [[synth.make_fastq(7, s, 3000, L=150, G=20000, d=5e-3, e=5e-3, revcomp=True)] for s in range(4)]


EDGE_FASTA = (
    b">single_N\n" + b"ACGTTGCAAGGCTTAACCGGTTAACGATCGATCGGATCGATTAGCN" + b"GATTACAGATTACAGATTACAGGATCCGATCGATCGGGATATCGA\n"
    b">NN_and_IUPAC\nACGTTGCAAGGCTTAACCGGTTAACGATCGATCGGATCNNGATTACAGATTACAGATTACAGGATCCGATCGRTCGGGATATCGAACGTGTGTGACGATGCATGCATGCATAGCTAGCTAGGGATC\n"
    b">lowercase\nacgttgcaaggcttaaccggttaacgatcgatcggatcgattagcgattacagattacagattacaggatccgatc\n"
    b">shorter_than_k\nACGTACGTACGTACGTACGT\n"
    b">exactly_k31\nACGTTGCAAGGCTTAACCGGTTAACGATCGA\n"
    b">polyA\n" + b"A" * 80 + b"\n"
    b">A_run_inside\nGATTACAGATTACAGGATCCGATCGATCGG" + b"A" * 45 + b"CGATCGGATCGATTAGCGATTACAGATTACAGATT\n"
    b">multi_line\nACGTTGCAAGGCTTAACCGGTTAACGATCGATCGGATCGATTAGCGATTAC\nAGATTACAGATTACAGGATCCGATCGATCGGGATATCGAACGTGTGTGACG\nATGCATGCATGCATAGCTAGCTAGGGATC\n"
)

EDGE_FASTQ_CRLF = (b"@r0\r\nACGTTGCAAGGCTTAACCGGTTAACGATCGATCGGATCGATTAGCGATTACAGATTACAG\r\n+\r\n" + b"I" * 60 + b"\r\n"
                   b"@r1\r\nGATTACAGATTACAGGATCCGATCGATCGGGATATCGAACGTGTGTGACGATGCATGCAT\r\n+\r\n" + b"I" * 60 + b"\r\n")


@pytest.fixture(scope="session")
def edge_samples():
    return [[EDGE_FASTA], [EDGE_FASTA[:400] + b"\n", EDGE_FASTQ_CRLF]]
