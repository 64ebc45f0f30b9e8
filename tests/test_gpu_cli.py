"""The C++ host (kmtricks_b200/bin/kmx: CLI + run-dir + plugin host over the C ABI) against the
unmodified reference CLI on the same fof: every file the two run directories share must be
byte-identical.  Also loads the reference's own example plugins, compiled unchanged."""
import filecmp
import os
import shutil
import subprocess
import tempfile

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KMX = os.path.join(ROOT, "kmtricks_b200", "bin", "kmx")
REF = os.path.join(ROOT, "oracle", "_ref", "bin", "kmtricks")
PLUG = os.path.join(ROOT, "oracle", "_ref", "plugins")


@pytest.fixture(scope="module")
def workdir():
    from kmtricks_b200 import synth
    from tests.conftest import EDGE_FASTA, EDGE_FASTQ_CRLF
    if not (os.path.exists(KMX) and os.path.exists(REF)):
        pytest.skip("kmx or the reference binary is not built")
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="kmx_cli_", dir=base)
    with open(f"{d}/fof.txt", "w") as f:
        for s in range(4):
            open(f"{d}/S{s}.fastq", "wb").write(synth.make_fastq(5, s, 20_000, L=150, G=100_000, d=4e-3, e=4e-3, revcomp=True))
            f.write(f"S{s}: {d}/S{s}.fastq" + (" ! 1" if s == 2 else "") + "\n")
        open(f"{d}/edge.fasta", "wb").write(EDGE_FASTA); open(f"{d}/crlf.fastq", "wb").write(EDGE_FASTQ_CRLF)
        f.write(f"E1 : {d}/edge.fasta ; {d}/crlf.fastq\n")
    yield d
    shutil.rmtree(d, ignore_errors=True)


def run_both(d, tag, args):
    common = ["pipeline", "--file", f"{d}/fof.txt", "--nb-partitions", "8", "--minimizer-size", "10", "--static-repart", "--keep-tmp"] + args
    a, b = f"{d}/ref_{tag}", f"{d}/kmx_{tag}"
    subprocess.run([REF] + common + ["--run-dir", a, "-t", "4"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([KMX] + common + ["--run-dir", b, "--threads", "3"], check=True)
    return a, b


def same_files(a, b, sub, must_exist=True):
    names = sorted(os.listdir(os.path.join(a, sub)))
    if must_exist:
        assert names, f"reference wrote nothing under {sub}"
    for n in names:
        pa, pb = os.path.join(a, sub, n), os.path.join(b, sub, n)
        if os.path.isdir(pa):
            same_files(a, b, os.path.join(sub, n), must_exist)
        else:
            assert os.path.exists(pb), f"missing {sub}/{n}"
            assert filecmp.cmp(pa, pb, shallow=False), f"{sub}/{n} differs"


@pytest.mark.parametrize("tag,args", [
    ("kmer_count", ["--kmer-size", "31", "--mode", "kmer:count:bin", "--hard-min", "2"]),
    ("kmer_pa_k63", ["--kmer-size", "63", "--mode", "kmer:pa:bin", "--hard-min", "1", "--soft-min", "3", "--share-min", "2", "--recurrence-min", "2"]),
    ("hash_bf", ["--kmer-size", "31", "--mode", "hash:bf:bin", "--hard-min", "2", "--bloom-size", "2000000", "--soft-min", "2", "--share-min", "1"]),
    ("hash_count", ["--kmer-size", "31", "--mode", "hash:count:bin", "--hard-min", "2", "--bloom-size", "2000000"]),
])
def test_run_dir_equals_reference(workdir, tag, args):
    a, b = run_both(workdir, tag, args)
    for sub in ("matrices", "merge_infos", "partition_infos", "counts", "repartition_gatb", "minimizers"):
        same_files(a, b, sub)
    assert filecmp.cmp(f"{a}/hash.info", f"{b}/hash.info", shallow=False)
    ca, cb = open(f"{a}/config_gatb/gatb.config", "rb").read(), open(f"{b}/config_gatb/gatb.config", "rb").read()
    assert len(ca) == len(cb) == 140 and ca[:32] == cb[:32] and ca[124:136] == cb[124:136]      # k, m, types; passes, partitions, bits per k-mer, banks
    if tag == "hash_bf":
        same_files(a, b, "fpr")


def test_template_example_plugin_fails_like_the_reference(workdir):
    """plugins/example/template_ex.cpp exports no `destroy`, so the reference refuses it
    (plugin_manager.hpp:73-79, PluginError); the kmx plugin host must fail the same loud way."""
    so = os.path.join(PLUG, "libtemplate_ex.so")
    if not os.path.exists(so):
        pytest.skip("reference plugins not built")
    common = ["pipeline", "--file", f"{workdir}/fof.txt", "--nb-partitions", "8", "--static-repart", "--kmer-size", "31",
              "--mode", "kmer:count:bin", "--plugin", so, "--plugin-config", "3"]
    r1 = subprocess.run([REF] + common + ["--run-dir", f"{workdir}/ref_tpl"], capture_output=True, text=True)
    r2 = subprocess.run([KMX] + common + ["--run-dir", f"{workdir}/kmx_tpl"], capture_output=True, text=True)
    assert r1.returncode != 0 and r2.returncode != 0
    assert "destroy" in r2.stderr


@pytest.mark.parametrize("lib,cfg,k", [("libbasic_ex.so", "2", "31"), ("libbasic_ex.so", "3", "45")])
def test_reference_example_plugins_load_unchanged(workdir, lib, cfg, k):
    """plugins/example/{basic,template}_ex.cpp compiled as they are (oracle/build_ref.sh) must load in
    the kmx plugin host and produce the matrices the reference produces with the same plugin."""
    so = os.path.join(PLUG, lib)
    if not os.path.exists(so):
        pytest.skip("reference plugins not built")
    a, b = run_both(workdir, f"plug_{lib[3:8]}_{k}", ["--kmer-size", k, "--mode", "kmer:count:bin", "--hard-min", "1",
                                                     "--plugin", so, "--plugin-config", cfg])
    same_files(a, b, "matrices")
    same_files(a, b, "merge_infos")


def test_repart_from_and_until_count(workdir):
    d = workdir
    a, b = run_both(d, "base", ["--kmer-size", "31", "--mode", "kmer:count:bin", "--hard-min", "2"])
    # reuse the reference's repartition file (--repart-from) and stop after counting
    subprocess.run([KMX, "pipeline", "--file", f"{d}/fof.txt", "--run-dir", f"{d}/kmx_rf", "--nb-partitions", "8", "--kmer-size", "31",
                    "--mode", "kmer:count:bin", "--hard-min", "2", "--repart-from", a, "--until", "count"], check=True)
    same_files(a, f"{d}/kmx_rf", "counts")
    assert os.listdir(f"{d}/kmx_rf/matrices") == []


def test_cli_errors_are_loud(workdir):
    r = subprocess.run([KMX, "pipeline", "--file", f"{workdir}/nope.txt", "--run-dir", f"{workdir}/x", "--nb-partitions", "4"], capture_output=True, text=True)
    assert r.returncode != 0 and "error" in r.stderr.lower()
    r = subprocess.run([KMX, "pipeline", "--file", f"{workdir}/fof.txt", "--run-dir", f"{workdir}/x"], capture_output=True, text=True)
    assert r.returncode != 0


def test_streamed_blocks_and_gz_input(workdir):
    """The host streams every FASTQ in blocks of whole records (--block-mib; here 1 MiB blocks over 6.3 MB files, several
    lanes) and inflates .gz on the fly: same run directory as the reference, which reads the plain files."""
    import gzip
    d = workdir
    a, _ = run_both(d, "base2", ["--kmer-size", "31", "--mode", "kmer:count:bin", "--hard-min", "2"])
    with open(f"{d}/fof_gz.txt", "w") as f:
        for line in open(f"{d}/fof.txt"):
            sid, rest = line.split(":", 1)
            if sid.strip() in ("S1", "S3"):
                src = rest.split("!")[0].strip()
                with open(src, "rb") as i, gzip.open(src + ".gz", "wb", compresslevel=1) as o:
                    shutil.copyfileobj(i, o)
                line = line.replace(src, src + ".gz")
            f.write(line)
    subprocess.run([KMX, "pipeline", "--file", f"{d}/fof_gz.txt", "--run-dir", f"{d}/kmx_blocks", "--nb-partitions", "8", "--minimizer-size", "10",
                    "--static-repart", "--keep-tmp", "--kmer-size", "31", "--mode", "kmer:count:bin", "--hard-min", "2", "--block-mib", "1",
                    "--threads", "3"], check=True)
    for sub in ("matrices", "merge_infos", "partition_infos", "counts"):
        same_files(a, f"{d}/kmx_blocks", sub)


def test_plugin_applies_in_pa_mode_too(workdir):
    """The reference calls the plugin for every merged row whatever the output mode (merge.hpp:249-257): here presence/absence
    rows with the unchanged basic_ex plugin."""
    so = os.path.join(PLUG, "libbasic_ex.so")
    if not os.path.exists(so):
        pytest.skip("reference plugins not built")
    a, b = run_both(workdir, "plug_pa", ["--kmer-size", "31", "--mode", "kmer:pa:bin", "--hard-min", "1", "--plugin", so, "--plugin-config", "2"])
    same_files(a, b, "matrices")
    same_files(a, b, "merge_infos")


def test_hash_bft_end_to_end_and_bf_files(workdir):
    """--mode hash:bft:bin from the command line (the reference CLI cannot reach it at this commit): every
    matrices/matrix_P.cmbf equals the file the reference's own HashMerger::write_as_bft writes from the reference's count
    files, and every filters/<id>.bf holds, after its header and the u64 bit count, the sample's row of every partition."""
    import numpy as np
    d = workdir
    harness = os.path.join(ROOT, "oracle", "_ref", "bin", "bft_harness")
    if not os.path.exists(harness):
        pytest.skip("bft_harness not built")
    args = ["--kmer-size", "31", "--hard-min", "2", "--bloom-size", "2000000", "--soft-min", "2", "--share-min", "1"]
    a, _ = run_both(d, "hash_bf2", args + ["--mode", "hash:bf:bin"])
    b = f"{d}/kmx_bft"
    subprocess.run([KMX, "pipeline", "--file", f"{d}/fof.txt", "--nb-partitions", "8", "--minimizer-size", "10", "--static-repart", "--run-dir", b,
                    "--mode", "hash:bft:bin", "--threads", "3"] + args, check=True)
    ids = [l.split(":")[0].strip() for l in open(f"{d}/fof.txt") if l.strip()]
    W = int(np.frombuffer(open(f"{a}/hash.info", "rb").read()[16:24], dtype=np.uint64)[0])
    rows = []
    for p in range(8):
        files = [f"{a}/counts/partition_{p}/{i}.hash" for i in ids]
        subprocess.run([harness, "bft", f"{d}/bft_{p}", str(W * p), str(W * (p + 1) - 1), "2", "1", "1"] + files, check=True)
        want = open(f"{d}/bft_{p}", "rb").read()
        got = open(f"{b}/matrices/matrix_{p}.cmbf", "rb").read()
        assert got == want, f"bft matrix {p}"
        rows.append(got[49:])
    for s, i in enumerate(ids):
        bf = open(f"{b}/filters/{i}.bf", "rb").read()
        assert len(bf) == 112 + 8 + 8 * (W // 8)
        assert int(np.frombuffer(bf[112:120], dtype=np.uint64)[0]) == 8 * W
        for p in range(8):
            assert bf[120 + p * (W // 8):120 + (p + 1) * (W // 8)] == rows[p][s * (W // 8):(s + 1) * (W // 8)], f"{i}.bf partition {p}"


def test_two_gpus_from_the_command_line(workdir):
    """--devices 0-1: one in-process rank per GPU (samples sharded for stage 1, one NCCL exchange, partitions sharded for
    count + merge); the run directory equals the reference's.  Skipped on a one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    d = workdir
    with open(f"{d}/fof4.txt", "w") as f:                 # the four FASTQ samples (one file each), hard-min override kept
        for line in open(f"{d}/fof.txt"):
            if line.startswith("S"):
                f.write(line)
    for tag, args in (("mg_hash", ["--kmer-size", "31", "--mode", "hash:bf:bin", "--hard-min", "2", "--bloom-size", "2000000"]),
                      ("mg_kmer", ["--kmer-size", "63", "--mode", "kmer:count:bin", "--hard-min", "2", "--soft-min", "2"])):
        common = ["pipeline", "--file", f"{d}/fof4.txt", "--nb-partitions", "8", "--minimizer-size", "10", "--static-repart", "--keep-tmp"] + args
        subprocess.run([REF] + common + ["--run-dir", f"{d}/ref_{tag}", "-t", "4"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.run([KMX] + common + ["--run-dir", f"{d}/kmx_{tag}", "--devices", "0-1", "--threads", "2"], check=True)
        for sub in ("matrices", "merge_infos", "partition_infos", "counts"):
            same_files(f"{d}/ref_{tag}", f"{d}/kmx_{tag}", sub)


def test_balanced_repartition_is_consumed_by_the_reference(workdir):
    """--balanced-repart: minimizer loads estimated on the device from the head of the samples, minimizers assigned heaviest
    first to the lightest partition, repartition.minimRepart written in the reference's format.  The reference run with
    --repart-from that directory gives the same counts / matrices / .pinfo, and the partitions are better balanced than
    with the static (hash) map."""
    import numpy as np
    d = workdir
    with open(f"{d}/fof4b.txt", "w") as f:
        for line in open(f"{d}/fof.txt"):
            if line.startswith("S"):
                f.write(line)
    common = ["pipeline", "--file", f"{d}/fof4b.txt", "--nb-partitions", "8", "--minimizer-size", "10", "--keep-tmp", "--kmer-size", "31",
              "--mode", "kmer:count:bin", "--hard-min", "2"]
    subprocess.run([KMX] + common + ["--run-dir", f"{d}/kmx_bal", "--balanced-repart", "--threads", "2"], check=True)
    subprocess.run([KMX] + common + ["--run-dir", f"{d}/kmx_sta", "--static-repart", "--threads", "2"], check=True)
    subprocess.run([REF] + common + ["--run-dir", f"{d}/ref_bal", "--repart-from", f"{d}/kmx_bal", "-t", "4"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for sub in ("matrices", "merge_infos", "partition_infos", "counts"):
        same_files(f"{d}/ref_bal", f"{d}/kmx_bal", sub)

    def imbalance(run):
        tot = np.zeros(8)
        for n in os.listdir(f"{run}/partition_infos"):
            tot += np.array([int(x) for x in open(f"{run}/partition_infos/{n}").read().split()], dtype=float)
        return tot.max() / tot.mean()
    bal, sta = imbalance(f"{d}/kmx_bal"), imbalance(f"{d}/kmx_sta")
    assert bal < 1.05 and bal <= sta, (bal, sta)


@pytest.mark.parametrize("tag,args", [
    ("txt_kmer_count", ["--kmer-size", "31", "--mode", "kmer:count:text", "--hard-min", "2"]),
    ("txt_k63_pa", ["--kmer-size", "63", "--mode", "kmer:pa:text", "--hard-min", "1", "--soft-min", "2", "--recurrence-min", "2"]),
    ("txt_hash_count", ["--kmer-size", "31", "--mode", "hash:count:text", "--hard-min", "2", "--bloom-size", "2000000"]),
])
def test_text_matrices(workdir, tag, args):
    """<kmer|hash>:<count|pa>:text (write_as_text / write_as_pa_text, merge.hpp:288-316,531-573): matrix_P.<ext>.txt byte for byte."""
    a, b = run_both(workdir, tag, args)
    same_files(a, b, "matrices")
    same_files(a, b, "merge_infos")


@pytest.mark.parametrize("tag,args", [
    ("hist_kmer", ["--kmer-size", "31", "--mode", "kmer:count:bin", "--hard-min", "3", "--hist"]),
    ("hist_hash", ["--kmer-size", "31", "--mode", "hash:count:bin", "--hard-min", "2", "--bloom-size", "2000000", "--hist"]),
    ("hist_k63", ["--kmer-size", "63", "--mode", "kmer:pa:bin", "--hard-min", "1", "--hist"]),
])
def test_abundance_histograms(workdir, tag, args):
    """--hist: histograms/<id>.hist (every distinct key of the sample binned before the hard-min test, lower 1, upper 255,
    histogram.hpp:34-207 + io/hist_file.hpp) byte for byte, and the matrices stay what they are without --hist."""
    a, b = run_both(workdir, tag, args)
    same_files(a, b, "histograms")
    same_files(a, b, "matrices")
    same_files(a, b, "counts")
