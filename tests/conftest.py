import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def hcc_bam(golden_dir):
    return os.path.join(golden_dir, "hcc1395", "test_hcc1395.bam")


@pytest.fixture(scope="session")
def motif_fastas(tmp_path_factory):
    """The generated FASTA fixtures (tests/fasta_fixture.py): name -> path."""
    import fasta_fixture as ff
    d = tmp_path_factory.mktemp("fasta")
    return {"kat": ff.write_kat_fasta(str(d / "kat.fa")), "synth": ff.write_synth_fasta(str(d / "synth.fa")),
            "kat_no10": ff.write_kat_fasta(str(d / "kat_no10.fa"), drop="10")}


def motif_manifest():
    rows = []
    for line in open(os.path.join(ROOT, "tests", "golden", "motif", "MANIFEST.tsv")):
        bam, fa, out, rc, args, err = line.rstrip("\n").split("\t")
        rows.append((bam, fa, out, int(rc), args.split(), err))
    return rows
