"""GPU: the device BGZF/DEFLATE decoder against zlib on every flavour of block the feeder can meet
(dynamic / fixed / stored Huffman blocks, tiny and full-size BGZF blocks, several deflate levels)."""
import gzip
import os
import subprocess
import zlib

import numpy as np
import pytest

import bamio

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


# every test runs once per decoder: the warp-per-block kernel (1) and the lane-per-stream kernel (3) with each root-table width
KERNELS = [("1", "10"), ("3", "10")]


@pytest.fixture(autouse=True, params=KERNELS, ids=lambda k: f"variant{k[0]}-root{k[1]}")
def inflate_kernel(request, monkeypatch):
    monkeypatch.setenv("RTJX_INFLATE_VARIANT", request.param[0])
    monkeypatch.setenv("RTJX_INFLATE_ROOT", request.param[1])
    return request.param


def device_inflate(path, max_blocks=0):
    import regtools_b200 as rt
    ex = rt.JunctionsExtractor(path, ".", 0)
    data = ex.inflate_file(max_blocks)
    st = ex.stats()
    ex.close()
    return data, st


def host_inflate(path):
    return gzip.decompress(open(path, "rb").read())


@pytest.mark.parametrize("rel", ["hcc1395/test_hcc1395.bam", "hcc1395/test_hcc1395.2.bam", "kat/kat.bam", "kat/synth.bam"])
def test_fixture_bams(rel):
    path = os.path.join(GOLD, rel)
    got, st = device_inflate(path)
    want = host_inflate(path)
    assert len(got) == len(want)
    assert got == want
    assert st["kernel_launches"] == 1 and st["inflated_bytes"] == len(want)


@pytest.mark.parametrize("level", [1, 6, 9])
def test_generated_bam_levels(level, tmp_path):
    bam = str(tmp_path / f"l{level}.bam")
    subprocess.check_call([os.path.join(ROOT, "tools", "bamgen"), "gen", "--out", bam, "--config", "tiny", "--reads", "300000",
                           "--seed", str(level), "--level", str(level)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    got, st = device_inflate(bam)
    assert got == host_inflate(bam)


def _bgzf(data, level, strategy=zlib.Z_DEFAULT_STRATEGY):
    import struct
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    comp = c.compress(data) + c.flush()
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(comp) + 25) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def _bam_header():
    import struct
    text = b"@HD\tVN:1.4\n"
    return b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", 1) + struct.pack("<i", 2) + b"c\0" + struct.pack("<i", 1000)


def test_stored_fixed_and_repetitive_blocks(tmp_path):
    """Level 0 (stored), Z_FIXED (fixed Huffman), long overlapping matches (runs), incompressible noise."""
    rng = np.random.default_rng(0)
    payloads = [
        (bytes(rng.integers(0, 256, 40000, dtype=np.uint8)), 0, zlib.Z_DEFAULT_STRATEGY),          # stored
        (b"BAM\x01" + bytes(60000), 6, zlib.Z_DEFAULT_STRATEGY),                                   # dist-1 runs of 258
        (b"abcabcabd" * 5000, 9, zlib.Z_DEFAULT_STRATEGY),                                          # short-period matches
        (bytes(rng.integers(0, 4, 50000, dtype=np.uint8)), 6, zlib.Z_FIXED),                        # fixed Huffman
        (bytes(rng.integers(0, 256, 65280, dtype=np.uint8)), 6, zlib.Z_DEFAULT_STRATEGY),           # incompressible, max size
        (b"x", 6, zlib.Z_DEFAULT_STRATEGY), (bytes(rng.integers(60, 70, 3000, dtype=np.uint8)), 1, zlib.Z_HUFFMAN_ONLY),
        (bytes(np.repeat(rng.integers(0, 256, 300, dtype=np.uint8), 200)), 6, zlib.Z_RLE),
        # ~16k four-byte matches in one block: close to the most a block can hold (21845); the lane decoder's lists are sized for it
        (rng.integers(0, 256, (64, 4), dtype=np.uint8)[rng.integers(0, 64, 16000)].tobytes(), 9, zlib.Z_DEFAULT_STRATEGY),
    ]
    path = str(tmp_path / "mix.bam")
    hdr = _bam_header()
    with open(path, "wb") as f:
        f.write(_bgzf(hdr, 6))
        for data, level, strat in payloads:
            f.write(_bgzf(data, level, strat))
        f.write(bamio.EOF_BLOCK)
    got, st = device_inflate(path)
    want = hdr + b"".join(p[0] for p in payloads)
    assert got == want
    assert st["bgzf_blocks"] == len(payloads) + 1


def test_corrupt_block_is_reported(tmp_path):
    good = _bgzf(b"hello world, hello world, hello world" * 100, 6)
    bad = bytearray(good)
    bad[40] ^= 0x5A
    bad[41] ^= 0xA5
    path = str(tmp_path / "bad.bam")
    open(path, "wb").write(_bgzf(_bam_header(), 6) + bytes(bad) + bamio.EOF_BLOCK)
    try:
        got, _ = device_inflate(path)
    except RuntimeError as e:
        assert "device inflate failed" in str(e)
    else:       # a flipped bit may still decode to *something* of the right length; it must not be the original
        assert got != _bam_header() + b"hello world, hello world, hello world" * 100
