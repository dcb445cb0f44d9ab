"""Minimal BAM writer for test fixtures (BGZF via zlib, records from Python tuples)."""
import struct
import zlib

OPS = "MIDNSHP=XB"


def parse_cigar(s):
    out, num = [], ""
    for ch in s:
        if ch.isdigit():
            num += ch
        else:
            out.append((int(num) << 4) | OPS.index(ch))
            num = ""
    return out


def reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def record(tid, pos, cigar, flag=0, mapq=60, aux=b"", name=b"r", l_seq=None):
    """cigar: string or list of uint32 words; aux: raw bytes (e.g. b'XSA+')."""
    words = parse_cigar(cigar) if isinstance(cigar, str) else list(cigar)
    qlen = sum(w >> 4 for w in words if (w & 0xF) in (0, 1, 4, 7, 8))
    if l_seq is None:
        l_seq = qlen
    rlen = sum(w >> 4 for w in words if (w & 0xF) in (0, 2, 3, 7, 8))
    end = pos + (rlen if (words and not flag & 4) else 1)
    name = name + b"\0"
    body = struct.pack("<iiBBHHHiiii", tid, pos, len(name), mapq, reg2bin(pos, end) & 0xFFFF, len(words), flag, l_seq, -1, -1, 0)
    body += name + b"".join(struct.pack("<I", w) for w in words)
    body += bytes((l_seq + 1) // 2) + bytes([30]) * l_seq + aux
    return struct.pack("<i", len(body)) + body


def bgzf_block(data):
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = c.compress(data) + c.flush()
    bsize = len(comp) + 25
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def write_bam(path, contigs, records, block_size=0xFF00, eof=True, extra_blocks=()):
    """contigs: list of (name, length); records: list of bytes from record()."""
    text = "@HD\tVN:1.4\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in contigs)
    hdr = b"BAM\x01" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(contigs))
    for n, l in contigs:
        hdr += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l)
    with open(path, "wb") as f:
        f.write(bgzf_block(hdr))
        payload = b"".join(records)
        for o in range(0, len(payload), block_size):
            f.write(bgzf_block(payload[o:o + block_size]))
        for b in extra_blocks:
            f.write(b)
        if eof:
            f.write(EOF_BLOCK)
