"""Host utilities of the C ABI (no GPU): BGZF inflate / deflate against Python's gzip / zlib (BGZF is multi-member gzip,
utils/bgzf/bgzf-files.go:95-127,324-431) and the BAM header walk."""
import gzip
import struct
import zlib

import numpy as np
import pytest

from elprep_b200 import bgzf

EOF_BLOCK = bytes([0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0x00, 0xff, 0x06, 0x00, 0x42, 0x43, 0x02, 0x00, 0x1b, 0x00, 0x03, 0x00, 0, 0, 0, 0, 0, 0, 0, 0])


def py_bgzf_block(payload, extra_subfield=b""):
    """one BGZF block built by hand from the format description (SAMv1 section 4.1)"""
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    cdata = co.compress(payload) + co.flush()
    xlen = 6 + len(extra_subfield)
    bsize = 12 + xlen + len(cdata) + 8
    return (struct.pack("<BBBBIBBH", 31, 139, 8, 4, 0, 0, 255, xlen) + extra_subfield + b"BC" + struct.pack("<HH", 2, bsize - 1) + cdata
            + struct.pack("<II", zlib.crc32(payload) & 0xffffffff, len(payload)))


def blocks_of(z):
    out, x = [], 0
    while x < len(z):
        assert z[x:x + 4] == b"\x1f\x8b\x08\x04" and z[x + 12:x + 16] == b"BC\x02\x00"
        bsize = struct.unpack_from("<H", z, x + 16)[0] + 1
        out.append(z[x:x + bsize]); x += bsize
    assert x == len(z)
    return out


@pytest.mark.parametrize("n", [0, 1, 0xff00 - 1, 0xff00, 0xff00 + 1, 1_000_003])
@pytest.mark.parametrize("kind", ["text", "random"])
def test_roundtrip_and_gzip_compat(n, kind):
    rng = np.random.default_rng(n + len(kind))
    d = rng.integers(0, 4 if kind == "text" else 256, size=n, dtype=np.uint8)          # "random" is incompressible: stored blocks
    z = bgzf.deflate(d, n_threads=4)
    assert gzip.decompress(z.tobytes()) == d.tobytes()
    assert np.array_equal(bgzf.inflate(z, n_threads=3), d)
    bl = blocks_of(z.tobytes())
    assert bl[-1] == EOF_BLOCK and len(bl) == (n + 0xff00 - 1) // 0xff00 + 1
    assert all(struct.unpack_from("<I", b, len(b) - 4)[0] <= 0xff00 and len(b) <= 65536 for b in bl)
    assert np.array_equal(bgzf.inflate(bgzf.deflate(d, write_eof=False)), d)


def test_inflate_foreign_blocks_and_errors():
    rng = np.random.default_rng(5)
    parts = [rng.integers(0, 7, size=k, dtype=np.uint8).tobytes() for k in (10, 65536, 1, 30000)]        # 65536-byte blocks as the reference writes them
    z = b"".join(py_bgzf_block(p) for p in parts[:2]) + py_bgzf_block(parts[2], extra_subfield=b"XY" + struct.pack("<H", 3) + b"abc") + py_bgzf_block(parts[3]) + EOF_BLOCK
    assert bgzf.inflate(z).tobytes() == b"".join(parts)
    bad = bytearray(z); bad[40] ^= 0xff                                    # payload byte of block 0 -> CRC / stream error
    with pytest.raises(bgzf.BgzfError):
        bgzf.inflate(bytes(bad))
    with pytest.raises(bgzf.BgzfError):
        bgzf.inflate(z[:-5])                                               # truncated
    nobc = bytearray(py_bgzf_block(b"hello")); nobc[12:14] = b"QQ"         # "missing BC extra subfield in BGZF header"
    with pytest.raises(bgzf.BgzfError):
        bgzf.inflate(bytes(nobc))
    with pytest.raises(bgzf.BgzfError):
        bgzf.inflate(gzip.compress(b"plain gzip is not BGZF"))


def test_bam_header_size():
    text = b"@HD\tVN:1.6\tSO:unsorted\n@SQ\tSN:chr1\tLN:1000\n"
    refs = [(b"chr1\0", 1000), (b"chrUn_x\0", 77)]
    hdr = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs)) + b"".join(struct.pack("<i", len(n)) + n + struct.pack("<i", ln) for n, ln in refs)
    assert bgzf.bam_header_size(hdr + b"\x10\0\0\0rest") == (len(hdr), 2)
    with pytest.raises(bgzf.BgzfError):
        bgzf.bam_header_size(hdr[:-3])
    with pytest.raises(bgzf.BgzfError):
        bgzf.bam_header_size(b"BAX\1" + hdr[4:])
