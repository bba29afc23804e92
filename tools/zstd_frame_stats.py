"""What is inside the zstd frames of a Blosc-zstd chunk of a readout image?  Parses the block / literals / sequences
HEADERS of every frame (no decoding) so the device decoder can be laid out for the real workload: how many zstd
blocks per 256 KiB Blosc block, raw vs RLE vs compressed, Huffman literal bytes and stream count, sequences per block.
CPU only.  Usage: python tools/zstd_frame_stats.py"""
import struct
import sys
from collections import Counter
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from merfish3d_analysis_b200 import _capi, synthetic  # noqa: E402


def frame_blocks(f: bytes):
    assert struct.unpack_from("<I", f, 0)[0] == 0xFD2FB528
    fhd = f[4]
    single, dict_flag, fcs_flag = (fhd >> 5) & 1, fhd & 3, fhd >> 6
    ip = 5 + (0 if single else 1) + (4 if dict_flag == 3 else dict_flag)
    ip += (1 if single else 0) if fcs_flag == 0 else (2, 4, 8)[fcs_flag - 1]
    while True:
        bh = f[ip] | (f[ip + 1] << 8) | (f[ip + 2] << 16)
        ip += 3
        last, btype, size = bh & 1, (bh >> 1) & 3, bh >> 3
        info = {"type": ("raw", "rle", "compressed")[btype], "size": size}
        if btype == 2:
            p = f[ip : ip + size]
            lt, fmt = p[0] & 3, (p[0] >> 2) & 3
            if lt < 2:
                hdr, regen = (1, p[0] >> 3) if fmt in (0, 2) else ((2, (p[0] >> 4) | (p[1] << 4)) if fmt == 1 else (
                    3, (p[0] >> 4) | (p[1] << 4) | (p[2] << 12)))
                comp, streams = (regen if lt == 0 else 1), 0
            else:
                v = int.from_bytes(p[:5], "little")
                hdr, bits = (3, 10) if fmt < 2 else ((4, 14) if fmt == 2 else (5, 18))
                regen, comp = (v >> 4) & ((1 << bits) - 1), (v >> (4 + bits)) & ((1 << bits) - 1)
                streams = 1 if fmt == 0 else 4
            q = hdr + comp
            nseq = p[q]
            if nseq >= 128:
                nseq = ((nseq - 128) << 8) + p[q + 1] if nseq < 255 else p[q + 1] + (p[q + 2] << 8) + 0x7F00
            info.update(lit_type=("raw", "rle", "huffman", "huffman-repeat")[lt], lit_regen=regen, lit_comp=comp,
                        lit_streams=streams, nseq=nseq, seq_bytes=size - q)
        yield info
        ip += size if btype != 1 else 1
        if last:
            return


def main():
    matrix = synthetic.mhd4_codebook_matrix(16)
    vol = synthetic.make_stack(matrix, (16, 512, 512), 7)[0]  # one (16, 512, 512) chunk of one bit
    frame = _capi.blosc_encode_host(vol, 2, "zstd", 5, "bitshuffle")
    h = _capi.blosc_info(frame)
    nblocks = -(-h["nbytes"] // h["blocksize"])
    starts = struct.unpack_from(f"<{nblocks}i", frame, 16)
    print(f"chunk {vol.shape} uint16: {h['nbytes']} B -> {h['cbytes']} B, {nblocks} Blosc blocks of {h['blocksize']} B "
          f"(16 bit rows of {h['blocksize'] // 16} B each)")
    kinds, lit_kinds = Counter(), Counter()
    rows = []
    for j, s in enumerate(starts):
        (cb,) = struct.unpack_from("<i", frame, s)
        if cb == h["blocksize"]:
            kinds["stored"] += 1
            continue
        blocks = list(frame_blocks(frame[s + 4 : s + 4 + cb]))
        for b in blocks:
            kinds[b["type"]] += 1
            if b["type"] == "compressed":
                lit_kinds[(b["lit_type"], b["lit_streams"])] += 1
                rows.append(b)
        if j == 0:
            print("first Blosc block, its zstd blocks:")
            for b in blocks:
                print("   ", b)
    print("zstd blocks by type:", dict(kinds))
    print("literal sections (type, streams):", dict(lit_kinds))
    if rows:
        for key in ("size", "lit_regen", "lit_comp", "nseq", "seq_bytes"):
            v = np.array([r[key] for r in rows])
            print(f"  compressed blocks, {key:9s}: mean {v.mean():9.1f}  min {v.min():7d}  max {v.max():7d}")


if __name__ == "__main__":
    main()
