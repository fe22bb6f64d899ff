"""`xyz_pcc.bin` container (a-15) -- byte-identical layout to the reference.

Reference: src/ai_pcc/GausPcgc/kit/op.py:32-48 (pack_byte_stream_ls / unpack_byte_stream) and the
header written at src/gs_compress/HAC/utils/pcc_utils.py:198-203, parsed at :271-276:

    f16 posQ | i32 n_base | i32[n_base,3] base xyz | u8[n_base] base occupancy
    | u16 n_streams | n_streams x ( u32 len | bytes )

Streams are level-major coarse->fine, stage-minor s0..s3 (pcc_utils.py:180-183).  Little-endian.
"""
from __future__ import annotations

import struct
from typing import List, Tuple

import numpy as np


def pack_byte_stream_ls(byte_stream_ls: List[bytes]) -> bytes:
    if len(byte_stream_ls) > 0xFFFF:
        raise ValueError("too many streams for the u16 count of the reference container")
    parts = [struct.pack("<H", len(byte_stream_ls))]
    for s in byte_stream_ls:
        parts.append(struct.pack("<I", len(s)))
        parts.append(bytes(s))
    return b"".join(parts)


def unpack_byte_stream(stream: bytes) -> List[bytes]:
    (n,) = struct.unpack_from("<H", stream, 0)
    out, cur = [], 2
    for _ in range(n):
        (ln,) = struct.unpack_from("<I", stream, cur)
        if cur + 4 + ln > len(stream):
            raise ValueError("truncated xyz_pcc.bin stream table")
        out.append(stream[cur + 4:cur + 4 + ln])
        cur += 4 + ln
    return out


def write_file(posQ: float, base_xyz: np.ndarray, base_occ: np.ndarray, streams: List[bytes]) -> bytes:
    base_xyz = np.ascontiguousarray(base_xyz, dtype=np.int32).reshape(-1, 3)
    base_occ = np.ascontiguousarray(base_occ, dtype=np.uint8).reshape(-1)
    assert base_xyz.shape[0] == base_occ.shape[0]
    return b"".join([
        np.array(posQ, dtype=np.float16).tobytes(),
        np.array(base_xyz.shape[0], dtype=np.int32).tobytes(),
        base_xyz.tobytes(),
        base_occ.tobytes(),
        pack_byte_stream_ls(streams),
    ])


def read_file(blob: bytes) -> Tuple[np.float16, np.ndarray, np.ndarray, List[bytes]]:
    if len(blob) < 6:
        raise ValueError("truncated xyz_pcc.bin header")
    posQ = np.frombuffer(blob[:2], dtype=np.float16)[0]
    n0 = int(np.frombuffer(blob[2:6], dtype=np.int32)[0])
    if n0 < 0 or 6 + 13 * n0 + 2 > len(blob):
        raise ValueError("corrupt xyz_pcc.bin base section")
    base_xyz = np.frombuffer(blob[6:6 + 12 * n0], dtype=np.int32).reshape(-1, 3)
    base_occ = np.frombuffer(blob[6 + 12 * n0:6 + 13 * n0], dtype=np.uint8)
    streams = unpack_byte_stream(blob[6 + 13 * n0:])
    return posQ, base_xyz, base_occ, streams


# ---- container version 2 (opt-in): the same outer layout with ONE extra trailing stream that names the format.
# A version-1 file has 4 * levels streams; a version-2 file has 4 * levels + 1, the last being V2_MAGIC + u32 chunk_size + u32 number
# of coded voxels (an integrity check the reference layout has no room for: a decode that desynchronises raises).  Every
# other stream is `u16 bytes_of_chunk[chunks]` + the chunks' bytes, each chunk an independent range coder (csrc/attr_ac.cu) over
# chunk_len(n, chunk_size) = min(chunk_size, max(64, ceil(n / 64))) symbols of the stream's n (GausPcgcCodec.chunk_len: short
# streams take shorter chunks so that a small level is not a handful of long serial chains).  The reference cannot read it (torchac codes a stream as one coder).
V2_MAGIC = b"GPCGC-V2"


def v2_trailer(chunk_size: int, n_voxels: int) -> bytes:
    return V2_MAGIC + struct.pack("<II", int(chunk_size), int(n_voxels))


def split_v2(streams: List[bytes]):
    """(streams without the trailer, chunk_size, coded voxels) for a version-2 stream list, (streams, 0, None) for version 1"""
    if len(streams) % 4 == 1 and streams[-1][:len(V2_MAGIC)] == V2_MAGIC and len(streams[-1]) == len(V2_MAGIC) + 8:
        chunk, n_vox = struct.unpack("<II", streams[-1][len(V2_MAGIC):])
        return list(streams[:-1]), chunk, n_vox
    return streams, 0, None
