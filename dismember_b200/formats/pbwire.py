"""Minimal protobuf wire-format codec (varint / 32-bit / 64-bit / length-delimited).

Enough for the reference's three tiny schemas (tdm/src/main/protobuf/tree.proto,
store_kv.proto; deep-retrieval/src/main/protobuf/item_mapping.proto) without a
protoc step.  A message is decoded into ``{field_number: [raw values]}``.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple


def read_varint(b: bytes, p: int) -> Tuple[int, int]:
    v = 0
    shift = 0
    while True:
        c = b[p]
        p += 1
        v |= (c & 0x7F) << shift
        if not c & 0x80:
            return v, p
        shift += 7
        if shift > 70:
            raise ValueError("varint too long")


def write_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        c = v & 0x7F
        v >>= 7
        if v:
            out.append(c | 0x80)
        else:
            out.append(c)
            return bytes(out)


def to_int32(v: int) -> int:
    v &= 0xFFFFFFFFFFFFFFFF
    if v >= 1 << 63:
        v -= 1 << 64
    return v


def decode(b: bytes) -> Dict[int, List]:
    out: Dict[int, List] = {}
    p = 0
    n = len(b)
    while p < n:
        key, p = read_varint(b, p)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, p = read_varint(b, p)
        elif wt == 1:
            v = b[p:p + 8]
            p += 8
        elif wt == 2:
            ln, p = read_varint(b, p)
            v = b[p:p + ln]
            if len(v) != ln:
                raise ValueError("truncated length-delimited field")
            p += ln
        elif wt == 5:
            v = b[p:p + 4]
            p += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        out.setdefault(fno, []).append(v)
    return out


def packed_varints(b: bytes) -> List[int]:
    out = []
    p = 0
    while p < len(b):
        v, p = read_varint(b, p)
        out.append(to_int32(v))
    return out


def field_varint(fno: int, v: int) -> bytes:
    return write_varint(fno << 3) + write_varint(v)


def field_bytes(fno: int, v: bytes) -> bytes:
    return write_varint((fno << 3) | 2) + write_varint(len(v)) + v


def field_float(fno: int, v: float) -> bytes:
    return write_varint((fno << 3) | 5) + struct.pack("<f", v)
