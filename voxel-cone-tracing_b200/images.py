"""Dependency-free image decode / encode for the asset path.

The reference decodes its textures with the vendored stb_image v2.26 (`stbi_load(path, &w, &h, &n, 0)`,
/root/reference/Voxel_Cone_Tracing_Final/Model.h:150) and uploads them with the channel count the file has (RED / RGB /
RGBA, Model.h:159-169), top row first.  This module restates the part of that decoder the asset path needs, on numpy +
zlib only: PNG (8-bit grey / grey+alpha / RGB / RGBA, 1-8 bit grey and palette incl. tRNS, non-interlaced), binary PPM / PGM and
uncompressed or RLE true-colour / grey TGA.  Anything else (JPEG, interlaced or 16-bit PNG, ...) raises
UnsupportedImage; objloader then falls back to Pillow if it is installed.  Output: uint8 array (h, w, c), c in {1, 3, 4}
(grey+alpha is expanded to RGBA, as the reference's 1/3/4-channel upload has no two-channel case), row 0 = top row.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np


class UnsupportedImage(ValueError):
    pass


# ------------------------------------------------------------------------------------------------ PNG
_PNG_SIG = b"\x89PNG\r\n\x1a\n"


def _unfilter(raw, h, stride, bpp):
    """PNG scanline filters 0-4 (RFC 2083 section 6).  Sub is a running sum, Up a row add; Average and Paeth depend on the
    pixel to the left and are walked pixel by pixel (vectorised over the bpp bytes of a pixel)."""
    out = np.zeros((h, stride), dtype=np.uint8)
    prev = np.zeros(stride, dtype=np.int32)
    pos = 0
    for y in range(h):
        ft = raw[pos]
        line = np.frombuffer(raw, dtype=np.uint8, count=stride, offset=pos + 1).astype(np.int32)
        pos += stride + 1
        if ft == 0:
            cur = line
        elif ft == 1:
            cur = line.reshape(-1, bpp).cumsum(axis=0).reshape(-1) & 255
        elif ft == 2:
            cur = (line + prev) & 255
        elif ft in (3, 4):
            cur = np.zeros(stride, dtype=np.int32)
            left = np.zeros(bpp, dtype=np.int32)
            upleft = np.zeros(bpp, dtype=np.int32)
            for x in range(0, stride, bpp):
                up = prev[x:x + bpp]
                if ft == 3:
                    pred = (left + up) >> 1
                else:
                    p = left + up - upleft
                    pa, pb, pc = np.abs(p - left), np.abs(p - up), np.abs(p - upleft)
                    pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, upleft))
                left = (line[x:x + bpp] + pred) & 255
                cur[x:x + bpp] = left
                upleft = up
        else:
            raise UnsupportedImage(f"PNG filter type {ft}")
        out[y] = cur
        prev = cur
    return out


def decode_png(data: bytes) -> np.ndarray:
    if data[:8] != _PNG_SIG:
        raise UnsupportedImage("not a PNG")
    pos, idat, plte, trns, hdr = 8, [], None, None, None
    while pos < len(data):
        n, kind = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        pos += 12 + n
        if kind == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif kind == b"PLTE":
            plte = np.frombuffer(body, dtype=np.uint8).reshape(-1, 3)
        elif kind == b"tRNS":
            trns = np.frombuffer(body, dtype=np.uint8)
        elif kind == b"IDAT":
            idat.append(body)
        elif kind == b"IEND":
            break
    if hdr is None:
        raise UnsupportedImage("PNG without IHDR")
    w, h, depth, ctype, _, _, interlace = hdr
    packed = depth in (1, 2, 4) and ctype in (0, 3)        # several grey / palette samples per byte
    if (depth != 8 and not packed) or interlace != 0 or ctype not in (0, 2, 3, 4, 6):
        raise UnsupportedImage(f"PNG depth {depth} / colour type {ctype} / interlace {interlace}")
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    if packed:
        stride = (w * depth + 7) // 8
        rows = _unfilter(zlib.decompress(b"".join(idat)), h, stride, 1)
        bits = np.unpackbits(rows, axis=1)[:, :w * depth].reshape(h, w, depth)
        px = (bits * (1 << np.arange(depth - 1, -1, -1))).sum(-1).astype(np.uint8)[..., None]
        if ctype == 0:
            px = (px.astype(np.uint16) * 255 // ((1 << depth) - 1)).astype(np.uint8)
    else:
        px = _unfilter(zlib.decompress(b"".join(idat)), h, w * ch, ch).reshape(h, w, ch)
    if ctype == 3:
        if plte is None:
            raise UnsupportedImage("palette PNG without PLTE")
        rgb = plte[px[..., 0]]
        if trns is not None:
            a = np.full(256, 255, dtype=np.uint8)
            a[:len(trns)] = trns
            return np.ascontiguousarray(np.concatenate([rgb, a[px[..., 0]][..., None]], -1))
        return np.ascontiguousarray(rgb)
    if ctype == 4:       # grey + alpha -> RGBA (the upload path knows 1, 3 and 4 channels)
        return np.ascontiguousarray(np.concatenate([np.repeat(px[..., :1], 3, -1), px[..., 1:]], -1))
    return np.ascontiguousarray(px)


def encode_png(img: np.ndarray) -> bytes:
    """uint8 (h, w), (h, w, 1|3|4) -> PNG bytes (filter 0, one IDAT)."""
    a = np.ascontiguousarray(img, dtype=np.uint8)
    if a.ndim == 2:
        a = a[..., None]
    h, w, c = a.shape
    ctype = {1: 0, 3: 2, 4: 6}[c]
    raw = np.concatenate([np.zeros((h, 1), np.uint8), a.reshape(h, w * c)], axis=1).tobytes()

    def chunk(kind, body):
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)
    return (_PNG_SIG + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) +
            chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


# ------------------------------------------------------------------------------------------------ PPM / PGM
def decode_pnm(data: bytes) -> np.ndarray:
    if data[:2] not in (b"P5", b"P6"):
        raise UnsupportedImage("not a binary PGM / PPM")
    fields, pos = [], 2
    while len(fields) < 3:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            pos = data.index(b"\n", pos) + 1
            continue
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        fields.append(int(data[pos:end])); pos = end
    pos += 1
    w, h, mx = fields
    if mx != 255:
        raise UnsupportedImage("PNM maxval != 255")
    c = 1 if data[:2] == b"P5" else 3
    return np.frombuffer(data, dtype=np.uint8, count=w * h * c, offset=pos).reshape(h, w, c).copy()


# ------------------------------------------------------------------------------------------------ TGA
def decode_tga(data: bytes) -> np.ndarray:
    if len(data) < 18:
        raise UnsupportedImage("short TGA")
    idlen, cmap, itype = data[0], data[1], data[2]
    w, h, bpp, desc = struct.unpack("<HHBB", data[12:18])
    if cmap != 0 or itype not in (2, 3, 10, 11) or bpp not in (8, 24, 32) or w == 0 or h == 0:
        raise UnsupportedImage(f"TGA type {itype} / {bpp} bpp / colour map {cmap}")
    c = bpp // 8
    pos = 18 + idlen
    if itype in (2, 3):
        px = np.frombuffer(data, dtype=np.uint8, count=w * h * c, offset=pos).reshape(h * w, c)
    else:                                   # run-length packets
        out = np.empty((w * h, c), dtype=np.uint8)
        k = 0
        while k < w * h:
            head = data[pos]; pos += 1
            n = (head & 127) + 1
            if head & 128:
                out[k:k + n] = np.frombuffer(data, dtype=np.uint8, count=c, offset=pos); pos += c
            else:
                out[k:k + n] = np.frombuffer(data, dtype=np.uint8, count=n * c, offset=pos).reshape(n, c); pos += n * c
            k += n
        px = out
    px = px.reshape(h, w, c)
    if c >= 3:
        px = px[..., [2, 1, 0] + ([3] if c == 4 else [])]      # BGR(A) -> RGB(A)
    if not (desc & 0x20):
        px = px[::-1]                                            # bottom-left origin -> top row first
    if desc & 0x10:
        px = px[:, ::-1]
    return np.ascontiguousarray(px)


def load_image(path: str) -> np.ndarray:
    """File -> uint8 (h, w, c), c in {1, 3, 4}, top row first (the stbi_load convention, Model.h:150)."""
    data = open(path, "rb").read()
    if data[:8] == _PNG_SIG:
        return decode_png(data)
    if data[:2] in (b"P5", b"P6"):
        return decode_pnm(data)
    if path.lower().endswith(".tga"):
        return decode_tga(data)
    raise UnsupportedImage(f"{path}: format not handled by the built-in decoders")


def save_png(img: np.ndarray, path: str) -> None:
    with open(path, "wb") as f:
        f.write(encode_png(img))
