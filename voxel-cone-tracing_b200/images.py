"""Dependency-free image decode / encode for the asset path.

The reference decodes its textures with the vendored stb_image v2.26 (`stbi_load(path, &w, &h, &n, 0)`,
/root/reference/Voxel_Cone_Tracing_Final/Model.h:152) and uploads them with the channel count the file has (RED / RGB /
RGBA, Model.h:159-169), top row first.  This module restates that decoder on numpy + zlib for the formats a scene's
textures come in: PNG (every colour type and bit depth, tRNS, Adam7 interlace), binary PGM / PPM, TGA (true-colour,
grey and colour-mapped, raw and RLE, 15/16/24/32 bit), BMP (1/4/8-bit palettised, 16/24/32-bit incl. bit fields) and
baseline / progressive JPEG.  It follows stb_image's RESULTS, quirks included (16-bit samples keep their high byte, the
TGA right-to-left bit is ignored, a PNM maxval below 255 is not rescaled, stb's own integer IDCT, chroma up-sampling and
YCbCr conversion for JPEG): tests/test_images_vs_stb.py compares every decoder byte for byte with the reference's
stb_image compiled from /root/reference (oracle/ref_stb.py).  Anything else (GIF, PSD, HDR, PIC, 12-bit or arithmetic
JPEG) raises UnsupportedImage; objloader then falls back to Pillow if it is installed.

Output: uint8 array (h, w, c), row 0 = top row.  c is what stbi_load reports (1, 2, 3 or 4) when `native_channels` is
set; by default grey+alpha is expanded to RGBA, because the reference's upload knows only 1, 3 and 4 channels
(Model.h:161-166 leaves `format` unset for n == 2).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np


class UnsupportedImage(ValueError):
    pass


def _finish(px, native_channels):
    px = np.ascontiguousarray(px, dtype=np.uint8)
    if px.shape[2] == 2 and not native_channels:
        px = np.concatenate([np.repeat(px[..., :1], 3, -1), px[..., 1:]], -1)
    return px


# ------------------------------------------------------------------------------------------------ PNG
_PNG_SIG = b"\x89PNG\r\n\x1a\n"
_ADAM7 = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)]   # x0, y0, dx, dy
_PNG_DEPTHS = {0: (1, 2, 4, 8, 16), 2: (8, 16), 3: (1, 2, 4, 8), 4: (8, 16), 6: (8, 16)}


def _unfilter(raw, pos, h, stride, bpp):
    """PNG scanline filters 0-4 (RFC 2083 section 6).  None / Sub / Up are numpy row operations (Sub is a running sum per
    byte lane); Average and Paeth depend on the reconstructed byte to the left and are walked byte by byte on plain
    Python integers, which is several times faster than small numpy slices."""
    out = np.zeros((h, stride), dtype=np.uint8)
    prev = bytes(stride)
    lanes = (-stride) % bpp
    for y in range(h):
        if pos + 1 + stride > len(raw):
            raise UnsupportedImage("PNG: not enough pixel data")
        ft = raw[pos]
        line = raw[pos + 1:pos + 1 + stride]
        pos += stride + 1
        if ft == 0:
            cur = bytes(line)
        elif ft == 1:
            a = np.frombuffer(line + bytes(lanes), dtype=np.uint8).astype(np.int32)
            cur = (a.reshape(-1, bpp).cumsum(axis=0).reshape(-1)[:stride] & 255).astype(np.uint8).tobytes()
        elif ft == 2:
            cur = ((np.frombuffer(line, dtype=np.uint8).astype(np.int32) + np.frombuffer(prev, dtype=np.uint8)) & 255).astype(np.uint8).tobytes()
        elif ft == 3:
            c = bytearray(stride)
            for x in range(min(bpp, stride)):
                c[x] = (line[x] + (prev[x] >> 1)) & 255
            for x in range(bpp, stride):
                c[x] = (line[x] + ((c[x - bpp] + prev[x]) >> 1)) & 255
            cur = bytes(c)
        elif ft == 4:
            c = bytearray(stride)
            for x in range(min(bpp, stride)):
                c[x] = (line[x] + prev[x]) & 255                 # left = upper left = 0: the predictor is `up`
            for x in range(bpp, stride):
                left, up, ul = c[x - bpp], prev[x], prev[x - bpp]
                p = left + up - ul
                pa, pb, pc = abs(p - left), abs(p - up), abs(p - ul)
                c[x] = (line[x] + (left if pa <= pb and pa <= pc else up if pb <= pc else ul)) & 255
            cur = bytes(c)
        else:
            raise UnsupportedImage(f"PNG filter type {ft}")
        out[y] = np.frombuffer(cur, dtype=np.uint8)
        prev = cur
    return out, pos


def _png_samples(raw, pos, w, h, depth, ch):
    """one (sub)image: unfilter + unpack to (h, w, ch) integer samples of `depth` bits"""
    stride = (w * ch * depth + 7) // 8
    rows, pos = _unfilter(raw, pos, h, stride, max(1, ch * depth // 8))
    if depth == 8:
        px = rows.reshape(h, w, ch).astype(np.uint16)
    elif depth == 16:
        px = rows.reshape(h, w, ch, 2).astype(np.uint16)
        px = (px[..., 0] << 8) | px[..., 1]
    else:
        bits = np.unpackbits(rows, axis=1)[:, :w * depth].reshape(h, w, depth)
        px = (bits.astype(np.uint16) << np.arange(depth - 1, -1, -1, dtype=np.uint16)).sum(-1, dtype=np.uint16)[..., None]
    return px, pos


def decode_png(data: bytes, native_channels: bool = False) -> np.ndarray:
    if data[:8] != _PNG_SIG:
        raise UnsupportedImage("not a PNG")
    pos, idat, plte, trns, hdr = 8, [], None, None, None
    while pos + 8 <= len(data):
        n, kind = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        pos += 12 + n
        if kind == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif kind == b"PLTE":
            plte = np.frombuffer(body, dtype=np.uint8).reshape(-1, 3)
        elif kind == b"tRNS":
            trns = bytes(body)
        elif kind == b"IDAT":
            idat.append(body)
        elif kind == b"IEND":
            break
    if hdr is None:
        raise UnsupportedImage("PNG without IHDR")
    w, h, depth, ctype, _, _, interlace = hdr
    if ctype not in _PNG_DEPTHS or depth not in _PNG_DEPTHS[ctype] or interlace not in (0, 1) or w == 0 or h == 0:
        raise UnsupportedImage(f"PNG depth {depth} / colour type {ctype} / interlace {interlace}")
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    raw = zlib.decompress(b"".join(idat))
    if interlace:
        px = np.zeros((h, w, ch), dtype=np.uint16)
        at = 0
        for x0, y0, dx, dy in _ADAM7:
            pw, ph = (w - x0 + dx - 1) // dx, (h - y0 + dy - 1) // dy
            if pw > 0 and ph > 0:
                sub, at = _png_samples(raw, at, pw, ph, depth, ch)
                px[y0::dy, x0::dx] = sub
    else:
        px, _ = _png_samples(raw, 0, w, h, depth, ch)
    if ctype == 3:
        if plte is None:
            raise UnsupportedImage("palette PNG without PLTE")
        pal = np.zeros((256, 4), dtype=np.uint8)
        pal[:, 3] = 255
        pal[:len(plte), :3] = plte
        if trns is not None:
            pal[:len(trns), 3] = np.frombuffer(trns, dtype=np.uint8)
        out = pal[px[..., 0] & 255]
        return _finish(out if trns is not None else out[..., :3], native_channels)
    if depth < 8:
        px = px * (255 // ((1 << depth) - 1))                    # stb: 1 -> 0xff, 2 -> 0x55, 4 -> 0x11
    if trns is not None and ctype in (0, 2):                      # colour key: one more channel, 0 where the pixel equals it
        key = np.array(struct.unpack(">" + "H" * ch, trns[:2 * ch]), dtype=np.uint16)
        if depth < 8:
            key = (key & 255) * (255 // ((1 << depth) - 1))
        elif depth == 8:
            key = key & 255
        full = np.uint16(65535 if depth == 16 else 255)
        alpha = np.where((px == key).all(-1), np.uint16(0), full)
        px = np.concatenate([px, alpha[..., None]], -1)
    if depth == 16:
        px = px >> 8                                             # stbi__convert_16_to_8 keeps the high byte
    return _finish(px, native_channels)


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
def decode_pnm(data: bytes, native_channels: bool = False) -> np.ndarray:
    """binary PGM / PPM.  Like stb_image: maxval must be <= 255 and the samples are NOT rescaled by it."""
    if data[:2] not in (b"P5", b"P6"):
        raise UnsupportedImage("not a binary PGM / PPM")
    fields, pos = [], 2
    while len(fields) < 3:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            while pos < len(data) and data[pos:pos + 1] not in (b"\n", b"\r"):
                pos += 1
            continue
        end = pos
        while data[end:end + 1].isdigit():
            end += 1
        if end == pos:
            raise UnsupportedImage("PNM header")
        fields.append(int(data[pos:end])); pos = end
    pos += 1                                                     # the single whitespace byte after maxval
    w, h, mx = fields
    if mx > 255 or w == 0 or h == 0:
        raise UnsupportedImage("PNM maxval > 255")
    c = 1 if data[:2] == b"P5" else 3
    if len(data) - pos < w * h * c:
        raise UnsupportedImage("PNM: not enough pixel data")
    return np.frombuffer(data, dtype=np.uint8, count=w * h * c, offset=pos).reshape(h, w, c).copy()


# ------------------------------------------------------------------------------------------------ TGA
def _tga_unpack(raw, bpp, grey):
    """(n, bytes per pixel) file pixels -> (n, c) channels.  16-bit grey is luminance + alpha; 15/16-bit colour is
    5-5-5 scaled by 255/31 (no alpha, as stb_image reads it); 24/32-bit are stored BGR(A)."""
    if bpp == 8:
        return raw
    if bpp in (15, 16) and not grey:
        v = raw[:, 0].astype(np.uint16) | (raw[:, 1].astype(np.uint16) << 8)
        rgb = np.stack([(v >> 10) & 31, (v >> 5) & 31, v & 31], -1)
        return ((rgb * 255) // 31).astype(np.uint8)
    if bpp == 16:
        return raw
    return raw[:, [2, 1, 0] + ([3] if bpp == 32 else [])]


def decode_tga(data: bytes, native_channels: bool = False) -> np.ndarray:
    if len(data) < 18:
        raise UnsupportedImage("short TGA")
    idlen, cmap, itype = data[0], data[1], data[2]
    pal_start, pal_len, pal_bits = struct.unpack("<HHB", data[3:8])
    w, h, bpp, desc = struct.unpack("<HHBB", data[12:18])
    rle, base = itype >= 8, itype & 7
    if cmap not in (0, 1) or base not in (1, 2, 3) or (base == 1) != (cmap == 1) or w == 0 or h == 0:
        raise UnsupportedImage(f"TGA type {itype} / colour map {cmap}")
    grey = base == 3
    if cmap:
        if bpp not in (8, 16) or pal_bits not in (8, 15, 16, 24, 32):
            raise UnsupportedImage(f"TGA colour map: {bpp}-bit indices, {pal_bits}-bit entries")
    elif bpp not in ((8, 16) if grey else (15, 16, 24, 32)):
        raise UnsupportedImage(f"TGA {bpp} bpp")
    pos = 18 + idlen
    pal = None
    if cmap:
        eb = (pal_bits + 7) // 8
        pal = _tga_unpack(np.frombuffer(data, dtype=np.uint8, count=pal_len * eb, offset=pos).reshape(pal_len, eb), pal_bits, False)
        pos += pal_len * eb
    pb = (bpp + 7) // 8
    n = w * h
    if not rle:
        if len(data) - pos < n * pb:
            raise UnsupportedImage("TGA: not enough pixel data")
        raw = np.frombuffer(data, dtype=np.uint8, count=n * pb, offset=pos).reshape(n, pb)
    else:                                   # run-length packets
        raw = np.empty((n, pb), dtype=np.uint8)
        k = 0
        while k < n:
            head = data[pos]; pos += 1
            m = min((head & 127) + 1, n - k)
            if head & 128:
                raw[k:k + m] = np.frombuffer(data, dtype=np.uint8, count=pb, offset=pos); pos += pb
            else:
                raw[k:k + m] = np.frombuffer(data, dtype=np.uint8, count=m * pb, offset=pos).reshape(m, pb); pos += m * pb
            k += m
    if cmap:
        idx = raw[:, 0].astype(np.int64) if pb == 1 else raw[:, 0].astype(np.int64) | (raw[:, 1].astype(np.int64) << 8)
        idx = idx - pal_start                                    # stb: index relative to the first stored entry ...
        px = pal[np.where((idx >= 0) & (idx < pal_len), idx, 0)]    # ... and entry 0 when out of range
    else:
        px = _tga_unpack(raw, bpp, grey)
    px = px.reshape(h, w, -1)
    if not (desc & 0x20):
        px = px[::-1]                                            # bottom-left origin -> top row first
    # bit 4 (right-to-left) is ignored, as stb_image v2.26 ignores it
    return _finish(px, native_channels)


# ------------------------------------------------------------------------------------------------ BMP
def _bmp_channel(v, mask):
    """stb_image's mask extraction: move the mask's top bit to bit 7, keep `bits` bits, replicate them to 8 bits."""
    if mask == 0:
        return None
    hi = mask.bit_length() - 1
    bits = bin(mask).count("1")
    if bits > 8:
        raise UnsupportedImage("BMP: channel mask wider than 8 bits")
    x = (v & np.uint32(mask)).astype(np.int64)
    x = (x >> (hi - 7)) if hi >= 7 else (x << (7 - hi))
    x = x >> (8 - bits)
    mul = (0, 0xff, 0x55, 0x49, 0x11, 0x21, 0x41, 0x81, 0x01)[bits]
    shr = (0, 0, 0, 1, 0, 2, 4, 6, 0)[bits]
    return ((x * mul) >> shr).astype(np.uint8)


def decode_bmp(data: bytes, native_channels: bool = False) -> np.ndarray:
    """Windows bitmaps the way stb_image reads them: core (12-byte) / INFO / V3 / V4 / V5 headers, 1/4/8-bit palettised,
    16/24/32-bit with the default or explicit channel masks, bottom-up or top-down; RLE is not supported (nor by stb).
    A 32-bit BI_RGB image whose alpha bytes are all zero comes out opaque.  With a 12-byte (OS/2) header stb_image's
    palette-size formula loses the last four entries (it then reads uninitialised memory); they are black here."""
    if data[:2] != b"BM" or len(data) < 26:
        raise UnsupportedImage("not a BMP")
    offset, hsz = struct.unpack("<II", data[10:18])
    if hsz not in (12, 40, 56, 108, 124):
        raise UnsupportedImage(f"BMP header size {hsz}")
    mr = mg = mb = ma = 0
    zero_alpha_means_opaque = False
    extra = 14
    if hsz == 12:
        w, h, planes, bpp = struct.unpack("<HHHH", data[18:26])
        compress = 0
    else:
        w, h, planes, bpp, compress = struct.unpack("<iiHHI", data[18:34])
        if compress in (1, 2):
            raise UnsupportedImage("BMP RLE")
        if hsz in (40, 56):
            if bpp in (16, 32):
                if compress == 0:
                    if bpp == 32:
                        mr, mg, mb, ma = 0xff0000, 0xff00, 0xff, 0xff000000
                        zero_alpha_means_opaque = True
                    else:
                        mr, mg, mb = 31 << 10, 31 << 5, 31
                elif compress == 3:
                    at = 14 + hsz
                    mr, mg, mb = struct.unpack("<III", data[at:at + 12])
                    extra += 12
                    if mr == mg == mb:
                        raise UnsupportedImage("bad BMP masks")
                else:
                    raise UnsupportedImage("bad BMP compression")
        else:
            mr, mg, mb, ma = struct.unpack("<IIII", data[54:70])
    if planes != 1 or w <= 0 or h == 0:
        raise UnsupportedImage("bad BMP")
    flip = h > 0
    h = abs(h)
    n = 3 if (bpp == 24 and ma == 0xff000000) else (4 if ma else 3)
    if bpp < 16:
        psize = (offset - extra - 24) // 3 if hsz == 12 else (offset - extra - hsz) >> 2
        if psize <= 0 or psize > 256 or bpp not in (1, 4, 8):
            raise UnsupportedImage("bad BMP palette / depth")
        es = 3 if hsz == 12 else 4
        pal = np.zeros((256, 3), dtype=np.uint8)
        pal[:psize] = np.frombuffer(data, dtype=np.uint8, count=psize * es, offset=14 + hsz).reshape(psize, es)[:, 2::-1]
        width = (w * bpp + 7) // 8
        stride = (width + 3) & ~3
        rows = np.frombuffer(data, dtype=np.uint8, count=stride * h, offset=offset).reshape(h, stride)[:, :width]
        if bpp == 8:
            idx = rows
        elif bpp == 4:
            idx = np.stack([rows >> 4, rows & 15], -1).reshape(h, -1)[:, :w]
        else:
            idx = np.unpackbits(rows, axis=1)[:, :w]
        px = pal[idx]
        if n == 4:
            px = np.concatenate([px, np.full((h, w, 1), 255, np.uint8)], -1)
    else:
        if bpp not in (16, 24, 32):
            raise UnsupportedImage(f"BMP {bpp} bpp")
        pb = bpp // 8
        stride = (w * pb + 3) & ~3
        if len(data) - offset < stride * h:
            raise UnsupportedImage("BMP: not enough pixel data")
        rows = np.frombuffer(data, dtype=np.uint8, count=stride * h, offset=offset).reshape(h, stride)[:, :w * pb].reshape(h, w, pb)
        if bpp == 24 or (bpp == 32 and (mb, mg, mr, ma) == (0xff, 0xff00, 0xff0000, 0xff000000)):
            px = rows[..., [2, 1, 0] + ([3] if bpp == 32 else [])]
            if bpp == 24 and n == 4:
                px = np.concatenate([px, np.full((h, w, 1), 255, np.uint8)], -1)
            alpha = px[..., 3] if bpp == 32 else None
            if n == 3:
                px = px[..., :3]
        else:
            if not (mr and mg and mb):
                raise UnsupportedImage("bad BMP masks")
            v = rows[..., 0].astype(np.uint32) | (rows[..., 1].astype(np.uint32) << 8)
            if bpp == 32:
                v |= (rows[..., 2].astype(np.uint32) << 16) | (rows[..., 3].astype(np.uint32) << 24)
            chans = [_bmp_channel(v, m) for m in (mr, mg, mb)]
            alpha = _bmp_channel(v, ma) if ma else None
            px = np.stack(chans + ([alpha] if n == 4 else []), -1)
        if n == 4 and zero_alpha_means_opaque and alpha is not None and not alpha.any():
            px = px.copy(); px[..., 3] = 255
    if flip:
        px = px[::-1]
    return _finish(px, native_channels)


# ------------------------------------------------------------------------------------------------ JPEG
_ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14,
                    21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53,
                    60, 61, 54, 47, 55, 62, 63], dtype=np.int64)


def _f2f(x):
    """stb's fixed-point constant: (int)(x * 4096 + 0.5) with x a float literal"""
    return int(float(np.float32(x) * np.float32(4096)) + 0.5)


_C = {k: _f2f(v) for k, v in dict(a=0.5411961, b=-1.847759065, c=0.765366865, d=1.175875602, e=0.298631336, f=2.053119869,
                                  g=3.072711026, h=1.501321110, i=-0.899976223, j=-2.562915447, k=-1.961570560,
                                  l=-0.390180644).items()}


def _idct_1d(s0, s1, s2, s3, s4, s5, s6, s7):
    """The even / odd butterfly of the integer 'islow' inverse DCT (IJG jidctint), constants scaled by 1 << 12.
    Returns the four even sums x0..x3 and the four odd sums t0..t3; the caller adds its rounding bias and shifts."""
    C = _C
    p1 = (s2 + s6) * C["a"]
    t2 = p1 + s6 * C["b"]
    t3 = p1 + s2 * C["c"]
    t0 = (s0 + s4) * 4096
    t1 = (s0 - s4) * 4096
    x0, x3, x1, x2 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    t0, t1, t2, t3 = s7, s5, s3, s1
    p3, p4, p1, p2 = t0 + t2, t1 + t3, t0 + t3, t1 + t2
    p5 = (p3 + p4) * C["d"]
    t0, t1, t2, t3 = t0 * C["e"], t1 * C["f"], t2 * C["g"], t3 * C["h"]
    p1 = p5 + p1 * C["i"]
    p2 = p5 + p2 * C["j"]
    p3 = p3 * C["k"]
    p4 = p4 * C["l"]
    return x0, x1, x2, x3, t0 + p1 + p3, t1 + p2 + p4, t2 + p2 + p3, t3 + p1 + p4


def _idct_blocks(coef):
    """(n, 64) dequantised coefficients in natural order -> (n, 8, 8) uint8 samples; columns first with two extra bits of
    precision (bias 512, >> 10), then rows (bias 65536 + (128 << 17), >> 17), clamped -- stb_image's arrangement."""
    d = coef.reshape(-1, 8, 8).astype(np.int64)
    x0, x1, x2, x3, t0, t1, t2, t3 = _idct_1d(*(d[:, k, :] for k in range(8)))
    x0, x1, x2, x3 = x0 + 512, x1 + 512, x2 + 512, x3 + 512
    v = np.stack([x0 + t3, x1 + t2, x2 + t1, x3 + t0, x3 - t0, x2 - t1, x1 - t2, x0 - t3], 1) >> 10
    x0, x1, x2, x3, t0, t1, t2, t3 = _idct_1d(*(v[:, :, k] for k in range(8)))
    bias = 65536 + (128 << 17)
    x0, x1, x2, x3 = x0 + bias, x1 + bias, x2 + bias, x3 + bias
    o = np.stack([x0 + t3, x1 + t2, x2 + t1, x3 + t0, x3 - t0, x2 - t1, x1 - t2, x0 - t3], 2) >> 17
    return np.clip(o, 0, 255).astype(np.uint8)


def _huffman_lut(sizes, values):
    """16-bit prefix table: entry = (code length << 8) | symbol, 0 where no code matches (T.81 Annex C code assignment)"""
    lut = np.zeros(65536, dtype=np.int32)
    code, k = 0, 0
    for length in range(1, 17):
        for _ in range(sizes[length - 1]):
            if k >= len(values) or code >= (1 << length):
                raise UnsupportedImage("bad JPEG Huffman table")
            lo = code << (16 - length)
            lut[lo:lo + (1 << (16 - length))] = (length << 8) | values[k]
            code += 1; k += 1
        code <<= 1
    return lut.tolist()


def _i16(v):
    return ((v + 32768) & 0xFFFF) - 32768          # stb keeps coefficients in `short`


class _Jpeg:
    pass


def _entropy_segments(data, pos):
    """Entropy-coded data from `pos` to the next marker that is not RSTn: the un-stuffed bytes of every restart interval
    and the position of that marker."""
    segs, cur = [], bytearray()
    n = len(data)
    while pos < n:
        nxt = data.find(b"\xff", pos)
        if nxt < 0 or nxt + 1 >= n:
            cur += data[pos:]
            pos = n
            break
        cur += data[pos:nxt]
        m = data[nxt + 1]
        if m == 0:
            cur.append(0xFF); pos = nxt + 2
        elif 0xD0 <= m <= 0xD7:
            segs.append(bytes(cur)); cur = bytearray(); pos = nxt + 2
        elif m == 0xFF:
            pos = nxt + 1                                # fill byte
        else:
            pos = nxt
            break
    segs.append(bytes(cur))
    return segs, pos


def _windows(seg):
    """32-bit big-endian window starting at every byte (zero padded past the end, as stb feeds zeros after a marker)"""
    a = np.frombuffer(seg + b"\0" * 12, dtype=np.uint8).astype(np.int64)
    return ((a[:-3] << 24) | (a[1:-2] << 16) | (a[2:-1] << 8) | a[3:]).tolist()


def _scan(j, comps, ss, se, ah, al, segs):
    """One scan (baseline: whole blocks; progressive: one spectral band / bit plane, T.81 Annex G) into the components'
    coefficient lists.  Interleaved scans walk MCUs, single-component scans walk that component's own blocks."""
    zz = _ZIGZAG.tolist()
    prog = j.progressive
    if len(comps) == 1:
        c = comps[0]
        units = [[(c, by * c.bw + bx)] for by in range((c.y + 7) >> 3) for bx in range((c.x + 7) >> 3)]
    else:
        units = []
        for my in range(j.mcu_y):
            for mx in range(j.mcu_x):
                units.append([(c, (my * c.v + y) * c.bw + mx * c.h + x) for c in comps for y in range(c.v) for x in range(c.h)])
    per = j.restart_interval or len(units)
    ui = 0
    for seg in segs:
        if ui >= len(units):
            break
        w = _windows(seg)
        limit = len(seg) * 8 + 64
        pos = 0
        eobrun = 0
        for c in comps:
            c.pred = 0
        for unit in units[ui:ui + per]:
            for c, b in unit:
                coef = c.coef
                base = b * 64
                if pos > limit:
                    raise UnsupportedImage("JPEG: entropy-coded data ends early")
                if not prog:
                    dq = c.dq
                    e = c.dc_lut[(w[pos >> 3] >> (16 - (pos & 7))) & 0xFFFF]
                    if not e:
                        raise UnsupportedImage("bad JPEG Huffman code")
                    pos += e >> 8
                    t = e & 255
                    if t:
                        v = (w[pos >> 3] >> (32 - (pos & 7) - t)) & ((1 << t) - 1)
                        pos += t
                        if v < (1 << (t - 1)):
                            v -= (1 << t) - 1
                        c.pred += v
                    coef[base] = _i16(c.pred * dq[0])
                    k = 1
                    lut = c.ac_lut
                    while k < 64:
                        e = lut[(w[pos >> 3] >> (16 - (pos & 7))) & 0xFFFF]
                        if not e:
                            raise UnsupportedImage("bad JPEG Huffman code")
                        pos += e >> 8
                        s = e & 15
                        r = (e >> 4) & 15
                        if s == 0:
                            if r != 15:
                                break
                            k += 16
                            continue
                        k += r
                        v = (w[pos >> 3] >> (32 - (pos & 7) - s)) & ((1 << s) - 1)
                        pos += s
                        if v < (1 << (s - 1)):
                            v -= (1 << s) - 1
                        if k > 63:
                            break
                        z = zz[k]
                        coef[base + z] = _i16(v * dq[z])
                        k += 1
                elif ss == 0:                                   # progressive DC: first pass or one more bit
                    if se != 0:
                        raise UnsupportedImage("JPEG: DC and AC in one progressive scan")
                    if ah == 0:
                        e = c.dc_lut[(w[pos >> 3] >> (16 - (pos & 7))) & 0xFFFF]
                        if not e:
                            raise UnsupportedImage("bad JPEG Huffman code")
                        pos += e >> 8
                        t = e & 255
                        if t:
                            v = (w[pos >> 3] >> (32 - (pos & 7) - t)) & ((1 << t) - 1)
                            pos += t
                            if v < (1 << (t - 1)):
                                v -= (1 << t) - 1
                            c.pred += v
                        coef[base] = _i16(c.pred << al)
                    else:
                        if (w[pos >> 3] >> (31 - (pos & 7))) & 1:
                            coef[base] = _i16(coef[base] + (1 << al))
                        pos += 1
                elif ah == 0:                                   # progressive AC, first pass over the band
                    if eobrun:
                        eobrun -= 1
                        continue
                    k = ss
                    lut = c.ac_lut
                    while k <= se:
                        e = lut[(w[pos >> 3] >> (16 - (pos & 7))) & 0xFFFF]
                        if not e:
                            raise UnsupportedImage("bad JPEG Huffman code")
                        pos += e >> 8
                        s = e & 15
                        r = (e >> 4) & 15
                        if s == 0:
                            if r < 15:
                                eobrun = 1 << r
                                if r:
                                    eobrun += (w[pos >> 3] >> (32 - (pos & 7) - r)) & ((1 << r) - 1)
                                    pos += r
                                eobrun -= 1
                                break
                            k += 16
                            continue
                        k += r
                        v = (w[pos >> 3] >> (32 - (pos & 7) - s)) & ((1 << s) - 1)
                        pos += s
                        if v < (1 << (s - 1)):
                            v -= (1 << s) - 1
                        if k > 63:
                            break
                        coef[base + zz[k]] = _i16(v << al)
                        k += 1
                else:                                           # progressive AC refinement (T.81 G.1.2.3)
                    bit = 1 << al
                    k = ss
                    if eobrun:
                        eobrun -= 1
                        while k <= se:
                            z = base + zz[k]
                            k += 1
                            p = coef[z]
                            if p:
                                if (w[pos >> 3] >> (31 - (pos & 7))) & 1 and not (p & bit):
                                    coef[z] = _i16(p + bit if p > 0 else p - bit)
                                pos += 1
                        continue
                    lut = c.ac_lut
                    while k <= se:
                        e = lut[(w[pos >> 3] >> (16 - (pos & 7))) & 0xFFFF]
                        if not e:
                            raise UnsupportedImage("bad JPEG Huffman code")
                        pos += e >> 8
                        s = e & 15
                        r = (e >> 4) & 15
                        if s == 0:
                            if r < 15:
                                eobrun = (1 << r) - 1
                                if r:
                                    eobrun += (w[pos >> 3] >> (32 - (pos & 7) - r)) & ((1 << r) - 1)
                                    pos += r
                                r = 64                           # the rest of the band only gets correction bits
                        else:
                            if s != 1:
                                raise UnsupportedImage("bad JPEG Huffman code")
                            s = bit if (w[pos >> 3] >> (31 - (pos & 7))) & 1 else -bit
                            pos += 1
                        while k <= se:
                            z = base + zz[k]
                            k += 1
                            p = coef[z]
                            if p:
                                if (w[pos >> 3] >> (31 - (pos & 7))) & 1 and not (p & bit):
                                    coef[z] = _i16(p + bit if p > 0 else p - bit)
                                pos += 1
                            else:
                                if r == 0:
                                    coef[z] = _i16(s)
                                    break
                                r -= 1
        ui += per


def _upsample(plane, c, j, W, H):
    """One component to full resolution the way stb_image walks it: per output row a near and a far source row
    (near weighs 3, far 1), then a horizontal 3:1 filter; factors other than 1 and 2 fall back to pixel repetition."""
    hs, vs = j.h_max // c.h, j.v_max // c.v
    wl = (W + hs - 1) // hs
    near, far = np.empty(H, np.int64), np.empty(H, np.int64)
    ystep, l0, l1, ypos = vs >> 1, 0, 0, 0
    for y in range(H):
        bot = ystep >= (vs >> 1)
        near[y], far[y] = (l1, l0) if bot else (l0, l1)
        ystep += 1
        if ystep >= vs:
            ystep, l0 = 0, l1
            ypos += 1
            if ypos < c.y:
                l1 += 1
    n = plane[near, :wl].astype(np.int64)
    if hs == 1 and vs == 1:
        return n[:, :W].astype(np.uint8)
    if hs == 1 and vs == 2:
        return ((3 * n + plane[far, :wl] + 2) >> 2)[:, :W].astype(np.uint8)
    if hs == 2 and vs in (1, 2):
        out = np.empty((H, 2 * wl), dtype=np.int64)
        if vs == 1:
            if wl == 1:
                out[:, 0] = out[:, 1] = n[:, 0]
            else:
                out[:, 0] = n[:, 0]
                out[:, 1] = (n[:, 0] * 3 + n[:, 1] + 2) >> 2
                m = 3 * n[:, 1:-1] + 2
                out[:, 2:-2:2] = (m + n[:, :-2]) >> 2
                out[:, 3:-2:2] = (m + n[:, 2:]) >> 2
                out[:, -2] = (n[:, -2] * 3 + n[:, -1] + 2) >> 2
                out[:, -1] = n[:, -1]
        else:
            t = 3 * n + plane[far, :wl]
            if wl == 1:
                out[:, 0] = out[:, 1] = (t[:, 0] + 2) >> 2
            else:
                out[:, 0] = (t[:, 0] + 2) >> 2
                out[:, 1:-1:2] = (3 * t[:, :-1] + t[:, 1:] + 8) >> 4
                out[:, 2::2] = (3 * t[:, 1:] + t[:, :-1] + 8) >> 4
                out[:, -1] = (t[:, -1] + 2) >> 2
        return out[:, :W].astype(np.uint8)
    return np.repeat(n, hs, axis=1)[:, :W].astype(np.uint8)


def _ycc_to_rgb(y, cb, cr):
    """stb_image's 20-bit fixed-point YCbCr -> RGB (the green chroma term is truncated to 16 fractional bits first)"""
    fx = lambda x: int(np.float32(x) * np.float32(4096.0) + np.float32(0.5)) << 8
    yf = (y.astype(np.int64) << 20) + (1 << 19)
    cr = cr.astype(np.int64) - 128
    cb = cb.astype(np.int64) - 128
    gb = (cb * -fx(0.34414)) & 0xFFFF0000
    gb = np.where(gb >= (1 << 31), gb - (1 << 32), gb)          # the mask is applied to a 32-bit int
    r = (yf + cr * fx(1.40200)) >> 20
    g = (yf + cr * -fx(0.71414) + gb) >> 20
    b = (yf + cb * fx(1.77200)) >> 20
    return np.clip(np.stack([r, g, b], -1), 0, 255).astype(np.uint8)


def _blinn(x, y):
    t = x.astype(np.int64) * y.astype(np.int64) + 128
    return ((t + (t >> 8)) >> 8).astype(np.uint8)


def decode_jpeg(data: bytes, native_channels: bool = False) -> np.ndarray:
    """Baseline and progressive Huffman JPEG, 8 bit, 1 / 3 / 4 components, any sampling factors, restart intervals.
    Returns (h, w, 1) for grey files and (h, w, 3) otherwise, like stbi_load with req_comp = 0."""
    if data[:2] != b"\xff\xd8":
        raise UnsupportedImage("not a JPEG")
    j = _Jpeg()
    j.progressive, j.restart_interval, j.jfif, j.adobe = False, 0, False, -1
    dequant, dc_luts, ac_luts = {}, {}, {}
    comps = None
    pos, n = 2, len(data)
    while True:
        while pos < n and data[pos] != 0xFF:
            pos += 1                                             # padding between segments
        while pos < n and data[pos] == 0xFF:
            pos += 1
        if pos >= n:
            break
        m = data[pos]; pos += 1
        if m == 0xD9:
            break
        if m == 0 or 0xD0 <= m <= 0xD7:
            continue
        L = struct.unpack(">H", data[pos:pos + 2])[0]
        body = data[pos + 2:pos + L]
        pos += L
        if m == 0xDB:
            at = 0
            while at < len(body):
                pq, tq = body[at] >> 4, body[at] & 15
                at += 1
                if pq > 1 or tq > 3:
                    raise UnsupportedImage("bad JPEG DQT")
                if pq:
                    q = np.frombuffer(body, dtype=">u2", count=64, offset=at).astype(np.int64); at += 128
                else:
                    q = np.frombuffer(body, dtype=np.uint8, count=64, offset=at).astype(np.int64); at += 64
                t = np.zeros(64, dtype=np.int64)
                t[_ZIGZAG] = q
                dequant[tq] = t.tolist()
        elif m == 0xC4:
            at = 0
            while at < len(body):
                tc, th = body[at] >> 4, body[at] & 15
                if tc > 1 or th > 3:
                    raise UnsupportedImage("bad JPEG DHT")
                sizes = list(body[at + 1:at + 17])
                cnt = sum(sizes)
                lut = _huffman_lut(sizes, list(body[at + 17:at + 17 + cnt]))
                (ac_luts if tc else dc_luts)[th] = lut
                at += 17 + cnt
        elif m == 0xDD:
            j.restart_interval = struct.unpack(">H", body[:2])[0]
        elif m == 0xE0 and body[:5] == b"JFIF\0":
            j.jfif = True
        elif m == 0xEE and len(body) >= 12 and body[:6] == b"Adobe\0":
            j.adobe = body[11]
        elif m in (0xC0, 0xC1, 0xC2):
            if comps is not None:
                raise UnsupportedImage("JPEG with several frames")
            j.progressive = m == 0xC2
            prec, H, W, nc = struct.unpack(">BHHB", body[:6])
            if prec != 8 or H == 0 or W == 0 or nc not in (1, 3, 4) or len(body) != 6 + 3 * nc:
                raise UnsupportedImage("JPEG: only 8-bit frames with a known height and 1, 3 or 4 components")
            comps = []
            for k in range(nc):
                c = _Jpeg()
                c.id, hv, c.tq = body[6 + 3 * k:9 + 3 * k]
                c.h, c.v = hv >> 4, hv & 15
                if not 1 <= c.h <= 4 or not 1 <= c.v <= 4 or c.tq > 3:
                    raise UnsupportedImage("bad JPEG sampling factors")
                comps.append(c)
            j.rgb = nc == 3 and [c.id for c in comps] == [82, 71, 66]
            j.h_max, j.v_max = max(c.h for c in comps), max(c.v for c in comps)
            j.mcu_x = (W + 8 * j.h_max - 1) // (8 * j.h_max)
            j.mcu_y = (H + 8 * j.v_max - 1) // (8 * j.v_max)
            for c in comps:
                c.x = (W * c.h + j.h_max - 1) // j.h_max
                c.y = (H * c.v + j.v_max - 1) // j.v_max
                c.bw, c.bh = j.mcu_x * c.h, j.mcu_y * c.v
                c.coef = [0] * (c.bw * c.bh * 64)
        elif m == 0xDA:
            if comps is None:
                raise UnsupportedImage("JPEG: scan before frame header")
            ns = body[0]
            if not 1 <= ns <= len(comps) or len(body) != 4 + 2 * ns:
                raise UnsupportedImage("bad JPEG SOS")
            scan = []
            for k in range(ns):
                cid, tabs = body[1 + 2 * k], body[2 + 2 * k]
                c = next((c for c in comps if c.id == cid), None)
                if c is None or (tabs >> 4) > 3 or (tabs & 15) > 3:
                    raise UnsupportedImage("bad JPEG SOS component")
                c.dc_lut, c.ac_lut = dc_luts.get(tabs >> 4), ac_luts.get(tabs & 15)
                scan.append(c)
            ss, se, a = body[1 + 2 * ns:4 + 2 * ns]
            ah, al = a >> 4, a & 15
            if j.progressive:
                if ss > 63 or se > 63 or ss > se or ah > 13 or al > 13 or (ss > 0 and ns != 1):
                    raise UnsupportedImage("bad JPEG progressive scan")
            elif ss != 0 or ah or al:
                raise UnsupportedImage("bad JPEG SOS")
            for c in scan:
                if c.tq not in dequant or (c.dc_lut is None and ss == 0 and ah == 0) or (c.ac_lut is None and (se > 0 or not j.progressive)):
                    raise UnsupportedImage("JPEG: scan uses a table that was never defined")
                c.dq = dequant[c.tq]
            segs, pos = _entropy_segments(data, pos)
            _scan(j, scan, ss, 63 if not j.progressive else se, ah, al, segs)
        elif 0xE0 <= m <= 0xEF or m == 0xFE:
            pass
        else:
            raise UnsupportedImage(f"JPEG marker 0x{m:02x} (arithmetic coding, lossless and hierarchical modes are not handled)")
    if comps is None:
        raise UnsupportedImage("JPEG without a frame")
    planes = []
    for c in comps:
        coef = np.array(c.coef, dtype=np.int64).reshape(-1, 64)
        if j.progressive:
            coef = ((coef * np.array(dequant[c.tq], dtype=np.int64) + 32768) & 0xFFFF) - 32768
        blocks = _idct_blocks(coef).reshape(c.bh, c.bw, 8, 8)
        plane = blocks.transpose(0, 2, 1, 3).reshape(c.bh * 8, c.bw * 8)
        planes.append(_upsample(plane, c, j, W, H))
    if len(comps) == 1:
        return planes[0][..., None]
    if len(comps) == 3:
        if j.rgb or (j.adobe == 0 and not j.jfif):
            return np.ascontiguousarray(np.stack(planes, -1))
        return _ycc_to_rgb(*planes)
    if j.adobe == 0:                                             # CMYK
        return np.stack([_blinn(planes[k], planes[3]) for k in range(3)], -1)
    rgb = _ycc_to_rgb(*planes[:3])
    if j.adobe == 2:                                             # YCCK
        return np.stack([_blinn(255 - rgb[..., k], planes[3]) for k in range(3)], -1)
    return rgb


def decode(data: bytes, name: str = "", native_channels: bool = False) -> np.ndarray:
    """stbi_load's format detection order for the formats handled here: JPEG, PNG, BMP, PNM by signature, TGA last
    (it has no signature; stb tests it after everything else)."""
    if data[:2] == b"\xff\xd8":
        return decode_jpeg(data, native_channels)
    if data[:8] == _PNG_SIG:
        return decode_png(data, native_channels)
    if data[:2] == b"BM":
        return decode_bmp(data, native_channels)
    if data[:2] in (b"P5", b"P6"):
        return decode_pnm(data, native_channels)
    if data[:4] in (b"GIF8", b"8BPS", b"#?RA") or data[:4] == b"\x53\x80\xf6\x34":
        raise UnsupportedImage(f"{name}: GIF / PSD / HDR / PIC are not handled by the built-in decoders")
    return decode_tga(data, native_channels)


def load_image(path: str, native_channels: bool = False) -> np.ndarray:
    """File -> uint8 (h, w, c), top row first (the stbi_load convention, Model.h:152)."""
    with open(path, "rb") as f:
        data = f.read()
    return decode(data, path, native_channels)


def encode_pnm(img: np.ndarray) -> bytes:
    """uint8 (h, w) / (h, w, 1) -> binary PGM, (h, w, 3) -> binary PPM"""
    a = np.ascontiguousarray(img, dtype=np.uint8)
    if a.ndim == 2:
        a = a[..., None]
    h, w, c = a.shape
    if c not in (1, 3):
        raise ValueError("PNM holds one or three channels")
    return (b"P5" if c == 1 else b"P6") + f"\n{w} {h}\n255\n".encode() + a.tobytes()


def save_pnm(img: np.ndarray, path: str) -> None:
    with open(path, "wb") as f:
        f.write(encode_pnm(img))


def save_png(img: np.ndarray, path: str) -> None:
    with open(path, "wb") as f:
        f.write(encode_png(img))
