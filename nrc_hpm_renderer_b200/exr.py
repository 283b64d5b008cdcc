"""OpenEXR scanline I/O for RGBA float frames -- the on-disk format of the reference's frames.

The reference writes its output image with tinyexr's `SaveEXR(data, w, h, 4, /*fp16=*/0, path)`
(src/NrcHpmRenderer.cu:437-493, src/McHpmRenderer.cpp ExportOutputImageToFile): single-part scanline file, channels
A, B, G, R as 32-bit FLOAT, ZIP compression in blocks of 16 lines, increasing-Y line order -- the bundled
`reference/<scene>/0.exr` files have exactly this header.  `write_exr` produces the same flavour, `read_exr` reads it
(plus NONE / ZIPS compression and HALF / UINT channels, which other writers of small test images use).

Stand-alone (struct + zlib + numpy): neither tinyexr nor OpenEXR exists in this image.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

MAGIC = 20000630
NO_COMPRESSION, RLE_COMPRESSION, ZIPS_COMPRESSION, ZIP_COMPRESSION = 0, 1, 2, 3
_LINES_PER_BLOCK = {NO_COMPRESSION: 1, RLE_COMPRESSION: 1, ZIPS_COMPRESSION: 1, ZIP_COMPRESSION: 16}
_PIXEL_DTYPE = {0: np.dtype("<u4"), 1: np.dtype("<f2"), 2: np.dtype("<f4")}


class ExrFormatError(RuntimeError):
    pass


def _zip_reorder_decode(raw: bytes) -> bytes:
    """inverse of the EXR ZIP pre-processing: delta predictor, then the two interleaved halves"""
    t = np.frombuffer(raw, np.uint8).astype(np.int64)
    if t.size == 0:
        return b""
    t[1:] -= 128
    t = (np.cumsum(t) & 0xFF).astype(np.uint8)
    half = (t.size + 1) // 2
    out = np.empty_like(t)
    out[0::2] = t[:half]
    out[1::2] = t[half:]
    return out.tobytes()


def _zip_reorder_encode(raw: bytes) -> bytes:
    s = np.frombuffer(raw, np.uint8)
    t = np.concatenate([s[0::2], s[1::2]]).astype(np.int64)
    d = t.copy()
    d[1:] = (t[1:] - t[:-1] + 128 + 256) & 0xFF
    return d.astype(np.uint8).tobytes()


def _rle_decode(raw: bytes, expected: int) -> bytes:
    out = bytearray()
    i = 0
    while i < len(raw):
        c = struct.unpack_from("b", raw, i)[0]
        i += 1
        if c < 0:
            out += raw[i:i - c]
            i += -c
        else:
            out += raw[i:i + 1] * (c + 1)
            i += 1
    if len(out) != expected:
        raise ExrFormatError("RLE block has the wrong size")
    return bytes(out)


def _read_header(buf: bytes):
    magic, version = struct.unpack_from("<iI", buf, 0)
    if magic != MAGIC:
        raise ExrFormatError("not an OpenEXR file")
    if version & 0xFF != 2 or version & 0x1A00:        # tiled (0x200), non-image (0x800), multi-part (0x1000)
        raise ExrFormatError(f"unsupported OpenEXR flavour (version word {version:#x}); only single-part scanline files")
    p, attrs = 8, {}
    while buf[p] != 0:
        e = buf.index(b"\0", p); name = buf[p:e].decode("latin-1"); p = e + 1
        e = buf.index(b"\0", p); typ = buf[p:e].decode("latin-1"); p = e + 1
        size = struct.unpack_from("<i", buf, p)[0]; p += 4
        attrs[name] = (typ, buf[p:p + size]); p += size
    return attrs, p + 1


def _channels(blob: bytes):
    chans, p = [], 0
    while blob[p] != 0:
        e = blob.index(b"\0", p); name = blob[p:e].decode("latin-1"); p = e + 1
        ptype, _plinear, xs, ys = struct.unpack_from("<iB3xii", blob, p); p += 16
        if xs != 1 or ys != 1:
            raise ExrFormatError("sub-sampled channels are not supported")
        chans.append((name, ptype))
    return chans


def read_exr(path: str) -> np.ndarray:
    """-> float32 [H][W][4] in R, G, B, A order (missing channels: RGB 0, A 1)"""
    buf = open(path, "rb").read()
    attrs, p = _read_header(buf)
    chans = _channels(attrs["channels"][1])
    comp = attrs["compression"][1][0]
    if comp not in _LINES_PER_BLOCK:
        raise ExrFormatError(f"unsupported compression {comp} (NONE, RLE, ZIPS, ZIP are read)")
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    W, H = x1 - x0 + 1, y1 - y0 + 1
    lpb = _LINES_PER_BLOCK[comp]
    n_blocks = (H + lpb - 1) // lpb
    offsets = struct.unpack_from(f"<{n_blocks}Q", buf, p)
    line_bytes = sum(_PIXEL_DTYPE[t].itemsize for _, t in chans) * W
    planes = {name: np.zeros((H, W), np.float32) for name, _ in chans}
    for off in offsets:
        y, size = struct.unpack_from("<ii", buf, off)
        data = buf[off + 8:off + 8 + size]
        rows = min(lpb, y1 - y + 1)
        expected = rows * line_bytes
        if size < expected:
            if comp in (ZIP_COMPRESSION, ZIPS_COMPRESSION):
                data = _zip_reorder_decode(zlib.decompress(data))
            elif comp == RLE_COMPRESSION:
                data = _zip_reorder_decode(_rle_decode(data, expected))
        if len(data) != expected:
            raise ExrFormatError("scanline block has the wrong size")
        q = 0
        for r in range(rows):
            for name, t in chans:                       # channels are stored in alphabetical order, one row each
                dt = _PIXEL_DTYPE[t]
                planes[name][y - y0 + r] = np.frombuffer(data, dt, W, q).astype(np.float32)
                q += W * dt.itemsize
    out = np.zeros((H, W, 4), np.float32)
    out[..., 3] = 1.0
    for k, name in enumerate("RGBA"):
        if name in planes:
            out[..., k] = planes[name]
    return out


def _attr(name: str, typ: str, value: bytes) -> bytes:
    return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(value)) + value


def write_exr(path: str, image: np.ndarray, compression: int = ZIP_COMPRESSION) -> None:
    """image: [H][W][4] (R, G, B, A) -> the file tinyexr's SaveEXR(..., 4 components, fp32) writes"""
    img = np.ascontiguousarray(image, dtype=np.float32)
    if img.ndim != 3 or img.shape[2] != 4:
        raise ValueError("write_exr expects an [H][W][4] RGBA image")
    if compression not in (NO_COMPRESSION, ZIPS_COMPRESSION, ZIP_COMPRESSION):
        raise ValueError("write_exr supports NONE, ZIPS and ZIP compression")
    H, W, _ = img.shape
    chl = b"".join(n.encode() + b"\0" + struct.pack("<iB3xii", 2, 0, 1, 1) for n in "ABGR") + b"\0"
    box = struct.pack("<4i", 0, 0, W - 1, H - 1)
    header = struct.pack("<iI", MAGIC, 2)
    header += _attr("channels", "chlist", chl) + _attr("compression", "compression", bytes([compression]))
    header += _attr("dataWindow", "box2i", box) + _attr("displayWindow", "box2i", box)
    header += _attr("lineOrder", "lineOrder", b"\0") + _attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    header += _attr("screenWindowCenter", "v2f", struct.pack("<2f", 0, 0)) + _attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0"
    lpb = _LINES_PER_BLOCK[compression]
    abgr = img[..., [3, 2, 1, 0]].transpose(0, 2, 1)           # [H][channel A,B,G,R][W]
    chunks = []
    for y in range(0, H, lpb):
        raw = np.ascontiguousarray(abgr[y:y + lpb]).astype("<f4").tobytes()
        data = raw
        if compression != NO_COMPRESSION:
            z = zlib.compress(_zip_reorder_encode(raw), 6)
            if len(z) < len(raw):
                data = z
        chunks.append(struct.pack("<ii", y, len(data)) + data)
    table_at = len(header)
    pos = table_at + 8 * len(chunks)
    offsets = []
    for c in chunks:
        offsets.append(pos); pos += len(c)
    with open(path, "wb") as f:
        f.write(header + struct.pack(f"<{len(offsets)}Q", *offsets) + b"".join(chunks))
