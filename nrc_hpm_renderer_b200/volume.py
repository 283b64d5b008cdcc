"""Stand-alone reader for the bundled OpenVDB cloud -> dense 8-bit density grid.

The reference loads ``data/volume/wdas_cloud_quarter.vdb`` through OpenVDB and turns it into a
dense RGBA8 3-D texture (reference src/Texture3D.cpp:12-117, src/HpmScene.cpp:44):

* every *active* voxel / active tile value inside ``file_bbox`` is written to ``data[x][y][z]``
  (Texture3D.cpp:57-72), inactive space stays 0;
* the maximum must be exactly 1.0 (Texture3D.cpp:74);
* the texel is ``uint8(v * 255.0f)`` -- truncation, not rounding (Texture3D.cpp:106) -- at linear
  index ``i + W*j + W*H*k`` (x fastest, Texture3D.cpp:107).

OpenVDB itself (v10.0.0, needs TBB/Boost/Blosc) is not available, so this module parses the one
file flavour the reference ships: file version 223, ``Tree_float_5_4_3``, per-grid compression flag
``COMPRESS_ACTIVE_MASK`` only (no zip / blosc, no half floats).  Layout facts follow
openvdb/io/Archive.cc (header, grid descriptors), tree/RootNode.h, tree/InternalNode.h:2206-2258
(readTopology), tree/LeafNode.h (readTopology / readBuffers) and io/Compression.h:463-600
(readCompressedValues and its seven metadata codes).

The product keeps ONE byte per voxel (the reference replicates it into RGBA8, 4 B/voxel).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

# io/Compression.h:53-74
_COMPRESS_ZIP, _COMPRESS_ACTIVE_MASK, _COMPRESS_BLOSC = 0x1, 0x2, 0x4
(_NO_MASK_OR_INACTIVE_VALS, _NO_MASK_AND_MINUS_BG, _NO_MASK_AND_ONE_INACTIVE_VAL,
 _MASK_AND_NO_INACTIVE_VALS, _MASK_AND_ONE_INACTIVE_VAL, _MASK_AND_TWO_INACTIVE_VALS,
 _NO_MASK_AND_ALL_VALS) = range(7)


class VdbFormatError(RuntimeError):
    pass


@dataclass
class DenseVolume:
    """Dense density grid as the tracker consumes it."""
    data: np.ndarray          # uint8, shape (D, H, W): data[k, j, i] == texel (i, j, k)
    bbox_min: tuple           # file_bbox_min (index space)
    bbox_max: tuple
    active_voxels: int
    max_value: float

    @property
    def dims(self):           # (W, H, D) == (x, y, z) extents, reference VkExtent3D order
        d, h, w = self.data.shape
        return (w, h, d)


class _Cursor:
    def __init__(self, buf: bytes, pos: int = 0):
        self.buf, self.pos = buf, pos

    def take(self, n: int) -> bytes:
        b = self.buf[self.pos:self.pos + n]
        if len(b) != n:
            raise VdbFormatError("unexpected end of file")
        self.pos += n
        return b

    def u32(self): return struct.unpack("<I", self.take(4))[0]
    def i32(self): return struct.unpack("<i", self.take(4))[0]
    def i64(self): return struct.unpack("<q", self.take(8))[0]
    def i8(self): return struct.unpack("<b", self.take(1))[0]
    def f32(self): return struct.unpack("<f", self.take(4))[0]
    def string(self): return self.take(self.u32()).decode("latin-1")

    def mask(self, nbits: int) -> np.ndarray:
        """util::NodeMask::load -- 64-bit little-endian words, bit n = word n>>6, bit n&63."""
        raw = np.frombuffer(self.take(nbits // 8), dtype=np.uint8)
        return np.unpackbits(raw, bitorder="little").astype(bool)

    def floats(self, n: int) -> np.ndarray:
        return np.frombuffer(self.take(4 * n), dtype="<f4")


def _read_metamap(c: _Cursor) -> dict:
    out = {}
    for _ in range(c.u32()):
        name, typ = c.string(), c.string()
        payload = c.take(c.u32())
        if typ == "string":
            out[name] = payload.decode("latin-1")
        elif typ == "vec3i":
            out[name] = struct.unpack("<3i", payload)
        elif typ == "int64":
            out[name] = struct.unpack("<q", payload)[0]
        elif typ == "float":
            out[name] = struct.unpack("<f", payload)[0]
        elif typ == "bool":
            out[name] = bool(payload[0])
        else:
            out[name] = payload
    return out


def _read_compressed_values(c: _Cursor, value_mask: np.ndarray, background: float) -> np.ndarray:
    """io::readCompressedValues with COMPRESS_ACTIVE_MASK (Compression.h:463-600).

    Returns the node's value array with only ACTIVE positions guaranteed meaningful (inactive
    positions are reconstructed too, but the reference never looks at them)."""
    n = value_mask.size
    meta = c.i8()
    inactive1 = background
    inactive0 = background if meta == _NO_MASK_OR_INACTIVE_VALS else -background
    if meta in (_NO_MASK_AND_ONE_INACTIVE_VAL, _MASK_AND_ONE_INACTIVE_VAL, _MASK_AND_TWO_INACTIVE_VALS):
        inactive0 = c.f32()
        if meta == _MASK_AND_TWO_INACTIVE_VALS:
            inactive1 = c.f32()
    selection = None
    if meta in (_MASK_AND_NO_INACTIVE_VALS, _MASK_AND_ONE_INACTIVE_VAL, _MASK_AND_TWO_INACTIVE_VALS):
        selection = c.mask(n)
    if meta == _NO_MASK_AND_ALL_VALS:
        return c.floats(n).copy()
    vals = np.full(n, inactive0, dtype=np.float32)
    if selection is not None:
        vals[selection] = inactive1
    vals[value_mask] = c.floats(int(value_mask.sum()))
    return vals


def read_vdb_dense(path: str) -> DenseVolume:
    """Parse the VDB and return the dense u8 grid exactly as Texture3D::FromVDB would build it."""
    buf = open(path, "rb").read()
    c = _Cursor(buf)
    if c.take(8) != b" BDV\x00\x00\x00\x00":
        raise VdbFormatError("not a VDB file")
    version = c.u32()
    if version < 222:
        raise VdbFormatError(f"file version {version} < 222 not supported")
    c.u32(); c.u32()                      # library major / minor
    if c.take(1) != b"\x01":
        raise VdbFormatError("file has no grid offsets")
    c.take(36)                            # uuid
    _read_metamap(c)                      # file-level metadata
    n_grids = c.u32()
    grid = None
    for _ in range(n_grids):
        name, gtype = c.string(), c.string()
        c.string()                        # instance parent name
        grid_pos, block_pos, end_pos = c.i64(), c.i64(), c.i64()
        if gtype.startswith("Tree_float_5_4_3") and grid is None:
            grid = (name, grid_pos, block_pos, end_pos)
        c.pos = end_pos
    if grid is None:
        raise VdbFormatError("No density volume found in vdb file")     # Texture3D.cpp:42
    _, grid_pos, block_pos, end_pos = grid

    c.pos = grid_pos
    compression = c.u32()
    if compression & (_COMPRESS_ZIP | _COMPRESS_BLOSC):
        raise VdbFormatError("zip/blosc-compressed VDBs are not supported by the stand-alone reader")
    if not compression & _COMPRESS_ACTIVE_MASK:
        raise VdbFormatError("expected COMPRESS_ACTIVE_MASK")
    meta = _read_metamap(c)
    if meta.get("is_saved_as_half_float"):
        raise VdbFormatError("half-float VDBs are not supported")
    bmin, bmax = meta["file_bbox_min"], meta["file_bbox_max"]
    map_type = c.string()
    if map_type != "UniformScaleMap":
        raise VdbFormatError(f"unexpected transform {map_type}")
    c.take(5 * 3 * 8)                     # ScaleMap: 5 x Vec3d
    if c.i32() != 1:
        raise VdbFormatError("expected one buffer per leaf")

    W, H, D = (bmax[0] - bmin[0] + 1, bmax[1] - bmin[1] + 1, bmax[2] - bmin[2] + 1)
    dense = np.zeros((W, H, D), dtype=np.float32)      # [x][y][z] like the reference's vector
    lo = np.array(bmin, dtype=np.int64)

    def fill_box(origin, size, value):
        a = np.maximum(np.array(origin) - lo, 0)
        b = np.minimum(np.array(origin) + size - lo, (W, H, D))
        if np.all(b > a):
            dense[a[0]:b[0], a[1]:b[1], a[2]:b[2]] = value

    # ---- topology (RootNode -> level-5 -> level-4 -> leaf masks) ----
    background = c.f32()
    n_tiles, n_children = c.u32(), c.u32()
    active = 0
    for _ in range(n_tiles):
        org = struct.unpack("<3i", c.take(12)); val = c.f32(); on = c.take(1)[0]
        if on:
            fill_box(org, 4096, val); active += 4096 ** 3
    leaves = []                                         # (origin, value_mask) in file order

    def coords(n, log2dim):
        m = (1 << log2dim) - 1
        return (n >> (2 * log2dim)) & m, (n >> log2dim) & m, n & m

    def read_internal(origin, log2dim, child_dim):
        nonlocal active
        nbits = 1 << (3 * log2dim)
        child_mask, value_mask = c.mask(nbits), c.mask(nbits)
        vals = _read_compressed_values(c, value_mask, background)
        for n in np.nonzero(value_mask & ~child_mask)[0]:        # active tiles
            x, y, z = coords(int(n), log2dim)
            fill_box((origin[0] + x * child_dim, origin[1] + y * child_dim, origin[2] + z * child_dim),
                     child_dim, vals[n])
            active += child_dim ** 3
        for n in np.nonzero(child_mask)[0]:
            x, y, z = coords(int(n), log2dim)
            corg = (origin[0] + x * child_dim, origin[1] + y * child_dim, origin[2] + z * child_dim)
            if child_dim == 128:
                read_internal(corg, 4, 8)
            else:
                leaves.append((corg, c.mask(512)))

    for _ in range(n_children):
        org = struct.unpack("<3i", c.take(12))
        read_internal(org, 5, 128)
    if c.pos != block_pos:
        raise VdbFormatError(f"topology ended at {c.pos}, expected blockPos {block_pos}")

    # ---- leaf buffers ----
    lx, ly, lz = np.meshgrid(np.arange(8), np.arange(8), np.arange(8), indexing="ij")
    lx, ly, lz = lx.ravel(), ly.ravel(), lz.ravel()              # offset = x<<6 | y<<3 | z
    for org, _topo_mask in leaves:
        vmask = c.mask(512)
        vals = _read_compressed_values(c, vmask, background)
        active += int(vmask.sum())
        gx, gy, gz = org[0] - lo[0] + lx, org[1] - lo[1] + ly, org[2] - lo[2] + lz
        ok = vmask & (gx >= 0) & (gx < W) & (gy >= 0) & (gy < H) & (gz >= 0) & (gz < D)
        dense[gx[ok], gy[ok], gz[ok]] = vals[ok]
    if c.pos != end_pos:
        raise VdbFormatError(f"buffers ended at {c.pos}, expected endPos {end_pos}")

    max_val = float(dense.max())
    if max_val != 0.0 and max_val != 1.0:
        raise VdbFormatError("VDB is not normalized")               # Texture3D.cpp:74
    u8 = (dense * np.float32(255.0)).astype(np.uint8)               # truncation, Texture3D.cpp:106
    # store z-major so that flat index == i + W*j + W*H*k (Texture3D.cpp:107)
    return DenseVolume(np.ascontiguousarray(u8.transpose(2, 1, 0)), tuple(bmin), tuple(bmax), active, max_val)


def save_volume(vol: DenseVolume, path: str) -> None:
    np.savez_compressed(path, data=vol.data, bbox_min=np.array(vol.bbox_min), bbox_max=np.array(vol.bbox_max),
                        active_voxels=np.int64(vol.active_voxels), max_value=np.float32(vol.max_value))


def load_volume(path: str) -> DenseVolume:
    z = np.load(path)
    return DenseVolume(np.ascontiguousarray(z["data"]), tuple(int(v) for v in z["bbox_min"]),
                       tuple(int(v) for v in z["bbox_max"]), int(z["active_voxels"]), float(z["max_value"]))


def synthetic_cloud(dims=(128, 96, 160), seed: int = 7) -> DenseVolume:
    """Procedural stand-in (sum of soft blobs) for boxes that have no VDB; NOT the benchmark volume."""
    w, h, d = dims
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.linspace(-1, 1, d), np.linspace(-1, 1, h), np.linspace(-1, 1, w), indexing="ij")
    f = np.zeros((d, h, w), dtype=np.float32)
    for _ in range(24):
        cx, cy, cz = rng.uniform(-0.6, 0.6, 3)
        r = rng.uniform(0.15, 0.45)
        f += np.exp(-(((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2) / (r * r))).astype(np.float32)
    f = np.clip(f - 0.6, 0, None)
    f /= f.max()
    u8 = (f * np.float32(255.0)).astype(np.uint8)
    return DenseVolume(u8, (0, 0, 0), (w - 1, h - 1, d - 1), int((u8 > 0).sum()), 1.0)
