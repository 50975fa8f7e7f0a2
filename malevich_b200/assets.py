"""Asset IO and deterministic stand-in assets (host side, numpy only).

* `.octrn` reader: restates the on-disk format the reference reads through the binary-only
  octarine libs (reference main.c:526-559; format in SURVEY.md App. B).
* sRGB->linear texture re-quantisation: restates reference main.c:546-558 with
  math.h:322-334,386-395 (decode, srgb_to_linear, encode-by-truncation).
* Stand-ins for the 12 assets that are missing from the reference checkout
  (`.MISSING_LARGE_BLOBS`): every 8-bit texture, emily_head_mesh, locomotive_mesh and the
  radiance panorama.  All stand-ins are pure functions of their arguments so the oracle and
  the GPU path consume identical bytes.
"""
from __future__ import annotations

import os
import struct

import numpy as np

ASSET_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")

OCTRN_MAGIC = b"eniratco"
OCTRN_TYPE_IMAGE = 0
OCTRN_TYPE_MESH = 1
FORMAT_R32G32B32A32_FLOAT = 0x4801


class OctrnError(ValueError):
    pass


def _read_header(blob: bytes, want_type: int) -> None:
    if len(blob) < 16 or blob[:8] != OCTRN_MAGIC:
        raise OctrnError("not an octarine asset (bad magic)")
    (asset_type,) = struct.unpack_from("<I", blob, 8)
    if asset_type != want_type:
        raise OctrnError(f"asset type {asset_type}, expected {want_type}")


def read_octrn_mesh(path: str):
    """-> (vertex_buffer float32 [num_vertices, 8], index_buffer uint32 [num_indices]).

    Layout (reference main.c:531-535): 32-byte vertices (pos3, normal3, uv2) followed by u32 indices.
    """
    with open(path, "rb") as f:
        blob = f.read()
    _read_header(blob, OCTRN_TYPE_MESH)
    size_of_data, num_vertices, num_indices = struct.unpack_from("<III", blob, 16)
    if 28 + size_of_data != len(blob) or num_vertices * 32 + num_indices * 4 != size_of_data:
        raise OctrnError("mesh header inconsistent with file size")
    vb = np.frombuffer(blob, dtype=np.float32, count=num_vertices * 8, offset=28).reshape(num_vertices, 8).copy()
    ib = np.frombuffer(blob, dtype=np.uint32, count=num_indices, offset=28 + num_vertices * 32).copy()
    return vb, ib


def read_octrn_image(path: str):
    """-> (texels, width, height, format). RGBA32F images come back as float32 [h, w, 4],
    8-bit ones as uint32 [h, w]."""
    with open(path, "rb") as f:
        blob = f.read()
    _read_header(blob, OCTRN_TYPE_IMAGE)
    size_of_data, fmt, width, height, depth, array_size, mips, flags = struct.unpack_from("<QIHHHHHH", blob, 16)
    if 40 + size_of_data != len(blob):
        raise OctrnError("image header inconsistent with file size")
    if fmt == FORMAT_R32G32B32A32_FLOAT:
        data = np.frombuffer(blob, dtype=np.float32, count=width * height * 4, offset=40).reshape(height, width, 4).copy()
    else:
        data = np.frombuffer(blob, dtype=np.uint32, count=width * height, offset=40).reshape(height, width).copy()
    return data, width, height, fmt


def write_octrn_image(path: str, texels: np.ndarray) -> None:
    """Writes an image asset the readers here and in host/octrn.c accept: float32 [h, w, 4] as R32G32B32A32_FLOAT,
    uint32 [h, w] as a 4-byte-per-texel format. (Stand-ins for the missing *_tex.octrn files go to disk this way when the
    C host loads a scene through the reference's own init() table.)"""
    t = np.ascontiguousarray(texels)
    fmt = FORMAT_R32G32B32A32_FLOAT if t.dtype == np.float32 else 0x1C01
    h, w = t.shape[0], t.shape[1]
    with open(path, "wb") as f:
        f.write(OCTRN_MAGIC + struct.pack("<II", OCTRN_TYPE_IMAGE, 0))
        f.write(struct.pack("<QIHHHHHH", t.nbytes, fmt, w, h, 1, 1, 1, 0))
        f.write(t.tobytes())


def write_octrn_mesh(path: str, vb: np.ndarray, ib: np.ndarray) -> None:
    v = np.ascontiguousarray(vb, dtype=np.float32).reshape(-1, 8)
    i = np.ascontiguousarray(ib, dtype=np.uint32)
    with open(path, "wb") as f:
        f.write(OCTRN_MAGIC + struct.pack("<II", OCTRN_TYPE_MESH, 0))
        f.write(struct.pack("<III", v.nbytes + i.nbytes, v.shape[0], i.shape[0]))
        f.write(v.tobytes())
        f.write(i.tobytes())


def load_mesh(name: str):
    return read_octrn_mesh(os.path.join(ASSET_DIR, name + ".octrn"))


def load_irradiance():
    data, w, h, fmt = read_octrn_image(os.path.join(ASSET_DIR, "ninomaru_teien_panorama_irradiance.octrn"))
    assert fmt == FORMAT_R32G32B32A32_FLOAT
    return data


# ---------------------------------------------------------------------------------------------
# sRGB -> linear re-quantisation (reference main.c:546-558)

def _srgb_to_linear_lut() -> np.ndarray:
    """256-entry byte->byte table equal to encode(srgb_to_linear(decode(b))) of the reference."""
    b = np.arange(256, dtype=np.uint32)
    normalizer = np.float32(1.0 / 255.0)                       # math.h:328  f32 normalizer = 1.0/255.0
    v = (b.astype(np.float32) * normalizer).astype(np.float32)  # math.h:329  (u32 -> f32) * f32
    vd = v.astype(np.float64)
    lin = np.where(vd <= 0.04045, vd / 12.92, np.power((vd + 0.055) / 1.055, 2.4))  # math.h:386-395 (double math)
    lin32 = lin.astype(np.float32)                             # f32 result
    return (lin32 * np.float32(255.0)).astype(np.float32).astype(np.uint32)  # math.h:323 (u32)(c*255.f), truncation


def srgb_texture_to_linear(tex_u32: np.ndarray) -> np.ndarray:
    """All four channels go through the curve, alpha included (reference main.c:551-554)."""
    lut = _srgb_to_linear_lut()
    t = tex_u32.astype(np.uint32)
    return (lut[t & 0xFF] | (lut[(t >> 8) & 0xFF] << 8) | (lut[(t >> 16) & 0xFF] << 16) | (lut[(t >> 24) & 0xFF] << 24)).astype(np.uint32)


# ---------------------------------------------------------------------------------------------
# stand-ins

def standin_texture_srgb(seed: int = 0, size: int = 1024) -> np.ndarray:
    """R8G8B8A8 stand-in for the missing *_tex.octrn files:
    R = 255x/(size-1), G = 255y/(size-1), B = checker(16 px) ? 200 : 60, A = 255; `seed` rotates the
    channels / checker phase so that each scene object gets a distinct texture. Row 0 = top."""
    y, x = np.mgrid[0:size, 0:size].astype(np.uint32)
    r = (255 * x) // (size - 1)
    g = (255 * y) // (size - 1)
    chk = (((x + 5 * seed) >> 4) ^ ((y + 3 * seed) >> 4)) & 1
    b = np.where(chk == 1, 200, 60).astype(np.uint32)
    chans = [r, g, b]
    k = seed % 3
    chans = chans[k:] + chans[:k]
    return (chans[0] | (chans[1] << 8) | (chans[2] << 16) | (np.uint32(255) << 24)).astype(np.uint32)


def standin_texture(seed: int = 0, size: int = 1024) -> np.ndarray:
    """Stand-in texture after the reference's load path (is_in_srgb=true, main.c:1320)."""
    return srgb_texture_to_linear(standin_texture_srgb(seed, size))


def _hash_noise(h: int, w: int, seed: int) -> np.ndarray:
    y, x = np.mgrid[0:h, 0:w].astype(np.uint32)
    v = (x * np.uint32(0x9E3779B1)) ^ (y * np.uint32(0x85EBCA77)) ^ np.uint32(seed * 0xC2B2AE3D & 0xFFFFFFFF)
    v ^= v >> 15
    v = (v * np.uint32(0x2C1B3C6D)).astype(np.uint32)
    v ^= v >> 12
    v = (v * np.uint32(0x297A2D39)).astype(np.uint32)
    v ^= v >> 15
    return (v >> 8).astype(np.float32) * np.float32(1.0 / (1 << 24))


def standin_radiance(irradiance: np.ndarray, width: int = 2048, height: int = 1024, seed: int = 1234) -> np.ndarray:
    """RGBA32F stand-in for ninomaru_teien_panorama_radiance: bilinear upsample of the real
    irradiance panorama times (1 + 0.5 * hash_noise)."""
    ih, iw, _ = irradiance.shape
    xs = (np.arange(width, dtype=np.float64) + 0.5) * iw / width - 0.5
    ys = (np.arange(height, dtype=np.float64) + 0.5) * ih / height - 0.5
    x0 = np.floor(xs).astype(np.int64)
    y0 = np.floor(ys).astype(np.int64)
    fx = (xs - x0)[None, :, None]
    fy = (ys - y0)[:, None, None]
    x0c, x1c = np.clip(x0, 0, iw - 1), np.clip(x0 + 1, 0, iw - 1)
    y0c, y1c = np.clip(y0, 0, ih - 1), np.clip(y0 + 1, 0, ih - 1)
    irr = irradiance.astype(np.float64)
    top = irr[y0c][:, x0c] * (1 - fx) + irr[y0c][:, x1c] * fx
    bot = irr[y1c][:, x0c] * (1 - fx) + irr[y1c][:, x1c] * fx
    up = top * (1 - fy) + bot * fy
    noise = _hash_noise(height, width, seed).astype(np.float64)[:, :, None]
    out = up * (1.0 + 0.5 * noise)
    out[:, :, 3] = 1.0
    return out.astype(np.float32)


def synthetic_grid_layer(layer: int, width: int, height: int, nx: int = 1250, ny: int = 500):
    """Config 5 geometry: one wavy grid of nx*ny quads facing the default camera at
    distance d = 2 + 0.25*layer + 0.05*sin(40u+layer)*cos(31v), 10% overscan.
    -> (vb float32 [(nx+1)*(ny+1), 8], ib uint32 [nx*ny*6]); index count is a multiple of 8 for the
    default 1250x500 (3 750 000)."""
    u = np.linspace(0.0, 1.0, nx + 1, dtype=np.float64)[None, :]
    v = np.linspace(0.0, 1.0, ny + 1, dtype=np.float64)[:, None]
    tan_theta = np.tan(np.deg2rad(37.5))
    d = 2.0 + 0.25 * layer + 0.05 * np.sin(40.0 * u + layer) * np.cos(31.0 * v)
    vb = np.zeros((ny + 1, nx + 1, 8), dtype=np.float32)
    vb[:, :, 0] = 3.5 - d
    vb[:, :, 1] = 1.0 + (u - 0.5) * 2.2 * d * tan_theta * width / height
    vb[:, :, 2] = 1.0 + (v - 0.5) * 2.2 * d * tan_theta
    vb[:, :, 3] = 1.0
    vb[:, :, 6] = u
    vb[:, :, 7] = v
    j, i = np.mgrid[0:ny, 0:nx].astype(np.uint32)
    a = j * np.uint32(nx + 1) + i
    b = a + 1
    c = a + np.uint32(nx + 1)
    dd = c + 1
    ib = np.stack([a, b, c, b, dd, c], axis=-1).reshape(-1).astype(np.uint32)
    return vb.reshape(-1, 8), ib


def pad_indices_to_8(ib: np.ndarray) -> np.ndarray:
    """The reference requires index_count % 8 == 0 (main.c:670) and pads its embedded scenes with
    degenerate (0,0,0) triangles (main.c:251-253); do the same, keeping index_count % 3 == 0."""
    n = ib.shape[0]
    assert n % 3 == 0
    while n % 24:
        n += 3
    out = np.zeros(n, dtype=np.uint32)
    out[: ib.shape[0]] = ib
    return out


def uv_sphere(center=(2.3, 1.0, 1.0), radius: float = 0.5, n_lat: int = 256, n_lon: int = 512):
    """EMILY stand-in: indexed UV sphere with outward normals, wound so the reference keeps the
    outside (signed_area <= 0, main.c:856)."""
    lat = np.linspace(0.0, np.pi, n_lat + 1, dtype=np.float64)[:, None]
    lon = np.linspace(0.0, 2.0 * np.pi, n_lon + 1, dtype=np.float64)[None, :]
    nxn = np.sin(lat) * np.cos(lon)
    nyn = np.sin(lat) * np.sin(lon)
    nzn = np.cos(lat) * np.ones_like(lon)
    vb = np.zeros((n_lat + 1, n_lon + 1, 8), dtype=np.float32)
    vb[:, :, 0] = center[0] + radius * nxn
    vb[:, :, 1] = center[1] + radius * nyn
    vb[:, :, 2] = center[2] + radius * nzn
    vb[:, :, 3] = nxn
    vb[:, :, 4] = nyn
    vb[:, :, 5] = nzn
    vb[:, :, 6] = (lon / (2.0 * np.pi)) * np.ones_like(lat)
    vb[:, :, 7] = (lat / np.pi) * np.ones_like(lon)
    j, i = np.mgrid[0:n_lat, 0:n_lon].astype(np.uint32)
    a = j * np.uint32(n_lon + 1) + i
    b = a + 1
    c = a + np.uint32(n_lon + 1)
    d = c + 1
    ib = np.stack([a, c, b, b, c, d], axis=-1).reshape(-1).astype(np.uint32)
    return vb.reshape(-1, 8), pad_indices_to_8(ib)


def torus_knot(center=(1.6, 1.0, 1.0), scale: float = 0.45, tube: float = 0.12, p: int = 2, q: int = 3,
               n_u: int = 4096, n_v: int = 128):
    """LOCOMOTIVE stand-in: (p,q) torus knot tube, n_u*n_v*2 triangles (1 048 576 by default),
    indexed, outward normals."""
    u = np.linspace(0.0, 2.0 * np.pi, n_u + 1, dtype=np.float64)
    r = 2.0 + np.cos(q * u)
    cx, cy, cz = r * np.cos(p * u), r * np.sin(p * u), -np.sin(q * u)
    cpos = np.stack([cx, cy, cz], -1) / 3.0
    tang = np.gradient(cpos, u, axis=0, edge_order=2)
    tang /= np.linalg.norm(tang, axis=1, keepdims=True)
    ref = np.array([0.0, 0.0, 1.0])
    nrm = np.cross(tang, ref)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    bin_ = np.cross(tang, nrm)
    v = np.linspace(0.0, 2.0 * np.pi, n_v + 1, dtype=np.float64)
    cv, sv = np.cos(v)[None, :, None], np.sin(v)[None, :, None]
    n3 = nrm[:, None, :] * cv + bin_[:, None, :] * sv
    pos = cpos[:, None, :] * scale + n3 * tube * scale
    vb = np.zeros((n_u + 1, n_v + 1, 8), dtype=np.float32)
    vb[:, :, 0:3] = pos + np.asarray(center)[None, None, :]
    vb[:, :, 3:6] = n3
    vb[:, :, 6] = (u / (2 * np.pi))[:, None]
    vb[:, :, 7] = (v / (2 * np.pi))[None, :]
    j, i = np.mgrid[0:n_u, 0:n_v].astype(np.uint32)
    a = j * np.uint32(n_v + 1) + i
    b = a + 1
    c = a + np.uint32(n_v + 1)
    d = c + 1
    ib = np.stack([a, c, b, b, c, d], axis=-1).reshape(-1).astype(np.uint32)
    return vb.reshape(-1, 8), pad_indices_to_8(ib)


def suprematist_scene():
    """The reference's embedded, asset-free SUPREMATISM scene (main.c:232-254): 11 vertices of
    (pos4, color3, pad) and 24 indices, the last 9 of which are (0,0,0) padding triangles."""
    v = np.array([
        [0.34107, 0.12215, 0.5, 1.0, 0.07500, 0.08200, 0.06300, 0.0],
        [0.95357, 0.12500, 0.5, 1.0, 0.07500, 0.08200, 0.06300, 0.0],
        [0.96250, 0.86931, 0.5, 1.0, 0.07500, 0.08200, 0.06300, 0.0],
        [0.33928, 0.86505, 0.5, 1.0, 0.07500, 0.08200, 0.06300, 0.0],
        [0.09464, 0.12500, 0.75, 1.0, 0.14100, 0.29000, 0.60800, 0.0],
        [0.69285, 0.39772, 0.75, 1.0, 0.14100, 0.29000, 0.60800, 0.0],
        [0.09107, 0.60937, 0.75, 1.0, 0.14100, 0.29000, 0.60800, 0.0],
        [0.00000, 0.00000, 0.25, 1.0, 0.96100, 0.96100, 0.92900, 0.0],
        [1.00000, 0.00000, 0.25, 1.0, 0.96100, 0.96100, 0.92900, 0.0],
        [1.00000, 1.00000, 0.25, 1.0, 0.96100, 0.96100, 0.92900, 0.0],
        [0.00000, 1.00000, 0.25, 1.0, 0.96100, 0.96100, 0.92900, 0.0],
    ], dtype=np.float64).astype(np.float32)
    ib = np.array([0, 1, 2, 2, 3, 0, 4, 5, 6, 7, 8, 9, 9, 10, 7] + [0] * 9, dtype=np.uint32)
    return v, ib


def fullscreen_quad():
    """The reference's embedded fullscreen quad (main.c:255-270) drawn with fullscreen_vs."""
    v = np.array([
        [0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0],
        [1.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0],
        [1.0, 1.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0],
        [0.0, 1.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0],
    ], dtype=np.float32)
    ib = np.array([0, 1, 2, 2, 3, 0] + [0] * 18, dtype=np.uint32)
    return v, ib
