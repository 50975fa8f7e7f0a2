"""Mip chain + trilinear sampling -- an EXTENSION: the reference samples level 0 bilinearly and has no
mips (common_shader_core.h:195-199), so there is no reference behaviour to pin ("parity unpinned"). What is pinned:
the chain against a numpy restatement of its definition (bit-exact), basic_trilinear_ps == basic_ps wherever nothing is
minified (bit-exact), and the level selection / blend against images the reference-exact bilinear path renders from the
individual levels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _np_mip_chain(level0: np.ndarray) -> list:
    """Definition of mlv_texture_generate_mips: per channel (a+b+c+d+2)>>2 over 2x2, odd extents repeat the last row/column."""
    chain = [level0]
    while chain[-1].shape != (1, 1):
        s = chain[-1]
        sh, sw = s.shape
        dh, dw = max(1, sh >> 1), max(1, sw >> 1)
        ys0, ys1 = np.minimum(2 * np.arange(dh), sh - 1), np.minimum(2 * np.arange(dh) + 1, sh - 1)
        xs0, xs1 = np.minimum(2 * np.arange(dw), sw - 1), np.minimum(2 * np.arange(dw) + 1, sw - 1)
        out = np.zeros((dh, dw), np.uint32)
        for shift in (0, 8, 16, 24):
            ch = (s >> np.uint32(shift)) & np.uint32(0xff)
            acc = ch[np.ix_(ys0, xs0)] + ch[np.ix_(ys0, xs1)] + ch[np.ix_(ys1, xs0)] + ch[np.ix_(ys1, xs1)] + np.uint32(2)
            out |= (acc >> np.uint32(2)) << np.uint32(shift)
        chain.append(out)
    return chain


def _noise_texture(w, h, seed):
    rng = np.random.default_rng(seed)
    # smooth-ish content so that neighbouring mip levels differ visibly but not wildly
    y, x = np.mgrid[0:h, 0:w]
    r = (127 + 120 * np.sin(x * 0.21 + seed)).astype(np.uint32)
    g = (127 + 120 * np.cos(y * 0.13)).astype(np.uint32)
    b = rng.integers(0, 256, (h, w), dtype=np.uint32)
    return (r | (g << 8) | (b << 16) | (np.uint32(255) << 24)).astype(np.uint32)


def _quad_scene(size, tex, ps):
    """A quad over the whole render target in clip space (identity clip_from_world, w = 1): UV is an affine function of
    the pixel, so the level of detail is the same everywhere: log2(texture extent / size)."""
    from malevich_b200 import basic_vs, scenes
    vb = np.zeros((4, 8), np.float32)
    for i, (x, y) in enumerate(((-1, -1), (1, -1), (-1, 1), (1, 1))):
        vb[i] = (x, y, 0.5, 0, 0, 1, (x + 1) / 2, (y + 1) / 2)
    ib = np.array([0, 1, 2, 2, 1, 3, 0, 2, 1, 2, 3, 1] + [0] * 12, np.uint32)  # both windings (one is culled) + padding to a multiple of 8
    cb = np.stack([np.eye(4, dtype=np.float32)] * 3)
    return scenes.Scene("quad", size, size, [scenes.SceneObject(vb, ib, basic_vs, ps, tex, "quad")], cb)


def _render(scene):
    from malevich_b200 import Device, scenes
    with Device(scene.width, scene.height) as dev:
        scenes.render(dev, scene)
        col, _ = dev.present()
    return col


def _rgb(img):
    return np.stack([(img >> 16) & 0xff, (img >> 8) & 0xff, img & 0xff], -1).astype(np.int32)


@pytest.mark.parametrize("w,h", [(256, 256), (37, 21), (64, 1), (1, 1), (100, 60)])
def test_mip_chain_matches_its_definition_bit_for_bit(w, h):
    from malevich_b200 import Device, Texture2D
    tex = Texture2D(_noise_texture(w, h, w * 31 + h), generate_mips=True)
    with Device(64, 64) as dev:
        got = dev.read_texture_mips(tex)
    want = _np_mip_chain(tex.p_data)
    assert len(got) == len(want)
    for level, (g, r) in enumerate(zip(got, want)):
        assert g.shape == r.shape and np.array_equal(g, r), f"level {level}"


def test_trilinear_without_minification_or_without_a_chain_is_basic_ps():
    from malevich_b200 import Texture2D, basic_ps, basic_trilinear_ps, scenes
    import cases
    t64 = _noise_texture(64, 64, 5)
    a = _render(_quad_scene(128, Texture2D(t64), basic_ps))
    assert len(np.unique(a)) > 1000  # the quad really is textured
    assert np.array_equal(a, _render(_quad_scene(128, Texture2D(t64, generate_mips=True), basic_trilinear_ps)))  # magnified: level 0, no blend
    sc = cases.SMALL["toon_320x200"]()
    ref = _render(sc)
    for o in sc.objects:
        o.pixel_shader = basic_trilinear_ps  # textures without a chain: one level
    assert np.array_equal(ref, _render(sc))


@pytest.mark.parametrize("size,lod", [(64, 2.0), (32, 3.0), (96, np.log2(256 / 96)), (48, np.log2(256 / 48)), (80, np.log2(256 / 80))])
def test_trilinear_selects_and_blends_the_levels(size, lod):
    """256^2 texture on a size^2 quad: lod = log2(256/size). Expected image = the blend (in linear light) of the images the
    bilinear path renders from levels floor(lod) and floor(lod)+1 as stand-alone textures."""
    from malevich_b200 import Texture2D, basic_ps, basic_trilinear_ps
    t = _noise_texture(256, 256, 11)
    chain = _np_mip_chain(t)
    got = _rgb(_render(_quad_scene(size, Texture2D(t, generate_mips=True), basic_trilinear_ps)))
    l0 = int(np.floor(lod + 1e-4))
    f = float(lod - l0) if lod - l0 > 1e-4 else 0.0
    img0 = _rgb(_render(_quad_scene(size, Texture2D(chain[l0]), basic_ps)))
    if f == 0.0:
        want, tol = img0, 1
    else:
        img1 = _rgb(_render(_quad_scene(size, Texture2D(chain[l0 + 1]), basic_ps)))
        lin = lambda e: ((e / 255.0 + 0.055) / 1.055) ** 2.4  # inverse of srgb_from_linear_approx (math.h:419-422)
        mix = lin(img0) * (1 - f) + lin(img1) * f
        want, tol = np.rint(np.maximum(1.055 * mix ** (1 / 2.4) - 0.055, 0) * 255).astype(np.int32), 2  # two quantised inputs
    diff = np.abs(got - want).max(-1)
    assert (got != _rgb(_render(_quad_scene(size, Texture2D(t), basic_ps)))).any(), "trilinear result does not differ from level 0 at all"
    assert diff.max() <= tol + 1 and (diff <= tol).mean() >= 0.995, (diff.max(), (diff <= tol).mean())
