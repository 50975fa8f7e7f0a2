"""malevich_b200 -- B200-native (sm_100a) implementation of Malevich's draw pipeline.

Only what the hot path needs: `csrc/` (CUDA kernels + the C-ABI of include/malevich_b200.h),
`device` (host-side mirror of the reference's pipeline-state interface), `scenes` (the caller,
`render()`), `assets`/`camera` (host-side inputs). There is no CPU fallback: creating a `Device`
fails loudly when the CUDA extension is missing or no B200 is visible.
"""
from . import _lib  # noqa: F401
from ._lib import MalevichError  # noqa: F401
from .device import (CommandList, Device, PixelShader, Texture2D, VertexShader, basic_ps, basic_trilinear_ps, basic_vs, env_lighting_ps,  # noqa: F401
                     fullscreen_vs, passthrough_ps, passthrough_vs, vertex_lighting_vs)
