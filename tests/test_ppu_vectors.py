"""Hand-derived PPU scenes (tests/ppu_vectors.py) through the oracle (CPU tier) and through the CUDA renderer (GPU tier)."""
import numpy as np
import pytest

from helpers import SHADE_WORDS
from ppu_vectors import SCENES
from pokegym_b200 import _capi
from pokegym_b200.state_file import field_offsets, parse_state, serialize_state


def _render(handle, scene):
    vram, oam, regs, params, expected = scene.arrays()
    st = parse_state(handle.save_state(0))
    st.raw["vram"], st.raw["oam"], st.raw["lcd_regs"], st.raw["scanline_params"] = vram, oam, regs, params.reshape(-1)
    handle.load_template(handle.add_state_template(serialize_state(st)))
    handle.debug_render_frame(0)
    off, ln = field_offsets(9)["screen"]
    got = np.frombuffer(handle.save_state(0)[off:off + ln], dtype="<u4").reshape(144, 160)
    want = SHADE_WORDS[expected]
    if not np.array_equal(got, want):
        ys, xs = np.nonzero(got != want)
        lut = {int(w): i for i, w in enumerate(SHADE_WORDS)}
        raise AssertionError(f"{scene.name}: {len(ys)} pixels differ; first at (x={xs[0]}, y={ys[0]}): got shade {lut.get(int(got[ys[0], xs[0]]))}, "
                             f"expected {expected[ys[0], xs[0]]}")


@pytest.mark.parametrize("make", SCENES, ids=lambda f: f.__name__)
def test_oracle_renders_hand_derived_scene(oracle_lib, roms, make):
    _render(_capi.Handle(oracle_lib, 1, roms("pokelike")), make())


@pytest.mark.gpu
@pytest.mark.parametrize("make", SCENES, ids=lambda f: f.__name__)
def test_cuda_renders_hand_derived_scene(cuda_lib, roms, make):
    _render(_capi.Handle(cuda_lib, 1, roms("pokelike"), 0), make())
