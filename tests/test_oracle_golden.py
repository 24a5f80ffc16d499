"""CPU: pin the oracle against the reference's own torch oracle (tests/torch_impl.cpp) through the
committed golden vectors (tests/golden/torch_impl_*.npz, made by make_torch_impl_golden.py).
Mirrors the reference's TileIntersectionTest (tests/test_garden_data.cpp:531-569, exact equality)
and its SH forward/backward checks (tests/test_numerical_gradients.cpp:158-229, 1e-4)."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["c3", "c1"])
def test_intersect_exact_vs_torch_impl(orc, name):
    g = np.load(os.path.join(G, f"torch_impl_isect_{name}.npz"))
    C = g["means2d"].shape[0]
    tw, th, ts = int(g["tile_width"]), int(g["tile_height"]), int(g["tile_size"])
    tpg, ids, flat = orc.isect_tiles(g["means2d"], g["radii"], g["depths"], C, ts, tw, th, True)
    assert np.array_equal(tpg, g["tiles_per_gauss"])
    assert np.array_equal(flat, g["flatten_ids"])
    assert np.array_equal(ids, g["isect_ids"])
    # offsets: definition check (first sorted position with (cam, tile) >= id; n_isects after the last)
    off = orc.isect_offsets(ids, C, tw, th).reshape(-1)
    bits = 0
    n = tw * th
    while n:
        bits += 1
        n >>= 1
    key = ids >> 32
    flat_tile = (key >> bits) * (tw * th) + (key & ((1 << bits) - 1))
    expect = np.searchsorted(flat_tile, np.arange(C * tw * th), side="left")
    assert np.array_equal(off, expect.astype(np.int32))


@pytest.mark.parametrize("deg", [0, 1, 2, 3, 4])
def test_sh_vs_torch_impl(orc, deg):
    g = np.load(os.path.join(G, "torch_impl_sh.npz"))
    col = orc.sh_fwd(deg, g["dirs"], g["coeffs"])
    np.testing.assert_allclose(col, g[f"colors_deg{deg}"], rtol=1e-4, atol=1e-4)
    v_coeffs, v_dirs = orc.sh_bwd(deg, g["dirs"], g["coeffs"], None, g["v_colors"])
    np.testing.assert_allclose(v_coeffs, g[f"v_coeffs_deg{deg}"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(v_dirs, g[f"v_dirs_deg{deg}"], rtol=1e-4, atol=1e-4)


def test_sh_masks_and_extra_coeffs(orc):
    g = np.load(os.path.join(G, "torch_impl_sh.npz"))
    masks = (np.arange(g["dirs"].shape[0]) % 3) != 0
    col = orc.sh_fwd(2, g["dirs"], g["coeffs"], masks)
    assert np.all(col[~masks] == 0)  # untouched rows (zero-initialised by the wrapper)
    np.testing.assert_allclose(col[masks], g["colors_deg2"][masks], rtol=1e-4, atol=1e-4)
    v_coeffs, _ = orc.sh_bwd(2, g["dirs"], g["coeffs"], masks, g["v_colors"])
    assert np.all(v_coeffs[~masks] == 0) and np.all(v_coeffs[:, 9:] == 0)


def test_quat_to_rotmat_vs_torch_impl(orc):
    g = np.load(os.path.join(G, "torch_impl_quat.npz"))
    np.testing.assert_allclose(orc.quat_to_rotmat(g["quats"]), g["rotmats"], rtol=1e-5, atol=1e-5)
