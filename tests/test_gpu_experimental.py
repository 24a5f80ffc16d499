"""B200: reduction variants of the blend backward that have NOT been validated on hardware yet (written
without a GPU at hand at the end of round 1).  Skipped unless GSB_TEST_EXPERIMENTAL=1, so that the default
GPU suite only contains paths that have run green on a B200; first thing to run in the next round:

    GSB_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q
"""
import os

import numpy as np
import pytest
import torch

import scenes
from test_gpu_parity import rel, to_dev

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("GSB_TEST_EXPERIMENTAL") != "1",
                                 reason="experimental kernel variants: set GSB_TEST_EXPERIMENTAL=1")]


@pytest.mark.parametrize("mode", ["transpose"])
def test_blend_bwd_reduction_variant_matches_shuffle_reduction(native, orc, cuda_device, monkeypatch, mode):
    sc = scenes.scene_b(N=30000, width=640, height=360, view=3, scale_mul=2.0)
    W, H = sc["width"], sc["height"]
    rng = np.random.default_rng(11)
    vrc = torch.from_numpy(rng.standard_normal((1, H, W, 3)).astype(np.float32)).to(cuda_device)
    vra = torch.from_numpy(rng.standard_normal((1, H, W, 1)).astype(np.float32)).to(cuda_device)
    ref = orc.render_pipeline(sc, "f32", False)
    t = to_dev(sc, cuda_device)
    colors = torch.from_numpy(ref["colors"]).to(cuda_device)
    off = torch.from_numpy(ref["tile_offsets"]).to(cuda_device)
    flat = torch.from_numpy(ref["flatten_ids"]).to(cuda_device)
    args = (t["means"], t["quats"], t["scales"], colors, t["opacities"][None], t.get("background"), None, W, H, 16,
            t["viewmats"], t["Ks"], off, flat)
    _, alphas, last_ids = native.rasterize_to_pixels_from_world_3dgs_fwd(*args)
    out = {}
    for m in ("shuffle", mode):
        monkeypatch.setenv("GSB_BWD_REDUCE", m)
        out[m] = native.rasterize_to_pixels_from_world_3dgs_bwd(*args, alphas, last_ids, vrc, vra)
        torch.cuda.synchronize()
    for a, b, name in zip(out["shuffle"], out[mode], ("v_means", "v_quats", "v_scales", "v_colors", "v_opacities")):
        assert rel(b.cpu().numpy(), a.cpu().numpy()) < 2e-5, name
