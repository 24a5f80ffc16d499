"""CPU: self-consistency of the oracle on the rows the reference's tests do not pin
(UT projection, from-world blend fwd/bwd): float32 vs float64 builds, finite differences of the
forward against the analytic backward, and structural properties of the blend."""
import numpy as np
import pytest

import scenes


@pytest.fixture(scope="module")
def small(orc):
    sc = scenes.scene_small(N=600, width=96, height=64, sh_degree=2, view=3)
    rng = np.random.default_rng(1)
    vrc = rng.standard_normal((1, sc["height"], sc["width"], 3)).astype(np.float32)
    vra = rng.standard_normal((1, sc["height"], sc["width"], 1)).astype(np.float32)
    o32 = orc.render_pipeline(sc, "f32", True, vrc, vra)
    o64 = orc.render_pipeline(sc, "f64", True, vrc, vra)
    return sc, vrc, vra, o32, o64


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_f32_vs_f64(small):
    sc, _, _, o32, o64 = small
    assert (o32["radii"] != o64["radii"]).sum() <= 2  # ceil() borderline flips only
    if np.array_equal(o32["flatten_ids"], o64["flatten_ids"]):
        assert rel(o32["renders"], o64["renders"]) < 1e-5
        for k in ("v_means", "v_quats", "v_scales", "v_colors", "v_opacities"):
            assert rel(o32[k], o64[k]) < 1e-3, k


def test_blend_properties(small):
    sc, _, _, o32, _ = small
    a = o32["alphas"]
    assert np.all(a >= 0) and np.all(a <= 1.0)
    # transmittance never drops below the 1e-4 stop threshold by more than one max-alpha step
    assert np.all(1 - a > 1e-4 * (1 - 0.999) - 1e-9)
    off = o32["tile_offsets"].reshape(-1)
    li = o32["last_ids"][0]
    ts = 16
    tw = (sc["width"] + ts - 1) // ts
    n = len(o32["flatten_ids"])
    for ty in range((sc["height"] + ts - 1) // ts):
        for tx in range(tw):
            t = ty * tw + tx
            lo, hi = off[t], (off[t + 1] if t + 1 < len(off) else n)
            blk = li[ty * ts:(ty + 1) * ts, tx * ts:(tx + 1) * ts]
            touched = blk[blk != 0]
            assert np.all((touched >= lo) & (touched < hi))


def test_backward_matches_finite_differences(orc, small):
    """d(sum(render * vrc) + sum(alpha * vra)) / d(param) by central differences in float64."""
    sc, vrc, vra, _, o64 = small
    base = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in sc.items()}

    def loss(scn):
        # keep the isect lists fixed (projection / binning is not differentiated, Ops.h:67)
        r, a, _ = orc.raster_fwd(scn["means"], scn["quats"], scn["scales"], o64["colors"], scn["opacities"][None],
                                 scn["background"], None, scn["width"], scn["height"], 16, scn["viewmats"], scn["Ks"],
                                 o64["tile_offsets"], o64["flatten_ids"], precision="f64")
        return float((r.astype(np.float64) * vrc).sum() + (a.astype(np.float64) * vra).sum())

    vis = np.nonzero(np.abs(o64["v_means"]).sum(-1) > 1e-3)[0][:6]
    checked = 0
    for g in vis:
        for name, key, dim in (("means", "v_means", 0), ("scales", "v_scales", 1), ("opacities", "v_opacities", None),
                               ("quats", "v_quats", 2)):
            ana = o64[key][0, g] if dim is None else o64[key][g, dim]
            if abs(ana) <= 0.05:
                continue
            ok = False
            nums = []
            # the forward is only piecewise smooth (1/255 cut, T <= 1e-4 stop): accept if any step size agrees
            for eps in (1e-4, 5e-5, 2e-4):
                p = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in base.items()}
                m = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in base.items()}
                if dim is None:
                    p[name][g] += eps; m[name][g] -= eps
                    h = float(p[name][g]) - float(m[name][g])
                else:
                    p[name][g, dim] += eps; m[name][g, dim] -= eps
                    h = float(p[name][g, dim]) - float(m[name][g, dim])
                num = (loss(p) - loss(m)) / h
                nums.append(num)
                if abs(num - ana) <= 0.03 * abs(ana) + 0.02:
                    ok = True
                    break
            assert ok, (name, g, nums, ana)
            checked += 1
    assert checked >= 6
