"""small fwd+bwd through every kernel, for compute-sanitizer (memcheck / racecheck / synccheck)"""
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import __graft_entry__ as ge, scenes
pkg=ge.load_package(); pkg.load()
dev=torch.device('cuda:0')
for sc in (scenes.scene_small(N=2500, width=150, height=70, view=3), scenes.scene_a(N=3000, width=96, height=80)):
    t={k:torch.from_numpy(v).to(dev) for k,v in sc.items() if isinstance(v,np.ndarray)}
    P={k:t[k].clone().requires_grad_(True) for k in ('means','quats','scales','opacities','sh_coeffs')}
    out=pkg.rasterize(P['means'],P['quats'],P['scales'],P['opacities'],P['sh_coeffs'],sc['sh_degree'],t['viewmats'],t['Ks'],sc['width'],sc['height'],bg_color=t.get('background'))
    (out.render_colors.sum()+out.alpha.sum()).backward()
    pkg.quats_to_rotmats(t['quats'])
torch.cuda.synchronize(); print('sanitize workload ok', out.n_isects)
