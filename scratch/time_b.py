"""quick per-op timing of config B on the GPU (scratch; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import __graft_entry__ as ge
import scenes
pkg = ge.load_package(); pkg.load()
N = int(os.environ.get('N', 1000000))
sc = scenes.scene_b(N=N)
dev = torch.device('cuda:0')
t = {k: torch.from_numpy(v).to(dev) for k,v in sc.items() if isinstance(v,np.ndarray)}
W,H = sc['width'], sc['height']; tw,th=(W+15)//16,(H+15)//16
def ev(): 
    e=torch.cuda.Event(enable_timing=True); e.record(); return e
def run(timing=True):
    T={}
    e0=ev()
    radii, means2d, depths, conics, _ = pkg.projection_ut_3dgs_fused(t['means'],t['quats'],t['scales'],t['opacities'],t['viewmats'],t['Ks'],W,H,0.3,0.01,1e4,0.0)
    e1=ev()
    campos = torch.linalg.inv(t['viewmats'])[:, :3, 3]
    dirs = t['means'][None]-campos[:,None]; masks=(radii>0).all(-1)
    colors = pkg.spherical_harmonics_fwd(3, dirs.reshape(-1,3), t['sh_coeffs'], masks.reshape(-1))
    colors_act = torch.clamp_min(colors+0.5,0).reshape(1,-1,3)
    e2=ev()
    tpg, ids, flat = pkg.intersect_tile(means2d, radii, depths, 1, 16, tw, th, True)
    off = pkg.intersect_offset(ids,1,tw,th)
    e3=ev()
    r,a,li = pkg.rasterize_to_pixels_from_world_3dgs_fwd(t['means'],t['quats'],t['scales'],colors_act,t['opacities'][None],t['background'],None,W,H,16,t['viewmats'],t['Ks'],off,flat)
    e4=ev()
    vr=torch.ones_like(r); va=torch.ones_like(a)
    g = pkg.rasterize_to_pixels_from_world_3dgs_bwd(t['means'],t['quats'],t['scales'],colors_act,t['opacities'][None],t['background'],None,W,H,16,t['viewmats'],t['Ks'],off,flat,a,li,vr,va)
    e5=ev()
    vc = g[3].reshape(-1,3)*((colors+0.5)>0)
    v_coeffs, v_dirs = pkg.spherical_harmonics_bwd(16,3,dirs.reshape(-1,3),t['sh_coeffs'],masks.reshape(-1),vc.contiguous(),True)
    e6=ev()
    torch.cuda.synchronize()
    return dict(proj=e0.elapsed_time(e1), sh_fwd=e1.elapsed_time(e2), isect=e2.elapsed_time(e3), blend_fwd=e3.elapsed_time(e4), blend_bwd=e4.elapsed_time(e5), sh_bwd=e5.elapsed_time(e6), total=e0.elapsed_time(e6)), dict(n_isects=int(flat.shape[0]), visible=int(masks.sum()), alpha_mean=float(a.mean()), maxtile=int((torch.diff(off.reshape(-1).long(), append=torch.tensor([flat.shape[0]],device=dev))).max()))
for i in range(3): run()
res=[run() for _ in range(10)]
info=res[0][1]
print('N',N,info)
for k in res[0][0]:
    v=[r[0][k] for r in res]; print(f'{k:10s} median {np.median(v):8.3f} ms  min {np.min(v):8.3f}')
