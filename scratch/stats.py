"""run config B once with the instrumented library (scratch/statlib) and print pair statistics"""
import sys, os, ctypes, shutil
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
# swap in the instrumented C-ABI lib
L='gaussian-splatting-cuda_b200/lib/libgsb200.so'
shutil.copy(L, '/tmp/libgsb200_orig.so'); shutil.copy('scratch/statlib/libgsb200.so', L)
try:
    import __graft_entry__ as ge, scenes
    pkg=ge.load_package(); pkg.load()
    cabi=ctypes.CDLL(pkg.CABI_PATH)
    sc=scenes.scene_b(N=int(os.environ.get('N',1000000)))
    dev=torch.device('cuda:0')
    t={k:torch.from_numpy(v).to(dev) for k,v in sc.items() if isinstance(v,np.ndarray)}
    P={k:t[k].clone().requires_grad_(True) for k in ('means','quats','scales','opacities','sh_coeffs')}
    out=pkg.rasterize(P['means'],P['quats'],P['scales'],P['opacities'],P['sh_coeffs'],3,t['viewmats'],t['Ks'],sc['width'],sc['height'],bg_color=t['background'])
    out.render_colors.abs().mean().backward(); torch.cuda.synchronize()
    buf=(ctypes.c_ulonglong*16)(); cabi.gsb_debug_stats(buf); s=list(buf)
    I=out.n_isects
    print('isects',I,'warp-pairs total',4*I)
    print('fwd: scanned %d (%.1f%% of 4I) candidates %d (%.1f%% of scanned) any-pass events %d (%.1f%% of cand) pass-lanes/event %.1f'%(s[0],100*s[0]/(4*I),s[1],100*s[1]/max(s[0],1),s[2],100*s[2]/max(s[1],1),s[3]/max(s[2],1)))
    print('bwd: scanned %d (%.1f%% of 4I) candidates %d (%.1f%% of scanned) events %d (%.1f%% of cand) pass-lanes/event %.1f'%(s[4],100*s[4]/(4*I),s[5],100*s[5]/max(s[4],1),s[6],100*s[6]/max(s[5],1),s[7]/max(s[6],1)))
finally:
    shutil.copy('/tmp/libgsb200_orig.so', L)
