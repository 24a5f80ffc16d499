"""numpy prototype of the reformulated blend math (quadratic forms) vs the oracle."""
import sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from oracle import oracle as orc
import scenes

def quat_to_R(q):  # math (row-major) rotation, normalised
    q = q/np.linalg.norm(q,axis=-1,keepdims=True)
    w,x,y,z = q[...,0],q[...,1],q[...,2],q[...,3]
    R = np.stack([1-2*(y*y+z*z), 2*(x*y-w*z), 2*(x*z+w*y),
                  2*(x*y+w*z), 1-2*(x*x+z*z), 2*(y*z-w*x),
                  2*(x*z-w*y), 2*(y*z+w*x), 1-2*(x*x+y*y)],-1).reshape(q.shape[:-1]+(3,3))
    return R

def prep(means, quats, scales, viewmat, dt):
    means=means.astype(dt); quats=quats.astype(dt); scales=scales.astype(dt)
    Rc = viewmat[:3,:3].astype(dt); t = viewmat[:3,3].astype(dt)
    Rg = quat_to_R(quats)                          # [N,3,3]
    M = np.swapaxes(Rg,-1,-2)/scales[...,None]     # S^-1 R^T
    A = M @ Rc.T                                   # [N,3,3]
    mu_c = means @ Rc.T + t
    zc = mu_c[:,2]; uc = mu_c[:,0]/zc; vc = mu_c[:,1]/zc
    A0,A1,A2 = A[:,:,0],A[:,:,1],A[:,:,2]
    G = A0*uc[:,None] + A1*vc[:,None] + A2
    gro = -zc[:,None]*G
    E0 = np.cross(A0,gro); E1 = np.cross(A1,gro)
    n = np.stack([(E0*E0).sum(-1), 2*(E0*E1).sum(-1), (E1*E1).sum(-1)],-1)
    d = np.stack([(G*G).sum(-1), 2*(G*A0).sum(-1), 2*(G*A1).sum(-1), (A0*A0).sum(-1), 2*(A0*A1).sum(-1), (A1*A1).sum(-1)],-1)
    return uc,vc,n,d

def power_mine(uc,vc,n,d,u,v):
    du=u-uc; dv=v-vc
    N = n[:,0]*du*du + n[:,1]*du*dv + n[:,2]*dv*dv
    D = d[:,0] + d[:,1]*du + d[:,2]*dv + d[:,3]*du*du + d[:,4]*du*dv + d[:,5]*dv*dv
    return -0.5*N/D

def power_ref(means,quats,scales,viewmat,u,v,dt):
    # reference formulation in dtype dt
    means=means.astype(dt); quats=quats.astype(dt); scales=scales.astype(dt)
    Rc = viewmat[:3,:3].astype(dt); t = viewmat[:3,3].astype(dt)
    Rg = quat_to_R(quats); M = np.swapaxes(Rg,-1,-2)/scales[...,None]
    o = -Rc.T @ t
    c = np.stack([u,v,np.ones_like(u)],-1); c = c/np.linalg.norm(c,axis=-1,keepdims=True)
    dvec = c @ Rc   # R^-1 c = Rc^T c
    gro = np.einsum('nij,nj->ni', M, o-means)
    grd = np.einsum('nij,nj->ni', M, dvec); grd = grd/np.linalg.norm(grd,axis=-1,keepdims=True)
    k = np.cross(grd,gro)
    return -0.5*(k*k).sum(-1)

if __name__=='__main__':
    sc = scenes.scene_b(N=20000, seed=1, view=1)
    out = orc.render_pipeline(sc, precision='f64')
    vis = out['masks'][0]
    idx = np.nonzero(vis)[0]
    rng = np.random.default_rng(0)
    fx=sc['Ks'][0,0,0]; fy=sc['Ks'][0,1,1]; cx=sc['Ks'][0,0,2]; cy=sc['Ks'][0,1,2]
    m2d = out['means2d'][0][idx]; rad = out['radii'][0][idx]
    # random pixel inside each gaussian's bbox
    px = np.floor(m2d[:,0] + (rng.random(len(idx))*2-1)*rad[:,0]*0.7)+0.5
    py = np.floor(m2d[:,1] + (rng.random(len(idx))*2-1)*rad[:,1]*0.7)+0.5
    V = sc['viewmats'][0]
    for dt in (np.float64, np.float32):
        u = ((px-cx)/fx).astype(dt); v=((py-cy)/fy).astype(dt)
        uc,vc,n,d = prep(sc['means'][idx],sc['quats'][idx],sc['scales'][idx],V,dt)
        pm = power_mine(uc,vc,n,d,u,v)
        pr = power_ref(sc['means'][idx],sc['quats'][idx],sc['scales'][idx],V,u,v,dt)
        if dt==np.float64: truth = pr
        sel = truth > -8
        print(dt.__name__, 'n=',sel.sum(), 'mine-vs-truth max abs', np.abs(pm-truth)[sel].max(), 'ref-form-vs-truth max abs', np.abs(pr-truth)[sel].max(),
              'p99 mine', np.percentile(np.abs(pm-truth)[sel],99), 'p99 ref', np.percentile(np.abs(pr-truth)[sel],99))
