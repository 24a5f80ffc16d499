"""prototype: quadratic-form blend fwd + moment-based bwd vs oracle (f64), then float32 emulation."""
import sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests'); sys.path.insert(0,'scratch')
import numpy as np
from oracle import oracle as orc
import scenes
from proto import quat_to_R

def quat_cast32(R):  # GLM quat_cast on float32 math rotation R (row-major), returns (w,x,y,z) float32
    R=R.astype(np.float32)
    m=lambda c,r: R[r,c]
    fx=m(0,0)-m(1,1)-m(2,2); fy=m(1,1)-m(0,0)-m(2,2); fz=m(2,2)-m(0,0)-m(1,1); fw=m(0,0)+m(1,1)+m(2,2)
    bi=0; fb=fw
    if fx>fb: fb=fx; bi=1
    if fy>fb: fb=fy; bi=2
    if fz>fb: fb=fz; bi=3
    bv=np.float32(np.sqrt(np.float32(fb+np.float32(1)))*np.float32(0.5)); mult=np.float32(0.25)/bv
    if bi==0: q=(bv,(m(1,2)-m(2,1))*mult,(m(2,0)-m(0,2))*mult,(m(0,1)-m(1,0))*mult)
    elif bi==1: q=((m(1,2)-m(2,1))*mult,bv,(m(0,1)+m(1,0))*mult,(m(2,0)+m(0,2))*mult)
    elif bi==2: q=((m(2,0)-m(0,2))*mult,(m(0,1)+m(1,0))*mult,bv,(m(1,2)+m(2,1))*mult)
    else: q=((m(0,1)-m(1,0))*mult,(m(2,0)+m(0,2))*mult,(m(1,2)+m(2,1))*mult,bv)
    return np.array(q,np.float32)

def mat3_cast(q, dt):  # GLM mat3_cast (no normalisation), returns math matrix (row-major)
    w,x,y,z=[dt(v) for v in q]
    return np.array([[1-2*(y*y+z*z), 2*(x*y-w*z), 2*(x*z+w*y)],
                     [2*(x*y+w*z), 1-2*(x*x+z*z), 2*(y*z-w*x)],
                     [2*(x*z-w*y), 2*(y*z+w*x), 1-2*(x*x+y*y)]],dt)

def cam_B(viewmat):
    q=quat_cast32(viewmat[:3,:3])
    d=np.float32((q*q).sum())
    qi=np.array([q[0]/d,-q[1]/d,-q[2]/d,-q[3]/d],np.float32)
    B=mat3_cast(qi,np.float32)      # reference R_inv in float32
    return B.astype(np.float64), viewmat[:3,3].astype(np.float64)

def prep(means, quats, scales, B, t):
    dt=np.float64
    means=means.astype(dt); quats=quats.astype(dt); scales=scales.astype(dt)
    Binv=np.linalg.inv(B)
    Rg = quat_to_R(quats)
    M = np.swapaxes(Rg,-1,-2)/scales[...,None]
    A = M @ B
    mu_c = means @ Binv.T + t
    zc = mu_c[:,2]; uc = mu_c[:,0]/zc; vc = mu_c[:,1]/zc
    A0,A1,A2 = A[:,:,0],A[:,:,1],A[:,:,2]
    G = A0*uc[:,None] + A1*vc[:,None] + A2
    gro = -zc[:,None]*G
    E0 = np.cross(A0,gro); E1 = np.cross(A1,gro)
    n = np.stack([(E0*E0).sum(-1), 2*(E0*E1).sum(-1), (E1*E1).sum(-1)],-1)
    d = np.stack([(G*G).sum(-1), 2*(G*A0).sum(-1), 2*(G*A1).sum(-1), (A0*A0).sum(-1), 2*(A0*A1).sum(-1), (A1*A1).sum(-1)],-1)
    return dict(uc=uc,vc=vc,n=n,d=d,A=A,G=G,gro=gro,E0=E0,E1=E1,zc=zc,M=M,Rg=Rg,Binv=Binv,mu_c=mu_c)

def finalize(P, vn, vd, vuc, vvc, quats, scales, B):
    """chain rule from coefficient grads to (mean, quat, scale) grads. all float64."""
    A=P['A']; A0,A1,A2=A[:,:,0],A[:,:,1],A[:,:,2]; G=P['G']; gro=P['gro']; E0=P['E0']; E1=P['E1']; zc=P['zc']; uc=P['uc']; vc=P['vc']
    c=lambda a: a[:,None]
    vE0 = 2*c(vn[:,0])*E0 + c(vn[:,1])*2*E1
    vE1 = 2*c(vn[:,2])*E1 + c(vn[:,1])*2*E0
    vG = 2*c(vd[:,0])*G + 2*c(vd[:,1])*A0 + 2*c(vd[:,2])*A1
    vA0 = 2*c(vd[:,1])*G + 2*c(vd[:,3])*A0 + 2*c(vd[:,4])*A1 + np.cross(gro, vE0)
    vA1 = 2*c(vd[:,2])*G + 2*c(vd[:,5])*A1 + 2*c(vd[:,4])*A0 + np.cross(gro, vE1)
    vgro = np.cross(vE0, A0) + np.cross(vE1, A1)
    vG = vG - c(zc)*vgro
    vzc = -(G*vgro).sum(-1)
    vA0 = vA0 + c(uc)*vG; vA1 = vA1 + c(vc)*vG; vA2 = vG
    vuc = vuc + (A0*vG).sum(-1); vvc = vvc + (A1*vG).sum(-1)
    vxc = vuc/zc; vyc = vvc/zc; vzc = vzc - (uc*vuc + vc*vvc)/zc
    vmuc = np.stack([vxc,vyc,vzc],-1)
    vmean = vmuc @ P['Binv']          # mu_c = Binv mu + t -> v_mu = Binv^T v_muc
    vA = np.stack([vA0,vA1,vA2],-1)   # [N,3(row),3(col)]
    vM = vA @ B.T                      # A = M B -> vM = vA B^T
    s=scales.astype(np.float64); Rg=P['Rg']
    # M = diag(1/s) Rg^T : M_ij = Rg_ji / s_i
    vs = -(vM*np.swapaxes(Rg,-1,-2)).sum(-1)/(s*s)
    vRgT = vM/s[...,None]; vRg=np.swapaxes(vRgT,-1,-2)
    # quaternion vjp (with normalisation)
    q=quats.astype(np.float64); inv=1/np.linalg.norm(q,axis=-1); qn=q*inv[:,None]
    w,x,y,z=qn[:,0],qn[:,1],qn[:,2],qn[:,3]
    g=lambda i,j: vRg[:,i,j]
    vq=np.stack([
      2*(x*(g(2,1)-g(1,2)) + y*(g(0,2)-g(2,0)) + z*(g(1,0)-g(0,1))),
      2*(-2*x*(g(1,1)+g(2,2)) + y*(g(1,0)+g(0,1)) + z*(g(2,0)+g(0,2)) + w*(g(2,1)-g(1,2))),
      2*(x*(g(1,0)+g(0,1)) - 2*y*(g(0,0)+g(2,2)) + z*(g(2,1)+g(1,2)) + w*(g(0,2)-g(2,0))),
      2*(x*(g(2,0)+g(0,2)) + y*(g(2,1)+g(1,2)) - 2*z*(g(0,0)+g(1,1)) + w*(g(1,0)-g(0,1)))],-1)
    vq=(vq-(vq*qn).sum(-1,keepdims=True)*qn)*inv[:,None]
    return vmean, vq, vs

def blend(sc, out, vrc, vra, dt):
    """my formulation: fwd + bwd, per pixel python loops (small scenes only). dt = accumulation dtype for per-pair math"""
    W,H,ts=sc['width'],sc['height'],16
    tw,th=(W+ts-1)//ts,(H+ts-1)//ts
    B,t=cam_B(sc['viewmats'][0])
    N=sc['means'].shape[0]
    P=prep(sc['means'],sc['quats'],sc['scales'],B,t)
    fx,fy,cx,cy=[np.float64(v) for v in (sc['Ks'][0,0,0],sc['Ks'][0,1,1],sc['Ks'][0,0,2],sc['Ks'][0,1,2])]
    uc=P['uc'].astype(dt); vc=P['vc'].astype(dt); n=P['n'].astype(dt); d=P['d'].astype(dt)
    op=sc['opacities'].astype(dt); col=out['colors'][0].astype(dt)
    bg=sc['background'][0].astype(dt) if sc['background'] is not None else None
    flat=out['flatten_ids']; off=out['tile_offsets'][0].reshape(-1); nI=len(flat)
    img=np.zeros((H,W,3),dt); alp=np.zeros((H,W),dt); last=np.zeros((H,W),np.int32)
    mom=np.zeros((N,15),np.float64)
    for tile in range(tw*th):
        ty,tx=divmod(tile,tw)
        rs=off[tile]; re=off[tile+1] if tile+1<tw*th else nI
        if re<=rs and bg is None: 
            pass
        g=flat[rs:re]
        ys=np.arange(ty*ts,min((ty+1)*ts,H)); xs=np.arange(tx*ts,min((tx+1)*ts,W))
        PX,PY=np.meshgrid(xs,ys)
        u=((PX+0.5-cx)/fx).astype(dt).reshape(-1); v=((PY+0.5-cy)/fy).astype(dt).reshape(-1)
        npx=len(u)
        du=u[:,None]-uc[g][None]; dv=v[:,None]-vc[g][None]
        Nq=n[g,0]*du*du+n[g,1]*du*dv+n[g,2]*dv*dv
        Dq=d[g,0]+d[g,1]*du+d[g,2]*dv+d[g,3]*du*du+d[g,4]*du*dv+d[g,5]*dv*dv
        power=(dt(-0.5)*Nq/Dq).astype(dt)
        visg=np.exp(power); araw=op[g][None]*visg
        alpha=np.minimum(dt(0.999),araw)
        ok=alpha>=dt(1/255.)
        T=np.ones(npx,dt); C=np.zeros((npx,3),dt); li=np.zeros(npx,np.int32); done=np.zeros(npx,bool)
        contrib=np.zeros((npx,len(g)),bool)
        for k in range(len(g)):
            a=alpha[:,k]; act=ok[:,k]&~done
            nT=T*(1-a)
            stop=act&(nT<=dt(1e-4)); done|=stop; act&=~stop
            C[act]+=col[g[k]][None]*(a[act]*T[act])[:,None]
            li[act]=rs+k; T[act]=nT[act]; contrib[:,k]=act
        Tf=T.copy()
        img[ys[0]:ys[-1]+1,xs[0]:xs[-1]+1]=(C+(Tf[:,None]*bg[None] if bg is not None else 0)).reshape(len(ys),len(xs),3)
        alp[ys[0]:ys[-1]+1,xs[0]:xs[-1]+1]=(1-Tf).reshape(len(ys),len(xs))
        last[ys[0]:ys[-1]+1,xs[0]:xs[-1]+1]=li.reshape(len(ys),len(xs))
        # backward
        if vrc is None: continue
        vc_=vrc[0][ys[0]:ys[-1]+1,xs[0]:xs[-1]+1].reshape(-1,3).astype(dt); va_=vra[0][ys[0]:ys[-1]+1,xs[0]:xs[-1]+1].reshape(-1).astype(dt)
        buf=np.zeros((npx,3),dt); T=Tf.copy()
        for k in range(len(g)-1,-1,-1):
            act=contrib[:,k]
            if not act.any(): continue
            a=alpha[:,k]; ra=1/(1-a)
            T=np.where(act,T*ra,T)
            fac=a*T
            valpha=((col[g[k]][None]*T[:,None]-buf*ra[:,None])*vc_).sum(-1)+Tf*ra*va_
            if bg is not None: valpha+=-Tf*ra*(bg[None]*vc_).sum(-1)
            gate=act&(araw[:,k]<=dt(0.999))
            gq=np.where(gate,araw[:,k]*valpha,0).astype(dt)      # = v_power
            w1=dt(-0.5)*gq/Dq[:,k]; w2=-w1*Nq[:,k]/Dq[:,k]
            a_du=du[:,k]; a_dv=dv[:,k]
            m=np.stack([w1*a_du,w1*a_dv,w1*a_du*a_du,w1*a_du*a_dv,w1*a_dv*a_dv,
                        w2,w2*a_du,w2*a_dv,w2*a_du*a_du,w2*a_du*a_dv,w2*a_dv*a_dv,
                        gq, *(np.where(act,fac,0)[:,None]*vc_).T],-1).astype(dt)
            mom[g[k]]+=m.sum(0,dtype=dt).astype(np.float64)
            buf=np.where(act[:,None],buf+col[g[k]][None]*fac[:,None],buf)
    res=dict(img=img,alpha=alp,last=last)
    if vrc is not None:
        n64=P['n']; d64=P['d']
        vn=mom[:,2:5]; vd=mom[:,5:11]
        vuc=-(2*n64[:,0]*mom[:,0]+n64[:,1]*mom[:,1] + d64[:,1]*mom[:,5]+2*d64[:,3]*mom[:,6]+d64[:,4]*mom[:,7])
        vvc=-(n64[:,1]*mom[:,0]+2*n64[:,2]*mom[:,1] + d64[:,2]*mom[:,5]+d64[:,4]*mom[:,6]+2*d64[:,5]*mom[:,7])
        vmean,vq,vs=finalize(P,vn,vd,vuc,vvc,sc['quats'],sc['scales'],B)
        res.update(v_means=vmean,v_quats=vq,v_scales=vs,v_opacities=mom[:,11]/sc['opacities'].astype(np.float64),v_colors=mom[:,12:15])
    return res

def cmp(name,a,b):
    a=np.asarray(a,np.float64); b=np.asarray(b,np.float64)
    print(f'  {name:12s} max|d| {np.abs(a-b).max():.3e}  max|b| {np.abs(b).max():.3e}  relL2 {np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-30):.3e}')

if __name__=='__main__':
    sc=scenes.scene_small(N=1500, view=1)
    rng=np.random.default_rng(0)
    o64=orc.render_pipeline(sc,precision='f64')
    vrc=rng.standard_normal(o64['renders'].shape).astype(np.float32); vra=rng.standard_normal(o64['alphas'].shape).astype(np.float32)
    o64=orc.render_pipeline(sc,precision='f64',with_bwd=True,v_render_colors=vrc,v_render_alphas=vra)
    o32=orc.render_pipeline(sc,precision='f32',with_bwd=True,v_render_colors=vrc,v_render_alphas=vra)
    print('isects',len(o64['flatten_ids']), 'same isect lists', np.array_equal(o64['flatten_ids'],o32['flatten_ids']))
    print('oracle f32 vs f64:')
    for k,kk in [('renders','img'),('alphas','alpha')]: cmp(k,o32[k][0].squeeze(),o64[k][0].squeeze())
    for k in ['v_means','v_quats','v_scales','v_opacities','v_colors']: cmp(k,o32[k].squeeze(),o64[k].squeeze())
    for dt in (np.float64,np.float32):
        print('mine',dt.__name__,'vs oracle f64:')
        r=blend(sc,o64,vrc,vra,dt)
        cmp('renders',r['img'],o64['renders'][0]); cmp('alphas',r['alpha'],o64['alphas'][0,...,0]); print('  last_ids mismatches',(r['last']!=o64['last_ids'][0]).sum())
        for k in ['v_means','v_quats','v_scales','v_opacities','v_colors']: cmp(k,r[k],o64[k].squeeze())
