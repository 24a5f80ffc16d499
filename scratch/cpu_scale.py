import sys,time,os; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, bench
from oracle import oracle as orc
orc.build()
print('max_threads', orc.max_threads(), 'cpu.max', open('/sys/fs/cgroup/cpu.max').read().strip() if os.path.exists('/sys/fs/cgroup/cpu.max') else None, 'affinity', len(os.sched_getaffinity(0)))
sc, what = bench.cpu_sample_scene(1000000)
target = np.full((1, sc["height"], sc["width"], 3), 0.5, np.float32)
for th in (8, 32, 64, 128):
    orc.set_threads(th)
    bench.oracle_step(orc, sc, target)
    t0=time.perf_counter(); o=orc.render_pipeline(sc, precision="f32"); t1=time.perf_counter()
    v=np.ones_like(o['renders']); va=np.zeros_like(o['alphas'])
    g=orc.raster_bwd(sc['means'],sc['quats'],sc['scales'],o['colors'],sc['opacities'][None],sc['background'],None,sc['width'],sc['height'],16,sc['viewmats'],sc['Ks'],o['tile_offsets'],o['flatten_ids'],o['alphas'],o['last_ids'],v,va); t2=time.perf_counter()
    print(th, 'fwd pipeline %.3f s  raster_bwd %.3f s'%(t1-t0,t2-t1))
