import sys, time; sys.path.insert(0, '.')
import numpy as np
import gl_slam_b200 as g
from gl_slam_b200 import scene
from oracle import oracle
ctx = g.Context(0)
for name, kw in [('tiny', dict(n_cam=6,n_pt=120,track_len=4,seed=11,outlier_frac=0.05,rot_sigma=0.01)), ('C2', None), ('C1', None)]:
    p = scene.make_scene(**kw) if kw else scene.config(name)
    # linearize parity
    Lg = ctx.linearize(p, 1e4); Lo = oracle.linearize(p, 1e4)
    def rel(a,b): return float(np.abs(a-b).max()/max(np.abs(b).max(),1e-300))
    print(name, 'cost', Lg.cost, Lo.cost, 'res', rel(Lg.residuals,Lo.residuals), 'jc', rel(Lg.jac_cam,Lo.jac_cam), 'jp', rel(Lg.jac_pt,Lo.jac_pt),
          'gc', rel(Lg.grad_cam,Lo.grad_cam), 'gp', rel(Lg.grad_pt,Lo.grad_pt), 'Hc', rel(Lg.hess_cam,Lo.hess_cam), 'Hp', rel(Lg.hess_pt,Lo.hess_pt),
          'Sd', rel(Lg.schur_diag,Lo.schur_diag), 'rhs', rel(Lg.schur_rhs,Lo.schur_rhs))
    t=time.time(); rg, sg = ctx.solve(p); tg=time.time()-t
    ro, so = oracle.solve(p)
    n=min(len(sg['cost']),len(so['cost']))
    print('  iters', sg['n_iters'], so['n_iters'], 'stop', sg['stop_reason'], so['stop_reason'], 'maxrel cost', max(abs(a-b)/abs(b) for a,b in zip(sg['cost'][:n],so['cost'][:n])),
          'cam', float(np.abs(rg.cam-ro.cam).max()), 'pt', float(np.abs(rg.pt-ro.pt).max()), 'cg', sg['cg_iters'][:8], 'time %.3f'%tg,
          't ms', [round(sg[k],3) for k in ('t_setup_ms','t_linearize_ms','t_schur_ms','t_solve_ms','t_update_ms')])
cam0,X,uv,gt = scene.pose_only_scene(500, 7)
pc, ps = ctx.pose_only(cam0,X,uv,scene.KITTI_K); oc, os_ = oracle.pose_only(cam0,X,uv,scene.KITTI_K)
print('pose iters', ps['n_iters'], os_['n_iters'], 'cam diff', np.abs(pc-oc).max(), 'gt diff', np.abs(pc-gt).max(), 'cost', ps['final_cost'], os_['final_cost'], 'ms', ps['t_total_ms'])
print('launches', g.kernel_launch_count())
# diagnose C2 parity per iteration, PCG tolerance sweep
p = scene.config('C2')
ro, so = oracle.solve(p)
for tol in (1e-13, 1e-15):
    rg, sg = ctx.solve(p, g.options(cg_rel_tol=tol, cg_max_iters=2000))
    print('tol', tol, 'iters', sg['n_iters'], ['%.1e'%(abs(a-b)/abs(b)) for a,b in zip(sg['cost'],so['cost'])], sg['cg_iters'][:12])
p2 = scene.config('C2', outlier_frac=0.0)
ro, so = oracle.solve(p2); rg, sg = ctx.solve(p2)
print('no outliers: iters', sg['n_iters'], so['n_iters'], ['%.1e'%(abs(a-b)/abs(b)) for a,b in zip(sg['cost'],so['cost'])], 'cam', np.abs(rg.cam-ro.cam).max(), 'pt', np.abs(rg.pt-ro.pt).max())
