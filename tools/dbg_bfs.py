import sys, ctypes, torch, numpy as np, time
sys.path.insert(0,'.')
from d3net_b200 import scenes, chain, pointgroup_ops as ops, PG_OP, _native
nb = scenes.make_batch(8, 150000, config_id=2)
b = chain.batch_to_device(nb, torch.device('cuda'))
sem = b['semantic_preds']; obj = torch.nonzero(sem>0).view(-1)
bi = b['locs_scaled'][:,0].int()[obj].contiguous(); bo = chain.get_batch_offsets(bi, 8)
xyz = (b['locs'][obj] + b['pt_offsets'][obj]).contiguous(); sem_ = sem[obj].int().contiguous()
idx, sl = ops.ballquery_batch_p(xyz, bi, bo, 0.03, 300)
print('n', xyz.shape[0], 'nA', idx.numel(), 'capped', int((sl[:,1]>=1000).sum()))
for it in range(3):
    torch.cuda.synchronize(); t=time.time()
    ci, co, g = PG_OP.bfs_cluster_impl(sem_, idx, sl, 50)
    torch.cuda.synchronize(); dt=time.time()-t
    d=(ctypes.c_longlong*5)(); _native.lib().pg_bfs_cluster_debug(d)
    print('bfs ms', dt*1e3, 'generic', g, 'dbg', list(d), 'nC', co.numel()-1)
