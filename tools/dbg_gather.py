import torch, time
import torch.nn.functional as F
dev='cuda'
N, S = 1200000, 1934572
feats = torch.randn(N,16,device=dev); coords=torch.randn(N,3,device=dev)
idx = torch.randint(0,N,(S,),device=dev)
def t(fn, name):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize(); print(name, a.elapsed_time(b)/10, 'ms')
t(lambda: feats[idx], 'feats[idx] C=16')
t(lambda: torch.index_select(feats,0,idx), 'index_select C=16')
t(lambda: F.embedding(idx, feats), 'embedding C=16')
t(lambda: torch.gather(feats,0,idx[:,None].expand(-1,16)), 'gather C=16')
t(lambda: coords[idx], 'coords[idx] C=3')
t(lambda: torch.index_select(coords,0,idx), 'index_select C=3')
t(lambda: F.embedding(idx, coords), 'embedding C=3')
i32 = idx.int()
t(lambda: i32.long(), 'int->long')
t(lambda: feats[i32], 'feats[int32 idx]')
