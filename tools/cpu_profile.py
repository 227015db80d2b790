"""cProfile of the host side of the proposal chain (where does the CPU spend the step?)."""
import cProfile, pstats, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from d3net_b200 import chain, scenes, pointgroup_ops as ops, dist as pgdist
nb = scenes.make_batch(8, 150000, config_id=2, with_feats=False)
batch = chain.batch_to_device(nb, torch.device("cuda"))
rand6 = torch.full((6,), 0.5, device="cuda")
def step():
    out = chain.proposal_chain(ops, batch, rand6)
    return pgdist.pack_proposals(out, batch, 256)
for _ in range(3): step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(10): step()
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) * 100)
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(22)
