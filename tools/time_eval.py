"""Stage-by-stage wall/GPU timing of trainer.recommend (debug aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from igcn_cf_b200 import engine

workload = sys.argv[1] if len(sys.argv) > 1 else 'yelp-lightgcn'
shape, kind, l2_reg, dropout = bench.WORKLOADS[workload]
dev = torch.device('cuda:0')
ds = bench.build_dataset(shape, dev)
model, trainer = bench.build_model(ds, kind, dropout, l2_reg, dev, use_graph=False)
trainer.eval('val')
torch.cuda.synchronize()
def T(f, name, n=3):
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
        print('%-28s %.3f ms' % (name, (time.perf_counter() - t0) * 1e3))
    return r
model.eval()
def rep():
    model._bump()
    with torch.no_grad():
        return model.get_rep().contiguous()
r = T(rep, 'get_rep')
mask = trainer._mask_csr('val')
T(lambda: mask.tiles(ds.n_items, None), 'mask.tiles')
users = trainer.test_users
T(lambda: engine.score_topk(r, users, model.n_users, model.n_items, 20, mask=mask, impl='tc'), 'score_topk tc')
T(lambda: engine._tc_scorer.topk(r, users, model.n_users, model.n_items, 20, mask), 'scorer.topk')
T(lambda: trainer.recommend('val'), 'recommend')
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); trainer.recommend('val'); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(12)
