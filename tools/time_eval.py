"""Times the evaluation pass of a bench workload under several item-split settings of the tensor-core scoring
path:   python tools/time_eval.py [workload] [splits ...]      (0 = engine default plan, s > 0 = s uniform splits)
Prints one line per setting: device ms per recommend pass (CUDA events), users sent to the exact kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else 'yelp-lightgcn'
    settings = sys.argv[2:] or ['0', '1', '2']
    shape, kind, l2_reg, dropout = bench.WORKLOADS[workload]
    dev = torch.device('cuda:0')
    ds = bench.build_dataset(shape, dev)
    model, trainer = bench.build_model(ds, kind, dropout, l2_reg, dev, use_graph=False)
    from igcn_cf_b200 import engine
    ref = None
    for s in settings:
        os.environ['IGCN_TC_SPLITS'] = s
        for _ in range(3):
            model._bump()
            rec = trainer.recommend_local('val')
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            model._bump()
            rec = trainer.recommend_local('val')
        e1.record()
        torch.cuda.synchronize()
        same = True if ref is None else bool(torch.equal(ref, rec))
        ref = rec if ref is None else ref
        print('%s splits=%s: %.3f ms/pass, fallback users %d, same lists as first setting: %s'
              % (workload, s, e0.elapsed_time(e1) / reps, int(engine._tc_scorer.last_fallback.item()), same), flush=True)


if __name__ == '__main__':
    main()
