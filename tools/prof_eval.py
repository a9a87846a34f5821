"""Driver for ncu captures of the evaluation path: python tools/prof_eval.py [workload] [impl]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else 'gowalla-lightgcn'
    shape, kind, l2_reg, dropout = bench.WORKLOADS[workload]
    dev = torch.device('cuda:0')
    ds = bench.build_dataset(shape, dev)
    model, trainer = bench.build_model(ds, kind, dropout, l2_reg, dev, use_graph=True)
    if len(sys.argv) > 2:
        trainer.config['score_impl'] = sys.argv[2]
    model.train()
    for _ in range(int(os.environ.get('PROF_TRAIN_STEPS', 300))):      # evaluate a partly trained model, as the bench does
        trainer.step.run()
    for i in range(2):
        model._bump()
        torch.cuda.synchronize()
        if i == 1:
            torch.cuda.nvtx.range_push('eval')           # ncu --nvtx --nvtx-include "eval/" captures this evaluation only
        print(trainer.eval('val')[0])
        torch.cuda.synchronize()
        if i == 1:
            torch.cuda.nvtx.range_pop()


if __name__ == '__main__':
    main()
