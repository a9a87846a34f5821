"""Small driver for ncu captures: build one workload, run a few eager training steps and one
evaluation.  Usage: python tools/prof_step.py [workload] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else 'yelp-lightgcn'
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    shape, kind, l2_reg, dropout = bench.WORKLOADS[workload]
    dev = torch.device('cuda:0')
    ds = bench.build_dataset(shape, dev)
    model, trainer = bench.build_model(ds, kind, dropout, l2_reg, dev, use_graph=False)
    model.train()
    for _ in range(steps):
        trainer.step.run()
    torch.cuda.synchronize()
    print('loss', trainer.step.loss.item())
    print(trainer.eval('val')[0])


if __name__ == '__main__':
    main()
