"""Time decomposition of igcn_tc_candidates: runs the scoring of a bench workload once per IGCN_TC_EXPERIMENT
variant (eval_tc.cu: 0 production, 2 no filter, 3 filter without hits, 4 no item-image stream, 5 no TMEM reads,
6 neither, 7 = experimental threshold-in-MMA filter with valid results) and prints the CUDA-event time of the candidates launch alone.  Variants other than 0 produce invalid
lists, so only igcn_tc_pack + igcn_tc_candidates are called here.
IGCN_TC_DEBUG flags (TcArgs.dbg) can be appended as variant:flags.
    python tools/tc_floor.py [workload] [variant[:flags] ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else 'yelp-lightgcn'
    variants = sys.argv[2:] or ['0', '3', '2', '4', '5', '6']
    shape, kind, l2_reg, dropout = bench.WORKLOADS[workload]
    dev = torch.device('cuda:0')
    ds = bench.build_dataset(shape, dev)
    model, trainer = bench.build_model(ds, kind, dropout, l2_reg, dev, use_graph=False)
    from igcn_cf_b200 import engine
    from igcn_cf_b200._lib import call, ptr, stream_ptr
    rep = model.get_rep().detach()
    n_users, n_items, D, k = ds.n_users, ds.n_items, rep.shape[1], 20
    users = torch.arange(n_users, device=dev)
    mask = trainer._mask_csr('val')
    scorer = engine.TcScorer()
    n_head, n_splits = scorer.plan_ctas((n_users + 127) // 128)
    ws = scorer._workspace(n_users, n_items, D, n_splits, k, dev)
    tile_ptr, entries = mask.tiles(n_items, None)
    for v in variants:
        os.environ['IGCN_TC_EXPERIMENT'], os.environ['IGCN_TC_DEBUG'] = (v.split(':') + ['0'])[:2]
        # packed per variant: variant 7 needs the threshold multiplier in the item image
        call('igcn_tc_pack', ptr(rep), rep.numel(), ptr(users), n_users, n_users, n_items, D, ptr(ws['maxabs']),
             ptr(ws['a_img']), ptr(ws['b_img']), ptr(ws['center']), ptr(ws['center_scratch']), stream_ptr())
        def launch():
            call('igcn_tc_candidates', ptr(ws['a_img']), ptr(ws['b_img']), n_users, n_items, D, n_splits, n_head, 0, n_items,
                 None, ptr(tile_ptr), ptr(entries), ptr(ws['cand_items']), ptr(ws['cand_cnt']), ptr(ws['cand_thr']), None,
                 stream_ptr())
        for _ in range(3):
            launch()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            launch()
        e1.record()
        torch.cuda.synchronize()
        print('%s variant %s (n_head %d, n_splits %d): %.3f ms' % (workload, v, n_head, n_splits, e0.elapsed_time(e1) / reps), flush=True)


if __name__ == '__main__':
    main()
