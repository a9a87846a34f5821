"""Time decomposition of igcn_tc_candidates: runs the scoring of a bench workload once per IGCN_TC_EXPERIMENT
variant (eval_tc.cu: 0 production, 2 no filter, 3 compare-free path only, 5 no TMEM reads) and prints the CUDA-event time
of the candidates launch alone, for the natural and the popularity scan order, plus the filter statistics of the
production kernel.  Variants other than 0 produce invalid lists, so only igcn_tc_pack + igcn_tc_candidates are called.
IGCN_TC_DEBUG flags (TcArgs.dbg) can be appended as variant:flags.
    python tools/tc_floor.py [workload] [variant[:flags] ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else 'yelp-lightgcn'
    variants = sys.argv[2:] or ['0', '3', '2', '5']
    shape, kind, l2_reg, dropout = bench.WORKLOADS[workload]
    dev = torch.device('cuda:0')
    ds = bench.build_dataset(shape, dev)
    model, trainer = bench.build_model(ds, kind, dropout, l2_reg, dev, use_graph=True)
    from igcn_cf_b200 import engine
    from igcn_cf_b200._lib import call, ptr, stream_ptr
    model.train()
    for _ in range(int(os.environ.get('TC_FLOOR_TRAIN_STEPS', 300))):     # embeddings with some structure, like after the bench's steps
        trainer.step.run()
    model.eval()
    rep = model.get_rep().detach()
    n_users, n_items, D, k = ds.n_users, ds.n_items, rep.shape[1], 20
    users = torch.arange(n_users, device=dev)
    mask = trainer._mask_csr('val')
    scorer = engine.TcScorer()
    n_head, n_splits = scorer.plan_ctas((n_users + 127) // 128, (n_items + 255) // 256)
    ws = scorer._workspace(n_users, n_items, D, n_splits, k, dev)
    for oname, order in (('natural', None), ('popularity', trainer.item_order())):
        tile_ptr, entries = mask.tiles(n_items, None, order)
        perm = None if order is None else order.perm
        call('igcn_tc_pack', ptr(rep), rep.numel(), ptr(users), n_users, n_users, n_items, D, ptr(perm), ptr(ws['maxabs']),
             ptr(ws['a_img']), ptr(ws['b_img']), ptr(ws['center']), ptr(ws['center_scratch']), stream_ptr())
        stats = torch.zeros(5, dtype=torch.int64, device=dev)

        def launch(st=None):
            call('igcn_tc_candidates', ptr(ws['a_img']), ptr(ws['b_img']), n_users, n_items, D, n_splits, n_head, 0, n_items,
                 None, ptr(tile_ptr), ptr(entries), ptr(ws['cand_items']), ptr(ws['cand_cnt']), ptr(ws['cand_thr']), None,
                 ptr(st), stream_ptr())
        for v in variants:
            os.environ['IGCN_TC_EXPERIMENT'], os.environ['IGCN_TC_DEBUG'] = (v.split(':') + ['0'])[:2]
            for _ in range(3):
                launch()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                launch()
            e1.record()
            torch.cuda.synchronize()
            print('%s order %-10s variant %s (n_head %d, n_splits %d): %.3f ms' % (workload, oname, v, n_head, n_splits,
                                                                                 e0.elapsed_time(e1) / reps), flush=True)
        os.environ['IGCN_TC_EXPERIMENT'] = '0'
        launch(stats)
        torch.cuda.synchronize()
        ch, slow, groups, hits, comp = stats.tolist()
        print('%s order %-10s stats: chunks %d, off the compare-free path %.1f%%, groups compared per chunk %.3f, '
              'candidates appended %.4f%% of scores, compactions %d' % (workload, oname, ch, 100.0 * slow / max(1, ch),
                                                                       groups / max(1, ch), 100.0 * hits / max(1, ch * 1024), comp), flush=True)


if __name__ == '__main__':
    main()
