"""`dgl.ops.gspmm` restricted to the ('mul', 'sum') mode used on the hot path."""
import torch


def gspmm(g, op, reduce_op, lhs_data=None, rhs_data=None):
    if op != 'mul' or reduce_op != 'sum':
        raise NotImplementedError('shim only covers gspmm(mul, sum)')
    n = g.num_nodes
    adj = torch.sparse_coo_tensor(torch.stack([g.dst, g.src]), rhs_data, (n, n))
    return torch.sparse.mm(adj, lhs_data)
