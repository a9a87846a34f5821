"""Minimal stand-in for the two DGL entry points the reference's hot path uses.

TEST INFRASTRUCTURE ONLY.  The reference calls `dgl.graph((column, row), num_nodes=n)` and
`dgl.ops.gspmm(g, 'mul', 'sum', lhs_data=X, rhs_data=vals)` (/root/reference/model.py:99-102,
428-431, 439-442).  Edges are (src=column, dst=row), so the result is out[row] = sum_e
vals[e] * X[column[e]] == (A @ X)[row]: the documented u_mul_e_sum semantics.
"""
from . import ops  # noqa: F401


class _Graph:
    def __init__(self, src, dst, num_nodes):
        self.src, self.dst, self.num_nodes = src, dst, num_nodes


def graph(edges, num_nodes=None, device=None):
    src, dst = edges
    return _Graph(src, dst, num_nodes)
