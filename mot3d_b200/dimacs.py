"""Interchange with muSSP-style min-cost-flow graphs: arc lists in the layout muSSP requires
(node 1 = source, node V = sink, even ids = pre-nodes, odd ids = post-nodes; zip!/muSSP-master/muSSP/Sink.cpp:23-28,
Graph.cpp:89-101) -> the solver's inputs. This is format conversion on the host (numpy); the solve itself runs in
the CUDA kernel.

The solver sweeps nodes layer by layer, so detections are renumbered by their longest-path layer (for graphs made
by build_graph that is the frame order; tracklet graphs and public MOT graphs carry no frame index)."""
import numpy as np
import torch

from . import _lib
from .batch import DetectionBatch, GraphBatch

INF = _lib.INF_COST


def parse_text(lines):
    """["p min V E", "a tail head w", ...] -> (V, tail, head, cost int64 in 1e-7 units); decimal-exact."""
    n = int(lines[0].split()[2])
    tail, head, cost = [], [], []
    for l in lines[1:]:
        if not l or l[0] != "a":
            continue
        _, s, t, x = l.split()
        neg = x.startswith("-")
        whole, _, frac = x.lstrip("+-").partition(".")
        frac = (frac + "0000000")[:7]
        v = int(whole or 0) * 10 ** 7 + int(frac)
        tail.append(int(s))
        head.append(int(t))
        cost.append(-v if neg else v)
    return n, np.array(tail, np.int64), np.array(head, np.int64), np.array(cost, np.int64)


def graph_from_arcs(n_nodes, tail, head, cost, device=None):
    """1-based arcs of a muSSP-layout graph -> (GraphBatch ready for solve_graphs, perm) where perm[k] = original
    detection index (0-based, node 2+2m <-> m) of the k-th detection in solver order."""
    tail = np.asarray(tail, dtype=np.int64)
    head = np.asarray(head, dtype=np.int64)
    cost = np.asarray(cost, dtype=np.int64)
    if n_nodes % 2 or n_nodes < 2:
        raise ValueError("a tracking graph has 2N+2 nodes")
    n = (n_nodes - 2) // 2
    S, T = 1, n_nodes
    en = np.full(n, INF, dtype=np.int64)
    ex = np.full(n, INF, dtype=np.int64)
    cn = np.full(n, INF, dtype=np.int64)
    m = tail == S
    if not (head[m] % 2 == 0).all():
        raise ValueError("source arcs must enter pre-nodes (even ids)")
    en[(head[m] - 2) // 2] = cost[m]
    m2 = head == T
    if not (tail[m2] % 2 == 1).all():
        raise ValueError("sink arcs must leave post-nodes (odd ids)")
    ex[(tail[m2] - 3) // 2] = cost[m2]
    m3 = (tail % 2 == 0) & (head == tail + 1) & (tail != S)
    cn[(tail[m3] - 2) // 2] = cost[m3]
    if (cn >= INF).any():
        raise ValueError("every detection needs its pre->post arc")
    m4 = ~(m | m2 | m3)
    if not ((tail[m4] % 2 == 1) & (head[m4] % 2 == 0)).all():
        raise ValueError("transition arcs must go from a post-node to a pre-node")
    src = (tail[m4] - 3) // 2
    dst = (head[m4] - 2) // 2
    c = cost[m4]
    # longest-path layers
    level = np.zeros(n, dtype=np.int64)
    for _ in range(n + 1):
        new = level.copy()
        if len(src):
            np.maximum.at(new, dst, level[src] + 1)
        if (new == level).all():
            break
        level = new
    else:
        raise ValueError("the transition arcs contain a cycle")
    perm = np.argsort(level, kind="stable")
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    s2, d2 = inv[src], inv[dst]
    o = np.lexsort((s2, d2))
    s2, d2, c2 = s2[o], d2[o], c[o]
    in_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(in_ptr, d2 + 1, 1)
    in_ptr = np.cumsum(in_ptr)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def t(a, dt):
        return torch.from_numpy(np.ascontiguousarray(a.astype(dt))).to(dev)
    gb = GraphBatch()
    gb.direction = _lib.DIR_IN
    gb.tkey = t(level[perm], np.int32)
    gb.entry_cost, gb.node_cost, gb.exit_cost = t(en[perm], np.int64), t(cn[perm], np.int64), t(ex[perm], np.int64)
    gb.row_ptr = t(in_ptr, np.int32)
    gb.col = t(s2 if len(s2) else np.zeros(1, np.int64), np.int32)
    gb.cost = t(c2 if len(c2) else np.zeros(1, np.int64), np.int64)
    gb.n_edges = int(len(s2))
    b = DetectionBatch(t(np.array([0, n]), np.int32), gb.tkey, None, None)
    b._max_nodes = n
    gb.dets = b
    return gb, perm
