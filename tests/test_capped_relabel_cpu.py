"""CPU restatement of the rule k_maxflow_cluster uses for the local-optimisation cut (csrc/pxb_maxflow_cluster.cu): the
FIRST global relabel may stop after `cap` BFS levels and give everything it has not reached the label cap + 2 -- a valid
(not exact) labelling --, a capped round never ends the cut, every later relabel is exact; the cut is read off the last
exact relabel (a node is on the SINK side iff it still reaches the sink in the residual graph). Checked here on random
graphs against scipy's maximum flow: same flow value, same (maximal) sink side. Integer capacities, so no rounding ties."""
import numpy as np
import pytest
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import maximum_flow

INF = 10**9


def relabel(n, arcs, cap_res, sink_cap, cap_levels):
    """distance-to-sink labels by backward BFS over residual arcs; cap_levels > 0: stop after that many levels"""
    h = [1 if sink_cap[u] > 0 else INF for u in range(n)]
    level = 1
    while True:
        if cap_levels and level > cap_levels:
            return [cap_levels + 2 if x == INF else x for x in h], False
        found = False
        for u in range(n):
            if h[u] != INF:
                continue
            if any(cap_res[a] > 0 and h[v] == level for a, v in arcs[u]):
                h[u] = level + 1
                found = True
        if not found:
            return h, True
        level += 1


def min_cut_push_relabel(n, edges, excess0, sink_cap0, cap_levels, cycles):
    arcs = [[] for _ in range(n)]  # arcs[u] = [(arc id, head)], mirror of arc a is a ^ 1
    cap_res = []
    for u, v, c in edges:
        arcs[u].append((len(cap_res), v))
        cap_res.append(c)
        arcs[v].append((len(cap_res), u))
        cap_res.append(c)
    excess, sink_cap = list(excess0), list(sink_cap0)
    flow = 0
    for rnd in range(10 * n):
        capped = rnd == 0 and cap_levels > 0
        h, exact = relabel(n, arcs, cap_res, sink_cap, cap_levels if capped else 0)
        active = [u for u in range(n) if excess[u] > 0 and h[u] != INF]
        if not active and exact:
            return flow, {u for u in range(n) if h[u] != INF}
        for _ in range(cycles if not capped else max(1, cycles // 8)):
            busy = False
            for u in range(n):
                if excess[u] <= 0 or h[u] == INF:
                    continue
                busy = True
                d = min(excess[u], sink_cap[u])
                if d > 0:
                    sink_cap[u] -= d
                    excess[u] -= d
                    flow += d
                if excess[u] <= 0:
                    continue
                res = [(h[v], a, v) for a, v in arcs[u] if cap_res[a] > 0]
                if not res or min(res)[0] == INF:
                    h[u] = INF  # stranded
                    continue
                hv, a, v = min(res)
                if h[u] > hv:
                    d = min(excess[u], cap_res[a])
                    cap_res[a] -= d
                    cap_res[a ^ 1] += d
                    excess[u] -= d
                    excess[v] += d
                else:
                    h[u] = hv + 1 if hv + 1 < n + 2 else INF
            if not busy:
                break
    raise AssertionError("no convergence")


def reference_cut(n, edges, excess0, sink_cap0):
    S, T = n, n + 1
    rows, cols, vals = [], [], []
    for u, v, c in edges:
        rows += [u, v]
        cols += [v, u]
        vals += [c, c]
    for u in range(n):
        if excess0[u] > 0:
            rows.append(S), cols.append(u), vals.append(excess0[u])
        if sink_cap0[u] > 0:
            rows.append(u), cols.append(T), vals.append(sink_cap0[u])
    g = csr_matrix((vals, (rows, cols)), shape=(n + 2, n + 2), dtype=np.int32)
    res = maximum_flow(g, S, T)
    residual = (g - res.flow).toarray()  # residual[u, v] > 0: arc u -> v still has capacity (reverse arcs included by -flow)
    reach = {T}
    stack = [T]
    while stack:
        v = stack.pop()
        for u in range(n + 2):
            if u not in reach and residual[u, v] > 0:
                reach.add(u)
                stack.append(u)
    return res.flow_value, {u for u in reach if u < n}


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("cap_levels", [0, 1, 4])
def test_capped_first_relabel_gives_the_same_cut(seed, cap_levels):
    rng = np.random.default_rng(seed)
    n = 60
    pts = rng.uniform(size=(n, 2))
    edges = []
    for u in range(n):  # a 3-nearest-neighbour graph: long chains like the kernel's neighbourhood graphs
        d = np.linalg.norm(pts - pts[u], axis=1)
        for v in np.argsort(d)[1:4]:
            if u < v:
                edges.append((u, int(v), int(rng.integers(1, 4))))
    t = rng.integers(-6, 7, size=n)
    left = pts[:, 0] < 0.5
    t = np.where(left, np.abs(t), -np.abs(t))  # sources on one half, sinks on the other: flow has to cross the graph
    excess0 = [int(max(x, 0)) for x in t]
    sink_cap0 = [int(max(-x, 0)) for x in t]
    flow_ref, sink_ref = reference_cut(n, edges, excess0, sink_cap0)
    flow, sink_side = min_cut_push_relabel(n, edges, excess0, sink_cap0, cap_levels, cycles=64)
    assert flow == flow_ref
    assert sink_side == sink_ref
