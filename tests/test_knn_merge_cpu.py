"""CPU restatement of k_knn_graph's list logic (csrc/pxb_fit.cu): candidates visited in ascending index order and inserted
in front of the first STRICTLY larger entry (everything behind moves down one slot), slices of the candidate range merged
in slice order the same way. The result must be the k smallest (distance, index) pairs -- the list one sequential scan
returns -- also with exactly equal distances (duplicated points). The old insertion compared a displaced entry again and
let it leap-frog an equal neighbour: `leapfrog=True` restates that bug and the test shows it breaks the tie rule."""
import numpy as np
import pytest


def insert(lst, d, j, k, leapfrog=False):
    """lst: list of (d, j) of length k (unused slots (inf, -1)); the kernel's unrolled compare-and-swap pass"""
    cd, cj = d, j
    placed = False
    for q in range(k):
        if (placed and not leapfrog) or cd < lst[q][0]:
            lst[q], (cd, cj) = (cd, cj), lst[q]
            placed = True


def scan(d2, lo, hi, i, radius2, k, leapfrog=False):
    lst = [(np.inf, -1)] * k
    for j in range(lo, hi):
        if j != i and d2[j] <= radius2 and d2[j] < lst[k - 1][0]:
            insert(lst, d2[j], j, k, leapfrog)
    return lst


def sliced(d2, i, radius2, k, slices):
    n = len(d2)
    chunk = (n + slices - 1) // slices
    lists = [scan(d2, min(n, s * chunk), min(n, s * chunk + chunk), i, radius2, k) for s in range(slices)]
    best = lists[0]
    for other in lists[1:]:
        for d, j in other:
            if j < 0 or not d < best[k - 1][0]:
                break
            insert(best, d, j, k)
    return [j for _, j in best if j >= 0]


@pytest.mark.parametrize("n,k,slices,seed", [(257, 5, 8, 0), (100, 8, 8, 1), (999, 12, 4, 2), (7, 8, 8, 3)])
def test_sliced_scan_equals_the_global_rule(n, k, slices, seed):
    rng = np.random.default_rng(seed)
    pts = np.round(rng.uniform(0, 12, size=(n, 2)))  # integer coordinates: many exactly equal distances, duplicated points
    radius2 = 25.0
    for i in range(0, n, max(1, n // 40)):
        d2 = ((pts - pts[i]) ** 2).sum(1)
        d2_self = d2.copy()
        d2_self[i] = np.inf
        want = [int(j) for j in np.argsort(d2_self, kind="stable")[:k] if d2_self[j] <= radius2]
        assert sliced(d2, i, radius2, k, slices) == want
        assert [j for _, j in scan(d2, 0, n, i, radius2, k) if j >= 0] == want  # one slice: the sequential scan itself


def test_the_old_insertion_broke_ties():
    d2 = np.array([9.0, 5.0, 5.0, 1.0, 7.0])  # candidates 1 and 2 tie; candidate 3 displaces both
    lst = scan(d2, 0, 5, -1, 100.0, 4, leapfrog=True)
    assert [j for _, j in lst] == [3, 2, 1, 4]  # 1 leap-frogged its equal neighbour 2
    lst = scan(d2, 0, 5, -1, 100.0, 4)
    assert [j for _, j in lst] == [3, 1, 2, 4]
