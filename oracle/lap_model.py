"""TEST INFRASTRUCTURE: a numpy model of the WARP-COOPERATIVE assignment solver of brille_b200/csrc/sortpairs.cu.

The device solver does not run the reference's sequential loops (lapjv.hpp:281-538); it replaces every scan over the columns by
a lane-parallel formulation that provably yields the same decisions, ties included:

* column minima            -> per-column arg-min with lowest-row tie-break; a row keeps the LARGEST column that chose it
                              (the reference walks the columns downwards and the first claim wins)            lapjv.hpp:311-335
* dual transfer            -> the reference's tolerant running minimum ``if (h < mn + eps) mn = h`` is a recurrence, not a
                              minimum; it is evaluated 32 columns at a time: flag the columns below mn + eps, take the first,
                              update mn, re-flag the columns behind it                                        lapjv.hpp:340-358
* row bidding              -> (umin, j1) = lexicographic min of (h_j, j); (usubmin, j2) = the same over j != j1
                              (what find_umins_plain computes, lapjv.hpp:74-99)                               lapjv.hpp:365-410
* shortest augmenting path -> the "new minimum or tie with the running minimum" columns of a scan are the prefix-minimum
                              records of d over the to-do list: found with a prefix-min, applied in list order; a relaxation
                              step flags (improved, ties the minimum, unassigned) per column, the first unassigned tie ends
                              the path, the ties before it are moved to the ready list in list order         lapjv.hpp:416-520

This file restates those formulations with numpy (vector operations where the device uses the lanes of a warp) so that they can
be checked on the CPU against the sequential restatement (oracle/sort_oracle.c: oracle_lapjv_batch), including cost matrices
full of ties.  It is never imported by the product.
"""
from __future__ import annotations

import numpy as np

DBL_MAX = np.finfo(np.float64).max


def _eps_scan(h, skip, eps):
    """mn after ``for j: if j != skip and h[j] < mn + eps: mn = h[j]`` starting from DBL_MAX, 32 columns at a time."""
    mn = DBL_MAX
    n = len(h)
    for base in range(0, n, 32):
        hh = h[base:base + 32]
        idx = np.arange(base, base + len(hh))
        start = 0
        while True:
            with np.errstate(over="ignore"):
                flag = (hh < mn + eps) & (idx != skip) & (np.arange(len(hh)) >= start)
            if not flag.any():
                break
            first = int(np.argmax(flag))
            mn = float(hh[first])
            start = first + 1
    return mn


def _lexmin(h, exclude=-1):
    """(value, index) of the minimum with the lowest index among equals, ignoring column ``exclude``."""
    hh = h.copy()
    if exclude >= 0:
        hh[exclude] = np.inf
        if not np.isfinite(hh).any():
            return DBL_MAX, -1
    j = int(np.argmin(hh))  # numpy: first occurrence
    return float(hh[j]), j


def solve(cost):
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    n = cost.shape[0]
    if n == 1:
        return np.zeros(1, np.int32), np.zeros(1, np.int32)
    total = 0.0
    for x in cost.ravel():  # the reference sums in storage order with one accumulator (lapjv.hpp:305-306)
        total += float(x)
    eps = total / float(10000 * n)
    # ---- column minima -----------------------------------------------------------------------------------------
    imin = np.argmin(cost, axis=0)  # first row among equals
    v = cost[imin, np.arange(n)].copy()
    matches = np.bincount(imin, minlength=n)
    rowsol = np.full(n, -1, dtype=np.int64)
    np.maximum.at(rowsol, imin, np.arange(n))
    colsol = np.where(rowsol[imin] == np.arange(n), imin, -1).astype(np.int64)
    # ---- dual transfer ------------------------------------------------------------------------------------------
    free = []
    for i in range(n):
        if matches[i] == 0:
            free.append(i)
        elif matches[i] == 1:
            j1 = int(rowsol[i])
            v[j1] = v[j1] - _eps_scan(cost[i] - v, j1, eps)
    # ---- row bidding (two sweeps) -----------------------------------------------------------------------------
    for _ in range(2):
        todo, free = free, []
        k = 0
        while k < len(todo):
            i = todo[k]
            k += 1
            h = cost[i] - v
            umin, j1 = _lexmin(h)
            usub, j2 = _lexmin(h, j1)
            i0 = int(colsol[j1])
            vnew = v[j1] - (usub + eps - umin)
            lowers = vnew < v[j1]
            if lowers:
                v[j1] = vnew
            elif i0 != -1:
                j1 = j2
                i0 = int(colsol[j2])
            rowsol[i] = j1
            colsol[j1] = i
            if i0 != -1:
                if lowers:
                    k -= 1
                    todo[k] = i0
                else:
                    free.append(i0)
    # ---- shortest augmenting paths ----------------------------------------------------------------------------
    for fr in free:
        d = cost[fr] - v
        pred = np.full(n, fr, dtype=np.int64)
        cl = np.arange(n)
        low = up = 0
        last = 0
        mn = 0.0
        end = -1
        while end < 0:
            if up == low:
                last = low - 1
                # prefix-minimum records of d over the to-do list cl[low:], in list order
                dd = d[cl[low:]]
                run = np.minimum.accumulate(dd)
                prev = np.concatenate([[np.inf], run[:-1]])
                strict = dd < prev
                event = dd <= prev
                event[0] = strict[0] = True
                for p in np.nonzero(event)[0]:
                    kk = low + int(p)
                    j = int(cl[kk])
                    if strict[p]:
                        up = low
                        mn = float(dd[p])
                    cl[kk] = cl[up]
                    cl[up] = j
                    up += 1
                ready = cl[low:up]
                un = np.nonzero(colsol[ready] == -1)[0]
                if len(un):
                    end = int(ready[un[0]])
                    break
            j1 = int(cl[low])
            low += 1
            i = int(colsol[j1])
            hh = cost[i, j1] - v[j1] - mn
            ks = np.arange(up, n)
            js = cl[ks]
            v2 = cost[i, js] - v[js] - hh
            upd = v2 < d[js]
            tie = upd & (v2 == mn)
            term = tie & (colsol[js] == -1)
            stop = int(np.argmax(term)) if term.any() else len(ks)
            # columns before the stop (and the stopping column's predecessor) take the relaxation
            sel = np.nonzero(upd[:stop])[0]
            pred[js[sel]] = i
            d[js[sel]] = v2[sel]
            if stop < len(ks):
                pred[js[stop]] = i
                end = int(js[stop])
            for p in np.nonzero(tie[:stop])[0]:  # ties join the ready list, in list order
                kk = int(ks[p])
                j = int(js[p])
                cl[kk] = cl[up]
                cl[up] = j
                up += 1
        ready = cl[:last + 1]
        v[ready] = v[ready] + d[ready] - mn
        while True:
            i = int(pred[end])
            colsol[end] = i
            j1 = end
            end = int(rowsol[i])
            rowsol[i] = j1
            if i == fr:
                break
    return rowsol.astype(np.int32), colsol.astype(np.int32)


def solve_batch(costs):
    costs = np.asarray(costs, dtype=np.float64)
    rows = np.zeros(costs.shape[:2], np.int32)
    cols = np.zeros(costs.shape[:2], np.int32)
    for k, c in enumerate(costs):
        rows[k], cols[k] = solve(c)
    return rows, cols
