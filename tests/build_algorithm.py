"""A sequential numpy model of the DEVICE construction algorithm (simplepath_b200/csrc/build_kernels.cu) — not of the
reference: level-synchronous nodes, bounds as "last position among the equal minima" keys with the zero-interval rule,
Hoare's partition as a closed-form permutation from one prefix sum, pre-order numbering from chain lengths.  It lets the CPU
suite check the algorithm's claims against the oracle's literal restatement of BVHAccelerator::construct
(tests/test_build_algorithm.py); the CUDA kernels themselves are checked on the GPU (tests/test_gpu_build.py)."""
from __future__ import annotations

import numpy as np

from simplepath_b200.capi import NODE_DTYPE

INVALID = 0xFFFFFFFF
F32 = np.float32


def ordered(f) -> int:
    u = int(F32(f).view(np.uint32))
    if (u << 1) & 0xFFFFFFFF == 0:
        u = 0
    return (~u) & 0xFFFFFFFF if u & 0x80000000 else u | 0x80000000


def unordered(k: int, negative_zero: int):
    u = (k & 0x7FFFFFFF) if k & 0x80000000 else (~k) & 0xFFFFFFFF
    if u == 0 and negative_zero:
        u = 0x80000000
    return np.uint32(u).view(np.float32)


def sign(f) -> int:
    return int(F32(f).view(np.uint32)) >> 31


def build(bounds: np.ndarray, non_triangle, first_id: int):
    """Returns (nodes, order, n_internal, max_depth) as spcu_build_bvh would."""
    n = len(bounds)
    perm = np.arange(n)
    seg = np.zeros(n, dtype=np.int64)
    first, last, parent, left_run, state, child0, box = [0], [n], [INVALID], [0], [0], [0], [None]
    begin, end, max_depth, level = 0, 1, 0, 0
    while begin < end and n:
        keys = {}
        for pos in range(n):  # k_bounds: 64-bit keys, atomicMin / atomicMax
            s = seg[pos]
            if s < 0:
                continue
            x = bounds[perm[pos]]
            k = keys.setdefault(s, [2 ** 64 - 1] * 3 + [0] * 3 + [INVALID] * 3)
            for a in range(3):
                k[a] = min(k[a], (ordered(x[a]) << 32) | ((0x7FFFFFFF - pos) << 1) | sign(x[a]))
                k[3 + a] = max(k[3 + a], (ordered(x[3 + a]) << 32) | (pos << 1) | sign(x[3 + a]))
                if not (x[a] == 0 and x[3 + a] == 0):
                    k[6 + a] = min(k[6 + a], pos)
        dim, at = {}, {}
        for node in range(begin, end):  # k_decide
            k = keys[node]
            lo, hi = [0] * 3, [0] * 3
            for a in range(3):
                l_, h_ = unordered(k[a] >> 32, k[a] & 1), unordered(k[3 + a] >> 32, k[3 + a] & 1)
                if l_ == 0:
                    j = 0x7FFFFFFF - ((k[a] & 0xFFFFFFFF) >> 1)
                    if j < k[6 + a]:
                        l_ = bounds[perm[j]][3 + a]
                lo[a], hi[a] = l_, h_
            box[node] = np.array(lo + hi, dtype=F32)
            if last[node] - first[node] <= 4:
                state[node] = 1
                continue
            size = [abs(F32(hi[a]) - F32(lo[a])) for a in range(3)]
            d = (0 if size[0] > size[2] else 2) if size[0] > size[1] else (1 if size[1] > size[2] else 2)
            dim[node], at[node], state[node] = d, (F32(lo[d]) + F32(hi[d])) / F32(2), 3
        flag = np.zeros(n + 1, dtype=np.int64)
        for pos in range(n):  # k_flags
            s = seg[pos]
            if s >= 0 and state[s] == 3:
                x, d = bounds[perm[pos]], dim[s]
                flag[pos] = (F32(x[d]) + F32(x[3 + d])) / F32(2) < at[s]
        prefix = np.concatenate([[0], np.cumsum(flag[:n])])
        mid = {}
        for node in range(begin, end):  # k_split
            if state[node] != 3:
                continue
            trues = prefix[last[node]] - prefix[first[node]]
            if trues == 0 or trues == last[node] - first[node]:
                state[node] = 1
                continue
            m, c0 = first[node] + trues, len(first)
            state[node], child0[node], mid[node] = 2, c0, m
            first += [first[node], m]
            last += [m, last[node]]
            parent += [node << 1, (node << 1) | 1]
            left_run += [left_run[node] + 1, 0]
            state += [0, 0]
            child0 += [0, 0]
            box += [None, None]
        left_misplaced, right_misplaced = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
        for pos in range(n):  # k_scatter
            s = seg[pos]
            if s < 0 or state[s] != 2:
                continue
            f0, m = first[s], mid[s]
            before = prefix[pos] - prefix[f0]
            if pos < m and not flag[pos]:
                left_misplaced[f0 + (pos - f0 - before)] = pos
            elif pos >= m and flag[pos]:
                right_misplaced[f0 + (m - f0 - before - 1)] = pos
        new_perm = perm.copy()
        for pos in range(n):  # k_bounds<true>: the partition as a gather
            s = seg[pos]
            if s < 0:
                continue
            if state[s] != 2:
                seg[pos] = -1
                continue
            f0, m = first[s], mid[s]
            before = prefix[pos] - prefix[f0]
            src = pos
            if pos < m and not flag[pos]:
                src = right_misplaced[f0 + (pos - f0 - before)]
            elif pos >= m and flag[pos]:
                src = left_misplaced[f0 + (m - f0 - before - 1)]
            new_perm[pos] = perm[src]
            seg[pos] = child0[s] + (1 if pos >= m else 0)
        perm = new_perm
        if len(first) > end:
            max_depth = level + 1
        begin, end, level = end, len(first), level + 1
    count = np.zeros(n + 1, dtype=np.int64)
    for v in range(len(first)):  # k_chain_counts
        if state[v] == 2 and state[child0[v]] != 2:
            count[first[v]] = left_run[v] + 1
    pre = np.concatenate([[0], np.cumsum(count[:n])])
    n_internal = int(pre[n]) if n else 0
    nodes = np.zeros(n_internal, dtype=NODE_DTYPE)
    for v in range(1, len(first)):  # k_emit
        p, w = parent[v] >> 1, parent[v] & 1
        dst = pre[first[p]] + left_run[p]
        nodes[dst]["box"][6 * w:6 * w + 6] = box[v]
        if state[v] == 2:
            nodes[dst]["child"][w], nodes[dst]["count"][w] = pre[first[v]] + left_run[v], 0
        else:
            mixed = non_triangle is not None and any(non_triangle[perm[q]] for q in range(first[v], last[v]))
            nodes[dst]["child"][w] = ~(first_id + first[v])
            nodes[dst]["count"][w] = (last[v] - first[v]) | (0x80000000 if mixed else 0)
    return nodes, perm, n_internal, max_depth
