"""CPU model of the force kernel's L1 behaviour (design evidence for round 2, not product code).

Question: the kernel is bound by the L1 data pipe and ~35 % of its wavefronts are fills (L1 hit rate 63 %,
profiles/r01c_force_ncu.txt).  How much of that is the ORDER in which blocks reach an SM?  The model builds a
thermalised-looking FCC configuration, cell-sorts it like cells.cu, forms the inner-list rows (r <= cutoff + delta),
and replays the gather stream of one SM through an LRU cache of 32-byte sectors:
  * 4 warps per block, 8 locals per warp, `resident` blocks per SM interleaved step by step;
  * "round-robin": SM s runs blocks s, s + nsm, s + 2 nsm, ...  (what a plain grid does in steady state);
  * "striped":     SM s runs a contiguous range of blocks     (block-index permutation or per-SM queues).
List words are streamed through the same cache (8 sectors per warp step pair) as pollution.

    python scripts/l1_model.py [L=46] [nsm=12] [resident=10] [l1_kb=200]
"""
import sys
from collections import OrderedDict

import numpy as np
from scipy.spatial import cKDTree


def config(L, density=0.7, sigma=0.12, seed=0):
    s = 1.0 / (density * 0.25) ** (1.0 / 3.0)
    n = int(L / s)
    s = L / n
    hs = 0.5 * s
    g = np.arange(n) * s
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    base = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    offs = np.array([[0, 0, 0], [0, hs, hs], [hs, 0, hs], [hs, hs, 0]])
    q = (base[:, None, :] + offs[None, :, :]).reshape(-1, 3)
    q = q + np.random.default_rng(seed).normal(0, sigma, q.shape)
    return np.mod(q, L)


def cell_order(q, L, SL, curve="x"):
    ms = int(L / SL)
    edge = L / ms
    c = np.minimum((q / edge).astype(int), ms - 1)
    if curve == "x":
        key = c[:, 0] + ms * (c[:, 1] + ms * c[:, 2])
    else:  # morton over cells
        key = np.zeros(len(q), dtype=np.int64)
        for b in range(8):
            for d in range(3):
                key |= ((c[:, d] >> b) & 1).astype(np.int64) << (3 * b + d)
    return np.argsort(key, kind="stable")


class LRU:
    """LRU over 128-byte LINES (tags); a line holds 4 sector-valid bits, a sector miss in a present line fills
    that sector only -- the sectored organisation of the L1 (a line of 4 records of which one is wanted still
    costs a whole tag)."""

    def __init__(self, cap_sectors):
        self.cap, self.d, self.hit, self.miss = max(cap_sectors // 4, 1), OrderedDict(), 0, 0

    def touch(self, k, count=True):
        d = self.d
        if isinstance(k, tuple):  # list words: whole lines, streamed
            line, bit = k, 15
        else:
            line, bit = k >> 2, 1 << (k & 3)
        v = d.get(line)
        if v is not None:
            d.move_to_end(line)
            if v & bit:
                self.hit += count
            else:
                d[line] = v | bit
                self.miss += count
        else:
            d[line] = bit
            self.miss += count
            if len(d) > self.cap:
                d.popitem(last=False)


def replay(rows, blocks_of_sm, resident, cap_sectors, list_pollution=True):
    """blocks_of_sm: block ids in the order the SM receives them; block = 4 warps x 8 locals"""
    cache = LRU(cap_sectors)
    pending = list(blocks_of_sm)
    active = []  # [warp rows (8 arrays), step, nsteps, stream id]
    stream_id = 0

    def start(b):
        nonlocal stream_id
        out = []
        for w in range(4):
            rr = rows[32 * b + 8 * w: 32 * b + 8 * w + 8]
            if not rr:
                continue
            ns = max((len(r) + 3) // 4 for r in rr)
            stream_id += 1
            out.append([rr, 0, ns, stream_id])
        return out

    slots = []
    while pending and len(slots) < resident:
        slots.append(start(pending.pop(0)))
    while slots:
        for bi in range(len(slots)):
            blk = slots[bi]
            for wp in blk:
                rr, k, ns, sid = wp
                if k >= ns:
                    continue
                for r in rr:
                    for j in r[4 * k:4 * k + 4]:
                        cache.touch(int(j))
                if list_pollution and k % 2 == 0:
                    for x in range(2):  # 256 B of list words per step pair (two lines), never reused
                        cache.touch(("L", sid, k, x), count=False)
                wp[1] = k + 1
            if all(wp[1] >= wp[2] for wp in blk):
                slots[bi] = start(pending.pop(0)) if pending else None
        slots = [s for s in slots if s]
    return cache.hit / max(cache.hit + cache.miss, 1)


def main():
    L = float(sys.argv[1]) if len(sys.argv) > 1 else 46.0
    nsm = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    resident = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    l1_kb = float(sys.argv[4]) if len(sys.argv) > 4 else 200.0
    SL, keep = 3.3, 3.1
    q = config(L)
    n = len(q)
    cap = int(l1_kb * 1024 / 32)
    print("N %d, %d SMs x %d resident blocks, L1 %.0f KB (%d sectors)" % (n, nsm, resident, l1_kb, cap))
    for curve in ("x", "morton"):
        order = cell_order(q, L, SL, curve)
        qs = q[order]
        tree = cKDTree(qs, boxsize=L)
        nb = tree.query_ball_point(qs, keep, return_sorted=True)
        rows = [np.array([j for j in r if j != i], dtype=np.int64) for i, r in enumerate(nb)]
        nblocks = n // 32
        per = nblocks // nsm
        for name, blocks in (("round-robin", list(range(0, per * nsm, nsm))),
                             ("striped", list(range(per * (nsm // 2), per * (nsm // 2) + per)))):
            hr = replay(rows, blocks, resident, cap)
            print("  cells %-6s blocks %-11s : L1 sector hit rate %.3f (mean row %.1f)" % (
                curve, name, hr, np.mean([len(r) for r in rows])))


if __name__ == "__main__":
    main()
