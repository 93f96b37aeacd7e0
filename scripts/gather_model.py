"""CPU model of the force kernel's gather pattern (design evidence, not product code).

For a disordered FCC configuration at density 0.7 it builds the full neighbour list (r <= 3.3) and counts, for
several thread/row organisations, (a) the distinct 128-byte lines a warp touches per gather instruction
(positions stored as 32-byte records) and (b) the padding (idle lane-steps) the organisation costs.
"""
import sys

import numpy as np
from scipy.spatial import cKDTree


def fcc(L, density=0.7, sigma=0.12, seed=0):
    s = 1.0 / (density * 0.25) ** (1.0 / 3.0)
    n = int(L / s)
    s = L / n
    hs = 0.5 * s
    g = np.arange(n) * s
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    base = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    offs = np.array([[0, 0, 0], [0, hs, hs], [hs, 0, hs], [hs, hs, 0]])
    q = (base[:, None, :] + offs[None, :, :]).reshape(-1, 3)
    rng = np.random.default_rng(seed)
    q = q + rng.normal(0, sigma, q.shape)
    return np.mod(q, L)


def order_cells(q, L, edge_target, sub=1, mode="x"):
    """slot order: coarse cells x-fastest; inside a cell optionally by sub-cell (sub^3, x-fastest or morton)."""
    ms = int(L / edge_target)
    edge = L / ms
    c = np.minimum((q / edge).astype(int), ms - 1)
    cell = c[:, 0] + ms * (c[:, 1] + ms * c[:, 2])
    key = cell.astype(np.int64)
    if sub > 1:
        f = np.minimum(((q / edge - c) * sub).astype(int), sub - 1)
        if mode == "morton" and sub == 2:
            sk = f[:, 0] + 2 * f[:, 1] + 4 * f[:, 2]
        else:
            sk = f[:, 0] + sub * (f[:, 1] + sub * f[:, 2])
        key = key * (sub ** 3) + sk
    order = np.argsort(key, kind="stable")
    return order, cell[order], ms


def main():
    L = float(sys.argv[1]) if len(sys.argv) > 1 else 33.6
    SL = 3.3
    q = fcc(L)
    n = len(q)
    print("N", n)
    for sub, mode in [(1, "x"), (2, "morton"), (3, "x")]:
        order, cell, ms = order_cells(q, L, SL, sub, mode)
        qs = q[order]
        tree = cKDTree(qs, boxsize=L)
        nb = tree.query_ball_point(qs, SL, return_sorted=True)
        rows = [np.array([j for j in r if j != i], dtype=np.int64) for i, r in enumerate(nb)]
        cnt = np.array([len(r) for r in rows])
        print("sub=%d %s: mean row %.1f max %d" % (sub, mode, cnt.mean(), cnt.max()))
        # periodic wrap makes the low-index cells see high-index slots; the "line" statistics are unaffected
        REC = 4  # 32-byte records per 128-byte line
        # --- S0: lane = particle, 32 consecutive slots per warp, rows ascending
        nw = n // 32
        tot_lines = 0
        tot_steps = 0
        used = 0
        for w in range(0, nw, 7):
            rr = rows[32 * w:32 * w + 32]
            ml = max(len(r) for r in rr)
            used += sum(len(r) for r in rr)
            for k in range(ml):
                js = [r[k] for r in rr if k < len(r)]
                tot_lines += len(set(j // REC for j in js))
            tot_steps += ml
        print("  S0 thread/particle : lines/gather %.1f  lane efficiency %.2f" % (tot_lines / tot_steps, used / (32.0 * tot_steps)))
        # --- S4: 4 lanes per particle (8 particles per warp), lane c takes entries 4k+c
        for G in (2, 4, 8):
            PW = 32 // G
            nw4 = n // PW
            tot_lines = 0
            tot_steps = 0
            used = 0
            for w in range(0, nw4, 29):
                rr = rows[PW * w:PW * w + PW]
                ml = max((len(r) + G - 1) // G for r in rr)
                used += sum(len(r) for r in rr)
                for k in range(ml):
                    js = []
                    for r in rr:
                        js.extend(r[G * k:G * k + G])
                    tot_lines += len(set(j // REC for j in js))
                tot_steps += ml
            print("  S%d %d lanes/particle : lines/gather %.1f  lane efficiency %.2f" % (G, G, tot_lines / tot_steps, used / (32.0 * tot_steps)))
        # --- S2: lane = particle, rows segmented by partner CELL, each segment padded to the warp maximum
        tot_lines = 0
        tot_steps = 0
        used = 0
        for w in range(0, nw, 7):
            rr = rows[32 * w:32 * w + 32]
            used += sum(len(r) for r in rr)
            cells = sorted(set(int(c) for r in rr for c in cell[r]))
            for cc in cells:
                seg = [r[cell[r] == cc] for r in rr]
                ml = max(len(s) for s in seg)
                for k in range(ml):
                    js = [s[k] for s in seg if k < len(s)]
                    tot_lines += len(set(j // REC for j in js))
                tot_steps += ml
        print("  S2 cell-aligned    : lines/gather %.1f  lane efficiency %.2f" % (tot_lines / tot_steps, used / (32.0 * tot_steps)))


if __name__ == "__main__":
    main()
