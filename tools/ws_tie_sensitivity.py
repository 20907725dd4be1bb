"""How much does the one unpinned choice of the watershed restatement matter?

The oracle (and the GPU kernel) order heap entries by the strict total order (value, age, pixel index). The library the
reference calls (scikit-image 0.19, ``_watershed_cy.pyx`` + ``heap_general.pxi``; not installed here) compares only
(value, age) and leaves exact ties to the mechanics of its array binary heap: push = append + sift-up while strictly
smaller than the parent; pop = move the last entry to the root + sift-down towards the strictly smaller child (left child
tested first). Ages are a global push counter, so only the marker seeds (all age 0) can tie. This tool floods the same
(dist, marker, mask) three ways --

  total  : (value, age, index)                       -- what oracle/ and csrc/postproc.cu implement
  heap   : (value, age) with the heap mechanics above, as recalled from the library source (NOT authoritative)
  reverse: (value, age, -index)                      -- the opposite tie-break, a worst-case bracket

-- and counts the label pixels that differ. ``python tools/ws_tie_sensitivity.py [size] [nuclei] [seeds...]``"""
import pathlib
import sys
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from cellvit_b200 import synth  # noqa: E402
from oracle import postproc_oracle as po  # noqa: E402


class ArrayHeap:
    """Binary heap over (value, age, index) tuples comparing (value, age) only; structure as in heap_general.pxi."""

    def __init__(self):
        self.a = []

    @staticmethod
    def smaller(x, y):
        return x[0] < y[0] if x[0] != y[0] else x[1] < y[1]

    def push(self, item):
        a = self.a
        a.append(item)
        child = len(a) - 1
        while child > 0:
            parent = (child + 1) // 2 - 1
            if not self.smaller(a[child], a[parent]):
                break
            a[child], a[parent] = a[parent], a[child]
            child = parent

    def pop(self):
        a = self.a
        top = a[0]
        last = a.pop()
        n = len(a)
        if n:
            a[0] = last
            i = 0
            while True:
                l, r = 2 * i + 1, 2 * i + 2
                smallest = i
                if l < n:
                    if self.smaller(a[l], a[i]):
                        smallest = l
                    if r < n and self.smaller(a[r], a[smallest]):
                        smallest = r
                else:
                    break
                if smallest == i:
                    break
                a[i], a[smallest] = a[smallest], a[i]
                i = smallest
        return top

    def __len__(self):
        return len(self.a)


def flood(dist, marker, mask, mode):
    """Marker-controlled priority flood, connectivity 1 (neighbour order up, left, right, down), label at push time."""
    import heapq
    H, W = dist.shape
    Wp = W + 2
    img = np.zeros((H + 2, Wp), np.float64); img[1:-1, 1:-1] = dist
    out = np.zeros((H + 2, Wp), np.int32); out[1:-1, 1:-1] = marker * (mask != 0)
    msk = np.zeros((H + 2, Wp), np.uint8); msk[1:-1, 1:-1] = mask != 0
    img, out, msk = img.ravel().tolist(), out.ravel(), msk.ravel().tolist()
    outl = out.tolist()
    seeds = np.nonzero(out)[0].tolist()
    nbrs = (-Wp, -1, 1, Wp)
    age = 0
    if mode == "heap":
        h = ArrayHeap()
        for p in seeds:
            h.push((img[p], 0, p))
        while len(h):
            _, _, p = h.pop()
            lab = outl[p]
            for d in nbrs:
                n = p + d
                if not msk[n] or outl[n]:
                    continue
                age += 1
                outl[n] = lab
                h.push((img[n], age, n))
    else:
        sgn = 1 if mode == "total" else -1
        h = [(img[p], 0, sgn * p) for p in seeds]
        heapq.heapify(h)
        while h:
            _, _, sp = heapq.heappop(h)
            p = sgn * sp
            lab = outl[p]
            for d in nbrs:
                n = p + d
                if not msk[n] or outl[n]:
                    continue
                age += 1
                outl[n] = lab
                heapq.heappush(h, (img[n], age, sgn * n))
    return np.asarray(outl, np.int32).reshape(H + 2, Wp)[1:-1, 1:-1]


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    nuclei = int(sys.argv[2]) if len(sys.argv) > 2 else 700
    seeds = [int(s) for s in sys.argv[3:]] or [0, 1, 2, 3]
    for seed in seeds:
        t = synth.synthetic_nuclei(size, nuclei, seed=seed)
        labels, st = po.proc_np_hv(t["np_bin"], t["hv"], 40, want_intermediates=True)
        dist, marker, blb = st["dist"], st["marker"], st["blb"]
        sv = dist[marker > 0]
        _, counts = np.unique(sv, return_counts=True)
        tied_seeds = int(counts[counts > 1].sum())
        t0 = time.perf_counter()
        res = {m: flood(dist, marker, blb, m) for m in ("total", "heap", "reverse")}
        assert np.array_equal(res["total"], labels), "python flood (total order) != C oracle"
        n_mask = int((blb != 0).sum())
        print(f"seed {seed}: {size}x{size}, {len(np.unique(labels)) - 1} instances, {n_mask} mask px, {int((marker > 0).sum())} seeds "
              f"({tied_seeds} share their dist value with another seed); label pixels differing from the total order: "
              f"heap-mechanics {int((res['heap'] != labels).sum())}, reverse-index {int((res['reverse'] != labels).sum())} "
              f"[{time.perf_counter() - t0:.1f} s]")


if __name__ == "__main__":
    main()
