import sys, numpy as np
sys.path.insert(0, '/root/repo/scratch')
from sim_l1 import *

LOCS = [(40, 60), (130, 300), (256, 256), (200, 200), (330, 120), (420, 440), (256, 40), (470, 250)]

def run(fn, **kw):
    h = t = 0
    for (x0, y0) in LOCS:
        lru = fn(x0=x0, y0=y0, **kw)
        h += lru.hit; t += lru.tot
    return round(h / t, 3)

def warp_per_ray(TX, TY, ctas, x0, y0, lines=1400, n_rounds=2, drift=0, ref=0, far=True):
    lru = LRU(lines)
    for it in range(n_rounds):
        keys = []
        for c in range(ctas):
            cx = (x0 + (97 * c if far else TX * c) + TX * it * (1 if far else ctas)) % (W - TX)
            cy = (y0 + (53 * c if far else 0)) % (H - TY)
            xs, ys = np.meshgrid(np.arange(TX) + cx, np.arange(TY) + cy, indexing='ij')
            keys.append(pixels(ref, xs.ravel(), ys.ravel()))
        nw = TX * TY
        for step in range(D // 4 + drift):
            for c in range(ctas):
                for w in range(nw):
                    k0 = 4 * (step - (w % (drift + 1)))
                    if 0 <= k0 < D:
                        for v in range(V - 1):
                            lru.access(keys[c][w, k0:k0 + 4, v].tolist())
    return lru

def tile_sched(TX, TY, KB, x0, y0, ctas=1, lines=1400, n_tiles=2, view_outer=True, ref=0):
    lru = LRU(lines)
    for it in range(n_tiles):
        keys = []
        for c in range(ctas):
            cx = (x0 + 97 * c + TX * it) % (W - TX); cy = (y0 + 53 * c) % (H - TY)
            xs, ys = np.meshgrid(np.arange(TX) + cx, np.arange(TY) + cy, indexing='ij')
            keys.append(pixels(ref, xs.ravel(), ys.ravel()))
        for kb in range(0, D, KB):
            if view_outer:
                for v in range(V - 1):
                    for k in range(kb, kb + KB):
                        for c in range(ctas): lru.access(keys[c][:, k, v].tolist())
            else:
                for k in range(kb, kb + KB):
                    for v in range(V - 1):
                        for c in range(ctas): lru.access(keys[c][:, k, v].tolist())
    return lru

if __name__ == '__main__':
    out = 0; tot = 0
    for (x0, y0) in LOCS:
        xs, ys = np.meshgrid(np.arange(8) + x0, np.arange(8) + y0, indexing='ij')
        key = pixels(0, xs.ravel(), ys.ravel())
        fx, fy = key % 1024, (key // 1024) % 1024
        o = ((fx == 0) | (fy == 0)).mean(); print((x0, y0), 'outside fraction', round(o, 3))
    print('today (1x4 x 8 far CTAs, lockstep)', run(warp_per_ray, TX=1, TY=4, ctas=8))
    print('today, drift 3', run(warp_per_ray, TX=1, TY=4, ctas=8, drift=3))
    for (TX, TY, ctas) in [(4, 4, 2), (4, 8, 1), (8, 4, 1), (8, 8, 1)]:
        print('warp/ray CTA', TX, TY, 'x', ctas, run(warp_per_ray, TX=TX, TY=TY, ctas=ctas), 'drift2', run(warp_per_ray, TX=TX, TY=TY, ctas=ctas, drift=2))
    for (TX, TY, KB, ctas) in [(8, 8, 16, 1), (16, 8, 16, 1), (16, 16, 16, 1), (16, 16, 64, 1), (32, 16, 16, 1), (32, 32, 16, 1)]:
        print('tile', TX, TY, 'KB', KB, 'ctas', ctas, 'view-outer', run(tile_sched, TX=TX, TY=TY, KB=KB, ctas=ctas),
              'plane-outer', run(tile_sched, TX=TX, TY=TY, KB=KB, ctas=ctas, view_outer=False))
