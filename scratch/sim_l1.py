"""L1 hit-rate model of plane-sweep feature gathers under different ray/plane/view schedules (no GPU)."""
import sys, numpy as np
from collections import OrderedDict
sys.path.insert(0, '/root/repo')
from raynet_b200.synth import ring_cameras

H = W = 512; D = 64; V = 9; PAD = 11
cams = ring_cameras(V, H, W)
bbox = np.array([-1, -1, -1, 1, 1, 1], np.float64)

def rays_of(ref, xs, ys):
    cam = cams[ref]
    Pinv = cam.P_pinv; C = cam.center[:3, 0].astype(np.float64)
    px = np.stack([xs, ys, np.ones_like(xs)], 0).astype(np.float64)
    X = Pinv @ px
    dirs = (X[:3] / X[3]).T - C
    t1 = (bbox[:3] - C) / dirs; t2 = (bbox[3:] - C) / dirs
    tn = np.minimum(t1, t2).max(1); tf = np.maximum(t1, t2).min(1)
    return C + tn[:, None] * dirs, C + tf[:, None] * dirs

def pixels(ref, xs, ys):
    """-> int64 key [n, D, V-1] unique per (view, fy, fx)"""
    rs, re = rays_of(ref, xs, ys)
    k = np.arange(D) / (D - 1)
    pts = rs[:, None, :] + k[None, :, None] * (re - rs)[:, None, :]          # n, D, 3
    keys = []
    for j in range(1, V):
        v = (ref + j) % V
        P = cams[v].P
        q = pts @ P[:, :3].T + P[:, 3]
        fx = np.clip(np.rint(q[..., 0] / q[..., 2]).astype(np.int64) + PAD // 2 + 1, 0, W)   # shift approx
        fy = np.clip(np.rint(q[..., 1] / q[..., 2]).astype(np.int64) + PAD // 2 + 1, 0, H)
        keys.append((v * 1024 + fy) * 1024 + fx)
    return np.stack(keys, -1)

class LRU:
    def __init__(s, lines): s.d = OrderedDict(); s.cap = lines; s.hit = s.tot = 0
    def access(s, keys):
        for k in keys:
            s.tot += 1
            if k in s.d: s.d.move_to_end(k); s.hit += 1
            else:
                s.d[k] = 1
                if len(s.d) > s.cap: s.d.popitem(last=False)

def tiled_xy(t):
    st, rem = t >> 12, t & 4095
    sty, stx = st % (H >> 6), st // (H >> 6)
    tile, inn = rem >> 6, rem & 63
    return stx * 64 + (tile >> 3) * 8 + (inn >> 3), sty * 64 + (tile & 7) * 8 + (inn & 7)

def sim_current(ref, sm=37, n_waves=6, lines=1400):
    """8 CTAs x 4 warps per SM; CTA b -> SM b % 148 (first-order model); warps round-robin per plane group."""
    lru = LRU(lines)
    for wave in range(n_waves):
        rays = []
        for c in range(8):
            b = sm + 148 * (wave * 8 + c) + 30000
            rays += [4 * b + w for w in range(4)]
        t = np.array(rays); xs, ys = tiled_xy(t)
        key = pixels(ref, xs, ys)            # 32, D, 8
        for k0 in range(0, D, 4):
            for w in range(32):
                for v in range(V - 1):
                    lru.access(key[w, k0:k0 + 4, v].tolist())
    return lru.hit / lru.tot

def sim_tile(ref, TX, TY, KB, ctas=1, lines=1400, x0=200, y0=200, n_tiles=3, view_outer=True):
    """CTA = TX x TY ray tile; plane blocks of KB planes; per block: view-outer, plane-inner, all rays of the tile per step."""
    lru = LRU(lines)
    for it in range(n_tiles):
        keys = []
        for c in range(ctas):
            xs, ys = np.meshgrid(np.arange(TX) + x0 + 97 * c + TX * it, np.arange(TY) + y0 + 53 * c, indexing='ij')
            keys.append(pixels(ref, xs.ravel(), ys.ravel()))
        for kb in range(0, D, KB):
            if view_outer:
                for v in range(V - 1):
                    for k in range(kb, kb + KB):
                        for c in range(ctas):
                            lru.access(keys[c][:, k, v].tolist())
            else:
                for k in range(kb, kb + KB):
                    for v in range(V - 1):
                        for c in range(ctas):
                            lru.access(keys[c][:, k, v].tolist())
    return lru.hit / lru.tot

if __name__ == '__main__':
    # epipolar step per plane
    xs, ys = np.meshgrid(np.arange(200, 208), np.arange(200, 208), indexing='ij')
    key = pixels(0, xs.ravel(), ys.ravel())
    fx, fy = key % 1024, (key // 1024) % 1024
    step = np.hypot(np.diff(fx, axis=1), np.diff(fy, axis=1))
    print('epipolar px per plane, per view:', step.mean(axis=(0, 1)).round(2))
    print('current schedule hit:', round(sim_current(0), 3))
    for (TX, TY, KB, ctas) in [(8, 8, 16, 1), (8, 8, 16, 2), (16, 8, 16, 1), (16, 16, 16, 1), (16, 16, 8, 1), (8, 8, 64, 1), (16, 8, 8, 2), (16, 16, 64, 1), (32, 8, 16, 1), (8, 4, 16, 4)]:
        print('tile', TX, TY, 'KB', KB, 'ctas', ctas, 'view-outer hit', round(sim_tile(0, TX, TY, KB, ctas), 3),
              'plane-outer hit', round(sim_tile(0, TX, TY, KB, ctas, view_outer=False), 3))

def sim_warp_per_ray(ref, TX, TY, ctas, lines=1400, n_rounds=3, x0=200, y0=200, drift=0):
    """warp per ray as today (4 planes x all views per step), but a CTA = TX x TY compact tile of rays, `ctas` CTAs per SM
    from distant tiles, all warps in approximate lockstep (drift: warp w lags by (w % (drift+1)) plane groups)."""
    lru = LRU(lines)
    for it in range(n_rounds):
        keys = []
        for c in range(ctas):
            xs, ys = np.meshgrid(np.arange(TX) + x0 + 97 * c + TX * it, np.arange(TY) + y0 + 53 * c, indexing='ij')
            keys.append(pixels(ref, xs.ravel(), ys.ravel()))
        nw = TX * TY
        for step in range(D // 4 + drift):
            for c in range(ctas):
                for w in range(nw):
                    k0 = 4 * (step - (w % (drift + 1)))
                    if 0 <= k0 < D:
                        for v in range(V - 1):
                            lru.access(keys[c][w, k0:k0 + 4, v].tolist())
    return lru.hit / lru.tot

if __name__ == '__main__':
    for (TX, TY, ctas, drift) in [(1, 4, 8, 0), (2, 4, 4, 0), (4, 4, 2, 0), (4, 8, 1, 0), (8, 4, 1, 0), (4, 8, 1, 2), (4, 8, 1, 4), (4, 4, 2, 2), (8, 8, 1, 0), (8, 8, 1, 3)]:
        print('warp/ray CTA', TX, 'x', TY, 'ctas/SM', ctas, 'drift', drift, 'hit', round(sim_warp_per_ray(0, TX, TY, ctas, drift=drift), 3))
    for lines in (800, 1400, 1800):
        print('lines', lines, 'CTA 4x8', round(sim_warp_per_ray(0, 4, 8, 1, lines=lines), 3))
