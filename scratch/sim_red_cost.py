"""Simulate the L1TEX cost model (measured by scratch/mb_l1.cu) of the accumulator gather / RED for
different lane <-> (ray, voxel) assignments on the C3 rig.  cost(LDG) = distinct 128-B lines,
cost(RED) = 1.5 * sum over sectors of the max multiplicity of one address inside the sector."""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from oracle import oracle as orc
from raynet_b200.synth import SyntheticScene, camera_arrays

G, H, W, M = 256, 512, 512, 768
scene = SyntheticScene(9, H, W, (G, G, G), neighbors=8)
bbox = np.array([-1, -1, -1, 1, 1, 1], np.float32); grid = np.array([G] * 3, np.int32)

def brick(x, y, z, mode):
    if mode == 'b442':   # current: 4x4x2 line of 2x2x2 sectors
        bl = G // 2; by = G // 4
        return ((x >> 2) * by + (y >> 2)) * bl * 32 + (z >> 1) * 32 + ((x >> 1) & 1) * 16 + (x & 1) * 4 + ((y >> 1) & 1) * 8 + (y & 1) * 2 + (z & 1)
    if mode == 'row':
        return (x * G + y) * G + z

def analyse(img, x0, y0, tw, th, label):
    order = scene.view_order(img)
    P, P_inv, centre = camera_arrays([scene.get_image(j) for j in order])
    xs, ys = np.meshgrid(np.arange(x0, x0 + tw), np.arange(y0, y0 + th), indexing='ij')
    ids = (xs * H + ys).astype(np.int32).ravel()          # x-major tile
    s, e = orc.sample_in_bbox(ids, H, P_inv, centre, bbox)
    idx, cnt = orc.voxel_traversal(bbox, grid, s, e, M)
    lin = brick(idx[..., 0], idx[..., 1], idx[..., 2], 'b442').astype(np.int64)
    n = len(ids)
    valid = np.arange(M)[None, :] < cnt[:, None]
    lin = np.where(valid, lin, -1)
    tot = int(cnt.sum())
    def cost(groups):
        # groups: list of arrays of lin values (<=32 each, -1 invalid)
        lines = 0; red = 0.0; ninstr = 0
        for g in groups:
            g = g[g >= 0]
            if len(g) == 0: continue
            ninstr += 1
            lines += len(np.unique(g >> 5))
            sec = g >> 3
            u, c = np.unique(g, return_counts=True)
            us = u >> 3
            # per sector max multiplicity
            order_ = np.argsort(us, kind='stable')
            m = {}
            for a, b in zip(us, c):
                m[a] = max(m.get(a, 0), b)
            red += 1.5 * sum(m.values())
        return lines, red, ninstr
    res = {}
    # A: warp per ray, 32 consecutive voxels
    groups = [lin[r, i:i + 32] for r in range(n) for i in range(0, int(cnt[r]), 32)]
    res['A ray x32'] = cost(groups)
    # B: 2 rays (adjacent in y) x 16 voxels
    lt = lin.reshape(tw, th, M)
    cn = cnt.reshape(tw, th)
    groups = []
    for a in range(tw):
        for b in range(0, th, 2):
            L = int(cn[a, b:b + 2].max())
            for i in range(0, L, 16):
                groups.append(lt[a, b:b + 2, i:i + 16].ravel())
    res['B 1x2 rays x16'] = cost(groups)
    groups = []
    for a in range(0, tw, 2):
        for b in range(0, th, 2):
            L = int(cn[a:a + 2, b:b + 2].max())
            for i in range(0, L, 8):
                groups.append(lt[a:a + 2, b:b + 2, i:i + 8].ravel())
    res['C 2x2 rays x8'] = cost(groups)
    groups = []
    for a in range(0, tw, 4):
        for b in range(0, th, 4):
            L = int(cn[a:a + 4, b:b + 4].max())
            for i in range(0, L, 2):
                groups.append(lt[a:a + 4, b:b + 4, i:i + 2].ravel())
    res['D 4x4 rays x2'] = cost(groups)
    groups = []
    for a in range(0, tw, 8):
        for b in range(0, th, 4):
            L = int(cn[a:a + 8, b:b + 4].max())
            for i in range(0, L, 1):
                groups.append(lt[a:a + 8, b:b + 4, i:i + 1].ravel())
    res['E 8x4 rays x1'] = cost(groups)
    # distinct voxels in the whole tile (what a CTA-level merge could reach)
    allv = lin[lin >= 0]
    print('%s: %d rays, %d visits, mean L %.0f, distinct voxels %d (%.2f visits each), distinct sectors %d' % (
        label, n, tot, tot / n, len(np.unique(allv)), tot / len(np.unique(allv)), len(np.unique(allv >> 3))))
    for k, (l, r, ni) in res.items():
        print('   %-16s LDG lines/32 visits %.2f   RED cycles/32 visits %.2f   lane utilisation %.2f' % (k, 32.0 * l / tot, 32.0 * r / tot, tot / (32.0 * ni)))

for img in (0, 1, 4):
    analyse(img, 256, 256, 16, 16, 'image %d centre' % img)
analyse(0, 64, 200, 16, 16, 'image 0 off-centre')
