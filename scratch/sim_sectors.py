"""How many distinct 32-byte accumulator sectors (2x2x2 voxel bricks) one gather / RED warp instruction touches, for
warp = 32 consecutive voxels of ONE ray (today) against 16 + 16 voxels of two neighbouring rays or 8 x 4 of four."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import oracle
from raynet_b200.synth import ring_cameras
H = W = 512; G = 256; M = 768
cams = ring_cameras(9, H, W)
bbox = np.array([-1, -1, -1, 1, 1, 1], np.float32)
grid = np.array([G, G, G], np.int32)
rng = np.random.RandomState(0)
res = {k: [0, 0] for k in ('1 ray x 32', '2 rays (y, y+1) x 16', '2 rays (x, x+1) x 16', '4 rays (2x2) x 8', '4 rays (1x4) x 8')}
for ref in (0, 3):
    cam = cams[ref]
    P_inv = cam.P_pinv.astype(np.float32); C = cam.center.astype(np.float32)
    for _ in range(60):
        x0, y0 = rng.randint(40, W - 40), rng.randint(40, H - 40)
        pix = [(x0 + dx, y0 + dy) for dx in range(2) for dy in range(4)]
        ids = np.array([x * H + y for (x, y) in pix], np.int32)
        starts, ends = oracle.sample_in_bbox(ids, H, P_inv, C, bbox)
        idx, cnt = oracle.voxel_traversal(bbox, grid, starts, ends, M)
        sect = [((idx[r, :cnt[r], 0] >> 1) * 128 + (idx[r, :cnt[r], 1] >> 1)) * 128 + (idx[r, :cnt[r], 2] >> 1) for r in range(len(pix))]
        def rays(*which): return [sect[pix.index(w)] for w in which]
        def count(group, per):   # instruction k covers voxels [k*per, (k+1)*per) of every ray of the group
            L = max(len(s) for s in group); n = tot = 0
            for k in range(0, L, per):
                u = set()
                lanes = 0
                for s in group:
                    u.update(s[k:k + per].tolist()); lanes += len(s[k:k + per])
                n += len(u); tot += lanes
            return n, tot
        for name, group, per in (('1 ray x 32', rays((x0, y0)), 32), ('2 rays (y, y+1) x 16', rays((x0, y0), (x0, y0 + 1)), 16),
                                 ('2 rays (x, x+1) x 16', rays((x0, y0), (x0 + 1, y0)), 16),
                                 ('4 rays (2x2) x 8', rays((x0, y0), (x0, y0 + 1), (x0 + 1, y0), (x0 + 1, y0 + 1)), 8),
                                 ('4 rays (1x4) x 8', rays((x0, y0), (x0, y0 + 1), (x0, y0 + 2), (x0, y0 + 3)), 8)):
            n, tot = count(group, per)
            res[name][0] += n; res[name][1] += tot
for k, (n, tot) in res.items():
    print('%-24s sectors per voxel %.3f  (per 32-lane instruction %.1f)' % (k, n / tot, 32 * n / tot))

# ---- lane layouts under the hypothesis that REDs coalesce only within 8-lane quarters ---------------------------------
print()
res2 = {}
rng = np.random.RandomState(1)
for ref in (0, 3):
    cam = cams[ref]
    P_inv = cam.P_pinv.astype(np.float32); C = cam.center.astype(np.float32)
    for _ in range(60):
        x0, y0 = rng.randint(40, W - 40), rng.randint(40, H - 40)
        ids = np.array([x0 * H + y0 + d for d in range(4)], np.int32)
        starts, ends = oracle.sample_in_bbox(ids, H, P_inv, C, bbox)
        idx, cnt = oracle.voxel_traversal(bbox, grid, starts, ends, M)
        sect = [((idx[r, :cnt[r], 0] >> 1) * 128 + (idx[r, :cnt[r], 1] >> 1)) * 128 + (idx[r, :cnt[r], 2] >> 1) for r in range(4)]
        L = min(cnt)
        L -= L % 32
        layouts = {
            'one ray, lanes = 32 consecutive slots': lambda k, lane: (0, k * 32 + lane),
            '4 rays, ray = lane >> 3, slot = lane & 7': lambda k, lane: (lane >> 3, k * 8 + (lane & 7)),
            '4 rays, ray = lane & 3, slot = lane >> 2': lambda k, lane: (lane & 3, k * 8 + (lane >> 2)),
            '4 rays, quarter = 4 rays x 2 slots (ray = (lane >> 1) & 3)': lambda k, lane: ((lane >> 1) & 3, k * 8 + (lane >> 3) * 2 + (lane & 1)),
        }
        for name, f in layouts.items():
            per = 32 if name.startswith('one') else 8
            full = quarter = n = 0
            for k in range(L // per):
                a = [sect[f(k, lane)[0]][f(k, lane)[1]] for lane in range(32)]
                full += len(set(a)); quarter += sum(len(set(a[q * 8:q * 8 + 8])) for q in range(4)); n += 1
            r = res2.setdefault(name, [0, 0, 0]); r[0] += full; r[1] += quarter; r[2] += n
for name, (full, quarter, n) in res2.items():
    print('%-62s sectors / instruction: whole warp %.1f, summed over 8-lane quarters %.1f' % (name, full / n, quarter / n))
