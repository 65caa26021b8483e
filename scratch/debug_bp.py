import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from rig import case_c1, sigmoid
from oracle import oracle as orc
from raynet_b200.engine import RayPotentialEngine
c = case_c1()
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
eng = RayPotentialEngine(c.M, c.D, c.V, 32, c.H, c.W, 11, c.bbox, c.grid, max_rays=c.N, use_distributed=False)
eng.set_voxel_grid(c.vgrid)
eng.add_image(d(c.ray_idxs), d(c.features), d(c.P), d(c.P_inv), d(c.centre))
eng.finalize_frontend()
print("class sizes", eng.class_sizes)
o = orc.frontend(c.ray_idxs, c.features, c.P, c.P_inv, c.centre, c.vgrid, c.bbox, c.grid, c.M, c.D, c.V, 32, c.H, c.W, 11)
prior = np.float32(np.log(0.05) - np.log(0.95))
acc_prev = np.full(tuple(c.grid), prior, np.float64)
msgs = np.zeros((c.N, c.M), np.float32)
for it in range(3):
    eng.bp_iteration()
    new = np.full(tuple(c.grid), prior, np.float64)
    orc.bp_iteration(o["S_vox"], o["idx"], o["cnt"], c.grid, acc_prev, new, msgs, acc_f64=True)
    acc_prev = new
    gm = eng.messages().cpu().numpy()
    ga = eng.accumulator().cpu().numpy()
    em = np.abs(sigmoid(gm) - sigmoid(msgs))
    ea = np.abs(sigmoid(ga) - sigmoid(new))
    r, i = np.unravel_index(em.argmax(), em.shape)
    print("sweep", it, "max msg err", em.max(), "at ray", r, "voxel", i, "count", o["cnt"][r], "acc err", ea.max(),
          "n bad rays", (em.max(1) > 1e-5).sum())
    if em.max() > 1e-5:
        print(" gpu", gm[r, max(0,i-3):i+4], "\n ref", msgs[r, max(0,i-3):i+4])
        bad = np.where(em.max(1) > 1e-5)[0]
        print(" bad ray counts", o["cnt"][bad][:20], "first bad voxel", [int(np.argmax(em[b] > 1e-5)) for b in bad[:20]])

print("==== detail")
eng.reset()
eng.add_image(d(c.ray_idxs), d(c.features), d(c.P), d(c.P_inv), d(c.centre))
eng.finalize_frontend()
eng.bp_iteration()
gm = eng.messages().cpu().numpy()
msgs = np.zeros((c.N, c.M), np.float32)
new = np.full(tuple(c.grid), prior, np.float64)
orc.bp_iteration(o["S_vox"], o["idx"], o["cnt"], c.grid, np.full(tuple(c.grid), prior, np.float64), new, msgs, acc_f64=True)
em = np.abs(sigmoid(gm) - sigmoid(msgs)).max(1)
bad = np.where(em > 1e-5)[0]
good = np.where(em <= 1e-5)[0]
cnt = o["cnt"]
print("bad L%4 hist", np.bincount(cnt[bad] % 4, minlength=4), "good L%4 hist", np.bincount(cnt[good] % 4, minlength=4))
print("bad L%32 ", sorted(set((cnt[bad] % 32).tolist()))[:40])
print("bad L    ", sorted(set(cnt[bad].tolist())))
print("good L   ", sorted(set(cnt[good].tolist())))
hdr = eng.hdr[:c.N].cpu().numpy()
sg = (hdr[:, 1] >> 16) & 7
print("bad signs hist", np.bincount(sg[bad], minlength=8), "good signs hist", np.bincount(sg[good], minlength=8))
order = eng.order[:c.N].cpu().numpy()
pos = np.empty(c.N, int); pos[order] = np.arange(c.N)
print("bad positions in order (mod 4):", np.bincount(pos[bad] % 4, minlength=4), " good:", np.bincount(pos[good] % 4, minlength=4))
b = bad[0]
print("ray", b, "L", cnt[b], "pos", pos[b])
print(" gpu msgs", gm[b, :cnt[b] + 4])
print(" ref msgs", msgs[b, :cnt[b] + 4])
