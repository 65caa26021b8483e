import sys
sys.path.insert(0, '/root/repo/scratch')
from sim_l1b import *
for (TX, TY, ctas) in [(8, 2, 2), (16, 1, 2), (2, 8, 2), (4, 4, 2), (8, 4, 1), (16, 2, 1), (32, 1, 1), (8, 1, 4), (4, 2, 4), (8,3,1), (8, 6, 1)]:
    print('warp/ray CTA', TX, TY, 'x', ctas, run(warp_per_ray, TX=TX, TY=TY, ctas=ctas), 'drift2', run(warp_per_ray, TX=TX, TY=TY, ctas=ctas, drift=2), flush=True)
