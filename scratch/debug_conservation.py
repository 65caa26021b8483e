import sys, os, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_gpu_fullsize import FullRun, PRIOR
run = FullRun(torch, sys.argv[1] if len(sys.argv) > 1 else 'c2')
eng = run.eng
n = eng.n_rays
valid = (torch.arange(eng.R, device='cuda')[None, :] < eng.count[:n, None]) & (eng.count[:n] > 1)[:, None]
lin = eng.lin[:n][valid].long()
for sweep in range(3):
    eng.bp_iteration()
    m = eng.msgs[:n][valid].double()
    ref = torch.full((eng.GB,), PRIOR, dtype=torch.float64, device='cuda')
    ref.index_add_(0, lin, m)
    got = eng.acc_prev.double()
    d = (got - ref)
    print('sweep', sweep, 'max|d|', float(d.abs().max()), 'sum d', float(d.sum()), 'sum|m|', float(m.abs().sum()), 'max|acc|', float(got.abs().max()),
          'n big', int((d.abs() > 1e-3).sum()))
    big = torch.nonzero(d.abs() > 1e-3).flatten()[:10]
    for b in big.tolist():
        print('   off', b, 'got', float(got[b]), 'ref', float(ref[b]))
