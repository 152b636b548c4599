"""Error levels of the line operators against the oracle (not a test: prints rel. L2 per operator and shape)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from common import grid_periodic, grid_tanh, smooth_field, rel_l2
from oracle import fdm, operators as O
from tlab_b200 import opr

dev = torch.device("cuda:0")
for nx, ny, nz in [(1024, 32, 32), (2048, 16, 32), (32, 512, 32), (32, 1024, 32), (32, 16, 1024), (32, 16, 2048)]:
    x, y, z = grid_periodic(nx), grid_tanh(ny), grid_periodic(nz)
    go = [fdm.Plan(x, True, True), fdm.Plan(y, False, False), fdm.Plan(z, True, True)]
    gg = [opr.FdmPlan(x, True, True), opr.FdmPlan(y, False, False), opr.FdmPlan(z, True, True)]
    B = O.Burgers(go, 1.0 / 5000.0, [1.0])
    opr.OPR_Burgers_Initialize(gg, 1.0 / 5000.0, [1.0])
    a = smooth_field((nz, ny, nx), (x, y, z), seed=1); b = smooth_field((nz, ny, nx), (x, y, z), seed=2)
    u, v = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    bcs = [[0, 0], [0, 0]]
    d = int(np.argmax([nx, ny, nz]))
    P = [opr.OPR_Partial_X, opr.OPR_Partial_Y, opr.OPR_Partial_Z][d]
    Bg = [opr.OPR_Burgers_X, opr.OPR_Burgers_Y, opr.OPR_Burgers_Z][d]
    r1, r2, r3, r4 = (torch.empty_like(u) for _ in range(4))
    P(opr.OPR_P2_P1, nx, ny, nz, bcs, gg[d], u, r1, r2)
    ref = O.opr_partial(d, O.OPR_P2_P1, bcs, go[d], a)
    P(opr.OPR_P1, nx, ny, nz, bcs, gg[d], u, r4)
    Bg(opr.OPR_B_U_IN, 0, nx, ny, nz, bcs, u, v, r3)
    print("dir %s n=%4d  P2 %.1e  P1(P2_P1) %.1e  P1 %.1e  Burgers %.1e" % ("xyz"[d], (nx, ny, nz)[d], rel_l2(r1.cpu().numpy(), ref[0]),
          rel_l2(r2.cpu().numpy(), ref[1]), rel_l2(r4.cpu().numpy(), O.opr_partial(d, O.OPR_P1, bcs, go[d], a)),
          rel_l2(r3.cpu().numpy(), B.apply(d, 0, bcs, a, b))), flush=True)
