"""Worker for the 2-GPU parity test (launched with torchrun): z-slab RK step vs the single-domain oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import grid_periodic, grid_tanh, smooth_field, rel_l2  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tlab_b200 import lib as tl, opr, dns as GD, mpi
    tl.check(tl.load().tlab_gpu_init(local))
    for kv in [t for t in os.environ.get("TLAB_TUNE", "").split(",") if t]:
        k, v = kv.split("=")
        tl.check(tl.load().tlab_gpu_set_tuning(k.encode(), int(v)))
    mpi.init_from_torch_distributed()
    nx, ny, nz = [int(v) for v in os.environ.get("TLAB_SHAPE", "32,32,32").split(",")]
    kmax, koff = mpi.slab(nz, rank, world)
    x, y, z = grid_periodic(nx), grid_tanh(ny), grid_periodic(nz)
    gg = [opr.FdmPlan(x, True, True, name="x"), opr.FdmPlan(y, False, False, name="y"), opr.FdmPlan(z, True, True, name="z")]
    D, N = GD.DNS_BCS_DIRICHLET, GD.DNS_BCS_NEUMANN
    kw = dict(visc=1.0 / 5000.0, schmidt=[1.0], buoyancy_type="linear", buoyancy_params=(1.0, 0.0),
              buoyancy_vector=(0.0, 1.0, 0.0), bcs_flow_jmin=(D, D, D), bcs_flow_jmax=(N, D, N),
              bcs_scal_jmin=(D,), bcs_scal_jmax=(N,))
    g = GD.Dns(gg, kmax=kmax, **kw)
    wall = np.sin(0.5 * np.pi * y / y[-1])[None, :, None]
    full = [0.5 * smooth_field((nz, ny, nx), (x, y, z), seed=31 + i) * wall for i in range(3)]
    sc = 0.5 + 0.1 * smooth_field((nz, ny, nx), (x, y, z), seed=40) * wall
    for i in range(3):
        g.set("q%d" % (i + 1), full[i][koff:koff + kmax])
    g.set("s1", sc[koff:koff + kmax])
    for _ in range(2):
        g.runge_kutta(1e-3)
    mine = [g.get("q%d" % (i + 1)) for i in range(3)] + [g.get("s1")]
    errs = None
    if rank == 0:
        from oracle import fdm, dns as OD
        go = [fdm.Plan(x, True, True, name="x"), fdm.Plan(y, False, False, name="y"), fdm.Plan(z, True, True, name="z")]
        o = OD.Dns(go, **kw)
        for i in range(3):
            o.q[i][...] = full[i]
        o.s[0][...] = sc
        for _ in range(2):
            o.runge_kutta(1e-3)
        ref = o.q + o.s
    gathered = []
    for f in mine:
        t = torch.from_numpy(f).cuda()
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        gathered.append(torch.cat(out, dim=0).cpu().numpy())
    if rank == 0:
        errs = [rel_l2(a, b) for a, b in zip(gathered, ref)]
        import ctypes
        cnt = {}
        for key in ("p2p_exchanges", "nccl_exchanges", "splitz_ops"):
            c = ctypes.c_longlong()
            tl.check(tl.load().tlab_gpu_get_counter(key.encode(), ctypes.byref(c)))
            cnt[key] = c.value
        print("DIST_PATH p2p=%d nccl=%d splitz=%d" % (cnt["p2p_exchanges"], cnt["nccl_exchanges"], cnt["splitz_ops"]), flush=True)
        print("DIST_ERRS", " ".join("%.3e" % e for e in errs), flush=True)
        assert max(errs) <= 1e-11, errs
    g.close()
    mpi.finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
