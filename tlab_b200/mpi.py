"""Domain decomposition plumbing: z slabs, one process per GPU.

Mirrors the reference's rank layout (src/base/tlab_mpi_procs.f90:40-58, ims_npro_k = world size, ims_npro_i = 1)
and the K-transpose index maps (src/base/tlab_mpi_transpose.f90:301-325).  torch.distributed is used only to
exchange the NCCL unique id (the Fortran host would MPI_Bcast it); the transposes of the product run inside
libtlab_gpu.so on its own NCCL communicator.

The functions `pack_k`, `unpack_k`, `trp_k_forward_ref`, `trp_k_backward_ref` restate the pack/unpack layout
of csrc/trp.cu on torch tensors so that the host-side logic can be tested with the gloo backend on CPUs.
"""
import ctypes

import numpy as np
import torch

from . import lib as _lib


def init_from_torch_distributed():
    """Create the library's NCCL communicator over the default torch.distributed process group."""
    import torch.distributed as dist
    L = _lib.load()
    rank, world = dist.get_rank(), dist.get_world_size()
    idbuf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0 and world > 1:
        raw = (ctypes.c_ubyte * 128)()
        _lib.check(L.tlab_mpi_get_unique_id(raw))
        idbuf = torch.tensor(list(raw), dtype=torch.uint8)
    if world > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = idbuf.to(dev)
        dist.broadcast(t, src=0)
        idbuf = t.cpu()
    raw = (ctypes.c_ubyte * 128)(*[int(v) for v in idbuf.tolist()])
    _lib.check(L.tlab_mpi_init(rank, world, raw))
    return rank, world


def finalize():
    _lib.load().tlab_mpi_finalize()


def slab(nz, rank, world):
    """(kmax, offset) of this rank's z slab; the reference requires nz % world == 0 (tlab_mpi_procs.f90:44-58)."""
    if nz % world:
        raise ValueError("Kmax(*) must divide the grid size in z")
    kmax = nz // world
    return kmax, rank * kmax


# ---- layout restatement (torch, device agnostic) -------------------------------------------------------------
def pack_k(a, P):
    """a: slab (kmax, nxy) [C order = Fortran a(nxy, kmax)] -> send buffer (P, kmax, nxy/P)."""
    kmax, nxy = a.shape
    nl = nxy // P
    return a.reshape(kmax, P, nl).permute(1, 0, 2).contiguous()


def unpack_k(buf):
    """receive buffer (P, kmax, nl) -> slab (kmax, P*nl)."""
    P, kmax, nl = buf.shape
    return buf.permute(1, 0, 2).reshape(kmax, P * nl).contiguous()


def trp_k_forward_ref(a, group=None):
    """TLabMPI_Trp_ExecK_Forward with torch.distributed.all_to_all_single: slab (kmax, nxy) -> pencil (nz, nxy/P)."""
    import torch.distributed as dist
    P = dist.get_world_size(group)
    send = pack_k(a, P)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)
    kmax, nl = send.shape[1], send.shape[2]
    return recv.reshape(P * kmax, nl)            # block q holds planes q*kmax .. (q+1)*kmax - 1


def trp_k_backward_ref(b, kmax, group=None):
    """TLabMPI_Trp_ExecK_Backward: pencil (nz, nl) -> slab (kmax, nl*P)."""
    import torch.distributed as dist
    P = dist.get_world_size(group)
    nl = b.shape[1]
    send = b.reshape(P, kmax, nl).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)
    return unpack_k(recv)
