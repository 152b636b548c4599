"""Torch restatement of the K-transpose pack/unpack layout of csrc/trp.cu (src/base/tlab_mpi_transpose.f90:301-325),
used by the gloo world-size-2 CPU test only.  Test infrastructure: the product's transposes run inside libtlab_gpu.so."""
import torch


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
