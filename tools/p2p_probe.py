"""Probe (torchrun, >= 2 ranks): CUDA IPC between the per-GPU processes of one box and the NVLink
bandwidth a plain peer copy reaches.  The data-parallel path maps every rank's arena into every
other rank with exactly these calls (cudaIpcGetMemHandle / cudaIpcOpenMemHandle)."""
import os
import time

import torch
import torch.distributed as dist
from cuda.bindings import runtime as rt


class _Arr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def ck(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError("cuda error %s" % err)
    return res[1] if len(res) == 2 else res[1:]


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    nbytes = 256 << 20
    ptr = ck(rt.cudaMalloc(nbytes))
    mine = torch.as_tensor(_Arr(ptr, nbytes), device="cuda")
    mine.fill_(rank + 1)
    h = ck(rt.cudaIpcGetMemHandle(ptr))
    hb = bytes(h.reserved)
    hs = [None] * world
    dist.all_gather_object(hs, hb)
    peer = (rank + 1) % world
    ph = rt.cudaIpcMemHandle_t()
    ph.reserved = hs[peer]
    pptr = ck(rt.cudaIpcOpenMemHandle(ph, rt.cudaIpcMemLazyEnablePeerAccess))
    theirs = torch.as_tensor(_Arr(pptr, nbytes), device="cuda")
    torch.cuda.synchronize(); dist.barrier()
    ok = int(theirs[:1024].float().mean().item()) == peer + 1
    res = {}
    for name, dst, src in (("push", theirs, mine), ("pull", mine, theirs)):
        for _ in range(3):
            dst.copy_(src)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dst.copy_(src)
        e1.record(); torch.cuda.synchronize()
        res[name] = nbytes * 10 / (e0.elapsed_time(e1) / 1e3) / 1e9
        dist.barrier()
    # NCCL all-reduce of a dW-sized fp32 buffer for comparison
    g = torch.zeros(290000 * 256, device="cuda")
    for _ in range(3):
        dist.all_reduce(g)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(10):
        dist.all_reduce(g)
    torch.cuda.synchronize()
    ar_ms = (time.perf_counter() - t0) / 10 * 1e3
    print("rank %d/%d ipc_ok=%s push %.0f GB/s pull %.0f GB/s  nccl allreduce 297MB %.3f ms" %
          (rank, world, ok, res["push"], res["pull"], ar_ms), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
