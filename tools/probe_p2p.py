#!/usr/bin/env python
"""Probe which peer-memory mechanisms work between the ranks of one box (run under torchrun, 2+ GPUs)."""
import os
import sys
import torch
import torch.distributed as dist


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    res = {}
    # 1. symmetric memory
    try:
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(1024, dtype=torch.float32, device=torch.device("cuda", lr))
        t.fill_(float(rank + 1))
        hdl = symm.rendezvous(t, dist.group.WORLD)
        torch.cuda.synchronize()
        dist.barrier()
        peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
        res["symm"] = float(peer[0].item())
        res["symm_ptrs"] = len(hdl.buffer_ptrs)
        res["symm_signal"] = len(hdl.signal_pad_ptrs)
    except Exception as e:  # noqa: BLE001
        res["symm_err"] = repr(e)[:200]
    # 2. legacy CUDA IPC through torch storages
    try:
        x = torch.full((1024,), float(rank + 10), device="cuda")
        info = x.untyped_storage()._share_cuda_()
        infos = [None] * world
        dist.all_gather_object(infos, info)
        pi = infos[(rank + 1) % world]
        st = torch.UntypedStorage._new_shared_cuda(*pi)
        y = torch.empty(0, dtype=torch.float32, device=st.device).set_(st, 0, (1024,))
        torch.cuda.synchronize()
        dist.barrier()
        res["ipc"] = float(y[0].item())
        res["ipc_dev"] = str(y.device)
    except Exception as e:  # noqa: BLE001
        res["ipc_err"] = repr(e)[:200]
    res["can_access_peer"] = torch.cuda.can_device_access_peer(lr, (lr + 1) % world)
    print(rank, res, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    os._exit(0)


if __name__ == "__main__":
    main()
