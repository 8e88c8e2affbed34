"""One process per GPU (torchrun): device selection and NCCL initialisation shared by every CLI."""
import os

import torch
import torch.distributed as dist


def dist_init():
    """-> (rank, world, device).  Under torchrun (WORLD_SIZE > 1) initialises the NCCL process group once."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, torch.device("cuda", local)


def dist_finish():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
