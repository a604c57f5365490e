import torch.distributed as dist


def _on():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if _on() else 1


def get_rank():
    return dist.get_rank() if _on() else 0


def get_local_rank():
    return get_rank()


def is_main_process():
    return get_rank() == 0


def synchronize():
    if get_world_size() > 1:
        dist.barrier()


def gather(data, dst=0):
    if get_world_size() == 1:
        return [data]
    out = [None] * get_world_size() if get_rank() == dst else None
    dist.gather_object(data, out, dst=dst)
    return out or []


def all_gather(data):
    if get_world_size() == 1:
        return [data]
    out = [None] * get_world_size()
    dist.all_gather_object(out, data)
    return out
