"""Multi-GPU data parallelism of the E-step: whole contigs are the unit, assigned to ranks by
longest-processing-time-first; the only exchange is one sum all-reduce of the raw statistics vector
[LL | E0(N) E1(N) | RL CL RU CU AD] (7N+1 doubles) per EM iteration, after which rank 0 runs the M-step (it is a serial
search; replicating it on every rank of one host only makes the ranks fight for cores) and broadcasts the parameters."""
import numpy as np


def lpt_shards(lengths, n_ranks):
    """owner[i] = rank that holds sequence i (longest first onto the least loaded rank; ties -> lowest rank)"""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * n_ranks
    owner = [0] * len(lengths)
    for i in order:
        g = min(range(n_ranks), key=lambda r: (load[r], r))
        owner[i] = g
        load[g] += int(lengths[i])
    return owner


def stats_len(n_states):
    return 7 * n_states + 1


def pack_raw(LL, E, RL, CL, RU, CU, AD):
    """raw statistics vector in the layout psmc_b200_device_stats() / psmc_b200_unpack_stats() use"""
    return np.concatenate([[LL], np.asarray(E).ravel(), RL, CL, RU, CU, AD]).astype(np.float64)


def all_reduce_raw(raw_tensor, group=None):
    """the one collective of an EM iteration (NCCL on GPUs, gloo in the CPU tests)"""
    import torch.distributed as dist
    dist.all_reduce(raw_tensor, op=dist.ReduceOp.SUM, group=group)
    return raw_tensor


def broadcast_params(params_tensor, src=0, group=None):
    """parameters found by the M-step on rank `src` -> every rank (NCCL on GPUs, gloo in the CPU tests)"""
    import torch.distributed as dist
    dist.broadcast(params_tensor, src, group=group)
    return params_tensor
