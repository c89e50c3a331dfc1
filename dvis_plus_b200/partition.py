"""Spatial SM partitioning for the two stages of the clip pipeline (CUDA green contexts, driver API through cuda-python).

The per-frame stage (big bandwidth / tensor bound kernels) and the temporal stage (a dependent chain of ~1 100 tiny kernels per
clip) overlap across clips (pipeline.GraphedClipRunner).  On one shared set of SMs every tiny kernel that needs a large
shared-memory configuration drains big-kernel CTAs from the SMs it lands on (tests/perf/interference_probe.py: 0.7 - 2 us of
the big stream lost per tiny kernel; the pipelined step is 19.2 ms against 14.5 ms for the per-frame stage alone).  Giving the
temporal stage a small SM partition of its own removes the interference at the price of those SMs.

sm_partition_streams(n_small) -> (stream_big, stream_small, info): torch ExternalStreams living in two green contexts with
disjoint SM sets.  Raises RuntimeError when the driver / device cannot partition."""
import torch


def _check(res, what):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"{what} failed: {err}")
    return res[1:] if len(res) > 2 else res[1]


def sm_partition_streams(n_small, device=0, small_priority=-1):
    from cuda.bindings import driver as drv
    torch.cuda.init()
    torch.zeros(1, device=f"cuda:{device}")                       # primary context active
    dev = _check(drv.cuDeviceGet(device), "cuDeviceGet")
    sm = _check(drv.cuDeviceGetDevResource(dev, drv.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM), "cuDeviceGetDevResource")
    groups, n_groups, remaining = _check(drv.cuDevSmResourceSplitByCount(1, sm, 0, n_small), "cuDevSmResourceSplitByCount")
    if n_groups < 1:
        raise RuntimeError("the device did not produce an SM group")
    small, big = groups[0], remaining
    streams, keep = [], []
    for res, prio in ((big, 0), (small, small_priority)):
        desc = _check(drv.cuDevResourceGenerateDesc([res], 1), "cuDevResourceGenerateDesc")
        gctx = _check(drv.cuGreenCtxCreate(desc, dev, drv.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM), "cuGreenCtxCreate")
        st = _check(drv.cuGreenCtxStreamCreate(gctx, drv.CUstream_flags.CU_STREAM_NON_BLOCKING, prio), "cuGreenCtxStreamCreate")
        keep.append((gctx, st))
        streams.append(torch.cuda.ExternalStream(int(st), device=device))
    info = {"sms_big": int(big.sm.smCount), "sms_small": int(small.sm.smCount), "_keep": keep}
    return streams[0], streams[1], info
