#!/usr/bin/env python
"""Aggregate host->device / device->host bandwidth of the box when N ranks copy at once (the ceiling of the
end-to-end host-buffer path at N GPUs), with the NUMA placement of every GPU and of the pinned buffers.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe_multi.py

Each rank moves 1 GiB between a pinned host buffer (smarl_host_alloc_pinned: bound to the GPU's NUMA node when the
kernel allows it) and its GPU; first every rank alone (the others wait), then all ranks at the same time.  One
cudaMemcpyAsync per copy."""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from safe_multiagent_rl_b200 import _lib  # noqa: E402
from safe_multiagent_rl_b200 import dist as sd  # noqa: E402


def main():
    rank, world, local = sd.init_from_env()
    torch.cuda.set_device(local)
    lib = _lib.load()
    n = 1 << 30
    raw, node = C.c_void_p(), C.c_int32(-1)
    _lib.check(lib.smarl_host_alloc_pinned(C.byref(raw), n, C.byref(node)))
    h = torch.from_numpy(np.ctypeslib.as_array((C.c_uint8 * n).from_address(raw.value)))
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def rate(fn, it=4):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(it):
            fn()
        torch.cuda.synchronize()
        return n * it / (time.perf_counter() - t0) / 1e9

    def both():
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
    solo = [0.0, 0.0]
    for r in range(world):                                  # one rank at a time
        barrier()
        if r == rank:
            solo = [rate(lambda: d.copy_(h, non_blocking=True)), rate(lambda: h.copy_(d, non_blocking=True))]
    barrier()
    together_h2d = rate(lambda: d.copy_(h, non_blocking=True))
    barrier()
    together_d2h = rate(lambda: h.copy_(d, non_blocking=True))
    barrier()
    together_both = rate(both)
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
    except Exception:
        bus = -1
    row = torch.tensor([solo[0], solo[1], together_h2d, together_d2h, together_both, float(node.value), float(bus)],
                       dtype=torch.float64, device="cuda")
    rows = [row.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(rows, row)
    if rank == 0:
        print(f"## PCIe probe, {world} rank(s) on one box, 1 GiB pinned copies\n")
        print("| rank | GPU PCI bus | pinned-buffer NUMA node | H2D alone GB/s | D2H alone | H2D all ranks at once | D2H all at once | H2D while D2H (each way) |")
        print("|---:|---:|---:|---:|---:|---:|---:|---:|")
        for r, x in enumerate(rows):
            x = x.cpu().tolist()
            print(f"| {r} | {int(x[6]):#04x} | {int(x[5])} | {x[0]:.1f} | {x[1]:.1f} | {x[2]:.1f} | {x[3]:.1f} | {x[4]:.1f} |")
        tot = torch.stack(rows).sum(0).cpu().tolist()
        print(f"| sum | | | {tot[0]:.1f} | {tot[1]:.1f} | **{tot[2]:.1f}** | **{tot[3]:.1f}** | {tot[4]:.1f} |")
        print("\n```")
        for cmd in (["nvidia-smi", "topo", "-m"], ["grep", "-E", "Cpus_allowed_list|Mems_allowed_list", "/proc/self/status"],
                    ["sh", "-c", "ls /sys/devices/system/node/ | grep node; cat /sys/devices/system/node/node*/cpulist"]):
            try:
                print(subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout.strip())
            except Exception as ex:
                print(cmd, "failed:", ex)
        print("```")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
