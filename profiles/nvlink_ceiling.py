#!/usr/bin/env python
"""Raw NVLink all-gather ceiling of the box, for bench.py's `roofline.nvlink` / `config.merge`.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           profiles/nvlink_ceiling.py [total_GB]

The merged cloud of the 1 h stream is 3.6e8 x (16 + 14) B = 10.8 GB on EVERY rank; rank r produces 1/N of
it and must receive the other (N-1)/N.  This script moves exactly those bytes with no compute at all:
  nccl        ncclAllGather, in place (torch.distributed.all_gather_into_tensor)
  ce_push     every rank copies its slice into each peer's buffer with the copy engines
              (one cudaMemcpyAsync per peer on its own stream; symmetric-memory peer mappings)
  ce_pull     every rank copies each peer's slice into its own buffer
Timing: CUDA events, cross-rank barrier before and after, max over ranks.  The best of the three is the
ceiling any merged-cloud assembly (NCCL or fused into the kernel) can reach on this box.
"""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    total = int(float(sys.argv[1]) * 1e9) if len(sys.argv) > 1 else 10_800_000_000
    per = total // world // 256 * 256
    total = per * world
    buf = symm.empty(total, dtype=torch.uint8, device=dev)
    h = symm.rendezvous(buf, dist.group.WORLD)
    buf.fill_(rank + 1)
    mine = buf[rank * per:(rank + 1) * per]
    peers = [r for r in range(world) if r != rank]
    peer_bufs = {r: h.get_buffer(r, (total,), torch.uint8) for r in peers}
    streams = [torch.cuda.Stream(dev) for _ in peers]

    def timed(fn, reps=5):
        best = None
        for it in range(reps + 1):
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if it:
                best = float(t.item()) if best is None else min(best, float(t.item()))
        return best

    def nccl():
        dist.all_gather_into_tensor(buf, mine)

    def ce(push):
        cur = torch.cuda.current_stream()
        for s, r in zip(streams, peers):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                if push:
                    peer_bufs[r][rank * per:(rank + 1) * per].copy_(mine, non_blocking=True)
                else:
                    buf[r * per:(r + 1) * per].copy_(peer_bufs[r][r * per:(r + 1) * per], non_blocking=True)
        for s in streams:
            cur.wait_stream(s)
        h.barrier()

    recv = per * (world - 1)
    res = {"n_gpus": world, "total_bytes": total, "bytes_received_per_rank": recv, "multicast": bool(int(getattr(h, "multicast_ptr", 0) or 0))}
    for name, fn in (("nccl", nccl), ("ce_push", lambda: ce(True)), ("ce_pull", lambda: ce(False))):
        try:
            ms = timed(fn)
            res[name] = {"ms": ms, "ingress_GBps_per_rank": recv / ms / 1e6}
        except Exception as e:                    # noqa: BLE001
            res[name] = {"error": repr(e)}
    ok = [v for k, v in res.items() if isinstance(v, dict) and "ms" in v]
    res["ceiling_ms"] = min(v["ms"] for v in ok) if ok else None
    res["ceiling_ingress_GBps_per_rank"] = max(v["ingress_GBps_per_rank"] for v in ok) if ok else None
    # the buffer really holds every rank's slice
    chk = all(int(buf[r * per].item()) == r + 1 and int(buf[(r + 1) * per - 1].item()) == r + 1 for r in range(world))
    res["gathered_ok"] = chk
    if rank == 0:
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, f"nvlink_ceiling_n{world}.json"), "w") as f:
            json.dump(res, f, indent=1)
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
