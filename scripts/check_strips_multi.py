"""Multi-process strip check (run under torchrun on N GPUs): the sharded run must equal the oracle bit for bit.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/check_strips_multi.py
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

import sandengine_b200 as se  # noqa: E402
from sandengine_b200.distributed import StripSimulation  # noqa: E402
from sandengine_b200.grids import synthetic_grid  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rules = se.parse_path(REPO / "data" / "materials.yaml")
    ok = True
    for (W, H, steps, halo, T, dsync) in [(1024, 2048, 100, 8, 4, True), (1024, 2048, 100, 8, 4, False), (512, 1024, 77, 16, 8, True),
                                          (260, 512, 50, 4, 1, True)]:
        strip = StripSimulation(rules, (W, H), halo_rows=halo, device=local, temporal_block=T, device_sync=dsync)
        g = synthetic_grid(W, H, 7)
        strip.upload_cells(g[strip.row_begin:strip.row_end])
        strip.params.frame = 1
        strip.step(steps)
        mine = torch.from_numpy(strip.download_cells().astype(np.int64)).cuda()
        parts = [torch.empty((strip.plan.rows(r)[1] - strip.plan.rows(r)[0], W), dtype=torch.int64, device="cuda") for r in range(world)]
        dist.all_gather(parts, mine) if len({p.shape for p in parts}) == 1 else None
        if len({p.shape for p in parts}) != 1:
            # ragged strips: gather through the object path
            objs = [None] * world
            dist.all_gather_object(objs, mine.cpu().numpy())
            full = np.concatenate(objs, axis=0)
        else:
            full = torch.cat(parts, 0).cpu().numpy()
        if rank == 0:
            from oracle.build_oracle import load_oracle
            ref, _, _ = load_oracle().run(g, 1, steps, blocks=True)
            same = np.array_equal(full.astype(np.uint32), ref)
            print(f"[check_strips] world={world} {W}x{H} steps={steps} halo={halo} T={T} device_sync={dsync}: {'OK' if same else 'MISMATCH'}", file=sys.stderr, flush=True)
            ok = ok and same
        strip.close()
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
