"""N > 1 host logic on CPU: two (and three) processes over the gloo backend run the strip plan + ghost-row
schedule of sandengine_b200.distributed (StripPlan: rows, ghosts, chunks) with the ORACLE as the per-strip
stepper and gloo send/recv as the exchange, and must reproduce the single-grid oracle bit for bit.
The GPU data path (se_sim_halo_exchange_async) is covered by the -m gpu strip tests and scripts/check_strips_multi.py."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, W, H, halo, steps, seed, out_dir):
    sys.path.insert(0, str(REPO))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), OMP_NUM_THREADS="1")
    import torch
    import torch.distributed as dist
    from oracle.build_oracle import load_oracle
    from sandengine_b200.distributed import StripPlan
    from sandengine_b200.grids import synthetic_grid

    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = StripPlan(W, H, world, halo)
    b, e = plan.rows(rank)
    gt, gb = plan.ghosts(rank)
    orc = load_oracle()
    local = np.zeros((gt + (e - b) + gb, W), np.uint32)
    local[gt:gt + (e - b)] = synthetic_grid(W, H, seed, row_begin=b, row_end=e)
    gy0 = b - gt

    def exchange():
        # push my boundary rows into the neighbours' ghost rows (same geometry as se_sim_halo_push)
        reqs, bufs = [], []
        if rank > 0:
            n_gb = plan.ghosts(rank - 1)[1]
            reqs.append(dist.isend(torch.from_numpy(local[gt:gt + n_gb].astype(np.int64)), rank - 1))
            buf = torch.empty((gt, W), dtype=torch.int64); bufs.append(("top", buf)); reqs.append(dist.irecv(buf, rank - 1))
        if rank < world - 1:
            n_gt = plan.ghosts(rank + 1)[0]
            reqs.append(dist.isend(torch.from_numpy(local[gt + (e - b) - n_gt:gt + (e - b)].astype(np.int64)), rank + 1))
            buf = torch.empty((gb, W), dtype=torch.int64); bufs.append(("bot", buf)); reqs.append(dist.irecv(buf, rank + 1))
        for r in reqs:
            r.wait()
        for where, buf in bufs:
            if where == "top":
                local[:gt] = buf.numpy().astype(np.uint32)
            else:
                local[gt + (e - b):] = buf.numpy().astype(np.uint32)

    exchange()
    frame = 1
    for k in plan.chunks(steps):
        for _ in range(k):
            frame += 1
            orc.step_blocks_strip(local, gy0, H, frame)
        exchange()
    np.save(Path(out_dir) / f"strip{rank}.npy", local[gt:gt + (e - b)])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,W,H,halo,steps", [(2, 48, 64, 4, 37), (2, 33, 50, 2, 21), (3, 40, 90, 6, 40)])
def test_strip_schedule_over_gloo(tmp_path, oracle, world, W, H, halo, steps):
    import torch.multiprocessing as mp
    from sandengine_b200.grids import synthetic_grid
    port = 29600 + (os.getpid() % 300) + world
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, halo, steps, 11, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    got = np.concatenate([np.load(tmp_path / f"strip{r}.npy") for r in range(world)], axis=0)
    ref, _, _ = oracle.run(synthetic_grid(W, H, 11), 1, steps, blocks=True)
    assert np.array_equal(got, ref)


def test_strip_plan_arithmetic():
    from sandengine_b200.distributed import StripPlan
    p = StripPlan(16384, 16384, 8, 64)
    assert [p.rows(r) for r in range(8)] == [(2048 * r, 2048 * (r + 1)) for r in range(8)]
    assert p.ghosts(0) == (0, 64) and p.ghosts(3) == (64, 64) and p.ghosts(7) == (64, 0)
    assert p.steps_per_exchange == 120 and p.chunks(200) == [120, 80] and p.chunks(0) == []
    assert StripPlan(16384, 16384, 8, 34).steps_per_exchange == 64 and StripPlan(64, 64, 2, 2).steps_per_exchange == 2 and StripPlan(64, 64, 2, 4).steps_per_exchange == 6
    q = StripPlan(100, 50, 3, 4)      # ragged: even boundaries, last strip takes the odd tail
    rows = [q.rows(r) for r in range(3)]
    assert rows[0][0] == 0 and rows[-1][1] == 50 and all(b % 2 == 0 for b, _ in rows) and all(rows[i][1] == rows[i + 1][0] for i in range(2))
    assert StripPlan(64, 64, 1, 0).chunks(10) == [10]
    with pytest.raises(ValueError):
        StripPlan(64, 64, 2, 3)
    with pytest.raises(ValueError):
        StripPlan(64, 16, 4, 8)


def test_a_strip_snapshot_is_refused_by_the_single_device_loader(tmp_path, native_lib):
    """a strip file holds owned rows only; `snapshot.load` must not build a strip without neighbours from it
    (no GPU needed: the refusal comes before any device object is made).  `load_strip` is exercised on GPUs by
    scripts/check_strips_multi.py."""
    import json

    import numpy as np

    import sandengine_b200 as se
    from sandengine_b200 import snapshot
    rules = se.parse_path(REPO / "data" / "materials.yaml", compile=False)
    meta = {"format": snapshot.FORMAT, "width": 64, "height": 128, "row_begin": 64, "row_end": 128, "halo_rows": 8, "frame": 5,
            "lighting": False, "rules_sha256": snapshot.rules_digest(rules)}
    np.savez_compressed(tmp_path / "strip.npz", meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), cells=np.zeros((64, 64), np.uint32))
    with pytest.raises(ValueError, match="load_strip"):
        snapshot.load(rules, tmp_path / "strip.npz")
