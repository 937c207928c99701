"""Multi-process strip check (run under torchrun on N GPUs): the sharded run must equal the oracle bit for bit.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/check_strips_multi.py
Covers the fused ghost-row push of the tile kernel (runs of steps), the stream-ordered exchange of the per-step kernels
(single steps, modifications, a rule set without a transition table), mixtures of both, and lit strips.
Rank 0 prints one line per case and writes them to gpurun_out/check_strips_<N>.log (kept under profiles/).
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

import sandengine_b200 as se  # noqa: E402
from sandengine_b200.distributed import StripSimulation  # noqa: E402
from sandengine_b200.grids import synthetic_grid  # noqa: E402


def gather_rows(strip, mine, world):
    objs = [None] * world
    dist.all_gather_object(objs, mine)
    return np.concatenate(objs, axis=0)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rules = se.parse_path(REPO / "data" / "materials.yaml")
    lines, ok = [], True

    def report(name, same):
        nonlocal ok
        ok = ok and same
        lines.append(f"[check_strips] world={world} {name}: {'OK' if same else 'MISMATCH'}")
        print(lines[-1], file=sys.stderr, flush=True)

    from oracle.build_oracle import load_oracle
    # ---- lighting off: (W, H, schedule of se_sim_step calls, halo, T) ----
    for (W, H, sched, halo, T) in [(1024, 2048, [100], 8, 4), (1024, 2048, [20, 20, 7, 1, 1, 40], 34, 8), (512, 1024, [77], 16, 8),
                                   (260, 512, [50], 4, 1), (2048, 4096, [64, 1, 1, 1, 33], 34, 0), (16384, 512 * world, [25], 34, 0)]:
        strip = StripSimulation(rules, (W, H), halo_rows=halo, device=local, temporal_block=T)
        g = synthetic_grid(W, H, 7)
        strip.upload_cells(g[strip.row_begin:strip.row_end])
        strip.params.frame = 1
        for n in sched:
            strip.step(n)
        full = gather_rows(strip, strip.download_cells(), world)
        if rank == 0:
            ref, _, _ = load_oracle().run(g, 1, sum(sched), blocks=True)
            report(f"{W}x{H} steps={sched} halo={halo} T={T}", bool(np.array_equal(full, ref)))
        strip.close()
        dist.barrier()
    # ---- modifications every third frame (broadcast to every strip, global coordinates) ----
    from test_gpu_parity import make_mods
    W, H, steps = 768, 1024, 30
    rng = np.random.default_rng(77)
    g = synthetic_grid(W, H, 17)
    mods = [make_mods(se, s, 11, W, H, rng) if s % 3 == 0 else np.zeros(0, se.MOD_DTYPE) for s in range(steps)]
    strip = StripSimulation(rules, (W, H), halo_rows=8, device=local)
    strip.upload_cells(g[strip.row_begin:strip.row_end])
    strip.params.frame = 1
    for k in range(steps):
        if len(mods[k]):
            strip.sim.push_modifications(mods[k])
        strip.step(1)
    full = gather_rows(strip, strip.download_cells(), world)
    if rank == 0:
        ref, _, _ = load_oracle().run(g, 1, steps, mods_per_step=mods)
        report(f"{W}x{H} modifications every 3rd of {steps} frames", bool(np.array_equal(full, ref)))
    strip.close()
    dist.barrier()
    # ---- lit strips: ids bit-exact, light within 1e-6 ----
    W, H, steps = 512, 768, 40
    g = synthetic_grid(W, H, 21)
    L0 = np.random.default_rng(23).random((H, W, 4), dtype=np.float32)
    strip = StripSimulation(rules, (W, H), halo_rows=8, device=local, lighting=True)
    strip.upload_cells(g[strip.row_begin:strip.row_end])
    strip.upload_light(np.ascontiguousarray(L0[strip.row_begin:strip.row_end]))
    strip.params.frame = 1
    strip.step(steps)
    full = gather_rows(strip, strip.download_cells(), world)
    fullL = gather_rows(strip, strip.download_light(), world)
    if rank == 0:
        ref, refL, _ = load_oracle().run(g, 1, steps, light=L0)
        err = float(np.abs(fullL - refL).max())
        report(f"{W}x{H} lighting on, {steps} steps (light max-abs err {err:.1e})", bool(np.array_equal(full, ref)) and err <= 1e-6)
    strip.close()
    dist.barrier()
    # ---- strip snapshots: save on every rank after 21 steps, restore with load_strip, 30 more steps ----
    import tempfile
    from sandengine_b200 import snapshot
    W, H = 512, 1024
    g = synthetic_grid(W, H, 31)
    strip = StripSimulation(rules, (W, H), halo_rows=16, device=local)
    strip.upload_cells(g[strip.row_begin:strip.row_end])
    strip.params.frame = 1
    strip.step(21)
    tmp = Path(tempfile.gettempdir()) / f"se_strip_snapshot_{os.getpid()}_{rank}.npz"
    snapshot.save(strip, tmp)
    strip.close()
    dist.barrier()
    strip = snapshot.load_strip(rules, tmp, device=local)
    tmp.unlink()
    frame_ok = strip.params.frame == 22
    strip.step(30)
    full = gather_rows(strip, strip.download_cells(), world)
    if rank == 0:
        ref, _, _ = load_oracle().run(g, 1, 51, blocks=True)
        report(f"{W}x{H} strip snapshot after 21 steps, restored with load_strip, 30 more steps", frame_ok and bool(np.array_equal(full, ref)))
    strip.close()
    dist.barrier()
    if rank == 0:
        out = REPO / "gpurun_out"
        out.mkdir(exist_ok=True)
        (out / f"check_strips_{world}.log").write_text("\n".join(lines) + "\n")
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
