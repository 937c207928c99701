"""Strip sharding of one grid over the GPUs of a node: one process per GPU,
`torch.distributed` for the plumbing (rendezvous, IPC-handle exchange, barriers), ghost rows pushed
device-to-device over NVLink by the native library -- no data-path collective.

The exchange itself lives in the library: the tile kernel (K1b) stores the rows next to a strip boundary into the
neighbour's ghost rows as part of its store phase and publishes per-tile flags in the neighbour's memory, so a run of
steps needs neither a separate copy nor a host synchronisation; the per-step kernels (modifications, lighting, rule
sets without a transition table) recompute the ghost rows redundantly and `se_sim_step` exchanges on the stream
(cuStreamWriteValue32 / cuStreamWaitValue32) when they are used up.  `StripSimulation.step(n)` is therefore just
`Simulation.step(n)` on every rank.  `StripPlan` holds the row arithmetic and the schedule of the redundant-ghost scheme
and is shared with the CPU (gloo) emulation the tests use.

The grid is cut into horizontal strips on EVEN rows.  A strip keeps `halo_rows` ghost rows towards each
neighbour and re-computes them redundantly: the Margolus update is block-local and its RAND depends
only on (block position, frame) (falling_sand.glsl:698), so both neighbours compute identical values
for the rows they share.  A step invalidates the outermost still-valid ghost row only when its block is cut
by the buffer edge, and the Margolus ROW offset changes every other frame (operations.glsl:25-34:
frame%4 -> 1,1,0,0), so n steps invalidate at most floor(n/2)+1 rows: G ghost rows are good for 2(G-1)
steps; then the owners push fresh copies.  `StripPlan` holds that arithmetic and is shared by the GPU path and by the CPU (gloo)
emulation the tests use.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple


@dataclass(frozen=True)
class StripPlan:
    width: int
    height: int
    world: int
    halo_rows: int
    lighting: bool = False

    def __post_init__(self):
        if self.halo_rows % 2:
            raise ValueError("halo_rows must be even (strip buffers must start on even rows)")
        if self.world > 1 and self.halo_rows < 2:
            raise ValueError("need at least 2 ghost rows to shard")
        if self.world > 1 and self.height < 2 * self.world:
            raise ValueError("grid too small to shard")
        if self.world > 1:
            smallest = min(e - b for b, e in (self.rows(r) for r in range(self.world)))
            if smallest < self.halo_rows:
                raise ValueError("halo_rows exceeds a strip's own rows")

    def rows(self, rank: int) -> Tuple[int, int]:
        """Owned global rows [begin, end) of `rank`; boundaries are even."""
        blocks = (self.height + 1) // 2
        b0 = (blocks * rank) // self.world
        b1 = (blocks * (rank + 1)) // self.world
        return 2 * b0, min(self.height, 2 * b1)

    def ghosts(self, rank: int) -> Tuple[int, int]:
        b, e = self.rows(rank)
        return (min(self.halo_rows, b) if rank > 0 else 0), (min(self.halo_rows, self.height - e) if rank < self.world - 1 else 0)

    @property
    def steps_per_exchange(self) -> int:
        """n with floor(n/2) + 1 <= halo_rows, rounded down to a multiple of 8 (whole fused launches) when possible.
        With lighting the 8-neighbour light stencil (operations.glsl:114-160) eats one ghost row per step: n = halo_rows."""
        if self.lighting:
            return self.halo_rows
        n = 2 * (self.halo_rows - 1)
        return n // 8 * 8 if n >= 8 else n

    def chunks(self, n_steps: int) -> List[int]:
        """Step counts between ghost exchanges."""
        if self.world == 1:
            return [n_steps] if n_steps else []
        g = self.steps_per_exchange
        return [g] * (n_steps // g) + ([n_steps % g] if n_steps % g else [])


class StripSimulation:
    """One rank's strip of a sharded `Simulation`.  `step(n)` is `Simulation.step(n)` on every rank: the library keeps the ghost
    rows current (pushed by the boundary tiles of the tile kernel in runs of steps, exchanged on the stream when the per-step
    kernels have used them up).  With lighting=True the light field gets ghost rows too (one is used up per step)."""

    def __init__(self, rules, size, halo_rows: int = 32, device=None, temporal_block: int = 0, device_sync: bool = True,
                 running_census: bool = False, lighting: bool = False, device_share: int = 1):
        import torch
        import torch.distributed as dist

        from .simulation import Simulation

        self.dist = dist
        self.device_sync = device_sync
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.plan = StripPlan(int(size[0]), int(size[1]), self.world, int(halo_rows) if self.world > 1 else 0, lighting=bool(lighting))
        self.row_begin, self.row_end = self.plan.rows(self.rank)
        dev = torch.cuda.current_device() if device is None else device
        self.sim = Simulation(rules, size, lighting=bool(lighting), lit_strip=bool(lighting) and self.world > 1, device=dev, row_begin=self.row_begin, row_end=self.row_end,
                              halo_rows=self.plan.halo_rows, temporal_block=temporal_block, running_census=running_census,
                              device_share=device_share)
        if self.world > 1:
            mine = self.sim.ipc_export()
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine)
            if self.rank > 0:
                h, lr, gt, gb = everyone[self.rank - 1]
                self.sim.ipc_attach(0, h, lr, gt, gb)
            if self.rank < self.world - 1:
                h, lr, gt, gb = everyone[self.rank + 1]
                self.sim.ipc_attach(1, h, lr, gt, gb)
            if lighting:
                lights = [None] * self.world
                dist.all_gather_object(lights, self.sim.ipc_export_light())
                if self.rank > 0:
                    self.sim.ipc_attach_light(0, lights[self.rank - 1])
                if self.rank < self.world - 1:
                    self.sim.ipc_attach_light(1, lights[self.rank + 1])
            dist.barrier()

    @property
    def params(self):
        return self.sim.params

    def upload_cells(self, owned_rows) -> None:
        self.sim.upload_cells(owned_rows)       # marks the ghost rows stale: the next step exchanges first

    def download_cells(self, out=None):
        return self.sim.download_cells(out)

    def upload_light(self, owned_rows) -> None:
        self.sim.upload_light(owned_rows)

    def download_light(self):
        return self.sim.download_light()

    def exchange(self) -> None:
        """Explicit exchange (not needed around step()): everyone has finished computing -> push boundary rows into the
        neighbours' ghosts -> everyone has landed.  device_sync: the ordering is enforced by stream-ordered flag
        writes/waits on peer memory (no host in the loop); otherwise by two host barriers."""
        if self.world == 1:
            return
        if self.device_sync:
            self.sim.halo_exchange_async()
            return
        self.sim.synchronize()
        self.dist.barrier()
        self.sim.halo_push()
        self.sim.synchronize()
        self.dist.barrier()

    def step(self, n_steps: int) -> None:
        """Every rank calls this with the same n_steps; the library keeps the ghost rows current (see the module text)."""
        self.sim.step(int(n_steps))

    def census(self):
        return self.sim.census()

    def close(self):
        self.sim.close()
