#!/usr/bin/env python
"""bench.py -- Gcell-updates/s of the Margolus block update at 16384^2 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl ours|reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Definitions (DESIGN.md "Measurement"):
  step      one Margolus step (one `Simulation::run`, simulation.rs:195) over the whole S x S grid
  value     S*S*K / t, t = CUDA-event time of K steps issued through ONE se_sim_step(K) call, inputs resident in
            HBM, max over ranks, median of --reps repetitions; default rule set (data/materials.yaml), lighting off, no modifications,
            counter-hash initial grid seed 3, frames 2.. after W warm-up steps
  e2e       the same K steps through the public per-frame API with HOST buffers, every step: push the frame's
            modification record(s) from host memory (H2D), Simulation.run(), read the step's result -- the
            per-material census (256 x u64) -- back to the host (D2H); wall clock around the loop (sync both sides)
  roofline  algorithmic bytes = 8 B per cell-update (packed u32 in + out) / average launch duration of the
            dominant kernel, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the C restatement of the reference shader (oracle/, "port": the GLSL cannot run in this image)
            timed on the host cores over a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

SEED = 3
METRIC = "Gcell-updates/s at 16384^2 (Margolus block update, default rule set, lighting off)"
UNIT = "Gcell-updates/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=64)
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--halo", type=int, default=34, help="ghost rows per inner strip side; good for 2*(halo-1) steps")
    ap.add_argument("--temporal-block", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reps", type=int, default=5, help="timed repetitions of --steps steps; the median is reported")
    ap.add_argument("--no-running-census", action="store_true", help="e2e leg: recount the census every frame instead of SE_FLAG_RUNNING_CENSUS")
    return ap.parse_args()


def measured_peak():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  nvidia-smi needs
    ~100 ms to produce its first row, so sampling starts before the warm-up and every row is time-stamped; the
    rows inside [mark_begin, mark_end] are the ones reported (the same kernel runs during the warm-up, whose rows
    are used if the timed region was too short to catch one)."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def digest(rows):
            sm, mx, pw, reasons = [], [], [], set()
            for _, r in rows:
                parts = [p.strip() for p in r.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
                except ValueError:
                    continue
                for nm, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            sm.sort()
            return sm, mx, pw, reasons
        inside = [r for r in self.rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or 1e30)]
        window = "timed region"
        if not inside:
            inside = [r for r in self.rows if self.t1 is None or r[0] <= self.t1 + 0.05][-10:] or self.rows[-10:]
            window = "warm-up + timed region (timed region shorter than one sample)"
        sm, mx, pw, reasons = digest(inside)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def cpu_port_rate(size_sample: int, budget_s: float, seed: int = SEED, one_thread: bool = True):
    """Oracle (C restatement of the shader, OpenMP over all host cores) on a bounded sample of the workload."""
    import numpy as np

    from oracle.build_oracle import load_oracle
    from sandengine_b200.grids import synthetic_grid

    orc = load_oracle()
    g = synthetic_grid(size_sample, size_sample, seed)
    cores = len(os.sched_getaffinity(0))
    frame = orc.run_blocks(g, 1, 2)   # warm-up (page-in, thread pool) ...
    tw = time.perf_counter()
    while time.perf_counter() - tw < 1.0:   # ... and ~1 s of untimed work: host clocks / worker threads ramp up (seen: 0.07 -> 0.12 Gcell/s)
        frame = orc.run_blocks(g, frame, 2)
    steps, t0 = 0, time.perf_counter()
    while True:
        frame = orc.run_blocks(g, frame, 4)
        steps += 4
        dt = time.perf_counter() - t0
        if dt >= budget_s or steps >= 4000:
            break
    rate = size_sample * size_sample * steps / dt / 1e9
    # the same code on ONE thread (SURVEY.md 8d asks for it): 1024^2, ~2 s
    one = None
    try:
        if not one_thread:
            raise RuntimeError("skipped")
        import ctypes
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(1)
        g1 = synthetic_grid(1024, 1024, seed)
        f1 = orc.run_blocks(g1, 1, 2)
        n1, t1 = 0, time.perf_counter()
        while time.perf_counter() - t1 < 2.0:
            f1 = orc.run_blocks(g1, f1, 4)
            n1 += 4
        one = round(1024 * 1024 * n1 / (time.perf_counter() - t1) / 1e9, 4)
        gomp.omp_set_num_threads(cores)
    except Exception:
        pass
    return {"value": round(rate, 4), "unit": UNIT, "cores": cores, "kind": "port", "value_1_thread": one,
            "sample": f"{size_sample}x{size_sample} crop-sized grid of the same generator (seed {seed}), {steps} steps in {dt:.1f} s, "
                      f"C restatement of the reference shader (per-block form), gcc -O2 -fopenmp, {cores} threads; "
                      "llvmpipe GL 4.3 is not available in this image"}


def load_ref_shader():
    """The reference's own compute shader compiled for the CPU (oracle/build_ref.py -> oracle/_ref/).  Only the
    PREBUILT library is loaded (`__graft_entry__.build()` makes it in the build container and it travels with the
    snapshot): bench.py never reads the reference's sources.  None when it does not exist -- the callers then fall back
    to the oracle port and say so."""
    try:
        from oracle import build_ref
        so = build_ref.prebuilt_default()
        return build_ref.RefShader(so) if so is not None else None
    except Exception:
        return None


REF_WHAT = ("the reference's own compute shader (shaders/compute/falling_sand.glsl, text unmodified) compiled for the CPU through a "
            "GLSL-types shim (oracle/build_ref.py; g++ -O2 -march=x86-64-v3 -fopenmp, one invocation per cell like the GL "
            "dispatch; it always computes light and colour too -- the shader has no switch for them); Mesa llvmpipe GL 4.3 is "
            "not available in this image")


def cpu_reference_rate(size_sample: int, budget_s: float, seed: int = SEED):
    """oracle/_ref on a bounded sample of the workload, all host cores; None when oracle/_ref does not exist."""
    from sandengine_b200.grids import synthetic_grid

    ref = load_ref_shader()
    if ref is None:
        return None
    cores = len(os.sched_getaffinity(0))
    g = synthetic_grid(size_sample, size_sample, seed)
    ref.create(size_sample, size_sample)
    ref.upload_ids(g)
    ref.frame = 1
    ref.step(1)                                            # page-in, thread pool ...
    tw = time.perf_counter()
    while time.perf_counter() - tw < 1.0:                  # ... and ~1 s of untimed work (host clocks / worker threads ramp up)
        ref.step(1)
    steps, t0 = 0, time.perf_counter()
    while True:
        ref.step(1)
        steps += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or steps >= 4000:
            break
    rate = size_sample * size_sample * steps / dt / 1e9
    # the same library on ONE thread: 256^2 (BASELINE configs[0]), ~2 s
    one = None
    try:
        small = min(256, size_sample)
        ref.set_threads(1)
        ref.create(small, small)
        ref.upload_ids(synthetic_grid(small, small, seed))
        ref.frame = 1
        ref.step(1)
        n1, t1 = 0, time.perf_counter()
        while time.perf_counter() - t1 < 2.0:
            ref.step(1)
            n1 += 1
        one = round(small * small * n1 / (time.perf_counter() - t1) / 1e9, 6)
    except Exception:
        pass
    finally:
        ref.set_threads(cores)
    return {"value": round(rate, 6), "unit": UNIT, "cores": cores, "kind": "reference", "value_1_thread": one,
            "sample": f"{size_sample}x{size_sample} grid of the same generator (seed {seed}), {steps} steps in {dt:.1f} s, {cores} threads; " + REF_WHAT}


def run_reference(args):
    """Reference arm: the reference's own implementation of the path on the host cores -- its compute shader compiled
    for the CPU (oracle/_ref) when that exists, else the oracle port (no GL in this image either way)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each "step" of this arm is one Margolus step over a bounded sample of the workload; K + W of them.  The sample
    # is the largest power-of-two square whose projected run (probed with 2 steps) stays under ~2 minutes.
    from oracle.build_oracle import load_oracle
    from sandengine_b200.grids import synthetic_grid

    cores = len(os.sched_getaffinity(0))
    K, Wm = max(1, args.steps), max(0, args.warmup)
    ref = load_ref_shader()
    # torchrun exports OMP_NUM_THREADS=1 to its workers: this arm uses every host core it is allowed to run on whatever
    # the launcher says, so the N = 1 and the N > 1 lines measure the same baseline
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(cores)
    except Exception:
        pass
    if ref is not None:
        ref.set_threads(cores)
        kind = "reference"
        ladder = (2048, 1024, 512, 256, 128, 64)           # 256^2 is BASELINE configs[0], the reference's own CPU-sized case

        def start(g):
            ref.create(g.shape[1], g.shape[0]); ref.upload_ids(g); ref.frame = 1

        def steps(n):
            ref.step(n)
    else:
        kind = "port"
        ladder = (4096, 2048, 1024)
        orc = load_oracle()
        state = {}

        def start(g):
            state["g"], state["f"] = g.copy(), 1

        def steps(n):
            state["f"] = orc.run_blocks(state["g"], state["f"], n)
    for sample in ladder:
        sample = min(sample, args.size)
        g = synthetic_grid(sample, sample, SEED)
        start(g); steps(1)                                 # page-in, thread pool
        t0 = time.perf_counter()
        steps(2)
        if (time.perf_counter() - t0) / 2 * (K + Wm) <= 120.0 or sample == min(ladder[-1], args.size):
            break
    start(g)
    if Wm:
        steps(Wm)
    t0 = time.perf_counter()
    steps(K)
    dt = time.perf_counter() - t0
    rate = sample * sample * K / dt / 1e9
    what = REF_WHAT if kind == "reference" else ("C restatement of the reference shader (oracle/sand_oracle.c, per-block form), gcc -O2 -fopenmp; "
                                                 "oracle/_ref (the reference's shader compiled for the CPU) was not found; the reference's GLSL needs GL 4.3 "
                                                 "(llvmpipe) which this image does not have")
    cpu = {"value": round(rate, 6), "unit": UNIT, "cores": cores, "kind": kind, "sample": f"{sample}x{sample}, {K} steps, {cores} threads; " + what}
    if kind == "reference":                                # for context: the hand-optimised port of the same algorithm (ids only), ~3 s
        try:
            port = cpu_port_rate(min(args.size, 2048), 3.0, one_thread=False)
            cpu["port"] = {"value": port["value"], "unit": UNIT, "cores": port["cores"], "sample": port["sample"]}
        except Exception as e:  # noqa: BLE001
            cpu["port"] = {"error": repr(e)}
    line = {
        "impl": "reference", "metric": METRIC, "value": round(rate, 6), "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": Wm,
        "ms_per_step": round(dt / K * 1e3, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": f"bounded sample of the {args.size}^2 workload: {sample}x{sample} grid on {cores} host threads, same generator (seed {SEED}), "
                               "default rule set, no modifications" + (", lighting off" if kind == "port" else " (the shader relaxes light and shades colour every step)"),
                   "rules": "data/materials.yaml"},
        "cpu_baseline": cpu,
        "e2e": {"value": round(rate, 6), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def committed_checksum(S: int, total_steps: int):
    """Checksum (sandengine_b200.grids.grid_checksum) of the ORACLE's grid after `total_steps` steps of the bench workload,
    from tests/golden/bench_checksums.json (generator: tests/golden/make_bench_checksums.py); None when not committed."""
    p = REPO / "tests" / "golden" / "bench_checksums.json"
    try:
        return json.loads(p.read_text()).get(f"{S}x{S}:seed{SEED}:steps{total_steps}")
    except Exception:
        return None


def profile_traffic(S: int, world: int):
    """Physical DRAM bytes per cell-update of se_step_tiles from the committed ncu capture of this configuration
    (profiles/r2_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one launch / cell-updates of that launch)."""
    try:
        t = json.loads((REPO / "profiles" / "r2_traffic.json").read_text())
        return t.get(f"{S}:gpus{world}")
    except Exception:
        return None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: sandengine_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.gpus != world and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    import sandengine_b200 as se
    from sandengine_b200.distributed import StripSimulation
    from sandengine_b200.grids import synthetic_grid

    sampler = ClockSampler(local_rank)     # started early: nvidia-smi takes a few 100 ms to deliver its first row
    if rank == 0:
        sampler.start()
    S = args.size
    K, Wm, R = args.steps, max(args.warmup, 3), max(1, args.reps)
    rules = se.parse_path(REPO / "data" / "materials.yaml")
    strip = StripSimulation(rules, (S, S), halo_rows=args.halo, device=local_rank, temporal_block=args.temporal_block,
                            running_census=not args.no_running_census)
    sim = strip.sim
    stream = torch.cuda.Stream()          # a real (non-legacy) stream: events below are recorded on the launching stream
    torch.cuda.set_stream(stream)
    sim.set_stream(stream.cuda_stream)

    rows = strip.row_end - strip.row_begin
    host = torch.empty((rows, S), dtype=torch.int32).pin_memory()
    host_np = host.numpy().view(np.uint32)
    synthetic_grid(S, S, SEED, row_begin=strip.row_begin, row_end=strip.row_end, out=host_np)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reset():
        strip.upload_cells(host_np)
        sim.params.frame = 1

    def reduce_max(values):
        t = torch.tensor(values, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    # ---------------- device-resident timing (value): R repetitions of K steps, each timed on the device ----------------
    reset()
    strip.step(Wm)
    barrier()
    l0 = sim.launch_count
    reps = []
    sampler.mark_begin()
    for _ in range(R):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        strip.step(K)
        ev1.record(stream)
        barrier()
        reps.append(ev0.elapsed_time(ev1) / 1e3)
    sampler.mark_end()
    launches = (sim.launch_count - l0) / R
    print(f"[bench] rank {rank}: device times {[round(x * 1e3, 3) for x in reps]} ms for {K} steps each, rows {strip.row_begin}..{strip.row_end}", file=sys.stderr, flush=True)
    clocks = sampler.stop() if rank == 0 else None
    reps = reduce_max(reps)                              # per repetition: the slowest rank
    t_dev = sorted(reps)[len(reps) // 2]                 # median repetition
    value = S * S * K / t_dev / 1e9

    # ---------------- parity: sharding-independent checksum + census of the grid after W + R*K steps ----------------
    total_steps = Wm + R * K
    cs_local = sim.checksum()
    cs = torch.tensor([cs_local - (1 << 64) if cs_local >= (1 << 63) else cs_local], dtype=torch.int64, device="cuda")
    census = torch.from_numpy(sim.census().astype(np.int64)).cuda()
    if world > 1:
        dist.all_reduce(cs, op=dist.ReduceOp.SUM)        # wraps modulo 2^64: the strips' checksums add up to the grid's
        dist.all_reduce(census, op=dist.ReduceOp.SUM)
    checksum = int(cs.item()) & 0xFFFFFFFFFFFFFFFF
    want = committed_checksum(S, total_steps)
    if want is None:
        parity = {"status": "unchecked", "why": f"no committed oracle checksum for {S}x{S} after {total_steps} steps (tests/golden/bench_checksums.json)"}
    else:
        parity = {"status": "bit-exact" if int(want["checksum"]) == checksum else "MISMATCH",
                  "against": "oracle (C restatement pinned by the reference's own shader), full grid, committed checksum"}
        if "census" in want and [int(x) for x in want["census"]] != [int(x) for x in census.tolist()[:len(want["census"])]]:
            parity["status"] = "MISMATCH"
    parity.update({"checksum": f"{checksum:016x}", "after_steps": total_steps, "census": [int(x) for x in census.tolist()[:11]],
                   "how": "sum over cells of mix(global index, id) mod 2^64 (se_sim_checksum), summed over ranks: independent of the sharding"})

    # ---------------- end to end through the per-frame API with host buffers ----------------
    reset()
    terminator = np.zeros(1, dtype=se.MOD_DTYPE)   # the frame's modification UBO: "no modifications" (mod_size == 0)
    LAG = 4                                          # frames the host reads behind the device: the queue never drains
    census_ring = torch.zeros((2 * LAG, 256), dtype=torch.int64).pin_memory()   # results of the last frames (pinned, D2H target)
    landed = [torch.cuda.Event() for _ in range(2 * LAG)]
    census_log = np.zeros((K, 11), np.int64)
    for k in range(Wm):                              # the warm-up frames go through the same per-frame calls as the timed ones
        sim.push_modifications(terminator)
        strip.step(1)
        sim.census_async(census_ring[k % (2 * LAG)].data_ptr())
    sim.census_wait()
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        sim.push_modifications(terminator)           # H2D of the step's input (32 B record, staged through pinned memory)
        strip.step(1)                                # Simulation.run()
        sim.census_async(census_ring[k % (2 * LAG)].data_ptr())   # D2H of the step's result (256 x u64), in stream order
        landed[k % (2 * LAG)].record(stream)
        if k >= LAG:                                 # the result of frame k - LAG has landed (or is waited for) and is read on the host
            landed[(k - LAG) % (2 * LAG)].synchronize()
            census_log[k - LAG] = census_ring.numpy()[(k - LAG) % (2 * LAG), :11]
    for k in range(max(0, K - LAG), K):              # the last frames' results
        landed[k % (2 * LAG)].synchronize()
        census_log[k] = census_ring.numpy()[k % (2 * LAG), :11]
    sim.census_wait()
    barrier()
    t_e2e = reduce_max([time.perf_counter() - t0])[0]
    e2e_value = S * S * K / t_e2e / 1e9

    # whole job from a host grid to a host grid (upload + K steps + download), reported beside e2e
    barrier()
    t0 = time.perf_counter()
    sim.upload_cells_ptr(host.data_ptr())
    sim.params.frame = 1
    strip.step(K)
    sim.download_cells_ptr(host.data_ptr())            # download into the pinned buffer the grid was uploaded from
    barrier()
    t_job = reduce_max([time.perf_counter() - t0])[0]

    if rank == 0:
        peak, peak_src = measured_peak()
        # dominant kernel: se_step_tiles (K1b) fuses several Margolus steps per launch; algorithmic bytes of a
        # launch = 8 B x cells this rank owns x steps fused, i.e. 8 B x cells x K over one repetition
        local_cells = S * sim.owned_shape[0]
        per_launch_s = t_dev / max(launches, 1)
        achieved = 8.0 * local_cells * K / t_dev / 1e9
        tr = profile_traffic(S, world) if launches < K else None
        traffic = tr["dram_bytes_per_cell_update"] * local_cells * K / max(launches, 1) if tr else None
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": round(t_dev / K * 1e3, 5), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{S}x{S} grid, default rule set (data/materials.yaml), lighting off, no modifications, "
                                   f"counter-hash initial grid seed {SEED}, frames {2 + Wm}..{1 + Wm + R * K}",
                       "parallelism": f"strips{world}" + (f" halo{args.halo}, ghost rows pushed by the boundary tiles of se_step_tiles over NVLink (peer st.global + per-tile flags)" if world > 1 else ""),
                       "l2": "inputs larger than L2 (1 GiB cell buffer at 16384^2); no explicit flush" if S * S * 4 > 2 * 126e6 else
                             "L2-RESIDENT workload: the cell buffer fits the 126 MB L2",
                       "cells_bytes": S * S * 4},
            "timing": {"repetitions": R, "ms_per_repetition": [round(x * 1e3, 4) for x in reps], "reported": "median repetition (each repetition: max over ranks of the CUDA-event time of one se_sim_step(K) call)",
                       "timed_region_ms": round(sum(reps) * 1e3, 3)},
            "parity": parity,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "physical_frac": round(traffic / per_launch_s / 1e9 / peak, 4) if traffic else None,
                         "traffic_unit": ("bytes per launch = (dram__bytes_read.sum + dram__bytes_write.sum of the ncu --set full capture in " + tr["source"] +
                                          ") per cell-update x cell-updates of a timed launch") if tr else "null: no committed ncu capture of this configuration",
                         "kernel": "se_step_tiles" if launches < K else ("se_step_lut_global" if K == 1 else "se_step_inplace"), "peak_source": peak_src,
                         "algorithmic_bytes_per_cell_update": 8, "launches_per_repetition": launches,
                         "steps_per_launch": round(K / max(launches, 1), 2), "avg_launch_us": round(per_launch_s * 1e6, 2),
                         "note": "temporal blocking: physical DRAM bytes are ~8/T per cell-update (physical_frac), so the algorithmic "
                                 "fraction can exceed 1.0; the kernel is instruction-issue bound, not HBM bound"},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": 32, "d2h_bytes_per_step": 2048,
                    "what": "per frame: push modification record (H2D) + Simulation.run() + per-material census (D2H, asynchronous: "
                            "every frame's result is read on the host four frames behind the device, all of them inside the timed region)" + ("" if args.no_running_census else "; census kept up to date by the step kernel (SE_FLAG_RUNNING_CENSUS)"),
                    "job_roundtrip": {"value": round(S * S * K / t_job / 1e9, 2), "unit": UNIT,
                                      "what": f"upload grid from pinned host memory + {K} steps + download grid",
                                      "h2d_bytes": S * rows * 4, "d2h_bytes": S * rows * 4}},
            "gpu_launches": int(round(launches * R)),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:      # the CPU baseline is reported at N = 1 only
            try:
                # the reference's own shader compiled for the CPU when oracle/_ref exists ("reference"), with the
                # hand-optimised port of the same algorithm beside it; the port alone otherwise
                refcpu = cpu_reference_rate(min(S, 1024), args.cpu_seconds)
                port = cpu_port_rate(min(S, 4096), args.cpu_seconds if refcpu is None else min(args.cpu_seconds, 6.0))
                if refcpu is not None:
                    refcpu["port"] = port
                    line["cpu_baseline"] = refcpu
                else:
                    line["cpu_baseline"] = port
            except Exception as e:   # the baseline is a reported number, never a dependency of the GPU path
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": len(os.sched_getaffinity(0)), "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    strip.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    # Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON line on
    # stdout, so everything else is routed to stderr and the line is written to the real stdout.
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _real_stdout
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
    _real_stdout.flush()
