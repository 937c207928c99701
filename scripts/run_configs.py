"""Secondary BASELINE.json configs (not bench lines): timing + parity spot checks, one JSON line each on stderr/stdout.
  python scripts/run_configs.py 1        # 256^2, 1000 steps (BASELINE configs[0]): timing + comparison with the reference shader's output
  python scripts/run_configs.py 2        # 4096^2, lighting off, 10k steps (L2-resident)
  python scripts/run_configs.py 4        # 4096^2, brush stamps + explosion every frame, lighting on
  torchrun ... scripts/run_configs.py 5  # synthetic 64-material rule set, 65536^2 over the ranks (table in global memory; env SE_LUT_FORCE_MODE=0: generated code)
  SE_CFG5_SIZE=16384 python scripts/run_configs.py 5      # the same rule set on one GPU
"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import sandengine_b200 as se  # noqa: E402
from sandengine_b200.grids import hashi, synthetic_grid  # noqa: E402


def frame_mods(k, w, h, n_mats_selectable):
    """configs[3]: per frame 4 brush stamps (CIRCLE / SQUARE, size 3..32) + 1 explosion (CIRCLE of EMPTY, size 16..64)."""
    m = np.zeros(5, se.MOD_DTYPE)
    hv = hashi(np.arange(k * 16, k * 16 + 16, dtype=np.uint32))
    for i in range(4):
        m[i]["position"] = (int(hv[3 * i] % w), int(hv[3 * i + 1] % h))
        m[i]["mod_shape"] = int(hv[3 * i + 2] & 1)
        m[i]["mod_size"] = 3 + int((hv[3 * i + 2] >> 1) % 30)
        m[i]["mod_matID"] = int(n_mats_selectable[int((hv[3 * i + 2] >> 8) % len(n_mats_selectable))])
    m[4]["position"] = (int(hv[12] % w), int(hv[13] % h))
    m[4]["mod_shape"] = 0
    m[4]["mod_size"] = 16 + int(hv[14] % 49)
    m[4]["mod_matID"] = 0
    return m


def config1():
    """configs[0]: 256 x 256, default rule set, 1000 steps, seed 1 -- the reference's own CPU-sized case.  Lighting on (the
    reference's shader always relaxes light); final ids are compared with the output of the reference's shader compiled
    for the CPU (tests/golden/ref_shader_goldens.json), light with its sub-sampled copy; timing per frame and batched."""
    import hashlib
    S, K = 256, 1000
    gold = json.loads((REPO / "tests" / "golden" / "ref_shader_goldens.json").read_text())["cases"]["default_256x256_seed1_lit_1000"]
    gold_light = np.load(REPO / "tests" / "golden" / "ref_shader_arrays.npz")["default_256x256_seed1_lit_1000/light_sub8"]
    rules = se.parse_path(REPO / "data" / "materials.yaml")
    out = {"config": 1, "workload": "256x256, default rules, 1000 steps, seed 1 (BASELINE configs[0])"}
    for lighting in (True, False):
        sim = se.Simulation(rules, (S, S), lighting=lighting)
        st = torch.cuda.Stream(); torch.cuda.set_stream(st); sim.set_stream(st.cuda_stream)
        g = synthetic_grid(S, S, 1)
        for batched in (False, True):
            sim.upload_cells(g); sim.params.frame = 1
            if lighting: sim.upload_light(np.zeros((S, S, 4), np.float32))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(st)
            if batched: sim.step(K)
            else:
                for _ in range(K): sim.run()
            e1.record(st); torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / 1e3
            ids = sim.download_cells()
            ok = hashlib.sha256(np.ascontiguousarray(ids, np.uint32).tobytes()).hexdigest() == gold["ids_sha256"]["1000"]
            key = ("lit" if lighting else "unlit") + ("_batched" if batched else "_per_frame")
            out[key] = {"gcell_per_s": round(S * S * K / t / 1e9, 3), "us_per_step": round(t / K * 1e6, 2), "ids_equal_reference_shader": bool(ok)}
            if lighting:
                out[key]["light_max_abs_err_vs_reference_shader"] = float(np.abs(sim.download_light()[::8, ::8] - gold_light).max())
        sim.close()
    out["note"] = "launch-latency bound at this size (65 536 cells): us_per_step is the figure of merit, not Gcell/s"
    print(json.dumps(out), flush=True)


def config2():
    S, K = 4096, 10000
    rules = se.parse_path(REPO / "data" / "materials.yaml")
    sim = se.Simulation(rules, (S, S))
    st = torch.cuda.Stream(); torch.cuda.set_stream(st); sim.set_stream(st.cuda_stream)
    sim.upload_cells(synthetic_grid(S, S, 2)); sim.params.frame = 1
    sim.step(64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st); sim.step(K); e1.record(st); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 1e3
    census = sim.census()
    print(json.dumps({"config": 2, "workload": "4096x4096, default rules, lighting off, 10000 steps, seed 2", "gcell_per_s": round(S * S * K / t / 1e9, 1),
                      "ms_per_step": round(t / K * 1e3, 5), "note": "L2-RESIDENT: the 64 MiB cell buffers fit the 126 MB L2",
                      "census_after": [int(c) for c in census[:11]]}), flush=True)


def config4():
    S, K = 4096, 200
    rules = se.parse_path(REPO / "data" / "materials.yaml")
    sel = [m.id for m in rules.selectable_materials]
    sim = se.Simulation(rules, (S, S), lighting=True)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st); sim.set_stream(st.cuda_stream)
    g = synthetic_grid(S, S, 4)
    sim.upload_cells(g); sim.upload_light(np.zeros((S, S, 4), np.float32)); sim.params.frame = 1
    mods = [frame_mods(k, S, S, sel) for k in range(K + 8)]
    for k in range(8):
        sim.push_modifications(mods[k]); sim.run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(8, K + 8):
        sim.push_modifications(mods[k]); sim.run()
    sim.synchronize()
    t = time.perf_counter() - t0
    out = {"config": 4, "workload": "4096x4096, default rules, 4 stamps + 1 explosion every frame, lighting on, seed 4", "steps": K,
           "gcell_per_s": round(S * S * K / t / 1e9, 1), "ms_per_step": round(t / K * 1e3, 4),
           "roofline_frac_40B": round(40.0 * S * S * K / t / 1e9 / 6549.8, 4)}
    # parity spot check against the oracle on a 512x512 crop-sized run of the same generator
    from oracle.build_oracle import load_oracle
    s = 512
    g2 = synthetic_grid(s, s, 4)
    sim2 = se.Simulation(rules, (s, s), lighting=True)
    sim2.upload_cells(g2); sim2.upload_light(np.zeros((s, s, 4), np.float32)); sim2.params.frame = 1
    m2 = [frame_mods(k, s, s, sel) for k in range(60)]
    for k in range(60):
        sim2.push_modifications(m2[k]); sim2.run()
    ref, refL, _ = load_oracle().run(g2, 1, 60, light=np.zeros((s, s, 4), np.float32), mods_per_step=m2)
    out["parity_512_60steps"] = {"ids_bit_exact": bool(np.array_equal(sim2.download_cells(), ref)),
                                 "light_max_abs_err": float(np.abs(sim2.download_light() - refL).max())}
    print(json.dumps(out), flush=True)


def config5():
    import torch.distributed as dist
    from sandengine_b200.distributed import StripSimulation
    from sandengine_b200.synth_rules import synthetic_rule_set
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    S = int(os.environ.get("SE_CFG5_SIZE", 65536))
    K = 64
    text, ids, mix = synthetic_rule_set(64, 28, seed=5)
    rules = se.parse_string(text)
    mode = int(rules.cuda_header.split("#define SE_LUT_MODE ")[1].split()[0])
    strip = StripSimulation(rules, (S, S), halo_rows=34, device=local)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st); strip.sim.set_stream(st.cuda_stream)
    rows = strip.row_end - strip.row_begin
    g = synthetic_grid(S, S, 5, mix=mix, ids=ids, row_begin=strip.row_begin, row_end=strip.row_end)
    strip.upload_cells(g); strip.params.frame = 1
    strip.step(32)
    times = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        e0.record(st); strip.step(K); e1.record(st); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / 1e3)
    t = torch.tensor(times, dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = t.median()
    census = torch.from_numpy(strip.census().astype(np.int64)).cuda()
    if world > 1: dist.all_reduce(census)
    if rank == 0:
        tt = float(t.item())
        print(json.dumps({"config": 5, "workload": f"synthetic 64-material rule set (28 rules, LEFT/RIGHT, depth-6 types), {S}x{S} over {world} GPU(s), seed 5",
                          "kernel": {0: "se_step_inplace (generated code, one step per launch)", 1: "se_step_tiles (table in shared memory)",
                                     2: "se_step_tiles (one transition table per view in global memory: 2 x 64^4 x 4 B = 134 MB, L2 / HBM resident)"}[mode],
                          "steps": K, "repetitions": 3, "gcell_per_s": round(S * S * K / tt / 1e9, 1),
                          "ms_per_step": round(tt / K * 1e3, 4), "roofline_frac_8B_per_gpu": round(8.0 * S * S * K / tt / 1e9 / 6549.8 / world, 4),
                          "cells_total": int(census.sum().item()), "cells_expected": S * S}), flush=True)
    strip.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    {"1": config1, "2": config2, "4": config4, "5": config5}[sys.argv[1]]()
