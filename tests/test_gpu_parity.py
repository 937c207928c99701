"""Parity of the CUDA path (through the C ABI) against the oracle -- the gate for every kernel.
Bit-exact for cell state; light within 1e-6 absolute (f32, same expression order, no FMA)."""
import hashlib
import os

import numpy as np
import pytest

import yaml_cases as Y
from conftest import DEFAULT_YAML
from sandengine_b200.grids import kat_grid, synthetic_grid

pytestmark = pytest.mark.gpu

LIGHT_ATOL = 1e-6   # recommended tolerance for the f32 lighting values


def sha16(g):
    return hashlib.sha256(g.astype(np.uint8).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def se(native_lib):
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import sandengine_b200
    return sandengine_b200


@pytest.fixture(scope="module")
def rich(se):
    from oracle.build_oracle import load_oracle
    return se.parse_string(Y.RICH_YAML), load_oracle(Y.RICH_YAML)


def run_gpu(se, rules, grid, n_steps, frame0=1, lighting=False, light0=None, mods_per_step=None, chunk=None, temporal_block=0):
    H, W = grid.shape
    sim = se.Simulation(rules, (W, H), lighting=lighting, temporal_block=temporal_block)
    sim.upload_cells(grid)
    if lighting and light0 is not None:
        sim.upload_light(light0)
    sim.params.frame = frame0
    if mods_per_step is None:
        if chunk is None:
            sim.step(n_steps)
        else:
            done = 0
            while done < n_steps:
                k = min(chunk, n_steps - done)
                sim.step(k)
                done += k
    else:
        for s in range(n_steps):
            if mods_per_step[s] is not None and len(mods_per_step[s]):
                sim.push_modifications(mods_per_step[s])
            sim.run()
    out = sim.download_cells()
    light = sim.download_light() if lighting else None
    frame = sim.params.frame
    sim.close()
    return out, light, frame


def test_survey_state_kats(se, default_rules):
    for n, steps, final in [(16, 40, "e4ab9d8c55d01623"), (32, 100, "d2f2ab20d7d3a8bc"), (20, 60, "298014800d6a9b1b")]:
        c, _, f = run_gpu(se, default_rules, kat_grid(n), steps)
        assert f == 1 + steps and sha16(c) == final


@pytest.mark.parametrize("w,h,seed,steps", [(256, 256, 1, 1000), (64, 48, 2, 200), (33, 21, 7, 120), (1, 9, 3, 40), (9, 1, 4, 40),
                                            (2, 2, 5, 16), (130, 66, 11, 257), (515, 77, 13, 64)])
def test_default_rules_vs_oracle(se, default_rules, oracle, w, h, seed, steps):
    """configs[0]: 256x256, default rule set, 1000 steps, fixed seed -- plus ragged / degenerate sizes."""
    g = synthetic_grid(w, h, seed)
    ref, _, _ = oracle.run(g, 1, steps, blocks=True)
    got, _, _ = run_gpu(se, default_rules, g, steps)
    assert np.array_equal(got, ref)


def test_every_intermediate_step_256(se, default_rules, oracle):
    g = synthetic_grid(256, 256, 1)
    sim = se.Simulation(default_rules, (256, 256))
    sim.upload_cells(g)
    sim.params.frame = 1
    ref = g.copy()
    frame = 1
    for k in range(100):
        sim.run()
        frame = oracle.run_blocks(ref, frame, 1)
        assert np.array_equal(sim.download_cells(), ref), f"diverged at step {k + 1}"
    sim.close()


def test_frame_phase_alignment(se, default_rules, oracle):
    """Start frames covering all four Margolus offsets, and a large frame number."""
    g = synthetic_grid(96, 80, 21)
    for f0 in (1, 2, 3, 4, 5, 1000, 1_000_003):
        ref, _, _ = oracle.run(g, f0, 9, blocks=True)
        got, _, f = run_gpu(se, default_rules, g, 9, frame0=f0)
        assert f == f0 + 9 and np.array_equal(got, ref)


def test_frame1_clears(se, default_rules, oracle):
    g = synthetic_grid(64, 64, 3)
    got, _, f = run_gpu(se, default_rules, g, 1, frame0=0)
    assert f == 1 and not got.any()
    got, _, f = run_gpu(se, default_rules, g, 5, frame0=0)
    assert f == 5 and not got.any()


def test_unknown_ids(se, default_rules, oracle):
    g = synthetic_grid(64, 64, 5)
    g[10, 10] = 11; g[11, 20] = 200; g[30, 31] = 255; g[40, 40] = 70000; g[41, 41] = 0xFFFFFFFF
    ref, _, _ = oracle.run(g, 1, 20)
    got, _, _ = run_gpu(se, default_rules, g, 20)
    assert np.array_equal(got, ref)


def test_rich_rules_left_right_vs_oracle(se, rich):
    rules, orc = rich
    for (w, h, seed, steps) in [(128, 96, 5, 300), (61, 47, 6, 150)]:
        g = synthetic_grid(w, h, seed, mix=Y.RICH_MIX, ids=Y.RICH_IDS)
        ref, _, _ = orc.run(g, 1, steps, blocks=True)
        got, _, _ = run_gpu(se, rules, g, steps)
        assert np.array_equal(got, ref)
        assert len(np.unique(ref)) > 5


def make_mods(se, frame, n_mats, w, h, rng):
    from sandengine_b200 import MOD_DTYPE
    k = int(rng.integers(0, 7))
    m = np.zeros(k, MOD_DTYPE)
    for i in range(k):
        m[i]["position"] = (int(rng.integers(-4, w + 4)), int(rng.integers(-4, h + 4)))
        m[i]["mod_shape"] = int(rng.integers(0, 2))
        m[i]["mod_size"] = int(rng.integers(1, 12))
        m[i]["mod_matID"] = int(rng.integers(0, n_mats))
    return m


def test_modifications_vs_oracle(se, default_rules, oracle):
    """configs[3] semantics at small size: stamps + explosions every frame, incl. size-0 terminator and unknown ids."""
    from sandengine_b200 import MOD_DTYPE
    rng = np.random.default_rng(42)
    w, h, steps = 96, 72, 60
    g = synthetic_grid(w, h, 4)
    mods = [make_mods(se, s, 11, w, h, rng) for s in range(steps)]
    mods[3] = np.zeros(4, MOD_DTYPE)
    mods[3][0] = ((20, 20), 0, 6, 3, (0, 0, 0)); mods[3][1] = ((30, 30), 1, 0, 4, (0, 0, 0)); mods[3][2] = ((40, 40), 1, 5, 4, (0, 0, 0))
    mods[5] = np.zeros(2, MOD_DTYPE)
    mods[5][0] = ((50, 20), 0, 6, 4, (0, 0, 0)); mods[5][1] = ((50, 20), 0, 3, 77, (0, 0, 0))   # unknown id cancels the centre
    mods[7] = np.zeros(300, MOD_DTYPE)   # more than 256: extras silently dropped (simulation.rs:205)
    for i in range(300):
        mods[7][i] = ((i % w, (i * 7) % h), i % 2, 1, 3 + (i % 8), (0, 0, 0))
    ref, _, _ = oracle.run(g, 1, steps, mods_per_step=mods)
    got, _, _ = run_gpu(se, default_rules, g, steps, mods_per_step=mods)
    assert np.array_equal(got, ref)


def test_lighting_vs_oracle(se, default_rules, oracle):
    g = kat_grid(16); g[5][5] = 6; g[9][12] = 8
    ref, refL, _ = oracle.run(g, 1, 40, light=np.zeros((16, 16, 4), np.float32))
    got, gotL, _ = run_gpu(se, default_rules, g, 40, lighting=True, light0=np.zeros((16, 16, 4), np.float32))
    assert sha16(got) == "ca0e8b7d36edd6b2" and np.array_equal(got, ref)
    assert np.abs(gotL - refL).max() <= LIGHT_ATOL
    # larger, random light, with modifications
    rng = np.random.default_rng(7)
    w, h, steps = 80, 64, 50
    g = synthetic_grid(w, h, 9)
    L0 = rng.random((h, w, 4), dtype=np.float32)
    mods = [make_mods(se, s, 11, w, h, rng) for s in range(steps)]
    ref, refL, _ = oracle.run(g, 1, steps, light=L0, mods_per_step=mods)
    got, gotL, _ = run_gpu(se, default_rules, g, steps, lighting=True, light0=L0, mods_per_step=mods)
    assert np.array_equal(got, ref)
    err = np.abs(gotL - refL).max()
    assert err <= LIGHT_ATOL, err
    # large enough for se_light's interior CTAs (32 x 32 tiles whose ring is inside the grid) next to rim CTAs
    w, h, steps = 200, 150, 12
    g = synthetic_grid(w, h, 10)
    L0 = rng.random((h, w, 4), dtype=np.float32)
    L0[rng.random((h, w)) < 0.2, 3] = 0.0
    ref, refL, _ = oracle.run(g, 1, steps, light=L0)
    got, gotL, _ = run_gpu(se, default_rules, g, steps, lighting=True, light0=L0)
    assert np.array_equal(got, ref)
    err = np.abs(gotL - refL).max()
    assert err <= LIGHT_ATOL, err


def test_lighting_frame1(se, default_rules, oracle):
    g = synthetic_grid(32, 32, 2)
    L0 = np.random.default_rng(1).random((32, 32, 4), dtype=np.float32)
    ref, refL, _ = oracle.run(g, 0, 3, light=L0)
    got, gotL, _ = run_gpu(se, default_rules, g, 3, frame0=0, lighting=True, light0=L0)
    assert np.array_equal(got, ref) and np.abs(gotL - refL).max() <= LIGHT_ATOL


def test_4096_short_horizon_vs_oracle(se, default_rules, oracle):
    """configs[1] size (4096^2): oracle comparison at steps 1..5, 32 and 100."""
    g = synthetic_grid(4096, 4096, 2)
    sim = se.Simulation(default_rules, (4096, 4096))
    sim.upload_cells(g)
    sim.params.frame = 1
    ref = g.copy()
    frame = 1
    for target in (1, 2, 3, 4, 5, 32, 100):
        k = target - (frame - 1)
        sim.step(k)
        frame = oracle.run_blocks(ref, frame, k)
        assert np.array_equal(sim.download_cells(), ref), f"step {target}"
    sim.close()


def test_census_and_conservation_properties(se, default_rules):
    """Size-independent properties at a BASELINE-size grid (16384 x 2048 strip-sized): census equals a host
    bincount, and materials that no rule creates or destroys (rock, radioactive, toxic_sludge, sand, dirt, water) are conserved."""
    w, h = 16384, 2048
    g = synthetic_grid(w, h, 3)
    sim = se.Simulation(default_rules, (w, h))
    sim.upload_cells(g)
    sim.params.frame = 1
    c0 = sim.census()
    assert np.array_equal(c0[:11], np.bincount(g.ravel(), minlength=11).astype(np.uint64))
    sim.step(200)
    c1 = sim.census()
    out = sim.download_cells()
    assert np.array_equal(c1[:11], np.bincount(out.ravel(), minlength=11).astype(np.uint64))
    assert c1.sum() == w * h
    for mat in (4, 6, 8, 5, 10):       # rock, radioactive, toxic_sludge, water, dirt: only ever swapped
        assert c1[mat] == c0[mat]
    assert c1[3] == c0[3]               # sand
    assert c1[1] == 0 and c1[2] == 0    # no NULL / WALL leaks into the grid
    assert c1[7] <= c0[7]               # smoke only dissolves
    sim.close()


def test_step_chunking_is_equivalent(se, default_rules):
    g = synthetic_grid(300, 200, 8)
    a, _, _ = run_gpu(se, default_rules, g, 64)
    b, _, _ = run_gpu(se, default_rules, g, 64, chunk=1)
    c, _, _ = run_gpu(se, default_rules, g, 64, chunk=7)
    assert np.array_equal(a, b) and np.array_equal(a, c)


def test_synthetic_64_material_rules_1024(se):
    """configs[4] rule set at crop size (1024^2): 64 materials, deep inheritance, LEFT and RIGHT rules."""
    from oracle.build_oracle import load_oracle
    from sandengine_b200.synth_rules import synthetic_rule_set
    text, ids, mix = synthetic_rule_set(64, 28, seed=5)
    rules = se.parse_string(text)
    orc = load_oracle(text)
    g = synthetic_grid(1024, 1024, 5, mix=mix, ids=ids)
    ref, _, _ = orc.run(g, 1, 100, blocks=True)
    got, _, _ = run_gpu(se, rules, g, 100)
    assert np.array_equal(got, ref) and not np.array_equal(got, g)


def _strip_pair_run(se, rules, g, steps, halo, n_strips, device_sync=False, per_step=False, auto=False):
    """n strips of one grid on ONE device in ONE process (se_sim_attach_local): exercises ghost rows, the
    missing-row logic and se_sim_halo_push without torch.distributed."""
    from sandengine_b200.distributed import StripPlan
    H, W = g.shape
    plan = StripPlan(W, H, n_strips, halo)
    sims = []
    for r in range(n_strips):
        b, e = plan.rows(r)
        s = se.Simulation(rules, (W, H), row_begin=b, row_end=e, halo_rows=halo, device_share=n_strips)
        s.upload_cells(g[b:e])
        s.params.frame = 1
        sims.append(s)
    for r, s in enumerate(sims):
        if r > 0:
            s.attach_local(0, sims[r - 1])
        if r < n_strips - 1:
            s.attach_local(1, sims[r + 1])

    def exchange():
        if device_sync:      # stream-ordered flags on (here: same-device) peer memory, no host synchronisation
            for s in sims: s.halo_exchange_async()
            return
        for s in sims: s.synchronize()
        for s in sims: s.halo_push()
        for s in sims: s.synchronize()
    if auto:                # no explicit exchange at all: se_sim_step keeps the ghost rows current (fused push / stream exchange)
        # runs of steps (tile kernel, fused push), single steps (K1c / K1a, stream-ordered exchange) and mixtures of both:
        # a per-step kernel right after a run must find the neighbours' last push complete
        sched = [1] * steps if per_step else [steps // 3, 1, 1, steps - steps // 3 - 3, 1]
        for k in sched:
            for s in sims: s.step(k)
    else:
        exchange()
        for k in plan.chunks(steps):
            for s in sims:
                if per_step:            # one se_sim_step(1) per frame: the K1c (single-step table kernel) path
                    for _ in range(k): s.step(1)
                else:
                    s.step(k)
            exchange()
    out = np.concatenate([s.download_cells() for s in sims], axis=0)
    for s in sims: s.close()
    return out


@pytest.mark.parametrize("n_strips,halo,h", [(2, 4, 64), (3, 8, 100), (4, 2, 64), (2, 16, 130)])
def test_strips_equal_single_grid(se, default_rules, oracle, n_strips, halo, h):
    g = synthetic_grid(96, h, 17)
    steps = 45
    ref, _, _ = oracle.run(g, 1, steps, blocks=True)
    got = _strip_pair_run(se, default_rules, g, steps, halo, n_strips)
    assert np.array_equal(got, ref)
    got = _strip_pair_run(se, default_rules, g, steps, halo, n_strips, device_sync=True)
    assert np.array_equal(got, ref)
    got = _strip_pair_run(se, default_rules, g, steps, halo, n_strips, device_sync=True, per_step=True)
    assert np.array_equal(got, ref)
    got = _strip_pair_run(se, default_rules, g, steps, halo, n_strips, auto=True)
    assert np.array_equal(got, ref)
    got = _strip_pair_run(se, default_rules, g, steps, halo, n_strips, auto=True, per_step=True)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("T", [1, 2, 4, 6, 8, 16])
def test_tiled_table_kernel_vs_oracle(se, default_rules, oracle, T):
    """K1b (transition table + shared-memory tiles, T fused steps per launch) against the oracle; T = 1 is K1a.
    Sizes straddle tile boundaries (tile 256 wide) and include WALL / NULL cells (slow path inside tiles)."""
    for (w, h, seed, steps) in [(512, 300, 31, 97), (260, 1000, 32, 50), (1024, 64, 33, 64)]:
        g = synthetic_grid(w, h, seed)
        rng = np.random.default_rng(seed)
        g[rng.integers(0, h, 200), rng.integers(0, w, 200)] = 2
        g[rng.integers(0, h, 50), rng.integers(0, w, 50)] = 1
        g[rng.integers(0, h, 50), rng.integers(0, w, 50)] = 9
        ref, _, _ = oracle.run(g, 1, steps, blocks=True)
        got, _, f = run_gpu(se, default_rules, g, steps, temporal_block=T)
        assert f == 1 + steps
        assert np.array_equal(got, ref), f"T={T} size {w}x{h}"


def test_tiled_with_modifications_interleaved(se, default_rules, oracle):
    """Steps with modifications fall back to K1a for that step and return to the tiled kernel afterwards."""
    rng = np.random.default_rng(5)
    w, h, steps = 512, 256, 40
    g = synthetic_grid(w, h, 12)
    mods = [make_mods(se, s, 11, w, h, rng) if s % 3 == 0 else None for s in range(steps)]
    ref, _, _ = oracle.run(g, 1, steps, mods_per_step=[m if m is not None else np.zeros(0, se.MOD_DTYPE) for m in mods])
    sim = se.Simulation(default_rules, (w, h))
    sim.upload_cells(g)
    sim.params.frame = 1
    s = 0
    while s < steps:
        if mods[s] is not None and len(mods[s]):
            sim.push_modifications(mods[s])
        run = 1
        while s + run < steps and mods[s + run] is None:
            run += 1
        sim.step(run)      # first step consumes the modifications, the rest are plain (tiled)
        s += run
    assert np.array_equal(sim.download_cells(), ref)
    sim.close()


def test_tiled_4096_vs_per_step_kernel(se, default_rules):
    """configs[1] size: the tiled kernel and the per-step kernel agree bit for bit after 256 steps at 4096^2."""
    g = synthetic_grid(4096, 4096, 2)
    a, _, _ = run_gpu(se, default_rules, g, 256, temporal_block=1)
    b, _, _ = run_gpu(se, default_rules, g, 256, temporal_block=4)
    c, _, _ = run_gpu(se, default_rules, g, 256, temporal_block=8, chunk=37)
    assert np.array_equal(a, b) and np.array_equal(a, c)


def test_census_async_overlaps_steps(se, default_rules):
    """se_sim_census_async: results equal a host bincount of the state at the time of the call, also when later
    steps (which ping-pong / overwrite buffers) are enqueued before the wait."""
    import torch
    w, h = 1024, 768
    g = synthetic_grid(w, h, 19)
    for T in (0, 1):      # tiled (ping-pong) and per-step in-place kernels
        sim = se.Simulation(default_rules, (w, h), temporal_block=T)
        sim.upload_cells(g); sim.params.frame = 1
        ring = torch.zeros((4, 256), dtype=torch.int64).pin_memory()
        expect = []
        for k in range(4):
            sim.step(3)
            expect.append(np.bincount(sim.download_cells().ravel(), minlength=256))
            sim.census_async(ring[k].data_ptr())
            sim.step(2)                      # enqueued before the census is consumed
            sim.step(1)
        sim.census_wait()
        for k in range(4):
            assert np.array_equal(ring[k].numpy(), expect[k]), (T, k)
        sim.close()


def test_single_step_table_kernel_k1c(se, default_rules, oracle):
    """K1c (se_step_lut_global): every frame issued as its own se_sim_step(1), incl. odd widths that disable the
    8-byte path, WALL/NULL cells and all four Margolus phases."""
    for (w, h, seed, steps) in [(516, 130, 41, 41), (1024, 64, 42, 24), (260, 258, 43, 19)]:
        g = synthetic_grid(w, h, seed)
        rng = np.random.default_rng(seed)
        g[rng.integers(0, h, 100), rng.integers(0, w, 100)] = 2
        g[rng.integers(0, h, 30), rng.integers(0, w, 30)] = 1
        ref, _, _ = oracle.run(g, 1, steps, blocks=True)
        got, _, _ = run_gpu(se, default_rules, g, steps, chunk=1)
        assert np.array_equal(got, ref), (w, h)


def test_colour_shading(se, default_rules, oracle):
    """K5 (next row L3): colour = f(material, position) with simplex noise built on fract(sin(p) * 43758.5453).
    That hash amplifies last-bit differences between CUDA sinf and glibc sinf by ~4e4 and is discontinuous, so parity
    is stated as a tolerance on the distribution: EMPTY cells exact, >= 99 % of channels within 2e-3 of the CPU
    restatement, alpha exact (noise never touches it)."""
    w, h = 192, 160
    g = synthetic_grid(w, h, 23)
    _, _, ref = oracle.step_cells(g, 2, want_color=True)          # colour of the cells AFTER frame 2
    after, _, _ = oracle.run(g, 1, 1)
    sim = se.Simulation(default_rules, (w, h))
    sim.upload_cells(g); sim.params.frame = 1
    sim.run()
    assert np.array_equal(sim.download_cells(), after)
    col = sim.download_color()
    err = np.abs(col - ref)
    assert np.array_equal(col[..., 3], ref[..., 3])
    assert err[after == 0].max() == 0.0
    frac_ok = float((err[..., :3] <= 2e-3).mean())
    assert frac_ok >= 0.99, (frac_ok, float(err.max()), float(np.median(err)))
    rgba8 = sim.download_color(rgba8=True)
    expect8 = np.rint(np.clip(col, 0, 1) * 255).astype(np.uint32)
    packed = expect8[..., 0] | (expect8[..., 1] << 8) | (expect8[..., 2] << 16) | (expect8[..., 3] << 24)
    assert np.array_equal(rgba8, packed)
    sim.close()


def test_snapshot_resume_is_bit_identical(se, default_rules, tmp_path):
    from sandengine_b200 import snapshot
    g = synthetic_grid(256, 128, 29)
    L0 = np.random.default_rng(2).random((128, 256, 4), dtype=np.float32)
    for lighting in (False, True):
        a = se.Simulation(default_rules, (256, 128), lighting=lighting)
        a.upload_cells(g); a.params.frame = 1
        if lighting: a.upload_light(L0)
        a.step(23)
        snapshot.save(a, tmp_path / "s.npz")
        a.step(40)
        b = snapshot.load(default_rules, tmp_path / "s.npz")
        assert b.params.frame == 24
        b.step(40)
        assert np.array_equal(a.download_cells(), b.download_cells())
        if lighting:
            assert np.array_equal(a.download_light(), b.download_light())
        a.close(); b.close()
    other = se.parse_string(Y.RICH_YAML)
    with pytest.raises(ValueError):
        snapshot.load(other, tmp_path / "s.npz")


def test_expression_rule_set_on_gpu(se):
    """Rule conditions with arithmetic, pos, frame, rand.z / rand.w, vector components (tests/yaml_cases.EXPR_YAML):
    generic generated-code kernel (K1a) against the oracle."""
    from oracle.build_oracle import load_oracle
    rules = se.parse_string(Y.EXPR_YAML)
    orc = load_oracle(Y.EXPR_YAML)
    for (w, h, seed, steps, f0) in [(256, 192, 5, 120, 1), (67, 41, 6, 60, 1001)]:
        g = synthetic_grid(w, h, seed, mix=Y.EXPR_MIX, ids=Y.EXPR_IDS)
        ref, _, _ = orc.run(g, f0, steps, blocks=True)
        got, _, _ = run_gpu(se, rules, g, steps, frame0=f0)
        assert np.array_equal(got, ref) and not np.array_equal(got, g)


def test_16384_full_grid_short_horizon_vs_oracle(se, default_rules, oracle):
    """BASELINE size (16384^2, configs[2]): the full grid against the oracle after 16 steps (two T-blocks of the
    tiled kernel) -- SURVEY.md 8d item 3 -- plus census conservation."""
    S = 16384
    g = synthetic_grid(S, S, 3)
    sim = se.Simulation(default_rules, (S, S))
    sim.upload_cells(g)
    sim.params.frame = 1
    c0 = sim.census()
    sim.step(16)
    got = sim.download_cells()
    c1 = sim.census()
    sim.close()
    frame = oracle.run_blocks(g, 1, 16)          # in place on g
    assert frame == 17
    assert np.array_equal(got, g)
    for mat in (3, 4, 5, 6, 8, 10):
        assert c1[mat] == c0[mat]


def test_gpu_matches_committed_state_hashes(se):
    """The CUDA path against the committed fixtures tests/golden/oracle_state_sha256.json (no oracle run involved)."""
    import importlib.util
    import json
    from pathlib import Path
    here = Path(__file__).parent / "golden"
    spec = importlib.util.spec_from_file_location("make_oracle_state_hashes", here / "make_oracle_state_hashes.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = json.loads((here / "oracle_state_sha256.json").read_text())
    for name, text, g, steps, _ in mod.cases():
        rules = se.parse_string(text)
        got, _, _ = run_gpu(se, rules, g, steps)
        assert hashlib.sha256(got.astype(np.uint32).tobytes()).hexdigest() == want[name]["final"], name


def test_running_census(se, oracle, monkeypatch):
    """SE_FLAG_RUNNING_CENSUS: cells stay bit-exact and every census (sync and async, after K1c steps, after runs that
    fall back to a recount, with WALL / NULL / unknown ids, odd widths) equals a host recount."""
    import torch
    default_rules = se.parse_path(DEFAULT_YAML)
    for (w, h, seed) in [(516, 130, 41), (1024, 768, 42), (260, 258, 43)]:
        g = synthetic_grid(w, h, seed)
        rng = np.random.default_rng(seed)
        g[rng.integers(0, h, 100), rng.integers(0, w, 100)] = 2
        g[rng.integers(0, h, 30), rng.integers(0, w, 30)] = 1
        g[3, 3] = 77; g[5, 9] = 4000000000
        sim = se.Simulation(default_rules, (w, h), running_census=True)
        sim.upload_cells(g); sim.params.frame = 1
        ref = g.copy(); frame = 1
        ring = torch.zeros((4, 256), dtype=torch.int64).pin_memory()
        for k in range(40):
            n = 1 if k % 7 else 9                       # mostly the per-frame kernel, sometimes a tiled run (invalidates)
            sim.step(n)
            frame = oracle.run_blocks(ref, frame, n)
            want = np.bincount(np.minimum(ref, 255).ravel(), minlength=256)
            if k % 2:
                assert np.array_equal(sim.census(), want), (w, h, k)
            else:
                sim.census_async(ring[k % 4].data_ptr()); sim.census_wait()
                assert np.array_equal(ring[k % 4].numpy(), want), (w, h, k)
        assert np.array_equal(sim.download_cells(), ref)
        sim.close()


@pytest.mark.parametrize("lighting", [False, True])
def test_push_brush_matches_the_oracle_circle_stamp(se, default_rules, oracle, lighting):
    """`Simulation.push_brush()` (sandengine-core/src/lib.rs:59-67: one CIRCLE of brushSize / brushMaterial at mousePos * size) against the
    oracle's stamp (falling_sand.glsl:749-794, the f32 distance test): every brush size 0..24 at positions inside, on the edge of and
    outside the grid, one stamp per frame with the simulation running in between."""
    from sandengine_b200 import MOD_DTYPE
    w, h = 128, 96
    g = synthetic_grid(w, h, 61)
    L0 = np.zeros((h, w, 4), np.float32)
    sim = se.Simulation(default_rules, (w, h), lighting=lighting)
    sim.upload_cells(g); sim.params.frame = 1
    if lighting: sim.upload_light(L0)
    selectable = default_rules.selectable_materials
    ref, refL, frame = g.copy(), (L0.copy() if lighting else None), 1
    spots = [(0.5, 0.5), (0.0, 0.0), (0.999, 0.999), (0.25, 0.9), (1.2, 0.5), (-0.1, 0.3), (0.7, 0.01)]
    for k in range(50):
        size = k % 25
        mouse = spots[k % len(spots)]
        mat = selectable[k % len(selectable)]
        sim.params.mousePos = list(mouse); sim.params.brushSize = size; sim.params.brushMaterial = mat
        sim.push_brush()
        sim.run()
        m = np.zeros(1, MOD_DTYPE)
        m[0] = ((int(mouse[0] * w), int(mouse[1] * h)), 0, size, mat.id, (0, 0, 0))
        ref, refL, frame = oracle.run(ref, frame, 1, light=refL, mods_per_step=[m])
        assert np.array_equal(sim.download_cells(), ref), (k, size, mouse)
    if lighting:
        assert np.abs(sim.download_light() - refL).max() <= LIGHT_ATOL
    sim.close()


@pytest.mark.parametrize("running_census", [False, True])
def test_brush_held_down_stays_on_the_per_frame_table_kernel(se, default_rules, oracle, running_census):
    """A brush held down (modifications every frame, the reference's normal interactive use) runs on se_step_lut_global[_census]_mods:
    cells bit-exact, the running census equal to a recount every frame, also with unknown ids, WALL / NULL cells, stamps that hang over
    the grid's edge, more than 32 records, and frames with runs of steps behind the modified one."""
    from sandengine_b200 import MOD_DTYPE
    for (w, h, seed) in [(516, 130, 51), (1024, 512, 52)]:
        rng = np.random.default_rng(seed)
        g = synthetic_grid(w, h, seed)
        g[rng.integers(0, h, 60), rng.integers(0, w, 60)] = 2
        g[rng.integers(0, h, 20), rng.integers(0, w, 20)] = 1
        g[7, 11] = 99; g[40, 300] = 3000000000
        sim = se.Simulation(default_rules, (w, h), running_census=running_census)
        sim.upload_cells(g); sim.params.frame = 1
        frames = 36
        mods = []
        for k in range(frames):
            m = make_mods(se, k, 11, w, h, rng)
            if k == 5:                                   # 40 records: more than one warp-load of the list
                m = np.zeros(40, MOD_DTYPE)
                for i in range(40):
                    m[i] = ((int(rng.integers(0, w)), int(rng.integers(0, h))), i % 2, 1 + i % 9, int(rng.integers(0, 11)), (0, 0, 0))
            if k == 9:                                   # a stamp larger than the grid
                m = np.zeros(1, MOD_DTYPE); m[0] = ((w // 2, h // 2), 1, 5000, 4, (0, 0, 0))
            mods.append(m)
        steps_per_frame = [1 if k % 6 else 3 for k in range(frames)]     # sometimes a run of steps behind the modified one
        ref = g.copy(); frame = 1
        for k in range(frames):
            if len(mods[k]):
                sim.push_modifications(mods[k])
            sim.step(steps_per_frame[k])
            ref, _, frame = oracle.run(ref, frame, steps_per_frame[k], mods_per_step=[mods[k]] + [np.zeros(0, MOD_DTYPE)] * (steps_per_frame[k] - 1))
            want = np.bincount(np.minimum(ref, 255).ravel(), minlength=256)
            assert np.array_equal(sim.census(), want), (w, h, k)
        assert np.array_equal(sim.download_cells(), ref)
        sim.close()


@pytest.mark.parametrize("n_strips,halo,w,h", [(2, 4, 96, 64), (3, 6, 200, 150), (4, 2, 64, 64)])
def test_lit_strips(se, default_rules, oracle, n_strips, halo, w, h):
    """Strips with lighting in one process (se_sim_attach_local): ids bit-exact, light bit-exact against the oracle's
    full-grid run; ghost rows of ids AND light are exchanged every `halo` steps (StripPlan(lighting=True))."""
    from sandengine_b200.distributed import StripPlan
    rng = np.random.default_rng(23)
    g = synthetic_grid(w, h, 21)
    L0 = rng.random((h, w, 4), dtype=np.float32)
    L0[rng.random((h, w)) < 0.2, 3] = 0.0
    steps = 3 * halo + 1
    ref, refL, _ = oracle.run(g, 1, steps, light=L0)
    plan = StripPlan(w, h, n_strips, halo, lighting=True)
    sims = []
    for r in range(n_strips):
        b, e = plan.rows(r)
        s = se.Simulation(default_rules, (w, h), lighting=True, lit_strip=True, row_begin=b, row_end=e, halo_rows=halo, device_share=n_strips)
        s.upload_cells(g[b:e]); s.upload_light(np.ascontiguousarray(L0[b:e])); s.params.frame = 1
        sims.append(s)
    for r, s in enumerate(sims):
        if r > 0: s.attach_local(0, sims[r - 1])
        if r < n_strips - 1: s.attach_local(1, sims[r + 1])

    def exchange():
        for s in sims: s.synchronize()
        for s in sims: s.halo_push()
        for s in sims: s.synchronize()
    exchange()
    for k in plan.chunks(steps):
        for s in sims: s.step(k)
        exchange()
    got = np.concatenate([s.download_cells() for s in sims], axis=0)
    gotL = np.concatenate([s.download_light() for s in sims], axis=0)
    for s in sims: s.close()
    assert np.array_equal(got, ref)
    assert np.abs(gotL - refL).max() <= LIGHT_ATOL


def test_fused_light(se, oracle, monkeypatch):
    """se_light_fused: ids bit-exact and light within tolerance against the oracle, with modifications, frame 1 included."""
    default_rules = se.parse_path(DEFAULT_YAML)
    rng = np.random.default_rng(29)
    for (w, h, steps, frame0) in [(200, 150, 24, 1), (80, 64, 30, 1), (33, 17, 10, 0), (516, 130, 12, 1)]:
        g = synthetic_grid(w, h, 31)
        g[rng.integers(0, h, 20), rng.integers(0, w, 20)] = 2
        g[rng.integers(0, h, 10), rng.integers(0, w, 10)] = 1
        L0 = rng.random((h, w, 4), dtype=np.float32)
        L0[rng.random((h, w)) < 0.2, 3] = 0.0
        mods = [make_mods(se, s, 11, w, h, rng) if s % 2 else np.zeros(0, se.MOD_DTYPE) for s in range(steps)]
        ref, refL, _ = oracle.run(g, frame0, steps, light=L0, mods_per_step=mods)
        sim = se.Simulation(default_rules, (w, h), lighting=True, fused_light=True)
        sim.upload_cells(g); sim.upload_light(L0); sim.params.frame = frame0
        for s in range(steps):
            if len(mods[s]): sim.push_modifications(mods[s])
            sim.run()
        got, gotL = sim.download_cells(), sim.download_light()
        sim.close()
        assert np.array_equal(got, ref), (w, h)
        assert np.abs(gotL - refL).max() <= LIGHT_ATOL, (w, h)


@pytest.mark.parametrize("seed,n_mat,n_rules", [(201, 7, 23), (216, 12, 22)])
def test_table_kernels_on_random_eligible_rule_sets(se, seed, n_mat, n_rules):
    """K1b / K1c with table-eligible rule sets OTHER than the default one (different material count, table size,
    rand.y thresholds): random mirrored-only sets from synth_rules against the oracle."""
    from oracle.build_oracle import load_oracle
    from sandengine_b200.synth_rules import synthetic_rule_set
    text, ids, mix = synthetic_rule_set(n_mat, n_rules, seed=seed, kinds=("mirrored",))
    rules = se.parse_string(text)
    assert "#define SE_LUT_ELIGIBLE 1" in rules.cuda_header
    orc = load_oracle(text)
    for (w, h, steps) in [(512, 300, 67), (260, 130, 33)]:
        g = synthetic_grid(w, h, seed, mix=mix, ids=ids)
        ref, _, _ = orc.run(g, 1, steps, blocks=True)
        got, _, _ = run_gpu(se, rules, g, steps)                      # K1b
        assert np.array_equal(got, ref), ("K1b", w, h)
        sim = se.Simulation(rules, (w, h))
        sim.upload_cells(g); sim.params.frame = 1
        for _ in range(steps): sim.step(1)                            # K1c
        assert np.array_equal(sim.download_cells(), ref), ("K1c", w, h)
        sim.close()


def test_two_table_lut(se, monkeypatch):
    """K1b / K1c for rule sets with Left/Right rules through one table per view (SE_LUT_LR=1 at rule-compile time)."""
    from oracle.build_oracle import load_oracle
    from sandengine_b200.synth_rules import synthetic_rule_set
    for seed, n_mat, n_rules in [(401, 9, 20), (505, 11, 21)]:
        text, ids, mix = synthetic_rule_set(n_mat, n_rules, seed=seed)
        rules = se.parse_string(text)
        assert "#define SE_LUT_TWO_TABLES 1" in rules.cuda_header
        orc = load_oracle(text)
        for (w, h, steps) in [(512, 300, 67), (260, 130, 33)]:
            g = synthetic_grid(w, h, seed, mix=mix, ids=ids)
            ref, _, _ = orc.run(g, 1, steps, blocks=True)
            got, _, _ = run_gpu(se, rules, g, steps)
            assert np.array_equal(got, ref), ("K1b", seed, w, h)
            sim = se.Simulation(rules, (w, h))
            sim.upload_cells(g); sim.params.frame = 1
            for _ in range(steps): sim.step(1)
            assert np.array_equal(sim.download_cells(), ref), ("K1c", seed, w, h)
            sim.close()


@pytest.mark.parametrize("n_strips,halo", [(2, 4), (3, 8)])
def test_strips_with_modifications(se, default_rules, oracle, n_strips, halo):
    """modifications are broadcast to every strip (global coordinates, clipped by the kernels); the
    sharded result stays bit-identical to the oracle's full-grid run."""
    from sandengine_b200.distributed import StripPlan
    rng = np.random.default_rng(77)
    w, h, steps = 96, 100, 40
    g = synthetic_grid(w, h, 17)
    mods = [make_mods(se, s, 11, w, h, rng) if s % 3 == 0 else np.zeros(0, se.MOD_DTYPE) for s in range(steps)]
    ref, _, _ = oracle.run(g, 1, steps, mods_per_step=mods)
    plan = StripPlan(w, h, n_strips, halo)
    sims = []
    for r in range(n_strips):
        b, e = plan.rows(r)
        s = se.Simulation(default_rules, (w, h), row_begin=b, row_end=e, halo_rows=halo, device_share=n_strips)
        s.upload_cells(g[b:e]); s.params.frame = 1
        sims.append(s)
    for r, s in enumerate(sims):
        if r > 0: s.attach_local(0, sims[r - 1])
        if r < n_strips - 1: s.attach_local(1, sims[r + 1])

    def exchange():
        for s in sims: s.synchronize()
        for s in sims: s.halo_push()
        for s in sims: s.synchronize()
    exchange()
    for k in range(steps):                       # one frame at a time: exchange after every step keeps it simple
        for s in sims:
            if len(mods[k]): s.push_modifications(mods[k])
            s.step(1)
        exchange()
    got = np.concatenate([s.download_cells() for s in sims], axis=0)
    for s in sims: s.close()
    assert np.array_equal(got, ref)
