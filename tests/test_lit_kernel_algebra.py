"""The two algebraic shortcuts of se_step_lit (csrc/kernels/sand_kernels.cuh, K3f) against the shader's own formulation, in IEEE
f32 on the CPU (numpy adds and multiplies are correctly rounded and never fused):

  phase C  math.glsl:168-176 accumulates, neighbour by neighbour,   falloff = (w == 0) ? max_falloff : w;  avg.w += falloff;
           max_falloff = max(falloff, max_falloff).   The kernel adds w with the packed add that also adds z, then adds max_falloff
           under a predicate when w == 0, and keeps the running maximum as max(max_falloff, w).
  phase A  falling_sand.glsl:498: term.rgb = (light.rgb * keep) * a with keep in {0, 1}.  The kernel multiplies by (keep ? a : 0).

Both must give the same f32 values for every input the light field can hold (alpha >= 0, finite), including zeros of either
sign and denormals; the GPU parity tests then only have to show that the kernel evaluates what is modelled here."""
import numpy as np

F = np.float32


def specials(rng, shape):
    """f32 values that stress the rules: exact zeros of both signs, denormals, tiny, ordinary and large magnitudes."""
    pool = np.array([0.0, -0.0, 1e-45, 3e-39, 1.1754944e-38, 1e-30, 1e-7, 0.25, 0.5, 0.999999, 1.0, 3.0, 1e30], dtype=F)
    pick = pool[rng.integers(0, len(pool), shape)]
    rnd = rng.random(shape, dtype=F)
    return np.where(rng.random(shape) < 0.6, pick, rnd).astype(F)


def test_zero_alpha_rule_as_a_predicated_add():
    rng = np.random.default_rng(2024)
    n = 400_000
    w = specials(rng, (n, 8))
    w = np.abs(w) * np.where(rng.random((n, 8)) < 0.02, F(-1), F(1))        # a few negative alphas: not produced by the shader, still equal
    z = specials(rng, (n, 8))
    # the shader
    avg_w = np.zeros(n, F); mf = np.zeros(n, F)
    for k in range(8):
        falloff = np.where(w[:, k] == 0, mf, w[:, k]).astype(F)
        avg_w = (avg_w + falloff).astype(F)
        mf = np.maximum(falloff, mf).astype(F)
    # the kernel (SE_LF_ACC): packed add of (z, w), predicated add of the running maximum, maximum with the raw alpha
    s_w = np.zeros(n, F); s_z = np.zeros(n, F); m2 = np.zeros(n, F)
    for k in range(8):
        s_z = (s_z + z[:, k]).astype(F)
        s_w = (s_w + w[:, k]).astype(F)
        s_w = np.where(w[:, k] == 0, (s_w + m2).astype(F), s_w).astype(F)
        m2 = np.maximum(m2, w[:, k]).astype(F)
    assert np.array_equal(avg_w, s_w)                                      # == treats +0 and -0 alike ...
    nz = avg_w != 0
    assert np.array_equal(avg_w[nz].view(np.uint32), s_w[nz].view(np.uint32))   # ... everything else is bit-equal
    assert np.array_equal(mf, m2)


def test_keep_times_alpha_folded_into_one_factor():
    rng = np.random.default_rng(7)
    n = 1_000_000
    x = specials(rng, n); a = specials(rng, n)                             # light.rgb >= 0, alpha >= 0
    for keep in (F(0.0), F(1.0)):
        with np.errstate(over="ignore"):                                   # 1e30 * 1e30 overflows alike on both sides
            shader = ((x * keep).astype(F) * a).astype(F)
            kernel = (x * (a if keep == 1 else np.zeros(n, F))).astype(F)
        assert np.array_equal(shader, kernel)                              # == treats +0 and -0 alike; everything else is bit-equal
        nz = shader != 0
        assert np.array_equal(shader[nz].view(np.uint32), kernel[nz].view(np.uint32))
