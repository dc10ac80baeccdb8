"""CPU: the oracle's portable sin/cos/atan2 (oracle/portable_sincos.hpp — the algorithms both the oracle and the CUDA
kernels use instead of the platform libm) are accurate, and swapping them for glibc's, or compiling the oracle with
FMA contraction, changes the reference algorithm's outputs only at the level of its own ill-conditioning."""
import numpy as np


def _ulps(a, exact_ld, ref):
    return (np.abs(a.astype(np.longdouble) - exact_ld) / np.spacing(np.abs(ref))).max()


def test_portable_sincos_and_atan2_accuracy(oracle_mod):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-10, 10, 400_000), rng.uniform(-1e5, 1e5, 100_000), rng.uniform(-0.8, 0.8, 100_000),
                        [0.0, -0.0, np.pi / 2, np.pi, 1e-300, 0.785398163397448, 0.7853981633974484]])
    s, c = oracle_mod.sincos(x)
    sg, cg = oracle_mod.sincos(x, "glibc")
    xl = x.astype(np.longdouble)
    assert _ulps(s, np.sin(xl), sg) < 2.0 and _ulps(c, np.cos(xl), cg) < 2.0
    assert np.abs(s * s + c * c - 1.0).max() < 1e-15
    y = rng.normal(size=300_000) * rng.choice([1e-3, 1.0, 1e3], 300_000)
    xx = rng.normal(size=300_000)
    a = oracle_mod.atan2(y, xx)
    assert _ulps(a, np.arctan2(y.astype(np.longdouble), xx.astype(np.longdouble)), oracle_mod.atan2(y, xx, "glibc")) < 2.0
    sp = oracle_mod.atan2([0.0, 0.0, 1.0, -1.0, -0.0, 2.0], [1.0, -1.0, 0.0, 0.0, 3.0, 1.0])
    assert np.array_equal(sp[:4], [0.0, np.pi, np.pi / 2, -np.pi / 2]) and np.signbit(sp[4]) and abs(sp[5] - np.arctan(2.0)) < 1e-15


def test_reference_noise_floor_between_oracle_builds(oracle_mod, scene2k):
    """Same source, three builds (portable math / glibc math / glibc + FMA contraction): cost and per-point SDF agree
    to rounding; the gradient agrees to ~1e-5 because of a few points whose minimiser sits where the robot is at rest."""
    sc = scene2k
    co = sc.coeffs_colmajor()
    res = {}
    for v in ("default", "glibc", "fma"):
        o = oracle_mod.Oracle("star", threads=min(8, oracle_mod.num_procs()), variant=v)
        o.set_points(sc.points)
        res[v] = o.cost_grad(sc.T, co, per_point=True)
    c0, gT0, gC0, pp0, _ = res["glibc"]
    for v in ("default", "fma"):
        c, gT, gC, pp, _ = res[v]
        assert abs(c - c0) <= 1e-12 * abs(c0)
        assert np.abs(pp[:, 0] - pp0[:, 0]).max() <= 1e-9  # per-point sdf
        assert np.linalg.norm(gC - gC0) / np.linalg.norm(gC0) < 1e-4
        assert np.median(np.abs(pp[:, 1] - pp0[:, 1])) < 1e-7  # typical t* agreement
        worst = np.argsort(-np.abs(pp[:, 1] - pp0[:, 1]))[:3]
        D = sc.T.sum()
        assert np.all((pp0[worst, 1] > D - 0.01) | (pp0[worst, 1] < 0.01) | (np.abs(pp[worst, 1] - pp0[worst, 1]) < 1e-6))
