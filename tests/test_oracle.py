"""Pins the oracle: (1) every exact known answer the reference's tests hold for the path
(tests/golden/known_answers.json, transcribed from /root/reference/test), (2) the reference's
finite-difference gradchecks at its own tolerance (src/gradcheck.jl:9, rtol=atol=1e-5,
Float64), (3) the C restatement of the scatter fold against the NumPy one."""
import json
import os

import numpy as np
import pytest

import cases
from oracle import c_oracle
from oracle import yota_oracle as O
from yota_b200 import jl

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "known_answers.json")))
G = {c["name"]: c for c in GOLD}


def _idx(spec):
    return tuple(O.colon if i == ":" else (np.array(i, dtype=np.int64) if isinstance(i, list) else i) for i in spec)


def test_ungetindex_known_answers():
    c = G["ungetindex_scalar"]
    assert np.array_equal(O.jl_ungetindex(tuple(c["x_dims"]), np.float64, c["dy"], _idx(c["index"])), c["expect"])
    c = G["ungetindex_dups"]
    assert np.array_equal(O.jl_ungetindex(tuple(c["x_dims"]), np.float64, np.array(c["dy"]), _idx(c["index"])), c["expect"])
    c = G["ungetindex_cols"]
    got = O.jl_ungetindex(tuple(c["x_dims"]), np.float64, np.ones(c["dy_ones"]), _idx(c["index"]))
    assert np.array_equal(got, np.array(c["expect"]))
    # the plain-C fold gives the same bits
    assert np.array_equal(c_oracle.scatter_add_cols(np.ones((4, 3), np.float32), [1, 3, 1], 5), np.array(c["expect"], np.float32))


def test_getindex_pullbacks_exact():
    x = cases.rng(0).random((3, 4, 5))
    for name in ("getindex_linear", "getindex_cartesian"):
        c = G[name]
        _, g = O.grad(lambda x: x[tuple(c["index"])] if len(c["index"]) > 1 else x[c["index"][0]], x, dtype=np.float64)
        e = np.zeros_like(x)
        e[tuple(c["onehot"])] = 1
        assert np.array_equal(g[1], e), name
    _, g = O.grad(lambda x: jl.sum(x[jl.colon, 1, jl.colon]), x, dtype=np.float64)
    e = np.zeros_like(x)
    e[:, 0, :] = 1
    assert np.array_equal(g[1], e)


def test_seed_and_trivial_cases():
    c = G["seed_vjp"]
    for seed in (np.ones(3), "auto"):
        val, g = O.grad(lambda x: 2 * x, np.array(c["x"]), seed=seed, dtype=np.float64)
        assert np.array_equal(val, c["val"]) and g[0] is O.ZeroTangent and np.array_equal(g[1], c["grad"])
    _, g = O.grad(lambda x, y: x, 42.0, 54.0)
    assert g == (O.ZeroTangent, 1, O.ZeroTangent)
    _, g = O.grad(lambda x, y: y, 42.0, 54.0)
    assert g == (O.ZeroTangent, O.ZeroTangent, 1)
    _, g = O.grad(lambda x: 1, 42.0)
    assert g == (O.ZeroTangent, O.ZeroTangent)
    with pytest.raises(ValueError):
        O.grad(lambda x: 2 * x, np.ones(3))  # vector-valued without a seed (src/grad.jl:204-205)


def test_transpose_ones():
    for d in G["transpose_ones"]["dims"]:
        a = cases.rng(1).random(tuple(d))
        for f in (jl.transpose, jl.adjoint):
            _, g = O.grad(lambda x: jl.sum(f(x)), a, dtype=np.float64)
            assert np.array_equal(g[1], np.ones_like(a))


def test_update_known_answer():
    c = G["update_array"]
    x = np.array(c["x"])
    assert np.array_equal(O.update(x, np.array(c["gx"])), c["expect"])


@pytest.mark.parametrize("name,f,args", cases.gradcheck_cases(), ids=lambda v: v if isinstance(v, str) else "")
def test_gradcheck(name, f, args):
    assert O.gradcheck(f, *args)


def test_bcast_grad_sizes():  # test/test_grad.jl:141-146,163-167: size(g) == size(arg)
    for f in (cases.sum_bcast, cases.sum_prod_bcast):
        for sa, sb in cases.BCAST_SHAPES:
            r = cases.rng(3)
            a, b = r.random(sa), r.random(sb)
            _, g = O.grad(f, a, b, dtype=np.float64)
            assert g[1].shape == a.shape and g[2].shape == b.shape


def test_cat_gradcheck():  # test/test_grad.jl:98-104
    assert O.gradcheck(lambda a, b: jl.sum(jl.vcat(a, b)), np.ones((2, 3)), np.ones((3, 3)))
    assert O.gradcheck(lambda a, b: jl.sum(jl.hcat(a, b)), np.ones((2, 3)), np.ones((2, 4)))
    arrs = [i * np.ones((3, 2, i, 2)) for i in range(1, 4)]
    assert O.gradcheck(lambda a, b, c: jl.sum(jl.cat(a, b, c, dims=3)), *arrs)


def test_configs_gradcheck_small():
    """The benchmark tapes at toy sizes, FD-checked in Float64."""
    a = [x.astype(np.float64) if x.dtype.kind == "f" else x for x in cases.c1_args(6, 5, 3, 4)]
    assert O.gradcheck(cases.c1_mlp, *a)
    a = [x.astype(np.float64) for x in cases.c2_args(4, 5)]
    assert O.gradcheck(cases.c2_chain, *a)
    a = [x.astype(np.float64) for x in cases.c3_args(3, 5, 4)]
    assert O.gradcheck(cases.make_c3(3, 4), *a)
    a = [x.astype(np.float64) for x in cases.c4_args(4, 3, 3)]
    assert O.gradcheck(cases.make_c4(3), *a)
    E, idx, V = cases.c5_args(3, 5, 8)
    assert O.gradcheck(cases.c5_embed, E.astype(np.float64), idx, V.astype(np.float64))


def test_accumulation_order_is_reverse_tape_order():
    """src/grad.jl:83-95 + 217-219: a variable used at tape positions p1<p2<p3 receives
    acc = d(p3); acc = d(p2) + acc; acc = d(p1) + acc."""
    x = np.array([1e8, 1.0, -1e8], dtype=np.float32)
    f = lambda a: jl.sum(a * np.float32(1e8) + a * np.float32(1.0) + a * np.float32(-1e8))  # noqa: E731
    _, g = O.grad(f, x)
    expect = np.float32(1e8) + (np.float32(1.0) + np.float32(-1e8))
    assert g[1][0] == expect


def test_c_scatter_matches_numpy_fold():
    r = cases.rng(7)
    dy = r.standard_normal((5, 200)).astype(np.float32)
    idx = r.integers(1, 12, size=200)
    a = c_oracle.scatter_add_cols(dy, idx, 11)
    b = O.jl_ungetindex((5, 11), np.float32, dy, (O.colon, idx))
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # -0.0 contributions: 0.0f + (-0.0f) = +0.0f
    z = c_oracle.scatter_add_cols(np.full((1, 1), -0.0, np.float32), [1], 1)
    assert not np.signbit(z[0, 0])


def test_examples_converge():
    """test/test_examples.jl:12-36 (loss drops >= 10x in 100 gradient steps) and :46-70 (W, b
    recovered to atol 0.1 with the update! rule x - 0.1 gx), on the oracle."""
    Yv, X, theta = cases.linreg_data()
    first, _ = O.grad(cases.linreg_obj, Yv, X, theta)
    last = first
    for _ in range(100):
        last, g = O.grad(cases.linreg_obj, Yv, X, theta)
        theta = theta - 0.01 * g[3]
    assert last < first / 10
    X, Yv, W_true, b_true, W, b = cases.linreg2_data()
    for _ in range(100):
        _, g = O.grad(cases.linear2_loss, W, b, X, Yv)
        W, b = W - 0.1 * g[1], b - 0.1 * g[2]
    assert np.allclose(W, W_true, atol=0.1) and np.allclose(b, b_true, atol=0.1)


def test_broadcast_pullback_equals_elementwise_pullbacks():
    """test/test_rulesets.jl:21-36: the pullback of the broadcast rule applied to ones equals the
    per-element scalar pullbacks applied to 1.0 (sin: cos(x_i)), exactly."""
    xs = cases.rng(3).random(2)
    _, g = O.grad(lambda x: jl.sum(jl.sin(x)), xs, dtype=np.float64)
    assert np.array_equal(g[1], np.cos(xs))
    inc = lambda x: jl.sum(jl.sin(x) + 1.0)  # noqa: E731  (sin_inc, :13)
    _, g = O.grad(inc, xs, dtype=np.float64)
    assert np.array_equal(g[1], np.cos(xs))
