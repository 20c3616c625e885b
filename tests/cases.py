"""Loss functions and inputs shared by the CPU and GPU tests.

Each case mirrors a reference test (file:line cited) or a BASELINE.json config scaled down
so the oracle finishes in seconds.  Functions are written once against `yota_b200.jl` and run
unchanged on the oracle's arrays and on traced device arrays (dispatch on the array type)."""
import numpy as np

from yota_b200 import jl

colon = jl.colon


def rng(seed):
    return np.random.default_rng(seed)


# ---- reference test/test_grad.jl ------------------------------------------------------
def loss_simple(W, b, x):  # :12, :29, :59
    return jl.sum(W @ x + b)


def loss_tanh(W, b, x):  # :30
    return jl.sum(jl.tanh(W @ x + b))


def loss_abs2(W, b, x, y):  # :31
    return jl.sum(jl.abs2(jl.tanh(W @ x + b) - y))


def loss_double_broadcast(W, b, x):  # :13
    return jl.sum(jl.sin(W @ x) + b)


def loss_double_broadcast2(b, x):  # :14
    return jl.sum(x * x + b)


def loss_kw_mean(W, b, x):  # :15
    return jl.mean(W @ x + b, dims=1)[1]


def loss_sumsum(x):  # :67
    return jl.sum(jl.sum(x, dims=1))


def multipath_loss(y, c):  # :37-54
    f = jl.tanh(y[jl.range_(1, 5), colon])
    c_ = f * c
    h_ = jl.tanh(c_)
    return jl.sum(h_)


def sum_bcast(x, y):  # :127
    return jl.sum(x + y)


def sum_prod_bcast(x, y):  # :149
    return jl.sum(x * y)


def literal_pow(x):  # :172
    return jl.sum(x ** 2)


def linreg_obj(Y, X, b):  # test/test_examples.jl:9
    return jl.mean((Y - X @ b) ** 2.0)


BCAST_SHAPES = [((3, 4), (3,)), ((3, 4), (3, 1)), ((3, 4), (1, 4)), ((3,), (3, 4)), ((3, 1), (3, 4)), ((1, 4), (3, 4))]


def gradcheck_cases():
    """(name, f, args) -- the reference's FD-checked cases (Float64 there)."""
    r = rng(11)
    a3 = (r.random((4, 3)), r.random(4), r.random(3))
    a4 = a3 + (r.random(4),)
    k = (r.random((3, 4)), r.random(3), r.random(4))
    out = [
        ("simple", loss_simple, a3), ("tanh", loss_tanh, a3), ("abs2", loss_abs2, a4),
        ("multipath", multipath_loss, (r.random((20, 4)), r.random((5, 4)))),
        ("kw_simple", loss_simple, k), ("double_bcast", loss_double_broadcast, k),
        ("double_bcast2", loss_double_broadcast2, (r.random(3), r.random(3))),
        ("kw_mean", loss_kw_mean, k), ("sumsum", loss_sumsum, (r.random((2, 3)),)),
        ("literal_pow", literal_pow, (r.random((3, 4)),)),
        ("linreg", linreg_obj, (r.random(10), r.random((10, 3)), r.random(3))),
    ]
    for i, (sa, sb) in enumerate(BCAST_SHAPES):
        out.append((f"bcast_add{i}", sum_bcast, (r.random(sa), r.random(sb))))
        out.append((f"bcast_mul{i}", sum_prod_bcast, (r.random(sa), r.random(sb))))
    return out


# ---- BASELINE.json configs, parametrised by size -------------------------
def c1_mlp(W1, b1, W2, b2, x, y):
    return jl.sum((W2 @ jl.tanh(W1 @ x + b1) + b2 - y) ** 2)


def c1_args(d_in=784, d_h=128, d_out=10, B=64, seed=1):
    r = rng(seed)
    f = np.float32
    return ((r.standard_normal((d_h, d_in)) / np.sqrt(d_in)).astype(f), (0.01 * r.standard_normal(d_h)).astype(f),
            (r.standard_normal((d_out, d_h)) / np.sqrt(d_h)).astype(f), (0.01 * r.standard_normal(d_out)).astype(f),
            r.standard_normal((d_in, B)).astype(f), r.random((d_out, B)).astype(f))


def c2_chain(x, W, b):
    return jl.sum(jl.tanh(x * W + b) ** 2)


def c2_args(n=8192, m=None, seed=2):
    r = rng(seed)
    m = m or n
    f = np.float32
    return ((0.5 * r.standard_normal((n, m), dtype=f)), (0.5 * r.standard_normal((n, m), dtype=f)),
            (0.1 * r.standard_normal(n, dtype=f)))


def make_c3(n_layers, batch):
    def c3_mlp(*a):
        Ws, bs, x, Y = a[:n_layers], a[n_layers:2 * n_layers], a[-2], a[-1]
        h = x
        for l in range(n_layers - 1):
            h = jl.tanh(Ws[l] @ h + bs[l])
        z = Ws[-1] @ h + bs[-1]
        return -jl.sum(Y * jl.logsoftmax(z)) / batch

    return c3_mlp


def c3_args(n_layers=8, width=4096, batch=8192, seed=3):
    r = rng(seed)
    f = np.float32
    Ws = [(r.standard_normal((width, width), dtype=f) / np.sqrt(width)).astype(f) for _ in range(n_layers)]
    bs = [(0.01 * r.standard_normal(width, dtype=f)).astype(f) for _ in range(n_layers)]
    x = r.standard_normal((width, batch), dtype=f)
    labels = r.integers(0, width, size=batch)
    Y = np.zeros((width, batch), dtype=f, order="F")
    Y[labels, np.arange(batch)] = 1
    return tuple(Ws) + tuple(bs) + (x, Y)


def make_c4(T):
    def c4_rnn(Wx, Wh, b, x):
        h = None
        for t in range(1, T + 1):
            xt = x[colon, colon, t]
            z = Wx @ xt + b if h is None else Wx @ xt + Wh @ h + b
            h = jl.tanh(z)
        return jl.sum(h ** 2)

    return c4_rnn


def c4_args(H=1024, B=128, T=128, seed=4):
    r = rng(seed)
    f = np.float32
    return ((0.9 * r.standard_normal((H, H), dtype=f) / np.sqrt(H)).astype(f),
            (0.9 * r.standard_normal((H, H), dtype=f) / np.sqrt(H)).astype(f),
            (0.01 * r.standard_normal(H, dtype=f)).astype(f), r.standard_normal((H, B, T), dtype=f))


def c5_embed(E, idx, V):
    return jl.sum(E[colon, idx] * V)


def c5_args(rows=256, n_table=1 << 20, n_idx=1 << 22, seed=5, zipf=False):
    r = rng(seed)
    f = np.float32
    E = r.standard_normal((rows, n_table), dtype=f)
    if zipf:
        idx = np.minimum(r.zipf(1.1, size=n_idx), n_table).astype(np.int64)
    else:
        idx = r.integers(1, n_table + 1, size=n_idx).astype(np.int64)
    V = r.standard_normal((rows, n_idx), dtype=f)
    return (np.asfortranarray(E), idx, np.asfortranarray(V))


# ---- the reference's end-to-end fixtures (test/test_examples.jl) ------------------------
def linear2_loss(W, b, X, y):  # test/test_examples.jl:44, the Linear2 struct flattened to (W, b)
    return jl.mean((W @ X + b - y) ** 2.0)


def linreg_data(n=1000, d=10, seed=1):  # test/test_examples.jl:14-24
    r = rng(seed)
    X = r.standard_normal((n, d))
    b = r.standard_normal(d)
    Yv = X @ b + 0.1 * r.standard_normal(n)
    return Yv, X, r.standard_normal(d)


def linreg2_data(n_vars=10, n_out=2, n_obs=1000, seed=2):  # test/test_examples.jl:48-56
    r = rng(seed)
    X = r.standard_normal((n_vars, n_obs))
    W_true, b_true = r.random((n_out, n_vars)), r.random(n_out)
    Yv = W_true @ X + b_true[:, None] + 0.1 * r.standard_normal((n_out, n_obs))
    return X, Yv, W_true, b_true, r.random((n_out, n_vars)), r.random(n_out)


# ---- the reference's examples: convnet (examples/flux/convnet.jl:53-66) and the MLP loss
# (examples/flux/mlp.jl:54,81) over the composite layers of yota_b200.jl ---------------------
def convnet_loss(w1, b1, Wd, bd, x, y):
    """Conv((3,3), C=>8, relu) -> MaxPool((2,2)) -> flatten -> Dense -> logitcrossentropy."""
    h = jl.relu(jl.conv(x, w1) + jl.reshape(b1, 1, 1, b1.shape[0], 1))
    h = jl.maxpool2(h)
    W, H, C, N = h.shape
    return jl.logitcrossentropy(Wd @ jl.reshape(h, W * H * C, N) + bd, y)


def convnet_args(W=10, H=10, C=3, N=4, CO=8, classes=5, seed=7):
    r = rng(seed)
    OW, OH = (W - 2) // 2, (H - 2) // 2
    lab = r.integers(0, classes, size=N)
    y = np.zeros((classes, N), order="F")
    y[lab, np.arange(N)] = 1
    return (r.standard_normal((3, 3, C, CO)) / 5, 0.1 * r.standard_normal(CO), r.standard_normal((classes, OW * OH * CO)) / 8,
            0.1 * r.standard_normal(classes), r.standard_normal((W, H, C, N)), y)


def mlp_relu_loss(W1, b1, W2, b2, x, y):  # examples/flux/mlp.jl:52-56, 81: Dense(relu) -> Dense, logitcrossentropy
    return jl.logitcrossentropy(W2 @ jl.relu(W1 @ x + b1) + b2, y)


def sum_of_f(x):  # sum(f, xs) with a non-primitive f: the nested tape of f (src/chainrules.jl:101-124), inlined
    return jl.sum(lambda t: jl.sin(t) * t + 1.0, x)
