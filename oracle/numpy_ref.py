"""TEST INFRASTRUCTURE - second, independent restatement in plain NumPy loops
(small cases only) used to cross-check oracle/tf_ops.py.  Same citations:
TF-1.3 semantics per SURVEY Appendix A at the reference call sites named in
oracle/tf_ops.py.  PARITY UNPINNED by the reference (it has no tests)."""
import numpy as np


def same_pad(n, k=3, s=2):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2, total - total // 2


def conv2d_3x3_s2_same(x, w, b):
    """x [N,H,W,Cin], w [3,3,Cin,Cout] -> [N,OH,OW,Cout]; explicit loops."""
    N, H, W, Cin = x.shape
    Cout = w.shape[3]
    OH, pt, _ = same_pad(H)
    OW, pl, _ = same_pad(W)
    y = np.zeros((N, OH, OW, Cout))
    for n in range(N):
        for oy in range(OH):
            for ox in range(OW):
                acc = b.astype(np.float64).copy()
                for ky in range(3):
                    iy = 2 * oy + ky - pt
                    if iy < 0 or iy >= H:
                        continue
                    for kx in range(3):
                        ix = 2 * ox + kx - pl
                        if ix < 0 or ix >= W:
                            continue
                        acc += x[n, iy, ix] @ w[ky, kx]
                y[n, oy, ox] = acc
    return y


def lrelu(x):
    return np.where(x > 0, x, 0.2 * x)   # == 0.6x + 0.4|x|


def batch_norm_train(x, gamma, beta, eps=1e-3):
    flat = x.reshape(-1, x.shape[-1])
    mean = flat.sum(0) / flat.shape[0]
    var = ((flat - mean) ** 2).sum(0) / flat.shape[0]
    return (x - mean) / np.sqrt(var + eps) * gamma + beta, mean, var


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_cell(x, c, h, kernel, bias, forget_bias=1.0):
    z = np.concatenate([x, h], 1) @ kernel + bias
    H = h.shape[1]
    i, j, f, o = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
    c2 = c * sigmoid(f + forget_bias) + sigmoid(i) * np.tanh(j)
    return c2, np.tanh(c2) * sigmoid(o)


def dynamic_rnn(x, lens, kernel, bias, c0, h0):
    R, T, _ = x.shape
    c, h = c0.copy(), h0.copy()
    out = np.zeros((R, T, h.shape[1]))
    for t in range(T):
        c2, h2 = lstm_cell(x[:, t], c, h, kernel, bias)
        for r in range(R):
            if t < lens[r]:
                c[r], h[r] = c2[r], h2[r]
                out[r, t] = h2[r]
    return out, h, c


def softmax_ce_loss(logits, onehot, lens):
    R, L, V = logits.shape
    tot, cnt = 0.0, 0.0
    for r in range(R):
        for t in range(L):
            if t < lens[r]:
                x = logits[r, t] - logits[r, t].max()
                logp = x - np.log(np.exp(x).sum())
                tot += -(onehot[r, t] * logp).sum()
                cnt += 1
    return tot / cnt


def sigmoid_ce_loss(logits, labels, lens):
    R, L, P = logits.shape
    tot, cnt = 0.0, 0.0
    for r in range(R):
        for t in range(L):
            if t < lens[r]:
                x, z = logits[r, t], labels[r, t]
                tot += (np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))).mean()
                cnt += 1
    return tot / cnt


def adam_clip_step(p, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, clip=20.0):
    norm = np.sqrt((g.astype(np.float64) ** 2).sum())
    g = g * (clip / max(norm, clip))
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return p - lr_t * m / (np.sqrt(v) + eps), m, v, norm


# Karel DSL vocabulary, derived from the token order of reference
# karel_env/dsl/dsl_prob.py:13-28 with INT expanded to R=0..R=19
# (karel_env/dsl/dsl_base.py:80-91).
KAREL_VOCAB = (['DEF', 'run', 'm(', 'm)', 'move', 'turnRight', 'turnLeft', 'pickMarker',
                'putMarker', 'r(', 'r)'] + ['R=%d' % i for i in range(20)] +
               ['REPEAT', 'c(', 'c)', 'i(', 'i)', 'e(', 'e)', 'IF', 'IFELSE', 'ELSE',
                'frontIsClear', 'leftIsClear', 'rightIsClear', 'markersPresent',
                'noMarkersPresent', 'not', 'w(', 'w)', 'WHILE'])
