"""numpy restatement of the reference's optimizer arithmetic (TEST INFRASTRUCTURE).

  * ``tf.keras.optimizers.Adam(lr, epsilon=1e-8)``   reference code/train.py:158-160,195-197
    -> TF's ApplyAdam dense functor (third-party, un-vendored, unpinned; published algorithm):
         alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
         m += (g - m) * (1 - beta1);  v += (g^2 - v) * (1 - beta2);  var -= (m * alpha) / (sqrt(v) + eps)
       all in float32 (the variables' dtype), t = iterations + 1.
  * ``tf.keras.experimental.CosineDecay(lr0, epochs)(epoch)``   code/train.py:92-100, evaluated once per epoch by a
    LearningRateScheduler: lr0 * 0.5 * (1 + cos(pi * min(epoch, epochs) / epochs)) in float32 (alpha = 0).
Parity is unpinned by the reference (no TF here); the cross-check is torch.optim.Adam, whose update differs only in
where epsilon sits (tests/test_cpu_train.py bounds that difference).
"""
import numpy as np

f32 = np.float32


def adam_step(param, grad, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    """One in-place ApplyAdam update on float32 arrays; ``step`` is 1-based.  Returns (param, m, v)."""
    lr, beta1, beta2, eps = f32(lr), f32(beta1), f32(beta2), f32(eps)
    b1p = f32(np.power(beta1, f32(step), dtype=f32))
    b2p = f32(np.power(beta2, f32(step), dtype=f32))
    alpha = f32(lr * np.sqrt(f32(1) - b2p, dtype=f32) / (f32(1) - b1p))
    g = grad.astype(f32)
    m += (g - m) * (f32(1) - beta1)
    v += (g * g - v) * (f32(1) - beta2)
    param -= (m * alpha) / (np.sqrt(v, dtype=f32) + eps)
    return param, m, v


def cosine_decay(lr0, decay_steps, step):
    step = min(f32(step), f32(decay_steps))
    frac = f32(step / f32(decay_steps))
    cosine = f32(f32(0.5) * (f32(1.0) + np.cos(f32(np.pi) * frac, dtype=f32)))
    return f32(f32(lr0) * cosine)
