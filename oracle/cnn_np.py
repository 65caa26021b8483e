"""CPU restatement of the reference's MV-CNN (test infrastructure only; raynet/models.py:90-111).

Keras semantics spelled out: `Conv2D(32, 3)` = 'valid' cross-correlation (no kernel flip) on
channels-last tensors with kernel layout [kh][kw][cin][cout] plus bias; `BatchNormalization` at
inference = gamma * (x - moving_mean) / sqrt(moving_variance + 1e-3) + beta; ReLU after the first
four of the five blocks.  Everything in float64 (the comparison tolerance covers float32
accumulation).  "parity unpinned": TensorFlow / Keras are not installed here and the reference has
no test or fixture for its CNN; tests/test_oracle_pinning.py pins this file's convolution against an
explicit loop restatement and the whole network against the same blocks written with torch's float64
conv2d / batch_norm / relu (an independent second implementation, not the reference).
"""
import numpy as np


def conv3x3_valid(x, kernel, bias):
    """x [N,H,W,Cin], kernel [3,3,Cin,Cout] -> [N,H-2,W-2,Cout]"""
    x = np.asarray(x, np.float64)
    win = np.lib.stride_tricks.sliding_window_view(x, (3, 3), axis=(1, 2))      # [N,H-2,W-2,Cin,3,3]
    return np.einsum("nhwcyx,yxco->nhwo", win, np.asarray(kernel, np.float64)) + np.asarray(bias, np.float64)


def simple_cnn_forward(x, weights, epsilon=1e-3):
    """weights in the Keras order of SimpleCNN.get_weights()."""
    x = np.asarray(x, np.float64)
    n_layers = len(weights) // 6
    for l in range(n_layers):
        k, b, g, be, mu, var = [np.asarray(w, np.float64) for w in weights[6 * l:6 * l + 6]]
        x = conv3x3_valid(x, k, b)
        x = g * (x - mu) / np.sqrt(var + epsilon) + be
        if l < n_layers - 1:
            x = np.maximum(x, 0.0)
    return x
