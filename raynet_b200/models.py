"""The MV-CNN feature extractor in front of the hot path.

Mirror of the reference's `create_simple_cnn` (raynet/models.py:90-111): five
`Conv2D(filters=32, kernel_size=3)` ('valid' padding, channels-last) each followed by
`BatchNormalization`, with `Activation("relu")` after the first four.  Views are zero-padded by
`padding` = 11 pixels first (forward_pass.py:181-198), so an (H, W) image gives an
(H + 12, W + 12, 32) feature map.  The object offers the two calls the forward passes use:
`predict(X)` like the Keras model (numpy in, numpy out) and `predict_features(scene, views)`, the
hook of RayNetForwardPass that leaves the feature volume ON THE DEVICE (no 316 MB host round trip
per call on the headline configuration).

Compute (one launch per layer, two code paths selected by `tensor_cores`):
  * raynet_b200/csrc/rn_cnn.cuh through `rn_conv3x3_bn_relu`: direct convolution in fp32 on the CUDA cores;
  * raynet_b200/csrc/rn_cnn_tc.cuh through `rn_conv3x3_bn_relu_tc` (the four 32 -> 32 layers; the 3 -> 32 first
    layer stays on the CUDA cores with a hi / lo epilogue, `rn_conv3x3_bn_relu_split`): implicit-GEMM convolution on
    tcgen05.mma kind::tf32 with 3 x TF32 split products -- float32-level accuracy on the tensor cores.
There is no CPU fallback; weights come from `set_weights` in the Keras order (kernel, bias, gamma, beta,
moving_mean, moving_variance per layer) or from `random_init` (no checkpoints are available offline).
"""
import numpy as np
import torch

from . import _lib
from .cuda_implementations.utils import current_stream_ptr, device

BN_EPSILON = 1e-3      # keras.layers.BatchNormalization default


class SimpleCNN(object):
    n_layers = 5
    filters = 32

    def __init__(self, channels=3, epsilon=BN_EPSILON, tensor_cores=True):
        self.channels = int(channels)
        self.epsilon = float(epsilon)
        self.tensor_cores = bool(tensor_cores)      # 32 -> 32 layers on tcgen05 (3 x TF32) instead of the CUDA cores
        self._layers = None       # list of dict(kernel, bias, gamma, beta, mean, var) numpy float32
        self._dev = None          # list of (kernel, scale, shift) CUDA tensors
        self.launches = 0
        self._scratch = None      # two device buffers for the intermediate layers
        self.last_h2d_bytes = 0   # bytes of zero-padded images uploaded by the last predict_features()
        self._staging = None      # pinned host staging buffer for views whose pixels are pageable
        self._padded = None       # device buffer of the zero-padded views (border zeroed once)

    # ------------------------------------------------------------------ weights
    @classmethod
    def random_init(cls, channels=3, seed=0, trained_like=True, tensor_cores=True):
        """Glorot-uniform kernels (Keras default); with trained_like=True the batch-norm statistics and
        biases are random too, so that every term of the folded affine map is exercised."""
        rng = np.random.default_rng(seed)
        m = cls(channels, tensor_cores=tensor_cores)
        weights = []
        cin = m.channels
        for _ in range(cls.n_layers):
            limit = np.sqrt(6.0 / (9 * cin + 9 * cls.filters))
            weights.append(rng.uniform(-limit, limit, size=(3, 3, cin, cls.filters)).astype(np.float32))
            if trained_like:
                weights += [rng.normal(0, 0.05, cls.filters).astype(np.float32), rng.uniform(0.5, 1.5, cls.filters).astype(np.float32),
                            rng.normal(0, 0.1, cls.filters).astype(np.float32), rng.normal(0, 0.1, cls.filters).astype(np.float32),
                            rng.uniform(0.5, 1.5, cls.filters).astype(np.float32)]
            else:
                weights += [np.zeros(cls.filters, np.float32), np.ones(cls.filters, np.float32), np.zeros(cls.filters, np.float32),
                            np.zeros(cls.filters, np.float32), np.ones(cls.filters, np.float32)]
            cin = cls.filters
        m.set_weights(weights)
        return m

    def set_weights(self, weights):
        """Keras order: per layer kernel [3,3,Cin,32], bias, gamma, beta, moving_mean, moving_variance."""
        assert len(weights) == 6 * self.n_layers
        layers = []
        cin = self.channels
        for l in range(self.n_layers):
            k, b, g, be, mu, var = [np.asarray(w, dtype=np.float32) for w in weights[6 * l:6 * l + 6]]
            assert k.shape == (3, 3, cin, self.filters), k.shape
            layers.append(dict(kernel=k, bias=b, gamma=g, beta=be, mean=mu, var=var))
            cin = self.filters
        self._layers = layers
        self._dev = None

    def get_weights(self):
        out = []
        for L in self._layers:
            out += [L["kernel"], L["bias"], L["gamma"], L["beta"], L["mean"], L["var"]]
        return out

    def _device_weights(self):
        if self._dev is None:
            dev = device()
            self._dev = []
            for L in self._layers:
                scale = L["gamma"].astype(np.float64) / np.sqrt(L["var"].astype(np.float64) + self.epsilon)
                shift = L["beta"].astype(np.float64) + scale * (L["bias"].astype(np.float64) - L["mean"].astype(np.float64))
                k = np.ascontiguousarray(L["kernel"])
                w_cat = None
                if k.shape[2] == self.filters:
                    # tensor-core layers: per tap [cout][cin] (K-major), split into an exactly-TF32 high part (low 13
                    # mantissa bits cleared) and the float32 remainder, stacked as 64 rows per tap (rn_cnn_tc.cuh)
                    w = np.ascontiguousarray(k.transpose(0, 1, 3, 2).reshape(9, self.filters, self.filters))
                    hi = (w.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)
                    lo = (w - hi).astype(np.float32)
                    w_cat = torch.from_numpy(np.ascontiguousarray(np.concatenate([hi, lo], axis=1))).to(dev)
                self._dev.append((torch.from_numpy(k).to(dev), torch.from_numpy(scale.astype(np.float32)).to(dev),
                                  torch.from_numpy(shift.astype(np.float32)).to(dev), w_cat))
        return self._dev

    # ------------------------------------------------------------------ inference
    def predict_device(self, x):
        """x: CUDA float32 [N, Hi, Wi, C] (already zero-padded) -> CUDA float32 [N, Hi-10, Wi-10, 32]."""
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[3] == self.channels
        x = x.contiguous()
        n, h, w = int(x.shape[0]), int(x.shape[1]), int(x.shape[2])
        if h < 2 * self.n_layers + 1 or w < 2 * self.n_layers + 1:
            raise AssertionError("input of %d x %d pixels is too small for five valid 3x3 convolutions" % (h, w))
        cin = self.channels
        st = current_stream_ptr()
        # intermediate layers ping-pong between cached buffers (the first layer's output is the largest; the
        # tensor-core path keeps every intermediate as a hi / lo pair); only the returned feature volume is a fresh tensor
        need = n * (h - 2) * (w - 2) * self.filters
        n_scratch = 4 if self.tensor_cores else 2
        if (self._scratch is None or len(self._scratch) != n_scratch or self._scratch[0].numel() < need
                or self._scratch[0].device != x.device):
            self._scratch = [torch.empty((need,), dtype=torch.float32, device=x.device) for _ in range(n_scratch)]
        x_lo = None
        for l, (k, scale, shift, w_cat) in enumerate(self._device_weights()):
            last = l == self.n_layers - 1
            shape = (n, h - 2, w - 2, self.filters)
            numel = n * (h - 2) * (w - 2) * self.filters
            if last:
                y = torch.empty(shape, dtype=torch.float32, device=x.device)
                y_lo = None
            elif self.tensor_cores:
                y = self._scratch[2 * (l & 1)][:numel].view(shape)
                y_lo = self._scratch[2 * (l & 1) + 1][:numel].view(shape)
            else:
                y = self._scratch[l & 1][:numel].view(shape)
                y_lo = None
            if self.tensor_cores and x_lo is not None:
                _lib.call("rn_conv3x3_bn_relu_tc", x.data_ptr(), x_lo.data_ptr(), w_cat.data_ptr(), scale.data_ptr(),
                          shift.data_ptr(), y.data_ptr(), y_lo.data_ptr() if y_lo is not None else None, n, h, w,
                          0 if last else 1, st)
            elif y_lo is not None:
                _lib.call("rn_conv3x3_bn_relu_split", x.data_ptr(), k.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                          y.data_ptr(), y_lo.data_ptr(), n, h, w, cin, 0 if last else 1, st)
            else:
                _lib.call("rn_conv3x3_bn_relu", x.data_ptr(), k.data_ptr(), scale.data_ptr(), shift.data_ptr(), y.data_ptr(),
                          n, h, w, cin, 0 if last else 1, st)
            self.launches += 1
            x, x_lo, h, w, cin = y, y_lo, h - 2, w - 2, self.filters
        return x

    def predict(self, X, batch_size=None):
        """Keras-style: numpy [N, Hi, Wi, C] -> numpy [N, Hi-10, Wi-10, 32]."""
        X = np.ascontiguousarray(X, dtype=np.float32)
        return self.predict_device(torch.from_numpy(X).to(device())).cpu().numpy()

    def predict_features(self, scene, view_indices, padding=11):
        """RayNetForwardPass hook: feature maps of the given views as ONE CUDA tensor
        [n, H+padding+1, W+padding+1, 32] (zero-padding as forward_pass.py:181-198).  The zero-padded stack
        lives on the device (its border is zeroed once); a view whose pixel buffer is page-locked float32 is
        uploaded straight from it (asynchronous DMA), anything else through a cached pinned staging buffer."""
        images = [scene.get_image(v).image for v in view_indices]
        H, W, C = images[0].shape
        dev = device()
        shape = (len(images), H + 2 * padding, W + 2 * padding, C)
        X = self._padded
        if X is None or tuple(X.shape) != shape or X.device != dev:
            X = self._padded = torch.zeros(shape, dtype=torch.float32, device=dev)
        bytes_up = 0
        for k, im in enumerate(images):
            src = None
            if isinstance(im, np.ndarray) and im.dtype == np.float32 and im.flags["C_CONTIGUOUS"]:
                t = torch.from_numpy(im)
                if t.is_pinned():
                    src = t
            if src is None:
                st = self._staging
                if st is None or tuple(st.shape) != (len(images), H, W, C):
                    st = self._staging = torch.empty((len(images), H, W, C), dtype=torch.float32).pin_memory()
                st[k].numpy()[...] = im
                src = st[k]
            X[k, padding:padding + H, padding:padding + W, :].copy_(src, non_blocking=True)
            bytes_up += src.numel() * 4
        self.last_h2d_bytes = bytes_up
        return self.predict_device(X)


def create_simple_cnn(input_shape=(None, None, 3), seed=0):
    """models.py:90 signature (input_shape = (H, W, C), sizes may be None)."""
    return SimpleCNN.random_init(channels=int(input_shape[-1]), seed=seed, trained_like=False)
