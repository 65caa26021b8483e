"""Device-array plumbing of the plug-in layer.

Mirrors raynet/cuda_implementations/utils.py:14-25 (`all_arrays_to_gpu`) and the small
part of pycuda.gpuarray the reference's callers rely on (forward_pass.py:646-664):
`.get()`, `.fill()`, slicing, `.gpudata`, `.shape`, `.dtype`, `len()`.
PyTorch provides device memory and streams only.
"""
import numpy as np
import torch

_TORCH_TO_NP = {
    torch.float32: np.dtype(np.float32),
    torch.float64: np.dtype(np.float64),
    torch.int32: np.dtype(np.int32),
    torch.int64: np.dtype(np.int64),
    torch.uint8: np.dtype(np.uint8),
}


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("raynet_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def current_stream_ptr():
    return torch.cuda.current_stream().cuda_stream


class GPUArray(object):
    """A thin pycuda.gpuarray.GPUArray look-alike over a CUDA torch.Tensor."""

    def __init__(self, tensor):
        assert isinstance(tensor, torch.Tensor) and tensor.is_cuda
        self.tensor = tensor

    @property
    def gpudata(self):
        return self.tensor.data_ptr()

    @property
    def ptr(self):
        return self.tensor.data_ptr()

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    @property
    def dtype(self):
        return _TORCH_TO_NP[self.tensor.dtype]

    @property
    def size(self):
        return self.tensor.numel()

    def __len__(self):
        return self.tensor.shape[0]

    def __getitem__(self, idx):
        return GPUArray(self.tensor[idx])

    def __setitem__(self, idx, value):
        if isinstance(value, GPUArray):
            value = value.tensor
        elif isinstance(value, np.ndarray):
            value = torch.from_numpy(np.ascontiguousarray(value)).to(self.tensor.device)
        self.tensor[idx] = value

    def get(self):
        return self.tensor.detach().cpu().numpy()

    def fill(self, value):
        self.tensor.fill_(value)
        return self

    def ravel(self):
        return GPUArray(self.tensor.reshape(-1))

    def reshape(self, *shape):
        return GPUArray(self.tensor.reshape(*shape))

    def copy(self):
        return GPUArray(self.tensor.clone())


def to_gpu(array):
    """pycuda.gpuarray.to_gpu: a fresh device copy of a numpy array."""
    array = np.ascontiguousarray(array)
    return GPUArray(torch.from_numpy(array).to(device()))


def as_gpu(x):
    """numpy -> fresh device copy; torch.Tensor / GPUArray -> wrapped as is."""
    if isinstance(x, GPUArray):
        return x
    if isinstance(x, np.ndarray):
        return to_gpu(x)
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            x = x.to(device())
        return GPUArray(x)
    return x


def all_arrays_to_gpu(f):
    """Decorator to copy all the numpy arrays to the gpu before function invocation
    (raynet/cuda_implementations/utils.py:14-25).  torch tensors are wrapped too."""
    def inner(*args, **kwargs):
        args = [as_gpu(a) for a in args]
        return f(*args, **kwargs)
    inner.__name__ = getattr(f, "__name__", "inner")
    inner.__doc__ = f.__doc__
    return inner


def ptr(x):
    """Device pointer of a GPUArray (must be contiguous)."""
    if x is None:
        return None
    assert x.tensor.is_contiguous(), "device arrays handed to the kernels must be C-contiguous"
    return x.tensor.data_ptr()
