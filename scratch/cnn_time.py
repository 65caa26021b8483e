"""Times the MV-CNN on 9 zero-padded 534x534x3 views (the C3 e2e input): CUDA-core fp32 path vs tcgen05 3xTF32 path."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
out = {}
for tc in (False, True):
    from raynet_b200.models import SimpleCNN
    orig = SimpleCNN.random_init
    SimpleCNN.random_init = classmethod(lambda cls, channels=3, seed=0, trained_like=True, tensor_cores=tc, _o=orig.__func__: _o(cls, channels, seed, trained_like, tensor_cores))
    r = bench.cnn_bar(bench.CONFIGS["c3"], torch.device("cuda"), 9)
    SimpleCNN.random_init = orig
    out["tcgen05_3xtf32" if tc else "cuda_cores_fp32"] = {"ms": r["ms"], "tflops": r["tflops"]}
print(json.dumps(out))
