"""gptst_b200 -- B200-native (sm_100a) implementation of the GPT-ST pre-training hot path.

Public surface (mirrors /root/reference/model/Pretrain_model/GPTST.py):

    from gptst_b200.GPTST import GPTST_Model          # drop-in for model/Pretrain_model/GPTST.py
    from gptst_b200 import ops                         # autograd functions over the C ABI
    from gptst_b200 import _lib                        # ctypes binding of libgptst_b200.so

The heavy blocks run hand-written CUDA kernels behind the C ABI declared in include/gptst_b200.h.
There is no CPU or PyTorch fallback: constructing the ops on a machine without the compiled library,
or calling them with non-CUDA tensors, raises.
"""
__version__ = "0.1.0"
