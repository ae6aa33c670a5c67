"""edadm -- host-side plumbing between PyTorch tensors and libedadm.so (the sm_100a kernels).

PyTorch is used for device memory, streams and torch.distributed only; all arithmetic of the
quantized-UNet hot path runs in the hand-written CUDA kernels behind the C ABI (include/edadm.h).
"""
from .native import lib, load_library, EdadmError  # noqa: F401
