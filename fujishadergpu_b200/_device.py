"""Device-array plumbing: accept torch CUDA tensors, CuPy arrays or anything exposing
``__cuda_array_interface__``; allocate results with torch; hand the caller back the array type
it passed in.  PyTorch is used for device memory and streams only -- no torch op is on the
compute path."""
from __future__ import annotations

import torch


def _is_cupy(x) -> bool:
    return type(x).__module__.split(".")[0] == "cupy"


def as_tensor(x) -> torch.Tensor:
    """Zero-copy view of a device array as a torch tensor (raises if it is not on a GPU)."""
    if isinstance(x, torch.Tensor):
        t = x
    elif hasattr(x, "__cuda_array_interface__"):
        t = torch.as_tensor(x, device="cuda")
    else:
        raise TypeError(
            f"expected a CUDA device array (torch tensor / CuPy / __cuda_array_interface__), got {type(x)!r}; "
            "fujishadergpu_b200 has no host path")
    if not t.is_cuda:
        raise TypeError("fujishadergpu_b200 kernels need a CUDA tensor (no CPU fallback)")
    return t


def as_f32_2d(x) -> torch.Tensor:
    t = as_tensor(x)
    if t.ndim != 2:
        raise ValueError(f"expected a 2-D block, got shape {tuple(t.shape)}")
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    if t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


def like_input(result: torch.Tensor, original):
    """Return ``result`` as the caller's array type (CuPy in, CuPy out)."""
    if _is_cupy(original):
        import cupy
        return cupy.asarray(result)
    return result


def stream_ptr(t: torch.Tensor) -> int:
    return int(torch.cuda.current_stream(t.device).cuda_stream)


_TORCH_DTYPE = {"float32": torch.float32, "int16": torch.int16, "uint8": torch.uint8}


def empty_out(t: torch.Tensor, shape, output_dtype: str = "float32") -> torch.Tensor:
    return torch.empty(tuple(int(s) for s in shape), dtype=_TORCH_DTYPE[str(output_dtype)], device=t.device)
