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


_SIDE_STREAMS: dict = {}


def run_concurrently(fns, device, n_streams: int = 3) -> list:
    """Enqueue independent jobs (callables returning device tensors) round-robin on a few side streams that
    fork from and join the current stream: small launches of different jobs overlap on the device.  The jobs'
    results may be used on the current stream afterwards; on a CPU device (tests) the jobs run in order."""
    dev = torch.device(device)
    if dev.type != "cuda" or len(fns) <= 1:
        return [fn() for fn in fns]
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), n_streams)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
    streams = _SIDE_STREAMS[key]
    main = torch.cuda.current_stream(dev)
    for s in streams:
        s.wait_stream(main)
    outs = []
    for i, fn in enumerate(fns):
        with torch.cuda.stream(streams[i % len(streams)]):
            outs.append(fn())
    for s in streams:
        main.wait_stream(s)
    for o in outs:          # the caching allocator must not hand these blocks out again before the main stream is done
        if isinstance(o, torch.Tensor) and o.is_cuda:
            o.record_stream(main)
    return outs
