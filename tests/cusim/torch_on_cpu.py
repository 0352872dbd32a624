"""TEST INFRASTRUCTURE (pytest plugin, `-p torch_on_cpu` with tests/cusim on PYTHONPATH): lets GPU tests that keep their
batches in torch CUDA tensors run against the simulated library -- "device" memory of the stand-in is host memory, so a
CPU tensor's data_ptr() is a valid device pointer there.  `.cuda()`, `.to("cuda")` and `device="cuda"` become no-ops."""
import functools

import torch


def _is_cuda(dev):
    return dev is not None and str(dev).startswith("cuda")


def _strip_device(fn):
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        if _is_cuda(kwargs.get("device")):
            kwargs.pop("device")
        return fn(*args, **kwargs)
    return wrapped


for _name in ("tensor", "full", "empty", "zeros", "ones", "arange", "as_tensor", "empty_like", "zeros_like"):
    setattr(torch, _name, _strip_device(getattr(torch, _name)))

_to = torch.Tensor.to


def _to_cpu(self, *args, **kwargs):
    args = tuple(a for a in args if not (isinstance(a, (str, torch.device)) and _is_cuda(a)))
    if _is_cuda(kwargs.get("device")):
        kwargs.pop("device")
    return _to(self, *args, **kwargs) if (args or kwargs) else self


torch.Tensor.to = _to_cpu
torch.Tensor.cuda = lambda self, *a, **k: self
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.is_available = lambda: True
torch.cuda.device_count = lambda: 1
torch.cuda.set_device = lambda *a, **k: None
