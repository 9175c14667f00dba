"""Flat-buffer optimiser for the training loop (reference training.py:21-28,124-136).

The reference averages gradients with one blocking all-reduce per parameter tensor, clips with
``clip_grad_norm_`` and steps ``torch.optim.Adam`` - ~60 small launches of each per step.  Here every
parameter and every gradient is a VIEW into one flat fp32 buffer:

  * the gradient exchange is ONE all-reduce of the flat gradient buffer (no packing copy),
  * the clip coefficient is one device scalar computed from one norm of that buffer (no host sync),
  * the update is ONE kernel, ``car_adam_step`` (include/car_b200.h), with torch.optim.Adam's arithmetic.

``state_dict()`` has torch.optim.Adam's layout (``{'state': {i: {'step','exp_avg','exp_avg_sq'}},
'param_groups': [...]}``), i.e. the ``'optimizer'`` entry of the reference's checkpoints (training.py:119).
"""
import torch
import torch.distributed as dist

from . import _lib


class FlatAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam runs car_adam_step on a CUDA device (no CPU path)")
        assert all(p.device == dev and p.dtype == torch.float32 for p in self.params)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        # every parameter starts on a 256-byte boundary of the flat buffer: the CUDA kernels read weights with
        # 16-byte vector loads and TMA, and torch's own kernels pick vectorised paths by pointer alignment
        ALIGN = 64                                                     # floats
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        self.flat = torch.zeros(n, device=dev)
        self.grad = torch.zeros(n, device=dev)
        self.exp_avg = torch.zeros(n, device=dev)
        self.exp_avg_sq = torch.zeros(n, device=dev)
        self.scale = torch.ones((), device=dev)
        for p, off in zip(self.params, self.offsets):
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)            # the parameter now lives in the flat buffer
            p.grad = self.grad[off:off + k].view_as(p)            # autograd accumulates into this view

    def zero_grad(self):
        self.grad.zero_()
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * off:
                p.grad = self.grad[off:off + p.numel()].view_as(p)

    def step(self, max_grad_norm=None, group=None, average=True):
        """all-reduce (if a process group is initialised) -> clip -> Adam, on the current stream.
        Returns the (pre-clip) gradient norm as a device scalar."""
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if world > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)      # ONE collective (training.py:21-28)
        avg = (1.0 / world) if (world > 1 and average) else 1.0
        norm = torch.linalg.vector_norm(self.grad) * avg
        if max_grad_norm is not None:
            coef = torch.clamp(max_grad_norm / (norm + 1e-6), max=1.0)         # torch.nn.utils.clip_grad_norm_
            self.scale.copy_(coef * avg)
        else:
            self.scale.fill_(avg)
        self.step_count += 1
        lib = _lib.load()
        with torch.cuda.device(self.flat.device):
            _lib.check(lib.car_adam_step(self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                         self.exp_avg_sq.data_ptr(), self.flat.numel(), self.lr, self.betas[0],
                                         self.betas[1], self.eps, self.weight_decay, self.step_count,
                                         self.scale.data_ptr(), torch.cuda.current_stream().cuda_stream), "car_adam_step")
        # the kernel wrote the parameters behind torch's back: bump their version counters so that caches keyed
        # on (data_ptr, _version) - CrossAttentionRenderer._packed_weights - see the update
        torch.autograd.graph.increment_version(self.params)
        return norm

    def state_dict(self):
        state = {}
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            k = p.numel()
            state[i] = {"step": torch.tensor(float(self.step_count)),
                        "exp_avg": self.exp_avg[off:off + k].view_as(p).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + k].view_as(p).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            st = sd["state"].get(i)
            if st is None:
                continue
            k = p.numel()
            self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
            self.step_count = int(st["step"])
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps = g["lr"], tuple(g["betas"]), g["eps"]
