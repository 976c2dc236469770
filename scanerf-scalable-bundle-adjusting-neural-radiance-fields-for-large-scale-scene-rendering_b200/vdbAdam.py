"""Drop-in for the reference's vdbAdam.py (the source is missing from the reference checkout;
the API below was recovered from __pycache__/vdbAdam.cpython-38.pyc):
a sparse Adam that only touches entries whose gradient is non-zero, built on
cuda.adam_step_cuda (cuda/adam_kernel.cu:23-94).

Reference quirk kept by default: the bytecode never increments `self.t` and the binding
passes the step by value, so the bias corrections stay those of step 1.  Pass
`bias_correction="standard"` for a counter that advances.  `fused_zero_grad=True` clears the
consumed gradients inside the update kernel, so `zero_grad()` costs nothing afterwards.

`with opt.table_backward(fused=True): loss.backward()` goes one step further for tables that are encoded
through the fused field kernels: the encode backward scatters each L2-resident slice of the table gradient
into a 64 MiB scratch and applies this optimiser's update to the slice right away
(snrf_field_encode_bwd_adam), so the 2 GiB gradient table never exists; `step()` afterwards only advances
the counter (and updates any parameter that was not reached that way).  Needs exactly ONE encode of the
table per backward (Adam of a sum is not a sum of Adams): a second one raises.
"""
import torch

from cuda import adam_step_cuda, adam_step_sparse
from hashgrid import _gradmode


class vdbAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0,
                 bias_correction="reference", fused_zero_grad=False):
        if weight_decay != 0.0:
            # the reference's kernel has no decay term either (cuda/adam_kernel.cu:23-69); refuse instead of ignoring
            raise ValueError("vdbAdam: weight_decay is not supported (the sparse update has no decay term)")
        self.param_groups = [{"lr": lr, "beta1": betas[0], "beta2": betas[1], "eps": eps,
                              "weight_decay": 0.0}]
        self._stepped = False
        self._applied = set()        # id(p) of parameters already updated inside this step's backward (fused mode)
        self._scratch = None
        # gradient scratch of the in-backward update: one slice of 2^23 entries = 64 MiB stays L2-resident
        self.scratch_log2 = 23
        self._grad_versions = {}
        self.t = 0
        self.bias_correction = bias_correction
        self.fused_zero_grad = fused_zero_grad
        self._clean = False
        self.params = [[p, torch.zeros_like(p), torch.zeros_like(p)] for p in params]

    def toCPU(self):
        for item in self.params:
            item[1], item[2] = item[1].cpu(), item[2].cpu()

    def toGPU(self, device="cuda"):
        for item in self.params:
            item[1], item[2] = item[1].to(device), item[2].to(device)

    # Checkpoint interchange with the dense torch.optim.Adam the reference keeps for the table (tile.py:301-303,
    # 562): same structure and moment names, so either side loads the other's `featureGrid_optimizer` entry.
    def state_dict(self):
        g = self.param_groups[0]
        # `step` = the number of updates whose bias correction has been applied: in the default "reference" mode the
        # counter never advances (the reference's quirk) and every update uses the step-1 corrections, so 1 is exported
        # once any update ran -- a torch.optim.Adam that loads this does not restart from step 0.  The moments are the
        # live tensors (as torch's own state_dict returns them).
        state = {i: {"step": torch.tensor(float(max(self.t, 1) if self._stepped else self.t)), "exp_avg": m, "exp_avg_sq": v}
                 for i, (_, m, v) in enumerate(self.params)}
        group = {"lr": g["lr"], "betas": (g["beta1"], g["beta2"]), "eps": g["eps"], "weight_decay": g["weight_decay"],
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "decoupled_weight_decay": False, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        group = sd["param_groups"][0]
        g = self.param_groups[0]
        g["lr"], g["eps"], g["weight_decay"] = group["lr"], group["eps"], group.get("weight_decay", 0.0)
        g["beta1"], g["beta2"] = group["betas"]
        for i, item in enumerate(self.params):
            st = sd["state"].get(i)
            if st is None:
                continue
            item[1] = st["exp_avg"].to(item[0].device, torch.float32).reshape(item[0].shape).contiguous().clone()
            item[2] = st["exp_avg_sq"].to(item[0].device, torch.float32).reshape(item[0].shape).contiguous().clone()
            self.t = int(float(st["step"]))
        self._clean = False

    def zero_grad(self):
        clean, self._clean = self._clean, False
        for p, _, _ in self.params:
            if p.grad is None:
                continue
            # `clean`: the fused update cleared the gradient behind its read.  Trust that only while nothing else has
            # written the tensor since (autograd accumulating a dense gradient from a validation / normal pass bumps
            # its version counter; a replaced .grad tensor has another identity).
            if clean and self._grad_versions.get(id(p)) == (id(p.grad), p.grad._version):
                continue
            p.grad.fill_(0)

    # ---- scatter + update fusion (hashgrid/_field.py: FieldEncodeFn.backward)
    def table_backward(self, fused=True):
        """Context for the training step's `loss.backward()`: the encode backward of this optimiser's tables either
        accumulates straight into `.grad` (fused=False) or applies the update on the spot (fused=True)."""
        self._applied.clear()
        return _gradmode.table_backward("fused", self) if fused else _gradmode.table_backward("direct")

    def owns(self, p):
        return any(q is p for q, _, _ in self.params)

    def begin_fused(self, p, small_levels=0):
        """-> (exp_avg, exp_avg_sq, hyper-parameters, step number, zeroed scratch) for the in-backward update of p.
        Scratch: one L2-resident slice (2^scratch_log2 entries); a table whose levels are larger than that gets one more
        level's worth in front of it when it has `small_levels` sparse coarse levels to reduce in one pass each."""
        if id(p) in self._applied:
            raise RuntimeError("vdbAdam: the table was encoded more than once in this backward; the in-backward update "
                               "needs exactly one encode per step (use table_backward(fused=False))")
        self._applied.add(id(p))
        m, v = next((m, v) for q, m, v in self.params if q is p)
        n = 4
        while n < min(p.numel() // 2, 1 << self.scratch_log2):
            n *= 2
        T = int(p.shape[1]) if p.dim() == 3 else 0
        if small_levels > 0 and T > n:
            n += T
        if self._scratch is None or self._scratch.device != p.device or self._scratch.shape[0] != n:
            self._scratch = torch.zeros(n, 2, dtype=torch.float32, device=p.device)   # stays all-zero between calls
        step = self.t + 1 if self.bias_correction == "standard" else max(self.t, 1)
        return m, v, self.param_groups[0], step, self._scratch

    def step(self):
        g = self.param_groups[0]
        self._stepped = True
        if self.bias_correction == "standard":
            self.t += 1
        for p, m, v in self.params:
            if id(p) in self._applied:      # updated inside the backward (table_backward(fused=True))
                continue
            if p.grad is None:
                continue
            if self.fused_zero_grad or self.bias_correction == "standard" or p.dim() != 2:
                adam_step_sparse(p.data, p.grad, m, v, g["lr"], g["beta1"], g["beta2"], g["eps"],
                                 max(self.t, 1), zero_grad=self.fused_zero_grad)
            else:
                adam_step_cuda(p.data, p.grad, m, v, g["lr"], g["beta1"], g["beta2"], g["eps"], self.t)
        self._applied.clear()
        self._clean = self.fused_zero_grad
        self._grad_versions = {id(p): (id(p.grad), p.grad._version) for p, _, _ in self.params if p.grad is not None}
