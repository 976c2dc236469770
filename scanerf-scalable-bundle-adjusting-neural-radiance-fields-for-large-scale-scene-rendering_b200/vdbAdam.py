"""Drop-in for the reference's vdbAdam.py (the source is missing from the reference checkout;
the API below was recovered from __pycache__/vdbAdam.cpython-38.pyc):
a sparse Adam that only touches entries whose gradient is non-zero, built on
cuda.adam_step_cuda (cuda/adam_kernel.cu:23-94).

Reference quirk kept by default: the bytecode never increments `self.t` and the binding
passes the step by value, so the bias corrections stay those of step 1.  Pass
`bias_correction="standard"` for a counter that advances.  `fused_zero_grad=True` clears the
consumed gradients inside the update kernel, so `zero_grad()` costs nothing afterwards.
"""
import torch

from cuda import adam_step_cuda, adam_step_sparse


class vdbAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0,
                 bias_correction="reference", fused_zero_grad=False):
        self.param_groups = [{"lr": lr, "beta1": betas[0], "beta2": betas[1], "eps": eps,
                              "weight_decay": weight_decay}]
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
        state = {i: {"step": torch.tensor(float(self.t)), "exp_avg": m, "exp_avg_sq": v} for i, (_, m, v) in enumerate(self.params)}
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
        if self._clean:           # gradients were cleared by the fused update
            self._clean = False
            return
        for p, _, _ in self.params:
            if p.grad is not None:
                p.grad.fill_(0)

    def step(self):
        g = self.param_groups[0]
        if self.bias_correction == "standard":
            self.t += 1
        for p, m, v in self.params:
            if p.grad is None:
                continue
            if self.fused_zero_grad or self.bias_correction == "standard" or p.dim() != 2:
                adam_step_sparse(p.data, p.grad, m, v, g["lr"], g["beta1"], g["beta2"], g["eps"],
                                 max(self.t, 1), zero_grad=self.fused_zero_grad)
            else:
                adam_step_cuda(p.data, p.grad, m, v, g["lr"], g["beta1"], g["beta2"], g["eps"], self.t)
        self._clean = self.fused_zero_grad
