"""FusedAdamW: torch.optim.AdamW semantics on one multi-tensor CUDA kernel (morec_adamw_multi).

Replaces the optimizer step of the reference loop (run.py:159-162 two parameter groups, :245-247 GradScaler
unscale / inf check / step): all parameters of all groups are covered by ONE kernel launch driven by a device-side
chunk table, with the gradient unscale and the found-inf skip fused in.

GradScaler: the class sets `_step_supports_amp_scaling`, so `scaler.step(optimizer)` (run.py:246) hands over
`optimizer.grad_scale` / `optimizer.found_inf` as DEVICE tensors and the update is skipped on the device -- the
`.item()` host wait of the stock path (GradScaler._maybe_opt_step) disappears.

State layout equals torch.optim.AdamW's (`state[p] = {'step', 'exp_avg', 'exp_avg_sq'}`, `step` a float32 scalar
tensor), so `optimizer.state_dict()` / `load_state_dict()` round-trip through the reference's checkpoints
(data_utils/utils.py:107-114, run.py:194).  The step count lives on the device and advances only when the update
is applied (GradScaler's rule), without a host round trip.
"""
import warnings

import numpy as np
import torch

from . import lib

# FusedAdamW.step declares the `grad_scaler` keyword (see its docstring); torch announces that contract's retirement on
# every call -- when it goes, GradScaler falls back to the grad_scale / found_inf attributes, which step() also honours
warnings.filterwarnings("ignore", message="GradScaler is going to stop passing itself")

_REC = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("s", "<u8"), ("n", "<i4"), ("lr", "<f4"),
                 ("wd", "<f4"), ("b1", "<f4"), ("b2", "<f4"), ("eps", "<f4")], align=True)
assert _REC.itemsize == 64


class FusedAdamW(torch.optim.Optimizer):
    _step_supports_amp_scaling = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._step_dev = None   # ONE device counter shared by every parameter's state['step']
        self._found_inf = None
        self._chunk = None
        self._host = [None, None]     # double-buffered pinned staging of the per-step tensor table ...
        self._host_ev = [None, None]  # ... each guarded by the event of the H2D copy that last read it
        self._flip = 0
        self._shadows = {}            # id(param) -> 16-bit tensor receiving a copy of the updated parameter
        self._shadow_owners = []      # ops.ShadowSet objects whose copies this optimizer keeps fresh

    # ---------------------------------------------------------------------------------------------------------
    def register_shadow(self, param, shadow, owner=None):
        """`shadow` (bf16 / fp16, same shape, 16-byte aligned) receives a copy of `param` after every update: the
        compute-dtype weight the next forward reads, written by the optimizer kernel instead of a separate cast pass.
        `owner` (an ops.ShadowSet) is told after each step that its copies are still fresh."""
        assert shadow.shape == param.shape and shadow.dtype in (torch.bfloat16, torch.float16) and shadow.is_contiguous()
        if any(s.dtype != shadow.dtype for s in self._shadows.values()):
            self._shadows = {}                    # the model switched its 16-bit compute dtype: start over
            self._shadow_owners = []
        self._shadows[id(param)] = shadow
        if owner is not None and all(o is not owner for o in self._shadow_owners):
            self._shadow_owners.append(owner)

    def clear_shadows(self):
        self._shadows = {}
        self._shadow_owners = []

    def _shared_step(self, dev):
        """the device step counter; (re)built from the per-parameter 'step' entries after a load_state_dict"""
        if self._step_dev is None or self._step_dev.device != dev:
            val = 0.0
            for st in self.state.values():
                if "step" in st:
                    val = float(st["step"])
                    break
            self._step_dev = torch.full((), val, device=dev, dtype=torch.float32)
            for st in self.state.values():
                if "exp_avg" in st:
                    st["step"] = self._step_dev
        return self._step_dev

    def state_dict(self):
        """torch.optim.AdamW's layout; every parameter gets its OWN copy of the (shared) step counter, as torch's
        optimizers increment each state['step'] tensor separately after loading it"""
        sd = super().state_dict()
        sd["state"] = {k: {n: (v.detach().clone() if n == "step" and torch.is_tensor(v) else v) for n, v in st.items()}
                       for k, st in sd["state"].items()}       # (new dicts: the packed state aliases the live one)
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._step_dev = None          # re-read the step from the loaded state on the next step()

    def _build_table(self):
        """per-tensor records (pointers change every step because autograd hands out fresh gradient tensors) +
        the prefix sum of chunk counts; ~200 rows, packed with numpy and uploaded with one async copy"""
        if self._chunk is None:
            self._chunk = lib.adamw_chunk_elems()
        rows = []
        dev = None
        for group in self.param_groups:
            lr, wd, (b1, b2), eps = group["lr"], group["weight_decay"], group["betas"], group["eps"]
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if not g.is_contiguous():
                    g = g.contiguous()
                    p.grad = g
                st = self.state[p]
                if "exp_avg" not in st:
                    assert p.dtype == torch.float32 and p.is_cuda and p.is_contiguous()
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                    st["step"] = self._shared_step(p.device)
                dev = p.device
                sh = self._shadows.get(id(p))
                rows.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                             sh.data_ptr() if sh is not None else 0, p.numel(), lr, wd, b1, b2, eps))
        n = len(rows)
        if n == 0:
            return None
        rec = np.zeros(n, dtype=_REC)
        for name, col in zip(_REC.names, zip(*rows)):
            rec[name] = col
        counts = (rec["n"].astype(np.int64) + self._chunk - 1) // self._chunk
        start = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(counts, out=start[1:])
        nbytes = rec.nbytes + start.nbytes
        i = self._flip
        self._flip ^= 1
        if self._host_ev[i] is not None:
            self._host_ev[i].synchronize()          # the copy issued two steps ago has long finished; never blocks in practice
        if self._host[i] is None or self._host[i].numel() < nbytes:
            self._host[i] = torch.empty(nbytes + 4096, dtype=torch.uint8).pin_memory()
        hb = self._host[i].numpy()
        hb[:rec.nbytes] = rec.view(np.uint8)
        hb[rec.nbytes:nbytes] = start.view(np.uint8)
        table = self._host[i][:nbytes].to(dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._host_ev[i] = ev
        return table, rec.nbytes, n, int(start[-1]), dev

    @torch.no_grad()
    def step(self, closure=None, grad_scale=None, found_inf=None, check_finite=False, grad_scaler=None):
        """grad_scale: optional device scalar S, every gradient is divided by it (GradScaler.unscale_ fused);
        found_inf: optional device scalar, a non-zero value skips the update (and the step count);
        check_finite: compute found_inf here (one read-only pass over the gradients) instead of receiving it;
        grad_scaler: `scaler.step(optimizer)` of torch.amp.GradScaler passes itself to optimizers that declare this
        keyword (the amp-aware optimizer contract, torch/amp/grad_scaler.py `step`): unscale, overflow check and the
        skip all run in this optimizer's kernels, and the verdict is handed back for `scaler.update()`.  Without the
        keyword GradScaler first runs its own `_amp_foreach_non_finite_check_and_unscale_` over every gradient
        (read + write of the whole 460 MB gradient set, measured 165 us per BERT-base step) and then supplies
        grad_scale / found_inf as attributes, which this method still honours."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        scaler_state = None
        if grad_scaler is not None and grad_scaler.is_enabled():
            from torch.amp.grad_scaler import OptState
            scaler_state = grad_scaler._per_optimizer_states[id(self)]
            if scaler_state["stage"] is OptState.UNSCALED:       # scaler.unscale_(opt) was called: gradients are true
                vals = list(scaler_state["found_inf_per_device"].values())
                found_inf = vals[0] if len(vals) == 1 else sum(v.to(vals[0].device) for v in vals)
                grad_scale = None
                scaler_state = None
            else:
                grad_scale = grad_scaler._get_scale_async()
                found_inf = None
                check_finite = True
        if grad_scale is None and scaler_state is None and grad_scaler is None:
            grad_scale = getattr(self, "grad_scale", None)
        if found_inf is None and grad_scaler is None:
            found_inf = getattr(self, "found_inf", None)
        built = self._build_table()
        if built is None:
            return loss
        table, off, n_tensors, n_chunks, dev = built
        if check_finite:
            if self._found_inf is None or self._found_inf.device != dev:
                self._found_inf = torch.zeros(1, device=dev)
            self._found_inf.zero_()
            found_inf = self._found_inf
        if found_inf is not None and (found_inf.dtype != torch.float32 or found_inf.device != dev):
            found_inf = found_inf.to(device=dev, dtype=torch.float32)
        if grad_scale is not None and (grad_scale.dtype != torch.float32 or grad_scale.device != dev):
            grad_scale = grad_scale.to(device=dev, dtype=torch.float32)
        p16 = lib.DT_BF16
        for s in self._shadows.values():
            p16 = lib.dtype_code(s)
            break
        lib.adamw_multi(table, table[off:], n_tensors, n_chunks, self._shared_step(dev), grad_scale=grad_scale,
                        found_inf=found_inf, check_finite=check_finite, p16_dtype=p16)
        if scaler_state is not None:          # what GradScaler.update() reads to grow / back off the scale
            scaler_state["found_inf_per_device"] = {dev: found_inf}
        # parameters changed behind torch's version counters: invalidate every weight-shadow cache except the ones
        # this very kernel refreshed (a skipped overflow step leaves parameters AND shadows untouched: still consistent)
        old = lib.PARAM_EPOCH
        lib.PARAM_EPOCH = old + 1
        for o in self._shadow_owners:
            if o.fresh_epoch == old and o.opt is not None and o.opt() is self:
                o.fresh_epoch = old + 1
        return loss

    @property
    def last_found_inf(self):
        return self._found_inf
