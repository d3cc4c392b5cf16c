"""FusedAdamW: torch.optim.AdamW semantics on one multi-tensor CUDA kernel (morec_adamw_multi).

Replaces the optimizer step of the reference loop (run.py:159-162 two parameter groups, :245-247 GradScaler
unscale / inf check / step): all parameters of all groups are covered by ONE kernel launch driven by a device-side
chunk table, with the gradient unscale and the found-inf skip fused in.
"""
import ctypes

import torch

from . import lib

_CHUNK = 65536


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._table = None
        self._table_key = None
        self._n_chunks = 0
        self._step = 0
        self._found_inf = None

    def _build_table(self):
        entries = []
        key = []
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                assert p.dtype == torch.float32 and p.is_cuda and p.is_contiguous()
                g = p.grad
                if not g.is_contiguous():
                    g = g.contiguous()
                    p.grad = g
                st = self.state[p]
                if "exp_avg" not in st:
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                key.append((p.data_ptr(), g.data_ptr(), group["lr"], group["weight_decay"]))
                n = p.numel()
                shadow = st.get("bf16_shadow")
                for off in range(0, n, _CHUNK):
                    m = min(_CHUNK, n - off)
                    entries.append(lib.AdamChunk(p.data_ptr() + 4 * off, g.data_ptr() + 4 * off,
                                                 st["exp_avg"].data_ptr() + 4 * off, st["exp_avg_sq"].data_ptr() + 4 * off,
                                                 (shadow.data_ptr() + 2 * off) if shadow is not None else None,
                                                 m, group["lr"], group["weight_decay"]))
        key = tuple(key)
        if key != self._table_key:
            arr = (lib.AdamChunk * len(entries))(*entries)
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8) if entries else torch.empty(0, dtype=torch.uint8)
            self._table = host.cuda()
            self._table_key = key
            self._n_chunks = len(entries)

    @torch.no_grad()
    def step(self, closure=None, inv_scale=None, check_finite=False):
        """inv_scale: optional device scalar multiplying every gradient (GradScaler.unscale_ fused);
        check_finite: set/obey found_inf like GradScaler.step (the update is skipped when a grad is inf/nan)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._build_table()
        if self._n_chunks == 0:
            return loss
        self._step += 1
        g0 = self.param_groups[0]
        if check_finite:
            if self._found_inf is None:
                self._found_inf = torch.zeros(1, device=self._table.device)
            self._found_inf.zero_()
        lib.adamw_multi(self._table, self._n_chunks, g0["betas"][0], g0["betas"][1], g0["eps"], self._step,
                        inv_scale=inv_scale, found_inf=self._found_inf if check_finite else None,
                        check_finite=check_finite)
        return loss

    @property
    def found_inf(self):
        return self._found_inf
