"""FusedAdamW: torch.optim.AdamW semantics on one multi-tensor CUDA kernel (morec_adamw_multi).

Replaces the optimizer step of the reference loop (run.py:159-162 two parameter groups, :245-247 GradScaler
unscale / inf check / step): all parameters of all groups are covered by ONE kernel launch driven by a device-side
chunk table, with the gradient unscale and the found-inf skip fused in.
"""
import ctypes

import torch

from . import lib

import numpy as np


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._step = 0
        self._found_inf = None
        self._chunk = None
        self._host = None      # pinned staging buffer for the per-step tensor table

    def _build_table(self):
        """per-tensor records (pointers change every step because autograd hands out fresh gradient tensors) +
        the prefix sum of chunk counts; ~200 rows, packed with numpy and uploaded with one async copy"""
        if self._chunk is None:
            self._chunk = lib.adamw_chunk_elems()
        rows = []
        for group in self.param_groups:
            lr, wd = group["lr"], group["weight_decay"]
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
                sh = st.get("bf16_shadow")
                rows.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                             sh.data_ptr() if sh is not None else 0, p.numel(), lr, wd))
        n = len(rows)
        if n == 0:
            return None
        rec = np.zeros(n, dtype=np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("s", "<u8"),
                                          ("n", "<i4"), ("lr", "<f4"), ("wd", "<f4")], align=True))
        assert rec.itemsize == 56
        cols = list(zip(*rows))
        for name, col in zip(("p", "g", "m", "v", "s", "n", "lr", "wd"), cols):
            rec[name] = col
        counts = (rec["n"].astype(np.int64) + self._chunk - 1) // self._chunk
        start = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(counts, out=start[1:])
        nbytes = rec.nbytes + start.nbytes
        if self._host is None or self._host.numel() < nbytes:
            self._host = torch.empty(nbytes + 4096, dtype=torch.uint8).pin_memory()
        hb = self._host.numpy()
        hb[:rec.nbytes] = rec.view(np.uint8)
        hb[rec.nbytes:nbytes] = start.view(np.uint8)
        dev = self._host[:nbytes].to(self.param_groups[0]["params"][0].device, non_blocking=True)
        return dev, rec.nbytes, n, int(start[-1])

    @torch.no_grad()
    def step(self, closure=None, inv_scale=None, check_finite=False):
        """inv_scale: optional device scalar multiplying every gradient (GradScaler.unscale_ fused);
        check_finite: set/obey found_inf like GradScaler.step (the update is skipped when a grad is inf/nan)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        built = self._build_table()
        if built is None:
            return loss
        table, off, n_tensors, n_chunks = built
        self._step += 1
        g0 = self.param_groups[0]
        if check_finite:
            if self._found_inf is None:
                self._found_inf = torch.zeros(1, device=table.device)
            self._found_inf.zero_()
        lib.adamw_multi(table, table[off:], n_tensors, n_chunks, g0["betas"][0], g0["betas"][1], g0["eps"], self._step,
                        inv_scale=inv_scale, found_inf=self._found_inf if check_finite else None,
                        check_finite=check_finite)
        return loss

    @property
    def found_inf(self):
        return self._found_inf
