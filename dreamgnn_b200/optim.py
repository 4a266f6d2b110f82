"""Gradient clipping + Adam as one fused step (train.py:297-300: `clip_grad_norm_` then `optimizer.step()`).

`FusedAdam` is a `torch.optim.Adam` whose `step()` runs `dg_adam_clip_step_f32` (csrc/optim.cu): two launches for the
whole parameter list instead of ~17 (per-tensor norms, their norm, the clip coefficient, the scaled gradients and the
eight multi-tensor passes of the foreach Adam update) -- ~250 us at the end of every iteration with nothing to overlap,
10 % of a step at the real-dataset shapes. Same update rule, hyper-parameters, `param_groups` and `state_dict` layout
(`step`, `exp_avg`, `exp_avg_sq` per parameter) as torch's Adam, so checkpoints interchange; the step counter and (if
given as a tensor) the learning rate live on the device, so the step is CUDA-graph capturable and a scheduler's change
takes effect under replay.

    opt = FusedAdam(model.parameters(), lr=0.002, weight_decay=1e-5)
    loss.backward()
    opt.clip_and_step(max_norm=1.0)        # == clip_grad_norm_(params, 1.0); opt.step()

There is no CPU path: parameters must be fp32 CUDA tensors.
"""
import torch as th

from . import _lib as L


class FusedAdam(th.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, **kw):
        for k in ('amsgrad', 'maximize', 'differentiable'):
            if kw.pop(k, False):
                raise ValueError('FusedAdam does not implement %s' % k)
        kw.pop('foreach', None)
        kw.pop('fused', None)
        kw['capturable'] = True            # the step counter is a device tensor in every mode
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, **kw)

    # ---- state ----------------------------------------------------------------------------------
    def _group_step(self, group, params):
        """One device step counter per group, shared by the `step` entry of every parameter's state (after
        load_state_dict the entries are separate tensors again: the first one wins and the rest are re-aliased)."""
        shared = None
        for p in group['params']:
            st = self.state.get(p)
            if st and 'step' in st:
                shared = st['step']
                break
        dev = params[0].device
        if shared is None:
            shared = th.zeros((), dtype=th.float32, device=dev)
        elif not (shared.is_cuda and shared.dtype == th.float32 and shared.dim() == 0):
            shared = th.as_tensor(float(shared), dtype=th.float32, device=dev).reshape(())
        for p in params:
            st = self.state[p]
            if st.get('step') is not shared:
                st['step'] = shared
            if 'exp_avg' not in st:
                st['exp_avg'] = th.zeros_like(p, memory_format=th.contiguous_format)
                st['exp_avg_sq'] = th.zeros_like(p, memory_format=th.contiguous_format)
        return shared

    # ---- the step -------------------------------------------------------------------------------
    @th.no_grad()
    def clip_and_step(self, max_norm=0.0):
        """`clip_grad_norm_(parameters, max_norm)` (max_norm <= 0: no clipping) followed by the Adam update. Returns the
        gradient norm before clipping as a device tensor. Clipping is by the norm over ONE parameter group (the
        reference has one); with several groups clip with nn.utils.clip_grad_norm_ and call `step()`."""
        lib = L.load()
        norm = None
        groups = [g for g in self.param_groups if any(p.grad is not None for p in g['params'])]
        if max_norm and max_norm > 0 and len(groups) > 1:
            raise ValueError('FusedAdam.clip_and_step clips by the global norm of ONE parameter group; '
                             'use nn.utils.clip_grad_norm_ + step() with several groups')
        for group in groups:
            params = [p for p in group['params'] if p.grad is not None]
            for p in params:
                if not (p.is_cuda and p.dtype == th.float32 and p.is_contiguous()):
                    raise RuntimeError('FusedAdam needs contiguous fp32 CUDA parameters (no CPU path)')
                if p.grad.is_sparse:
                    raise RuntimeError('FusedAdam does not support sparse gradients')
                if not p.grad.is_contiguous() or p.grad.dtype != th.float32:
                    p.grad = p.grad.to(th.float32).contiguous()
            step = self._group_step(group, params)
            n = len(params)
            arr = (L.AdamTensor * n)()
            for i, p in enumerate(params):
                st = self.state[p]
                a = arr[i]
                a.param, a.grad, a.exp_avg, a.exp_avg_sq, a.numel = (p.data_ptr(), p.grad.data_ptr(), st['exp_avg'].data_ptr(),
                                                                     st['exp_avg_sq'].data_ptr(), p.numel())
            dev = params[0].device
            lr = group['lr']
            lr_dev = None
            if isinstance(lr, th.Tensor):
                if not (lr.is_cuda and lr.dtype == th.float32):
                    raise RuntimeError('a tensor learning rate must be a float32 CUDA tensor')
                lr_dev, lr = lr, 0.0
            b1, b2 = group['betas']
            ws = L.workspace(lib.dg_adam_workspace_bytes(arr, n), dev)
            norm = th.empty((), dtype=th.float32, device=dev)
            L.check(lib.dg_adam_clip_step_f32(arr, n, L.ptr(step), L.ptr(lr_dev), float(lr), float(b1), float(b2),
                                              float(group['eps']), float(group['weight_decay']), float(max_norm or 0.0),
                                              L.ptr(norm), L.ptr(ws), ws.numel(), L.stream()), 'adam_clip_step')
        return norm

    def step(self, closure=None):
        loss = None
        if closure is not None:
            with th.enable_grad():
                loss = closure()
        self.clip_and_step(0.0)
        return loss
