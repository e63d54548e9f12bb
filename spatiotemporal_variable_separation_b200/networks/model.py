"""SeparableNetwork (counterpart of /root/reference/var_sep/networks/model.py).

``get_forecast`` keeps the reference signature and return values, but is organised for the GPU:
the latent rollout (strictly sequential, tiny) runs first, then *all* ``n_forecast`` frames are
decoded by ONE pass over a batch of ``n_forecast`` groups.  In training mode BatchNorm statistics
are taken per group and the running-stat EMA is applied group by group, i.e. exactly what
``n_forecast`` sequential ``decoder(...)`` calls do (model.py:74-83).
"""
import torch
import torch.nn as nn

from .. import ops
from .mlp_encdec import MLPEncoder


class SeparableNetwork(nn.Module):

    def __init__(self, Es, Et, t_resnet, decoder, nt_cond, skipco):
        super().__init__()
        assert isinstance(Es, nn.Module)
        assert isinstance(Et, nn.Module)
        assert isinstance(t_resnet, nn.Module)
        assert isinstance(decoder, nn.Module)
        self.Es = Es
        self.Et = Et
        self.decoder = decoder
        self.t_resnet = t_resnet
        self.nt_cond = nt_cond
        self.skipco = skipco
        self.__grad = True

    @property
    def grad(self):
        return self.__grad

    @grad.setter
    def grad(self, grad):
        assert isinstance(grad, bool)
        self.__grad = grad

    # ------------------------------------------------------------------------------------------
    # internal (NHWC, compute dtype) building blocks shared with train.py
    # ------------------------------------------------------------------------------------------
    def encoder_input(self, frames, windows):
        """frames [B,T,C,H,W] fp32; windows = list of first-frame indices -> one batch of
        len(windows) groups, each the nt_cond-frame window folded into channels (conv.py:90)."""
        B = frames.shape[0]
        nt = self.nt_cond
        if isinstance(self.Et, MLPEncoder):
            flat = [frames[:, t0:t0 + nt].reshape(B, -1) for t0 in windows]         # (t,c,h,w) order
            return ops.to_internal(torch.cat(flat, 0) if len(flat) > 1 else flat[0])
        _, T, Cf, H, W = frames.shape
        out = torch.empty((B * len(windows), H, W, nt * Cf), device=frames.device, dtype=ops.compute_dtype())
        for g, t0 in enumerate(windows):
            ops.frames_window(frames, t0, nt, out, g * B)
        return out

    def rollout(self, t_code, n_forecast):
        """model.py:78-83 without the decoder: -> (list of n_forecast codes, residual lists)."""
        codes, residuals = [t_code], []
        for _ in range(1, n_forecast):
            t_code, t_res = self.t_resnet.step(t_code)
            codes.append(t_code)
            residuals.append(t_res)
        return codes, residuals

    # ------------------------------------------------------------------------------------------
    # public API (model.py:52-89)
    # ------------------------------------------------------------------------------------------
    def get_forecast(self, cond, n_forecast, init_t_code=None, init_s_code=None):
        B = cond.shape[0]
        if init_s_code is None:
            s_int = self.Es.encode(self.encoder_input(cond, [0]), 1, self.skipco)
        elif self.skipco:
            s_int = (ops.to_internal(init_s_code[0]), [ops.to_internal(s) for s in init_s_code[1]])
        else:
            s_int = ops.to_internal(init_s_code)
        if self.skipco:
            s_int, skip_int = s_int
        else:
            skip_int = None
        t_int = self.Et.encode(self.encoder_input(cond, [0])) if init_t_code is None else ops.to_internal(init_t_code)
        return self.forecast_internal(s_int, skip_int, t_int, n_forecast, B)

    def forecast_internal(self, s_int, skip_int, t_int, n_forecast, B, extra_t=None):
        """Rollout + one grouped decode.  ``extra_t`` (optional) is an additional T code decoded as
        group 0 *before* the forecast groups (the auto-encoding call of train.py:79-82 precedes
        get_forecast, so its BatchNorm EMA update comes first).  Returns the reference tuple; with
        ``extra_t`` the reconstruction is returned as a fifth element."""
        codes, residuals = self.rollout(t_int, n_forecast)
        groups = list(codes) if extra_t is None else [extra_t] + list(codes)
        t_all = torch.cat(groups, 0) if len(groups) > 1 else groups[0]
        frames = self.decoder.decode_external(s_int, t_all, skip_int, groups=len(groups))
        frames = frames.view(len(groups), B, *frames.shape[1:])
        if extra_t is not None:
            recon, frames = frames[0], frames[1:]
        forecasts = frames.transpose(0, 1)                                   # [B, T, C, H, W] view
        t_codes = _external_codes(torch.stack(codes, 1))                    # [B, T, ...]
        s_code = _external_codes(s_int)
        t_residuals = [[_external_codes(r) for r in step] for step in residuals]
        if extra_t is not None:
            return forecasts, t_codes, s_code, t_residuals, recon
        return forecasts, t_codes, s_code, t_residuals


def _external_codes(h):
    """internal code(s) -> the reference's fp32 layout: [..,1,1,d] -> [..,d]; feature maps -> NCHW."""
    lead = h.shape[:-3]
    x = ops.to_external(h.reshape(-1, *h.shape[-3:]))                        # [n, C, H, W]
    if x.shape[-1] == 1 and x.shape[-2] == 1:
        return x.reshape(*lead, x.shape[1])
    return x.reshape(*lead, *x.shape[1:])
