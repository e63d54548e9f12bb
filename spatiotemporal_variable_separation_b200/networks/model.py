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
from .resnet import MLPResnet


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

    def rollout(self, t_int, n_forecast):
        """model.py:78-83 without the decoder.  Returns (t_all, t_codes, t_residuals):
        t_all       internal codes of all n_forecast steps stacked along the batch (step-major) for ONE grouped decode,
        t_codes     the reference's [B, T, ...] fp32 tensor,
        t_residuals list over steps of lists over blocks (reference layout).
        The MLP stepper runs as one fused launch per direction (ops.latent_rollout); the convolutional stepper
        (SST) is a sequence of grouped conv blocks."""
        B = t_int.shape[0]
        if isinstance(self.t_resnet, MLPResnet):
            t0 = _external_codes(t_int)                                        # [B, d] fp32
            codes, res = ops.latent_rollout(t0, self.t_resnet, n_forecast)      # [T,B,d], [nb,T-1,B,d]
            t_all = ops.to_internal(codes.reshape(n_forecast * B, -1))
            residuals = [[res[j, t] for j in range(res.shape[0])] for t in range(n_forecast - 1)]
            return t_all, codes.transpose(0, 1), residuals
        codes, residuals = [t_int], []
        for _ in range(1, n_forecast):
            t_int, t_res = self.t_resnet.step(t_int)
            codes.append(t_int)
            residuals.append([_external_codes(r) for r in t_res])
        t_all = torch.cat(codes, 0) if len(codes) > 1 else codes[0]
        return t_all, _external_codes(torch.stack(codes, 1)), residuals

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
        t_all, t_codes, t_residuals = self.rollout(t_int, n_forecast)
        n_groups = n_forecast
        if extra_t is not None:
            t_all = torch.cat([extra_t, t_all], 0)
            n_groups += 1
        frames = self.decoder.decode_external(s_int, t_all, skip_int, groups=n_groups)
        frames = frames.view(n_groups, B, *frames.shape[1:])
        if extra_t is not None:
            recon, frames = ops.split_first_group(frames)
        forecasts = frames.transpose(0, 1)                                   # [B, T, C, H, W] view
        s_code = _external_codes(s_int)
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
