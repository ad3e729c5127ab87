"""GaussianDiffusion with the reference's call surface, sampling on the B200 kernels.

Reference: diff_model.py:269-484 (schedules, tables, q_*/p_* helpers, `sample`, `ddim_sample`).
Host-side table construction is float64 torch on the CPU exactly as the reference does; the hot
loops (`ddim_sample`, `sample`) run UNetEngine + the fused update kernels, captured in a CUDA graph.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn.functional as F
from tqdm import tqdm

from . import _capi as capi
from ._model import UNetModelBase


def linear_beta_schedule(timesteps):                      # dm1:269-273
    scale = 1000 / timesteps
    return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)


def cosine_beta_schedule(timesteps, s=0.008):             # dm1:275-285
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps, dtype=torch.float64)
    acp = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    acp = acp / acp[0]
    betas = 1 - (acp[1:] / acp[:-1])
    return torch.clip(betas, 0, 0.999)


def ddim_timestep_tables(timesteps, ddim_timesteps, method):
    """(seq, prev_seq) of dm1:428-440, including the `T % n != 0` over-long table quirk."""
    if method == 'uniform':
        c = timesteps // ddim_timesteps
        seq = np.asarray(list(range(0, timesteps, c)))
    elif method == 'quad':
        seq = ((np.linspace(0, np.sqrt(timesteps * .8), ddim_timesteps)) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{method}"')
    seq = seq + 1
    prev = np.append(np.array([0]), seq[:-1])
    return seq, prev


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class GaussianDiffusionBase:
    _DEFAULT_SCHEDULE = 'cosine'

    def __init__(self, timesteps=1000, beta_schedule=None):
        if beta_schedule is None:
            beta_schedule = self._DEFAULT_SCHEDULE
        self.timesteps = timesteps
        if beta_schedule == 'linear':
            betas = linear_beta_schedule(timesteps)
        elif beta_schedule == 'cosine':
            betas = cosine_beta_schedule(timesteps)
        else:
            raise ValueError(f'unknown beta schedule {beta_schedule}')
        self.betas = betas
        self.alphas = 1. - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, axis=0)
        self.alphas_cumprod_prev = F.pad(self.alphas_cumprod[:-1], (1, 0), value=1.)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = torch.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = torch.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = self.betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = torch.log(
            torch.cat([self.posterior_variance[1:2], self.posterior_variance[1:]]))
        self.posterior_mean_coef1 = self.betas * torch.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = ((1.0 - self.alphas_cumprod_prev) * torch.sqrt(self.alphas)
                                     / (1.0 - self.alphas_cumprod))
        self.use_cuda_graph = True

    # ---- small table helpers (API surface of dm1:334-382; tensors in, tensors out) ----
    def _extract(self, a, t, x_shape):
        out = a.to(t.device).gather(0, t).float()
        return out.reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))

    def q_sample(self, x_start, t, noise=None):
        if noise is None:
            noise = torch.randn_like(x_start)
        return (self._extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
                + self._extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    def q_mean_variance(self, x_start, t):
        mean = self._extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
        variance = self._extract(1.0 - self.alphas_cumprod, t, x_start.shape)
        log_variance = self._extract(self.log_one_minus_alphas_cumprod, t, x_start.shape)
        return mean, variance, log_variance

    def q_posterior_mean_variance(self, x_start, x_t, t):
        mean = (self._extract(self.posterior_mean_coef1, t, x_t.shape) * x_start
                + self._extract(self.posterior_mean_coef2, t, x_t.shape) * x_t)
        var = self._extract(self.posterior_variance, t, x_t.shape)
        logvar = self._extract(self.posterior_log_variance_clipped, t, x_t.shape)
        return mean, var, logvar

    def predict_start_from_noise(self, x_t, t, noise):
        return (self._extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t
                - self._extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * noise)

    def p_mean_variance(self, model, x_t, t, clip_denoised=True):
        pred_noise = model(x_t, t)
        x_recon = self.predict_start_from_noise(x_t, t, pred_noise)
        if clip_denoised:
            x_recon = torch.clamp(x_recon, min=-1., max=1.)
        return self.q_posterior_mean_variance(x_recon, x_t, t)

    @torch.no_grad()
    def p_sample(self, model, x_t, t, clip_denoised=True):
        mean, _, logvar = self.p_mean_variance(model, x_t, t, clip_denoised=clip_denoised)
        noise = torch.randn_like(x_t)
        nonzero = (t != 0).float().view(-1, *([1] * (len(x_t.shape) - 1)))
        return mean + nonzero * (0.5 * logvar).exp() * noise

    # ---- per-step coefficient rows, evaluated with the reference's fp32 op order on the host ----
    def ddim_coefficients(self, seq, prev, n, eta):
        """rows i = n-1..0 (execution order) of [sqrt(1-a_t), sqrt(a_t), sqrt(a_p), sqrt(1-a_p-s^2), s, 0,0,0]
        from float64 alphas_cumprod gathered then cast to fp32 (dm1:336, 450-470)."""
        rows = []
        for i in reversed(range(n)):
            a_t = self.alphas_cumprod[int(seq[i])].float()
            a_p = self.alphas_cumprod[int(prev[i])].float()
            sigma = eta * torch.sqrt((1 - a_p) / (1 - a_t) * (1 - a_t / a_p))
            sigma = sigma.float() if torch.is_tensor(sigma) else torch.tensor(float(sigma))
            rows.append(torch.stack([torch.sqrt(1. - a_t), torch.sqrt(a_t), torch.sqrt(a_p),
                                     torch.sqrt(1 - a_p - sigma ** 2), sigma,
                                     torch.zeros(()), torch.zeros(()), torch.zeros(())]))
        return torch.stack(rows).float().contiguous()

    def ddpm_coefficients(self):
        """rows t = T-1..0 of [sqrt_recip_acp, sqrt_recipm1_acp, coef1, coef2, exp(.5*logvar)*(t!=0),0,0,0]
        (dm1:356-395), float64 tables cast to fp32 like `_extract`."""
        T = self.timesteps
        idx = torch.arange(T - 1, -1, -1)
        g = lambda a: a[idx].float()
        std = (0.5 * g(self.posterior_log_variance_clipped)).exp() * (idx != 0).float()
        z = torch.zeros(T)
        return torch.stack([g(self.sqrt_recip_alphas_cumprod), g(self.sqrt_recipm1_alphas_cumprod),
                            g(self.posterior_mean_coef1), g(self.posterior_mean_coef2), std, z, z, z], 1).contiguous()

    # ---- the hot loop ----
    def _run_steps(self, model, shape, ts_exec, coef, kind, clip_denoised, x_T, noise, keep_all, desc):
        """Shared driver of ddim_sample / sample: x <- step(x, model(x, t_i)) for every row of `coef`."""
        B = shape[0]
        device = next(model.parameters()).device
        if device.type != 'cuda':
            raise RuntimeError("advshadow_b200 samplers run on CUDA only (no CPU path); move the model to a GPU")
        n = coef.shape[0]
        step_fn = "advs_ddim_step" if kind == "ddim" else "advs_ddpm_step"
        with torch.cuda.device(device):
            coef_d = coef.to(device)
            step_dev = torch.zeros(1, dtype=torch.int32, device=device)
            if x_T is None:
                x_T = torch.randn(shape, device=device)
            x_T = x_T.to(device=device, dtype=torch.float32)
            if tuple(x_T.shape) != tuple(shape):
                raise ValueError(f"x_T must have shape {tuple(shape)}")
            needs_noise = bool((coef[:, 4] != 0).any())
            if noise is not None:
                noise = [z.to(device=device, dtype=torch.float32).contiguous() for z in noise]
                if len(noise) != n:
                    raise ValueError(f"noise must hold one tensor per step ({n})")
            traj = torch.empty((n,) + tuple(shape), dtype=torch.float32, device=device) if keep_all else None
            native = isinstance(model, UNetModelBase)
            ts_exec_t = torch.as_tensor(np.asarray(ts_exec), dtype=torch.int64, device=device)
            n_elems = int(np.prod(shape))
            if native:
                eng = model.engine(B, shape[2], shape[3])
                x, eps = eng.x, eng.eps
                x.copy_(x_T)
                table = eng.temb_table(ts_exec_t)
                total = table.shape[1]

                def one_step(z_ptr):
                    st = _stream()
                    capi.call("advs_select_row", table.data_ptr(), total, step_dev.data_ptr(), eng.temb_cur.data_ptr(),
                              B, st)
                    eng.run()
                    capi.call(step_fn, x.data_ptr(), eps.data_ptr(), z_ptr, x.data_ptr(), n_elems, coef_d.data_ptr(),
                              step_dev.data_ptr(), 1, 1 if clip_denoised else 0, st)

                graph = None
                if self.use_cuda_graph and not needs_noise and not keep_all and n > 1:
                    # One captured step per (engine, update kind, clip): every per-call quantity (embedding table,
                    # coefficient table, step counter) lives in a persistent device buffer the graph reads, so later
                    # calls with other schedules / step counts only refill the buffers and replay.
                    cache = eng.__dict__.setdefault("_step_graphs", {})
                    gkey = (kind, bool(clip_denoised))
                    st8 = cache.get(gkey)
                    if st8 is None or st8["cap"] < n:
                        cap = max(n, 64)
                        st8 = dict(cap=cap, table=torch.zeros(cap, total, dtype=torch.float32, device=device),
                                   coef=torch.zeros(cap, 8, dtype=torch.float32, device=device),
                                   step=torch.zeros(1, dtype=torch.int32, device=device), graph=None)
                        cache[gkey] = st8
                    st8["table"][:n].copy_(table)
                    st8["coef"][:n].copy_(coef_d)
                    table, coef_d, step_dev = st8["table"], st8["coef"], st8["step"]
                    step_dev.zero_()
                    if st8["graph"] is None:
                        # one eager step on a side stream first (lazy one-time attribute setup), then capture
                        s = torch.cuda.Stream(device=device)
                        s.wait_stream(torch.cuda.current_stream())
                        with torch.cuda.stream(s):
                            one_step(None)
                        torch.cuda.current_stream().wait_stream(s)
                        x.copy_(x_T)
                        step_dev.zero_()
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            one_step(None)
                        st8["graph"] = g
                        x.copy_(x_T)
                        step_dev.zero_()
                    graph = st8["graph"]
                for i in tqdm(range(n), desc=desc, total=n, disable=None):
                    if graph is not None:
                        graph.replay()
                        continue
                    z = None
                    if needs_noise:
                        z = noise[i] if noise is not None else torch.randn(shape, device=device)
                    one_step(z.data_ptr() if z is not None else None)
                    if traj is not None:
                        traj[i].copy_(x)
                out = x.clone()
            else:
                # foreign callable (x, t) -> eps: the model stays PyTorch, the update is the fused kernel
                x = x_T.clone().contiguous()
                for i in tqdm(range(n), desc=desc, total=n, disable=None):
                    t = torch.full((B,), int(ts_exec[i]), device=device, dtype=torch.long)
                    eps = model(x, t).float().contiguous()
                    z = None
                    if needs_noise:
                        z = noise[i] if noise is not None else torch.randn(shape, device=device)
                    capi.call(step_fn, x.data_ptr(), eps.data_ptr(), z.data_ptr() if z is not None else None,
                              x.data_ptr(), n_elems, coef_d.data_ptr(), step_dev.data_ptr(), 1,
                              1 if clip_denoised else 0, _stream())
                    if traj is not None:
                        traj[i].copy_(x)
                out = x
        return out, traj

    @torch.no_grad()
    def ddim_sample(self, model, image_size, batch_size=8, channels=3, ddim_timesteps=50,
                    ddim_discr_method="uniform", ddim_eta=0.0, clip_denoised=True, *,
                    x_T=None, noise=None, return_tensor=False):
        """DDIM sampling (dm1:416-474).  Returns a host numpy array (B,C,H,W) float32 like the reference;
        extension kwargs: `x_T` (start noise), `noise` (list of per-step z for eta>0),
        `return_tensor` (keep the result on the GPU)."""
        seq, prev = ddim_timestep_tables(self.timesteps, ddim_timesteps, ddim_discr_method)
        coef = self.ddim_coefficients(seq, prev, ddim_timesteps, ddim_eta)
        ts_exec = [int(seq[i]) for i in reversed(range(ddim_timesteps))]
        shape = (batch_size, channels, image_size, image_size)
        out, _ = self._run_steps(model, shape, ts_exec, coef, "ddim", clip_denoised, x_T, noise, False,
                                 'sampling loop time step')
        return out if return_tensor else out.cpu().numpy()

    @torch.no_grad()
    def p_sample_loop(self, model, shape, *, x_T=None, noise=None, keep="auto"):
        """DDPM ancestral loop (dm1:397-408).  The reference copies every step to the host; here the
        trajectory stays on the GPU and is fetched once.  keep='all' | 'last' | 'auto' (all if <= 2 GiB)."""
        coef = self.ddpm_coefficients()
        T = self.timesteps
        ts_exec = list(range(T - 1, -1, -1))
        nbytes = T * int(np.prod(shape)) * 4
        keep_all = keep == "all" or (keep == "auto" and nbytes <= (2 << 30))
        out, traj = self._run_steps(model, tuple(shape), ts_exec, coef, "ddpm", True, x_T, noise, keep_all,
                                    'sampling loop time step')
        if keep_all:
            host = traj.cpu().numpy()
            return [host[i] for i in range(T)]
        return _LastOnlyTrajectory(T, out.cpu().numpy())

    @torch.no_grad()
    def sample(self, model, image_size, batch_size=8, channels=3, **kw):
        return self.p_sample_loop(model, shape=(batch_size, channels, image_size, image_size), **kw)

    def train_losses(self, model, x_start, t, *_ignored_device):
        noise = torch.randn_like(x_start)
        x_noisy = self.q_sample(x_start, t, noise=noise)
        return F.mse_loss(noise, model(x_noisy, t))


class _LastOnlyTrajectory(list):
    """len()==T list whose only materialised entry is the final image (all callers read [-1]:
    main.py:125, gen.py:563); other indices raise instead of silently returning junk."""

    def __init__(self, T, last):
        super().__init__([None] * T)
        list.__setitem__(self, T - 1, last)

    def __getitem__(self, i):
        v = super().__getitem__(i)
        if v is None:
            raise IndexError("intermediate DDPM steps were not kept (pass keep='all' to sample())")
        return v
