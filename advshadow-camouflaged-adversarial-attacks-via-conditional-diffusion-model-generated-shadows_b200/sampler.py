"""ShadowSampler: the end-to-end hot path as one reusable object.

  noise x_T  --DDIM-n (UNet eps + fused update per step)-->  x_0  --masked shadow composite-->  shadowed image

The whole trajectory -- n UNet evaluations, n-1 plain updates and one fused "last update + composite" kernel --
is captured ONCE as a single CUDA graph and replayed per batch: every per-step quantity (time-embedding row,
coefficient row) is read from device tables through a device-side step counter that the graph itself resets.
The object owns every device buffer (static addresses), so the graph survives any number of batches and any
weight reload (the engine keeps its own packed copies of the parameters).  Reference call sites it replaces:
GaussianDiffusion.ddim_sample (dm1:416-474) followed by apply_shadow's blend (dm2:642-653, or the blurred-mask
flavours tools/train_shadow.py:244-265 / ddim2/test.py:851-870).
"""
import ctypes as C

import numpy as np
import torch

from . import _capi as capi
from ._diffusion import ddim_timestep_tables
from ._model import UNetModelBase

# apply_shadow flavours of the reference: does the disk mask go through the 5x5 Gaussian before compositing?
SHADOW_FLAVOURS = {"dm2": False,      # ddim2/diff_model2.py:615-654 (intensity 0.33, hard mask)
                   "ts": True,        # tools/train_shadow.py:224-266 (intensity 0.43, blurred mask)
                   "dt": True}        # ddim2/test.py:830-871 (intensity 0.051, blurred mask)


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class ShadowSampler:
    def __init__(self, model: UNetModelBase, diffusion, batch_size, image_size, ddim_timesteps=50,
                 ddim_discr_method="uniform", clip_denoised=True, precision=None, mask_channels=1, use_graph=True,
                 streams=1, shadow_flavour="dm2", graph_scope="trajectory", engine_options=None, pdl="auto", _instance=0,
                 _buffers=None):
        """`streams` > 1 splits the batch into that many independent sub-batches, each with its own engine and
        CUDA stream: the HBM-bound kernels of one sub-batch (GroupNorm apply, stem, ...) then overlap with the
        tensor-bound kernels of the other on the same SMs (they need no shared memory, the conv CTAs need it all).
        `shadow_flavour`: which apply_shadow of the reference the tail reproduces (SHADOW_FLAVOURS).  With the
        generated image injected as the adversarial image the shadow intensity drops out of the result, so the
        flavours differ only in the mask blur.
        `graph_scope`: "trajectory" (one CUDA graph holds all n steps and the composite) or "step" (one graph per
        step, replayed n times from the host; kept for comparison).
        `engine_options`: keyword arguments for UNetModel.engine (wide_prenorm, gemm_operands, ...; A/B measurements).
        `pdl`: programmatic dependent launch between the ~335 kernels of a step (advs_set_pdl): True / False, or
        "auto" = on for latency-bound engines (batch <= 2) in a single-process run (ADVS_PDL=0 / 1 overrides "auto").  Batch 1, 256x256, DDIM-50:
        209.6 -> 197.5 ms; nothing to gain from batch 8 up.  Multi-process runs keep it off: the round-2 `bench.py`
        runs on 2 and 4 GPUs -- the first to combine it with two sub-batch streams and the NCCL exchange -- ended
        after ~690 s without a result (what a 600 s collective time-out would look like), and the GPU budget ended
        before the cause could be isolated."""
        if not isinstance(model, UNetModelBase):
            raise TypeError("ShadowSampler needs an advshadow_b200 UNetModel")
        if streams > 1 and (batch_size % streams or batch_size // streams < 1):
            raise ValueError("batch_size must be a multiple of streams")
        if shadow_flavour not in SHADOW_FLAVOURS:
            raise ValueError(f"shadow_flavour must be one of {sorted(SHADOW_FLAVOURS)}")
        if graph_scope not in ("trajectory", "step"):
            raise ValueError("graph_scope must be 'trajectory' or 'step'")
        self.model, self.gd = model, diffusion
        self.B, self.S, self.n = batch_size, image_size, ddim_timesteps
        self.clip = 1 if clip_denoised else 0
        self.blur = 1 if SHADOW_FLAVOURS[shadow_flavour] else 0
        self.graph_scope = graph_scope
        self.precision = precision
        self.engine_options = dict(engine_options or {})
        import os
        single = int(os.environ.get("WORLD_SIZE", "1") or "1") <= 1
        env = os.environ.get("ADVS_PDL", "")          # an explicit ADVS_PDL=0/1 overrides the "auto" heuristic
        if pdl == "auto":
            self.pdl = (env != "0") if env in ("0", "1") else (batch_size // max(streams, 1) <= 2 and single)
        else:
            self.pdl = bool(pdl)
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("ShadowSampler runs on CUDA only (no CPU path)")
        self.C = model.in_channels
        self.children = []
        if streams > 1:
            with torch.cuda.device(self.device):
                shape = (batch_size, self.C, image_size, image_size)
                self.x_T = torch.zeros(shape, dtype=torch.float32, device=self.device)
                self.clean = torch.zeros(shape, dtype=torch.float32, device=self.device)
                self.fmask = torch.zeros(batch_size, mask_channels, image_size, image_size, dtype=torch.float32,
                                         device=self.device)
                self.centers = torch.zeros(batch_size, 2, dtype=torch.float32, device=self.device)
                self.radii = torch.zeros(batch_size, dtype=torch.float32, device=self.device)
                self.out = torch.zeros(shape, dtype=torch.float32, device=self.device)
                sub = batch_size // streams
                self.side = [torch.cuda.Stream(device=self.device) for _ in range(streams)]
                for i in range(streams):
                    sl = slice(i * sub, (i + 1) * sub)
                    bufs = dict(clean=self.clean[sl], fmask=self.fmask[sl], centers=self.centers[sl], radii=self.radii[sl],
                                out=self.out[sl])
                    with torch.cuda.stream(self.side[i]):
                        self.children.append(ShadowSampler(model, diffusion, sub, image_size, ddim_timesteps,
                                                           ddim_discr_method, clip_denoised, precision, mask_channels,
                                                           use_graph, streams=1, shadow_flavour=shadow_flavour,
                                                           graph_scope=graph_scope, engine_options=engine_options,
                                                           pdl=self.pdl, _instance=i + 1, _buffers=bufs))
                torch.cuda.synchronize(self.device)
            self.eng = self.children[0].eng
            self.launches_per_trajectory = sum(c.launches_per_trajectory for c in self.children)
            return
        with torch.cuda.device(self.device):
            self._instance = _instance
            self.eng = model.engine(batch_size, image_size, image_size, precision=precision, instance=_instance,
                                    **self.engine_options)
            seq, prev = ddim_timestep_tables(diffusion.timesteps, ddim_timesteps, ddim_discr_method)
            self.coef = diffusion.ddim_coefficients(seq, prev, ddim_timesteps, 0.0).to(self.device)
            self.ts = torch.tensor([int(seq[i]) for i in reversed(range(ddim_timesteps))], dtype=torch.int64)
            self.table = self.eng.temb_table(self.ts)
            self._weights_seen = (id(self.eng), self.eng.weights_version)
            self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
            shape = (batch_size, self.C, image_size, image_size)
            if _buffers is not None:      # views into the parent's batch buffers (multi-stream mode)
                self.clean, self.fmask, self.centers, self.radii, self.out = (_buffers[k] for k in (
                    "clean", "fmask", "centers", "radii", "out"))
            else:
                self.clean = torch.zeros(shape, dtype=torch.float32, device=self.device)
                self.fmask = torch.zeros(batch_size, mask_channels, image_size, image_size, dtype=torch.float32,
                                         device=self.device)
                self.centers = torch.zeros(batch_size, 2, dtype=torch.float32, device=self.device)
                self.radii = torch.zeros(batch_size, dtype=torch.float32, device=self.device)
                self.out = torch.zeros(shape, dtype=torch.float32, device=self.device)
            self.n_elems = int(np.prod(shape))
            self.graph = None
            if use_graph:
                self._capture()
        # launches per trajectory: per step select_row + UNet kernels + update + counter advance (the last update is
        # the fused update+composite kernel)
        self.launches_per_trajectory = self.n * (self.eng.n_kernels + 3)

    # ---- one DDIM step; `last` fuses the composite into the update ----
    def _one_step(self, last):
        eng, st = self.eng, _st()
        capi.call("advs_select_row", self.table.data_ptr(), self.table.shape[1], self.step_dev.data_ptr(),
                  eng.temb_cur.data_ptr(), self.B, st)
        eng.run()
        if last:
            capi.call("advs_ddim_step_composite", eng.x.data_ptr(), eng.eps.data_ptr(), eng.x.data_ptr(),
                      self.coef.data_ptr(), self.step_dev.data_ptr(), 1, self.clip, self.clean.data_ptr(),
                      self.centers.data_ptr(), self.radii.data_ptr(), self.fmask.data_ptr(), self.fmask.shape[1],
                      self.blur, self.out.data_ptr(), self.B, self.C, self.S, self.S, st)
        else:
            capi.call("advs_ddim_step", eng.x.data_ptr(), eng.eps.data_ptr(), None, eng.x.data_ptr(), self.n_elems,
                      self.coef.data_ptr(), self.step_dev.data_ptr(), 1, self.clip, st)

    def _trajectory(self):
        self.step_dev.zero_()
        for i in range(self.n):
            self._one_step(i == self.n - 1)

    def _capture(self):
        prev = capi.lib().advs_set_pdl(1 if self.pdl else 0)      # the captured graph keeps this setting
        try:
            self._capture_graphs()
        finally:
            capi.lib().advs_set_pdl(prev)

    def _capture_graphs(self):
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._one_step(False)     # eager once: one-time kernel attribute setup must not be captured
            self._one_step(True)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize(self.device)
        self.step_dev.zero_()
        if self.graph_scope == "trajectory":
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._trajectory()
        else:
            self.graph, self.graph_last = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._one_step(False)
            with torch.cuda.graph(self.graph_last):
                self._one_step(True)
        self.step_dev.zero_()

    def _sync_weights(self):
        """Weights changed since the tables were built (load_state_dict, optimiser step, in-place edit)?  Then the
        engine has re-packed them into its own buffers (model.engine does that) and the time-embedding table is
        recomputed in place; captured graphs only hold engine-owned addresses and stay valid."""
        eng = self.model.engine(self.B, self.S, self.S, precision=self.precision, instance=self._instance,
                                **self.engine_options)
        if eng is not self.eng:       # the model dropped its engine cache (.to() / release_engines()): rebuild
            self.eng = eng
            self.table = eng.temb_table(self.ts)
            self._weights_seen = (id(eng), eng.weights_version)
            if self.graph is not None:
                self._capture()
            return
        if self._weights_seen != (id(eng), eng.weights_version):
            self.table.copy_(eng.temb_table(self.ts))
            self._weights_seen = (id(eng), eng.weights_version)

    def load_x_T(self, x_T):
        """Copy the start noise (device tensor) into the engine state(s)."""
        if self.children:
            sub = self.B // len(self.children)
            for i, c in enumerate(self.children):
                c.eng.x.copy_(x_T[i * sub:(i + 1) * sub], non_blocking=True)
        else:
            self.eng.x.copy_(x_T, non_blocking=True)

    def run_device(self):
        """Trajectory + composite on buffers already resident in HBM (engine state holds x_T on entry)."""
        if self.children:
            main = torch.cuda.current_stream()
            for st, c in zip(self.side, self.children):
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    c.run_device()
            for st in self.side:
                main.wait_stream(st)
            return self.out
        self._sync_weights()
        if self.graph is None:
            prev = capi.lib().advs_set_pdl(1 if self.pdl else 0)
            try:
                self._trajectory()
            finally:
                capi.lib().advs_set_pdl(prev)
        elif self.graph_scope == "trajectory":
            self.graph.replay()
        else:
            self.step_dev.zero_()
            for _ in range(self.n - 1):
                self.graph.replay()
            self.graph_last.replay()
        return self.out

    def set_inputs(self, x_T, clean, fmask, centers, radii, non_blocking=True):
        """Copy one batch (host or device tensors) into the static buffers."""
        if self.children:
            for c in self.children:
                c._sync_weights()     # an engine rebuild moves the state buffer x_T is copied into
            self.x_T.copy_(x_T, non_blocking=non_blocking)
            self.load_x_T(self.x_T)
        else:
            self._sync_weights()
            self.eng.x.copy_(x_T, non_blocking=non_blocking)
        self.clean.copy_(clean, non_blocking=non_blocking)
        self.fmask.copy_(fmask, non_blocking=non_blocking)
        self.centers.copy_(centers, non_blocking=non_blocking)
        self.radii.copy_(radii, non_blocking=non_blocking)

    @torch.no_grad()
    def __call__(self, x_T, clean, fmask, centers, radii, out_host=None):
        """Host in, host out: H2D of the batch, n DDIM steps, composite, D2H of the shadowed images."""
        with torch.cuda.device(self.device):
            self.set_inputs(x_T, clean, fmask, centers, radii)
            out = self.run_device()
            if out_host is None:
                return out.cpu()
            out_host.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out_host
