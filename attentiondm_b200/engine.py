"""CUDA-graph sampler engine: one captured denoising step, replayed T times.

The reference's quantizer tables are indexed by a per-module Python counter
(index_seq, utils/quant_util.py:228-229,281).  Here every layer's scale / zero-
point / multiplier rows for all T steps live in ONE [T, total] device table; at
the start of each step `attndm_stage_tables` copies row `step` into a fixed
"current" buffer that every kernel reads, and advances the device-side step
counter -- so a single graph serves all steps with no host work in the loop.
"""
from __future__ import annotations

import os

import torch

from . import ops
from . import rowprog
from .denoising import ddim_coefficients


class SamplerEngine:
    _cache = {}

    @classmethod
    def for_model(cls, model, seq, betas, eta, x_shape):
        key = (id(model), tuple(seq), float(eta), tuple(x_shape), betas.data_ptr())
        eng = cls._cache.get(key)
        if eng is None or not eng.still_valid():
            cls._cache.clear()            # one live engine: graphs pin a lot of memory
            eng = cls(model, seq, betas, eta, x_shape)
            cls._cache[key] = eng
        return eng

    def __init__(self, model, seq, betas, eta, x_shape, use_graph=True):
        self.model = model
        self.seq = list(seq)
        self.T = len(self.seq)
        self.eta = float(eta)
        self.B, self.C, self.H, self.W = x_shape
        dev = next(model.parameters()).device
        self.dev = dev
        self.layers = [m for _, m in model.qconvs()]
        idx = {m.index_seq % max(1, m.args.timesteps) for m in self.layers}
        if len(idx) != 1:
            raise RuntimeError("SamplerEngine: QConv2d.index_seq counters are out of step with each other; "
                               "call model.reset_index_seq() (utils/quant_util.py:228-229 wraps per module)")
        self.start_index = idx.pop()
        if any(m.len_seq != self.T or m.args.timesteps != self.T for m in self.layers):
            raise RuntimeError("SamplerEngine: len(seq) must equal every layer's len_seq and args.timesteps")
        # ---- pack every layer's per-step rows + the DDIM coefficients + t into one table ----
        cols, off = [], 0
        self.slices = []
        for m in self.layers:
            tb = m._tables()
            w = tb["lay"]["width"]
            cols.append(tb["tab"])
            self.slices.append((off, w))
            off += w
        coef = ddim_coefficients(self.seq, betas, self.eta).to(dev)        # [T,8]
        self.coef_off = off
        off += 8
        tcol = coef[:, 5:6].expand(self.T, self.B).contiguous()             # t repeated B times
        self.t_off = off
        bq = (self.B + 3) // 4 * 4
        tpad = torch.zeros(self.T, bq, device=dev)
        tpad[:, :self.B] = tcol
        off += bq
        # the DDIM/t columns are indexed by the sampler step k, the layer tables by index_seq;
        # both advance together, offset by start_index
        layer_tab = torch.cat(cols, dim=1)
        if self.start_index:
            layer_tab = torch.roll(layer_tab, shifts=-self.start_index, dims=0)
        self._table_roll = self.start_index
        self.table = torch.cat([layer_tab, coef, tpad], dim=1).contiguous()
        self.cur = torch.zeros(off, dtype=torch.float32, device=dev)
        self.step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.versions = self._versions()
        self.x_cur = torch.zeros(self.B, self.H, self.W, self.C, device=dev)
        self.x0 = torch.zeros_like(self.x_cur)
        # device-side history rings: every step's x_t and x0 prediction, which generalized_steps returns as lists
        # (functions/denoising.py:34,40).  The DDIM kernel fills slot `step`; run() copies whole chunks to the host
        # on a side stream while later steps compute.
        self.hist_x = torch.empty(self.T, self.B, self.H, self.W, self.C, device=dev)
        self.hist_x0 = torch.empty_like(self.hist_x)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.noise = torch.zeros_like(self.x_cur) if self.eta != 0 else None
        self.ext_noise = False
        self.graph = None
        self.use_graph = use_graph
        self.launches_per_step = 0
        self.launches_per_pass = 0
        # fused per-sample programs (time_mlp of every block; all blocks on 1x1 maps): graph mode only
        self.fused = None
        self.fused_parts = [None]
        if use_graph and hasattr(model, "down_blocks"):
            if hasattr(model, "materialize_lazy_layers"):
                model.materialize_lazy_layers()
            sl = {id(m): s_ for m, s_ in zip(self.layers, self.slices)}
            self.fused_parts = [rowprog.build(model, self.B, sl, dev, T=self.T)]
            self.fused = self.fused_parts[0]
        # The time path (timestep embedding -> time_embed Linears -> every block's time_mlp; models/diffusion.py:
        # 157-161,273-277,347-351) depends on the step alone -- generalized_steps gives every sample of the batch the
        # same t (functions/denoising.py:26) -- so the captured step does not evaluate it B times: at the start of each
        # pass (load_input) it is evaluated ONCE per step for all T steps (two launches of the same kernels, each
        # step's time_mlp convs reading that step's table row), the results become extra table columns, and the step
        # fans the staged row out to the [B, O] tensors its consumers read.  Same kernels, same per-row arithmetic:
        # bit-identical to the per-step evaluation (ATTNDM_HOIST_TIME=0 restores it).
        self.hoist = False
        fp = self.fused
        if (fp is not None and fp.time_all is not None and os.environ.get("ATTNDM_HOIST_TIME", "1") != "0"
                and hasattr(model, "time_embedding")):
            w0 = self.table.shape[1]
            pad = (-w0) % 4
            self.temb_off = w0 + pad
            self.table = torch.cat([self.table, torch.zeros(self.T, pad + fp.temb_cols, device=dev)], dim=1).contiguous()
            self.cur = torch.zeros(self.table.shape[1], dtype=torch.float32, device=dev)
            fp.set_hoisted(self.temb_off)
            self.hoist = True

    def _versions(self):
        return tuple((m._tab_key, m._pack_key) for m in self.layers)

    def still_valid(self):
        try:
            for m in self.layers:
                m._tables()
        except Exception:
            return False
        # the fused trunk bakes the scalar parameters of every MixedPrecisionAttention into its ops: stale once
        # update_quantization_params ran (it drops the module's host cache) or the bit width changed
        for fp in self.fused_parts:
            plan = fp.trunk_plan if fp is not None else None
            for mpa, vals in (plan.attn_params if plan is not None else []):
                if mpa._host is None or float(mpa.num_heads) != vals[0]:
                    return False
                eff = mpa.get_effective_bits(None)
                if float(max(4, int(eff)) if eff <= 6 else 0) != vals[4] or float(max(3, int(eff - 1)) if eff <= 4 else 0) != vals[7]:
                    return False
        return self._versions() == self.versions

    # ---- one denoising step on the current stream ----
    def _set_fused(self, fp):
        self.model._fused, self.model._fused_cur = fp, self.cur
        for b in self._blocks:
            b._temb_fused = fp.temb.get(id(b)) if fp is not None else None

    def _forward_part(self, part=0):
        """UNet forward + DDIM update of the batch, on the current stream."""
        self._set_fused(self.fused_parts[part])
        x = self.x_cur
        eps = self.model.forward_nhwc(x, self.cur[self.t_off:self.t_off + self.B])
        ops.ddim_step(x, eps, self.cur[self.coef_off:], self.noise, x_next=x, x0_out=self.x0,
                      hist=(self.hist_x, self.hist_x0, self.step))
        return eps

    # ---- one denoising step on the current stream ----
    def _step_body(self):
        # The noise buffer is refreshed OUTSIDE the captured body (run_loaded / run): a graph that baked in
        # `normal_()` would overwrite a caller-supplied noise_fn's values, and one captured without it would
        # replay stale noise -- the graph must not depend on who fills the buffer.
        ops.stage_tables(self.table, self.step, self.cur, advance=True)
        return self._forward_part(0)

    def _with_staged(self, fn):
        for m, (o, w) in zip(self.layers, self.slices):
            m.use_staged_row(self.cur[o:o + w])
        saved = [m.index_seq for m in self.layers]
        self._blocks = (list(self.model.down_blocks) + list(self.model.up_blocks)) if hasattr(self.model, "down_blocks") else []
        try:
            return fn()
        finally:
            for m, s in zip(self.layers, saved):
                m.use_staged_row(None)
                m.index_seq = s
            if hasattr(self.model, "down_blocks"):
                self._set_fused(None)

    def _capture(self):
        from . import _ffi
        # warm-up on a side stream (lazy channel_proj, cuBLAS workspaces, weight packs), state restored after
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        x_save = self.x_cur.clone()
        with torch.cuda.stream(s):
            for _ in range(2):
                self.step.zero_()
                self._with_staged(self._step_body)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self.x_cur.copy_(x_save)
        self.step.zero_()
        g = torch.cuda.CUDAGraph()
        before = _ffi.launches
        with torch.cuda.graph(g):
            self._with_staged(self._step_body)
        self.launches_per_step = _ffi.launches - before
        self.graph = g
        self.x_cur.copy_(x_save)
        self.step.zero_()

    def load_input(self, x):
        """x: logical NCHW tensor (CUDA, or pinned host memory for the end-to-end path)."""
        self.x_cur.copy_(x.permute(0, 2, 3, 1), non_blocking=True)
        self.step.zero_()
        self.start_index = self.layers[0].index_seq % self.T
        self.done = 0
        if self.start_index != self._table_roll:
            raise RuntimeError("SamplerEngine: index_seq moved since the engine was built; rebuild it "
                               "(SamplerEngine.for_model) or call model.reset_index_seq()")
        if self.hoist:
            from . import _ffi
            before = _ffi.launches
            self._time_path_all_steps()
            self.launches_per_pass = _ffi.launches - before       # our kernels that run once per pass, outside the graph

    def _time_path_all_steps(self):
        """The time path of every sampler step of this pass (see __init__), on the current stream."""
        fp = self.fused
        t_all = self.table[:, self.t_off].contiguous()                      # t of step k = table[k, t_off]
        te = self.model.time_embedding(t_all)                               # [T, 1, 1, ted4]
        ted4 = te.shape[-1]
        te2 = te.view(self.T, 1, ted4).expand(self.T, 2, ted4).reshape(2 * self.T, ted4).contiguous()
        cols = fp.run_time_all(te2, self.table)
        self.table[:, self.temb_off:self.temb_off + fp.temb_cols].copy_(cols)

    def run_loaded(self, steps=None):
        """Replay `steps` (default T) denoising steps on the already loaded input; async."""
        n = self.T if steps is None else steps
        if n <= 0:
            return
        fresh_noise = self.noise is not None and not self.ext_noise      # eta > 0, engine-generated noise
        if self.use_graph:
            if self.graph is None:
                self._capture()
            for _ in range(n):
                if fresh_noise:
                    self.noise.normal_()
                self.graph.replay()
        else:
            for _ in range(n):
                if fresh_noise:
                    self.noise.normal_()
                self._with_staged(self._step_body)
        # mirror the reference's per-module counter: wrap at the START of a call, +1 at its end
        self.done = getattr(self, "done", 0) + n
        final = (self.start_index + self.done - 1) % self.T + 1
        for m in self.layers:
            m.index_seq = final

    def run(self, x, keep="all", noise_fn=None):
        """Reference-shaped result: (xs, x0_preds) lists of CPU tensors."""
        if tuple(x.shape) != (self.B, self.C, self.H, self.W):
            raise RuntimeError("SamplerEngine: input shape differs from the captured one")
        self.ext_noise = noise_fn is not None
        if self.ext_noise and self.noise is None:
            self.noise = torch.zeros_like(self.x_cur)
            self.graph = None
        main = torch.cuda.current_stream()
        main.wait_stream(self.copy_stream)                 # a previous run's history copies have left the rings
        if self.use_graph and self.graph is None:
            self.load_input(x)
            self._capture()
        self.load_input(x)
        T = self.T
        if keep == "last":
            if self.ext_noise:
                for k in range(T):
                    self.noise.copy_(noise_fn(k, ops.to_nchw(self.x_cur)).permute(0, 2, 3, 1))
                    self.run_loaded(1)
            else:
                self.run_loaded()
            hx, hx0 = _pinned_copy(self.x_cur), _pinned_copy(self.x0)
            main.synchronize()
            return [x, ops.to_nchw(hx)], [ops.to_nchw(hx0)]
        # keep == "all": ONE pinned block per call (fresh tensors for the caller, like the reference's .to('cpu')),
        # filled chunk by chunk from the history rings on the copy stream, overlapped with the following steps
        host = torch.empty((2,) + tuple(self.hist_x.shape), dtype=torch.float32, device="cpu", pin_memory=True)
        chunk = 1 if self.ext_noise else max(1, min(16, T // 6))
        k0 = 0
        while k0 < T:
            n = min(chunk, T - k0)
            if self.ext_noise:
                self.noise.copy_(noise_fn(k0, ops.to_nchw(self.x_cur)).permute(0, 2, 3, 1))
            self.run_loaded(n)
            ev = torch.cuda.Event()
            ev.record(main)
            self.copy_stream.wait_event(ev)
            with torch.cuda.stream(self.copy_stream):
                host[0, k0:k0 + n].copy_(self.hist_x[k0:k0 + n], non_blocking=True)
                host[1, k0:k0 + n].copy_(self.hist_x0[k0:k0 + n], non_blocking=True)
            k0 += n
        self.copy_stream.synchronize()
        main.synchronize()
        xs = [x] + [ops.to_nchw(host[0, k]) for k in range(T)]
        return xs, [ops.to_nchw(host[1, k]) for k in range(T)]


def _pinned_copy(t):
    h = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
    h.copy_(t, non_blocking=True)
    return h
