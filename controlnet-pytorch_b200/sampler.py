"""Sampling drivers for the hot path: the timestep loop of tools/sample_ddpm_controlnet.py:43-51 /
tools/sample_ldm_controlnet.py:45-50 with the per-step host work removed.

  * one denoising step (sampler prologue -> ControlNet forward -> fused scheduler update -> step counter bump) is
    captured ONCE into a CUDA graph and replayed for every timestep: t, the scheduler coefficients and the Philox
    step are read from device memory, so there is no H2D copy, no host sync and no CPU RNG inside the loop;
  * noise is Philox keyed by the GLOBAL element index, so the result is independent of how the batch is sharded;
  * data parallel: each rank owns a contiguous slice of the batch, no collective inside the loop, one all_gather of
    the final samples (NCCL over NVLink) at the end.
"""
import os

import torch

from . import ops
from . import runtime as rt

_XT_STEP = (1 << 40)   # Philox "step" reserved for drawing x_T


class DDPMSampler:
    def __init__(self, model, scheduler, seed=0, use_graph=True):
        self.model, self.scheduler, self.seed, self.use_graph = model, scheduler, int(seed), use_graph
        self._graph = None
        self._key = None

    # ---- x_T -------------------------------------------------------------------------------------------
    def draw_xT(self, shape, device, elem_offset=0):
        return ops.philox_normal(shape, device, self.seed, _XT_STEP, elem_offset)

    # ---- eager loop (injected z, parity tests, per-step callbacks) ----------------------------------------
    @torch.no_grad()
    def sample_eager(self, x_T, hint, steps=None, zs=None, elem_offset=0, callback=None):
        sch = self.scheduler
        steps = sch.num_timesteps if steps is None else steps
        table = sch.coef_table(x_T.device)
        xt, x0 = x_T.contiguous(), None
        t_dev = torch.empty((1,), device=x_T.device, dtype=torch.int64)
        for k, t in enumerate(reversed(range(steps))):
            t_dev.fill_(t)
            eps = self.model(xt, t_dev, hint)
            z = zs[k].to(xt.device) if (zs is not None and t > 0) else None
            xt, x0 = ops.sched_step(xt, eps, table[t], z=z, want_x0=True, seed=self.seed, step=k,
                                    elem_offset=elem_offset)
            if callback is not None:
                callback(t, xt, x0)
        return xt, x0

    # ---- graph-replayed loop ----------------------------------------------------------------------------
    def _param_stamp(self):
        """(version, address) of every parameter: a load_state_dict / optimizer step / .to() after capture changes it,
        and the graph (which baked the packed-weight caches derived from the old values) is rebuilt."""
        return tuple((p._version, p.data_ptr()) for p in self.model.parameters())

    def _hint_parts(self):
        return [self._hint[lo:hi] for lo, hi in self._bounds]

    def _refresh_hint(self):
        """The captured step reads the hint FEATURE tensors the warm-up cached (models/_engine.HintCache); after the
        static hint buffer has been overwritten they are recomputed in place, outside the graph, once per call."""
        fn = getattr(self.model, "_hint_feat", None)
        if fn is None:
            return
        mode = rt.get_mode()
        for part in self._hint_parts():
            fn(part, mode)

    def _capture(self, x_T, hint, steps, elem_offset):
        dev = x_T.device
        sch = self.scheduler
        self.xt = x_T.clone().contiguous()
        # the graph reads a hint buffer OWNED by the sampler: callers may free or overwrite theirs
        self._hint = hint.clone().contiguous()
        hint = self._hint
        self.x0 = torch.empty_like(self.xt)
        self.t_seq = torch.arange(steps - 1, -1, -1, device=dev, dtype=torch.int64)
        self.step_idx = torch.zeros((1,), device=dev, dtype=torch.int32)
        self.t_out = torch.zeros((1,), device=dev, dtype=torch.int64)
        self.coef = torch.zeros((6,), device=dev, dtype=torch.float32)
        table = sch.coef_table(dev)
        L = rt.lib()

        # Batch halves on parallel capture streams: every kernel is batch-invariant, so the result
        # is bit-identical to the unsplit step, while the GPU overlaps the MUFU-bound attention of one half with the
        # tensor- / HBM-bound convolutions and GroupNorms of the other (measured at B = 1024: 13.08 -> 12.65 ms).
        B = self.xt.shape[0]
        per = self.xt[0].numel()
        # Default: off.  The ControlNet forward's own fork of its branches (models/_controlnet_common.py) gives the overlap
        # without halving the per-kernel batch: MNIST B = 1024 12.59 split vs 12.61 two branches vs 11.84 ms with the
        # four-stream layout; CelebHQ (190 M parameters) B = 256: 31.2 unsplit, 31.1 two branches, 28.25 split, 27.79 ms
        # four streams (gpu_jobs/r2_34_cfg3.sh).  CNB_SAMPLER_SPLIT=2 keeps the split available.  Never both.
        nsplit = int(os.environ.get("CNB_SAMPLER_SPLIT", "1"))
        nsplit = max(1, min(nsplit, B))
        bounds = [shard_bounds(B, nsplit, k) for k in range(nsplit)]
        self._bounds = bounds
        self._split_streams = [torch.cuda.Stream(device=dev) for _ in range(nsplit)] if nsplit > 1 else []

        def part(lo, hi):
            xs = self.xt[lo:hi]
            eps = self.model(xs, self.t_out, hint[lo:hi])
            ops.sched_step(xs, eps, self.coef, z=None, seed=self.seed, step_dev=self.step_idx,
                           elem_offset=elem_offset + lo * per, out=xs, x0_out=self.x0[lo:hi])

        def one_step():
            rt.check(L.cnb_sampler_prologue(self.step_idx.data_ptr(), self.t_seq.data_ptr(), self.t_out.data_ptr(),
                                            table.data_ptr(), self.coef.data_ptr(), rt.stream()))
            if nsplit == 1:
                part(0, B)
            else:
                cur = torch.cuda.current_stream()
                for st, (lo, hi) in zip(self._split_streams, bounds):
                    st.wait_stream(cur)
                    with torch.cuda.stream(st):
                        part(lo, hi)
                for st in self._split_streams:
                    cur.wait_stream(st)
            rt.check(L.cnb_bump_index(self.step_idx.data_ptr(), 1, rt.stream()))

        from .models import _controlnet_common as _cc
        _cc.set_branch_parallel(False if (nsplit > 1 and os.environ.get("CNB_SPLIT_KEEP_BRANCH", "0") != "1") else None)
        try:
            if nsplit > 1:
                # cold caches: ONE unsplit forward on a single stream packs every weight / builds the t-embedding table
                # before the parts (which only read those caches) run concurrently on the split streams
                self.model(self.xt, self.t_seq[:1], hint)
                torch.cuda.synchronize(dev)
            # warm-up on a side stream: builds weight / hint caches and sets kernel attributes outside the capture
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                one_step()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize(dev)
            c0 = rt.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                one_step()
            self.launches_per_step = rt.launch_count() - c0
        finally:
            _cc.set_branch_parallel(None)
        self._graph = g
        self._nsteps, self._pos = steps, 0
        # keep every tensor the graph reads alive for as long as the graph: the hint features (HintCache may evict)
        cache = getattr(self.model, "_hint_cache", None)
        self._baked = cache.pinned() if cache is not None else []

    @torch.no_grad()
    def sample(self, x_T, hint, steps=None, elem_offset=0):
        """Run t = steps-1 .. 0 (SURVEY.md 3.5 semantics) with fused on-device noise; returns (x_{-1 mean}, x0)."""
        rt.require_cuda(x_T, hint)
        steps = self.scheduler.num_timesteps if steps is None else steps
        if not self.use_graph:
            return self.sample_eager(x_T, hint, steps, None, elem_offset)
        if hint.shape[0] != x_T.shape[0]:
            raise rt.CnbError("sample: x_T and hint must have the same batch size")
        key = (tuple(x_T.shape), tuple(hint.shape), steps, elem_offset, rt.get_mode(), str(x_T.device),
               self._param_stamp())
        if self._graph is None or key != self._key:
            self._capture(x_T, hint, steps, elem_offset)
            self._key = key
        else:
            self._hint.copy_(hint)
            self._refresh_hint()
        self.xt.copy_(x_T)
        self.replay_steps(steps, reset=True)
        out = self.xt.clone(), self.x0.clone()
        check_device_errors()
        return out

    def replay_steps(self, n, reset=True):
        """bench.py hook: replay the captured step n times (state continues from wherever it is)."""
        if reset:
            self.step_idx.zero_()
            self._pos = 0
        for _ in range(n):
            if self._pos >= self._nsteps:      # wrap around the schedule instead of running off the t table
                self.step_idx.zero_()
                self._pos = 0
            self._graph.replay()
            self._pos += 1


class LDMSampler(DDPMSampler):
    """tools/sample_ldm_controlnet.py:21-64: the same timestep loop on VAE latents, then `vae.decode` of the final
    latents only (:51-53), clamp to [-1, 1] and map to [0, 1] (:57-58).  Decoding runs in chunks so that the
    128x128x256-channel decoder activations of a large batch stay bounded (2.1 GB per 256 samples and tensor)."""

    def __init__(self, model, scheduler, vae, seed=0, use_graph=True, decode_chunk=256):
        super().__init__(model, scheduler, seed=seed, use_graph=use_graph)
        self.vae, self.decode_chunk = vae, int(decode_chunk)

    @torch.no_grad()
    def decode(self, latents, to_unit_range=True):
        rt.require_cuda(latents)
        outs = []
        for lo in range(0, latents.shape[0], self.decode_chunk):
            ims = self.vae.decode(latents[lo:lo + self.decode_chunk].contiguous())
            if to_unit_range:
                ims = (torch.clamp(ims, -1., 1.) + 1) / 2
            outs.append(ims)
        check_device_errors()
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    @torch.no_grad()
    def sample_images(self, x_T, hint, steps=None, elem_offset=0, zs=None, to_unit_range=True):
        """Returns (images, final latents).  zs: injected per-step noise (eager loop, parity tests)."""
        if zs is not None:
            xt, _ = self.sample_eager(x_T, hint, steps, zs, elem_offset)
        else:
            xt, _ = self.sample(x_T, hint, steps=steps, elem_offset=elem_offset)
        return self.decode(xt, to_unit_range), xt


class GraphedStudent:
    """Single-step students (tools/sample_consistency_controlnet_distilled.py:60-75,
    tools/sample_distribution_matching_controlnet_distilled.py) replayed from ONE captured CUDA graph per input shape:
    at small batch the forward is ~160 launches of a few microseconds each, so eager Python dispatch (3.9 ms) is the
    whole latency.  The hint block is captured inside the graph (it is part of every call in the reference as well);
    the consistency student's whole-batch early return (`all(sigma <= sigma_min)`, :81-82) is evaluated on the host
    before the replay."""

    def __init__(self, model):
        self.model = model
        self._graphs = {}

    @torch.no_grad()
    def __call__(self, x, t_or_sigma, hint):
        rt.require_cuda(x, hint)
        m = self.model
        cond = torch.as_tensor(t_or_sigma).to(x.device)
        is_sigma = torch.is_floating_point(cond)
        if is_sigma and hasattr(m, "sigma_min") and bool((cond <= m.sigma_min).all()):
            return x
        key = (tuple(x.shape), tuple(hint.shape), tuple(cond.shape), cond.dtype, rt.get_mode(), str(x.device))
        ent = self._graphs.get(key)
        if ent is None:
            sx, sh, sc = x.clone().contiguous(), hint.clone().contiguous(), cond.clone().contiguous()

            def run():
                if hasattr(m, "_hint_cache"):
                    m._hint_cache.clear()              # the hint block must be INSIDE the captured work
                if is_sigma:
                    m._skip_boundary_sync = True
                try:
                    return m(sx, sc, sh)
                finally:
                    if is_sigma:
                        m._skip_boundary_sync = False
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                run()                                   # warm-up: weight caches, kernel attributes
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(x.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = run()
            if hasattr(m, "_hint_cache"):
                m._hint_cache.clear()                  # the cached feature lives in the graph's private pool
            ent = (g, sx, sc, sh, out)
            self._graphs[key] = ent
        g, sx, sc, sh, out = ent
        sx.copy_(x)
        sc.copy_(cond)
        sh.copy_(hint)
        g.replay()
        res = out.clone()
        check_device_errors()
        return res


def check_device_errors():
    """A bounded mbarrier wait that expired makes a tensor-core kernel drain and leave garbage (csrc/tc_common.cuh); the
    latch is read (and cleared) at the end of every public sampling call so that corrupt samples never return rc = 0."""
    if rt.lib().cnb_tc_error_flag() != 0:
        raise rt.CnbError("a tcgen05 pipeline wait timed out on the device (hang guard): results are invalid")


def shard_bounds(total, world, rank):
    """Contiguous slice [lo, hi) of the batch owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@torch.no_grad()
def sample_data_parallel(model, scheduler, shape, hint_fn, steps=None, seed=0, group=None, use_graph=True,
                         gather=True, vae=None, sampler=None, x_T_fn=None):
    """One job over all ranks of `group` (tools/sample_ddpm_controlnet.py:21-51 sharded over the GPUs of a box): rank r
    draws (or receives) and denoises samples [lo, hi) of the global batch `shape[0]`, no collective inside the
    timestep loop, and the final samples are all-gathered ONCE (NCCL over NVLink).
    `hint_fn(lo, hi)` returns this shard's hint tensor on the local device (it may do the host -> device copy);
    `x_T_fn(lo, hi)`, when given, supplies the shard's initial noise the same way (the reference draws x_T on the host,
    :26-29); otherwise x_T is Philox noise keyed by the GLOBAL element index, independent of the number of ranks.
    `sampler` re-uses a DDPMSampler / LDMSampler (and its captured step graph) across calls.  With `vae` the shard's
    final latents are decoded locally (LDM path) and the images are what is gathered."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = shape[0]
    lo, hi = shard_bounds(B, world, rank)
    per = 1
    for d in shape[1:]:
        per *= d
    dev = torch.device("cuda", torch.cuda.current_device())
    smp = sampler
    if smp is None:
        smp = (DDPMSampler(model, scheduler, seed=seed, use_graph=use_graph) if vae is None else
               LDMSampler(model, scheduler, vae, seed=seed, use_graph=use_graph))
    if x_T_fn is not None:
        x_T = x_T_fn(lo, hi)
    else:
        x_T = smp.draw_xT((hi - lo,) + tuple(shape[1:]), dev, elem_offset=lo * per)
    xt, x0 = smp.sample(x_T, hint_fn(lo, hi), steps=steps, elem_offset=lo * per)
    if vae is not None:
        xt = smp.decode(xt)
    if not gather or world == 1:
        return xt
    return gather_shards(xt, B, group)


def gather_shards(x_local, total, group=None):
    """The one exchange step of the job: concatenate every rank's contiguous batch shard (shard_bounds order).
    all_gather_into_tensor when the shards are equal, padded all_gather otherwise.  Backend-agnostic (NCCL on the
    GPUs; the world_size-2 gloo test drives it on CPU tensors)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [b - a for a, b in (shard_bounds(total, world, r) for r in range(world))]
    tail = tuple(x_local.shape[1:])
    x_local = x_local.contiguous()
    if all(n == sizes[0] for n in sizes):
        out = torch.empty((total,) + tail, device=x_local.device, dtype=x_local.dtype)
        dist.all_gather_into_tensor(out, x_local, group=group)
        return out
    mx = max(sizes)
    padded = torch.zeros((mx,) + tail, device=x_local.device, dtype=x_local.dtype)
    padded[: x_local.shape[0]] = x_local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
