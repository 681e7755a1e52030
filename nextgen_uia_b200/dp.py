"""One-process-per-GPU data-parallel fine-tuning step (NCCL over NVLink 5 / NVSwitch).

The reference trains on a single GPU (`--device cuda:0`, src/models/biomedclip/finetune.py:101) with no
torch.distributed call anywhere, so the data-parallel layer is new:
  * the global batch is partitioned by rank; every rank holds a full replica of the frozen towers and of
    the adapters;
  * InfoNCE all-gathers the L2-normalised image/text features once per micro-step (losses.py) and forms
    the global loss, so feature gradients need no second collective;
  * adapter / LoRA gradients are all-reduced with SUM (the loss is already a mean over the global
    batch) — in per-layer buckets launched from autograd hooks on a side stream so the collective of
    layer i overlaps the backward kernels of layer i-1;
  * gradient clipping uses the post-all-reduce global norm; accumulation stays local to a rank
    (finetune.py:287,296-303 semantics).
"""
import math

import torch
import torch.distributed as dist


class _nvtx:
    """NVTX range around the phases of a step when NGU_NVTX=1; free otherwise."""
    import os as _os
    on = _os.environ.get("NGU_NVTX", "0") == "1"

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if self.on and torch.cuda.is_available():
            torch.cuda.nvtx.range_push("ngu/" + self.name)

    def __exit__(self, *a):
        if self.on and torch.cuda.is_available():
            torch.cuda.nvtx.range_pop()
        return False


def setup_mona(model, variant="baseline", bottleneck=64, num_layers=None):
    """finetune.py:165-177 `_setup_mona_finetuning`: freeze all, inject, thaw names containing 'mona'."""
    from .adapters.mona import inject_mona_variant_to_open_clip
    for p in model.parameters():
        p.requires_grad = False
    inject_mona_variant_to_open_clip(model, variant=variant, bottleneck_dim=bottleneck, num_layers=num_layers)
    for n, p in model.named_parameters():
        if "mona" in n.lower():
            p.requires_grad = True
    return model


def setup_lora(model, r=16, alpha=32, dropout=0.1, num_layers=None):
    """finetune.py:180-197 `_setup_lora_finetuning` (the wrapped projections' biases stay trainable, as in the reference)."""
    from .adapters.lora import inject_lora_to_biomedclip
    for p in model.parameters():
        p.requires_grad = False
    inject_lora_to_biomedclip(model, lora_r=r, lora_alpha=alpha, lora_dropout=dropout, num_layers=num_layers)
    for n, p in model.named_parameters():
        if "lora" in n.lower():
            p.requires_grad = True
    return model


class GradBuckets:
    """One flat fp32 gradient buffer; each bucket (group of parameters, e.g. one transformer layer) is a contiguous
    slice of it with an async SUM all-reduce.

    Parameters' .grad tensors are views into the flat buffer, so the collective runs on one contiguous slice per
    bucket and the fused optimiser walks a single buffer.  With `flatten_params=True` the parameters' storage is
    re-pointed into a matching flat fp32 buffer as well (values preserved)."""

    enabled = True

    def __init__(self, params, bucket_of, group=None, flatten_params=False):
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        order = {}
        for p in params:
            order.setdefault(bucket_of(p), []).append(p)
        self.buckets = order
        self.params = [p for ps in order.values() for p in ps]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat_grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_param = torch.empty(n, device=dev, dtype=torch.float32) if flatten_params else None
        self.flat = {}
        self.pending = {}
        self.launched = set()
        self.works = []
        self.comm_stream = None
        off = 0
        for key, ps in order.items():
            start = off
            for p in ps:
                k = p.numel()
                p.grad = self.flat_grad[off:off + k].view_as(p)
                if flatten_params:
                    with torch.no_grad():
                        self.flat_param[off:off + k].copy_(p.detach().reshape(-1))
                        p.data = self.flat_param[off:off + k].view_as(p)
                off += k
            self.flat[key] = self.flat_grad[start:off]
        self._mark_sinks()
        if self.world > 1:
            if self.params and self.params[0].is_cuda:
                self.comm_stream = torch.cuda.Stream()
            for key, ps in order.items():
                for p in ps:
                    p.register_post_accumulate_grad_hook(self._make_hook(key, len(ps)))

    def _mark_sinks(self):
        """Adapters whose parameters all sit in ONE bucket may accumulate their gradients straight into the flat buffer
        from their own backward (adapters/mona.py MonaFunction) and then call notify() instead of the autograd hooks."""
        for key, ps in self.buckets.items():
            owners = {}
            for p in ps:
                owners.setdefault(getattr(p, "_ngu_owner", None), []).append(p)
            for owner, ops_ in owners.items():
                if owner is not None:
                    for p in ops_:
                        p._ngu_sink = (self, key, len(ops_))

    def notify(self, key, count):
        """`count` parameters of bucket `key` have their gradients accumulated (called from a fused backward)."""
        total = len(self.buckets[key])
        c = self.pending.get(key, 0) + count
        if c >= total:
            self.pending[key] = 0
            if self.world > 1:
                self._launch(key)
        else:
            self.pending[key] = c

    def _make_hook(self, key, count):
        def hook(_p):
            if getattr(_p, "_ngu_sink", None) is not None:
                return  # this parameter's gradient is accumulated by its adapter's own backward (notify())
            c = self.pending.get(key, 0) + 1
            if c == count:
                self.pending[key] = 0
                self._launch(key)
            else:
                self.pending[key] = c
        return hook

    def _launch(self, key):
        if not self.enabled or key in self.launched:
            return
        self.launched.add(key)
        if self.comm_stream is None:  # CPU tensors (gloo): no stream juggling
            self.works.append(dist.all_reduce(self.flat[key], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            self.works.append(dist.all_reduce(self.flat[key], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self):
        for w in self.works:
            w.wait()
        self.works = []
        self.launched.clear()
        if self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)

    def zero(self):
        self.flat_grad.zero_()


def cosine_lr(base_lr, eta_min, t, t_max):
    """closed form of torch.optim.lr_scheduler.CosineAnnealingLR after t scheduler steps (finetune.py:255)."""
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * t / t_max)) / 2.0


class FusedAdamW:
    """clip_grad_norm_ + AdamW + cosine LR + zero_grad in two kernel launches over the flat buffers
    (finetune.py:244-255, :296-303).  State (exp_avg, exp_avg_sq) is flat fp32 like the parameters."""

    def __init__(self, buckets, lr, betas, eps, weight_decay, grad_clip, total_updates, lr_min):
        assert buckets.flat_param is not None, "FusedAdamW needs GradBuckets(flatten_params=True)"
        self.b = buckets
        self.base_lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.grad_clip, self.t_max, self.lr_min = grad_clip, total_updates, lr_min
        self.m = torch.zeros_like(buckets.flat_param)
        self.v = torch.zeros_like(buckets.flat_param)
        self.gsq = torch.zeros(1, device=buckets.flat_param.device, dtype=torch.float32)
        # device-resident loop state (include/ngu_b200.h ngu_guard_tick): updates applied, poison flag, updates skipped,
        # micro-steps run.  The LR schedule and the bias-correction step are derived from it INSIDE the kernel, so a skipped
        # update advances neither (finetune.py:281-288: `continue` also skips scheduler.step()) and a CUDA-graph replay
        # needs nothing from the host.
        self.state = torch.zeros(4, device=buckets.flat_param.device, dtype=torch.int64)

    @property
    def updates(self):
        """updates applied so far (host read: synchronises; for logging / tests)"""
        return int(self.state[0].item())

    @property
    def skipped(self):
        return int(self.state[2].item())

    @property
    def lr(self):
        return cosine_lr(self.base_lr, self.lr_min, self.updates, self.t_max)

    def note_loss(self, loss):
        """poison |= !isfinite(loss): called for EVERY micro-step, so a bad micro-batch early in an accumulation window
        cancels the window's update instead of reaching the moments."""
        from . import ops
        ops.guard_tick(self.state, 0, loss=loss)

    def end_micro_step(self):
        from . import ops
        ops.guard_tick(self.state, 2)

    def step(self, loss=None):
        from . import ops
        self.gsq.zero_()
        ops.sqnorm(self.b.flat_grad, self.gsq)     # always: a non-finite gradient norm also cancels the update
        ops.adamw_step(self.b.flat_param, self.b.flat_grad, self.m, self.v, lr=self.base_lr, betas=self.betas, eps=self.eps,
                       weight_decay=self.wd, step=1, max_norm=self.grad_clip or 0.0, gsq=self.gsq, loss=loss, zero_grad=True,
                       state=self.state, lr_min=self.lr_min, t_max=self.t_max)
        ops.guard_tick(self.state, 1, gsq=self.gsq)


class Trainer:
    """The training micro-step of finetune.py:272-303 on the B200 path, data-parallel when torch.distributed is up."""

    def __init__(self, model, temperature=0.07, lr=1e-4, betas=(0.9, 0.95), weight_decay=0.01, grad_clip=1.0,
                 accumulation_steps=1, total_updates=1000, lr_min=1e-8, fused_optimizer=True):
        from .losses import InfoNCELoss
        self.model = model
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.criterion = InfoNCELoss(temperature, gather_distributed=self.distributed)
        self.params = [p for p in model.parameters() if p.requires_grad]
        names = {id(p): n for n, p in model.named_parameters()}

        def bucket_of(p):  # one bucket per transformer layer
            parts = names[id(p)].split(".")
            for i, s in enumerate(parts):
                if s in ("blocks", "resblocks", "layer") and i + 1 < len(parts) and parts[i + 1].isdigit():
                    return int(parts[i + 1])
            return -1

        on_gpu = bool(self.params) and self.params[0].is_cuda
        self.fused = fused_optimizer and on_gpu
        if on_gpu:
            from .adapters.mona import BaselineMona
            for mod in model.modules():  # let Mona adapters write their gradients straight into the flat buffer
                if isinstance(mod, BaselineMona) and all(p.requires_grad for p in mod.parameters()):
                    for p in mod.parameters():
                        p._ngu_owner = mod
        self.buckets = GradBuckets(self.params, bucket_of, flatten_params=self.fused)
        # One-launch refresh of the bf16 shadows (and transposes) of every Mona projection weight after each update,
        # instead of four small cast launches per layer per step.
        self.castplan = None
        self.monaplan = None
        dt = next((getattr(m, "compute_dtype") for m in model.modules() if hasattr(m, "compute_dtype")), None)
        if on_gpu and self.fused and dt == torch.bfloat16:
            import os
            from .adapters.mona import BaselineMona
            from . import ops
            monas = [m for m in model.modules() if isinstance(m, BaselineMona) and all(p.requires_grad for p in m.parameters())]
            fusable = [m for m in monas if m.project1.weight.shape[0] == 64 and m.project1.weight.shape[1] % 128 == 0
                       and not hasattr(m.adapter_conv, "noise_estimator")]
            if fusable and len(fusable) == len(monas) and os.environ.get("NGU_MONA_FUSED", "1") != "0":
                # fused Mona path: one launch re-derives [W1*ln_w*gamma ; W1*gammax], the merged stencil, ... of every adapter
                self.monaplan = ops.MonaPrepPlan(fusable)
                monas = []
            entries = [(w, tr) for m in monas for w in (m.project1.weight, m.project2.weight) for tr in (False, True)]
            if entries:
                self.castplan = ops.CastPlan(entries, dt)
                self.castplan.fresh = False
                for m in monas:
                    m._ngu_castplan = self.castplan
        self._mona_stale = True
        # overlap the frozen text tower with the vision forward (NGU_TEXT_STREAM=0 disables)
        import os as _os
        text_frozen = hasattr(model, "text") and not any(p.requires_grad for p in model.text.parameters())
        self._text_stream = (torch.cuda.Stream() if (on_gpu and text_frozen and _os.environ.get("NGU_TEXT_STREAM", "1") != "0") else None)
        self._poisoned = False
        self.graph = None
        self.grad_clip = grad_clip
        self.accum = accumulation_steps
        self.micro = 0
        if self.fused:
            self.optimizer = FusedAdamW(self.buckets, lr, betas, 1e-8, weight_decay, grad_clip, total_updates, lr_min)
            self.scheduler = None
        else:
            self.optimizer = torch.optim.AdamW(self.params, lr=lr, betas=betas, weight_decay=weight_decay)
            self.scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(self.optimizer, T_max=total_updates, eta_min=lr_min)

    def micro_step(self, images, ids):
        """One micro-batch: encode, loss, backward (+ optimizer update at accumulation boundaries).
        Returns the (device) loss tensor; nothing here synchronises with the host."""
        m = self.model
        last = (self.micro + 1) % self.accum == 0
        from . import _segcap
        seg = _segcap.ACTIVE is not None and self.distributed   # segmented graph capture: no NCCL call inside a segment
        self.buckets.enabled = last and not seg  # all-reduce only on the micro-step that completes an update
        if self.castplan is not None and not self.castplan.fresh:
            if self.castplan.valid():
                self.castplan.run()
        if self.monaplan is not None and self._mona_stale and self.monaplan.valid():
            self.monaplan.run()
            self._mona_stale = False
        if self._text_stream is not None:
            # the frozen text tower has no dependency on the vision tower: run it on a second stream so its kernels fill the
            # tails / launch gaps of the vision forward (one fork-join; a parallel branch of the graph under capture)
            cur = torch.cuda.current_stream()
            self._text_stream.wait_stream(cur)
            with torch.cuda.stream(self._text_stream), _nvtx("encode_text"):
                ft = m.encode_text(ids)
            with _nvtx("encode_image"):
                fi = m.encode_image(images)
            cur.wait_stream(self._text_stream)
            ft.record_stream(cur)
        else:
            with _nvtx("encode_image"):
                fi = m.encode_image(images)
            with _nvtx("encode_text"):
                ft = m.encode_text(ids)
        with _nvtx("infonce"):
            loss = self.criterion(fi, ft)
        lossd = loss.detach().float().view(1)
        if self.fused:
            self.optimizer.note_loss(lossd)
        elif not bool(torch.isfinite(lossd).all()):
            # torch.optim path (CPU / fused_optimizer=False): the reference's host-side check, finetune.py:281-285
            self.micro += 1
            self._poisoned = True
            if last:
                self._finish_unfused_window()
            return loss.detach()
        with _nvtx("backward"):
            (loss / self.accum).backward()
        self.micro += 1
        if last:
            if seg:
                # one SUM all-reduce of the whole flat gradient buffer between the backward segment and the optimiser segment
                # (the per-layer overlap of the eager path is traded for graph launches: 5.4 MB, tens of microseconds)
                flat, grp = self.buckets.flat_grad, self.buckets.group
                _segcap.collective(lambda: dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=grp))
            else:
                self.buckets.wait()
            if self.fused:
                # global-norm clip + AdamW + cosine LR + zero_grad on device; a non-finite loss anywhere in the accumulation
                # window or a non-finite gradient norm skips the update (finetune.py:281-285 without the host sync)
                with _nvtx("optimizer"):
                    self.optimizer.step(loss=lossd)
                from . import ops as _ops
                _ops.bump_param_epoch()           # raw-pointer parameter update: invalidate low-precision shadows
                if self.castplan is not None:
                    self.castplan.fresh = False   # parameters changed: shadows are re-made at the next micro-step
                self._mona_stale = True
            else:
                self._finish_unfused_window()
        elif self.fused:
            self.optimizer.end_micro_step()
        return loss.detach()

    def _finish_unfused_window(self):
        if not self._poisoned:
            if self.grad_clip and self.grad_clip > 0:
                torch.nn.utils.clip_grad_norm_(self.params, max_norm=self.grad_clip)
            self.optimizer.step()
            self.scheduler.step()
        self._poisoned = False
        self.buckets.zero()

    def flush(self):
        """End of an epoch with leftover micro-batches (finetune.py:287 `or batch_idx + 1 == len(trainloader)`): apply the
        update for the partial accumulation window.  Gradients of the partial window are averaged over `accum` like the
        reference (it divides every micro-loss by accumulation_steps regardless)."""
        if self.micro % self.accum == 0:
            return
        self.micro += self.accum - self.micro % self.accum
        if self.distributed:
            for key in self.buckets.buckets:
                self.buckets.enabled = True
                self.buckets._launch(key)
        self.buckets.wait()
        if self.fused:
            self.optimizer.step()
            from . import ops as _ops
            _ops.bump_param_epoch()
            if self.castplan is not None:
                self.castplan.fresh = False
            self._mona_stale = True
        else:
            self._finish_unfused_window()

    # ---------------------------------------------------------------------------------------------------------
    # CUDA-graph replay of the whole micro-step: forward, backward, collectives, optimiser
    def capture(self, images, ids, warmup=3):
        """Capture one full training step (accumulation_steps == 1) into a CUDA graph with static input buffers.
        The LR schedule, bias-correction step, non-finite guard and dropout counter all live on the device
        (FusedAdamW.state), so replays need nothing from the host.  Parameters / optimiser state are restored after the
        warm-up iterations capture needs: capturing has no training side effects."""
        if not (self.fused and self.accum == 1):
            raise RuntimeError("Trainer.capture needs the fused optimiser and accumulation_steps == 1")
        from . import ops as _ops
        opt = self.optimizer
        _ops.set_seed_counter(opt.state[3:4])
        self.static_in = (torch.empty_like(images), torch.empty_like(ids))
        self.static_in[0].copy_(images)
        self.static_in[1].copy_(ids)
        snap = (self.buckets.flat_param.clone(), opt.m.clone(), opt.v.clone(), opt.state.clone())
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.micro_step(*self.static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():
            self.buckets.flat_param.copy_(snap[0]); opt.m.copy_(snap[1]); opt.v.copy_(snap[2]); opt.state.copy_(snap[3])
            self.buckets.flat_grad.zero_()
        _ops.bump_param_epoch()                 # every derived / low-precision shadow is re-made INSIDE the graph
        if self.castplan is not None:
            self.castplan.fresh = False
        self._mona_stale = True
        if self.distributed:
            # NCCL calls cannot be captured on this stack (the capture hangs): graph segments with the two collectives between
            # them (_segcap.py).  The process-group watchdog thread polls CUDA events meanwhile: thread-local capture mode.
            from . import _segcap
            import gc
            torch.cuda.synchronize()
            gc.collect()
            cap = torch.cuda.Stream()
            cap.wait_stream(torch.cuda.current_stream())
            self.graph = _segcap.SegmentedCapture("thread_local")
            with torch.cuda.stream(cap):
                with self.graph:
                    self.static_loss = self.micro_step(*self.static_in)
            torch.cuda.current_stream().wait_stream(cap)
            torch.cuda.synchronize()
            return self
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="global"):
            self.static_loss = self.micro_step(*self.static_in)
        return self

    def replay(self, images, ids):
        """One captured step on a new batch (device tensors of the captured shapes); returns the (static) loss tensor."""
        if images is not self.static_in[0]:
            self.static_in[0].copy_(images, non_blocking=True)
        if ids is not self.static_in[1]:
            self.static_in[1].copy_(ids, non_blocking=True)
        self.graph.replay()
        return self.static_loss


class DeviceFeeder:
    """Double-buffered host -> device input pipeline.

    The reference's loop does `images.to(device)` at the top of every iteration (finetune.py:272-276) behind a
    `DataLoader(pin_memory=True)`.  Here the copy of batch i+1 runs on a side stream into a second set of device
    buffers while batch i computes, so the PCIe transfer (154 MB of fp32 pixels per 256-image batch) is off the critical
    path.  Every batch is still copied exactly once; `h2d_bytes` counts what was moved.  Iterating yields tuples of
    device tensors that stay valid until the next-but-one `next()`."""

    def __init__(self, batches, device, depth=2):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.depth = depth
        self.slots = [None] * depth
        self.ready = [None] * depth
        self.free = [None] * depth
        self.h2d_bytes = 0
        self.head = 0          # next slot to fill
        self.tail = 0          # next slot to hand out
        self.inflight = 0
        self.stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        self.last = None
        self.exhausted = False

    def _fill(self):
        if self.exhausted or self.inflight >= self.depth:
            return
        try:
            host = next(self.it)
        except StopIteration:
            self.exhausted = True
            return
        k = self.head
        if not self.cuda:
            self.slots[k] = tuple(t.to(self.device) for t in host)
        else:
            if self.slots[k] is None or any(s.shape != h.shape or s.dtype != h.dtype for s, h in zip(self.slots[k], host)):
                self.slots[k] = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in host)
            with torch.cuda.stream(self.stream):
                if self.free[k] is not None:
                    self.stream.wait_event(self.free[k])   # the consumer's kernels that read this slot have finished
                for s, h in zip(self.slots[k], host):
                    s.copy_(h, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
                self.ready[k] = ev
        self.h2d_bytes += sum(h.numel() * h.element_size() for h in host)
        self.head = (k + 1) % self.depth
        self.inflight += 1

    def __iter__(self):
        return self

    def __next__(self):
        if self.cuda and self.last is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))  # everything enqueued so far used the previous slot
            self.free[self.last] = ev
            self.inflight -= 1
            self.last = None
        elif not self.cuda and self.last is not None:
            self.inflight -= 1
            self.last = None
        while self.inflight < self.depth and not self.exhausted:
            self._fill()
        if self.inflight == 0:
            raise StopIteration
        k = self.tail
        if self.cuda:
            torch.cuda.current_stream(self.device).wait_event(self.ready[k])
        self.tail = (k + 1) % self.depth
        self.last = k
        return self.slots[k]


class ScalarLog:
    """Asynchronous device -> host read of per-step scalars (the `loss.item()` of finetune.py:299) that does not stall
    the launch queue: `push` enqueues a copy into pinned memory, `pop_ready` returns the values whose copies have landed
    (normally everything but the step in flight), `drain` waits for the rest."""

    def __init__(self, capacity=8):
        self.ring = None
        self.cap = capacity
        self.pending = []      # (slot, event)
        self.next = 0
        self.d2h_bytes = 0

    def push(self, scalar):
        if not scalar.is_cuda:
            self.pending.append((float(scalar), None))
            return
        if self.ring is None:
            self.ring = torch.empty(self.cap, dtype=torch.float32).pin_memory()
        if len(self.pending) >= self.cap:
            raise RuntimeError("ScalarLog overflow: pop_ready()/drain() must be called at least every `capacity` pushes")
        k = self.next
        self.ring[k:k + 1].copy_(scalar.detach().float().view(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.pending.append((k, ev))
        self.next = (k + 1) % self.cap
        self.d2h_bytes += 4

    def pop_ready(self, keep_inflight=1):
        out = []
        while len(self.pending) > keep_inflight:
            k, ev = self.pending.pop(0)
            if ev is None:
                out.append(k)
            else:
                ev.synchronize()
                out.append(float(self.ring[k]))
        return out

    def drain(self):
        return self.pop_ready(0)
