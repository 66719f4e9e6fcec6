"""GAN training step: mirror of DeepSC-GAN/utlis/gan_train.py (gan_train_step :8-50, eval_step :53-83).

One forward of ``Transeiver_GAN`` under the tape, then the reference's three optimizer applications:
  A (:24-29)  loss   = CE(clean branch)                          on every variable except the generator ``g``;
  B (:35-37)  g_loss = 10 - CE(perturbed branch)                  on the generator (trainable_variables[104:108]);
  C (:39-44)  d_loss = l*CE(clean) + (1-l)*CE(perturbed)          on the receiver (channel decoder + semantic decoder;
              'g', 'encoder', 'channel_encoder' frozen - intent per SURVEY.md App. B Q14).  The reference freezes by
              ``layer.name`` (:41), and Keras auto-names a ``Channel_Encoder`` instance 'channel__encoder', so as written
              its channel encoder is NOT frozen in step C; ``step_c_literal_names=True`` reproduces that (step C then also
              updates the channel encoder with d_loss).  The default follows the stated intent.
All three gradients are taken at the pre-update weights (one persistent tape), and all three losses are linear in
CE_r and CE_p, so two backward sweeps fill two flat gradient buffers (d CE_r, d CE_p); with a process group the two
buffers are all-reduced together in ONE NCCL collective, and the lambda mix of step C is folded into the Adam kernel.
"""
from __future__ import annotations

import torch

from .. import _lib
from .. import optim as O
from ..models.modules import create_masks, differentiable, loss_function
from .eval import fgm_perturbation
from .trainer import make_optimizer  # noqa: F401  (re-export: the optimizer the missing driver builds)


def _is_generator(name: str) -> bool:
    return name.startswith("generator.")


def _is_receiver(name: str) -> bool:
    return name.startswith("channel_decoder.") or name.startswith("semantic_decoder.")


def _is_channel_encoder(name: str) -> bool:
    return name.startswith("channel_encoder.")


def gan_train_step(inp, tar, p, net, optim_net, lenmda, channel='AWGN', n_std=0.1, training=False, traingan=False, *,
                   noise=None, noise_r=None, h=None, h_r=None, p_draw=None, step_c_literal_names: bool = False,
                   _phase: str = "all", _state=None):
    """utlis/gan_train.py:8-50.  ``p`` is overwritten as in the reference (:13-14): a normal draw of std n_std
    normalised to unit Frobenius norm (``p_draw`` = injected unit-normal tensor; used only when traingan=False).
    The forward always runs with training=True and PNR_dB = 40 (:16-18).  Returns (loss, g_loss, d_loss).

    ``_phase`` (used by GraphedGanTrainStep under a process group): "grads" stops after the two backward sweeps and
    returns (ce_r, ce_p); "apply" takes that pair as ``_state`` and runs the three Adam applications with the 1/world
    factor, the caller having all-reduced ``optim_net.fp.grad_bucket`` in between."""
    fp = optim_net.fp
    step_c = (lambda n: _is_receiver(n) or _is_channel_encoder(n)) if step_c_literal_names else _is_receiver
    if _phase == "apply":
        ce_r, ce_p = _state
        return _apply_three(optim_net, fp, step_c, ce_r, ce_p, lenmda, O.mean_scale())
    tar_inp, tar_real = tar[:, :-1], tar[:, 1:]
    masks = create_masks(inp, tar_inp)
    enc_padding_mask, combined_mask, dec_padding_mask = masks
    dev = inp.device
    if p_draw is None:
        p_draw = torch.randn((inp.shape[0], inp.shape[1], 16), device=dev, dtype=torch.float32)
    p = _lib.power_normalize(p_draw.contiguous(), 1, factor=float(p_draw.numel()))   # p / ||p||_F (n_std cancels)
    assert fp.grad_bucket.shape[0] >= 2, "gan_train_step needs two flat gradient buffers (make_optimizer default)"
    fp.grad_bucket.zero_()
    not_g = fp.select(lambda n: not _is_generator(n))
    with differentiable():
        predictions_p, predictions_r, _, _ = net(inp, tar_inp, p, 40, channel=channel, n_std=n_std, training=True,
                                                 enc_padding_mask=enc_padding_mask, combined_mask=combined_mask,
                                                 dec_padding_mask=dec_padding_mask, traingan=traingan, noise=noise,
                                                 noise_r=noise_r, h=h, h_r=h_r)
        ce_r = loss_function(tar_real, predictions_r)
        ce_p = loss_function(tar_real, predictions_p)
        fp.point_grads(0)
        ce_r.backward(inputs=not_g, retain_graph=True)                       # d CE_r / d (everything but g)
        fp.point_grads(1)
        wanted = fp.select(lambda n: step_c(n) or (traingan and _is_generator(n)))
        ce_p.backward(inputs=wanted)                                         # d CE_p / d (step-C variables, generator)
    if _phase == "grads":
        return ce_r.detach(), ce_p.detach()
    scale = O.all_reduce_mean_scale(fp.grad_bucket)
    return _apply_three(optim_net, fp, step_c, ce_r.detach(), ce_p.detach(), lenmda, scale)


def _apply_three(optim_net, fp, step_c, ce_r, ce_p, lenmda, scale):
    """The reference's three apply_gradients calls (utlis/gan_train.py:24-44) on the (summed) flat gradient buffers."""
    g_r, g_p = fp.grad_bucket[0], fp.grad_bucket[1]
    optim_net.apply(fp.ranges(lambda n: not _is_generator(n)), g_r, scale)                      # step A
    optim_net.apply(fp.ranges(_is_generator), g_p, -scale)                                      # step B: 10 - CE_p
    optim_net.apply(fp.ranges(step_c), g_r, scale * float(lenmda), g_p, scale * (1.0 - float(lenmda)))          # step C
    return ce_r, 10 - ce_p, float(lenmda) * ce_r + (1 - float(lenmda)) * ce_p


def eval_step(inp, tar, net, channel='AWGN', n_std=0.1, epsilon=1, *, noise=None, noise2=None, noise2_r=None, h=None):
    """utlis/gan_train.py:53-83 (stale in the reference: it unpacks 3 of 4 outputs and omits PNR_dB; intent kept):
    FGM direction from the clean branch w.r.t. y_r, second forward with it, loss on the perturbed branch.
    Returns (loss, loss_p)."""
    from .eval import eval_step_FGM
    loss, loss_p, _, _ = eval_step_FGM(inp, tar, net, 0, channel=channel, n_std=n_std, epsilon=epsilon, noise=noise,
                                       noise2=noise2, noise2_r=noise2_r, h=h)
    return loss, loss_p


class GraphedGanTrainStep:
    """``gan_train_step`` captured once as a CUDA graph and replayed: one launch per training step instead of ~1,050.

    The eager step is launch-bound (its kernels are 3-30 us, shorter than the host path through autograd and ctypes:
    ~22 ms per step for ~10.7 ms of kernel time).  Everything that changes from step to step lives on the device: the
    channel noise and the unit-norm perturbation draw come from ``torch.randn`` inside the graph (graph-safe Philox),
    the dropout masks and Adam's bias correction read a device-side step counter (``_lib.STEP_DEV``), the batch is
    copied into a static buffer.  With a process group the step is two graphs with the flat-gradient all-reduce between
    them (one NCCL collective per step, issued on the stream, never captured).

        step = GraphedGanTrainStep(net, optim_net, lenmda=0.5, n_std=SNR_to_noise(3), traingan=True)
        loss, g_loss, d_loss = step(inp, tar)        # device scalars, overwritten by the next call
    """

    def __init__(self, net, optim_net, lenmda, channel='AWGN', n_std=0.1, traingan=True, warmup: int = 2):
        if channel != 'AWGN':
            # Channels.fading draws its coefficient on the host (models/transceiver.py:48-50 is one scalar per call): a
            # capture would bake ONE coefficient into every replay.  Use the eager gan_train_step for fading channels.
            raise ValueError("GraphedGanTrainStep supports channel='AWGN' only; call gan_train_step for fading channels")
        self.net, self.opt, self.lenmda, self.channel, self.n_std, self.traingan = net, optim_net, lenmda, channel, n_std, traingan
        self.warmup, self.graph, self.out = warmup, None, None
        dev = optim_net.fp.flat.device
        self.step_dev = torch.zeros((), device=dev, dtype=torch.int64)
        self.steps = 0

    def _one(self, phase: str = "all", state=None):
        kw = dict(channel=self.channel, n_std=self.n_std, training=True, traingan=self.traingan)
        if phase == "apply":
            return gan_train_step(None, None, None, self.net, self.opt, self.lenmda, _phase="apply", _state=state, **kw)
        shape = (self._inp.shape[0], self._inp.shape[1], 16)
        dev = self._inp.device
        z = torch.randn(shape, device=dev)
        z_r = torch.randn(shape, device=dev)
        p_draw = torch.randn(shape, device=dev)
        return gan_train_step(self._inp, self._tar, None, self.net, self.opt, self.lenmda, noise=z, noise_r=z_r,
                              p_draw=p_draw, _phase=phase, **kw)

    @staticmethod
    def _world() -> int:
        import torch.distributed as dist
        return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1

    def __call__(self, inp, tar=None):
        prev = (_lib.STEP_DEV, _lib.ADAM_APPLIES_PER_STEP)
        _lib.STEP_DEV, _lib.ADAM_APPLIES_PER_STEP = self.step_dev, 3
        try:
            if self.graph is None:
                self._inp = inp.clone()
                self._tar = self._inp if (tar is None or tar is inp) else tar.clone()   # the dataset yields (x, x)
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(self.warmup):                   # eager: fills caches, warms the allocator
                        self._one()
                        self.steps += 1
                torch.cuda.current_stream().wait_stream(side)
                self.graph = torch.cuda.CUDAGraph()
                if self._world() == 1:
                    self.graph_apply = None
                    with torch.cuda.graph(self.graph):
                        self.out = self._one()
                        self.step_dev += 1
                else:
                    # data-parallel: [forward + two backward sweeps] and [three Adam applications] are captured as two
                    # graphs and the flat-bucket all-reduce runs between them on the stream.  No NCCL kernel lives in a
                    # captured graph, so tearing the process group down never waits on a graph (the 8-GPU hang of round 1).
                    with torch.cuda.graph(self.graph):
                        self._ce = self._one("grads")
                    self.graph_apply = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph_apply):
                        self.out = self._one("apply", self._ce)
                        self.step_dev += 1
                # the capture advanced the host-side iteration count by one step without running it; the first replay
                # below is that step (the device counter is still 0)
            else:
                self.opt.iterations += 3
            self._inp.copy_(inp)
            if self._tar is not self._inp:
                self._tar.copy_(inp if tar is None else tar)
            elif tar is not None and tar is not inp:
                raise ValueError("this step was captured with tar = inp; build another GraphedGanTrainStep for a separate target")
            self.graph.replay()
            if self.graph_apply is not None:
                import torch.distributed as dist
                dist.all_reduce(self.opt.fp.grad_bucket, op=dist.ReduceOp.SUM)
                self.graph_apply.replay()
            self.steps += 1
            _lib.weights_changed()                                 # packed / padded weight caches of eager callers
        finally:
            _lib.STEP_DEV, _lib.ADAM_APPLIES_PER_STEP = prev
        return self.out
