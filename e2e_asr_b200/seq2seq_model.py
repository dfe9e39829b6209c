"""Seq2SeqModel: the multitask attention encoder-decoder training step.

Keeps the reference's constructor, hyper-parameters and public attributes
(seq2seq_model.py:26-216): `Seq2SeqModel(data_iter, isTraining, params)` builds the
model and runs `create_computational_graph()` on the next batch of `data_iter`
(TF built a graph once and re-ran it; here every call to
`create_computational_graph()` / `run_step()` executes one step eagerly on the
GPU), leaving `encoder_inputs, decoder_inputs, seq_len, seq_len_target, targets,
encoder_hidden_states, time_major_states, seq_len_encs, outputs, losses,
total_loss, updates` behind.

Additions the north-star asks for (absent from the reference):
auxiliary CTC heads on lower encoder layers (`params.ctc_tasks`), and data-parallel
gradient averaging over NCCL (`dist.GradAllReducer`).  The Adam update
(seq2seq_model.py:137,153-155) is not built yet: `updates` holds the clipped
gradients.
"""
import numpy as np
import torch

from . import ops
from ._lib import call
from .attn_decoder import AttnDecoder
from .base_params import BaseParams, Bunch
from .encoder import Encoder
from .losses import LossUtils
from .tf_utils import create_shifted_targets
from .variables import VariableStore


class Seq2SeqModel(BaseParams):
    """Implements the Attention-Enabled Encoder-Decoder model."""

    @classmethod
    def class_params(cls):
        params = Bunch()
        # Task specification (seq2seq_model.py:31-35)
        params['tasks'] = ['char']
        params['num_layers'] = {'char': 4}
        params['max_output'] = {'char': 120}
        # Optimization params
        params['learning_rate'] = 1e-3
        params['learning_rate_decay_factor'] = 0.5
        params['max_gradient_norm'] = 5.0
        # Loss params
        params['avg'] = True
        params['encoder_params'] = Encoder.class_params()
        params['decoder_params'] = {'char': AttnDecoder.class_params()}
        # --- additions (not in the reference) ---
        # auxiliary CTC heads: task -> vocabulary size (blank is added as the last class);
        # the encoder depth comes from num_layers[task]; labels from batch[task], batch[task+"_len"]
        params['ctc_tasks'] = {}
        # tf.clip_by_global_norm sees the embedding gradient as IndexedSlices
        params['tf_indexed_slices_norm'] = True
        # opt.apply_gradients (seq2seq_model.py:137,153-155) inside run_step; off by default so that the
        # parity tests and the fwd+bwd benchmark see the clipped gradients of fixed parameters
        params['apply_updates'] = False
        # seed of the (builder-defined) Philox dropout masks; the step counter is mixed in
        params['dropout_seed'] = 0
        return params

    def __init__(self, data_iter, isTraining=True, params=None, variables=None, device="cuda", reducer=None):
        self.params = self.class_params() if params is None else params
        params = self.params
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            # an indexed device everywhere: per-device registries (side streams, step-start events) are keyed by
            # str(tensor.device) = "cuda:N"
            self.device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self.variables = variables if variables is not None else VariableStore(self.device)
        self.encoder = Encoder(isTraining=isTraining, params=params.encoder_params, variables=self.variables)
        self.decoder = {}
        for task in params.tasks:
            self.decoder[task] = AttnDecoder(isTraining=isTraining, params=params.decoder_params[task],
                                             scope=task, variables=self.variables)
        self.data_iter = data_iter
        self.isTraining = isTraining
        self.reducer = reducer
        self.learning_rate = float(params.learning_rate)
        self.global_step = 0
        self.epoch = 0
        self._pinned = {}
        self._sq = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._sq_emb = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.ctc_stash = {}
        self._side_stream = None
        self._loss_scale = None
        self._loss_scale_value = None
        self._inflight = []
        self._adam = None
        self._seed_dev = None
        # loop bounds of the recurrences / decoder loop: False = the batch's maximum lengths (the reference's
        # dynamic_rnn / raw_rnn semantics), True = the padded shapes with masked extra steps (one captured step per
        # length bucket, GraphedStep(bucket=True))
        self.shape_bounds = False
        if data_iter is not None:
            self.create_computational_graph()

    # learning-rate / epoch bookkeeping ops of seq2seq_model.py:74-83
    def learning_rate_decay_op(self):
        self.learning_rate *= self.params.learning_rate_decay_factor
        return self.learning_rate

    def epoch_incr(self):
        self.epoch += 1
        return self.epoch

    # ------------------------------------------------------------------
    def _to_device(self, key, arr, dtype):
        """Host array -> device through reusable pinned staging buffers.  The H2D copy is asynchronous, so a staging
        buffer must not be rewritten before the copy that reads it has finished: every key owns TWO pinned buffers
        used alternately, each with the event of its last copy, which the host waits on before refilling it (with a
        run-ahead of one step that wait is already satisfied)."""
        buf = None
        if isinstance(arr, torch.Tensor):
            if arr.device == self.device:
                return arr.to(dtype)
            if arr.dtype == dtype and arr.is_contiguous() and arr.is_pinned():
                buf = arr               # a loader that hands out pinned batches: copy straight from its buffer
            else:
                arr = arr.numpy()
        slot = None
        if buf is None:
            arr = np.ascontiguousarray(arr)
            t = torch.from_numpy(arr)
            if t.dtype != dtype:
                t = t.to(dtype)
            ring = self._pinned.get(key)
            if ring is None or ring["bufs"][0].shape != t.shape or ring["bufs"][0].dtype != dtype:
                ring = self._pinned[key] = {"bufs": [torch.empty(t.shape, dtype=dtype).pin_memory() for _ in range(2)],
                                            "events": [None, None], "next": 0}
            slot = ring["next"]
            ring["next"] = 1 - slot
            if ring["events"][slot] is not None:
                ring["events"][slot].synchronize()
            buf = ring["bufs"][slot]
            buf.copy_(t)
        static = getattr(self, "_static_inputs", None)
        if static is not None:
            # graphed step: the captured kernels read these very buffers, so refill them in place
            dev = static.get(key)
            if dev is None or dev.shape != buf.shape or dev.dtype != dtype:
                dev = static[key] = torch.empty(buf.shape, dtype=dtype, device=self.device)
            dev.copy_(buf, non_blocking=True)
        else:
            dev = buf.to(self.device, non_blocking=True)
        if slot is not None:
            ev = torch.cuda.Event()
            ev.record()
            ring["events"][slot] = ev
        return dev

    def _len_tensor(self, key, arr):
        host = np.asarray(arr.cpu().numpy() if isinstance(arr, torch.Tensor) else arr, np.int64)
        t = self._to_device(key, host, torch.int64)
        t._host = host
        t._i32 = self._to_device(key + "/i32", host.astype(np.int32), torch.int32)
        return t

    def get_batch(self, batch):
        """Get a batch from the iterator (seq2seq_model.py:159-197).  Frame stacking
        (:164-183) is fused into the encoder's input staging kernel."""
        encoder_inputs = self._to_device("logmel", batch["logmel"], torch.float32)
        encoder_len = self._len_tensor("logmel_len", batch["logmel_len"])
        decoder_inputs, decoder_len = {}, {}
        for task in self.params.tasks:
            ids = batch[task]
            ids = ids.cpu().numpy() if isinstance(ids, torch.Tensor) else np.asarray(ids)
            decoder_inputs[task] = self._to_device(task, np.ascontiguousarray(ids.T), torch.int64)   # time major (:189)
            ln = batch[task + "_len"]
            if not self.isTraining:                                                                  # (:191-193)
                ln = np.ones_like(np.asarray(ln)) * self.params.max_output[task]
            decoder_len[task] = self._len_tensor(task + "_len", ln)
        for task in self.params.ctc_tasks:
            decoder_inputs[task] = self._to_device(task, batch[task], torch.int64)                   # [B, Lmax] labels
            decoder_len[task] = self._len_tensor(task + "_len", batch[task + "_len"])
        if not self.isTraining and "utt_id" in batch:
            decoder_inputs["utt_id"] = batch["utt_id"]
        return [encoder_inputs, decoder_inputs, encoder_len, decoder_len]

    # ------------------------------------------------------------------
    def create_computational_graph(self, batch=None, prepared=None):
        """One step (seq2seq_model.py:88-157).  `prepared` = a previous get_batch()
        result (device-resident inputs), otherwise `batch` (or the next batch of
        data_iter) is staged host -> device first."""
        params = self.params
        if prepared is None:
            if batch is None:
                batch = self.data_iter.get_next()
            prepared = self.get_batch(batch)
        self.encoder_inputs, self.decoder_inputs, self.seq_len, self.seq_len_target = prepared
        self._stage_step_seed()

        if self.isTraining:
            # Bound the host's run-ahead to ONE step (step N+1 is enqueued while step N runs): tensors handed to
            # the side streams return to the caching allocator through record_stream events; with an unbounded
            # backlog every allocation polls a growing event list and blocks that are still pending force fresh
            # cudaMalloc calls (measured: 34-66 ms per 17 ms GPU step).
            if torch.cuda.is_available():
                while len(self._inflight) >= 2:
                    self._wait(self._inflight.pop(0))
        self._step_core()
        if self.isTraining:
            self._step_tail()

    def _step_seed_value(self):
        """One Philox key per step: seed * 1000003 + global_step (see oracle train_step(dropout_seed=...))."""
        # (data parallel: every rank draws its own masks)
        rank = int(getattr(self.reducer, "rank", 0)) if self.reducer is not None else 0
        return (int(self.params.get('dropout_seed', 0)) + 7919 * rank) * 1000003 + int(self.global_step)

    def _stage_step_seed(self):
        """Writes this step's Philox key into the device word the dropout kernels read (host -> device through the
        pinned staging ring, asynchronous).  Done OUTSIDE a captured step, before it is launched: the captured kernels
        hold the word's address, so every replay draws fresh masks."""
        if self.device.type != "cuda" or not self.isTraining:
            return
        static, self._static_inputs = getattr(self, "_static_inputs", None), None
        try:
            fresh = self._to_device("step_seed", np.array([self._step_seed_value()], np.int64), torch.int64)
        finally:
            self._static_inputs = static
        if self._seed_dev is None:
            self._seed_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._seed_dev.copy_(fresh, non_blocking=True)

    def _step_core(self):
        """Forward, losses and backward of one step on the inputs held in self.encoder_inputs / decoder_inputs:
        everything up to the complete local gradient in the flat buffer.  No host synchronisation and no per-step
        host state inside, so it can be captured in a CUDA graph (GraphedStep)."""
        params = self.params
        self.targets, self.target_weights = {}, {}
        for task in params.tasks:
            self.targets[task], self.target_weights[task] = create_shifted_targets(
                self.decoder_inputs[task], self.seq_len_target[task]) if self.isTraining else (None, None)
        from ._lib import capture_checkpoint as ck
        ck("step start")
        if self.isTraining:
            self.variables.zero_grad()
        ck("zero_grad")
        if self.isTraining and getattr(params, "overlap_weight_grads", True):
            ops.enable_wgrad_stream(self.device)
            ops.mark_step_start(self.device)
        # the key is passed by value AND as the address of the device word holding it (ops.StepSeed)
        step_seed = ops.StepSeed(self._step_seed_value(), self._seed_dev)
        self.encoder.dropout_seed = step_seed
        self.encoder.shape_bounds = self.shape_bounds
        for i, task in enumerate(params.tasks):
            self.decoder[task].dropout_seed = step_seed
            self.decoder[task].dropout_stream = i
            self.decoder[task].shape_bounds = self.shape_bounds
        depth_of = dict((t, params.num_layers[t]) for t in list(params.tasks) + list(params.ctc_tasks))
        ctx = torch.enable_grad() if self.isTraining else torch.no_grad()
        with ctx:
            self.encoder_hidden_states, self.time_major_states, self.seq_len_encs = \
                self.encoder(self.encoder_inputs, self.seq_len, depth_of)
            ck("encoder forward")

            # The auxiliary CTC heads only need the encoder states: they run on a side stream,
            # concurrently with the (latency-bound) attention decoders; autograd replays their
            # backward on the same side stream.
            self.losses = {}
            main = torch.cuda.current_stream()
            sides = []
            if self.isTraining and params.ctc_tasks:
                if self._side_stream is None:
                    self._side_stream = {}
                for task, vocab in params.ctc_tasks.items():
                    # one stream per head: the heads are independent of each other as well
                    if task not in self._side_stream:
                        self._side_stream[task] = torch.cuda.Stream(device=self.device)
                        if ops.get_gemm_mode() != 0:
                            ops.ensure_workspace(self.device, nbytes=512 << 20, stream=self._side_stream[task])
                    side = self._side_stream[task]
                    sides.append(side)
                    ev = getattr(self.encoder, "layer_done", {}).get(params.num_layers[task])
                    if ev is not None:
                        side.wait_event(ev)      # the head starts when ITS encoder layer is done
                    else:
                        side.wait_stream(main)
                    with torch.cuda.stream(side):
                        d = params.num_layers[task]
                        # the encoder keeps the time-major view for tasks named "state" / "*_ctc" (encoder.py:143-144);
                        # any other CTC task name finds its layer among the batch-major attention states (the head
                        # takes either layout)
                        states_d = self.time_major_states.get(d)
                        if states_d is None:
                            states_d = self.encoder_hidden_states[d]
                        D = states_d.shape[2]
                        if ops.get_gemm_mode() != 0:
                            # operand pre-pass scratch of the head's largest product (d kernel = states^T . d logits:
                            # two planes of both operands); grows only on the first, un-captured step
                            rows = states_d.shape[0] * states_d.shape[1]
                            need = 2 * 4 * rows * (D + vocab + 9) + (1 << 20)
                            ops.ensure_workspace(self.device, nbytes=max(512 << 20, need), stream=side)
                        k = self.variables.get("model/ctc_%s/kernel" % task, (D, vocab + 1))
                        b = self.variables.get("model/ctc_%s/bias" % task, (vocab + 1,), ("zeros",))
                        self.ctc_stash[task] = {"consumer_stream": main}
                        self.losses[task] = LossUtils.ctc_head_loss(
                            states_d, k, b, self.seq_len_encs[d], self.decoder_inputs[task],
                            self.seq_len_target[task], self.ctc_stash[task],
                            max_label_len=self.decoder_inputs[task].shape[1] if self.shape_bounds else None)

            ck("ctc heads forward")
            self.outputs = {}
            for task in params.tasks:
                d = params.num_layers[task]
                self.outputs[task] = self.decoder[task](
                    self.decoder_inputs[task], self.seq_len_target[task],
                    self.encoder_hidden_states[d], self.seq_len_encs[d])

            ck("decoders forward")
            if not self.isTraining:
                return
            for task in params.tasks:
                self.losses[task] = LossUtils.cross_entropy_loss(
                    self.outputs[task], self.targets[task], self.seq_len_target[task])
            self.losses = {t: self.losses[t] for t in list(params.tasks) + list(params.ctc_tasks)}

        # Gradients, clipping (:148-151).  Adam (:137,153-155) is the "next" row.
        # total_loss = sum_task loss_task (/ n_tasks iff avg) (:140-144).  The sum is differentiated term by
        # term -- d total / d loss_task is the same constant for every task -- so the backward of the attention
        # decoder never waits for the CTC heads' forward on the side streams; autograd replays each head's
        # backward on the head's own stream and joins it where the encoder layer consumes its output.
        if getattr(params, "overlap_weight_grads", True):
            ops.enable_wgrad_stream(self.device)
        scale = 1.0 / float(len(self.losses)) if params.avg else 1.0
        if self._loss_scale is None or float(self._loss_scale_value) != scale:
            self._loss_scale = torch.full((), scale, dtype=torch.float32, device=self.device)
            self._loss_scale_value = scale
        roots = [self.losses[t] for t in self.losses]
        ck("losses")
        dp = self.reducer is not None and self.reducer.world_size > 1
        if dp:
            # data parallel: every group of weight gradients is all-reduced from the side stream that produced it,
            # while the rest of the backward pass runs (dist.GradAllReducer)
            self.reducer.begin_step(self.variables.gflat)
            ops.set_grad_ready_hook(self.reducer.grad_ready)
        try:
            torch.autograd.backward(roots, [self._loss_scale] * len(roots))
        finally:
            ops.set_grad_ready_hook(None)
        ck("backward")
        for side in sides:
            main.wait_stream(side)
        ops.sync_wgrad_stream(self.device)
        ck("joins")
        if dp:
            # what is left (auxiliary heads, encoder layer 1 if its hook did not fire) + one trailing float: the sum of
            # squares of the embedding gradient's IndexedSlices values, which tf.global_norm needs summed over ranks
            vs = self.variables
            self._emb_sq_local()
            vs.gflat[vs.used:vs.used + 1].copy_(self._sq_emb)
            self.reducer.finish(vs.used, extra=1)
        with torch.no_grad():
            self.total_loss = torch.stack([l.detach() for l in roots]).sum() * scale
        ck("total_loss")

    def _step_tail(self):
        """Clipping (with the data-parallel 1/n folded in), optional Adam, step bookkeeping (seq2seq_model.py:148-155).
        The gradient all-reduce itself is part of _step_core: it overlaps the backward pass."""
        params = self.params
        self.clip_gradients()
        self.updates = self.variables
        if params.get('apply_updates', False):
            self.apply_gradients()
        self.global_step += 1
        if torch.cuda.is_available():
            ev = torch.cuda.Event()
            ev.record()
            self._inflight.append(ev)
            if len(self._inflight) >= 2:          # step N-1 must be done before step N+1 is enqueued
                self._wait(self._inflight.pop(0))

    run_step = create_computational_graph

    def graphed_step(self, batch, bucket=False):
        """Captures the step for batches of this batch's shape in a CUDA graph; see GraphedStep.  bucket=True: one
        captured step serves every batch that fits the shapes of `batch` (a length bucket)."""
        return GraphedStep(self, batch, bucket=bucket)

    def _wait(self, event):
        """Blocks the host on a step-completion event; the time spent here is the host's slack (host_wait_s)."""
        import time
        t0 = time.perf_counter()
        event.synchronize()
        self.host_wait_s = getattr(self, "host_wait_s", 0.0) + time.perf_counter() - t0

    def _emb_sq_local(self):
        """Sum of squares of this rank's embedding-gradient IndexedSlices values (one row per looked-up token,
        duplicates not summed: what tf.global_norm sees) -> self._sq_emb; False if there are none."""
        st = ops._dev_state(self.device)
        first = True
        if self.params.tf_indexed_slices_norm:
            for task in self.params.tasks:
                du = self.decoder[task].stash.get("emb_values")
                if du is None:
                    continue
                call("e2e_sumsq", du.numel(), du, st["partials"], self._sq_emb, 1.0, 0 if first else 1)
                first = False
        if first:
            self._sq_emb.zero_()
        return not first

    def clip_gradients(self):
        """tf.clip_by_global_norm(gradients, max_gradient_norm) over the flat buffer.
        With `tf_indexed_slices_norm` the embedding gradient enters the norm as TF's
        IndexedSlices values (one row per looked-up token, duplicates not summed).
        Data parallel: the buffer holds the SUM of the n rank gradients; the norm of their mean is taken with
        sign = 1/n^2 and the 1/n itself is the clipping kernel's pre-scale (no averaging pass)."""
        vs = self.variables
        st = ops._dev_state(self.device)
        g = vs.flat_grads()
        n = self.reducer.world_size if self.reducer is not None else 1
        inv2 = 1.0 / float(n * n)
        call("e2e_sumsq", g.numel(), g, st["partials"], self._sq, inv2, 0)
        if self.params.tf_indexed_slices_norm:
            have = False
            for task in self.params.tasks:
                if self.decoder[task].stash.get("emb_values") is None:
                    continue
                ge = vs.grad(self.decoder[task].scope_name() + "/decoder/embedding")
                call("e2e_sumsq", ge.numel(), ge, st["partials"], self._sq, -inv2, 1)
                have = True
            if have:
                if n > 1:       # summed over ranks with the gradients (the float behind the flat buffer)
                    call("e2e_axpy", 1, inv2, vs.gflat[vs.used:vs.used + 1], self._sq)
                else:
                    self._emb_sq_local()
                    call("e2e_axpy", 1, 1.0, self._sq_emb, self._sq)
        call("e2e_clip_by_norm", g.numel(), g, self._sq, float(self.params.max_gradient_norm), self.grad_norm,
             1.0 / float(n), st["err"])

    def apply_gradients(self, beta1=0.9, beta2=0.999, epsilon=1e-8):
        """tf.train.AdamOptimizer(self.learning_rate).apply_gradients(clipped gradients) (seq2seq_model.py:137,
        153-155) as ONE fused kernel over the flat parameter / gradient / moment buffers.  TF applies the embedding
        gradient as IndexedSlices with duplicate indices summed first, which equals the dense update used here."""
        vs = self.variables
        n = vs.used
        if self._adam is None or self._adam["m"].numel() < n:
            self._adam = dict(m=torch.zeros(vs.capacity, dtype=torch.float32, device=self.device),
                              v=torch.zeros(vs.capacity, dtype=torch.float32, device=self.device), t=0)
        a = self._adam
        a["t"] += 1
        lr_t = self.learning_rate * (1.0 - beta2 ** a["t"]) ** 0.5 / (1.0 - beta1 ** a["t"])
        call("e2e_adam", (n + 3) // 4 * 4, vs.flat, vs.gflat, a["m"], a["v"], float(lr_t), float(beta1), float(beta2),
             float(epsilon))

    def gradients(self):
        """Clipped gradients keyed by TF variable name (host copies)."""
        return {n: self.variables.grad(n).detach().cpu().numpy().copy() for n in self.variables.names()}

    @classmethod
    def add_parse_options(cls, parser):
        # flag names and defaults of seq2seq_model.py:199-216
        parser.add_argument("-tasks", "--tasks", default="", type=str, help="Auxiliary task choices")
        parser.add_argument("-nlc", "--num_layers_char", default=4, type=int,
                            help="Output layer of encoder which is used for char.")
        parser.add_argument("-nlp", "--num_layers_phone", default=3, type=int,
                            help="Output layer of encoder which is used for phone.")
        parser.add_argument("-max_out_char", "--max_output_char", default=120, type=int,
                            help="Maximum length of char/word-piece sequence")
        parser.add_argument("-max_out_phone", "--max_output_phone", default=250, type=int,
                            help="Maximum length of phone sequence")
        parser.add_argument("-lr_decay", "--learning_rate_decay_factor", default=0.5, type=float,
                            help="Learning rate decay factor")
        parser.add_argument("-avg", "--avg", default=False, action="store_true", help="Average the loss")


class GraphedStep(object):
    """One training step replayed from a CUDA graph.

    The eager step issues ~200 kernel launches and ~12 ms of host work on several streams; when the host is slower
    than the 15 ms device step (8 ranks sharing the host's cores) the GPU waits for it.  The step's kernel sequence
    is a pure function of the batch SHAPE (B, padded T, U = max target length, max CTC label length), so it is
    captured once -- forward, losses, backward with its weight-gradient and CTC side streams as parallel graph
    branches -- and replayed with one launch.  Inputs are refilled in place (pinned staging -> the static device
    buffers the captured kernels read); the gradient all-reduce, clipping and Adam stay eager after the replay.

    Valid for batches with the captured shapes and the same maximum lengths.  Output dropout (out_prob, out_prob_dec
    < 1) is captured: the mask kernels read the step's Philox key from a device word the host rewrites before every
    replay.  Scheduled sampling (samp_prob > 0) is not (its ids are realised step by step from the host).
    `step(batch)` = Seq2SeqModel.run_step(batch).
    """

    def __init__(self, model, batch, warmup=2, bucket=False):
        if not model.isTraining:
            raise ValueError("GraphedStep captures the training step")
        self.bucket = bool(bucket)
        p = model.params
        if any(d.samp_prob > 0 for d in p.decoder_params.values()):
            raise NotImplementedError("GraphedStep: scheduled sampling realises its ids step by step from the host; "
                                      "use the eager run_step")
        self.model = model
        model._static_inputs = {}
        self.caps = {k: tuple(np.asarray(v).shape) for k, v in batch.items()
                     if hasattr(v, "shape") and np.asarray(v).ndim >= 2}
        self.prepared = model.get_batch(batch)
        self.signature = self._signature(self.prepared)
        model.shape_bounds = self.bucket or model.shape_bounds
        cur = torch.cuda.current_stream()
        # eager warm-up ON THE CAPTURE STREAM: lazy variables, workspaces, side streams and the allocator reach their
        # steady state, and autograd's cached gradient accumulators are bound to the stream that will be captured
        # (an accumulator created on the default stream would run -- uncaptured -- on the default stream)
        # high priority: the captured kernel nodes of the critical path (recurrences, decoder loop) inherit it and
        # are scheduled ahead of the weight-gradient / CTC branches that only have to finish by the end of the step
        self.stream = torch.cuda.Stream(device=model.device, priority=-1)
        self.stream.wait_stream(cur)
        # (the warm-up steps must leave the training state alone: no optimiser update, same step counter)
        apply_updates, global_step = p.get('apply_updates', False), model.global_step
        p['apply_updates'] = False
        try:
            with torch.cuda.stream(self.stream):
                for _ in range(warmup):
                    model.run_step(prepared=self.prepared)
        finally:
            p['apply_updates'] = apply_updates
            model.global_step = global_step
        while model._inflight:
            model._inflight.pop(0).synchronize()
        torch.cuda.synchronize()
        # drop the eager step's autograd graph before capturing
        model.losses, model.outputs, model.total_loss = {}, {}, None
        model.encoder_hidden_states = model.time_major_states = model.seq_len_encs = None
        for d in model.decoder.values():
            d.stash.clear()
        for st_ in getattr(model, "ctc_stash", {}).values():
            st_.clear()
        import gc
        gc.collect()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _launch_count()
        model.encoder_inputs, model.decoder_inputs, model.seq_len, model.seq_len_target = self.prepared
        failure = None
        try:
            with torch.cuda.graph(self.graph, stream=self.stream, capture_error_mode="relaxed"):
                try:
                    model._step_core()
                except BaseException as e:      # ending an aborted capture raises too and would mask this one
                    failure = e
        except BaseException as e:
            failure = failure or e
        if failure is not None:
            raise RuntimeError("GraphedStep: capturing the step failed: %r" % (failure,)) from failure
        self.launches_per_step = _launch_count() - n0
        cur.synchronize()

    def release(self):
        """Drops the captured graph and its memory pool.  With data parallelism the graph holds NCCL collectives as
        nodes: it MUST be released before torch.distributed.destroy_process_group(), which otherwise blocks forever."""
        self.graph = None
        self.prepared = None
        import gc
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def _signature(self, prepared):
        """What a batch must share with the captured one: the tensor shapes and -- unless the step was captured for a
        bucket, with loop bounds from the padded shapes -- the maximum lengths that bound its loops."""
        enc_in, dec_in, enc_len, dec_len = prepared
        sig = [tuple(enc_in.shape), None if self.bucket else int(enc_len._host.max())]
        for k in sorted(dec_len):
            sig.append((k, tuple(dec_in[k].shape), None if self.bucket else int(dec_len[k]._host.max())))
        return sig

    def pad_to_bucket(self, batch):
        """Zero / PAD-pads the batch's 2-D and 3-D arrays (logmel frames, id and label columns) up to the captured
        shapes; lengths are untouched, so the padding is masked everywhere.  Raises if the batch does not fit."""
        out = dict(batch)
        for k, cap in self.caps.items():
            v = batch[k]
            v = v.cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
            if v.shape == cap:
                continue
            if len(v.shape) != len(cap) or v.shape[0] != cap[0] or any(a > b for a, b in zip(v.shape, cap)):
                raise ValueError("GraphedStep: %s of shape %s does not fit the captured bucket %s" % (k, v.shape, cap))
            padded = np.zeros(cap, v.dtype)
            padded[tuple(slice(0, n) for n in v.shape)] = v
            out[k] = padded
        return out

    def prefetch(self, batch):
        """Starts the host -> device copy of the NEXT batch on a copy stream (into staging buffers, so it overlaps the
        step that is running); the following `step()` without argument consumes it with device-to-device copies."""
        model = self.model
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=model.device)
            self._staging, self._staged, self._consumed = {}, None, None
        if self.bucket:
            batch = self.pad_to_bucket(batch)
        if self._staged is not None:
            self._staged[1].synchronize()       # the pinned staging buffers are about to be rewritten
        if self._consumed is not None:
            self._copy_stream.wait_event(self._consumed)    # the previous staged batch has been copied out
        static = model._static_inputs
        model._static_inputs = self._staging
        try:
            with torch.cuda.stream(self._copy_stream):
                prepared = model.get_batch(batch)
                done = torch.cuda.Event()
                done.record()
        finally:
            model._static_inputs = static
        if self._signature(prepared) != self.signature:
            raise ValueError("GraphedStep: batch shape / maximum lengths differ from the captured step")
        self._staged = (prepared, done)

    def step(self, batch=None):
        model = self.model
        if batch is None and getattr(self, "_staged", None) is not None:
            _, done = self._staged
            torch.cuda.current_stream().wait_event(done)
            for key, src in self._staging.items():
                model._static_inputs[key].copy_(src, non_blocking=True)
            self._consumed = torch.cuda.Event()
            self._consumed.record()
            self._staged = None
        if batch is not None:
            if self.bucket:
                batch = self.pad_to_bucket(batch)
            prepared = model.get_batch(batch)          # refills the static buffers in place
            if self._signature(prepared) != self.signature or prepared[0] is not self.prepared[0]:
                raise ValueError("GraphedStep: batch shape / maximum lengths differ from the captured step; capture "
                                 "one GraphedStep per shape bucket or use run_step")
        while len(model._inflight) >= 2:
            model._wait(model._inflight.pop(0))
        model._stage_step_seed()               # this replay's dropout key
        self.graph.replay()
        model._step_tail()

    __call__ = step


def _launch_count():
    from ._lib import launch_count
    return launch_count()
