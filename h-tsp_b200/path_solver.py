"""B200PathSolver: drop-in for the reference's rl4cop PathSolver *inference* interface
(rl4cop/train_path_solver.py:170-311) -- same constructor-from-state_dict keys, `.device`,
`.eval()/.train()/.training`, and the same `forward(batch, source_nodes, target_nodes,
val_type, return_pi, group_size, greedy)` call -- with every tensor op replaced by the
sm_100a kernels of libhtsp_b200.so (C ABI, include/htsp_b200.h).  torch is used for device
memory, streams and the host RNG call the reference makes (`torch.randperm`, :256)."""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import _abi
from .weights import EMBED_DIM, FF_DIM, expected_keys, n_layers_of


class _EncLayerParams(torch.nn.Module):
    """Parameter container with the reference's names (models.py:12-28); never called."""

    def __init__(self, H: int, F: int):
        super().__init__()
        self.Wq = torch.nn.Linear(H, H, bias=False)
        self.Wk = torch.nn.Linear(H, H, bias=False)
        self.Wv = torch.nn.Linear(H, H, bias=False)
        self.multi_head_combine = torch.nn.Linear(H, H)
        self.feed_forward = torch.nn.Sequential(torch.nn.Linear(H, F), torch.nn.ReLU(),
                                                torch.nn.Linear(F, H))
        self.norm1 = torch.nn.BatchNorm1d(H)
        self.norm2 = torch.nn.BatchNorm1d(H)


class _EncParams(torch.nn.Module):
    def __init__(self, n_layers: int, H: int, F: int):
        super().__init__()
        self.init_projection_layer = torch.nn.Linear(2, H)
        self.attn_layers = torch.nn.ModuleList([_EncLayerParams(H, F) for _ in range(n_layers)])


class _DecParams(torch.nn.Module):
    def __init__(self, H: int):
        super().__init__()
        for name in ("Wq_graph", "Wq_source", "Wq_target", "Wq_first", "Wq_last", "Wk", "Wv"):
            setattr(self, name, torch.nn.Linear(H, H, bias=False))
        self.multi_head_combine = torch.nn.Linear(H, H)


class _Workspace:
    """Grow-only device scratch buffers keyed by (name, device, CUDA stream) -- the C ABI never allocates.  One buffer
    per stream: two streams driving the same model do not share scratch memory (kernels of one stream are ordered, so
    reuse within a stream is safe)."""

    def __init__(self):
        self._bufs: Dict[tuple, torch.Tensor] = {}

    def get(self, name: str, nbytes: int, device) -> torch.Tensor:
        device = torch.device(device)
        stream = torch.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0
        key = (name, str(device), stream)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf


class B200PathSolver(torch.nn.Module):
    """mode: "fp32" (CUDA-core parity anchor), "x3" (tensor cores, split-fp16, default once
    available) or "fast" (tensor cores, single-fp16 glimpse scores)."""

    def __init__(self, cfg=None, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 n_layers: Optional[int] = None, graph_size: int = 200, mode: str = "x3",
                 device="cuda"):
        super().__init__()
        if cfg is None:
            if n_layers is None:
                n_layers = n_layers_of(state_dict) if state_dict is not None else 6
            cfg = SimpleNamespace(node_dim=2, encoder_type="mha", n_layers=n_layers, n_heads=8,
                                  embedding_dim=EMBED_DIM, add_init_projection=True,
                                  tanh_clipping=10, n_decoding_neighbors=None,
                                  graph_size=graph_size, group_size=graph_size, precision=32,
                                  val_type="x8Aug_nTraj")
        assert cfg.encoder_type == "mha" and cfg.embedding_dim == EMBED_DIM and cfg.n_heads == 8
        assert cfg.node_dim == 2 and cfg.tanh_clipping == 10 and cfg.n_decoding_neighbors is None
        self.cfg = cfg
        self.mode = mode
        self.encoder = _EncParams(cfg.n_layers, EMBED_DIM, FF_DIM)
        self.decoder = _DecParams(EMBED_DIM)
        self._blob: Optional[torch.Tensor] = None
        self._ws = _Workspace()
        self.first_action_override: Optional[torch.Tensor] = None
        self.max_instances_per_launch = 8192
        self.to(device)
        if state_dict is not None:
            self.load_state_dict(state_dict, strict=False)
        self.eval()

    # -- reference surface ---------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return self.encoder.init_projection_layer.weight.device

    def load_state_dict(self, state_dict, strict: bool = True):
        out = super().load_state_dict(state_dict, strict=strict)
        self._blob = None
        return out

    def _apply(self, fn, *a, **k):
        self._blob = None
        return super()._apply(fn, *a, **k)

    # -- weights -------------------------------------------------------------------------
    def _param_versions(self):
        return tuple(int(p._version) for p in self.parameters()) + tuple(int(b._version) for b in self.buffers())

    def _packed_weights(self) -> torch.Tensor:
        # in-place updates (optimizer.step on the lower model during joint training, copy_ into a parameter) bump the
        # tensors' version counters: the packed copy is rebuilt when they moved
        if self._blob is not None and getattr(self, "_blob_versions", None) != self._param_versions():
            self._blob = None
        if self._blob is None:
            lib = _abi.lib()
            L = self.cfg.n_layers
            sd = super().state_dict()
            keys = expected_keys(L)
            host = [sd[k].detach().to("cpu", torch.float32).contiguous() for k in keys]
            assert len(host) == lib.htsp_n_weight_tensors(L)
            arr = (_abi.C.c_void_p * len(host))(*[t.data_ptr() for t in host])
            blob = torch.empty(lib.htsp_weights_blob_bytes(L), dtype=torch.uint8, device=self.device)
            with torch.cuda.device(self.device):
                st = torch.cuda.current_stream().cuda_stream
                _abi.check(lib.htsp_pack_weights(arr, len(host), L, blob.data_ptr(), st),
                           "htsp_pack_weights")
            self._blob = blob
            self._blob_versions = self._param_versions()
        return self._blob

    # -- stages (also used directly by the tests) ----------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def encode(self, x: torch.Tensor, train: bool = False, update_running_stats: bool = True) -> torch.Tensor:
        """MHAEncoder.forward (models.py:55-60): [Bp,N,2] -> [Bp,N,128].  train=True runs the BatchNorm1d layers as
        `model.train()` does (batch statistics, htsp_encode_train) and, like torch, moves running_mean / running_var
        (momentum 0.1, unbiased variance) and num_batches_tracked; the batch statistics of the call are kept in
        `self.last_bn_batch_stats` [L,4,128] (mean1, biased var1, mean2, biased var2)."""
        lib = _abi.lib()
        x = x.to(self.device, torch.float32).contiguous()
        Bp, N, _ = x.shape
        mode = _abi.MODES[self.mode]
        emb = torch.empty(Bp, N, EMBED_DIM, dtype=torch.float32, device=self.device)
        if train:
            L = self.cfg.n_layers
            layers = self.encoder.attn_layers
            gb = torch.stack([torch.stack([l.norm1.weight, l.norm1.bias, l.norm2.weight, l.norm2.bias])
                              for l in layers]).detach().to(self.device, torch.float32).contiguous()
            stats = torch.empty(L, 4, EMBED_DIM, dtype=torch.float32, device=self.device)
            nbytes = lib.htsp_encode_train_workspace_bytes(Bp, N, mode)
            if nbytes == 0:
                raise NotImplementedError("train-mode BatchNorm is served by the tensor-core encoder (mode x3 / fast)")
            ws = self._ws.get("enc", nbytes, self.device)
            with torch.cuda.device(self.device):
                _abi.check(lib.htsp_encode_train(self._packed_weights().data_ptr(), L, gb.data_ptr(), x.data_ptr(),
                                                 Bp, N, mode, emb.data_ptr(), stats.data_ptr(), ws.data_ptr(),
                                                 ws.numel(), self._stream()), "htsp_encode_train")
            self.last_bn_batch_stats = stats
            if update_running_stats:
                M = Bp * N
                with torch.no_grad():
                    for i, l in enumerate(layers):
                        for j, bn in enumerate((l.norm1, l.norm2)):
                            m = bn.momentum if bn.momentum is not None else 0.1
                            bn.running_mean.mul_(1 - m).add_(stats[i, 2 * j], alpha=m)
                            bn.running_var.mul_(1 - m).add_(stats[i, 2 * j + 1] * (M / (M - 1)), alpha=m)
                            bn.num_batches_tracked += 1
            return emb
        nbytes = lib.htsp_encode_workspace_bytes(Bp, N, mode)
        ws = self._ws.get("enc", nbytes, self.device)
        with torch.cuda.device(self.device):
            _abi.check(lib.htsp_encode(self._packed_weights().data_ptr(), self.cfg.n_layers,
                                       x.data_ptr(), Bp, N, mode, emb.data_ptr(), ws.data_ptr(),
                                       ws.numel(), self._stream()), "htsp_encode")
        return emb

    def rollout(self, x: torch.Tensor, emb: torch.Tensor, src: torch.Tensor, tgt: torch.Tensor,
                first: torch.Tensor, forced_pi: Optional[torch.Tensor] = None,
                want_logits: bool = False, sample_seed: Optional[int] = None, want_logp: bool = False):
        """POMO rollout -> (pi [Bp,G,N] i32, reward [Bp,G] f32, logits or None): greedy, or (sample_seed
        given) every row samples its next node from the policy (htsp_decode_sample)."""
        lib = _abi.lib()
        Bp, N, _ = x.shape
        G = first.numel()
        dev = self.device
        mode = _abi.MODES[self.mode]
        src = src.reshape(-1).to(dev, torch.int32).contiguous()
        tgt = tgt.reshape(-1).to(dev, torch.int32).contiguous()
        first = first.reshape(-1).to(dev, torch.int32).contiguous()
        assert src.numel() == Bp and tgt.numel() == Bp
        if forced_pi is not None:
            forced_pi = forced_pi.to(dev, torch.int32).contiguous()
            assert forced_pi.shape == (Bp, G, N)
        pi = torch.empty(Bp, G, N, dtype=torch.int32, device=dev)
        reward = torch.empty(Bp, G, dtype=torch.float32, device=dev)
        logits = (torch.empty(Bp, G, N - 2, N, dtype=torch.float32, device=dev)
                  if want_logits else None)
        nbytes = lib.htsp_decode_workspace_bytes(Bp, N, G, mode)
        ws = self._ws.get("dec", nbytes, dev)
        if sample_seed is not None and want_logp:
            # third return value: log-probability of every sampled tour (train_path_solver.py:394-414, forward only)
            assert forced_pi is None and not want_logits
            logp = torch.empty(Bp, G, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _abi.check(lib.htsp_decode_sample_logp(self._packed_weights().data_ptr(), self.cfg.n_layers,
                                                       x.data_ptr(), emb.data_ptr(), src.data_ptr(), tgt.data_ptr(),
                                                       first.data_ptr(), Bp, N, G, mode, int(sample_seed) & (2**64 - 1),
                                                       pi.data_ptr(), reward.data_ptr(), logp.data_ptr(), ws.data_ptr(),
                                                       ws.numel(), self._stream()), "htsp_decode_sample_logp")
            return pi, reward, logp
        if sample_seed is not None:
            assert forced_pi is None and not want_logits
            with torch.cuda.device(dev):
                _abi.check(lib.htsp_decode_sample(self._packed_weights().data_ptr(), self.cfg.n_layers,
                                                  x.data_ptr(), emb.data_ptr(), src.data_ptr(), tgt.data_ptr(),
                                                  first.data_ptr(), Bp, N, G, mode, int(sample_seed) & (2**64 - 1),
                                                  pi.data_ptr(), reward.data_ptr(), ws.data_ptr(), ws.numel(),
                                                  self._stream()), "htsp_decode_sample")
            return pi, reward, None
        with torch.cuda.device(dev):
            _abi.check(lib.htsp_decode(self._packed_weights().data_ptr(), self.cfg.n_layers,
                                       x.data_ptr(), emb.data_ptr(), src.data_ptr(),
                                       tgt.data_ptr(), first.data_ptr(), Bp, N, G, mode,
                                       _abi.ptr(forced_pi), pi.data_ptr(), reward.data_ptr(),
                                       _abi.ptr(logits), ws.data_ptr(), ws.numel(),
                                       self._stream()), "htsp_decode")
        return pi, reward, logits

    def rollout_debug(self, x, emb, src, tgt, first, forced_pi, dbg_step: int):
        """Test hook (htsp_decode_debug): teacher-forced rollout on the tcgen05 kernel that also
        returns the raw S [2,8,128,NP], normalised O [256,128] and raw U [256,NP] of instance 0
        at decode step `dbg_step`."""
        lib = _abi.lib()
        Bp, N, _ = x.shape
        G = first.numel()
        dev = self.device
        mode = _abi.MODES[self.mode]
        src = src.reshape(-1).to(dev, torch.int32).contiguous()
        tgt = tgt.reshape(-1).to(dev, torch.int32).contiguous()
        first = first.reshape(-1).to(dev, torch.int32).contiguous()
        if forced_pi is not None:
            forced_pi = forced_pi.to(dev, torch.int32).contiguous()
        pi = torch.empty(Bp, G, N, dtype=torch.int32, device=dev)
        reward = torch.empty(Bp, G, dtype=torch.float32, device=dev)
        NP = (N + 15) // 16 * 16
        dbg = torch.zeros(lib.htsp_decode_debug_floats(N), dtype=torch.float32, device=dev)
        nbytes = lib.htsp_decode_workspace_bytes(Bp, N, G, mode)
        ws = self._ws.get("dec", nbytes, dev)
        with torch.cuda.device(dev):
            _abi.check(lib.htsp_decode_debug(self._packed_weights().data_ptr(), self.cfg.n_layers,
                                             x.data_ptr(), emb.data_ptr(), src.data_ptr(),
                                             tgt.data_ptr(), first.data_ptr(), Bp, N, G, mode,
                                             _abi.ptr(forced_pi), pi.data_ptr(), reward.data_ptr(),
                                             ws.data_ptr(), ws.numel(), dbg.data_ptr(), int(dbg_step),
                                             self._stream()), "htsp_decode_debug")
        nS = 2 * 8 * 128 * NP
        S = dbg[:nS].reshape(2, 8, 128, NP)
        O = dbg[nS:nS + 256 * 128].reshape(256, 128)
        U = dbg[nS + 256 * 128:nS + 256 * 128 + 256 * NP].reshape(256, NP)
        self.last_debug_cycles = dbg[nS + 256 * 128 + 256 * NP:].reshape(32, 8)[:, :6]
        return pi, reward, S, O, U

    def select_best(self, reward: torch.Tensor, pi: torch.Tensor, n_aug: int):
        """train_path_solver.py:293-306 -> (best_len [B] f32, best_pi [B,N] i64, best_row [B])."""
        lib = _abi.lib()
        Bp, G, N = pi.shape
        B = Bp // n_aug
        dev = self.device
        best_len = torch.empty(B, dtype=torch.float32, device=dev)
        best_row = torch.empty(B, dtype=torch.int32, device=dev)
        best_pi = torch.empty(B, N, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _abi.check(lib.htsp_select_best(reward.data_ptr(), pi.data_ptr(), n_aug, B, G, N,
                                            best_len.data_ptr(), best_row.data_ptr(),
                                            best_pi.data_ptr(), self._stream()), "htsp_select_best")
        return best_len, best_pi, best_row

    def draw_first_action(self, N: int, G: int) -> torch.Tensor:
        """The reference's `torch.randperm(N, device=self.device)[None, :G]` (:256): same call,
        same generator, so a seeded caller sees the same POMO starts on both sides."""
        if self.first_action_override is not None:
            fa = self.first_action_override.reshape(-1)
            assert fa.numel() == G
            if fa.device != torch.device(self.device):
                # copied once: a host tensor would be copied, and the stream synchronised, in every call
                fa = fa.to(self.device)
                self.first_action_override = fa
            return fa
        return torch.randperm(N, device=self.device)[:G]

    # -- PathSolver.forward (train_path_solver.py:224-311) ---------------------------------
    @torch.no_grad()
    def forward(self, batch, source_nodes, target_nodes, val_type="x8Aug_2Traj", return_pi=False,
                group_size=2, greedy=True):
        # greedy=False (train_path_solver.py:266-271): rows sample their next node in the kernel; the seed
        # is drawn from torch's global generator, so torch.manual_seed controls it (the reference consumes the
        # same generator through multinomial: equal in distribution, not bit-wise)
        n_nodes = batch.shape[1]
        key_blocks = self.mode == "fast" and 208 < n_nodes <= 576          # csrc/decode_tc2.cu, NKB = 2 | 3
        if not greedy and (self.mode == "fp32" or not (key_blocks or (17 <= n_nodes <= 208 and int(group_size) <= 256))):
            raise NotImplementedError("sampling rollouts (greedy=False) are served by the tcgen05 kernels only: "
                                      "tensor-core mode, 17 <= N <= 208, group_size <= 256 (mode fast: N <= 576)")
        sample_seed = None if greedy else int(torch.randint(0, 2**62, (1,)).item())
        lib = _abi.lib()
        val_type = val_type or self.cfg.val_type
        dev = self.device
        batch = batch.to(dev, torch.float32).contiguous()
        B0, N, C = batch.shape
        assert C == 2
        G = int(group_size)
        assert G <= self.cfg.graph_size
        n_aug = 8 if "x8Aug" in val_type else 1
        for name, t in (("source_nodes", source_nodes), ("target_nodes", target_nodes)):
            # host tensors are checked here (the reference's gather raises IndexError); device tensors are not read
            # back: the kernels clamp every index and a flagged call returns NaN lengths (htsp_decode)
            if not t.is_cuda and t.numel() and (int(t.min()) < 0 or int(t.max()) >= N):
                raise IndexError(f"{name} outside [0, {N})")
        src = source_nodes.reshape(-1).to(dev)
        tgt = target_nodes.reshape(-1).to(dev)
        if n_aug == 8:
            x = torch.empty(8 * B0, N, 2, dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _abi.check(lib.htsp_augment8(batch.data_ptr(), B0, N, x.data_ptr(), self._stream()),
                           "htsp_augment8")
            # the reference repeat_interleaves the endpoints although the coordinates are
            # block-major (:236-238); reproduced as is
            src = torch.repeat_interleave(src, 8, dim=0)
            tgt = torch.repeat_interleave(tgt, 8, dim=0)
        else:
            x = batch
        first = self.draw_first_action(N, G)
        pi, reward = self._rollout_chunked(x, src, tgt, first, n_aug, sample_seed)
        if val_type == "noAug_1Traj":
            max_reward, best_pi = reward, pi.to(torch.int64)
        else:
            best_len, best_pi, _ = self.select_best(reward, pi, n_aug)
            max_reward = -best_len
            best_pi = best_pi.reshape(B0, 1, N) if n_aug == 1 else best_pi.reshape(1, B0, 1, N)
        if return_pi:
            return -max_reward, best_pi.squeeze()
        return -max_reward

    # -- PathSolver.training_step, forward half (train_path_solver.py:352-434) ----------------
    @torch.no_grad()
    def training_step_forward(self, batch: torch.Tensor, group_size: Optional[int] = None, norm_reward: bool = False,
                              source_nodes: Optional[torch.Tensor] = None, target_nodes: Optional[torch.Tensor] = None):
        """Everything `training_step` computes before `loss.backward()`: random endpoints (:362-366), the encoder with
        TRAIN-mode BatchNorm (batch statistics, running statistics updated), a sampling rollout from `randperm` starts
        that also returns the log-probability of every sampled tour (htsp_decode_sample_logp), the advantage
        (:402-410), `loss = (-advantage * log_prob).mean()` and the logged scalars.  Returns a dict with the tensors
        (reward, log_prob, advantage [B,G], pi [B,G,N]) and the scalars.  The backward pass is not built
        (DESIGN.md section 7): this is the forward half, for evaluation of a training batch and for parity."""
        if self.mode != "fast":
            raise NotImplementedError("training_step_forward needs mode='fast' (log-probabilities come from the second tcgen05 kernel)")
        dev = self.device
        batch = batch.to(dev, torch.float32).contiguous()
        B, N, _ = batch.shape
        G = int(group_size or self.cfg.group_size)
        assert G <= self.cfg.graph_size and G <= N
        if source_nodes is None or target_nodes is None:
            order = torch.argsort(torch.rand(B, N, device=dev))             # :362-366
            source_nodes, target_nodes = order[:, 0], order[:, 1]
        first = self.draw_first_action(N, G)                                # :374
        seed = int(torch.randint(0, 2**62, (1,)).item())
        emb = self.encode(batch, train=self.training)
        pi, reward, logp = self.rollout(batch, emb, source_nodes.reshape(-1), target_nodes.reshape(-1), first,
                                        sample_seed=seed, want_logp=True)
        r = reward
        if G != 1:
            adv = r - r.mean(dim=1, keepdim=True)
            if norm_reward:
                adv = adv / (r.std(dim=1, keepdim=True) + 1e-5)
        else:
            adv = r
        loss = (-adv * logp).mean()
        return {"loss": loss, "length": float(-r.max(dim=1).values.mean()), "advantage": adv, "log_prob": logp,
                "reward": r, "pi": pi, "entropy": -(logp.exp() * logp).mean(), "r_mean": r.mean(),
                "r_std": r.std(dim=1).mean() if G != 1 else r.new_zeros(()), "source_nodes": source_nodes,
                "target_nodes": target_nodes, "first_action": first, "sample_seed": seed}

    def _rollout_chunked(self, x, src, tgt, first, n_aug, sample_seed=None):
        Bp = x.shape[0]
        cap = self.max_instances_per_launch
        if Bp <= cap:
            emb = self.encode(x)
            pi, reward, _ = self.rollout(x, emb, src, tgt, first, sample_seed=sample_seed)
            return pi, reward
        pis, rs = [], []
        for s in range(0, Bp, cap):
            xs = x[s:s + cap]
            emb = self.encode(xs)
            seed_s = None if sample_seed is None else sample_seed + 0x9E3779B97F4A7C15 * (s // cap)   # own stream per chunk
            p, r, _ = self.rollout(xs, emb, src[s:s + cap], tgt[s:s + cap], first, sample_seed=seed_s)
            pis.append(p)
            rs.append(r)
        return torch.cat(pis), torch.cat(rs)
