"""DppoEngine: thin host wrapper over the C ABI (include/dppo_b200.h).

It owns the dppo_ctx and its workspace (a torch uint8 CUDA tensor -- device memory container
only), validates tensors, passes raw ``data_ptr()``s and the current CUDA stream, and raises
``RuntimeError`` with ``dppo_last_error()`` on any non-zero return.  No computation happens in
Python or in torch ops.
"""
import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L
from .layout import ParamLayout


class DppoEngine:
    """One engine per GPU / process (the replacement of one tf.Session + PPO graph,
    /root/reference/main.py:31, PPO.py:8-53)."""

    LOSS_NAMES = ("policyLoss", "valueLoss", "entropyLoss", "total_loss", "mean_kl", "clip_frac",
                  "vclip_frac", "rows")

    def __init__(self, obs_dim: int, act_dim: int, hidden: Sequence[int], max_rows: int, device: int = 0,
                 entcoeff: float = 0.01, vcoeff: float = 0.5, clip_param: float = 0.2,
                 gemm_mode: int = L.GEMM_FP32_SIMT, pd_type: int = L.PD_DIAG_GAUSSIAN,
                 value_loss_mode: int = L.VLOSS_PER_SAMPLE):
        """pd_type = L.PD_CATEGORICAL: act_dim is the number of classes, actions are int32 [B, 1] (PPO.py:26)."""
        self.lib = L.load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("DppoEngine needs a CUDA device (sm_100a); there is no CPU path")
        if len(hidden) > L.MAX_HIDDEN:
            raise ValueError(f"at most {L.MAX_HIDDEN} hidden layers")
        self.is_categorical = int(pd_type) == L.PD_CATEGORICAL
        self.pd_width = act_dim if self.is_categorical else 2 * act_dim
        self.act_cols = 1 if self.is_categorical else act_dim
        self.layout = ParamLayout(obs_dim, hidden, act_dim, self.pd_width)
        cfg = L.DppoConfig()
        cfg.pd_type = int(pd_type)
        cfg.value_loss_mode = int(value_loss_mode)
        cfg.obs_dim, cfg.act_dim, cfg.n_hidden = obs_dim, act_dim, len(hidden)
        for i, h in enumerate(hidden):
            cfg.hidden[i] = int(h)
        cfg.max_rows, cfg.device, cfg.gemm_mode = int(max_rows), int(device), int(gemm_mode)
        cfg.entcoeff, cfg.vcoeff, cfg.clip_param = entcoeff, vcoeff, clip_param
        self.cfg = cfg
        self.device = torch.device("cuda", device)
        self.obs_dim, self.act_dim, self.hidden, self.max_rows = obs_dim, act_dim, tuple(hidden), int(max_rows)
        self.n_params = int(self.lib.dppo_param_count(C.byref(cfg)))
        if self.n_params != self.layout.n_params:
            raise RuntimeError("C and Python parameter layouts disagree")
        self.grad_len = int(self.lib.dppo_grad_buffer_len(C.byref(cfg)))
        nbytes = int(self.lib.dppo_workspace_bytes(C.byref(cfg)))
        if nbytes == 0:
            raise RuntimeError("dppo_workspace_bytes: " + L.last_error(self.lib))
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        ctx = C.c_void_p()
        self._check(self.lib.dppo_create(C.byref(cfg), self.workspace.data_ptr(), nbytes, C.byref(ctx)))
        self.ctx = ctx
        self.world = 1
        # a dppo_ctx is not thread-safe (one workspace, one side stream): callers that share an engine between
        # threads (the reference's Worker.work / Chief.check threads funnel into one session) hold this lock per call
        import threading
        self.lock = threading.RLock()

    # ------------------------------------------------------------------ helpers
    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"libdppo_b200 error {rc}: {L.last_error(self.lib)}")

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _dev(self, t: torch.Tensor, numel: Optional[int] = None, name: str = "tensor"):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise TypeError(f"{name}: expected a contiguous float32 CUDA tensor")
        if t.device != self.device:
            raise ValueError(f"{name}: on {t.device}, engine is on {self.device}")
        if numel is not None and t.numel() != numel:
            raise ValueError(f"{name}: expected {numel} elements, got {t.numel()}")
        return t.data_ptr()

    def _act(self, t: torch.Tensor, rows: int, name: str = "act"):
        """action tensor: fp32 [rows, A] (DiagGaussianPd) or int32 [rows, 1] (CategoricalPd, PPO.py:26)"""
        want = torch.int32 if self.is_categorical else torch.float32
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == want and t.is_contiguous()):
            raise TypeError(f"{name}: expected a contiguous {want} CUDA tensor")
        if t.device != self.device:
            raise ValueError(f"{name}: on {t.device}, engine is on {self.device}")
        if t.numel() != rows * self.act_cols:
            raise ValueError(f"{name}: expected {rows * self.act_cols} elements, got {t.numel()}")
        return t.data_ptr()

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx.value:
            self.lib.dppo_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def new_grad_buffer(self):
        return torch.zeros(self.grad_len, dtype=torch.float32, device=self.device)

    # ------------------------------------------------------------------ hot path
    def gae(self, rew, predv, done, gamma: float, lam: float, mode: int = L.GAE_REFERENCE, normalize: bool = True,
            last_v=None, last_done=None, adv_out=None, ret_out=None):
        """[T,N] time-major buffers -> (adv, ret).  Worker.py:82-92."""
        T, N = rew.shape
        adv = adv_out if adv_out is not None else torch.empty_like(rew)
        ret = ret_out if ret_out is not None else torch.empty_like(rew)
        self._check(self.lib.dppo_gae(
            self.ctx, self._dev(rew, T * N, "rew"), self._dev(predv, T * N, "predv"), self._dev(done, T * N, "done"),
            self._dev(last_v, N, "last_v") if last_v is not None else None,
            self._dev(last_done, N, "last_done") if last_done is not None else None,
            T, N, float(gamma), float(lam), int(mode), int(bool(normalize)),
            self._dev(adv, T * N, "adv_out"), self._dev(ret, T * N, "ret_out"), self._stream()))
        return adv, ret

    def forward(self, params, obs):
        """(v[B], pdflat[B,2A]) = Model.FC forward.  Model.py:11-15."""
        B = obs.shape[0]
        v = torch.empty(B, dtype=torch.float32, device=self.device)
        pd = torch.empty(B, self.pd_width, dtype=torch.float32, device=self.device)
        self._check(self.lib.dppo_forward(self.ctx, self._dev(params, self.n_params, "params"),
                                          self._dev(obs, B * self.obs_dim, "obs"), B, v.data_ptr(), pd.data_ptr(),
                                          self._stream()))
        return v, pd

    def act(self, params, obs, seed: int, offset: int):
        """(action[B,A], v[B]) = piPD.sample(), pipredv.  PPO.py:23, Worker.py:146."""
        B = obs.shape[0]
        a = (torch.empty(B, 1, dtype=torch.int32, device=self.device) if self.is_categorical else
             torch.empty(B, self.act_dim, dtype=torch.float32, device=self.device))
        v = torch.empty(B, dtype=torch.float32, device=self.device)
        self._check(self.lib.dppo_act(self.ctx, self._dev(params, self.n_params, "params"),
                                      self._dev(obs, B * self.obs_dim, "obs"), B, int(seed), int(offset),
                                      a.data_ptr(), v.data_ptr(), self._stream()))
        return a, v

    def loss_grad(self, params_pi, params_old, obs, act, adv, ret, l_mul: float, grad_out=None, losses_out=None):
        """One worker tower: losses + d total_loss / d pipara.  PPO.py:21-46.
        Returns (losses[8], grad_buffer[n_params + 8])."""
        B = obs.shape[0]
        g = grad_out if grad_out is not None else self.new_grad_buffer()
        self._check(self.lib.dppo_loss_grad(
            self.ctx, self._dev(params_pi, self.n_params, "params_pi"), self._dev(params_old, self.n_params, "params_old"),
            self._dev(obs, B * self.obs_dim, "obs"), self._act(act, B),
            self._dev(adv, B, "adv"), self._dev(ret, B, "ret"), B, float(l_mul),
            self._dev(losses_out, L.NUM_LOSS_SLOTS, "losses_out") if losses_out is not None else None,
            self._dev(g, self.grad_len, "grad_out"), self._stream()))
        return g[self.n_params:], g

    # ------------------------------------------------------------------ collective + optimiser
    @staticmethod
    def unique_id() -> bytes:
        lib = L.load_library()
        buf = C.create_string_buffer(L.NCCL_UNIQUE_ID_BYTES)
        rc = lib.dppo_comm_unique_id(buf)
        if rc != 0:
            raise RuntimeError(f"dppo_comm_unique_id: {L.last_error(lib)}")
        return buf.raw

    def comm_init(self, rank: int, world: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, L.NCCL_UNIQUE_ID_BYTES)
        self._check(self.lib.dppo_comm_init(self.ctx, rank, world, buf))
        self.world = world

    def comm_peer_export(self, rank: int, world: int) -> bytes:
        """NCCL-free bootstrap of the fused peer-memory exchange, step 1: this rank's IPC handle blob."""
        buf = C.create_string_buffer(L.PEER_HANDLE_BYTES)
        self._check(self.lib.dppo_comm_peer_export(self.ctx, int(rank), int(world), buf))
        self._peer_world = int(world)
        return buf.raw

    def comm_peer_import(self, blobs):
        """step 2: the blobs of ALL ranks, in rank order (moved between the processes by the caller)."""
        data = b"".join(blobs)
        if len(data) != self._peer_world * L.PEER_HANDLE_BYTES:
            raise ValueError("comm_peer_import: need one handle blob per rank")
        self._check(self.lib.dppo_comm_peer_import(self.ctx, C.create_string_buffer(data, len(data))))
        self.world = self._peer_world

    def allreduce_avg(self, buf, n: Optional[int] = None):
        """In-place mean over ranks.  PPO._average_gradients, PPO.py:55-65."""
        n = buf.numel() if n is None else n
        self._dev(buf, None, "buf")
        self._check(self.lib.dppo_allreduce_avg(self.ctx, buf.data_ptr(), n, self._stream()))
        return buf

    def adam_apply(self, params, grad, m, v, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8):
        """tf.train.AdamOptimizer.apply_gradients.  PPO.py:20,53."""
        n = self.n_params
        self._dev(grad, None, "grad")
        self._check(self.lib.dppo_adam_apply(self.ctx, self._dev(params, n, "params"), grad.data_ptr(),
                                             self._dev(m, n, "m"), self._dev(v, n, "v"), n, int(step), float(lr),
                                             beta1, beta2, eps, self._stream()))

    def sync_old(self, params_pi, params_old):
        """oldpi <- pi.  PPO.py:47."""
        self._check(self.lib.dppo_sync_old(self.ctx, self._dev(params_pi, self.n_params, "params_pi"),
                                           self._dev(params_old, self.n_params, "params_old"), self._stream()))

    def update(self, params_pi, params_old, obs, act, rew, predv, done, gamma, lam, l_mul, grad_out=None,
               adam_m=None, adam_v=None, step: int = 0, lr: float = 0.0, beta1=0.9, beta2=0.999, eps=1e-8,
               gae_mode: int = L.GAE_REFERENCE):
        """One whole update on device-resident buffers in one call: GAE (side stream) -> forward -> loss -> backward ->
        allreduce (if a communicator exists) -> Adam (if adam_m/adam_v given).  Chief.py:64 + Worker.py:82-92."""
        T, N = rew.shape
        B = T * N
        g = grad_out if grad_out is not None else self.new_grad_buffer()
        self._check(self.lib.dppo_update(
            self.ctx, self._dev(params_pi, self.n_params, "params_pi"), self._dev(params_old, self.n_params, "params_old"),
            self._dev(obs, B * self.obs_dim, "obs"), self._act(act, B), self._dev(rew, B, "rew"),
            self._dev(predv, B, "predv"), self._dev(done, B, "done"), T, N, float(gamma), float(lam), int(gae_mode),
            float(l_mul), self._dev(adam_m, self.n_params, "adam_m") if adam_m is not None else None,
            self._dev(adam_v, self.n_params, "adam_v") if adam_v is not None else None, int(step), float(lr),
            beta1, beta2, eps, self._dev(g, self.grad_len, "grad_out"), self._stream()))
        return g[self.n_params:], g

    # ------------------------------------------------------------------ epochs x minibatches on the device
    def update_epochs(self, params_pi, params_old, obs, act, rew, predv, done, gamma, lam, l_mul, adam_m, adam_v,
                      step0: int, lr: float, n_epochs: int = 1, minibatch_rows: int = 0, shuffle: bool = False,
                      seed: int = 0, beta1=0.9, beta2=0.999, eps=1e-8, grad_out=None, losses_out=None,
                      gae_mode: int = L.GAE_REFERENCE):
        """GAE once, then n_epochs x ceil(T*N / minibatch_rows) x [forward, loss, backward, tower mean, Adam] with no host
        involvement.  n_epochs=UPDATE_STEPS, minibatch_rows=0 (whole rollout buffer) is Chief.py:64's loop.
        Returns (losses [n_epochs * n_minibatches, 8] device tensor, grad buffer of the last minibatch)."""
        T, N = rew.shape
        B = T * N
        mb = B if minibatch_rows in (0, None) else min(int(minibatch_rows), B)
        n_mb = (B + mb - 1) // mb
        g = grad_out if grad_out is not None else self.new_grad_buffer()
        if losses_out is None:
            losses_out = torch.empty(n_epochs * n_mb, L.NUM_LOSS_SLOTS, dtype=torch.float32, device=g.device)
        self._check(self.lib.dppo_update_epochs(
            self.ctx, self._dev(params_pi, self.n_params, "params_pi"), self._dev(params_old, self.n_params, "params_old"),
            self._dev(obs, B * self.obs_dim, "obs"), self._act(act, B), self._dev(rew, B, "rew"),
            self._dev(predv, B, "predv"), self._dev(done, B, "done"), T, N, float(gamma), float(lam), int(gae_mode),
            int(n_epochs), int(mb), int(bool(shuffle)), int(seed), float(l_mul),
            self._dev(adam_m, self.n_params, "adam_m"), self._dev(adam_v, self.n_params, "adam_v"), int(step0), float(lr),
            beta1, beta2, eps, self._dev(losses_out, n_epochs * n_mb * L.NUM_LOSS_SLOTS, "losses_out"),
            self._dev(g, self.grad_len, "grad_out"), self._stream()))
        return losses_out, g

    def minibatch_perm(self, n: int, seed: int, epoch: int):
        """Row permutation of [0, n) the scheduler uses for `epoch` (int32 device tensor)."""
        out = torch.empty(n, dtype=torch.int32, device=self.device)
        self._check(self.lib.dppo_minibatch_perm(self.ctx, int(n), int(seed), int(epoch), out.data_ptr(), self._stream()))
        return out

    # ------------------------------------------------------------------ host-buffer entry (e2e)
    def update_host(self, params_pi, params_old, obs_h, act_h, rew_h, predv_h, done_h, gamma, lam, l_mul,
                    grad_out=None, gae_mode: int = L.GAE_REFERENCE, adam_m=None, adam_v=None, step: int = 0,
                    lr: float = 0.0, beta1=0.9, beta2=0.999, eps=1e-8):
        """Whole update from pinned HOST buffers (what Chief._train feeds, Chief.py:54-64):
        chunked H2D overlapped with GAE + loss_grad (+ allreduce) (+ Adam when adam_m/adam_v are given), losses read
        back.  Synchronises."""
        T, N = rew_h.shape
        B = T * N
        for name, t, n in (("obs_h", obs_h, B * self.obs_dim), ("act_h", act_h, B * self.act_cols),
                           ("rew_h", rew_h, B), ("predv_h", predv_h, B), ("done_h", done_h, B)):
            want = torch.int32 if (self.is_categorical and name == "act_h") else torch.float32
            if t.is_cuda or t.dtype != want or not t.is_contiguous() or t.numel() != n:
                raise TypeError(f"{name}: expected a contiguous {want} host tensor with {n} elements")
            if not t.is_pinned():
                raise TypeError(f"{name}: host buffers must be pinned (torch.Tensor.pin_memory())")
        g = grad_out if grad_out is not None else self.new_grad_buffer()
        losses = (C.c_float * L.NUM_LOSS_SLOTS)()
        self._check(self.lib.dppo_update_host(
            self.ctx, self._dev(params_pi, self.n_params, "params_pi"), self._dev(params_old, self.n_params, "params_old"),
            obs_h.data_ptr(), act_h.data_ptr(), rew_h.data_ptr(), predv_h.data_ptr(), done_h.data_ptr(), T, N,
            float(gamma), float(lam), int(gae_mode), float(l_mul),
            self._dev(adam_m, self.n_params, "adam_m") if adam_m is not None else None,
            self._dev(adam_v, self.n_params, "adam_v") if adam_v is not None else None, int(step), float(lr),
            beta1, beta2, eps, losses, self._dev(g, self.grad_len, "grad_out"), self._stream()))
        return list(losses), g

    # ------------------------------------------------------------------ Chief._train in one call (host towers)
    def train_host(self, params_pi, params_old, towers, n_steps: int, l_mul: float, adam_m, adam_v, step0: int, lr: float,
                   beta1=0.9, beta2=0.999, eps=1e-8, grad_out=None):
        """Chief.py:54-64: `towers` = [(obs, act, adv, ret), ...] HOST tensors (one tuple per worker's rollout batch; pinned
        memory makes the upload asynchronous).  One upload, n_steps (= UPDATE_STEPS) x [forward, loss, backward, tower
        mean, Adam] on the device, oldpi cached.  Returns (losses [n_steps][8] as a list of lists, gradient buffer)."""
        k = len(towers)
        if k < 1:
            raise ValueError("train_host: no towers")
        ptrs = [(C.c_void_p * k)() for _ in range(4)]
        rows = (C.c_int * k)()
        keep = []
        for i, (obs, act, adv, ret) in enumerate(towers):
            want_act = torch.int32 if self.is_categorical else torch.float32
            n = int(obs.shape[0])
            for name, t, dt, numel in (("obs", obs, torch.float32, n * self.obs_dim), ("act", act, want_act, n * self.act_cols),
                                       ("adv", adv, torch.float32, n), ("ret", ret, torch.float32, n)):
                if not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != dt or not t.is_contiguous() or t.numel() != numel:
                    raise TypeError(f"train_host: tower {i} `{name}` must be a contiguous {dt} host tensor with {numel} elements")
            for j, t in enumerate((obs, act, adv, ret)):
                ptrs[j][i] = t.data_ptr()
            rows[i] = n
            keep.append((obs, act, adv, ret))
        g = grad_out if grad_out is not None else self.new_grad_buffer()
        losses = (C.c_float * (L.NUM_LOSS_SLOTS * int(n_steps)))()
        self._check(self.lib.dppo_train_host(
            self.ctx, self._dev(params_pi, self.n_params, "params_pi"), self._dev(params_old, self.n_params, "params_old"), k,
            ptrs[0], ptrs[1], ptrs[2], ptrs[3], rows, int(n_steps), float(l_mul), self._dev(adam_m, self.n_params, "adam_m"),
            self._dev(adam_v, self.n_params, "adam_v"), int(step0), float(lr), beta1, beta2, eps, losses,
            self._dev(g, self.grad_len, "grad_out"), self._stream()))
        flat = list(losses)
        return [flat[i * L.NUM_LOSS_SLOTS:(i + 1) * L.NUM_LOSS_SLOTS] for i in range(int(n_steps))], g

    # ------------------------------------------------------------------ checkpoint / resume
    def save_state(self, path: str, params_pi, params_old, adam_m=None, adam_v=None, step: int = 0):
        """save_state (Others/tf_util.py:276-279): pi, oldpi, Adam slots and the optimiser step to one file."""
        n = self.n_params
        self._check(self.lib.dppo_save_state(
            self.ctx, str(path).encode(), self._dev(params_pi, n, "params_pi"), self._dev(params_old, n, "params_old"),
            self._dev(adam_m, n, "adam_m") if adam_m is not None else None,
            self._dev(adam_v, n, "adam_v") if adam_v is not None else None, int(step), self._stream()))

    def load_state(self, path: str, params_pi, params_old, adam_m=None, adam_v=None) -> int:
        """load_state (Others/tf_util.py:271-273) into the given tensors; returns the saved optimiser step."""
        n = self.n_params
        step = C.c_int64(0)
        self._check(self.lib.dppo_load_state(
            self.ctx, str(path).encode(), self._dev(params_pi, n, "params_pi"), self._dev(params_old, n, "params_old"),
            self._dev(adam_m, n, "adam_m") if adam_m is not None else None,
            self._dev(adam_v, n, "adam_v") if adam_v is not None else None, C.byref(step), self._stream()))
        return int(step.value)

    # ------------------------------------------------------------------ CategoricalPd methods
    def categorical(self, op: str, logits, arg=None, seed: int = 0, offset: int = 0):
        """distributions.py:124-159 on a [B,ncat] logits tensor: neglogp(arg=int32 x[B]) / entropy / kl(arg=other logits)
        -> fp32 [B]; sample / mode -> int32 [B]."""
        B, n = logits.shape
        code = self.PD_OPS[op]
        out = torch.empty(B, dtype=torch.int32 if code >= 3 else torch.float32, device=self.device)
        a = None
        if code == 0:
            if not (arg.is_cuda and arg.dtype == torch.int32 and arg.is_contiguous() and arg.numel() == B):
                raise TypeError("categorical neglogp: x must be a contiguous int32 CUDA tensor with B elements")
            a = arg.data_ptr()
        elif code == 2:
            a = self._dev(arg, B * n, "other")
        self._check(self.lib.dppo_categorical(self.ctx, code, self._dev(logits, B * n, "logits"), a, B, n, int(seed),
                                              int(offset), out.data_ptr(), self._stream()))
        return out

    # ------------------------------------------------------------------ DiagGaussianPd methods
    PD_OPS = {"neglogp": 0, "entropy": 1, "kl": 2, "sample": 3, "mode": 4}

    def diag_gaussian(self, op: str, flat, arg=None, seed: int = 0, offset: int = 0):
        """distributions.py:193-205 on a [B,2A] flat tensor; returns [B] or [B,A]."""
        B, A2 = flat.shape
        A = A2 // 2
        out = torch.empty((B, A) if op in ("sample", "mode") else (B,), dtype=torch.float32, device=self.device)
        self._check(self.lib.dppo_diag_gaussian(
            self.ctx, self.PD_OPS[op], self._dev(flat, B * A2, "flat"),
            self._dev(arg, None, "arg") if arg is not None else None, B, A, int(seed), int(offset),
            out.data_ptr(), self._stream()))
        return out

    # ------------------------------------------------------------------ launch accounting
    def profile_reset(self, timing: bool = False):
        self._check(self.lib.dppo_profile_reset(self.ctx, int(timing)))

    def profile_read(self):
        """{family: (launches, total_ms)} since the last reset."""
        n = len(L.PROFILE_KINDS)
        cnt = (C.c_int64 * n)(); ms = (C.c_double * n)()
        k = self.lib.dppo_profile_read(self.ctx, cnt, ms, n)
        if k < 0:
            self._check(k)
        return {name: (int(cnt[i]), float(ms[i])) for i, name in enumerate(L.PROFILE_KINDS)}

    # ------------------------------------------------------------------ unit-test hook
    def dbg_gemm(self, mode, A, B, M, N, K, bias=None, mask=None, relu=False, want_bias_out=False, aux_out=None):
        """aux_out: a caller-provided tensor passed in the bias_out slot (mode 98: int64[128] kernel time line)."""
        Cm = torch.empty(M, N, dtype=torch.float32, device=self.device)
        bo = torch.empty(N, dtype=torch.float32, device=self.device) if want_bias_out else aux_out
        self._check(self.lib.dppo_dbg_gemm(self.ctx, mode, M, N, K, self._dev(A, None, "A"), self._dev(B, None, "B"),
                                           Cm.data_ptr(), bias.data_ptr() if bias is not None else None,
                                           mask.data_ptr() if mask is not None else None, int(relu),
                                           bo.data_ptr() if bo is not None else None, self._stream()))
        return (Cm, bo) if want_bias_out else Cm
