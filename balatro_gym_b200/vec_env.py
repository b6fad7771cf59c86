"""BalatroVecEnv — the vector-env entry point: N Balatro envs as device-resident record arrays
advanced by the sm_100a kernels through the C-ABI (include/bgym.h).

Mirrors the reference's step path for N envs at once:
    BalatroEnv.reset   balatro_gym/balatro_env_2.py:505-558
    BalatroEnv.step    balatro_gym/balatro_env_2.py:616-1064, 1174-1392
and the vector conventions of its consumers (SB3 `SubprocVecEnv`, hpc_train.py:57-72):
same-step autoreset, `step_async/step_wait`, observations as a dict of arrays.

PyTorch is used for device memory and streams only; there is no CPU fallback.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from . import layout as L
from . import _lib

# observation field -> (byte offset, torch dtype name, count)
_OBS_FIELDS = {}
for _name in L.OBS_DTYPE.names:
    _dt, _off = L.OBS_DTYPE.fields[_name][:2]
    _base = _dt.base
    _cnt = int(np.prod(_dt.shape)) if _dt.shape else 0
    _OBS_FIELDS[_name] = (_off, _base.str, _cnt)

_TORCH_DT = {"|i1": "int8", "|u1": "uint8", "<i2": "int16", "<i4": "int32", "<i8": "int64", "<f4": "float32",
             "<u8": "int64", "<u4": "int32", "<u2": "int16"}


def _field_view(torch, buf_u8, dtype_np, name):
    """Zero-copy typed view of one record field over a [N, record_bytes] uint8 tensor."""
    dt, off = dtype_np.fields[name][:2]
    base = dt.base
    size = dt.itemsize
    tdt = getattr(torch, _TORCH_DT[base.str])
    v = buf_u8[:, off:off + size]
    if base.itemsize > 1:
        v = v.view(tdt)
    elif base.str == "|i1":
        v = v.view(torch.int8)
    return v if dt.shape else v[:, 0]


class _ObsViews(dict):
    """The observation dict of a slab: zero-copy views of the record fields, plus 'action_mask' — the reference's
    int8 [N, 60] array — which is NOT stored in the records (they carry the packed word `action_mask_bits`) and
    is expanded only when somebody asks for it."""

    def __init__(self, views, expand):
        super().__init__(views)
        self._expand = expand

    def __getitem__(self, key):
        if key == "action_mask":
            return self._expand()
        return dict.__getitem__(self, key)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def __contains__(self, key):
        return key == "action_mask" or dict.__contains__(self, key)

    def __iter__(self):
        yield from dict.__iter__(self)
        yield "action_mask"

    def __len__(self):
        return dict.__len__(self) + 1

    def keys(self):
        return list(iter(self))

    def items(self):
        return [(k, self[k]) for k in self]

    def values(self):
        return [self[k] for k in self]


class BalatroVecEnv:
    """N environments on one GPU.

    Parameters
    ----------
    num_envs : number of environments in this slab
    device   : torch device (must be CUDA)
    seed     : base seed; env i of rank r gets seed `seed + env_offset + i` (>= 1)
    autoreset: re-initialise terminated envs inside the step kernel (same-step autoreset)
    env_offset: global index of the first env of this slab (multi-GPU: results do not depend on
               the number of GPUs because seeds are a function of the global env index)
    generator: None (the reference's reset state), "c3" or "c4": every episode — the ones started by
               reset() and the ones started by the in-kernel autoreset — begins from the synthetic state of
               BASELINE configs[2] / configs[3] (5 random jokers, card enhancements / editions / seals;
               "c4" also fills the two consumable slots); BGYM_FLAG_GEN_C3 / BGYM_FLAG_GEN_CONS in include/bgym.h
    """

    num_actions = L.NUM_ACTIONS

    def __init__(self, num_envs: int, device="cuda", seed: int = 1, autoreset: bool = True, env_offset: int = 0,
                 generator: Optional[str] = None):
        torch = _lib.require_cuda()
        self.torch = torch
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.BgymError("BalatroVecEnv needs a CUDA device; there is no CPU fallback")
        self.num_envs = int(num_envs)
        self.autoreset = bool(autoreset)
        self.base_seed = int(seed)
        self.env_offset = int(env_offset)
        if generator not in L.GENERATORS:
            raise ValueError(f"generator must be one of {list(L.GENERATORS)}")
        self.generator = generator
        self._gen_flags = L.GENERATORS[generator]
        n, dev = self.num_envs, self.device
        with torch.cuda.device(dev):
            # env state = dense record arrays (include/bgym.h): hot (144 B), cold (176 B) and the toggle records
            # (32 B: the select path's working set, holding THE copy of hot bytes 16..31); observations = 176-byte
            # records + selection records (16 B: THE copy of selected_cards and the mask word).  `hot` and `obs_buf`
            # (properties) give the record arrays with those fields folded back in.
            self._hot = torch.zeros((n, L.HOT_BYTES), dtype=torch.uint8, device=dev)
            self.tog = torch.zeros((n, L.TOG_BYTES), dtype=torch.uint8, device=dev)
            self.cold = torch.zeros((n, L.COLD_BYTES), dtype=torch.uint8, device=dev)
            self._obs = torch.zeros((n, L.OBS_BYTES), dtype=torch.uint8, device=dev)
            self.sel = torch.zeros((n, L.SEL_BYTES), dtype=torch.uint8, device=dev)
            self.obs_dirty = None     # [n] uint8, allocated by a HostMirror: which observation records a step rewrote
            self.info_buf = torch.zeros((n, L.INFO_BYTES), dtype=torch.uint8, device=dev)
            self.reward = torch.zeros(n, dtype=torch.float64, device=dev)
            self.terminated = torch.zeros(n, dtype=torch.uint8, device=dev)
            self.truncated = torch.zeros(n, dtype=torch.uint8, device=dev)
            self.actions = torch.zeros(n, dtype=torch.int32, device=dev)
            self._mask64 = torch.zeros(n, dtype=torch.int64, device=dev)
            # episode statistics (K6)
            self._ret_acc = torch.zeros(n, dtype=torch.float64, device=dev)
            self._len_acc = torch.zeros(n, dtype=torch.int32, device=dev)
            self.stats = torch.zeros(8, dtype=torch.float64, device=dev)
        self._obs_views: Optional[Dict[str, "torch.Tensor"]] = None
        self._hot_whole = True        # hot bytes 16..31 equal the toggle records (False after a step)
        self._obs_whole = True        # obs records' selected_cards / mask word equal the selection records
        self._obs_current = False     # obs_buf describes the current state (False after the state was changed from outside)
        self._step_count = 0
        self._pending_actions = None
        self.observation_keys = list(L.OBS_KEYS)

    # -- helpers ---------------------------------------------------------------------------------
    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def _ptr(self, t):
        return None if t is None else t.data_ptr()

    def default_seeds(self):
        torch = self.torch
        s = torch.arange(self.num_envs, dtype=torch.int64, device=self.device) + (self.base_seed + self.env_offset)
        s = s % (2 ** 32)
        s = torch.where(s == 0, torch.ones_like(s), s)  # seed 0 is "random" in the reference
        return s.to(torch.int64)

    # -- record arrays with their device-only side arrays folded in ------------------------------------
    def _sync(self, what, direction):
        with self.torch.cuda.device(self.device):
            if what == "state":
                rc = self.lib.bgym_sync_state(self._hot.data_ptr(), self.tog.data_ptr(), self.num_envs, direction, self._stream())
            else:
                rc = self.lib.bgym_sync_obs(self._obs.data_ptr(), self.sel.data_ptr(), self.num_envs, direction, self._stream())
        _lib.check(rc, f"bgym_sync_{what}")

    @property
    def hot(self):
        """[N, 144] hot records, whole (bgym_sync_state folds the toggle records' chunk back in first).  Code that
        WRITES into the returned tensor must call state_written() afterwards."""
        if not self._hot_whole:
            self._sync("state", L.SYNC_TO_RECORDS)
            self._hot_whole = True
        return self._hot

    @property
    def obs_buf(self):
        """[N, 176] observation records, whole (bgym_sync_obs folds the selection records back in first)."""
        if not self._obs_whole:
            self._sync("obs", L.SYNC_TO_RECORDS)
            self._obs_whole = True
        return self._obs

    def state_written(self):
        """The caller has rewritten hot / cold records from outside (through `hot`, `cold`, state_field views):
        rebuild the toggle records and re-emit every observation."""
        self._sync("state", L.SYNC_FROM_RECORDS)
        self._hot_whole = True
        return self.refresh_observations()

    @property
    def mask_words(self):
        """(pointer to env 0's legal-action word, byte stride): the selection records, which every step keeps current."""
        return self.sel.data_ptr() + 8, L.SEL_BYTES

    @property
    def obs(self) -> Dict[str, "object"]:
        """The 31-key observation dict of the reference: zero-copy views of the obs records — selected_cards and the
        mask word are views of the selection records, which is where a step keeps them current —, plus
        `action_mask_bits` (the packed legal-action word, int64) and `action_mask` — the reference's int8[N, 60]
        array, expanded from the word on access (it is not stored: 8 B instead of 60 B per record)."""
        if self._obs_views is None:
            self._obs_views = {k: _field_view(self.torch, self._obs, L.OBS_DTYPE, k) for k in L.OBS_KEYS
                               if k not in ("action_mask", "selected_cards")}
            self._obs_views["selected_cards"] = _field_view(self.torch, self.sel, L.SEL_DTYPE, "selected_cards")
            self._obs_views["action_mask_bits"] = _field_view(self.torch, self.sel, L.SEL_DTYPE, "action_mask_bits")
            shifts = self.torch.arange(L.NUM_ACTIONS, device=self.device)
            bits = self._obs_views["action_mask_bits"]
            self._obs_views = _ObsViews(self._obs_views, lambda: ((bits.unsqueeze(1) >> shifts) & 1).to(self.torch.int8))
        return self._obs_views

    def info_field(self, name):
        return _field_view(self.torch, self.info_buf, L.INFO_DTYPE, name)

    def state_field(self, name):
        """Zero-copy typed view of one state field (from the hot or the cold record array)."""
        if name in L.HOT_FIELD_NAMES:
            return _field_view(self.torch, self.hot, L.HOT_DTYPE, name)      # `hot`: synced first
        return _field_view(self.torch, self.cold, L.COLD_DTYPE, name)

    # -- reset -------------------------------------------------------------------------------------
    def reset(self, seeds=None, decks52=None, reset_mask=None):
        """Reset envs (all, or those with reset_mask != 0).

        seeds   : int64/uint32 tensor [N] (default: base_seed + global env index)
        decks52 : optional uint8 tensor [N, 52] of card codes = replay of the reference's shuffle
                  (deck[i] = code of the i-th card after `random.Random(seed).shuffle`)
        """
        torch = self.torch
        if seeds is None:
            seeds = self.default_seeds()
        # int32 storage holds the uint32 bit pattern
        seeds32 = (torch.as_tensor(seeds, device=self.device).to(torch.int64) & 0xFFFFFFFF)
        seeds32 = torch.where(seeds32 >= 2 ** 31, seeds32 - 2 ** 32, seeds32).to(torch.int32).contiguous()
        if decks52 is not None:
            decks52 = torch.as_tensor(decks52, device=self.device).to(torch.uint8).contiguous()
            assert decks52.shape == (self.num_envs, 52)
        if reset_mask is not None:
            reset_mask = torch.as_tensor(reset_mask, device=self.device).to(torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            rc = self.lib.bgym_reset(self._hot.data_ptr(), self.tog.data_ptr(), self.cold.data_ptr(), self._obs.data_ptr(),
                                     self.sel.data_ptr(), self._ptr(reset_mask), seeds32.data_ptr(), self._ptr(decks52),
                                     self.num_envs, self._gen_flags, self._stream())
        _lib.check(rc, "bgym_reset")
        self._keep = (seeds32, decks52, reset_mask)  # keep inputs alive until the stream has consumed them
        self._hot_whole = self._obs_whole = True     # a reset (masked or not) writes whole records for every env
        self._obs_current = True
        return self.obs

    def refresh_observations(self):
        """Re-emit every env's observation from its state (hot + toggle + cold records).  bgym_reset with an all-zero
        reset mask resets nothing and rewrites the observations.  After WRITING state records call state_written(),
        which rebuilds the toggle records first."""
        torch = self.torch
        zero = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
        seeds = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.bgym_reset(self._hot.data_ptr(), self.tog.data_ptr(), self.cold.data_ptr(), self._obs.data_ptr(),
                                     self.sel.data_ptr(), zero.data_ptr(), seeds.data_ptr(), None, self.num_envs, 0, self._stream())
        _lib.check(rc, "bgym_reset (observation refresh)")
        self._keep = (zero, seeds)
        self._hot_whole = self._obs_whole = True
        self._obs_current = True
        return self.obs

    # -- step --------------------------------------------------------------------------------------
    def step(self, actions=None, draws=None, random_policy: bool = False, want_info: bool = True):
        """One step for all envs.  Returns (obs, reward, terminated, truncated, info_buf).

        actions : int32 tensor [N] on the device (ignored with random_policy=True, in which case the
                  kernel samples a uniform legal action per env and writes it to self.actions)
        draws   : optional uint8 tensor [N, 256] of BgymDraws records (replay mode)
        """
        torch = self.torch
        if not random_policy:
            if actions is None:
                raise ValueError("actions is required unless random_policy=True")
            if actions.dtype != torch.int32 or actions.device != self.device or not actions.is_contiguous():
                actions = torch.as_tensor(actions, device=self.device).to(torch.int32).contiguous()
            act = actions
        else:
            act = self.actions
        flags = (L.FLAG_AUTORESET if self.autoreset else 0) | (L.FLAG_RANDOM_POLICY if random_policy else 0) | self._gen_flags
        if draws is not None:
            draws = torch.as_tensor(draws, device=self.device).contiguous()
            assert draws.dtype == torch.uint8 and draws.shape == (self.num_envs, L.DRAWS_BYTES)
        with torch.cuda.device(self.device):
            rc = self.lib.bgym_step(self._hot.data_ptr(), self.tog.data_ptr(), self.cold.data_ptr(), act.data_ptr(), self._ptr(draws),
                                    self._obs.data_ptr(), self.sel.data_ptr(), self._ptr(self.obs_dirty),
                                    self.reward.data_ptr(), self.terminated.data_ptr(), self.truncated.data_ptr(),
                                    self.info_buf.data_ptr() if want_info else None, self.num_envs, flags, self._stream())
        _lib.check(rc, "bgym_step")
        self._keep = (act, draws)
        self._hot_whole = self._obs_whole = False
        self._obs_current = True
        self._step_count += 1
        return self.obs, self.reward, self.terminated, self.truncated, self.info_buf

    # SB3 VecEnv-style split call
    def step_async(self, actions):
        self._pending_actions = actions

    def step_wait(self):
        out = self.step(self._pending_actions)
        self._pending_actions = None
        return out

    def sample_actions(self, seed: int = 0, out=None):
        """Uniform random legal action per env from the current observation's mask word."""
        out = self.actions if out is None else out
        with self.torch.cuda.device(self.device):
            rc = self.lib.bgym_sample_actions(*self.mask_words, out.data_ptr(), seed & 0xFFFFFFFF,
                                              self._step_count, self.num_envs, self._stream())
        _lib.check(rc, "bgym_sample_actions")
        return out

    def graphed_rollout_step(self, policy: str = "sampler", seed: int = 0):
        """One env-step (policy + step kernels) captured in a CUDA graph; returns a zero-argument callable that
        replays it on the current stream.  A step is six short launches chained by events: replaying them from a
        graph removes the launch gaps between them.  policy = 'sampler' (uniform legal action from the
        observation's mask word, step number kept in a device counter) or 'fused' (sampled inside the step
        kernels).  Results land in the usual buffers (self.obs_buf, self.reward, self.terminated, self.actions)."""
        torch = self.torch
        assert policy in ("sampler", "fused")
        if not hasattr(self, "_step_ctr"):
            self._step_ctr = torch.zeros(1, dtype=torch.int64, device=self.device)
        flags = (L.FLAG_AUTORESET if self.autoreset else 0) | (L.FLAG_RANDOM_POLICY if policy == "fused" else 0) | self._gen_flags
        if not self._obs_current:
            self.refresh_observations()

        def launch():
            st = self._stream()
            if policy == "sampler":
                rc = self.lib.bgym_sample_actions_ctr(*self.mask_words, self.actions.data_ptr(), seed & 0xFFFFFFFF,
                                                      self._step_ctr.data_ptr(), self.num_envs, st)
                _lib.check(rc, "bgym_sample_actions_ctr")
            rc = self.lib.bgym_step(self._hot.data_ptr(), self.tog.data_ptr(), self.cold.data_ptr(), self.actions.data_ptr(), None,
                                    self._obs.data_ptr(), self.sel.data_ptr(), self._ptr(self.obs_dirty),
                                    self.reward.data_ptr(), self.terminated.data_ptr(), self.truncated.data_ptr(), None,
                                    self.num_envs, flags, st)
            _lib.check(rc, "bgym_step")
            self._hot_whole = self._obs_whole = False

        # ONE capture stream per env object: bgym_step keeps its work lists per (device, stream), and a captured graph points
        # at them — re-capturing on the same stream reuses that scratch instead of taking a new slot every time
        if not hasattr(self, "_graph_stream"):
            self._graph_stream = torch.cuda.Stream(device=self.device)
        side = self._graph_stream
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):          # warm-up on the capture stream: per-stream scratch is allocated here
            for _ in range(2):
                launch()
        torch.cuda.current_stream(self.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            launch()
        self._graphs = getattr(self, "_graphs", []) + [graph]

        def replay():
            graph.replay()
            self._hot_whole = self._obs_whole = False

        return replay

    def action_masks(self):
        """uint64 mask word per env computed from the state (bit a = action a legal), as int64."""
        with self.torch.cuda.device(self.device):
            rc = self.lib.bgym_action_mask(self._hot.data_ptr(), self.tog.data_ptr(), self.cold.data_ptr(), self._mask64.data_ptr(),
                                           self.num_envs, self._stream())
        _lib.check(rc, "bgym_action_mask")
        return self._mask64

    def accumulate_stats(self):
        """Fold the last step's (reward, terminated) into the slab statistics (device side)."""
        with self.torch.cuda.device(self.device):
            rc = self.lib.bgym_episode_stats(self.reward.data_ptr(), self.terminated.data_ptr(), self._ret_acc.data_ptr(),
                                             self._len_acc.data_ptr(), self.stats.data_ptr(), self.num_envs, self._stream())
        _lib.check(rc, "bgym_episode_stats")
        return self.stats

    # -- checkpoint (save_state / load_state, balatro_env_2.py:1575-1615) ----------------------------
    def save_state(self):
        ckpt = {"hot": self.hot.clone(), "cold": self.cold.clone(), "obs": self.obs_buf.clone(), "step_count": self._step_count,
                "ret_acc": self._ret_acc.clone(), "len_acc": self._len_acc.clone(), "stats": self.stats.clone()}
        if hasattr(self, "_step_ctr"):
            ckpt["step_ctr"] = self._step_ctr.clone()
        return ckpt

    def load_state(self, ckpt):
        self._hot.copy_(ckpt["hot"])
        self.cold.copy_(ckpt["cold"])
        self._sync("state", L.SYNC_FROM_RECORDS)
        self._hot_whole = True
        self._step_count = ckpt["step_count"]
        self._ret_acc.copy_(ckpt["ret_acc"])
        self._len_acc.copy_(ckpt["len_acc"])
        if "stats" in ckpt:
            self.stats.copy_(ckpt["stats"])
        if "step_ctr" in ckpt:
            if not hasattr(self, "_step_ctr"):
                self._step_ctr = self.torch.zeros(1, dtype=self.torch.int64, device=self.device)
            self._step_ctr.copy_(ckpt["step_ctr"])
        if "obs" in ckpt:
            self._obs.copy_(ckpt["obs"])
            self._sync("obs", L.SYNC_FROM_RECORDS)
            self._obs_whole = True
            self._obs_current = True
        else:
            self.refresh_observations()

    # -- state injection (the C3 state generator of SURVEY §8d; reference: Appendix E) ---------------
    def inject_numpy(self, state_np: np.ndarray):
        """Overwrite all state records from a host array of L.STATE_DTYPE."""
        assert state_np.dtype == L.STATE_DTYPE and state_np.shape == (self.num_envs,)
        raw = state_np.view(np.uint8).reshape(self.num_envs, L.STATE_BYTES)
        self._hot.copy_(self.torch.from_numpy(np.ascontiguousarray(raw[:, :L.HOT_BYTES])))
        self.cold.copy_(self.torch.from_numpy(np.ascontiguousarray(raw[:, L.HOT_BYTES:])))
        self.state_written()

    def state_numpy(self) -> np.ndarray:
        """Host copy of all envs as combined {hot, cold} records (L.STATE_DTYPE)."""
        raw = np.concatenate([self.hot.cpu().numpy(), self.cold.cpu().numpy()], axis=1)
        return np.ascontiguousarray(raw).reshape(-1).view(L.STATE_DTYPE).copy()

    def obs_numpy(self) -> np.ndarray:
        return self.obs_buf.cpu().numpy().reshape(-1).view(L.OBS_DTYPE).copy()

    def info_numpy(self) -> np.ndarray:
        return self.info_buf.cpu().numpy().reshape(-1).view(L.INFO_DTYPE).copy()

    def randomize_c3(self, seed: int = 0):
        """Config-3 state generator, applied on the device right after reset():
        5 distinct random shop-eligible jokers, per-deck-card modifiers i.i.d. (enhancement != NONE
        w.p. 0.25 uniform over 8; edition w.p. 0.1 over {FOIL,HOLO,POLY}; seal w.p. 0.1 over 4)."""
        torch = self.torch
        n, dev = self.num_envs, self.device
        g = torch.Generator(device=dev)
        g.manual_seed(seed + 7919 * (self.env_offset + 1))
        # 5 distinct jokers out of ids 1..145: top-5 of random keys
        keys = torch.rand((n, 145), device=dev, generator=g)
        jk = (keys.topk(5, dim=1).indices + 1).to(torch.uint8)
        self.state_field("joker_id")[:, :5] = jk
        self.state_field("joker_n")[:] = 5
        enh = torch.where(torch.rand((n, 52), device=dev, generator=g) < 0.25,
                          torch.randint(1, 9, (n, 52), device=dev, generator=g), torch.zeros((n, 52), dtype=torch.int64, device=dev))
        ed = torch.where(torch.rand((n, 52), device=dev, generator=g) < 0.1,
                         torch.randint(1, 4, (n, 52), device=dev, generator=g), torch.zeros((n, 52), dtype=torch.int64, device=dev))
        seal = torch.where(torch.rand((n, 52), device=dev, generator=g) < 0.1,
                           torch.randint(1, 5, (n, 52), device=dev, generator=g), torch.zeros((n, 52), dtype=torch.int64, device=dev))
        deck_view = self.state_field("deck")                     # int16 view of the 52 card16 entries
        deck = deck_view.to(torch.int64) & 63
        deck = deck | (enh << 6) | (ed << 10) | (seal << 13)
        deck_view.copy_(torch.where(deck >= 2 ** 15, deck - 2 ** 16, deck).to(torch.int16))
        self.state_written()


class HostMirror:
    """Pinned-host copy of a slab's step results, kept current by OBSERVATION DELTAS instead of whole-array copies.

    What a host-driven loop over `BalatroVecEnv` needs per step is the reference's step() result on the host:
    observation, reward, terminated.  A step rewrites the 176-byte observation record only of the envs whose action was
    not a card toggle (include/bgym.h, BgymSel), so per step this class moves
        host -> device   actions                                                                4 B / env
        device -> host   selection (8 selected_cards flags as ONE byte, mask word), reward, terminated  1 + 8 + 8 + 1 B / env
        device -> host   the rewritten observation records, packed on the device (bgym_pack_dirty_obs) and written
                         into the pinned mirror by the GPU itself (bgym_scatter_dirty_obs, zero-copy stores, one
                         aligned 128-byte line per record + 32 B for envs in or entering the shop)   128 (+32) B / changed env
    against 189 B / env for whole arrays.  The device->host traffic of step t runs on a second stream while step t+1
    is launched (double-buffered snapshots); `wait()` returns when the mirror holds the last step's results.

        mirror = HostMirror(env); env.reset(); mirror.pull_all()
        mirror.actions[:] = ...                      # host policy writes this step's actions
        mirror.step()                                # H2D, step, D2H deltas (asynchronous)
        mirror.wait(); mirror.core / .shop / .sel_bits / .mask_words / .reward / .terminated are current

    Layout of the mirror (include/bgym.h): `core` [n, 128] = chunks 0..5, 8, 9 of the observation record, `shop` [n, 32]
    = chunks 6, 7; the selection record travels packed — `sel_bits` [n] uint8 (bit i = selected_cards[i]) and `mask_words`
    [n] int64 (the legal-action word) — 9 B instead of 16 B per env and step: like `action_mask`, `selected_cards` is
    expanded on access (`field('selected_cards')`, `sel`).  `obs_records()` reassembles whole L.OBS_DTYPE records (a
    host-side copy); `field(name)` gives a zero-copy view of one observation field where the mirror stores it as such."""

    def __init__(self, env: "BalatroVecEnv"):
        torch = env.torch
        self.env, self.torch = env, torch
        n, dev = env.num_envs, env.device
        pin = dict(pin_memory=True)
        self.actions = torch.zeros(n, dtype=torch.int32, **pin)
        self.core = torch.zeros((n, L.MIRROR_CORE_BYTES), dtype=torch.uint8, **pin)
        self.shop = torch.zeros((n, L.MIRROR_SHOP_BYTES), dtype=torch.uint8, **pin)
        self.sel_bits = torch.zeros(n, dtype=torch.uint8, **pin)
        self.mask_words = torch.zeros(n, dtype=torch.int64, **pin)
        self.reward = torch.zeros(n, dtype=torch.float64, **pin)
        self.terminated = torch.zeros(n, dtype=torch.uint8, **pin)
        self._d_act = torch.zeros(n, dtype=torch.int32, device=dev)
        if env.obs_dirty is None:
            env.obs_dirty = torch.zeros(n, dtype=torch.uint8, device=dev)     # from now on every step flags the records it rewrites
        self._scratch = torch.zeros((n + 4095) // 4096 + 4, dtype=torch.int32, device=dev)
        self.staging_bytes = 16 + ((n * 4 + 15) & ~15) + n * L.OBS_DELTA_BYTES
        self._snap = [{"staging": torch.zeros(self.staging_bytes, dtype=torch.uint8, device=dev),
                       "selbits": torch.empty(n, dtype=torch.uint8, device=dev), "maskw": torch.empty(n, dtype=torch.int64, device=dev),
                       "rew": torch.empty_like(env.reward), "term": torch.empty_like(env.terminated)} for _ in range(2)]
        self._prod = torch.empty(n, dtype=torch.int64, device=dev)
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._ready = [torch.cuda.Event() for _ in range(2)]
        self._copied = [torch.cuda.Event() for _ in range(2)]
        self._t = 0
        self.h2d_bytes_per_step = 4 * n
        self.dense_d2h_bytes_per_step = (1 + 8 + 8 + 1) * n

    # eight 0/1 bytes b0..b7 of a little-endian int64 -> the byte sum(b_i << i): the product with this constant puts b_i at
    # bit 56 + i (no two partial products share a bit, so nothing carries), i.e. the answer is the product's top byte
    _GATHER_BITS = 0x0102040810204080

    def _pack_sel(self, snap):
        """Selection records (16 B: selected_cards[8] | mask word) -> one byte of flags + the word, on the current stream."""
        n = self.env.num_envs
        selv = self.env.sel.view(self.torch.int64)              # [n, 2]
        self.torch.mul(selv[:, 0], self._GATHER_BITS, out=self._prod)
        snap["selbits"].copy_(self._prod.view(self.torch.uint8).view(n, 8)[:, 7], non_blocking=True)
        snap["maskw"].copy_(selv[:, 1], non_blocking=True)

    @property
    def sel(self) -> np.ndarray:
        """The selection records as the device keeps them (L.SEL_DTYPE), expanded from the packed mirror: a host-side copy."""
        out = np.zeros(self.env.num_envs, dtype=L.SEL_DTYPE)
        out["selected_cards"] = np.unpackbits(self.sel_bits.numpy()[:, None], axis=1, bitorder="little")
        out["action_mask_bits"] = self.mask_words.numpy().view(out["action_mask_bits"].dtype)
        return out

    def _pack(self, snap, stream, everything: bool):
        env = self.env
        with self.torch.cuda.device(env.device):
            rc = env.lib.bgym_pack_dirty_obs(env._obs.data_ptr(), None if everything else env.obs_dirty.data_ptr(),
                                             snap["staging"].data_ptr(), self._scratch.data_ptr(), env.num_envs, env.num_envs,
                                             stream.cuda_stream)
        _lib.check(rc, "bgym_pack_dirty_obs")

    def _scatter(self, snap, stream):
        env = self.env
        with self.torch.cuda.device(env.device):
            rc = env.lib.bgym_scatter_dirty_obs(snap["staging"].data_ptr(), env.num_envs, self.core.data_ptr(), self.shop.data_ptr(),
                                                stream.cuda_stream)
        _lib.check(rc, "bgym_scatter_dirty_obs")

    def pull_all(self):
        """Fill the whole mirror from the device arrays (after reset / state injection)."""
        env, torch = self.env, self.torch
        main = torch.cuda.current_stream(env.device)
        self._copy_stream.synchronize()
        self._pack(self._snap[0], main, True)
        self._scatter(self._snap[0], main)
        env.obs_dirty.zero_()
        self._pack_sel(self._snap[0])
        self.sel_bits.copy_(self._snap[0]["selbits"], non_blocking=True)
        self.mask_words.copy_(self._snap[0]["maskw"], non_blocking=True)
        self.reward.copy_(env.reward, non_blocking=True)
        self.terminated.copy_(env.terminated, non_blocking=True)
        main.synchronize()

    def step(self, want_info: bool = False):
        env, torch = self.env, self.torch
        b = self._t & 1
        self._t += 1
        snap = self._snap[b]
        main = torch.cuda.current_stream(env.device)
        self._d_act.copy_(self.actions, non_blocking=True)
        env.step(self._d_act, want_info=want_info)
        main.wait_event(self._copied[b])                       # the snapshot's previous contents have left the device
        self._pack(snap, main, False)
        self._pack_sel(snap)
        snap["rew"].copy_(env.reward, non_blocking=True)
        snap["term"].copy_(env.terminated, non_blocking=True)
        self._ready[b].record(main)
        cs = self._copy_stream
        with torch.cuda.stream(cs):
            cs.wait_event(self._ready[b])
            self.sel_bits.copy_(snap["selbits"], non_blocking=True)
            self.mask_words.copy_(snap["maskw"], non_blocking=True)
            self.reward.copy_(snap["rew"], non_blocking=True)
            self.terminated.copy_(snap["term"], non_blocking=True)
            self._scatter(snap, cs)
            self._copied[b].record(cs)

    def wait(self):
        self._copy_stream.synchronize()
        self.torch.cuda.current_stream(self.env.device).synchronize()

    def delta_counts(self, which: int = None):
        """(records, records with their shop chunks) the given (default: last) step's delta carried."""
        b = ((self._t - 1) & 1) if which is None else which
        st = self._snap[b]["staging"]
        cnt = int(st[:4].view(self.torch.int32).item())
        flagged = int((st[16:16 + 4 * cnt].view(self.torch.int32) < 0).sum().item()) if cnt > 0 else 0
        return cnt, flagged

    def dirty_count(self, which: int = None):
        return self.delta_counts(which)[0]

    def obs_records(self) -> np.ndarray:
        """The mirror as whole observation records (L.OBS_DTYPE), a host-side copy."""
        n = self.env.num_envs
        core = self.core.numpy().reshape(n, 8, 16)
        shop = self.shop.numpy().reshape(n, 2, 16)
        raw = np.zeros((n, L.OBS_BYTES // 16, 16), dtype=np.uint8)
        raw[:, list(L.MIRROR_CORE_CHUNKS)] = core
        raw[:, list(L.MIRROR_SHOP_CHUNKS)] = shop
        rec = raw.reshape(n, L.OBS_BYTES).reshape(-1).view(L.OBS_DTYPE)
        s = self.sel
        rec["selected_cards"] = s["selected_cards"]
        rec["action_mask_bits"] = s["action_mask_bits"]
        return rec

    def field(self, name):
        """numpy view of one observation field in the mirror: zero-copy for every field but shop_items / shop_costs (they
        straddle the core / shop split: assembled copy), 'action_mask' (expanded from the mask word) and 'selected_cards'
        (expanded from the flag byte)."""
        if name == "action_mask":
            return L.mask_from_bits(self.field("action_mask_bits"))
        if name == "action_mask_bits":
            return self.mask_words.numpy().view(L.SEL_DTYPE.fields["action_mask_bits"][0])
        if name == "selected_cards":
            return np.unpackbits(self.sel_bits.numpy()[:, None], axis=1, bitorder="little").view(L.SEL_DTYPE.fields["selected_cards"][0].base)
        if name in L.MIRROR_CORE_DTYPE.names:
            return self.core.numpy().reshape(-1).view(L.MIRROR_CORE_DTYPE)[name]
        return self.obs_records()[name]
