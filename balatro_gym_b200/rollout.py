"""On-device PPO rollout collection (BASELINE config 5).

What the reference does per rollout (train_balatro_agent.py:269-475 through stable_baselines3):
N subprocess envs step on CPU, observations are pickled to the trainer, `BalatroFeaturesExtractor`
(:42-119) one-hots the hand on the GPU, PPO samples an action per env, and after `n_steps` the
RolloutBuffer computes GAE(gamma=0.99, lambda=0.95) (:328-336).

Here the whole loop stays on the device: the env step kernels write observation records, a
featurize kernel turns them into the extractor's dense input, the policy MLP runs under bf16
autocast (cuBLAS — policy side, not part of the env path), a masked-categorical kernel samples
actions from the logits and the observation's legal-action word, and a GAE kernel closes the
rollout.  Nothing crosses PCIe.  Multi-GPU: one collector per rank over its env slab
(`BalatroVecEnv(env_offset=...)`); the only collective is the gradient all-reduce of the PPO
update (`ppo_update`), which is policy side.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import layout as L
from . import _lib

FEATURE_DIM = 448
_DT = {"float32": 0, "bfloat16": 1}


def _torch():
    return _lib.require_cuda()


def featurize(obs_records, out=None, dtype=None):
    """[n, 176] uint8 observation records -> [n, 448] features (bgym_featurize)."""
    torch = _torch()
    lib = _lib.load()
    n = obs_records.shape[0]
    assert obs_records.dtype == torch.uint8 and obs_records.shape[1] == L.OBS_BYTES and obs_records.is_contiguous()
    if out is None:
        out = torch.empty((n, FEATURE_DIM), dtype=dtype or torch.float32, device=obs_records.device)
    code = _DT[str(out.dtype).split(".")[-1]]
    with torch.cuda.device(obs_records.device):
        rc = lib.bgym_featurize(obs_records.data_ptr(), out.data_ptr(), n, code, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "bgym_featurize")
    return out


def policy_first_layer(obs_records, wt_hand, wt_joker, wt_game, bias, out=None):
    """[n, 176] observation records -> [n, 448] bf16 = relu of the first Linear of hand_net / joker_net /
    game_state_net (bgym_policy_first_layer): the hand block is a sum of eight weight rows, no one-hot is built.
    wt_* are the TRANSPOSED bf16 weights ([in, out], contiguous), bias the three biases back to back (fp32 [448])."""
    torch = _torch()
    lib = _lib.load()
    n = obs_records.shape[0]
    assert obs_records.dtype == torch.uint8 and obs_records.shape[1] == L.OBS_BYTES and obs_records.is_contiguous()
    assert wt_hand.shape == (416, 256) and wt_joker.shape == (10, 128) and wt_game.shape == (21, 64) and bias.shape == (448,)
    for w in (wt_hand, wt_joker, wt_game):
        assert w.dtype == torch.bfloat16 and w.is_contiguous()
    assert bias.dtype == torch.float32 and bias.is_contiguous()
    if out is None:
        out = torch.empty((n, 448), dtype=torch.bfloat16, device=obs_records.device)
    with torch.cuda.device(obs_records.device):
        rc = lib.bgym_policy_first_layer(obs_records.data_ptr(), wt_hand.data_ptr(), wt_joker.data_ptr(), wt_game.data_ptr(),
                                         bias.data_ptr(), out.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "bgym_policy_first_layer")
    return out


# ---- fused tcgen05 forward of layers 2.. (libbgym_policy.so, include/bgym_policy.h) -------------------------------
_POLICY_LAYER_KEYS = ["hand_net.0", "joker_net.0", "game_state_net.0", "hand_net.2", "joker_net.2", "game_state_net.2",
                      "combined_net.0", "combined_net.2", "pi.0", "vf.0", "pi.2", "vf.2", "pi.4", "vf.4"]
# input-column shift of a layer inside its K-block (the game scalars sit at columns 16..36 of the K-block they share with
# the joker ids), and (epilogue group, first accumulator column) of every layer: bgym_policy.cu::LAYERS
_POLICY_K_SHIFT = {2: 16}
_POLICY_COL_OF = {0: (0, 0), 1: (0, 256), 2: (0, 384), 3: (1, 0), 4: (1, 128), 5: (1, 192), 6: (2, 0), 7: (3, 0),
                  8: (4, 0), 9: (4, 256), 10: (5, 0), 11: (5, 256), 12: (6, 0), 13: (6, 64)}


def policy_program():
    """(steps as a numpy structured array, weight blob bytes, bias blob floats) of bgym_policy_program."""
    import ctypes as C
    lib = _lib.load_policy()
    step_dt = np.dtype([(k, "<i4") for k in ("offset", "bytes", "layer", "n0", "n", "kb", "a_kb", "col", "first", "last", "group", "_pad")])
    steps = np.zeros(80, dtype=step_dt)
    wb, bf = C.c_int64(0), C.c_int64(0)
    n = lib.bgym_policy_program(steps.ctypes.data, C.addressof(wb), C.addressof(bf))
    return steps[:n], int(wb.value), int(bf.value)


_PACK_INDEX = {}


def _pack_index(shapes):
    """Gather index of the packed weight blob (one int64 per bf16 element): position in the concatenation of the eleven
    flattened [out, in] weight matrices (+ one trailing zero element that padding points at), built once per set of layer
    shapes.  Swizzle: 16-byte chunk c of row r is stored at chunk c ^ (r & 7)."""
    key = tuple(shapes)
    if key in _PACK_INDEX:
        return _PACK_INDEX[key]
    steps, wbytes, bfloats = policy_program()
    base = np.concatenate([[0], np.cumsum([a * b for a, b in shapes])]).astype(np.int64)
    zero_pos = int(base[-1])
    index = np.full(wbytes // 2, zero_pos, dtype=np.int64)
    for st in steps:
        li, n0, n, kb = int(st["layer"]), int(st["n0"]), int(st["n"]), int(st["kb"])
        rows_total, cols_total = shapes[li]
        r = np.arange(n)[:, None, None]                   # tile row
        c = np.arange(8)[None, :, None]                   # 16-byte chunk of the row (8 bf16)
        e = np.arange(8)[None, None, :]                   # element of the chunk
        src_row, src_col = n0 + r, 64 * kb + 8 * c + e - _POLICY_K_SHIFT.get(li, 0)
        ok = (src_row < rows_total) & (src_col >= 0) & (src_col < cols_total)
        src = np.where(ok, base[li] + src_row * cols_total + src_col, zero_pos)
        dst = int(st["offset"]) // 2 + r * 64 + ((c ^ (r & 7)) * 8) + e
        index[np.broadcast_to(dst, src.shape).reshape(-1)] = src.reshape(-1)
    bias_at = [512 * _POLICY_COL_OF[li][0] + _POLICY_COL_OF[li][1] for li in range(len(shapes))]
    _PACK_INDEX[key] = (index, bias_at, bfloats)
    return _PACK_INDEX[key]


def pack_policy_weights(state_dict, device):
    """Pack the fourteen Linear layers into the blobs bgym_policy_forward reads (on the device: one gather): every program step's weight tile (rows [n0, n0 + n) x input columns [64 kb, 64 kb + 64), zero-padded) as n
    rows of 128 bytes of bf16, 128-byte swizzled; biases per epilogue group at the accumulator column they are added to."""
    torch = _torch()
    ws = [state_dict[k + ".weight"].detach() for k in _POLICY_LAYER_KEYS]
    index, bias_at, bfloats = _pack_index([tuple(w.shape) for w in ws])
    cache = _PACK_INDEX.setdefault(("dev", str(device)), {})
    if "index" not in cache:
        cache["index"] = torch.from_numpy(index).to(device)
    flat = torch.cat([w.to(device=device, dtype=torch.bfloat16).reshape(-1) for w in ws] + [torch.zeros(1, dtype=torch.bfloat16, device=device)])
    blob = flat[cache["index"]].contiguous()
    bias = torch.zeros(bfloats, dtype=torch.float32, device=device)
    for li, key in enumerate(_POLICY_LAYER_KEYS):
        b = state_dict[key + ".bias"].detach().to(device=device, dtype=torch.float32)
        bias[bias_at[li]:bias_at[li] + b.shape[0]] = b
    return blob, bias


def policy_forward_fused(obs_records, weights, bias, logits=None, value=None):
    """[n, 176] observation records -> (logits fp32 [n, 60], value fp32 [n]): the whole policy in one tcgen05 kernel
    (bgym_policy_forward).  Reads hand, joker_ids and the game scalars of the records: fields a step keeps current."""
    torch = _torch()
    lib = _lib.load_policy()
    n = obs_records.shape[0]
    assert obs_records.dtype == torch.uint8 and obs_records.shape[1] == L.OBS_BYTES and obs_records.is_contiguous()
    dev = obs_records.device
    logits = torch.empty((n, L.NUM_ACTIONS), dtype=torch.float32, device=dev) if logits is None else logits
    value = torch.empty(n, dtype=torch.float32, device=dev) if value is None else value
    with torch.cuda.device(dev):
        rc = lib.bgym_policy_forward(obs_records.data_ptr(), weights.data_ptr(), bias.data_ptr(), logits.data_ptr(), value.data_ptr(), n,
                                     torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise _lib.BgymError(f"bgym_policy_forward failed (rc={rc}): {lib.bgym_policy_last_error().decode()}")
    return logits, value


def masked_sample(logits, obs_records, seed: int = 0, step: int = 0, env_offset: int = 0, uniforms=None,
                  actions=None, logp=None, entropy=None):
    """Sample one legal action per env from softmax(logits | legal) (bgym_masked_sample).
    obs_records: [n, 176] WHOLE observation records (the mask word is read at byte 160) or [n, 16] selection records
    (BgymSel, byte 8) — the array a step keeps current.  Returns (actions int32 [n], logp float32 [n], entropy float32 [n])."""
    torch = _torch()
    lib = _lib.load()
    n = logits.shape[0]
    assert logits.shape[1] == L.NUM_ACTIONS and logits.is_contiguous()
    code = _DT[str(logits.dtype).split(".")[-1]]
    dev = logits.device
    actions = torch.empty(n, dtype=torch.int32, device=dev) if actions is None else actions
    logp = torch.empty(n, dtype=torch.float32, device=dev) if logp is None else logp
    entropy = torch.empty(n, dtype=torch.float32, device=dev) if entropy is None else entropy
    assert obs_records.dtype == torch.uint8 and obs_records.is_contiguous() and obs_records.shape[1] in (L.OBS_BYTES, L.SEL_BYTES)
    mask_off = 160 if obs_records.shape[1] == L.OBS_BYTES else 8
    with torch.cuda.device(dev):
        rc = lib.bgym_masked_sample(logits.data_ptr(), code, obs_records.data_ptr() + mask_off, obs_records.shape[1],
                                    None if uniforms is None else uniforms.data_ptr(), seed & 0xFFFFFFFF, step, env_offset,
                                    actions.data_ptr(), logp.data_ptr(), entropy.data_ptr(), n,
                                    torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "bgym_masked_sample")
    return actions, logp, entropy


def gae(rewards, values, dones, gamma: float = 0.99, lam: float = 0.95, advantages=None, returns=None):
    """rewards [T, n] f32, values [T+1, n] f32, dones [T, n] u8 -> (advantages, returns) [T, n] (bgym_gae)."""
    torch = _torch()
    lib = _lib.load()
    T, n = rewards.shape
    assert values.shape == (T + 1, n) and dones.shape == (T, n)
    assert rewards.dtype == torch.float32 and values.dtype == torch.float32 and dones.dtype == torch.uint8
    assert rewards.is_contiguous() and values.is_contiguous() and dones.is_contiguous()
    advantages = torch.empty_like(rewards) if advantages is None else advantages
    returns = torch.empty_like(rewards) if returns is None else returns
    with torch.cuda.device(rewards.device):
        rc = lib.bgym_gae(rewards.data_ptr(), values.data_ptr(), dones.data_ptr(), gamma, lam,
                          advantages.data_ptr(), returns.data_ptr(), T, n, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "bgym_gae")
    return advantages, returns


def make_policy(features_dim: int = 512, pi=(256, 256), vf=(256, 256), device="cuda", seed: int = 0):
    """Actor-critic with the reference's extractor topology (train_balatro_agent.py:48-81, heads :341).
    The reference declares joker_dim = 160 and game_state_dim = 32 but feeds 10 and 21 columns
    (:98, :102-113) — it cannot run as written; the widths here are the ones the data has."""
    torch = _torch()
    nn = torch.nn

    class BalatroPolicy(nn.Module):
        def __init__(self):
            super().__init__()
            self.hand_net = nn.Sequential(nn.Linear(416, 256), nn.ReLU(), nn.Linear(256, 128), nn.ReLU())
            self.joker_net = nn.Sequential(nn.Linear(10, 128), nn.ReLU(), nn.Linear(128, 64), nn.ReLU())
            self.game_state_net = nn.Sequential(nn.Linear(21, 64), nn.ReLU(), nn.Linear(64, 32), nn.ReLU())
            self.combined_net = nn.Sequential(nn.Linear(224, features_dim), nn.ReLU(),
                                              nn.Linear(features_dim, features_dim), nn.ReLU())

            def head(widths, out):
                layers, d = [], features_dim
                for w in widths:
                    layers += [nn.Linear(d, w), nn.Tanh()]
                    d = w
                return nn.Sequential(*layers, nn.Linear(d, out))
            self.pi = head(pi, L.NUM_ACTIONS)
            self.vf = head(vf, 1)

        def forward(self, feats):
            h = self.hand_net(feats[:, :416])
            j = self.joker_net(feats[:, 416:426])
            g = self.game_state_net(feats[:, 426:447])
            z = self.combined_net(torch.cat([h, j, g], dim=1))
            return self.pi(z), self.vf(z).squeeze(-1)

    # same seed -> same initial weights on every rank, without disturbing the caller's RNG stream
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        policy = BalatroPolicy()
    finally:
        torch.random.set_rng_state(state)
    return policy.to(device)


class RolloutCollector:
    """`n_steps` of experience for every env of a slab, collected without leaving the device.

    Buffers are time-major [T(+1), n, ...]: obs (raw 176-byte records; features are recomputed
    on demand, 5x smaller than storing bf16 features), actions, logp, values, rewards, dones,
    advantages, returns.
    """

    def __init__(self, vec, policy, n_steps: int = 128, gamma: float = 0.99, gae_lambda: float = 0.95,
                 seed: int = 0, autocast: bool = True, fused: bool = True):
        torch = vec.torch
        self.torch = torch
        self.vec, self.policy = vec, policy
        self.T, self.n = int(n_steps), vec.num_envs
        self.gamma, self.lam, self.seed = gamma, gae_lambda, seed
        self.autocast = autocast
        # fused = layers 2.. of the rollout forward in one tcgen05 kernel (libbgym_policy.so) instead of eleven library GEMMs
        self.fused = bool(fused and autocast)
        dev, T, n = vec.device, self.T, self.n
        self.obs = torch.empty((T + 1, n, L.OBS_BYTES), dtype=torch.uint8, device=dev)
        self.actions = torch.empty((T, n), dtype=torch.int32, device=dev)
        self.logp = torch.empty((T, n), dtype=torch.float32, device=dev)
        self.entropy = torch.empty((T, n), dtype=torch.float32, device=dev)
        self.values = torch.empty((T + 1, n), dtype=torch.float32, device=dev)
        self.rewards = torch.empty((T, n), dtype=torch.float32, device=dev)
        self.dones = torch.empty((T, n), dtype=torch.uint8, device=dev)
        self.advantages = torch.empty((T, n), dtype=torch.float32, device=dev)
        self.returns = torch.empty((T, n), dtype=torch.float32, device=dev)
        self._feats = torch.empty((n, FEATURE_DIM), dtype=torch.bfloat16 if autocast else torch.float32, device=dev)
        self._logits = torch.empty((n, L.NUM_ACTIONS), dtype=torch.float32, device=dev)
        self._value = torch.empty(n, dtype=torch.float32, device=dev)
        self.global_step = 0

    def refresh_inference_weights(self):
        """bf16 copies of the policy's weights for the no-grad rollout forward (call after each update;
        `collect` does it on entry)."""
        torch = self.torch
        with torch.no_grad():
            sd = self.policy.state_dict()
            self._w16 = {k: v.detach().to(torch.bfloat16).contiguous() for k, v in sd.items()}
            # first layer of the three sub-nets: transposed weights for the row-sum kernel
            self._first = (self._w16["hand_net.0.weight"].t().contiguous(), self._w16["joker_net.0.weight"].t().contiguous(),
                           self._w16["game_state_net.0.weight"].t().contiguous(),
                           torch.cat([sd["hand_net.0.bias"], sd["joker_net.0.bias"], sd["game_state_net.0.bias"]]).detach().float().contiguous())
            if self.fused:
                self._packed = pack_policy_weights(sd, self.vec.device)

    def _mlp(self, x, prefix, n_layers, last_plain=False, act="relu"):
        """Sequential of Linear(+activation) from the cached bf16 weights; ReLU layers use the cuBLASLt
        bias+ReLU epilogue (`torch._addmm_activation`), so an activation never makes its own HBM round trip."""
        torch = self.torch
        w = self._w16
        for li in range(n_layers):
            W, b = w[f"{prefix}.{2 * li}.weight"], w[f"{prefix}.{2 * li}.bias"]
            last = li == n_layers - 1
            if last and last_plain:
                x = torch.addmm(b, x, W.t())
            elif act == "relu":
                x = torch._addmm_activation(b, x, W.t(), use_gelu=False)
            else:
                x = torch.tanh_(torch.addmm(b, x, W.t()))
        return x

    def _mlp_from(self, x, prefix, first, n_layers):
        """layers first..n_layers-1 of a ReLU Sequential (the first one was done by policy_first_layer)"""
        torch = self.torch
        for li in range(first, n_layers):
            W, b = self._w16[f"{prefix}.{2 * li}.weight"], self._w16[f"{prefix}.{2 * li}.bias"]
            x = torch._addmm_activation(b, x, W.t(), use_gelu=False)
        return x

    def _forward(self, obs_records):
        torch = self.torch
        if not self.autocast:
            featurize(obs_records, out=self._feats)
            logits, value = self.policy(self._feats)
            return logits.float().contiguous(), value.float()
        # first Linear + ReLU of the three sub-nets straight from the records (no one-hot feature matrix), then their
        # second layers on strided column views of that activation (lda = 448)
        if self.fused:      # the whole policy in one tcgen05 kernel, straight from the records: activations never leave the SM
            return policy_forward_fused(obs_records, *self._packed, logits=self._logits, value=self._value)
        a = policy_first_layer(obs_records, *self._first, out=self._feats)
        h = self._mlp_from(a[:, :256], "hand_net", 1, 2)
        j = self._mlp_from(a[:, 256:384], "joker_net", 1, 2)
        g = self._mlp_from(a[:, 384:448], "game_state_net", 1, 2)
        z = self._mlp(torch.cat([h, j, g], dim=1), "combined_net", 2)
        logits = self._mlp(z, "pi", 3, last_plain=True, act="tanh")
        value = self._mlp(z, "vf", 3, last_plain=True, act="tanh").squeeze(-1)
        return logits.float().contiguous(), value.float()

    def collect(self):
        """Run T steps from the vec env's CURRENT observations (call vec.reset() once before the first)."""
        torch = self.torch
        vec = self.vec
        with torch.no_grad():
            if self.autocast:
                self.refresh_inference_weights()
            self.obs[0].copy_(vec.obs_buf)
            for t in range(self.T):
                logits, value = self._forward(self.obs[t])
                masked_sample(logits, self.obs[t], seed=self.seed, step=self.global_step, env_offset=vec.env_offset,
                              actions=self.actions[t], logp=self.logp[t], entropy=self.entropy[t])
                self.values[t].copy_(value)
                vec.step(self.actions[t], want_info=False)
                self.rewards[t].copy_(vec.reward)
                self.dones[t].copy_(vec.terminated)
                self.obs[t + 1].copy_(vec.obs_buf)
                self.global_step += 1
            _, value = self._forward(self.obs[self.T])
            self.values[self.T].copy_(value)
            gae(self.rewards, self.values, self.dones, self.gamma, self.lam, self.advantages, self.returns)
        return self

    def stats(self):
        """(env-steps, episodes finished, mean reward per step) of the last rollout, as python numbers."""
        return self.T * self.n, int(self.dones.sum().item()), float(self.rewards.mean().item())


def legal_mask(obs_records):
    """[B, 176] uint8 records -> [B, 60] bool from the packed legal-action word."""
    torch = _torch()
    bits = obs_records[:, 160:168].contiguous().view(torch.int64)
    return ((bits >> torch.arange(L.NUM_ACTIONS, device=obs_records.device)) & 1).bool()


def evaluate_actions(policy, obs_records, actions, autocast: bool = True):
    """log-prob, entropy, value of `actions` under the masked policy — the differentiable twin of
    the sampling kernel (torch ops, used by the PPO update)."""
    torch = _torch()
    feats = featurize(obs_records, dtype=torch.bfloat16 if autocast else torch.float32)
    if autocast:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits, value = policy(feats)
    else:
        logits, value = policy(feats)
    mask = legal_mask(obs_records)
    mask[:, 0] |= ~mask.any(dim=1)
    logits = logits.float().masked_fill(~mask, float("-inf"))
    logp_all = torch.log_softmax(logits, dim=1)
    logp = logp_all.gather(1, actions.long().unsqueeze(1)).squeeze(1)
    p = logp_all.exp()
    entropy = -(p * logp_all.masked_fill(~mask, 0.0)).sum(dim=1)
    return logp, entropy, value.float()


def ppo_update(policy, optimizer, rollout: RolloutCollector, n_epochs: int = 4, minibatch: int = 1 << 16,
               clip_range: float = 0.2, ent_coef: float = 0.01, vf_coef: float = 0.5, max_grad_norm: float = 0.5,
               generator=None):
    """Clipped-surrogate PPO over one rollout (the SB3 defaults the reference uses,
    train_balatro_agent.py:328-336).  With torch.distributed initialised, gradients are averaged
    over ranks with one all-reduce per minibatch on a flat buffer."""
    torch = _torch()
    from . import dist as bdist
    T, n = rollout.T, rollout.n
    N = T * n
    obs = rollout.obs[:T].view(N, L.OBS_BYTES)
    actions, old_logp = rollout.actions.view(N), rollout.logp.view(N)
    adv_all, ret_all = rollout.advantages.view(N), rollout.returns.view(N)
    params = [p for p in policy.parameters() if p.requires_grad]
    last = {}
    for _ in range(n_epochs):
        perm = torch.randperm(N, device=obs.device, generator=generator)
        for s in range(0, N, minibatch):
            idx = perm[s:s + minibatch]
            logp, entropy, value = evaluate_actions(policy, obs[idx], actions[idx], autocast=rollout.autocast)
            adv = adv_all[idx]
            adv = (adv - adv.mean()) / (adv.std() + 1e-8)
            ratio = torch.exp(logp - old_logp[idx])
            pg_loss = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1 - clip_range, 1 + clip_range)).mean()
            v_loss = torch.nn.functional.mse_loss(value, ret_all[idx])
            ent = entropy.mean()
            loss = pg_loss + vf_coef * v_loss - ent_coef * ent
            optimizer.zero_grad(set_to_none=True)
            loss.backward()
            bdist.allreduce_mean_grads(params)
            torch.nn.utils.clip_grad_norm_(params, max_grad_norm)
            optimizer.step()
            last = {"loss": loss.detach(), "pg_loss": pg_loss.detach(), "v_loss": v_loss.detach(), "entropy": ent.detach(),
                    "approx_kl": ((ratio - 1) - torch.log(ratio)).mean().detach()}
    return {k: float(v.item()) for k, v in last.items()}
