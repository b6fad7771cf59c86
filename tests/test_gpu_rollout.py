"""GPU tests of the rollout-collection kernels: each kernel against a plain PyTorch
fp32 statement of the same op (these are the floating-point kernels of the repo; tolerances are
written at each comparison), then the collector end to end."""
import numpy as np
import pytest

from balatro_gym_b200 import layout as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.fixture(scope="module")
def mixed_obs(torch):
    """Observation records of a slab spread over phases, antes and shop states."""
    from balatro_gym_b200 import BalatroVecEnv
    v = BalatroVecEnv(1 << 14, seed=5)
    v.reset()
    v.randomize_c3(5)
    for _ in range(150):
        v.step(random_policy=True)
    return v


def reference_features(torch, obs):
    """BalatroFeaturesExtractor.forward's preprocessing (train_balatro_agent.py:84-113) in plain torch fp32."""
    hand = obs["hand"].long()
    B = hand.shape[0]
    one_hot = torch.zeros(B, 8, 52, device=hand.device)
    for i in range(8):
        valid = hand[:, i] >= 0
        if valid.any():
            one_hot[valid, i, hand[valid, i]] = 1
    game = torch.cat([
        obs["chips_scored"].float().unsqueeze(1) / 1e6, obs["chips_needed"].float().unsqueeze(1) / 1e5,
        obs["progress_ratio"].float().unsqueeze(1), obs["money"].float().unsqueeze(1) / 100,
        obs["ante"].float().unsqueeze(1) / 10, obs["round"].float().unsqueeze(1) / 3,
        obs["hands_left"].float().unsqueeze(1) / 10, obs["discards_left"].float().unsqueeze(1) / 5,
        obs["hand_levels"].float() / 10, obs["phase"].float().unsqueeze(1) / 3], dim=1)
    return torch.cat([one_hot.view(B, -1), obs["joker_ids"].float().view(B, -1), game,
                      torch.zeros(B, 1, device=hand.device)], dim=1)


def test_featurize_matches_the_extractor_preprocessing(torch, mixed_obs):
    from balatro_gym_b200.rollout import featurize, FEATURE_DIM
    v = mixed_obs
    ref = reference_features(torch, v.obs)
    assert ref.shape == (v.num_envs, FEATURE_DIM)
    out = featurize(v.obs_buf)
    # one-hot and joker-id columns are exact; the scaled scalars are IEEE fp32 divisions here while
    # torch multiplies by the reciprocal of a scalar divisor on CUDA: tolerance 1 ulp (rtol 1.2e-7)
    assert torch.equal(out[:, :426], ref[:, :426])
    assert torch.allclose(out[:, 426:], ref[:, 426:], rtol=1.2e-7, atol=0)
    out16 = featurize(v.obs_buf, dtype=torch.bfloat16)
    assert torch.equal(out16[:, :416], ref[:, :416].to(torch.bfloat16))
    assert torch.allclose(out16.float(), ref, rtol=2 ** -8, atol=0)   # bf16: one rounding of the fp32 value
    assert int((ref[:, :416].sum(dim=1) != (v.obs["hand"] >= 0).sum(dim=1)).sum()) == 0
    assert len(torch.unique(v.obs["phase"])) >= 3                # the slab really is spread over phases


def test_policy_first_layer_matches_torch(torch, mixed_obs):
    """bgym_policy_first_layer (sum of eight weight rows for the 8-hot hand block, no feature matrix) against the plain
    torch statement: relu(Linear(features)) of the three sub-nets' first layers, fp32 reference on the bf16-rounded
    operands the kernel uses.  Tolerance: fp32 accumulation in another order + one bf16 rounding of the output."""
    from balatro_gym_b200.rollout import policy_first_layer, make_policy
    v = mixed_obs
    pol = make_policy(seed=3)
    sd = pol.state_dict()
    bf = lambda t: t.detach().to(torch.bfloat16)
    wt = (bf(sd["hand_net.0.weight"]).t().contiguous(), bf(sd["joker_net.0.weight"]).t().contiguous(), bf(sd["game_state_net.0.weight"]).t().contiguous())
    bias = torch.cat([sd["hand_net.0.bias"], sd["joker_net.0.bias"], sd["game_state_net.0.bias"]]).float().contiguous()
    out = policy_first_layer(v.obs_buf, *wt, bias)
    feats = reference_features(torch, v.obs).to(torch.bfloat16).float()            # the operands of the bf16 GEMM path
    ref = torch.cat([torch.relu(feats[:, :416] @ wt[0].float() + bias[:256]),
                     torch.relu(feats[:, 416:426] @ wt[1].float() + bias[256:384]),
                     torch.relu(feats[:, 426:447] @ wt[2].float() + bias[384:448])], dim=1)
    assert out.shape == (v.num_envs, 448) and out.dtype == torch.bfloat16
    err = (out.float() - ref).abs()
    assert float((err - 2 ** -8 * ref.abs()).max()) < 1e-3, float(err.max())       # bf16 output rounding + reordering
    assert float(ref.abs().max()) > 1.0                                              # the joker block is not tiny: ids up to 150


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_masked_sample_matches_torch(torch, mixed_obs, dtype):
    from balatro_gym_b200.rollout import masked_sample, legal_mask
    v = mixed_obs
    n = v.num_envs
    g = torch.Generator(device="cuda").manual_seed(1)
    logits = (torch.randn((n, 60), device="cuda", generator=g) * 3).to(getattr(torch, dtype)).contiguous()
    u = torch.rand(n, device="cuda", generator=g)
    a, lp, ent = masked_sample(logits, v.obs_buf, uniforms=u)
    mask = legal_mask(v.obs_buf)
    assert torch.equal(mask, v.obs["action_mask"].bool())
    x = logits.float().masked_fill(~mask, float("-inf"))
    logp_all = torch.log_softmax(x, dim=1)
    p = logp_all.exp()
    # every sampled action is legal
    assert bool(mask.gather(1, a.long().unsqueeze(1)).all())
    # log-prob and entropy: fp32 softmax statistics, tolerance 2e-5 absolute
    ref_lp = logp_all.gather(1, a.long().unsqueeze(1)).squeeze(1)
    assert float((lp - ref_lp).abs().max()) < 2e-5
    ref_ent = -(p * logp_all.masked_fill(~mask, 0)).sum(dim=1)
    assert float((ent - ref_ent).abs().max()) < 2e-5
    # inverse-CDF draw: the action's CDF interval contains u (up to 1e-5 of summation-order slack)
    cdf = p.cumsum(dim=1)
    hi = cdf.gather(1, a.long().unsqueeze(1)).squeeze(1)
    lo = hi - p.gather(1, a.long().unsqueeze(1)).squeeze(1)
    assert bool(((u >= lo - 1e-5) & (u <= hi + 1e-5)).all())
    # Philox path: deterministic in (seed, step, env index), different across steps, empirically unbiased
    a1, _, _ = masked_sample(logits, v.obs_buf, seed=7, step=3)
    a2, _, _ = masked_sample(logits, v.obs_buf, seed=7, step=3)
    a3, _, _ = masked_sample(logits, v.obs_buf, seed=7, step=4)
    assert torch.equal(a1, a2) and not torch.equal(a1, a3)
    half = n // 2
    b, _, _ = masked_sample(logits[half:].contiguous(), v.obs_buf[half:].contiguous(), seed=7, step=3, env_offset=half)
    assert torch.equal(b, a1[half:])                              # slab-invariant: keyed by the global env index
    flat = torch.zeros((n, 60), device="cuda")
    cnt = torch.zeros(60, device="cuda")
    for s in range(64):
        aa, _, _ = masked_sample(flat, v.obs_buf, seed=11, step=s)
        cnt += torch.bincount(aa.long(), minlength=60).float()
    expect = (mask.float() / mask.float().sum(dim=1, keepdim=True)).sum(dim=0) * 64
    sel = expect > 50
    assert float(((cnt[sel] - expect[sel]).abs() / expect[sel].sqrt()).max()) < 6.0   # within 6 sigma per action


def test_gae_matches_sb3_formula(torch):
    from balatro_gym_b200.rollout import gae
    T, n = 37, 5000
    g = torch.Generator(device="cuda").manual_seed(2)
    r = torch.randn((T, n), device="cuda", generator=g)
    val = torch.randn((T + 1, n), device="cuda", generator=g)
    d = (torch.rand((T, n), device="cuda", generator=g) < 0.1).to(torch.uint8)
    adv, ret = gae(r, val, d, 0.99, 0.95)
    # stable_baselines3 RolloutBuffer.compute_returns_and_advantage, in fp64 for the comparison
    r64, v64, nd = r.double(), val.double(), 1.0 - d.double()
    last = torch.zeros(n, device="cuda", dtype=torch.float64)
    ref = torch.zeros((T, n), device="cuda", dtype=torch.float64)
    for t in reversed(range(T)):
        delta = r64[t] + 0.99 * v64[t + 1] * nd[t] - v64[t]
        last = delta + 0.99 * 0.95 * nd[t] * last
        ref[t] = last
    assert float((adv.double() - ref).abs().max()) < 1e-4        # fp32 recursion over 37 steps
    assert float((ret.double() - (ref + v64[:T])).abs().max()) < 1e-4


def test_rollout_collector_end_to_end(torch):
    from balatro_gym_b200 import BalatroVecEnv
    from balatro_gym_b200.rollout import RolloutCollector, make_policy, evaluate_actions, legal_mask, ppo_update
    n, T = 4096, 16
    policy = make_policy(seed=0)

    def run():
        vec = BalatroVecEnv(n, seed=3)
        vec.reset()
        roll = RolloutCollector(vec, policy, n_steps=T, seed=9)
        return roll.collect()
    r1, r2 = run(), run()
    for name in ("obs", "actions", "rewards", "dones"):
        assert torch.equal(getattr(r1, name), getattr(r2, name)), name       # env side is bit-reproducible
    # every action taken was legal in the observation it was sampled from
    for t in range(T):
        assert bool(legal_mask(r1.obs[t]).gather(1, r1.actions[t].long().unsqueeze(1)).all()), t
    # the sampling kernel's log-prob agrees with the differentiable torch evaluation (bf16 forward: 2e-2)
    lp, ent, val = evaluate_actions(policy, r1.obs[3], r1.actions[3])
    assert float((lp.detach() - r1.logp[3]).abs().max()) < 2e-2
    assert float((val.detach() - r1.values[3]).abs().max()) < 2e-2
    steps, episodes, mean_r = r1.stats()
    assert steps == n * T and episodes >= 0 and np.isfinite(mean_r)
    # one PPO update runs and moves the parameters
    before = [p.detach().clone() for p in policy.parameters()]
    opt = torch.optim.Adam(policy.parameters(), lr=3e-4)
    out = ppo_update(policy, opt, r1, n_epochs=1, minibatch=16384)
    assert all(np.isfinite(v) for v in out.values()), out
    assert any(not torch.equal(a, b) for a, b in zip(before, policy.parameters()))


def test_fused_policy_forward_matches_torch(torch, mixed_obs):
    """bgym_policy_forward (fourteen layers in one tcgen05 kernel straight from the observation records: bf16 operands, fp32
    accumulation in tensor memory, bf16 activations between layers) against the same network in torch: fp32 math on the
    bf16-rounded weights and inputs, with the activations rounded to bf16 between layers as the kernel does.
    Tolerance: accumulation order + tanh.approx (2^-11) + one bf16 rounding per layer."""
    from balatro_gym_b200.rollout import make_policy, pack_policy_weights, policy_forward_fused, policy_program
    steps, wbytes, bfloats = policy_program()
    assert len(steps) == 72 and wbytes == int(steps["bytes"].sum()) and int(steps["last"].sum()) == 7
    v = mixed_obs
    pol = make_policy(seed=11)
    with torch.no_grad():       # make the last layers' outputs large enough to be a test
        pol.pi[4].weight.mul_(4.0); pol.vf[4].weight.mul_(4.0)
    sd = pol.state_dict()
    bf = lambda t: t.detach().to(torch.bfloat16)
    weights, bias = pack_policy_weights(sd, v.device)
    obs = v.obs_buf
    feats = reference_features(torch, v.obs).to(torch.bfloat16).float()            # the operands of the bf16 path
    lin = lambda x, key: x @ bf(sd[key + ".weight"]).float().t() + sd[key + ".bias"].float()
    r16 = lambda t: t.to(torch.bfloat16).float()
    for n in (obs.shape[0], 128, 129, 1):                                  # whole slab, one tile, ragged tiles
        logits, value = policy_forward_fused(obs[:n].contiguous(), weights, bias)
        torch.cuda.synchronize()
        x = feats[:n]
        h = r16(torch.relu(lin(x[:, :416], "hand_net.0"))); h = r16(torch.relu(lin(h, "hand_net.2")))
        j = r16(torch.relu(lin(x[:, 416:426], "joker_net.0"))); j = r16(torch.relu(lin(j, "joker_net.2")))
        g = r16(torch.relu(lin(x[:, 426:447], "game_state_net.0"))); g = r16(torch.relu(lin(g, "game_state_net.2")))
        z = r16(torch.relu(lin(torch.cat([h, j, g], dim=1), "combined_net.0")))
        z = r16(torch.relu(lin(z, "combined_net.2")))
        p = r16(torch.tanh(lin(z, "pi.0"))); p = r16(torch.tanh(lin(p, "pi.2"))); ref_logits = lin(p, "pi.4")
        q = r16(torch.tanh(lin(z, "vf.0"))); q = r16(torch.tanh(lin(q, "vf.2"))); ref_value = lin(q, "vf.4").squeeze(-1)
        assert logits.shape == (n, 60) and value.shape == (n,)
        err_l = float((logits - ref_logits).abs().max()); err_v = float((value - ref_value).abs().max())
        scale = float(ref_logits.abs().max())
        assert err_l < 0.03 * max(1.0, scale) and err_v < 0.03 * max(1.0, float(ref_value.abs().max())), (n, err_l, err_v, scale)
        assert scale > 0.3


def test_fused_policy_forward_cta_pairs(torch):
    """The cta_group::2 variant of the policy kernel (clusters of two CTAs, M = 256 MMAs, each CTA holding half of every weight
    tile; BGYM_POLICY_CTAS=2, read once per process): the same torch comparison in a child process."""
    import os
    import subprocess
    import sys
    if os.environ.get("BGYM_POLICY_CTAS") == "2":
        pytest.skip("already the child")
    env = dict(os.environ, BGYM_POLICY_CTAS="2")
    here = os.path.dirname(os.path.abspath(__file__))
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_rollout.py"), "-q", "-x", "-m", "gpu", "-k",
                          "fused_policy_forward_matches_torch"], env=env, capture_output=True, text=True, timeout=300, cwd=os.path.dirname(here))
    assert res.returncode == 0 and "1 passed" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
