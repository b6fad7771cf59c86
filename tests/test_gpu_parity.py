"""GPU parity tests: the CUDA path (through the C-ABI) against
  (1) golden traces recorded from the unmodified reference (replay mode, bit-exact),
  (2) the C oracle on identical seeded inputs (native Philox mode, bit-exact, 10^6+ env-steps),
  (3) size-independent properties at BASELINE.json's full sizes (2^24 hands, 2^20 envs).
All marked gpu; they run on the B200 box only."""
import numpy as np
import pytest

from conftest import load_trace, assert_records_equal, STATE_SKIP, reward_close
from balatro_gym_b200 import layout as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.fixture(params=["multi_pass", "one_launch"])
def step_path(request):
    """Both launch structures of bgym_step give identical results: `multi_pass` = main pass + three gather
    passes (what large slabs use), `one_launch` = the small-slab kernel (default for n <= 65536)."""
    import balatro_gym_b200 as b
    lib = b.load()
    assert lib.bgym_set_option(1, 0 if request.param == "multi_pass" else 1 << 40) == 0
    yield request.param
    assert lib.bgym_set_option(1, 65536) == 0


def _loaded_native_lib():
    import balatro_gym_b200 as b
    b.load()
    maps = open("/proc/self/maps").read()
    assert "libbgym.so" in maps, "native library not loaded"


# ---------------------------------------------------------------------------------------------
# (1) golden traces from the reference, replay mode
# ---------------------------------------------------------------------------------------------
class CudaStepper:
    def __init__(self, torch, E):
        from balatro_gym_b200 import BalatroVecEnv
        self.torch = torch
        self.v = BalatroVecEnv(E, autoreset=False)

    def reset(self, seeds, decks):
        t = self.torch
        self.v.reset(seeds=t.from_numpy(seeds % (2 ** 32)), decks52=t.from_numpy(decks))

    def set_state(self, init_state):
        cur = self.v.state_numpy()
        new = init_state.copy()
        for k in STATE_SKIP:
            new[k] = cur[k]
        self.v.inject_numpy(new)

    def step(self, actions, draws):
        t = self.torch
        d = t.from_numpy(np.ascontiguousarray(draws).view(np.uint8).reshape(len(actions), L.DRAWS_BYTES).copy())
        self.v.step(t.from_numpy(actions.astype(np.int32)).cuda(), draws=d.cuda())
        return (self.v.state_numpy(), self.v.obs_numpy(), self.v.reward.cpu().numpy(), self.v.terminated.cpu().numpy(),
                self.v.info_numpy())


@pytest.mark.parametrize("name", ["c1", "c3", "c4", "c4x"])
def test_cuda_replays_reference_trace(torch, name, step_path):
    from test_oracle_golden import replay
    tr = load_trace(name)
    n = replay(tr, CudaStepper(torch, tr["action"].shape[1]))
    assert n == int(tr["length"].sum())
    _loaded_native_lib()


# ---------------------------------------------------------------------------------------------
# (2) CUDA vs C oracle, native mode
# ---------------------------------------------------------------------------------------------
def _compare_step(v, ov, t, check_info=True):
    st, ob = v.state_numpy(), v.obs_numpy()
    assert_records_equal(ov.state, st, L.STATE_DTYPE, (), f"step {t} state")
    assert_records_equal(ov.obs, ob, L.OBS_DTYPE, (), f"step {t} obs")
    assert np.array_equal(ov.terminated, v.terminated.cpu().numpy()), t
    r = v.reward.cpu().numpy()
    # the pre-step ante is not kept; use a safe superset: tolerance only where rewards differ AND a hand was played
    inf = v.info_numpy()
    rc = (ov.reward == r) | (((inf["flags"] & L.F_PLAYED) != 0) & (np.abs(ov.reward - r) <= 1e-12 * np.maximum(1, np.abs(r))))
    assert rc.all(), (t, ov.reward[~rc][:4], r[~rc][:4])
    if check_info:
        assert_records_equal(ov.info, inf, L.INFO_DTYPE, (), f"step {t} info")


@pytest.mark.parametrize("n,steps,mode", [(4096, 260, "vanilla"), (4096, 260, "host_c3"), (1000, 120, "host_c3"),
                                          (4096, 260, "c3"), (4096, 400, "c4"), (1000, 120, "c4")])
def test_cuda_vs_oracle_native_rollout(torch, n, steps, mode, step_path):
    """Fused random-legal policy + autoreset on both sides: > 10^6 env-steps compared record by record.
    mode: vanilla = the reference's reset state; host_c3 = config-3 state scattered once from the host;
    c3 / c4 = the device-side state generator (BGYM_FLAG_GEN_C3 [| BGYM_FLAG_GEN_CONS]) applied by reset and by
    every in-kernel autoreset, mirrored by the oracle."""
    from balatro_gym_b200 import BalatroVecEnv
    from oracle import coracle
    gen = mode if mode in ("c3", "c4") else None
    gflags = L.GENERATORS[gen]
    v = BalatroVecEnv(n, seed=1, autoreset=True, generator=gen)
    v.reset()
    if mode == "host_c3":
        v.randomize_c3(seed=5)
    ov = coracle.OracleVec(n)
    coracle.reset(ov.state, ov.obs, np.arange(1, n + 1), flags=gflags)
    st0 = v.state_numpy()
    if mode != "host_c3":
        assert_records_equal(ov.state, st0, L.STATE_DTYPE, (), "reset state")   # native Philox shuffle (+ generator) parity
        assert_records_equal(ov.obs, v.obs_numpy(), L.OBS_DTYPE, (), "reset obs")
    ov.state[:] = st0
    rng = np.random.default_rng(3)
    for t in range(steps):
        if t % 7 == 3:   # explicit actions incl. masked / out-of-range ones
            act = rng.integers(-2, 62, size=n).astype(np.int32)
            v.step(torch.from_numpy(act).cuda())
            ov.step(act, flags=L.FLAG_AUTORESET | gflags)
        else:
            v.step(random_policy=True)
            oact = np.zeros(n, np.int32)
            coracle.step(ov.state, oact, ov.obs, ov.reward, ov.terminated, ov.truncated, ov.info, None,
                         flags=L.FLAG_AUTORESET | L.FLAG_RANDOM_POLICY | gflags)
            assert np.array_equal(oact, v.actions.cpu().numpy()), f"policy actions differ at step {t}"
        _compare_step(v, ov, t)
    st = v.state_numpy()
    assert int(st["episode"].sum()) > 0      # autoreset happened
    if gen:
        # every env still carries generator state, however many episodes it has been through
        assert ((st["deck"] >> 6) != 0).any(axis=1).all()
        if gen == "c4":     # consumables were used (the OTHER gather list's rare path ran)
            assert int((st["cons_n"] < 2).sum()) > n // 8


def test_generator_reset_law_and_replayed_decks(torch):
    """BGYM_FLAG_GEN_C3 | BGYM_FLAG_GEN_CONS at reset: CUDA == oracle bit for bit (native shuffle and a replayed
    permutation), and the generated state follows SURVEY 8(d) C3: 5 distinct shop-eligible jokers, enhancement w.p.
    1/4 over 8, edition w.p. 1/10 over 3, seal w.p. 1/10 over 4, two consumables over the 52 names."""
    from balatro_gym_b200 import BalatroVecEnv
    from oracle import coracle
    n = 1 << 16
    v = BalatroVecEnv(n, seed=1000, autoreset=False, generator="c4")
    v.reset()
    ov = coracle.OracleVec(n)
    coracle.reset(ov.state, ov.obs, np.arange(1000, n + 1000), flags=L.GENERATORS["c4"])
    st = v.state_numpy()
    assert_records_equal(ov.state, st, L.STATE_DTYPE, (), "generated reset state")
    assert_records_equal(ov.obs, v.obs_numpy(), L.OBS_DTYPE, (), "generated reset obs")
    jk = st["joker_id"][:, :5].astype(int)
    assert (st["joker_n"] == 5).all() and jk.min() >= 1 and jk.max() <= 145 and (st["joker_id"][:, 5:] == 0).all()
    assert (np.diff(np.sort(jk, axis=1), axis=1) > 0).all()                      # distinct
    cnt = np.bincount(jk.ravel(), minlength=146)[1:]
    e = n * 5 / 145.0
    assert abs(((cnt - e) ** 2 / e).sum() - 144) < 6 * (2 * 144) ** 0.5          # every id equally likely
    deck = st["deck"].astype(int)
    assert (np.sort(deck & 63, axis=1) == np.arange(52)).all()
    enh, ed, seal = (deck >> 6) & 15, (deck >> 10) & 7, (deck >> 13) & 7
    N = n * 52
    for arr, p, k in ((enh, 0.25, 8), (ed, 0.1, 3), (seal, 0.1, 4)):
        assert arr.max() == k
        assert abs((arr != 0).mean() - p) < 5 * (p * (1 - p) / N) ** 0.5
        c = np.bincount(arr.ravel(), minlength=k + 1)[1:]
        assert abs(((c - c.mean()) ** 2 / c.mean()).sum() - (k - 1)) < 6 * (2 * (k - 1)) ** 0.5 + 6
    cons = st["cons_id"][:, :2].astype(int)
    assert (st["cons_n"] == 2).all()
    ids = np.array(list(range(1, 23)) + list(range(30, 42)) + list(range(50, 68)))
    assert np.isin(cons, ids).all()
    cc = np.array([(cons == i).sum() for i in ids])
    e = n * 2 / 52.0
    assert abs(((cc - e) ** 2 / e).sum() - 51) < 6 * (2 * 51) ** 0.5
    # a replayed permutation (the reference's shuffle stream) keeps each card's generated modifiers
    rng = np.random.default_rng(0)
    m = 512
    decks = np.argsort(rng.random((m, 52)), axis=1).astype(np.uint8)
    v2 = BalatroVecEnv(m, seed=1000, autoreset=False, generator="c3")
    v2.reset(decks52=torch.from_numpy(decks))
    ov2 = coracle.OracleVec(m)
    coracle.reset(ov2.state, ov2.obs, np.arange(1000, m + 1000), decks52=decks, flags=L.GENERATORS["c3"])
    s2 = v2.state_numpy()
    assert_records_equal(ov2.state, s2, L.STATE_DTYPE, (), "generated reset state, replayed decks")
    assert (s2["deck"] & 63 == decks).all() and (s2["cons_n"] == 0).all()
    a = np.take_along_axis(s2["deck"], np.argsort(s2["deck"] & 63, axis=1), axis=1)      # by card identity
    b = np.take_along_axis(st["deck"][:m], np.argsort(st["deck"][:m] & 63, axis=1), axis=1)
    assert (a == b).all()


@pytest.mark.parametrize("n", [1, 31, 33, 97])
def test_ragged_and_tiny_slabs_match_the_oracle(torch, n, step_path):
    """Slab sizes that are not a multiple of the 32-env tile, down to a single env."""
    from balatro_gym_b200 import BalatroVecEnv
    from oracle import coracle
    v = BalatroVecEnv(n, seed=11, autoreset=True)
    v.reset()
    ov = coracle.OracleVec(n)
    ov.reset(np.arange(11, n + 11))
    assert_records_equal(ov.state, v.state_numpy(), L.STATE_DTYPE, (), "reset state")
    for t in range(150):
        v.step(random_policy=True)
        oact = np.zeros(n, np.int32)
        coracle.step(ov.state, oact, ov.obs, ov.reward, ov.terminated, ov.truncated, ov.info, None,
                     flags=L.FLAG_AUTORESET | 4)
        assert np.array_equal(oact, v.actions.cpu().numpy())
        _compare_step(v, ov, t)


def test_step_without_observations(torch, step_path):
    """BGYM_FLAG_NO_OBS with obs = NULL: same state, rewards and terminations as the observing step."""
    import balatro_gym_b200 as b
    from balatro_gym_b200 import BalatroVecEnv
    lib = b.load()
    n = 3000
    x, y = BalatroVecEnv(n, seed=4), BalatroVecEnv(n, seed=4)
    for v in (x, y):
        v.reset()
        v.randomize_c3(4)
    flags = L.FLAG_AUTORESET | 4
    for t in range(120):
        x.step(random_policy=True)
        rc = lib.bgym_step(y._hot.data_ptr(), y.tog.data_ptr(), y.cold.data_ptr(), y.actions.data_ptr(), None, None, None, None,
                           y.reward.data_ptr(), y.terminated.data_ptr(), y.truncated.data_ptr(), y.info_buf.data_ptr(), n,
                           flags | L.FLAG_NO_OBS, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        y._hot_whole = False
    torch.cuda.synchronize()
    assert torch.equal(x.tog, y.tog)
    for name in ("hot", "cold", "reward", "terminated", "actions", "info_buf"):
        assert torch.equal(getattr(x, name), getattr(y, name)), name
    # and obs = NULL without the flag is an argument error, not a crash
    assert lib.bgym_step(y._hot.data_ptr(), y.tog.data_ptr(), y.cold.data_ptr(), y.actions.data_ptr(), None, None, None, None,
                         y.reward.data_ptr(), y.terminated.data_ptr(), y.truncated.data_ptr(), None, n, flags, None) < 0


def test_refresh_observations_after_state_injection(torch):
    """State changed from outside (inject_numpy / load_state): the observation buffer is re-emitted from the state, so
    the next sampled actions are legal for the injected state."""
    from balatro_gym_b200 import BalatroVecEnv
    from oracle import coracle
    n = 2048
    v = BalatroVecEnv(n, seed=21, generator="c4")
    v.reset()
    for _ in range(40):
        v.step(random_policy=True)
    ck = v.save_state()
    st = v.state_numpy()
    w = BalatroVecEnv(n, seed=999)
    w.reset()
    w.inject_numpy(st)
    assert torch.equal(w.obs_buf, v.obs_buf)            # observations follow the injected state
    for _ in range(10):
        v.step(random_policy=True)
    v.load_state(ck)
    assert torch.equal(w.obs_buf, v.obs_buf) and torch.equal(w.hot, v.hot)
    a = v.sample_actions(seed=5)
    m = coracle.action_mask(v.state_numpy())
    assert ((m >> a.cpu().numpy().astype(np.uint64)) & 1).all()


def test_empty_slab_calls_are_noops(torch):
    import balatro_gym_b200 as b
    lib = b.load()
    z = torch.zeros(16, dtype=torch.uint8, device="cuda")
    p = z.data_ptr()
    assert lib.bgym_reset(p, p, p, p, p, None, p, None, 0, 0, None) == 0
    assert lib.bgym_step(p, p, p, p, None, p, p, None, p, p, p, None, 0, 0, None) == 0
    assert lib.bgym_action_mask(p, p, p, p, 0, None) == 0
    assert lib.bgym_sync_state(p, p, 0, 0, None) == 0 and lib.bgym_sync_obs(p, p, 0, 1, None) == 0
    assert lib.bgym_sample_actions(p, 16, p, 1, 0, 0, None) == 0
    assert lib.bgym_featurize(p, p, 0, 0, None) == 0
    assert lib.bgym_gae(p, p, p, 0.99, 0.95, p, p, 0, 0, None) == 0
    # argument errors come back as negative codes with a message, never a crash
    assert lib.bgym_sample_actions(p, 12, p, 1, 0, 4, None) < 0      # stride not a multiple of 8
    assert lib.bgym_step(None, p, p, p, None, p, p, None, p, p, p, None, 4, 0, None) < 0
    assert b"bgym_step" in lib.bgym_last_error()
    assert lib.bgym_featurize(p, p, 4, 7, None) < 0


def test_sampler_kernel_and_mask_kernel(torch):
    from balatro_gym_b200 import BalatroVecEnv
    from oracle import coracle
    n = 5000
    v = BalatroVecEnv(n, seed=77, autoreset=True)
    v.reset()
    for t in range(40):
        a = v.sample_actions(seed=123)
        obs = v.obs_numpy()
        exp = coracle.sample_actions(obs, 123, v._step_count)
        got = a.cpu().numpy()
        assert np.array_equal(exp, got)
        assert ((obs["action_mask_bits"] >> got.astype(np.uint64)) & 1).all()      # always legal
        m = v.action_masks().cpu().numpy().view(np.uint64)
        assert np.array_equal(m, obs["action_mask_bits"])
        assert np.array_equal(m, coracle.action_mask(v.state_numpy()))
        v.step(a)


def test_host_buffer_handle_api(torch):
    """bgym_vec_* (what a non-torch caller binds): host pointers in, host pointers out."""
    import ctypes as C
    from balatro_gym_b200 import _lib
    from oracle import coracle
    lib = _lib.load()
    n = 777
    h = C.c_void_p()
    _lib.check(lib.bgym_vec_create(C.byref(h), n, 0), "create")
    seeds = np.arange(10, 10 + n, dtype=np.uint32)
    obs = np.zeros(n, L.OBS_DTYPE)
    _lib.check(lib.bgym_vec_reset_host(h, seeds.ctypes.data, None, obs.ctypes.data), "reset_host")
    ov = coracle.OracleVec(n)
    ov.reset(seeds)
    assert_records_equal(ov.obs, obs, L.OBS_DTYPE, (), "host reset obs")
    reward = np.zeros(n); term = np.zeros(n, np.uint8); trunc = np.zeros(n, np.uint8); info = np.zeros(n, L.INFO_DTYPE)
    for t in range(30):
        act = coracle.sample_actions(obs, 5, t)
        _lib.check(lib.bgym_vec_step_host(h, act.ctypes.data, None, obs.ctypes.data, reward.ctypes.data, term.ctypes.data,
                                          trunc.ctypes.data, info.ctypes.data, L.FLAG_AUTORESET), "step_host")
        ov.step(act, flags=L.FLAG_AUTORESET)
        assert_records_equal(ov.obs, obs, L.OBS_DTYPE, (), f"host step {t}")
        assert np.array_equal(ov.terminated, term)
    st = np.zeros(n, L.STATE_DTYPE)
    _lib.check(lib.bgym_vec_get_state(h, st.ctypes.data), "get_state")
    assert_records_equal(ov.state, st, L.STATE_DTYPE, (), "host state")
    _lib.check(lib.bgym_vec_destroy(h), "destroy")


# ---------------------------------------------------------------------------------------------
# hand scoring
# ---------------------------------------------------------------------------------------------
def _random_hands(n, seed, jokers=True, mods=True):
    rng = np.random.default_rng(seed)
    keys = rng.random((n, 52))
    cards = np.argsort(keys, axis=1)[:, :8].astype(np.uint8)
    nc = rng.integers(1, 9, size=n).astype(np.uint8)
    m = np.zeros((n, 8), np.uint16)
    if mods:
        enh = np.where(rng.random((n, 8)) < 0.3, rng.integers(1, 9, (n, 8)), 0)
        ed = np.where(rng.random((n, 8)) < 0.15, rng.integers(1, 4, (n, 8)), 0)
        seal = np.where(rng.random((n, 8)) < 0.1, rng.integers(1, 5, (n, 8)), 0)
        m = (enh | (ed << 4) | (seal << 8)).astype(np.uint16)
    jk = np.zeros((n, 8), np.uint8)
    if jokers:
        jk[:, :5] = np.argsort(rng.random((n, 150)), axis=1)[:, :5] + 1
        jk[rng.random((n, 8)) < 0.3] = 0
    lv = rng.integers(1, 17, (n, 12)).astype(np.uint8)
    ctx = np.zeros(n, L.SCORE_CTX_DTYPE)
    ctx["hands_left"] = rng.integers(1, 5, n); ctx["discards_left"] = rng.integers(0, 4, n)
    ctx["deck_len"] = rng.integers(40, 53, n)
    return cards, m, nc, jk, lv, ctx


@pytest.mark.parametrize("table_names", [False, True])
def test_score_hands_cuda_vs_oracle(torch, table_names):
    from balatro_gym_b200 import score_hands
    from oracle import coracle
    n = 1 << 18
    cards, m, nc, jk, lv, ctx = _random_hands(n, 11)
    exp = coracle.score_hands(cards, m, nc, jk, lv, ctx, seed=99, flags=int(table_names))
    t = lambda a: torch.from_numpy(a.view(np.uint8).reshape(n, -1) if a.dtype.fields else a).cuda()
    out = score_hands(t(cards), t(m), t(nc), t(jk), t(lv), t(ctx), seed=99, table_names=table_names)
    for k in ("hand_type", "chips", "mult", "x_mult", "score", "money"):
        assert np.array_equal(exp[k], out[k].cpu().numpy()), k


def test_score_hands_replayed_reference_cases(torch, reference):
    """Reference O2 cases (jokers by name, tapped Misprint/Bloodstone draws) straight into the CUDA kernel."""
    from balatro_gym_b200 import score_hands
    from test_oracle_vs_reference import make_score_cases, _pack, check_scores
    for tn in (False, True):
        hands = make_score_cases(reference, 800, 1234 + tn, tn)
        cards, mods, nc, jk, lv, ctx = _pack(hands)
        t = lambda a: torch.from_numpy(a.view(np.uint8).reshape(len(hands), -1) if a.dtype.fields else a).cuda()
        out = score_hands(t(cards), t(mods), t(nc), t(jk), t(lv), t(ctx), table_names=tn)
        check_scores(hands, {k: v.cpu().numpy() for k, v in out.items()})


def test_known_answers_cuda(torch):
    from balatro_gym_b200 import score_hands
    from test_known_answers import known_arrays
    cards, n, ht, score = known_arrays()
    out = score_hands(torch.from_numpy(cards).cuda(), n_cards=torch.from_numpy(n).cuda())
    assert out["hand_type"].cpu().tolist() == ht.tolist()
    assert out["score"].cpu().tolist() == score.tolist()


def test_score_hands_full_size_properties(torch):
    """BASELINE config 2 at full size (2^24 five-card plays, no jokers, level 1): properties that do
    not need the oracle at that size + an oracle check on a 2^16 sample."""
    from balatro_gym_b200 import score_hands
    from oracle import coracle
    n = 1 << 24
    g = torch.Generator(device="cuda"); g.manual_seed(0x5EED)
    keys = torch.rand((n, 52), device="cuda", generator=g)
    cards = torch.zeros((n, 8), dtype=torch.uint8, device="cuda")
    cards[:, :5] = keys.topk(5, dim=1).indices.to(torch.uint8)
    del keys
    out = score_hands(cards, want_x_mult=False, want_money=False)
    ht = out["hand_type"]; chips = out["chips"].long(); mult = out["mult"].long(); score = out["score"]
    assert bool((score == chips * mult).all())                      # x_mult == 1 without jokers
    # card-order invariance (classification and chip sums are symmetric)
    perm = cards.clone(); perm[:, :5] = cards[:, [4, 2, 0, 3, 1]]
    out2 = score_hands(perm, want_x_mult=False, want_money=False)
    assert bool((out2["score"] == score).all()) and bool((out2["hand_type"] == ht).all())
    # hand-type mix of uniform 5-card draws (exact combinatorics): HC .5012 1P .4226 2P .0475 3K .0211 S .0039 F .0020 FH .0014 4K .00024 SF .000015
    frac = torch.bincount(ht.long(), minlength=12).double() / n
    expect = [0.501177, 0.422569, 0.047539, 0.021128, 0.003925, 0.001965, 0.001441, 0.000240, 0.0000154]
    for i, e in enumerate(expect):
        assert abs(float(frac[i]) - e) < 4 * (e * (1 - e) / n) ** 0.5 + 1e-6, (i, float(frac[i]), e)
    assert float(frac[9:].sum()) == 0.0
    # chips lower bound: base chips of the hand type + at least 5 cards x 2
    idx = torch.arange(0, n, 256, device="cuda")
    sub = cards[idx].cpu().numpy()
    exp = coracle.score_hands(sub)
    assert np.array_equal(exp["score"], score[idx].cpu().numpy())
    assert np.array_equal(exp["hand_type"], ht[idx].cpu().numpy())


# ---------------------------------------------------------------------------------------------
# native Philox mode: distributional tests (the reference's MT19937 shuffles are matched by replay)
# ---------------------------------------------------------------------------------------------
def test_native_shuffle_is_uniform(torch):
    from balatro_gym_b200 import BalatroVecEnv
    n = 1 << 17
    v = BalatroVecEnv(n, seed=12345, autoreset=False)
    v.reset()
    deck = (v.state_field("deck") & 63).long()
    assert bool((deck.sort(dim=1).values == torch.arange(52, device="cuda")).all())     # every deck is a permutation
    # position x card contingency table: chi-square against uniform, 51*51 dof
    table = torch.zeros((52, 52), dtype=torch.float64, device="cuda")
    pos = torch.arange(52, device="cuda").expand(n, 52)
    table.index_put_((pos.reshape(-1), deck.reshape(-1)), torch.ones(n * 52, dtype=torch.float64, device="cuda"), accumulate=True)
    e = n / 52.0
    chi2 = float(((table - e) ** 2 / e).sum())
    dof = 51 * 51
    assert abs(chi2 - dof) < 6 * (2 * dof) ** 0.5, chi2
    # different seeds give different decks; same seed reproduces
    v2 = BalatroVecEnv(n, seed=12345, autoreset=False)
    v2.reset()
    assert bool((v2.hot == v.hot).all()) and bool((v2.cold == v.cold).all())
    assert int((deck[1:] == deck[:-1]).all(dim=1).sum()) == 0


def test_full_size_rollout_properties(torch):
    """BASELINE configs 3/4 at full per-GPU size (2^20 envs): invariants that hold for every env after
    many fused random-policy steps with autoreset."""
    from balatro_gym_b200 import BalatroVecEnv
    n = 1 << 20
    v = BalatroVecEnv(n, seed=1, autoreset=True)
    v.reset()
    v.randomize_c3(seed=1)
    for t in range(96):
        v.step(random_policy=True, want_info=False)
    torch.cuda.synchronize()
    hand_n = v.state_field("hand_n").long(); phase = v.state_field("phase").long()
    deck = (v.state_field("deck") & 63).long()
    assert bool((deck.sort(dim=1).values == torch.arange(52, device="cuda")).all())
    assert bool((hand_n <= 8).all()) and bool((phase <= 2).all())
    hand = v.state_field("hand").long()
    # the hand_code cache of the hot record agrees with deck[hand[i]]
    codes = v.state_field("hand_code").long()
    look = torch.gather(deck, 1, hand.clamp(max=51))
    okc = torch.where(torch.arange(8, device="cuda")[None, :] < hand_n[:, None], codes == look, codes == 255)
    assert bool(okc.all())
    valid = torch.arange(8, device="cuda")[None, :] < hand_n[:, None]
    assert bool(((hand < 52) | ~valid).all()) and bool(((hand == 255) | valid).all())
    # hand slots hold distinct deck indices
    h = torch.where(valid, hand, torch.arange(100, 108, device="cuda")[None, :].expand(n, 8))
    assert bool((h.sort(dim=1).values.diff(dim=1) != 0).all())
    # the obs mask word equals the mask recomputed from the state, and every sampled action was legal
    m = v.action_masks()
    assert bool((m == v.obs["action_mask_bits"]).all())
    ob = v.obs
    assert bool((ob["money"] == v.state_field("money")).all())
    assert bool((ob["phase"].long() == phase).all())
    bits = ((m[:, None] >> torch.arange(60, device="cuda")[None, :]) & 1).to(torch.int8)
    assert bool((bits == ob["action_mask"]).all())
    # rewards are finite and terminations happened and were reset in place
    assert bool(torch.isfinite(v.reward).all())
    assert int(v.state_field("episode").long().sum()) > n // 4


def test_gym_facade_matches_reference_episode(torch, reference):
    """BalatroEnv (N=1 facade) against the unmodified reference on whole episodes: same seed -> same
    deck (replayed MT19937 shuffle), then identical obs / reward / termination for small-blind episodes
    (config 1: no further random draws are consumed)."""
    from balatro_gym_b200.env import BalatroEnv
    from oracle.refenv import RefEnv
    rng = np.random.default_rng(0)
    for seed in (3, 17, 91):
        ref = RefEnv(seed=seed, tap=False)
        env = BalatroEnv(seed=seed)
        ro, _ = ref.reset(seed)
        eo, _ = env.reset(seed=seed)
        for t in range(300):
            for k in L.OBS_KEYS:
                assert np.array_equal(np.asarray(ro[k]), np.asarray(eo[k])), (seed, t, k)
            if ref.env.state.phase == 1:       # shop draws need replay; stop the facade check there
                break
            a = 45 if t == 0 else int(rng.choice(np.flatnonzero(ro["action_mask"])))
            ro, rr, rt, _, ri = ref.step(a)
            eo, er, et, _, ei = env.step(a)
            if ref.env.state.phase == 1:
                break
            assert rr == er and rt == et, (seed, t, rr, er)
            assert ("error" in ri) == ("error" in ei)
            if rt:
                break


def test_gym_facade_unseeded_resets_continue_the_reference_streams(torch, reference):
    """`BalatroEnv(seed=s)` then plain `reset()` calls: every episode gets the NEXT shuffle of the seed's MT19937 stream
    (the reference rebuilds its RNG only when a seed is passed, balatro_env_2.py:507-509; the constructor's own reset
    takes the first shuffle), and a new key for the native in-game draws."""
    from balatro_gym_b200.env import BalatroEnv
    from oracle.refenv import card_code
    for seed in (5, 77):
        ref = reference.BalatroEnv(seed=seed)
        env = BalatroEnv(seed=seed)
        decks, keys = [], []
        for k in range(4):
            rd = [card_code(c) for c in ref.state.deck]
            st = env.state
            assert (st["deck"] & 63).tolist() == rd, (seed, k)
            decks.append(tuple(rd)); keys.append(int(st["rng_seed"]))
            ro, _ = ref.reset()
            eo, _ = env.reset()
            for key in L.OBS_KEYS:
                assert np.array_equal(np.asarray(ro[key]), np.asarray(eo[key])), (seed, k, key)
        assert len(set(decks)) == 4 and len(set(keys)) == 4          # a new deck and a new draw key every episode
        ref.reset(seed=seed); env.reset(seed=seed)                     # a seed restarts the streams
        assert (env.state["deck"] & 63).tolist() == [card_code(c) for c in ref.state.deck]
        assert tuple((env.state["deck"] & 63).tolist()) == decks[0]   # = the first shuffle of the seed's stream again
        env.close()


def _chi2_ok(counts, expected, dof=None):
    counts = np.asarray(counts, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64) * np.ones_like(counts)
    chi2 = float(((counts - expected) ** 2 / expected).sum())
    dof = dof if dof is not None else len(counts) - 1
    return abs(chi2 - dof) < 6 * (2 * dof) ** 0.5 + 6, chi2


def test_native_draw_distributions(torch):
    """Native Philox mode, distribution of the in-game draws (the reference's own draws are matched by
    replay): boss pick uniform over 28 (boss_blinds.py:522-532); shop inventory — third pack uniform over 3, three
    distinct jokers uniform over the 145 shop-eligible ids, voucher over 2, two cards over 52 (shop.py:112-139);
    lucky-card money roll p = 0.0667 per played lucky card (balatro_env_2.py:719-724); The Wheel's face-down roll
    p = 1/7 per hand card (boss_blinds.py:353)."""
    from balatro_gym_b200 import BalatroVecEnv
    n = 1 << 17
    dev = "cuda"
    full = lambda a: torch.full((n,), a, dtype=torch.int32, device=dev)
    # ---- boss pick ----
    v = BalatroVecEnv(n, seed=4242, autoreset=False)
    v.reset()
    v.step(full(47))
    boss = v.state_field("boss_type").long()
    assert int(boss.min()) == 1 and int(boss.max()) == 28
    ok, chi2 = _chi2_ok(torch.bincount(boss, minlength=29)[1:].cpu().numpy(), n / 28.0)
    assert ok, chi2
    # ---- shop inventory (skip blind -> round advance -> shop generation) ----
    v = BalatroVecEnv(n, seed=777, autoreset=False)
    v.reset()
    v.step(full(48))
    st = v.state_numpy()
    assert (st["phase"] == L.PHASE_SHOP).all() and (st["n_items"] == 9).all()
    it, iid = st["item_type"], st["item_id"].astype(int)
    assert (it[:, :3] == 1).all() and (it[:, 3:6] == 3).all() and (it[:, 6] == 4).all() and (it[:, 7:9] == 2).all()
    assert (iid[:, 0] == 0).all() and (iid[:, 1] == 1).all()
    ok, chi2 = _chi2_ok(np.bincount(iid[:, 2], minlength=5)[2:], n / 3.0); assert ok, ("third pack", chi2)
    jk = iid[:, 3:6]
    assert jk.min() >= 1 and jk.max() <= 145 and (np.diff(np.sort(jk, axis=1), axis=1) > 0).all()
    ok, chi2 = _chi2_ok(np.bincount(jk.ravel(), minlength=146)[1:], 3 * n / 145.0); assert ok, ("shop jokers", chi2)
    for pos in range(3):          # every position of the ordered sample is uniform too
        ok, chi2 = _chi2_ok(np.bincount(jk[:, pos], minlength=146)[1:], n / 145.0); assert ok, ("shop joker pos", pos, chi2)
    ok, chi2 = _chi2_ok(np.bincount(iid[:, 6], minlength=2), n / 2.0); assert ok, ("voucher", chi2)
    ok, chi2 = _chi2_ok(np.bincount(iid[:, 7:9].ravel(), minlength=52), 2 * n / 52.0); assert ok, ("shop cards", chi2)
    # ---- lucky money roll: every card LUCKY, small blind, select five cards, play ----
    v = BalatroVecEnv(n, seed=99, autoreset=False)
    v.reset()
    deck = v.state_field("deck")
    deck.copy_(((deck.to(torch.int32) & 63) | (L.card16(0, 8) & ~63)).to(torch.int16))
    v.refresh_observations()
    v.step(full(45))
    for s in range(5):
        v.step(full(2 + s))
    m0 = v.state_field("money").clone()
    v.step(full(0))
    gained = (v.state_field("money") - m0).long()
    played = (v.info_field("flags").long() & L.F_PLAYED) != 0
    assert bool(played.all())
    # a play that beats the blind also pays the round reward; look at hands that did not
    cont = (v.info_field("flags").long() & L.F_BEAT_BLIND) == 0
    g = gained[cont]
    assert bool((g % 20 == 0).all()) and int(g.max()) <= 100
    hits = float((g // 20).sum()); trials = 5.0 * int(cont.sum())
    p = 0.0667
    assert abs(hits / trials - p) < 5 * (p * (1 - p) / trials) ** 0.5, hits / trials
    # ---- The Wheel: 1 in 7 hand cards face down after a non-final play ----
    v = BalatroVecEnv(n, seed=31337, autoreset=False)
    v.reset()
    v.step(full(47))
    wheel = v.state_field("boss_type") == 3
    v.step(full(2)); v.step(full(0))
    cont = wheel & ((v.info_field("flags").long() & (L.F_PLAYED | L.F_BEAT_BLIND | L.F_FAILED)) == L.F_PLAYED)
    fd = v.state_field("face_down_mask").long()[cont]
    hn = v.state_field("hand_n").long()[cont]
    assert int(cont.sum()) > 2000
    bits = float(sum(((fd >> i) & 1).sum() for i in range(8))); trials = float(hn.sum())
    assert abs(bits / trials - 1 / 7) < 5 * ((1 / 7) * (6 / 7) / trials) ** 0.5, bits / trials


class HostHandleStepper:
    """The golden traces through the HOST-buffer entry points (bgym_vec_* with host pointers, draws included): what a
    non-torch caller binds."""

    def __init__(self, E):
        import ctypes as C
        from balatro_gym_b200 import _lib
        self.lib, self._lib, self.E = _lib.load(), _lib, E
        self.h = C.c_void_p()
        _lib.check(self.lib.bgym_vec_create(C.byref(self.h), E, 0), "create")
        self.obs = np.zeros(E, L.OBS_DTYPE); self.reward = np.zeros(E); self.term = np.zeros(E, np.uint8)
        self.trunc = np.zeros(E, np.uint8); self.info = np.zeros(E, L.INFO_DTYPE); self.state = np.zeros(E, L.STATE_DTYPE)

    def reset(self, seeds, decks):
        s = np.ascontiguousarray(seeds % (2 ** 32), dtype=np.uint32)
        d = np.ascontiguousarray(decks, dtype=np.uint8)
        self._lib.check(self.lib.bgym_vec_reset_host(self.h, s.ctypes.data, d.ctypes.data, self.obs.ctypes.data), "reset_host")

    def set_state(self, init_state):
        self._lib.check(self.lib.bgym_vec_get_state(self.h, self.state.ctypes.data), "get_state")
        new = init_state.copy()
        for k in STATE_SKIP:
            new[k] = self.state[k]
        self._lib.check(self.lib.bgym_vec_set_state(self.h, new.ctypes.data), "set_state")

    def step(self, actions, draws):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        d = np.ascontiguousarray(draws)
        self._lib.check(self.lib.bgym_vec_step_host(self.h, a.ctypes.data, d.ctypes.data, self.obs.ctypes.data, self.reward.ctypes.data,
                                                    self.term.ctypes.data, self.trunc.ctypes.data, self.info.ctypes.data, 0), "step_host")
        self._lib.check(self.lib.bgym_vec_get_state(self.h, self.state.ctypes.data), "get_state")
        return self.state, self.obs, self.reward, self.term, self.info

    def close(self):
        self.lib.bgym_vec_destroy(self.h)


@pytest.mark.parametrize("name", ["c4", "c4x"])
def test_host_handle_replays_reference_trace_with_draws(torch, name):
    """Whole reference episodes — boss blinds, shops, rerolls, consumables — through bgym_vec_step_host with the
    reference's recorded draws passed as HOST BgymDraws records."""
    from test_oracle_golden import replay
    tr = load_trace(name)
    st = HostHandleStepper(tr["action"].shape[1])
    try:
        n = replay(tr, st)
    finally:
        st.close()
    assert n == int(tr["length"].sum())
    # the trace really goes through the shop and boss blinds
    assert (tr["state"]["phase"] == L.PHASE_SHOP).any() and (tr["state"]["boss_type"] != 0).any()


def test_side_arrays_and_host_mirror_track_the_records(torch, step_path):
    """The toggle / selection arrays (BgymTog, BgymSel) against the whole records, and HostMirror — the pinned-host copy
    kept current by observation deltas — against the device arrays, over a rollout with autoresets, masked-out actions
    and the c4 generator.  The oracle (whole records, no side arrays) steps the same actions: after every step the
    synced device records equal its records, and the host mirror equals both."""
    from balatro_gym_b200 import BalatroVecEnv, HostMirror
    from oracle import coracle
    n = 6000
    v = BalatroVecEnv(n, seed=31, generator="c4")
    v.reset()
    ov = coracle.OracleVec(n)
    coracle.reset(ov.state, ov.obs, np.arange(31, 31 + n), flags=L.GENERATORS["c4"])
    m = HostMirror(v)
    m.pull_all()
    rng = np.random.default_rng(5)
    total_dirty = total_shop = 0
    for t in range(90):
        v.sample_actions(seed=17)
        act = v.actions.cpu().numpy().copy()
        wild = rng.random(n) < 0.05
        act[wild] = rng.integers(-2, 64, size=int(wild.sum()))          # rejected / out-of-range actions change nothing
        m.actions.copy_(torch.from_numpy(act))
        m.step(want_info=True)
        oobs, orew, oterm = ov.step(act.astype(np.int32), flags=L.FLAG_AUTORESET | L.GENERATORS["c4"])[:3]
        m.wait()
        total_dirty += m.delta_counts()[0]
        total_shop += m.delta_counts()[1]
        # side arrays vs whole records (device): the sync folds them in; nothing else changes
        tog = v.tog.cpu().numpy().reshape(-1).view(L.TOG_DTYPE)
        st = v.state_numpy()
        for name in L.TOG_OWNED_FIELDS:
            assert np.array_equal(tog[name], st[name]), (t, name)
        assert np.array_equal(tog["discards_left"], st["discards_left"]) and np.array_equal(tog["cons_n"], st["cons_n"])
        assert np.array_equal(tog["rng_seed"], st["rng_seed"])
        assert np.array_equal(tog["guard"] != 0, (st["ante"] > 100) | (st["chips_scored"] > 1000000000))
        assert_records_equal(ov.state, st, L.STATE_DTYPE, (), f"step {t} state")
        dev_obs = v.obs_numpy()
        assert_records_equal(oobs, dev_obs, L.OBS_DTYPE, (), f"step {t} obs")
        # host mirror vs device
        assert_records_equal(dev_obs, m.obs_records(), L.OBS_DTYPE, (), f"step {t} host mirror")
        for name in ("hand", "money", "hand_levels", "phase", "shop_items", "shop_costs", "selected_cards", "action_mask_bits"):
            assert np.array_equal(m.field(name), dev_obs[name]), (t, name)
        assert np.array_equal(m.field("action_mask"), L.mask_from_bits(dev_obs["action_mask_bits"]))
        assert np.array_equal(m.reward.numpy(), v.reward.cpu().numpy()) and np.array_equal(m.terminated.numpy(), v.terminated.cpu().numpy())
        assert np.array_equal(m.terminated.numpy() != 0, oterm != 0)
    if step_path == "multi_pass":
        assert 0 < total_dirty < 0.6 * 90 * n       # deltas, not whole arrays
    else:
        assert total_dirty == 90 * n                # the one-launch step rewrites (and flags) every record
    assert 0 < total_shop < 0.3 * 90 * n, (total_shop, total_dirty)   # envs outside the shop do not resend their (all-zero) shop chunks
