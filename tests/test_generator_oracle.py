"""CPU checks of the synthetic-state generator (BGYM_FLAG_GEN_C3 / BGYM_FLAG_GEN_CONS, include/bgym.h) as the
C oracle states it: the law SURVEY 8(d) C3/C4 asks for, survival across autoresets, and the untouched default."""
import numpy as np

from balatro_gym_b200 import layout as L
from oracle import coracle


def test_generator_law_on_the_oracle():
    n = 20000
    ov = coracle.OracleVec(n)
    coracle.reset(ov.state, ov.obs, np.arange(1, n + 1), flags=L.GENERATORS["c4"])
    st = ov.state
    jk = st["joker_id"][:, :5].astype(int)
    assert (st["joker_n"] == 5).all() and jk.min() >= 1 and jk.max() <= 145
    assert (np.diff(np.sort(jk, axis=1), axis=1) > 0).all()
    deck = st["deck"].astype(int)
    assert (np.sort(deck & 63, axis=1) == np.arange(52)).all()
    N = n * 52
    for arr, p, k in (((deck >> 6) & 15, 0.25, 8), ((deck >> 10) & 7, 0.1, 3), ((deck >> 13) & 7, 0.1, 4)):
        assert arr.max() == k and abs((arr != 0).mean() - p) < 5 * (p * (1 - p) / N) ** 0.5
    assert (st["cons_n"] == 2).all()
    ids = list(range(1, 23)) + list(range(30, 42)) + list(range(50, 68))
    assert set(np.unique(st["cons_id"][:, :2])) == set(ids)
    assert (ov.obs["joker_count"] == 5).all() and (ov.obs["consumable_count"] == 2).all()
    # without the flags nothing is generated
    ov0 = coracle.OracleVec(64)
    ov0.reset(np.arange(1, 65))
    assert (ov0.state["joker_n"] == 0).all() and ((ov0.state["deck"] >> 6) == 0).all() and (ov0.state["cons_n"] == 0).all()
    # c3 = no consumables, same jokers and card modifiers
    ov3 = coracle.OracleVec(64)
    coracle.reset(ov3.state, ov3.obs, np.arange(1, 65), flags=L.GENERATORS["c3"])
    assert (ov3.state["cons_n"] == 0).all()
    assert (ov3.state["deck"] == st["deck"][:64]).all() and (ov3.state["joker_id"] == st["joker_id"][:64]).all()


def test_generator_state_survives_autoreset_on_the_oracle():
    n = 512
    gf = L.GENERATORS["c4"]
    ov = coracle.OracleVec(n)
    coracle.reset(ov.state, ov.obs, np.arange(1, n + 1), flags=gf)
    act = np.zeros(n, np.int32)
    for _ in range(400):
        coracle.step(ov.state, act, ov.obs, ov.reward, ov.terminated, ov.truncated, ov.info, None,
                     flags=L.FLAG_AUTORESET | L.FLAG_RANDOM_POLICY | gf)
    st = ov.state
    assert int(st["episode"].sum()) > n            # every env has been through several episodes
    assert ((st["deck"] >> 6) != 0).any(axis=1).all()
    fresh = st["ep_len"] == 0                      # envs reset by the last step: untouched generator output
    assert fresh.any() and (st["joker_n"][fresh] == 5).all() and (st["cons_n"][fresh] == 2).all()
