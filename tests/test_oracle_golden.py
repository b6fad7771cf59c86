"""The C oracle replays the golden traces recorded from the unmodified reference
(tests/golden/make_golden.py) bit for bit: state records, observations, rewards, terminations,
score breakdowns.  CPU only."""
import numpy as np
import pytest

from conftest import load_trace, assert_records_equal, STATE_SKIP, reward_close
from oracle import coracle
from balatro_gym_b200 import layout as L


def replay(trace, stepper):
    """Drive `stepper` (an object with .state/.obs/.reward/.terminated/.info numpy views and
    reset/step) through a [T, E] trace and compare every live step against the reference."""
    T, E = trace["action"].shape
    stepper.reset(trace["seeds"], trace["decks"])
    stepper.set_state(trace["init_state"])           # C3 injection is part of the initial state
    live_steps = 0
    for t in range(T):
        act = trace["action"][t]
        live = act >= 0
        if not live.any():
            break
        state, obs, reward, term, info = stepper.step(act, trace["draws"][t])
        exc = trace["exc"][t].astype(bool)
        ok = live & ~exc
        idx = np.flatnonzero(ok)
        where = f"step {t}"
        assert_records_equal(trace["state"][t][idx], state[idx], L.STATE_DTYPE, STATE_SKIP, where)
        assert_records_equal(trace["obs"][t][idx], obs[idx], L.OBS_DTYPE, (), where)
        assert np.array_equal(trace["term"][t][idx], term[idx]), where
        rc = reward_close(trace["reward"][t][idx], reward[idx], trace["state"][t][idx]["ante"])
        assert rc.all(), (where, trace["reward"][t][idx][~rc], reward[idx][~rc])
        played = trace["info"][t][:, 5].astype(bool) & ok
        p = np.flatnonzero(played)
        assert np.array_equal(trace["info"][t][p, 0], info["final_score"][p]), where
        assert np.array_equal(trace["info"][t][p, 1], info["hand_type"][p]), where
        assert np.array_equal(trace["info"][t][p, 2], info["chips"][p]), where
        assert np.array_equal(trace["info"][t][p, 3], info["mult"][p]), where
        assert np.array_equal(trace["info"][t][idx, 4] != 0, info["error_code"][idx] != 0), where
        # the reference raised: SafeBalatroEnv convention (reward -100, terminated, state unchanged)
        x = np.flatnonzero(live & exc)
        assert (reward[x] == -100.0).all() and (term[x] == 1).all() and (info["error_code"][x] == L.ERR_REF_EXCEPTION).all()
        live_steps += int(live.sum())
    return live_steps


class OracleStepper:
    def __init__(self, E):
        self.v = coracle.OracleVec(E)

    def reset(self, seeds, decks):
        self.v.reset(seeds % (2 ** 32), decks52=decks)

    def set_state(self, init_state):
        keep = {k: self.v.state[k].copy() for k in STATE_SKIP}
        self.v.state[:] = init_state
        for k, val in keep.items():
            self.v.state[k] = val

    def step(self, actions, draws):
        self.v.step(actions, draws=np.ascontiguousarray(draws))
        return self.v.state, self.v.obs, self.v.reward, self.v.terminated, self.v.info


@pytest.mark.parametrize("name", ["c1", "c3", "c4", "c4x"])
def test_oracle_replays_reference_trace(name):
    tr = load_trace(name)
    n = replay(tr, OracleStepper(tr["action"].shape[1]))
    assert n == int(tr["length"].sum())


def consumables_used(tr):
    """(ids whose use succeeded, ids whose use made the reference raise) over a trace: a USE_CONSUMABLE step whose
    pre-step state lists consumable i at the used slot, reported without error / with an exception."""
    T, E = tr["action"].shape
    ok, raised = set(), set()
    for e in range(E):
        prev = tr["init_state"][e]
        for t in range(int(tr["length"][e])):
            a = int(tr["action"][t, e])
            if 10 <= a < 15 and prev["phase"] == L.PHASE_PLAY and a - 10 < prev["cons_n"]:
                cid = int(prev["cons_id"][a - 10])
                if tr["exc"][t, e]:
                    raised.add(cid)
                elif tr["info"][t, e, 4] == 0:
                    ok.add(cid)
            if not tr["exc"][t, e]:
                prev = tr["state"][t, e]
    return ok, raised


def test_c4x_trace_covers_every_consumable():
    """Every consumable id of the reference — tarots 1..22, planets 30..41, spectrals 50..67 and the enum-style
    tarots 101..122 The Emperor creates — is USED in the committed reference trace."""
    tr = load_trace("c4x")
    ok, raised = consumables_used(tr)
    every = set(range(1, 23)) | set(range(30, 42)) | set(range(50, 68)) | set(range(101, 123))
    assert every <= (ok | raised), sorted(every - (ok | raised))
    # where the reference raises: Hanged Man, Familiar, Grim, Incantation with a target; Sigil, Ouija
    assert {13, 113, 50, 51, 52, 56, 57} <= raised
    assert {59, 65} <= ok                                   # Immolate and Cryptid really rebuild the deck list
    st = tr["state"]
    assert int(st["deck_n"].max()) > 52 and int(st["deck_n"][st["deck_n"] > 0].min()) < 50
    assert int(st["deck_extra_n"].max()) == 4
