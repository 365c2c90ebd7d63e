"""GPU parity: bitboard game kernels (through the C ABI) vs the array-board oracle — bit-exact."""
import numpy as np
import pytest

import selfplay_b200 as S
from helpers import random_states
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[S.GAME_TTT, S.GAME_C4], ids=["ttt", "c4"])
def eng(request):
    e = S.Engine(game=request.param, num_games=4, evaluator=S.EVAL_DET)
    yield e
    e.close()


def test_rules_bit_exact_on_random_positions(eng):
    game = eng.game
    A = S.NUM_ACTIONS[game]
    n = 60000 if game == S.GAME_C4 else 20000
    states = random_states(game, n, seed=11 + game)
    rng = np.random.default_rng(5)
    actions = rng.integers(0, A + 1, size=n).astype(np.uint8)      # A itself is out of range -> illegal
    arr = O.states_array(states)
    out, err = eng.game_next_states(arr, actions)
    masks = eng.game_valid_actions(arr)
    enc = eng.game_encode(arr[:4000])
    n_illegal = 0
    for i, s in enumerate(states):
        want = O.next_state(game, s, int(actions[i])) if actions[i] < A else None
        if want is None:
            assert err[i] == -4, (i, s, actions[i])
            n_illegal += 1
        else:
            assert err[i] == 0
            got = O.state_from_record(out[i])
            assert got.key() == want.key(), (i, s, actions[i], got, want)
        va = O.valid_actions(game, s)
        assert masks[i] == sum(1 << a for a in va)
        if i < 4000:
            assert np.array_equal(enc[i], O.encode(game, s))
    assert n_illegal > n // 20


def test_connect4_antidiagonal_quirk_on_device():
    with S.Engine(game=S.GAME_C4, num_games=1, evaluator=S.EVAL_DET) as e:
        s = O.State()
        for a in [3, 2, 2, 1, 1, 0, 1, 0, 0, 6]:
            s = O.next_state(O.GAME_C4, s, a)
        out, err = e.game_next_states([s], [0])
        assert err[0] == 0 and out[0]["status"] == S.ONGOING       # four on the anti-diagonal is NOT a win
        s2 = O.State()
        for a in [0, 1, 1, 2, 2, 3, 2, 3, 3, 6]:
            s2 = O.next_state(O.GAME_C4, s2, a)
        out, err = e.game_next_states([s2], [3])
        assert err[0] == 0 and out[0]["status"] == S.WON           # main diagonal is


def test_synthetic_roots_generated_on_device_match_oracle():
    """bench.py builds its roots with the device rules; the reference arm builds them with the oracle."""
    import helpers
    from selfplay_b200.synth import synthetic_roots_device
    with S.Engine(game=S.GAME_C4, num_games=4, evaluator=S.EVAL_DET) as e:
        got = synthetic_roots_device(e, 300)
    want = O.states_array(helpers.synthetic_roots(O.GAME_C4, 300))
    assert got.tobytes() == want.tobytes()
