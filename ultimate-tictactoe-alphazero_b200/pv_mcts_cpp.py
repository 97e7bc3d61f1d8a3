"""Drop-in replacement for the reference's pv_mcts_cpp.py (same public names and defaults):
pv_mcts_scores_cpp / pv_mcts_action_cpp / check_cpp_compatibility / CPP_AVAILABLE.

The reference builds a Python inference closure (state -> tensor -> model -> numpy,
pv_mcts_cpp.py:37-78) and hands it to the C++ search.  Here a DualNetwork `model` never leaves the
GPU: leaves are gathered, evaluated (tcgen05 trunk + heads kernels) and backed up on the device.
Any other callable is honoured as a leaf evaluator with the reference's callback contract.
"""
import numpy as np

import uttt_cpp

CPP_AVAILABLE = True      # importing uttt_cpp above raises if the CUDA library is missing


def pv_mcts_scores_cpp(model, state, temperature, evaluate_count=50, batch_size=8):
    """pv_mcts_cpp.py:17-89 -> np.ndarray (float64) of scores over state.legal_actions()"""
    if hasattr(model, "eval"):
        model.eval()
    scores = uttt_cpp.pv_mcts_scores(model=model, state=state, temperature=temperature,
                                     evaluate_count=evaluate_count, batch_size=batch_size)
    return np.array(scores)


def _as_cpp_state(state):
    """pv_mcts_cpp.py:107-117: accept the Python game.State duck type as well"""
    if isinstance(state, uttt_cpp.State):
        return state
    return uttt_cpp.State(state.pieces, state.enemy_pieces, state.main_board_pieces,
                          state.main_board_enemy_pieces, state.active_board)


def pv_mcts_action_cpp(model, temperature=0, evaluate_count=50, batch_size=8):
    """pv_mcts_cpp.py:92-137 -> callable(state) -> action"""
    def action_func(state):
        cpp_state = _as_cpp_state(state)
        scores = pv_mcts_scores_cpp(model, cpp_state, temperature, evaluate_count, batch_size)
        legal_actions = cpp_state.legal_actions()
        if len(scores) != len(legal_actions):
            raise ValueError(f"Score size mismatch: scores={len(scores)}, legal_actions={len(legal_actions)}")
        total = np.sum(scores)
        scores = np.ones(len(scores)) / len(scores) if total == 0 else scores / total
        return np.random.choice(legal_actions, p=scores)
    return action_func


def check_cpp_compatibility():
    """pv_mcts_cpp.py:140-167 smoke check of the rules API"""
    try:
        state = uttt_cpp.State()
        legal_actions = state.legal_actions()
        print(f"OK game logic working (legal actions: {len(legal_actions)})")
        if legal_actions:
            state.next(legal_actions[0])
            print("OK state transition working")
        print(f"OK tensor conversion working (shape: {len(state.to_input_tensor())})")
        return True
    except Exception as e:  # noqa: BLE001 - mirrors the reference's catch-all smoke test
        print(f"XX uttt_cpp (B200) module test failed: {e}")
        return False


if __name__ == "__main__":
    check_cpp_compatibility()
