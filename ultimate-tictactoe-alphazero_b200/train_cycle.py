"""Drop-in for the reference's train_cycle.py (the caller of the hot path): the same four stages per iteration
through the same module and function names -- dual_network.dual_network(), self_play_cpp.self_play(),
train_network.train_network(), evaluate_network.evaluate_network(), evaluate_best_player.evaluate_best_player() -- so the
reference's own driver also runs unmodified with this directory first on sys.path (tests/test_abi_cpu.py checks every
name its import block needs, train_cycle.py:6-18 there).  There is no hybrid / pure-Python fallback: importing uttt_cpp
fails loudly when the CUDA library or an sm_100 GPU is missing.
"""
from dual_network import dual_network

import uttt_cpp  # noqa: F401  (ImportError without the CUDA library: no fallback)
from self_play_cpp import self_play
from train_network import train_network
from evaluate_network import evaluate_network
from evaluate_best_player import evaluate_best_player

TC_CYCLES = 10            # train_cycle.py:24


def train_cycle(cycles=None):
    """train_cycle.py:20-41"""
    dual_network()                                   # creates ./model/best.pth if it does not exist
    for i in range(TC_CYCLES if cycles is None else cycles):
        print('Train', i, '====================')
        self_play()
        print(f'>> Train {i}')
        train_network()
        update_best_player = evaluate_network()
        if update_best_player:
            evaluate_best_player()


if __name__ == '__main__':
    print(">> Using B200 CUDA backend")
    train_cycle()
