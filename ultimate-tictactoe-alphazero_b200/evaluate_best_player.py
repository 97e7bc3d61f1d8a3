"""Drop-in for the reference's evaluate_best_player.py (vs-random evaluation): EP_GAME_COUNT games of
the best network (temperature 0) against game.random_action, alternating colours, played concurrently through
the engine's batched search (see evaluate_network.py in this directory for the search-semantics note)."""
import numpy as np
import torch

from dual_network import DualNetwork, device
from evaluate_network import NetworkActor, RandomActor, play_matches

EP_GAME_COUNT = 10        # evaluate_best_player.py:21
EP_SEED = None


def evaluate_algorithm_of(label, actors, seed=None):
    """evaluate_best_player.py:53-70"""
    seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
    points = play_matches(actors, EP_GAME_COUNT, seed)
    print('\rEvaluate {}/{}'.format(EP_GAME_COUNT, EP_GAME_COUNT), end='')
    print('')
    average_point = sum(points) / EP_GAME_COUNT
    print(label, average_point)
    return average_point


def evaluate_best_player():
    """evaluate_best_player.py:73-98"""
    model = DualNetwork().to(device)
    model.load_state_dict(torch.load('./model/best.pth', map_location=device, weights_only=True))
    best = NetworkActor(model, 0.0, EP_GAME_COUNT)
    try:
        evaluate_algorithm_of('VS_Random', (best, RandomActor()), EP_SEED)
    finally:
        best.close()
    del model


if __name__ == '__main__':
    evaluate_best_player()
