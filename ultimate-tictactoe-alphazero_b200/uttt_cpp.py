"""Drop-in replacement for the reference's pybind11 module `uttt_cpp`
(cpp/python_bindings.cpp:49-107): same names, argument meaning and error behaviour, backed by
libuttt_b200.so.  `State` is an immutable value object over the 32-byte packed position;
`pv_mcts_scores` runs the search on the GPU (tree kernels), with the network forward on the GPU
too when `model` is a DualNetwork, or with the caller's callable as evaluator otherwise.
"""
import ctypes as C
import weakref

import numpy as np

import engine as _eng

_lib = _eng.load_library()          # ImportError if the CUDA library is missing: no fallback


class InferenceResult:
    """cpp/python_bindings.cpp:77-80"""

    def __init__(self, policy=None, value=0.0):
        self.policy = list(policy) if policy is not None else []
        self.value = float(value)


class State:
    """cpp/python_bindings.cpp:54-74 over the packed position (csrc/uttt_rules.cuh)."""
    __slots__ = ("_w",)

    def __init__(self, pieces=None, enemy_pieces=None, main_board_pieces=None, main_board_enemy_pieces=None,
                 active_board=-1):
        w = np.zeros(8, np.uint32)
        if pieces is not None:
            if enemy_pieces is None or main_board_pieces is None or main_board_enemy_pieces is None:
                raise TypeError("State() takes either no arguments or (pieces, enemy_pieces, main_board_pieces, "
                                "main_board_enemy_pieces, active_board)")
            p = np.asarray(pieces, dtype=np.int64).reshape(9, 9)
            e = np.asarray(enemy_pieces, dtype=np.int64).reshape(9, 9)
            mp = np.asarray(main_board_pieces, dtype=np.int64).reshape(9)
            me = np.asarray(main_board_enemy_pieces, dtype=np.int64).reshape(9)
            for arr in (p, e, mp, me):
                if ((arr != 0) & (arr != 1)).any():
                    raise ValueError("State arrays must contain only 0/1")
            act = int(active_board)
            if not -1 <= act <= 8:
                raise ValueError("active_board must be in [-1, 8]")
            for b in range(9):
                sh = 9 * (b % 3)
                w[b // 3] |= np.uint32(int(np.dot(p[b], 1 << np.arange(9))) << sh)
                w[3 + b // 3] |= np.uint32(int(np.dot(e[b], 1 << np.arange(9))) << sh)
            w[6] = np.uint32(int(np.dot(mp, 1 << np.arange(9))) | (int(np.dot(me, 1 << np.arange(9))) << 9)
                             | ((act + 1) << 18))
        self._w = w

    @classmethod
    def _from_packed(cls, w):
        s = cls.__new__(cls)
        s._w = np.array(w, dtype=np.uint32, copy=True).reshape(8)
        return s

    def packed(self):
        return self._w.copy()

    def _flags(self):
        f = C.c_int()
        _eng._check(_lib.uttt_state_flags(self._w, C.byref(f)))
        return f.value

    def is_lose(self):
        return bool(self._flags() & 1)

    def is_draw(self):
        return bool(self._flags() & 2)

    def is_done(self):
        return bool(self._flags() & 4)

    def is_first_player(self):
        return bool(self._flags() & 8)

    def next(self, action):
        out = np.zeros(8, np.uint32)
        _eng._check(_lib.uttt_state_next(self._w, int(action), out))
        return State._from_packed(out)

    def legal_actions(self):
        out = np.zeros(81, np.int32)
        n = C.c_int()
        _eng._check(_lib.uttt_state_legal_actions(self._w, out, C.byref(n)))
        return out[:n.value].tolist()

    def to_input_tensor(self):
        out = np.zeros(243, np.float32)
        _eng._check(_lib.uttt_state_encode(self._w, out))
        return out.tolist()

    def to_string(self):
        buf = C.create_string_buffer(1024)
        n = C.c_int()
        _eng._check(_lib.uttt_state_to_string(self._w, buf, 1024, C.byref(n)))
        return buf.value.decode()

    __str__ = to_string

    def _cells(self, base):
        return [[int((self._w[base + b // 3] >> (9 * (b % 3) + c)) & 1) for c in range(9)] for b in range(9)]

    @property
    def pieces(self):
        return self._cells(0)

    @property
    def enemy_pieces(self):
        return self._cells(3)

    @property
    def main_board_pieces(self):
        return [int((self._w[6] >> b) & 1) for b in range(9)]

    @property
    def main_board_enemy_pieces(self):
        return [int((self._w[6] >> (9 + b)) & 1) for b in range(9)]

    @property
    def active_board(self):
        return int((self._w[6] >> 18) & 15) - 1


# ------------------------------------------------------------------------------------------ search
_engine = None
_uploaded_key = None
_uploaded_model = None           # weakref to the module whose weights the engine holds
NUMERICS = _eng.DEFAULT_NUMERICS   # "bf16x3" (default: split-bf16 tcgen05 trunk, within 1e-2 of the fp32 reference forward),
                                 # "bf16" (plain bf16 operands, 3x faster), "fp32" (CUDA cores)


def _get_engine(sims, batch):
    global _engine, _uploaded_key, _uploaded_model
    if _engine is None or _engine.max_sims < sims or _engine.max_batch < batch:
        if _engine is not None:
            _engine.close()
        _engine = _eng.Engine(n_slots=64, max_sims=max(sims, 50), max_batch=max(batch, 8), max_games=64)
        _uploaded_key = None
        _uploaded_model = None
    return _engine


def _sync_weights(e, model):
    """upload the module's weights unless the engine provably holds them already: the SAME live module object (a weak
    reference: a new module that re-uses a dead one's address does not match) whose tensors have the same storage,
    the same in-place version counters and the same content fingerprint of four of them"""
    global _uploaded_key, _uploaded_model
    import torch
    sd = model.state_dict()
    tensors = list(sd.values())
    with torch.no_grad():          # a cheap content fingerprint of four tensors (writes through .data bump no version counter)
        probe = [sd[k] for k in ("conv_input.weight", "residual_blocks.7.conv2.weight", "policy_fc.weight", "value_fc2.weight")
                 if k in sd]
        fp = torch.stack([t.reshape(-1)[:: max(1, t.numel() // 7)].double().sum() for t in probe]).cpu().numpy().tobytes()
    key = (id(e), fp) + tuple((t.data_ptr(), t._version) for t in tensors)
    if _uploaded_model is None or _uploaded_model() is not model or key != _uploaded_key:
        e.upload_model(model)
        _uploaded_key = key
        _uploaded_model = weakref.ref(model)


def _is_network(model):
    try:
        import torch.nn as nn
        return isinstance(model, nn.Module) and hasattr(model, "residual_blocks") and hasattr(model, "policy_fc")
    except ImportError:
        return False


def pv_mcts_scores(model, state, temperature=0.0, evaluate_count=50, batch_size=8):
    """cpp/python_bindings.cpp:83-100 -> UTTT::pv_mcts_scores (cpp/uttt_mcts.cpp:84-196).

    model: a DualNetwork (evaluated on the GPU by the engine) or any callable
           list[State] -> iterable of (policy[81], value) exactly like the reference's callback.
    Returns list[float], one score per legal action of `state` in ascending action id."""
    if not isinstance(state, State):
        raise TypeError("state must be a uttt_cpp.State")
    evaluate_count, batch_size = int(evaluate_count), int(batch_size)
    e = _get_engine(evaluate_count, batch_size)
    roots = state._w.reshape(1, 8)
    if _is_network(model):
        _sync_weights(e, model)
        ev = _eng.evaluator_of(NUMERICS)
        scores, _, ns = e.mcts_search(roots, evaluate_count, batch_size, temperature, ev)
        return scores[0, :ns[0]].tolist()

    mb = e.max_batch

    def eval_fn(st, k):
        pol = np.zeros((len(st), mb, 81), np.float32)
        val = np.zeros((len(st), mb), np.float32)
        for i in range(len(st)):
            batch = [State._from_packed(st[i]) for _ in range(int(k[i]))]   # k queued copies (Q-M3)
            results = list(model(batch))
            if len(results) < len(batch):
                raise ValueError("inference callback returned %d results for %d states" % (len(results), len(batch)))
            for c, (p, v) in enumerate(results[:len(batch)]):
                p = np.asarray(p, dtype=np.float32).reshape(-1)
                pol[i, c, :min(81, p.size)] = p[:81]
                val[i, c] = float(v)
        return pol, val
    scores, _, ns = e.mcts_search_host(roots, evaluate_count, batch_size, temperature, eval_fn, per_copy=True)
    return scores[0, :ns[0]].tolist()


def boltzman(xs, temperature):
    """cpp/python_bindings.cpp:102-106 -> UTTT::boltzman (cpp/uttt_mcts.cpp:199-216)"""
    xs = np.ascontiguousarray(xs, dtype=np.float32).reshape(-1)
    out = np.zeros_like(xs)
    _eng._check(_lib.uttt_boltzman(xs, len(xs), float(temperature), out))
    return out.tolist()
