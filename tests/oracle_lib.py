"""ctypes loaders for the CPU checkers (TEST INFRASTRUCTURE ONLY).

`oracle()`  -> oracle/liboracle.so        (plain-C restatement; built on demand with gcc)
`ref()`     -> oracle/_ref/libref_harness.so (the compiled reference; only where it was built)
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")


class OrcState(C.Structure):
    _fields_ = [("pieces", C.c_int * 81), ("enemy", C.c_int * 81),
                ("main_pieces", C.c_int * 9), ("main_enemy", C.c_int * 9), ("active", C.c_int)]


_oracle = None
_ref = None


def build_oracle():
    src = os.path.join(ORACLE_DIR, "uttt_oracle.c")
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    if (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def oracle():
    global _oracle
    if _oracle is None:
        L = C.CDLL(build_oracle())
        SP = C.POINTER(OrcState)
        L.orc_init.argtypes = [SP]
        L.orc_pack.argtypes = [SP, u32p]
        L.orc_unpack.argtypes = [u32p, SP]
        for f in ("orc_is_lose", "orc_is_draw", "orc_is_done", "orc_is_first_player"):
            getattr(L, f).argtypes = [SP]
            getattr(L, f).restype = C.c_int
        L.orc_next.argtypes = [SP, C.c_int, SP]
        L.orc_legal_actions.argtypes = [SP, i32p]
        L.orc_legal_actions.restype = C.c_int
        L.orc_to_input_tensor.argtypes = [SP, f32p]
        L.orc_to_string.argtypes = [SP, C.c_char_p, C.c_int]
        L.orc_to_string.restype = C.c_int
        L.orc_philox4x32.argtypes = [C.c_uint32] * 6 + [u32p]
        L.orc_state_hash.argtypes = [SP]
        L.orc_state_hash.restype = C.c_uint32
        L.orc_hash_eval.argtypes = [SP, f32p, C.POINTER(C.c_float)]
        L.orc_pv_mcts_scores_hash.argtypes = [SP, C.c_float, C.c_int, C.c_int, f32p, i32p, i32p]
        L.orc_pv_mcts_scores_hash.restype = C.c_int
        L.orc_boltzman.argtypes = [f32p, C.c_int, C.c_float, f32p]
        L.orc_playouts.argtypes = [C.c_uint32, C.c_uint64, C.c_int, u64p, i32p, i32p]
        L.orc_playout.argtypes = [C.c_uint32, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), u8p]
        L.orc_selfplay_hash.argtypes = [C.c_uint32, C.c_uint64, C.c_int, C.c_int, u32p, u16p, u8p, i8p]
        L.orc_selfplay_hash.restype = C.c_int
        L.orc_pv_mcts_scores_table.argtypes = [SP, C.c_float, C.c_int, C.c_int, C.c_int, u32p, f32p, f32p, f32p, i32p,
                                               C.POINTER(C.c_int)]
        L.orc_pv_mcts_scores_table.restype = C.c_int
        L.orc_pv_mcts_scores_hash_record.argtypes = [SP, C.c_float, C.c_int, C.c_int, C.c_int, u32p, f32p, f32p,
                                                     C.POINTER(C.c_int), f32p]
        L.orc_pv_mcts_scores_hash_record.restype = C.c_int
        L.orc_py_mcts_counts_hash.argtypes = [SP, C.c_int, C.c_int, i32p]
        L.orc_py_mcts_counts_hash.restype = C.c_int
        L.orc_py_mcts_counts_table.argtypes = [SP, C.c_int, C.c_int, C.c_int, u32p, f32p, f32p, i32p, C.POINTER(C.c_int)]
        L.orc_py_mcts_counts_table.restype = C.c_int
        L.orc_az_search_hash.argtypes = [SP, C.c_int, i32p]
        L.orc_az_search_hash.restype = C.c_int
        _oracle = L
    return _oracle


def ref_available():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libref_harness.so"))


def ref():
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libref_harness.so"))
        L.ref_playouts.argtypes = [C.c_uint32, C.c_uint64, C.c_int, u64p, i32p, i32p]
        L.ref_state_probe.argtypes = [u32p, C.POINTER(C.c_int), C.POINTER(C.c_int), i32p, f32p]
        L.ref_state_next.argtypes = [u32p, C.c_int, u32p]
        L.ref_state_to_string.argtypes = [u32p, C.c_char_p, C.c_int]
        L.ref_state_to_string.restype = C.c_int
        L.ref_mcts_scores_hash.argtypes = [u32p, C.c_float, C.c_int, C.c_int, f32p, i32p]
        L.ref_mcts_scores_hash.restype = C.c_int
        L.ref_mcts_scores_table.argtypes = [u32p, C.c_float, C.c_int, C.c_int, C.c_int, u32p, f32p, f32p, f32p,
                                            C.POINTER(C.c_int)]
        L.ref_mcts_scores_table.restype = C.c_int
        L.ref_boltzman.argtypes = [f32p, C.c_int, C.c_float, f32p]
        L.ref_selfplay_hash.argtypes = [C.c_uint32, C.c_uint64, C.c_int, C.c_int, u32p, u16p, u8p, i8p]
        L.ref_selfplay_hash.restype = C.c_int
        L.ref_selfplay_null.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int]
        L.ref_selfplay_null.restype = C.c_long
        _ref = L
    return _ref


# ---------------------------------------------------------------- helpers
def state_from_packed(w):
    s = OrcState()
    oracle().orc_unpack(np.ascontiguousarray(w, dtype=np.uint32), C.byref(s))
    return s


def packed_from_state(s):
    w = np.zeros(8, np.uint32)
    oracle().orc_pack(C.byref(s), w)
    return w


def playout_states(seed, game):
    """All packed states visited by Philox playout `game` (incl. the terminal one)."""
    L = oracle()
    acts = np.zeros(81, np.uint8)
    dg, pl, rs = C.c_uint64(), C.c_int(), C.c_int()
    L.orc_playout(seed, game, C.byref(dg), C.byref(pl), C.byref(rs), acts)
    s = OrcState()
    L.orc_init(C.byref(s))
    out = [packed_from_state(s)]
    for t in range(pl.value):
        n = OrcState()
        L.orc_next(C.byref(s), int(acts[t]), C.byref(n))
        s = n
        out.append(packed_from_state(s))
    return np.stack(out), acts[:pl.value].copy()


def oracle_probe(w):
    """(flags, legal list, tensor[243]) for one packed state via the oracle."""
    L = oracle()
    s = state_from_packed(w)
    flags = (L.orc_is_lose(C.byref(s)) | (L.orc_is_draw(C.byref(s)) << 1) |
             (L.orc_is_done(C.byref(s)) << 2) | (L.orc_is_first_player(C.byref(s)) << 3))
    legal = np.zeros(81, np.int32)
    n = L.orc_legal_actions(C.byref(s), legal)
    t = np.zeros(243, np.float32)
    L.orc_to_input_tensor(C.byref(s), t)
    return flags, legal[:n].copy(), t


def oracle_mcts(w, temperature, sims, batch):
    L = oracle()
    s = state_from_packed(w)
    sc = np.zeros(81, np.float32)
    cn = np.zeros(81, np.int32)
    st = np.zeros(3, np.int32)
    n = L.orc_pv_mcts_scores_hash(C.byref(s), temperature, sims, batch, sc, cn, st)
    return sc[:n].copy(), cn[:n].copy(), st


def ref_mcts(w, temperature, sims, batch):
    sc = np.zeros(81, np.float32)
    st = np.zeros(2, np.int32)
    n = ref().ref_mcts_scores_hash(np.ascontiguousarray(w, dtype=np.uint32), temperature, sims, batch, sc, st)
    return sc[:n].copy(), st


def table_mcts(w, temperature, sims, batch, states, policy, value, use_ref=None):
    """The reference search (compiled reference if it was built here, else the C restatement) fed recorded rows
    (leaf state -> policy[81], value), consumed in order.  -> (scores, leaves without a row, rows left unused)"""
    states = np.ascontiguousarray(states, dtype=np.uint32).reshape(-1, 8)
    policy = np.ascontiguousarray(policy, dtype=np.float32).reshape(-1, 81)
    value = np.ascontiguousarray(value, dtype=np.float32).reshape(-1)
    n = len(value)
    if n == 0:      # ndpointer rejects empty arrays of the wrong shape: pass one dummy row that is never matched
        states, policy, value = np.full((1, 8), 0xFFFFFFFF, np.uint32), np.zeros((1, 81), np.float32), np.zeros(1, np.float32)
    sc = np.zeros(81, np.float32)
    miss = C.c_int(0)
    if use_ref is None:
        use_ref = ref_available()
    if use_ref:
        m = ref().ref_mcts_scores_table(np.ascontiguousarray(w, dtype=np.uint32), temperature, sims, batch, n, states, policy,
                                        value, sc, C.byref(miss))
    else:
        s = state_from_packed(w)
        cn = np.zeros(81, np.int32)
        m = oracle().orc_pv_mcts_scores_table(C.byref(s), temperature, sims, batch, n, states, policy, value, sc, cn,
                                              C.byref(miss))
    return sc[:m].copy(), miss.value & 0xFFFF, miss.value >> 16


def record_hash_mcts(w, temperature, sims, batch, cap=4096):
    """oracle search under the hash evaluator + the rows it evaluated, in order"""
    s = state_from_packed(w)
    st = np.zeros((cap, 8), np.uint32)
    pol = np.zeros((cap, 81), np.float32)
    val = np.zeros(cap, np.float32)
    n = C.c_int(0)
    sc = np.zeros(81, np.float32)
    m = oracle().orc_pv_mcts_scores_hash_record(C.byref(s), temperature, sims, batch, cap, st, pol, val, C.byref(n), sc)
    assert n.value <= cap
    return sc[:m].copy(), st[:n.value], pol[:n.value], val[:n.value]


def oracle_py_mcts(w, sims, batch):
    """root visit counts of the reference's pure-Python search (pv_mcts.py:74-180) under the hash evaluator"""
    s = state_from_packed(w)
    cn = np.zeros(81, np.int32)
    n = oracle().orc_py_mcts_counts_hash(C.byref(s), sims, batch, cn)
    return cn[:n].copy()


def table_py_mcts(w, sims, batch, states, policy, value):
    """the Python-semantics search (pv_mcts.py:74-180 restated, pinned by tests/golden/pymcts.npz) fed recorded rows
    -> (root visit counts, leaves without a row, rows left unused)"""
    states = np.ascontiguousarray(states, dtype=np.uint32).reshape(-1, 8)
    policy = np.ascontiguousarray(policy, dtype=np.float32).reshape(-1, 81)
    value = np.ascontiguousarray(value, dtype=np.float32).reshape(-1)
    n = len(value)
    if n == 0:
        states, policy, value = np.full((1, 8), 0xFFFFFFFF, np.uint32), np.zeros((1, 81), np.float32), np.zeros(1, np.float32)
    s = state_from_packed(w)
    cn = np.zeros(81, np.int32)
    miss = C.c_int(0)
    m = oracle().orc_py_mcts_counts_table(C.byref(s), sims, batch, n, states, policy, value, cn, C.byref(miss))
    return cn[:m].copy(), miss.value & 0xFFFF, miss.value >> 16


def oracle_az_search(w, sims):
    s = state_from_packed(w)
    cn = np.zeros(81, np.int32)
    n = oracle().orc_az_search_hash(C.byref(s), sims, cn)
    return cn[:n].copy()
