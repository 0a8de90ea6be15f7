"""ctypes front-end of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE.

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU legs;
nothing under dismember_b200/ may import it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h", ".inc"))]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    vp = C.c_void_p
    L.orc_tree_create.restype = vp
    L.orc_tree_create.argtypes = [C.c_int, C.c_int64, i32p, i32p, u8p, C.c_int64, i32p, i32p]
    L.orc_tree_destroy.argtypes = [vp]
    L.orc_tdm_id_to_code.argtypes = [vp, C.c_int, i32p, i32p, u8p]
    L.orc_tdm_model_create.restype = vp
    L.orc_tdm_model_create.argtypes = [C.c_int64, C.c_int, C.c_int, f32p]
    L.orc_tdm_deepfm_create.restype = vp
    L.orc_tdm_deepfm_create.argtypes = [C.c_int64, C.c_int, C.c_int, f32p]
    L.orc_tdm_model_destroy.argtypes = [vp]
    L.orc_otm_model_create.restype = vp
    L.orc_otm_model_create.argtypes = [C.c_int64, C.c_int, C.c_int, f64p]
    L.orc_otm_deepfm_create.restype = vp
    L.orc_otm_deepfm_create.argtypes = [C.c_int64, C.c_int, C.c_int, f64p]
    L.orc_otm_model_destroy.argtypes = [vp]
    L.orc_din_forward_f32_api.argtypes = [vp, C.c_int64, i32p, i32p, vp, C.c_int64, f32p]
    L.orc_din_forward_f64_api.argtypes = [vp, C.c_int64, i32p, i32p, vp, C.c_int64, f64p]
    L.orc_tdm_recommend_raw.argtypes = [vp, vp, i32p, C.c_int, C.c_int, vp, C.c_int, i32p, f32p, C.c_int]
    L.orc_tdm_recommend.argtypes = [vp, vp, i32p, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int,
                                    i32p, f32p, f64p]
    L.orc_tdm_retrieve_batch.argtypes = [vp, vp, C.c_int, i32p, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int,
                                         C.c_int, i32p, f32p, i32p]
    L.orc_tuned_create.restype = vp
    L.orc_tuned_create.argtypes = [vp, vp, C.c_int]
    L.orc_tuned_destroy.argtypes = [vp]
    L.orc_tuned_retrieve_batch.argtypes = [vp, C.c_int, i32p, C.c_int, C.c_int, C.c_int, C.c_int, i32p, f32p, i32p]
    L.orc_otm_beam_search.argtypes = [vp, i32p, C.c_int, C.c_int, C.c_int, i32p, f64p]
    L.orc_otm_recommend.argtypes = [vp, i32p, C.c_int, C.c_int, C.c_int, C.c_int, i32p, i32p, f64p, f64p]
    L.orc_otm_retrieve_batch.argtypes = [vp, C.c_int, i32p, C.c_int, C.c_int, C.c_int, C.c_int, i32p, C.c_int,
                                         i32p, f64p, i32p]
    L.orc_otm_pseudo_targets.argtypes = [vp, C.c_int, C.c_int, i32p, np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS"), i32p, C.c_int, C.c_int,
                                         C.c_int, C.c_int, i32p, f64p, i32p]
    L.orc_dr_model_create.restype = vp
    L.orc_dr_model_create.argtypes = [C.c_int] * 5 + [f64p, C.POINTER(vp), C.POINTER(vp), f64p, f64p, f64p, f64p, f64p]
    L.orc_dr_model_destroy.argtypes = [vp]
    L.orc_dr_beam_search.argtypes = [vp, i32p, C.c_int, i32p, f64p]
    L.orc_dr_rerank.argtypes = [vp, i32p, C.c_int, i32p, f64p]
    L.orc_dr_recommend.argtypes = [vp, i32p, C.c_int, C.c_int, i64p, i32p, i32p, f64p, f64p]
    L.orc_din_gradients_f32.argtypes = [C.c_int64, C.c_int, C.c_int, f32p, C.c_int64, i32p, i32p, vp, C.c_int64, f32p, f32p, f32p]
    L.orc_din_gradients_f64.argtypes = [C.c_int64, C.c_int, C.c_int, f64p, C.c_int64, i32p, i32p, vp, C.c_int64, f64p, f64p, f64p]
    L.orc_deepfm_gradients_f32.argtypes = [C.c_int64, C.c_int, C.c_int, f32p, C.c_int64, i32p, i32p, f32p, f32p, f32p]
    L.orc_adam_f32.argtypes = [f32p, f32p, f32p, f32p, C.c_int64, C.c_double, C.c_int]
    L.orc_adam_f32.restype = None
    L.orc_adam_f64.argtypes = [f64p, f64p, f64p, f64p, C.c_int64, C.c_double, C.c_int]
    L.orc_adam_f64.restype = None
    L.orc_cross_entropy_f64.argtypes = [C.c_int64, C.c_int, f64p, i32p, vp]
    L.orc_cross_entropy_f64.restype = C.c_double
    L.orc_dr_layer_grad.argtypes = [C.c_int] * 5 + [f64p, C.POINTER(vp), C.POINTER(vp), C.c_int, i32p, i32p, i32p, C.c_int, C.c_int,
                                    f64p, C.POINTER(vp), C.POINTER(vp), f64p]
    L.orc_sampled_softmax_f64.argtypes = [C.c_int, C.c_int, C.c_int, f64p, f64p, f64p, i32p, f64p, f64p, f64p]
    L.orc_sampled_softmax_f64.restype = C.c_double
    L.orc_dr_rerank_grad.argtypes = [C.c_int, C.c_int, C.c_int, f64p, f64p, f64p, f64p, f64p, C.c_int, i32p, i32p, C.c_int,
                                     f64p, f64p, f64p, f64p, f64p, f64p]
    L.orc_adam_eps_f64.argtypes = [f64p, f64p, f64p, f64p, C.c_int64, C.c_double, C.c_double, C.c_int]
    L.orc_adam_eps_f64.restype = None
    L.orc_arg_partition.argtypes = [f64p, C.c_int, C.c_int, i32p]
    L.orc_kmeans_tree.argtypes = [C.c_int, C.c_int, f64p, C.c_int, C.c_uint64, i32p]
    L.orc_softmax_f32.argtypes = [C.c_int, C.c_int, f32p, f32p]
    L.orc_softmax_grad_f32.argtypes = [C.c_int, C.c_int, f32p, f32p, f32p]
    L.orc_expf_api.restype = C.c_float
    L.orc_expf_api.argtypes = [C.c_float]
    L.orc_exp_api.restype = C.c_double
    L.orc_exp_api.argtypes = [C.c_double]
    _lib = L
    return L


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _ci32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Tree:
    """DistTree maps as flat arrays (see dismember_b200.formats.tree_file.TreeFile)."""

    def __init__(self, max_level, codes, node_ids, is_leaf, leaf_ids, leaf_codes):
        self.max_level = int(max_level)
        self._keep = [_ci32(codes), _ci32(node_ids), np.ascontiguousarray(is_leaf, np.uint8),
                      _ci32(leaf_ids), _ci32(leaf_codes)]
        k = self._keep
        self.h = lib().orc_tree_create(self.max_level, len(k[0]), k[0], k[1], k[2], len(k[3]), k[3], k[4])
        if not self.h:
            raise ValueError("orc_tree_create failed (code out of range)")

    @classmethod
    def from_treefile(cls, tf):
        return cls(tf.max_level, tf.codes, tf.node_ids, tf.is_leaf, tf.leaf_ids, tf.leaf_codes)

    def id_to_code(self, ids):
        ids = _ci32(ids)
        codes = np.empty(len(ids), np.int32)
        masked = np.empty(len(ids), np.uint8)
        lib().orc_tdm_id_to_code(self.h, len(ids), ids, codes, masked)
        return codes, masked

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_tree_destroy(self.h)
            self.h = None


class TdmModel:
    """DIN(Float) from the compact parameter vector."""

    def __init__(self, params, rows, E, T, deepfm=False):
        self.params = np.ascontiguousarray(params, np.float32)
        self.rows, self.E, self.T = int(rows), int(E), int(T)
        if deepfm:                                   # tdm/.../model/DeepFM.scala: [emb | W1 (T+1)x(T+1)E | b1 | W2 | b2]
            assert self.params.size == rows * E + (T + 1) * (T + 1) * E + 2 * (T + 1) + 1, "bad DeepFM parameter count"
            self.h = lib().orc_tdm_deepfm_create(self.rows, self.E, self.T, self.params)
        else:
            assert self.params.size == rows * E + E * E + 2 * E * E + 2 * E + 1, "bad DIN parameter count"
            self.h = lib().orc_tdm_model_create(self.rows, self.E, self.T, self.params)

    def forward(self, node, seq, mask_flat=None):
        node = _ci32(node).ravel()
        seq = _ci32(seq).reshape(len(node), self.T)
        out = np.empty(len(node), np.float32)
        m = None if mask_flat is None else _ci32(mask_flat)
        rc = lib().orc_din_forward_f32_api(self.h, len(node), node, seq, _ptr(m), 0 if m is None else len(m), out)
        if rc:
            raise IndexError("embeddingLookup failed: index out of range")
        return out

    def recommend_raw(self, tree: Tree, seq_ids, beam, use_mask=True, consumed=None):
        seq_ids = _ci32(seq_ids)
        cap = 2 * beam * (tree.max_level + 2) + 8
        items = np.empty(cap, np.int32)
        logits = np.empty(cap, np.float32)
        cons = None if consumed is None else _ci32(consumed)
        n = lib().orc_tdm_recommend_raw(tree.h, self.h, seq_ids, beam, int(use_mask), _ptr(cons),
                                        0 if cons is None else len(cons), items, logits, cap)
        if n < 0:
            raise IndexError(f"oracle error {n}")
        return items[:n].copy(), logits[:n].copy()

    def recommend(self, tree: Tree, seq_ids, topk, beam, use_mask=True, consumed=None, widen_beam=False):
        seq_ids = _ci32(seq_ids)
        items = np.empty(topk, np.int32)
        logits = np.empty(topk, np.float32)
        prob = np.empty(topk, np.float64)
        cons = None if consumed is None else _ci32(consumed)
        n = lib().orc_tdm_recommend(tree.h, self.h, seq_ids, beam, topk, int(use_mask), _ptr(cons),
                                    0 if cons is None else len(cons), int(widen_beam), items, logits, prob)
        if n < 0:
            raise IndexError(f"oracle error {n}")
        return items[:n].copy(), logits[:n].copy(), prob[:n].copy()

    def retrieve_batch(self, tree: Tree, seqs, beam, topk, use_mask=True, cons_off=None, cons=None,
                       widen_beam=False, n_threads=1):
        seqs = _ci32(seqs).reshape(-1, self.T)
        B = len(seqs)
        items = np.empty((B, topk), np.int32)
        logits = np.empty((B, topk), np.float32)
        counts = np.empty(B, np.int32)
        co = None if cons_off is None else np.ascontiguousarray(cons_off, np.int64)
        cc = None if cons is None else _ci32(cons)
        rc = lib().orc_tdm_retrieve_batch(tree.h, self.h, B, seqs, beam, topk, int(use_mask), _ptr(co), _ptr(cc),
                                          int(widen_beam), n_threads, items, logits, counts)
        if rc:
            raise IndexError(f"oracle error {rc}")
        return items, logits, counts

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_tdm_model_destroy(self.h)
            self.h = None


class TunedTdm:
    """oracle_tuned.c: the re-associated, batched CPU form of TDM retrieval over a DIN model (SURVEY 8(d)'s second CPU form).
    A BASELINE for bench.py, never a checker: logits are close to the reference's, not bit-equal."""

    def __init__(self, tree: Tree, model: TdmModel, n_threads=1):
        self.tree, self.model = tree, model              # keep the borrowed arrays alive
        self.T = model.T
        self.h = lib().orc_tuned_create(tree.h, model.h, int(n_threads))
        if not self.h:
            raise ValueError("tuned form: DIN(Float) models only")

    def retrieve_batch(self, seqs, beam, topk, use_mask=True, n_threads=1):
        seqs = _ci32(seqs).reshape(-1, self.T)
        B = len(seqs)
        items = np.empty((B, topk), np.int32)
        logits = np.empty((B, topk), np.float32)
        counts = np.empty(B, np.int32)
        rc = lib().orc_tuned_retrieve_batch(self.h, B, seqs, beam, topk, int(use_mask), n_threads, items, logits, counts)
        if rc:
            raise IndexError(f"oracle error {rc}")
        return items, logits, counts

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_tuned_destroy(self.h)
            self.h = None


class OtmModel:
    """DIN(Double) from the compact parameter vector."""

    def __init__(self, params, rows, E, T, deepfm=False):
        self.params = np.ascontiguousarray(params, np.float64)
        self.rows, self.E, self.T = int(rows), int(E), int(T)
        if deepfm:                                   # otm/.../model/DeepFM.scala: [emb | W1 (T+1)x(T+1)E | b1 | W2 | b2]
            assert self.params.size == rows * E + (T + 1) * (T + 1) * E + 2 * (T + 1) + 1, "bad DeepFM parameter count"
            self.h = lib().orc_otm_deepfm_create(self.rows, self.E, self.T, self.params)
        else:
            assert self.params.size == rows * E + E * E + 2 * E * E + 2 * E + 1, "bad DIN parameter count"
            self.h = lib().orc_otm_model_create(self.rows, self.E, self.T, self.params)

    def forward(self, node, seq, mask_flat=None):
        node = _ci32(node).ravel()
        seq = _ci32(seq).reshape(len(node), self.T)
        out = np.empty(len(node), np.float64)
        m = None if mask_flat is None else _ci32(mask_flat)
        rc = lib().orc_din_forward_f64_api(self.h, len(node), node, seq, _ptr(m), 0 if m is None else len(m), out)
        if rc:
            raise IndexError("embeddingLookup failed: index out of range")
        return out

    def pseudo_targets(self, seqs, target_off, targets, leaf_level, start_level, use_mask=True, M=None):
        """OTMTree.optimalPseudoTargets -> (ids [n_lvl, B, M], vals, counts [n_lvl, B]) for the levels start_level + 1 .. leaf_level"""
        seqs = _ci32(seqs).reshape(-1, self.T)
        B = len(seqs)
        off = np.ascontiguousarray(target_off, np.int64)
        tg = _ci32(targets)
        M = int(M or max(1, int(np.diff(off).max())))
        n_lvl = leaf_level - start_level
        ids = np.empty((n_lvl, B, M), np.int32)
        vals = np.empty((n_lvl, B, M), np.float64)
        cnt = np.zeros((n_lvl, B), np.int32)
        rc = lib().orc_otm_pseudo_targets(self.h, B, self.T, seqs, off, tg, leaf_level, start_level, int(use_mask), M, ids, vals, cnt)
        if rc:
            raise IndexError(f"oracle error {rc}")
        return ids, vals, cnt

    def beam_search(self, seq_leaf_ids, leaf_level, beam, use_mask=True):
        seq = _ci32(seq_leaf_ids)
        s = int(np.floor(np.log(beam) / np.log(2)))
        cap = 2 * max(beam, 1 << s) + 2
        ids = np.empty(cap, np.int32)
        sc = np.empty(cap, np.float64)
        n = lib().orc_otm_beam_search(self.h, seq, leaf_level, beam, int(use_mask), ids, sc)
        if n < 0:
            raise IndexError(f"oracle error {n}")
        return ids[:n].copy(), sc[:n].copy()

    def recommend(self, seq_leaf_ids, leaf_level, topk, beam, leaf_item, use_mask=True):
        seq = _ci32(seq_leaf_ids)
        items = np.empty(topk, np.int32)
        sc = np.empty(topk, np.float64)
        pr = np.empty(topk, np.float64)
        n = lib().orc_otm_recommend(self.h, seq, leaf_level, beam, topk, int(use_mask), _ci32(leaf_item), items, sc, pr)
        if n < 0:
            raise IndexError(f"oracle error {n}")
        return items[:n].copy(), sc[:n].copy(), pr[:n].copy()

    def retrieve_batch(self, seqs, leaf_level, beam, topk, leaf_item, use_mask=True, n_threads=1):
        seqs = _ci32(seqs).reshape(-1, self.T)
        B = len(seqs)
        items = np.empty((B, topk), np.int32)
        sc = np.empty((B, topk), np.float64)
        counts = np.empty(B, np.int32)
        rc = lib().orc_otm_retrieve_batch(self.h, B, seqs, leaf_level, beam, topk, int(use_mask), _ci32(leaf_item),
                                          n_threads, items, sc, counts)
        if rc:
            raise IndexError(f"oracle error {rc}")
        return items, sc, counts

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_otm_model_destroy(self.h)
            self.h = None


class DrModel:
    def __init__(self, num_item, K, D, T, E, layer_emb, layer_w, layer_b, rr_emb, rr_w, rr_b, sm_w, sm_b):
        c = lambda a: np.ascontiguousarray(a, np.float64)
        self.num_item, self.K, self.D, self.T, self.E = num_item, K, D, T, E
        self.layer_emb = c(layer_emb)
        self.layer_w = [c(w) for w in layer_w]
        self.layer_b = [c(b) for b in layer_b]
        self.rr_emb, self.rr_w, self.rr_b, self.sm_w, self.sm_b = c(rr_emb), c(rr_w), c(rr_b), c(sm_w), c(sm_b)
        wp = (C.c_void_p * D)(*[w.ctypes.data for w in self.layer_w])
        bp = (C.c_void_p * D)(*[b.ctypes.data for b in self.layer_b])
        self.h = lib().orc_dr_model_create(num_item, K, D, T, E, self.layer_emb, wp, bp, self.rr_emb, self.rr_w,
                                           self.rr_b, self.sm_w, self.sm_b)

    def beam_search(self, seq, beam):
        seq = _ci32(seq)
        paths = np.empty((max(beam, 1), self.D), np.int32)
        prob = np.empty(max(beam, 1), np.float64)
        n = lib().orc_dr_beam_search(self.h, seq, beam, paths, prob)
        if n < 0:
            raise IndexError(f"oracle error {n}")
        return paths[:n].copy(), prob[:n].copy()

    def rerank(self, seq, cand):
        cand = _ci32(cand)
        out = np.empty(len(cand), np.float64)
        rc = lib().orc_dr_rerank(self.h, _ci32(seq), len(cand), cand, out)
        if rc:
            raise IndexError(f"oracle error {rc}")
        return out

    def recommend(self, seq, topk, beam, path_off, path_items):
        ids = np.empty(topk, np.int32)
        sc = np.empty(topk, np.float64)
        pr = np.empty(topk, np.float64)
        n = lib().orc_dr_recommend(self.h, _ci32(seq), beam, topk, np.ascontiguousarray(path_off, np.int64),
                                   _ci32(path_items), ids, sc, pr)
        if n < 0:
            raise IndexError(f"oracle error {n}")
        return ids[:n].copy(), sc[:n].copy(), pr[:n].copy()

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_dr_model_destroy(self.h)
            self.h = None


def cross_entropy(logits, target, want_grad=True):
    """CrossEntropyCriterion (sizeAverage) on [R, C] Double logits -> (loss, gradInput)"""
    lg = np.ascontiguousarray(logits, np.float64)
    grad = np.empty_like(lg) if want_grad else None
    loss = lib().orc_cross_entropy_f64(lg.shape[0], lg.shape[1], lg, _ci32(target), grad.ctypes.data if want_grad else None)
    return loss, grad


def sampled_softmax(u, sm_w, sm_b, sampled, g_sm_w, g_sm_b):
    """SampledSoftmaxLoss forward + backward on user vectors u[n, E] -> (loss, gradInput); g_sm_w / g_sm_b are accumulated into"""
    u = np.ascontiguousarray(u, np.float64)
    sampled = _ci32(sampled).reshape(len(u), -1)
    gu = np.empty_like(u)
    loss = lib().orc_sampled_softmax_f64(len(u), u.shape[1], sampled.shape[1] - 1, u, sm_w, sm_b, sampled, gu, g_sm_w, g_sm_b)
    return loss, gu


def adam_eps(w, g, s, r, lr, eps, t):
    lib().orc_adam_eps_f64(w, g, s, r, w.size, lr, eps, t)


class DrTrainer:
    """One LocalOptimizer (deep-retrieval/.../optim/LocalOptimizer.scala) over a DrModel's arrays: layer Adam, rerank Adam and the
    SampledSoftmaxLoss's own Adam state, updated in place by step()."""

    def __init__(self, m: "DrModel", lr: float):
        self.m, self.lr = m, lr
        z = np.zeros_like
        self.layer = [m.layer_emb] + [x for d in range(m.D) for x in (m.layer_w[d], m.layer_b[d])]
        self.layer_s, self.layer_r = [z(x) for x in self.layer], [z(x) for x in self.layer]
        self.rr = [m.rr_emb, m.rr_w, m.rr_b]
        self.rr_s, self.rr_r = [z(x) for x in self.rr], [z(x) for x in self.rr]
        self.sm = [m.sm_w, m.sm_b]
        self.sm_g, self.sm_s, self.sm_r = [z(x) for x in self.sm], [z(x) for x in self.sm], [z(x) for x in self.sm]

    def layer_grad(self, seq, target, item_paths, P, parallelism=1):
        m = self.m
        g = [np.zeros_like(x) for x in self.layer]
        gw = (C.c_void_p * m.D)(*[g[1 + 2 * d].ctypes.data for d in range(m.D)])
        gb = (C.c_void_p * m.D)(*[g[2 + 2 * d].ctypes.data for d in range(m.D)])
        wp = (C.c_void_p * m.D)(*[w.ctypes.data for w in m.layer_w])
        bp = (C.c_void_p * m.D)(*[b.ctypes.data for b in m.layer_b])
        loss = np.zeros(m.D)
        seq = _ci32(seq).reshape(-1, m.T)
        rc = lib().orc_dr_layer_grad(m.num_item, m.K, m.D, m.T, m.E, m.layer_emb, wp, bp, len(seq), seq, _ci32(target), _ci32(item_paths), P,
                                     parallelism, g[0], gw, gb, loss)
        if rc:
            raise IndexError(f"oracle error {rc}")
        return g, loss

    def rerank_grad(self, seq, sampled):
        m = self.m
        seq = _ci32(seq).reshape(-1, m.T)
        sampled = _ci32(sampled).reshape(len(seq), -1)
        g = [np.zeros_like(x) for x in self.rr]
        loss = np.zeros(1)
        rc = lib().orc_dr_rerank_grad(m.num_item, m.T, m.E, m.rr_emb, m.rr_w, m.rr_b, m.sm_w, m.sm_b, len(seq), seq, sampled,
                                      sampled.shape[1] - 1, g[0], g[1], g[2], self.sm_g[0], self.sm_g[1], loss)
        if rc:
            raise IndexError(f"oracle error {rc}")
        return g, float(loss[0])

    def step(self, seq, target, item_paths, P, sampled, t, rerank_t=None, parallelism=1):
        """the body of LocalOptimizer.optimize's while loop (:62-84) -> (layer losses [D], rerank loss)"""
        g, loss = self.layer_grad(seq, target, item_paths, P, parallelism)
        for x, gx, s, r in zip(self.layer, g, self.layer_s, self.layer_r):
            lib().orc_adam_eps_f64(x.reshape(-1), gx.reshape(-1), s.reshape(-1), r.reshape(-1), x.size, self.lr, 1e-8, t)
        rloss = float("nan")
        if rerank_t:
            gr, rloss = self.rerank_grad(seq, sampled)
            for x, gx, s, r in zip(self.sm, self.sm_g, self.sm_s, self.sm_r):       # inside reRankCriterion.backward
                lib().orc_adam_eps_f64(x.reshape(-1), gx.reshape(-1), s.reshape(-1), r.reshape(-1), x.size, self.lr, 1e-7, rerank_t)
            for x, gx, s, r in zip(self.rr, gr, self.rr_s, self.rr_r):
                lib().orc_adam_eps_f64(x.reshape(-1), gx.reshape(-1), s.reshape(-1), r.reshape(-1), x.size, self.lr, 1e-8, rerank_t)
        return loss, rloss


def deepfm_gradients(params, rows, E, T, node, seq, labels):
    """one DeepFM training step's gradient of the compact vector [emb | W1 | b1 | W2 | b2] and the mean BCE loss (Float)"""
    params = np.ascontiguousarray(params, np.float32)
    node = _ci32(node).ravel()
    seq = _ci32(seq).reshape(len(node), T)
    grad = np.empty_like(params)
    loss = np.zeros(1, np.float32)
    rc = lib().orc_deepfm_gradients_f32(rows, E, T, params, len(node), node, seq, np.ascontiguousarray(labels, np.float32), grad, loss)
    if rc:
        raise IndexError(f"oracle error {rc}")
    return grad, loss[0]


def arg_partition(dist, position):
    """Utils.argPartition on a copy -> (partitioned values, indices)"""
    d = np.ascontiguousarray(dist, np.float64).copy()
    ix = np.arange(len(d), dtype=np.int32)
    rc = lib().orc_arg_partition(d, len(d), int(position), ix)
    if rc:
        raise ValueError(f"oracle error {rc}")
    return d, ix


def kmeans_tree(emb, iters, seed):
    """RecursiveCluster.run (kmeans) -> codes[n]"""
    emb = np.ascontiguousarray(emb, np.float64)
    codes = np.full(len(emb), -1, np.int32)
    rc = lib().orc_kmeans_tree(emb.shape[0], emb.shape[1], emb, int(iters), int(seed), codes)
    if rc:
        raise ValueError(f"oracle error {rc}")
    return codes


def softmax_f32(x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    lib().orc_softmax_f32(x.shape[0], x.shape[1], x, out)
    return out


def softmax_grad_f32(y, go):
    y = np.ascontiguousarray(y, np.float32)
    go = np.ascontiguousarray(go, np.float32)
    out = np.empty_like(y)
    lib().orc_softmax_grad_f32(y.shape[0], y.shape[1], y, go, out)
    return out


def din_gradients(params, rows, E, T, node, seq, mask_flat, labels):
    """zeroGrad + forward + BCEWithLogits(mean) + backward -> (grad of the compact vector, loss)."""
    params = np.ascontiguousarray(params)
    dt = params.dtype
    assert dt in (np.float32, np.float64)
    node = _ci32(node).ravel()
    seq = _ci32(seq).reshape(len(node), T)
    labels = np.ascontiguousarray(labels, dt).ravel()
    m = None if mask_flat is None else _ci32(mask_flat)
    grad = np.empty_like(params)
    loss = np.zeros(1, dt)
    fn = lib().orc_din_gradients_f32 if dt == np.float32 else lib().orc_din_gradients_f64
    rc = fn(rows, E, T, params, len(node), node, seq, _ptr(m), 0 if m is None else len(m), labels, grad, loss)
    if rc:
        raise IndexError(f"oracle error {rc}")
    return grad, loss[0]


def adam_step(w, g, s, r, lr, t):
    """in-place Adam.optimize on (w, s, r)."""
    fn = lib().orc_adam_f32 if w.dtype == np.float32 else lib().orc_adam_f64
    fn(w, np.ascontiguousarray(g, w.dtype), s, r, w.size, float(lr), int(t))
